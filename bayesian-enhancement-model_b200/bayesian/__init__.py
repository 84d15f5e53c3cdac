"""`bayesian` — drop-in for the reference's top-level package basicsr/bayesian (basicsr/bayesian/__init__.py:1-4)."""
from .base_layer import *          # noqa: F401,F403
from .base_layer import BaseLayer_  # noqa: F401
from .conv import *                # noqa: F401,F403
from .linear import *              # noqa: F401,F403
from .tools import *               # noqa: F401,F403
from . import functional           # noqa: F401
