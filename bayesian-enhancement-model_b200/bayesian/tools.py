"""Module-swap tools — drop-ins for basicsr/bayesian/tools.py:3-84, plus the Monte-Carlo configuration helper."""
from __future__ import annotations


def _layer_cls(name):
    from . import conv, linear
    table = {"Conv2dReparameterization": conv.Conv2dReparameterization,
             "Linear2dReparameterization": linear.Linear2dReparameterization,
             "LinearReparameterization": linear.LinearReparameterization}
    try:
        return table[name]
    except KeyError:
        # the reference does getattr(bayesian, name) (tools.py:5,25) and fails the same way for e.g. ConvTranspose2d
        raise AttributeError(f"module 'bayesian' has no attribute '{name}'")


_LINEAR_ARGS = ("in_features", "out_features")
_CONV_ARGS = ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups")


def _reparameterize(params, src, geometry):
    """Bayesian twin of the deterministic layer `src`: same geometry (the attributes named in `geometry`), the twin's class
    looked up by name as the reference does (`<ClassName>Reparameterization`, tools.py:5,25); with params["pretrain"] the
    posterior and prior means start from the deterministic weights (tools.py:11-18, 37-44)."""
    twin_cls = _layer_cls(type(src).__name__ + "Reparameterization")
    kwargs = {k: getattr(src, k) for k in geometry}
    twin = twin_cls(bias=src.bias is not None, decay=params["decay"], sigma_init=params["sigma_init"], **kwargs)
    if params["pretrain"]:
        pairs = [("weight", src.weight)] + ([("bias", src.bias)] if twin.bias else [])
        for which, value in pairs:
            for target in ("mu_", "prior_mu_"):
                getattr(twin, target + which).data.copy_(value.data)
    return twin


def bnn_linear_layer(params, d):
    return _reparameterize(params, d, _LINEAR_ARGS)


def bnn_conv_layer(params, d):
    return _reparameterize(params, d, _CONV_ARGS)


def convert2bnn(m, config):
    """Replace, in place and recursively, every leaf child whose class name contains "Linear" / "Conv" by its Bayesian twin
    (tools.py:52-63: containers are descended into, leaves are matched by class NAME, "Linear" tested first)."""
    for child_name, child in list(m.named_children()):
        if len(child._modules) > 0:
            convert2bnn(child, config)
            continue
        cls_name = type(child).__name__
        builder = bnn_linear_layer if "Linear" in cls_name else bnn_conv_layer if "Conv" in cls_name else None
        if builder is not None:
            setattr(m, child_name, builder(config, child))


def convert2bnn_selective(model, config):
    """convert2bnn on every sub-module flagged `.bayesian = True` (tools.py:47-50)."""
    flagged = [mod for mod in model.modules() if getattr(mod, "bayesian", False)]
    for mod in flagged:
        convert2bnn(mod, config)


def set_prediction_type(model, deterministic=True):
    """switch every layer that has a `.deterministic` flag (tools.py:65-73)"""
    for mod in model.modules():
        if hasattr(mod, "deterministic"):
            mod.deterministic = bool(deterministic)


def get_kl_loss(m):
    """sum of `.kl_loss()` over the layers that define it; None when there is none (tools.py:76-84)"""
    terms = [layer.kl_loss() for layer in m.modules() if hasattr(layer, "kl_loss")]
    if not terms:
        return None
    total = terms[0]
    for t in terms[1:]:
        total = total + t
    return total


# ---------------------------------------------------------------------------------------------------------------------
# extension: Monte-Carlo configuration of every Bayesian layer of a model
# ---------------------------------------------------------------------------------------------------------------------
def bayesian_layers(model):
    """Bayesian layers in execution-independent (registration) order; the index is the layer's Philox stream id."""
    from .base_layer import BaseLayer_
    return [m for m in model.modules() if isinstance(m, BaseLayer_)]


def set_mc_config(model, mc_samples=None, eps_source=None, seed=None, sample0=None):
    """mc_samples: weight samples batched per forward; eps_source: "torch" | "philox"; seed / sample0: Philox key and the
    global index of the first sample of the next forward (rank offset under sample sharding)."""
    for i, layer in enumerate(bayesian_layers(model)):
        layer.layer_id = i
        if mc_samples is not None:
            layer.mc_samples = int(mc_samples)
        if eps_source is not None:
            if eps_source not in ("torch", "philox"):
                raise ValueError(eps_source)
            layer.eps_source = eps_source
        if seed is not None:
            layer.mc_seed = int(seed)
        if sample0 is not None:
            layer.mc_sample0 = int(sample0)


# ---------------------------------------------------------------------------------------------------------------------
# extension: the resume gap of the reference. `prior_mu_* / prior_rho_*` are NON-persistent buffers and `step`,
# `prior_sigma_*` plain attributes (conv.py:39-52, linear.py:26-39), so they are absent from every checkpoint: a run resumed
# from `{'params': state_dict}` (basicsr/models/base_model.py:255-263) restarts the prior EMA from the freshly initialised
# prior and step 0, which changes the KL term of the first thousands of iterations. The state_dict itself must stay
# key-for-key the reference's (released checkpoints load with strict=True), so the missing state travels in a side dict.
# ---------------------------------------------------------------------------------------------------------------------
def prior_state_dict(model):
    """{'<layer name>.prior_mu_weight': tensor, ..., '<layer name>.step': int}: what the checkpoint does not hold"""
    out = {}
    for name, layer in model.named_modules():
        if not (hasattr(layer, "deterministic") and hasattr(layer, "prior_mu_weight")):
            continue
        for k in ("prior_mu_weight", "prior_rho_weight", "prior_mu_bias", "prior_rho_bias"):
            t = getattr(layer, k, None)
            if t is not None:
                out[f"{name}.{k}"] = t.detach().clone()
        out[f"{name}.step"] = int(getattr(layer, "step", 0))
    return out


def load_prior_state_dict(model, state):
    """inverse of prior_state_dict: restores the prior EMA tensors, `step`, and `prior_sigma_*` = log1p(exp(prior_rho_*))
    (what the next training forward would otherwise recompute from a reset prior). Returns the restored layer names."""
    import torch
    done = []
    for name, layer in model.named_modules():
        if f"{name}.step" not in state:
            continue
        for k in ("prior_mu_weight", "prior_rho_weight", "prior_mu_bias", "prior_rho_bias"):
            key = f"{name}.{k}"
            if key in state and getattr(layer, k, None) is not None:
                with torch.no_grad():
                    getattr(layer, k).copy_(state[key].to(getattr(layer, k).device))
        layer.step = int(state[f"{name}.step"])
        with torch.no_grad():
            layer.prior_sigma_weight = torch.log1p(torch.exp(layer.prior_rho_weight))
            if getattr(layer, "bias", False):
                layer.prior_sigma_bias = torch.log1p(torch.exp(layer.prior_rho_bias))
        done.append(name)
    return done
