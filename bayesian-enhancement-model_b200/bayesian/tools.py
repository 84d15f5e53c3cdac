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


def bnn_linear_layer(params, d):
    layer_fn = _layer_cls(d.__class__.__name__ + "Reparameterization")
    bnn_layer = layer_fn(in_features=d.in_features, out_features=d.out_features, bias=d.bias is not None,
                         decay=params["decay"], sigma_init=params["sigma_init"])
    if params["pretrain"]:
        bnn_layer.mu_weight.data.copy_(d.weight.data)
        bnn_layer.prior_mu_weight.data.copy_(d.weight.data)
        if bnn_layer.bias:
            bnn_layer.mu_bias.data.copy_(d.bias.data)
            bnn_layer.prior_mu_bias.data.copy_(d.bias.data)
    return bnn_layer


def bnn_conv_layer(params, d):
    layer_fn = _layer_cls(d.__class__.__name__ + "Reparameterization")
    bnn_layer = layer_fn(in_channels=d.in_channels, out_channels=d.out_channels, kernel_size=d.kernel_size,
                         stride=d.stride, padding=d.padding, dilation=d.dilation, groups=d.groups,
                         bias=d.bias is not None, decay=params["decay"], sigma_init=params["sigma_init"])
    if params["pretrain"]:
        bnn_layer.mu_weight.data.copy_(d.weight.data)
        bnn_layer.prior_mu_weight.data.copy_(d.weight.data)
        if bnn_layer.bias:
            bnn_layer.mu_bias.data.copy_(d.bias.data)
            bnn_layer.prior_mu_bias.data.copy_(d.bias.data)
    return bnn_layer


def convert2bnn_selective(model, config):
    for name, module in model.named_modules():
        if getattr(module, 'bayesian', False):
            convert2bnn(module, config)


def convert2bnn(m, config):
    for name, value in list(m._modules.items()):
        if m._modules[name]._modules:
            convert2bnn(m._modules[name], config)
        elif "Linear" in m._modules[name].__class__.__name__:
            setattr(m, name, bnn_linear_layer(config, m._modules[name]))
        elif "Conv" in m._modules[name].__class__.__name__:
            setattr(m, name, bnn_conv_layer(config, m._modules[name]))
        else:
            pass
    return


def set_prediction_type(model, deterministic=True):
    for name, module in model.named_modules():
        if hasattr(module, 'deterministic'):
            module.deterministic = bool(deterministic)


def get_kl_loss(m):
    kl_loss = None
    for layer in m.modules():
        if hasattr(layer, "kl_loss"):
            if kl_loss is None:
                kl_loss = layer.kl_loss()
            else:
                kl_loss += layer.kl_loss()
    return kl_loss


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
