"""BaseLayer_ — drop-in for basicsr/bayesian/base_layer.py:8-39, plus the sampling machinery shared by the three
reparameterised layers (the reference repeats it in conv.py:91-111, linear.py:67-88, 165-187)."""
from __future__ import annotations

from abc import abstractmethod

import torch
import torch.nn as nn

from . import functional as BF
from .. import _lib


class BaseLayer_(_lib.InvalidatesCaches, nn.Module):
    # --- Monte-Carlo controls (extensions; defaults reproduce the reference exactly) ------------------------------
    eps_source = "torch"   # "torch": eps_*.normal_() from torch's global generator, weight then bias (conv.py:107,110)
    #                        "philox": counter-based stream generated inside the sample kernel, keyed (mc_seed, layer_id, sample)
    mc_samples = 1         # S > 1: the batch holds S groups of images, each group gets its own weight sample
    mc_seed = 0
    mc_sample0 = 0         # global index of this call's first sample (rank offset when samples are sharded)
    layer_id = 0

    def __init__(self):
        super().__init__()

    def forward(self, input, eps_weight=None, eps_bias=None, pre_norm=None, post_act=None, residual=None):
        """eps_weight / eps_bias (extension): inject the noise instead of drawing it — used by the parity tests to replay
        the eps the reference layer left in its eps_* buffers.
        pre_norm (extension): a LayerNorm2d-like module (weight, bias, eps over the channel dim) that precedes this layer;
        post_act (extension): "silu" | "gelu_gate", the activation that follows it; residual (extension): tensor added
        to the result (the block's skip connection). Each is fused into the layer's kernel when the geometry allows and
        no gradient is needed (1x1: norm + residual, depthwise 3x3: activation); otherwise it is simply applied here."""
        inj = getattr(self, "_injected_eps", None)   # tests: {"weight": t, "bias": t} replayed from the reference
        if inj:
            eps_weight = inj.get("weight", eps_weight)
            eps_bias = inj.get("bias", eps_bias)
        needs_grad = torch.is_grad_enabled() and (input.requires_grad or self.mu_weight.requires_grad)
        self._ln = self._act = self._res = None
        if pre_norm is not None:
            if self._fuses_norm() and not needs_grad:
                self._ln = (pre_norm.weight, pre_norm.bias, pre_norm.eps)
            else:
                input = pre_norm(input)
        if not needs_grad:
            if post_act is not None and self._fuses_act():
                self._act = post_act
            if residual is not None and self._fuses_residual():
                self._res = residual
        fused_act, fused_res = self._act is not None, self._res is not None
        try:
            if not self.deterministic:
                out = self._forward_uncertain(input, eps_weight, eps_bias)
            else:
                out = self._forward_det(input)
        finally:
            self._ln = self._act = self._res = None
        if post_act is not None and not fused_act:
            out = BF.apply_act(out, post_act)
        if residual is not None and not fused_res:
            out = residual + out
        return out

    _ln = _act = _res = None

    def _fuses_act(self):
        return False

    def _fuses_residual(self):
        return False

    def _fuses_norm(self):
        return False

    @abstractmethod
    def _forward_uncertain(self, input, eps_weight=None, eps_bias=None):
        pass

    @abstractmethod
    def _forward_det(self, input):
        pass

    def kl_div(self, mu_q, sigma_q, mu_p, sigma_p):
        """KL(Q || P) between diagonal Gaussians, averaged over elements (base_layer.py:26-39)."""
        kl = torch.log(sigma_p) - torch.log(sigma_q) + (sigma_q ** 2 + (mu_q - mu_p) ** 2) / (2 * (sigma_p ** 2)) - 0.5
        return kl.mean()

    # ----------------------------------------------------------------------------------------------------------------
    # shared pieces
    # ----------------------------------------------------------------------------------------------------------------
    @property
    def sigma_weight(self):
        """log1p(exp(rho_weight)); the reference caches this attribute on every stochastic forward (conv.py:106) and
        reads it in kl_loss (:86) — computed on demand here so inference does not launch it per layer per sample."""
        return torch.log1p(torch.exp(self.rho_weight))

    @property
    def sigma_bias(self):
        return torch.log1p(torch.exp(self.rho_bias))

    def _sigma_cached(self, which="weight"):
        """detached log1p(exp(rho)) for the inference kernels, recomputed only when rho changes: optimizer steps bump the
        tensor version; train() / load_state_dict / .to() and explicit bem_b200.invalidate_caches() (needed after writes
        through `.data`, which leave the version alone) advance the package's cache generation"""
        rho = getattr(self, "rho_" + which)
        key = (rho._version, rho.data_ptr(), rho.device, _lib.cache_generation())
        cache = self.__dict__.setdefault("_sigma_cache", {})
        hit = cache.get(which)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, torch.log1p(torch.exp(rho.detach())).contiguous())
            cache[which] = hit
        return hit[1]

    def kl_loss(self):
        kl = self.kl_div(self.mu_weight, self.sigma_weight, self.prior_mu_weight, self.prior_sigma_weight)
        if self.bias:
            kl += self.kl_div(self.mu_bias, self.sigma_bias, self.prior_mu_bias, self.prior_sigma_bias)
        return kl

    def _update_prior(self):
        """training-mode prior EMA (conv.py:92-104): decay' = min(decay, (1 + step) / (10 + step))."""
        with torch.no_grad():
            _decay = min(self.decay, (1 + self.step) / (10 + self.step))
            self.prior_mu_weight = _decay * self.prior_mu_weight + (1 - _decay) * self.mu_weight
            self.prior_rho_weight = _decay * self.prior_rho_weight + (1 - _decay) * self.rho_weight
            self.prior_sigma_weight = torch.log1p(torch.exp(self.prior_rho_weight))
            if self.bias:
                self.prior_mu_bias = _decay * self.prior_mu_bias + (1 - _decay) * self.mu_bias
                self.prior_rho_bias = _decay * self.prior_rho_bias + (1 - _decay) * self.rho_bias
                self.prior_sigma_bias = torch.log1p(torch.exp(self.prior_rho_bias))
        self.step += 1

    def _draw_eps(self, which, given):
        """eps for `which` in ("weight", "bias"): (S, *shape) tensor, or None to let the sample kernel generate it."""
        buf = getattr(self, "eps_" + which)
        S = self.mc_samples
        if given is not None:
            return given.reshape((S,) + tuple(buf.shape)).to(buf.dtype)
        if self.eps_source == "philox":
            return None
        if S == 1:
            return buf.data.normal_().unsqueeze(0)      # in place, same generator call as the reference
        return torch.randn((S,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device)

    def _sample(self, which, given):
        """-> (w, eps_used) with w: (S, *shape). stream ids: 2*layer_id for weights, 2*layer_id + 1 for biases."""
        mu, rho = getattr(self, "mu_" + which), getattr(self, "rho_" + which)
        arena = getattr(self, "_arena", None)
        if (arena is not None and given is None and self.eps_source == "philox" and self.mc_samples == arena[which].shape[0]
                and not (torch.is_grad_enabled() and mu.requires_grad)):
            return arena[which], None   # (S, *shape), drawn for the whole network by MCArena.draw (one launch per sample set)
        eps = self._draw_eps(which, given)
        sid = 2 * int(self.layer_id) + (1 if which == "bias" else 0)
        w, used = BF.sample_weights(mu, rho, eps, self.mc_samples, self.mc_seed, sid, self.mc_sample0)
        if eps is None and self.mc_samples == 1:
            getattr(self, "eps_" + which).data.copy_(used[0])   # leave the eps in the buffer like the reference does
        return w, used
