"""Linear2dReparameterization / LinearReparameterization — drop-ins for basicsr/bayesian/linear.py:8-104, 106-203."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Parameter

from . import functional as BF
from .base_layer import BaseLayer_


class _LinearBase(BaseLayer_):
    def __init__(self, in_features, out_features, bias=True, sigma_init=0.05, decay=0.9998):
        super().__init__()
        self.deterministic = False   # set to True to get deterministic output
        self.in_features = in_features
        self.out_features = out_features
        self.bias = bias
        self.decay = decay
        self.sigma_init = sigma_init
        self.step = 0
        self.mu_weight = Parameter(torch.Tensor(out_features, in_features))
        self.rho_weight = Parameter(torch.Tensor(out_features, in_features))
        self.register_buffer('eps_weight', torch.Tensor(out_features, in_features), persistent=False)
        self.register_buffer('prior_mu_weight', torch.Tensor(out_features, in_features), persistent=False)
        self.register_buffer('prior_rho_weight', torch.Tensor(out_features, in_features), persistent=False)
        if bias:
            self.mu_bias = Parameter(torch.Tensor(out_features))
            self.rho_bias = Parameter(torch.Tensor(out_features))
            self.register_buffer('eps_bias', torch.Tensor(out_features), persistent=False)
            self.register_buffer('prior_mu_bias', torch.Tensor(out_features), persistent=False)
            self.register_buffer('prior_rho_bias', torch.Tensor(out_features), persistent=False)
        self.init_parameters()

    def init_parameters(self):
        rho_init = math.log(math.expm1(abs(self.sigma_init)) + 1e-20)
        nn.init.xavier_uniform_(self.mu_weight)
        self.rho_weight.data.fill_(rho_init)
        self.prior_mu_weight.data.copy_(self.mu_weight.data)
        self.prior_rho_weight.data.copy_(self.rho_weight.data)
        if self.bias:
            self.mu_bias.data.fill_(0)
            self.rho_bias.data.fill_(rho_init)
            self.prior_mu_bias.data.copy_(self.mu_bias.data)
            self.prior_rho_bias.data.copy_(self.rho_bias.data)


class Linear2dReparameterization(_LinearBase):
    """1x1 convolution over (B, C, H, W) with a sampled (out, in) weight (linear.py:8-104) -> bem_bayes_pointwise."""

    def _forward_uncertain(self, input, eps_weight=None, eps_bias=None):
        if self.training:
            self._update_prior()
        S = self.mc_samples
        needs_grad = torch.is_grad_enabled() and (self.mu_weight.requires_grad or input.requires_grad)
        if (not needs_grad) and (self.eps_source == "torch" or eps_weight is not None):
            eps_w = self._draw_eps("weight", eps_weight)
            b = self._sample("bias", eps_bias)[0] if self.bias else None
            return BF.pointwise_conv_sampled(input, self.mu_weight, self._sigma_cached(), eps_w, b, S, ln=self._ln,
                                             residual=self._res)
        w, _ = self._sample("weight", eps_weight)
        b = self._sample("bias", eps_bias)[0] if self.bias else None
        return BF.pointwise_conv(input, w, b, S, ln=self._ln, residual=self._res)

    def _fuses_norm(self):
        return True

    def _fuses_residual(self):
        return True

    def _forward_det(self, input):
        return BF.pointwise_conv(input, self.mu_weight.unsqueeze(0), self.mu_bias.unsqueeze(0) if self.bias else None, 1,
                                 ln=self._ln, residual=self._res)


class LinearReparameterization(_LinearBase):
    """F.linear on the last axis with a sampled weight (linear.py:106-203). No shipped arch instantiates it
    (channel_first=True everywhere). At inference it runs on the same tcgen05 1x1 kernel as Linear2d: the rows of
    the (.., in_features) input are the kernel's pixels, so the input is presented channel-major (one transposing copy each
    way — the layer's natural layout is the transpose of the kernel's); with a gradient the contraction is the library GEMM
    on the weights sampled by bem_bayes_sample."""

    def _contract(self, input, w, b, S):
        """w: (S, out, in), b: (S, out) | None"""
        needs_grad = torch.is_grad_enabled() and (input.requires_grad or w.requires_grad)
        if input.is_cuda and input.dtype == torch.float32 and not needs_grad and input.numel() > 0:
            xs = input.reshape(S, -1, self.in_features).transpose(1, 2).contiguous()        # (S, in, rows)
            out = BF.pointwise_conv(xs, w, b, S)                                             # (S, out, rows)
            return out.transpose(1, 2).reshape(*input.shape[:-1], self.out_features)
        if S == 1:
            return F.linear(input, w[0], None if b is None else b[0])
        xs = input.reshape(S, -1, self.in_features)
        out = torch.baddbmm(b.unsqueeze(1), xs, w.transpose(1, 2)) if b is not None else torch.bmm(xs, w.transpose(1, 2))
        return out.reshape(*input.shape[:-1], self.out_features)

    def _forward_uncertain(self, input, eps_weight=None, eps_bias=None):
        if self.training:
            self._update_prior()
        S = self.mc_samples
        w, _ = self._sample("weight", eps_weight)
        b = self._sample("bias", eps_bias)[0] if self.bias else None
        return self._contract(input, w, b, S)

    def _forward_det(self, input):
        return self._contract(input, self.mu_weight.unsqueeze(0), self.mu_bias.unsqueeze(0) if self.bias else None, 1)
