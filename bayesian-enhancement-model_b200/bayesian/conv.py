"""Conv2dReparameterization — drop-in for basicsr/bayesian/conv.py:10-128.

Same constructor, parameter / buffer names (`mu_weight, rho_weight[, mu_bias, rho_bias]` are the only state_dict keys;
`eps_*`, `prior_mu_*`, `prior_rho_*` are non-persistent buffers, conv.py:57-69), same `.deterministic`, `.step`,
`.kl_loss()`. The forward runs on the sm_100a kernels:
  * 1x1, stride 1, no padding, groups 1   -> bem_bayes_pointwise (eps drawn by torch: sample fused into the weight load)
  * depthwise 3x3, stride 1, padding 1     -> bem_bayes_depthwise
  * any other geometry (not used by the BEM archs) -> bem_bayes_sample, then the library convolution on the sample
"""
from __future__ import annotations

import collections
import math
from itertools import repeat

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Parameter

from . import functional as BF
from .base_layer import BaseLayer_


def get_kernel_size(x, n):
    if isinstance(x, collections.abc.Iterable):
        return tuple(x)
    return tuple(repeat(x, n))


class Conv2dReparameterization(BaseLayer_):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 sigma_init=0.05, decay=0.9998):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('invalid in_channels size')
        if out_channels % groups != 0:
            raise ValueError('invalid in_channels size')
        self.deterministic = False   # set to True to get deterministic output
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.bias = bias
        self.decay = decay
        self.sigma_init = sigma_init
        self.step = 0

        kernel_size = get_kernel_size(kernel_size, 2)
        shape = (out_channels, in_channels // groups, kernel_size[0], kernel_size[1])
        self.mu_weight = Parameter(torch.Tensor(*shape))
        self.rho_weight = Parameter(torch.Tensor(*shape))
        self.register_buffer('eps_weight', torch.Tensor(*shape), persistent=False)
        self.register_buffer('prior_mu_weight', torch.Tensor(*shape), persistent=False)
        self.register_buffer('prior_rho_weight', torch.Tensor(*shape), persistent=False)
        if self.bias:
            self.mu_bias = Parameter(torch.Tensor(out_channels))
            self.rho_bias = Parameter(torch.Tensor(out_channels))
            self.register_buffer('eps_bias', torch.Tensor(out_channels), persistent=False)
            self.register_buffer('prior_mu_bias', torch.Tensor(out_channels), persistent=False)
            self.register_buffer('prior_rho_bias', torch.Tensor(out_channels), persistent=False)
        self.init_parameters()

    def init_parameters(self):
        rho_init = math.log(math.expm1(abs(self.sigma_init)) + 1e-20)
        nn.init.kaiming_normal_(self.mu_weight, mode='fan_in', nonlinearity='leaky_relu')
        self.rho_weight.data.fill_(rho_init)
        self.prior_mu_weight.data.copy_(self.mu_weight.data)
        self.prior_rho_weight.data.copy_(self.rho_weight.data)
        if self.bias:
            self.mu_bias.data.fill_(0)
            self.rho_bias.data.fill_(rho_init)
            self.prior_mu_bias.data.copy_(self.mu_bias.data)
            self.prior_rho_bias.data.copy_(self.rho_bias.data)

    # ------------------------------------------------------------------------------------------------
    def _geometry(self):
        k = get_kernel_size(self.kernel_size, 2)
        st, pd, dl = (get_kernel_size(v, 2) for v in (self.stride, self.padding, self.dilation))
        if k == (1, 1) and st == (1, 1) and pd == (0, 0) and self.groups == 1:
            return "pointwise"
        if (k == (3, 3) and st == (1, 1) and pd == (1, 1) and dl == (1, 1) and self.groups == self.in_channels
                and self.in_channels == self.out_channels):
            return "depthwise3"
        return "general"

    def _conv(self, input, w, b, S):
        """w: (S, Cout, Cin/g, kh, kw), b: (S, Cout) | None"""
        geo = self._geometry()
        if geo == "pointwise":
            return BF.pointwise_conv(input, w.reshape(S, self.out_channels, self.in_channels), b, S, ln=self._ln,
                                     residual=self._res)
        if geo == "depthwise3":
            return BF.depthwise_conv3x3(input, w.reshape(S, self.out_channels, 3, 3), b, S, act=self._act)
        # geometry outside the BEM hot path: library convolution on the sampled weights (S samples as S x groups groups)
        if S == 1:
            return F.conv2d(input, w[0], None if b is None else b[0], self.stride, self.padding, self.dilation, self.groups)
        Bx = input.shape[0] // S
        xi = input.reshape(S, Bx, *input.shape[1:]).transpose(0, 1).reshape(Bx, S * input.shape[1], *input.shape[2:])
        out = F.conv2d(xi, w.reshape(S * self.out_channels, *w.shape[2:]), None if b is None else b.reshape(-1),
                       self.stride, self.padding, self.dilation, S * self.groups)
        return out.reshape(Bx, S, self.out_channels, *out.shape[2:]).transpose(0, 1).reshape(S * Bx, self.out_channels, *out.shape[2:])

    def _forward_uncertain(self, input, eps_weight=None, eps_bias=None):
        if self.training:
            self._update_prior()
        S = self.mc_samples
        needs_grad = torch.is_grad_enabled() and (self.mu_weight.requires_grad or input.requires_grad)
        if (not needs_grad) and self._geometry() == "pointwise" and (self.eps_source == "torch" or eps_weight is not None):
            # inference fast path: the sampled weight never exists in memory
            eps_w = self._draw_eps("weight", eps_weight)
            b = self._sample("bias", eps_bias)[0] if self.bias else None
            oc, ic = self.out_channels, self.in_channels
            return BF.pointwise_conv_sampled(input, self.mu_weight.view(oc, ic), self._sigma_cached().view(oc, ic),
                                             eps_w.reshape(S, oc, ic), b, S, ln=self._ln, residual=self._res)
        w, _ = self._sample("weight", eps_weight)
        b = self._sample("bias", eps_bias)[0] if self.bias else None
        return self._conv(input, w, b, S)

    def _fuses_norm(self):
        return self._geometry() == "pointwise"

    def _fuses_residual(self):
        return self._geometry() == "pointwise"

    def _fuses_act(self):
        return self._geometry() == "depthwise3"

    def _forward_det(self, input):
        w = self.mu_weight.unsqueeze(0)
        b = self.mu_bias.unsqueeze(0) if self.bias else None
        return self._conv(input, w, b, 1)
