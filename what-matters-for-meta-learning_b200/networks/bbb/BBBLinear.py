"""Drop-in ``BBBLinear`` (reference networks/bbb/BBBLinear.py:34-100): same constructor, parameters, initialisation draws
and ``forward(input, sample=True)`` / ``kl_loss()`` contract."""
import os

import torch
from torch.nn import Parameter

from .misc import ModuleWrapper

_PRIORS = {'prior_mu': 0, 'prior_sigma': 0.1, 'posterior_mu_initial': (0, 0.1), 'posterior_rho_initial': (-3, 0.1)}


class BBBLinear(ModuleWrapper):
    def __init__(self, in_features, out_features, bias=True, priors=None, device="cpu"):
        super().__init__()
        self.in_features, self.out_features, self.use_bias, self.device = in_features, out_features, bias, device
        priors = dict(_PRIORS) if priors is None else priors
        self.prior_mu, self.prior_sigma = priors['prior_mu'], priors['prior_sigma']
        self.posterior_mu_initial, self.posterior_rho_initial = priors['posterior_mu_initial'], priors['posterior_rho_initial']
        self.W_mu = Parameter(torch.empty((out_features, in_features), device=device))
        self.W_rho = Parameter(torch.empty((out_features, in_features), device=device))
        if bias:
            self.bias_mu = Parameter(torch.empty((out_features), device=device))
            self.bias_rho = Parameter(torch.empty((out_features), device=device))
        else:
            self.register_parameter('bias_mu', None)
            self.register_parameter('bias_rho', None)
        self.reset_parameters()

    def reset_parameters(self):
        self.W_mu.data.normal_(*self.posterior_mu_initial)
        self.W_rho.data.normal_(*self.posterior_rho_initial)
        if self.use_bias:
            self.bias_mu.data.normal_(*self.posterior_mu_initial)
            self.bias_rho.data.normal_(*self.posterior_rho_initial)

    def forward(self, input, sample=True):
        from b200np import bbb
        return bbb.linear_forward(self, input, self.training or sample)

    def kl_loss(self):
        if getattr(self, "_kl", None) is None:
            raise RuntimeError("kl_loss() before a sampling forward")
        return self._kl.view(())


if os.environ.get("B200NP_BBB", "0") != "1":
    from .._refload import reference_module
    _ref = reference_module("bbb.BBBLinear")
    if _ref is not None:
        BBBLinear = _ref.BBBLinear      # noqa: F811  (default: the reference's own class)
