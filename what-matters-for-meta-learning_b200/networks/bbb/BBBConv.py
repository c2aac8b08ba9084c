"""Drop-in ``BBBConv2d`` (reference networks/bbb/BBBConv.py:37-105): same constructor, parameters (``W_mu``, ``W_rho``,
``bias_mu``, ``bias_rho``), initialisation draws and ``forward(input, sample=True)`` / ``kl_loss()`` contract."""
import os

import torch
from torch.nn import Parameter

from .misc import ModuleWrapper

_PRIORS = {'prior_mu': 0, 'prior_sigma': 0.1, 'posterior_mu_initial': (0, 0.1), 'posterior_rho_initial': (-3, 0.1)}


class BBBConv2d(ModuleWrapper):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True, priors=None,
                 device="cpu"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, 1
        self.use_bias, self.device = bias, device
        priors = dict(_PRIORS) if priors is None else priors
        self.prior_mu, self.prior_sigma = priors['prior_mu'], priors['prior_sigma']
        self.posterior_mu_initial, self.posterior_rho_initial = priors['posterior_mu_initial'], priors['posterior_rho_initial']
        self.W_mu = Parameter(torch.empty((out_channels, in_channels, *self.kernel_size), device=device))
        self.W_rho = Parameter(torch.empty((out_channels, in_channels, *self.kernel_size), device=device))
        if bias:
            self.bias_mu = Parameter(torch.empty((out_channels), device=device))
            self.bias_rho = Parameter(torch.empty((out_channels), device=device))
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
        return bbb.conv2d_forward(self, input, self.training or sample)

    def kl_loss(self):
        """The KL of the LAST forward's sample (the reference recomputes it from W_sigma of the last forward, :102-105);
        it is produced by the same kernel that samples the weights."""
        if getattr(self, "_kl", None) is None:
            raise RuntimeError("kl_loss() before a sampling forward")
        return self._kl.view(())


if os.environ.get("B200NP_BBB", "0") != "1":
    from .._refload import reference_module
    _ref = reference_module("bbb.BBBConv")
    if _ref is not None:
        BBBConv2d = _ref.BBBConv2d      # noqa: F811  (default: the reference's own class)
