"""Bayes-by-backprop layers of the "MR" model variants on libb200np kernels.

The reference (networks/bbb/BBBConv.py:83-105, BBBLinear.py) samples ``w = mu + eps * log1p(exp(rho))`` with ``eps`` drawn
on the HOST (``torch.empty(size).normal_(0, 1).to(device)``, :86) -- so with the same CPU seed the noise is reproducible
and parity is exact, not statistical -- and adds ``calculate_kl(prior, posterior)`` (:32-34, :102-105) to the loss.
Here sampling, sigma and the KL reduction are ONE kernel (``b200np_bbb_sample_kl_fwd``), its backward folds dL/dw and
dL/dkl into dL/dmu and dL/drho, and the sampled weight feeds the same convolution / GEMM kernels as everywhere else.
"""
import torch
from torch.autograd import Function

from . import engine, ops
from .lib import ACT_NONE
from .mmaml import Conv3x3S2Fn


class BBBSampleFn(Function):
    """(mu, rho, eps) -> (w, kl): w = mu + eps * softplus(rho); kl [1] as networks/bbb/BBBConv.py:102-105 sums it."""

    @staticmethod
    def forward(ctx, mu, rho, eps, prior_mu, prior_sigma):
        mu, rho, eps = mu.contiguous(), rho.contiguous(), eps.contiguous()
        w, sigma, kl = ops.bbb_sample_kl_fwd(mu, rho, eps, prior_mu, prior_sigma)
        ctx.prior = (prior_mu, prior_sigma)
        ctx.save_for_backward(mu, rho, eps, sigma)
        ctx.mark_non_differentiable(sigma)
        return w, kl, sigma

    @staticmethod
    def backward(ctx, dw, dkl, _dsigma):
        mu, rho, eps, sigma = ctx.saved_tensors
        dw = None if dw is None else dw.contiguous()
        dkl = None if dkl is None else dkl.contiguous()
        dmu, drho = ops.bbb_sample_kl_bwd(dw, dkl, mu, rho, eps, sigma, *ctx.prior)
        return dmu, drho, None, None, None


def sample(layer, mu, rho):
    """The host-side draw of the reference (CPU generator, then .to(device)), then the fused sample + KL kernel."""
    eps = torch.empty(mu.size()).normal_(0, 1).to(mu.device)
    return BBBSampleFn.apply(mu, rho, eps, float(layer.prior_mu), float(layer.prior_sigma))


def conv2d_forward(layer, x, sample_weights):
    """BBBConv2d.forward (BBBConv.py:83-100) for the 3x3 / stride 2 / padding 1 convolutions of the MR encoders
    (networks/CNPMR.py:29-52); NCHW in, NCHW out like the reference layer."""
    if not (x.is_cuda and x.dtype == torch.float32):
        raise RuntimeError("BBBConv2d: expected a float32 CUDA tensor (the B200 path has no CPU fallback)")
    if layer.kernel_size != (3, 3) or layer.stride not in (2, (2, 2)) or layer.padding not in (1, (1, 1)) or layer.dilation not in (1, (1, 1)):
        raise NotImplementedError("B200 BBBConv2d covers 3x3 / stride 2 / padding 1 (the MR encoders' configuration)")
    kl = None
    if sample_weights:
        w, kl_w, sig_w = sample(layer, layer.W_mu, layer.W_rho)
        layer.W_sigma = sig_w
        kl = kl_w
        if layer.use_bias:
            b, kl_b, sig_b = sample(layer, layer.bias_mu, layer.bias_rho)
            layer.bias_sigma = sig_b
            kl = kl + kl_b
        else:
            b = None
    else:
        w, b = layer.W_mu, (layer.bias_mu if layer.use_bias else None)
    layer._kl = kl
    N, C, H, W = x.shape
    xh = x.reshape(N, H, W, 1) if C == 1 else NchwToNhwcFn.apply(x)
    if b is None:
        b = ops.zeros((layer.out_channels,), x)
    y = Conv3x3S2Fn.apply(engine.PRECISION, xh, w, b)
    return NhwcToNchwFn.apply(y)


def linear_forward(layer, x, sample_weights):
    """BBBLinear.forward: F.linear(x, W, b) with sampled parameters."""
    kl = None
    if sample_weights:
        w, kl, sig_w = sample(layer, layer.W_mu, layer.W_rho)
        layer.W_sigma = sig_w
        if layer.use_bias:
            b, kl_b, sig_b = sample(layer, layer.bias_mu, layer.bias_rho)
            layer.bias_sigma = sig_b
            kl = kl + kl_b
        else:
            b = None
    else:
        w, b = layer.W_mu, (layer.bias_mu if layer.use_bias else None)
    layer._kl = kl
    if b is None:
        b = ops.zeros((layer.out_features,), x)
    return engine.LinearFn.apply(ACT_NONE, engine.PRECISION, x, None, w, b)


class NhwcToNchwFn(Function):
    """NHWC -> NCHW (and its transpose back for the gradient): the layer keeps the reference's NCHW interface because
    nn.ReLU / nn.MaxPool2d of the surrounding nn.Sequential (networks/CNPMR.py:29-52) consume its output."""

    @staticmethod
    def forward(ctx, y):
        N, H, W, C = y.shape
        return ops.nhwc_to_nchw_flat(y.contiguous()).view(N, C, H, W)

    @staticmethod
    def backward(ctx, d):
        N, C, H, W = d.shape
        dx = torch.empty((N, H, W, C), device=d.device, dtype=d.dtype)
        return ops.nchw_flat_to_nhwc(d.contiguous().view(N, -1), None, dx, mask=False)


class NchwToNhwcFn(Function):
    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        out = torch.empty((N, H, W, C), device=x.device, dtype=x.dtype)
        return ops.nchw_flat_to_nhwc(x.contiguous().view(N, -1), None, out, mask=False)

    @staticmethod
    def backward(ctx, d):
        N, H, W, C = d.shape
        return ops.nhwc_to_nchw_flat(d.contiguous()).view(N, C, H, W)
