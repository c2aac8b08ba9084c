"""Forward, backward and (through b200np/second_order.py) second-order backward of the MMAML conv nets on libb200np
kernels.

``GatedConvModel`` (networks/gated_conv_net.py:167-212): four blocks of 3x3 stride-2 conv -> batch-statistics
BatchNorm (``training=True`` always, no affine) -> FiLM ``x * (1 + gamma) + beta`` (:154-159) -> ReLU with 32/64/128/256
channels, mean over the 8x8 map, Linear, tanh.  ``ConvEmbeddingModel`` (networks/conv_embedding_model.py:99-184): the
same conv stack with affine BatchNorm, mean over the map, Linear 256->128 + ReLU, average (or max) over the task's
samples, one Linear head per FiLM layer.

Every op is an ``autograd.Function`` over C-ABI kernels: the convolution is im2col + the tcgen05 GEMM (forward, weight
gradient and data gradient read torch's ``[Cout, Cin, 3, 3]`` weight as it is, any channel count), normalisation +
scale/shift + ReLU is one fused kernel after a deterministic two-level statistics pass.

Second order: the reference's MMAML learner differentiates through the inner-loop gradient
(trainer/meta_learner_reg.py:116-122, ``first_order=False`` at train.py:99).  Under ``create_graph=True`` autograd runs
``backward`` with gradients enabled; the Functions below then return their gradients through the differentiable
backward ops of ``b200np/second_order.py`` (same kernels, closed under differentiation) instead of the fused
single-pass kernels, so the outer gradient contains the second-order terms.
"""
import torch
from torch.autograd import Function
from . import engine, ops, second_order
from .engine import _p
from .lib import ACT_NONE, ACT_RELU, ACT_TANH


class Conv3x3S2Fn(Function):
    """x NHWC [N,H,W,Cin], w [Cout,Cin,3,3], b [Cout] -> NHWC [N,H/2,W/2,Cout] (stride 2, padding 1)."""

    @staticmethod
    def forward(ctx, prec, x_in, w_in, b):
        x, w = x_in.contiguous(), w_in.contiguous()
        N, H, W, Cin = x.shape
        Cout = w.shape[0]
        K = Cin * 9
        col = ops.im2col3x3s2(x)
        M = col.shape[0]
        y = ops.empty((N, H // 2, W // 2, Cout), x)
        ops.gemm(_p(col), _p(w), _p(y), M, Cout, K, K, 1, 1, K, Cout, bias=_p(b), prec=prec)
        ctx.prec, ctx.xshape = prec, tuple(x.shape)
        ctx.save_for_backward(col, w, x_in, w_in)      # the inputs themselves: they carry the graph in second-order mode
        return y

    @staticmethod
    def backward(ctx, dy):
        col, w, x_in, w_in = ctx.saved_tensors
        prec = ctx.prec
        if torch.is_grad_enabled():                     # create_graph=True
            dx, dw, db = second_order.conv3x3s2_backward(prec, x_in, w_in, dy, ctx.needs_input_grad[1], col=col)
            return None, dx, dw, db
        N, H, W, Cin = ctx.xshape
        Cout, K = w.shape[0], Cin * 9
        M = col.shape[0]
        dy = dy.contiguous()
        dw = ops.empty(tuple(w.shape), w)
        ops.gemm(_p(dy), _p(col), _p(dw), Cout, K, M, 1, Cout, K, 1, K, prec=prec)          # dW = dY^T col
        db = ops.colsum(dy, M, Cout, Cout)
        dx = None
        if ctx.needs_input_grad[1]:
            dcol = ops.empty((M, K), dy)
            ops.gemm(_p(dy), _p(w), _p(dcol), M, K, Cout, Cout, 1, K, 1, K, prec=prec)      # dcol = dY W
            dx = ops.col2im3x3s2(dcol, ctx.xshape)
        return None, dx, dw, db


class BnActFn(Function):
    """relu?(batch_norm(x) * (scale + plus_one) + shift) over the last dim's channels; scale / shift [C] or None."""

    @staticmethod
    def forward(ctx, x_in, scale_in, shift, plus_one, relu, eps, run_mean, run_var, momentum):
        x = x_in.contiguous()
        scale = None if scale_in is None else scale_in.contiguous()
        shift = None if shift is None else shift.contiguous()
        y, mean, rstd = ops.bn_act_fwd(x, scale, shift, plus_one, relu, eps, run_mean, run_var, momentum)
        ctx.cfg = (plus_one, relu, scale is not None, shift is not None)
        ctx.save_for_backward(x_in, y, mean, rstd, scale_in)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, scale = ctx.saved_tensors
        plus_one, relu, has_scale, has_shift = ctx.cfg
        if torch.is_grad_enabled():                     # create_graph=True: the gate (y > 0) is a constant
            dx, dscale, dshift = second_order.BnActBwdP.apply(x, dy, scale, y.detach(), mean, rstd, plus_one, relu)
        else:
            dx, dscale, dshift = ops.bn_act_bwd(dy.contiguous(), y, x.contiguous(), mean, rstd,
                                                None if scale is None else scale.contiguous(), plus_one, relu)
        return dx, (dscale if has_scale else None), (dshift if has_shift else None), None, None, None, None, None, None


def _nhwc_input(x):
    """NCHW images -> NHWC (a view for one channel, one transposing copy kernel otherwise)."""
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32):
        raise RuntimeError("MMAML conv nets: expected a float32 CUDA tensor (the B200 path has no CPU fallback)")
    N, C, H, W = x.shape
    if C == 1:
        return x.reshape(N, H, W, 1)
    return x.permute(0, 2, 3, 1).contiguous()


def _mean_rows(t, groups):
    """[groups * n, D] -> [groups, D] mean over n (context-aggregation kernel, mode 0)."""
    D = t.shape[-1]
    return engine.AggregateFn.apply(0, t.reshape(groups, -1, D))


def gated_conv_forward(model, x, params, embeddings):
    """GatedConvModel.forward(x, params, embeddings) (gated_conv_net.py:161-212), stride-2 / affine-FiLM variant."""
    prec = engine.PRECISION
    h = _nhwc_input(x)
    N = h.shape[0]
    for i in range(1, 5):
        w, b = params[f"features.layer{i}_conv.weight"], params[f"features.layer{i}_conv.bias"]
        bn = getattr(model.features, f"layer{i}_bn")
        h = Conv3x3S2Fn.apply(prec, h, w, b)
        gamma = beta = None
        if embeddings is not None:
            e = embeddings[i - 1].reshape(-1)
            Cc = h.shape[-1]
            gamma, beta = e[:Cc], e[Cc:2 * Cc]        # torch.split(embedding, C, dim=-1), :156
        h = BnActFn.apply(h, gamma, beta, 1.0, True, 1e-5, bn.running_mean, bn.running_var, 0.1)   # F.batch_norm defaults, as called at :186-189
    feat = _mean_rows(h.reshape(N * h.shape[1] * h.shape[2], h.shape[3]), N)      # mean over the map, :203-205
    return engine.LinearFn.apply(ACT_TANH, prec, feat, None, params["classifier.fully_connected.weight"],
                                 params["classifier.fully_connected.bias"])


def conv_embedding_forward(model, x, params, return_task_embedding=False):
    """ConvEmbeddingModel.forward (conv_embedding_model.py:99-184): convolutional, batch-norm, avgpool_after_conv,
    no RNN aggregation."""
    prec = engine.PRECISION
    h = _nhwc_input(x)
    N = h.shape[0]
    for i in range(1, model._num_conv + 1):
        bn = getattr(model.conv, f"bn{i}")
        h = Conv3x3S2Fn.apply(prec, h, params[f"conv.conv{i}.weight"], params[f"conv.conv{i}.bias"])
        h = BnActFn.apply(h, params[f"conv.bn{i}.weight"], params[f"conv.bn{i}.bias"], 0.0, True, 1e-5,
                          bn.running_mean, bn.running_var, 0.1)
    feat = _mean_rows(h.reshape(N * h.shape[1] * h.shape[2], h.shape[3]), N)
    hid = engine.LinearFn.apply(ACT_RELU, prec, feat, None, params["linear.weight"], params["linear.bias"])
    if model._embedding_pooling == "avg":
        pooled = engine.AggregateFn.apply(0, hid.reshape(1, N, -1))               # avg_pool1d over the samples, :152
    elif model._embedding_pooling == "max":
        pooled = engine.AggregateFn.apply(1, hid.reshape(1, N, -1))
    else:
        raise NotImplementedError
    outs = [engine.LinearFn.apply(ACT_NONE, prec, pooled, None, params[f"_embeddings.{j}.weight"],
                                  params[f"_embeddings.{j}.bias"]) for j in range(len(model._embeddings))]
    return (outs, pooled) if return_task_embedding else outs
