"""Differentiable backward passes: what the autograd Functions of the path return when they are differentiated with
``create_graph=True``.

The reference's MAML-family trainers take gradients THROUGH the inner-loop gradient step
(trainer/meta_learner_reg.py:116-130 ``torch.autograd.grad(loss, params, create_graph=not first_order)``, and train.py:99
builds the MMAML learner with ``first_order=False``).  A custom Function whose backward launches raw kernels hands
autograd gradients without a graph, i.e. it silently drops those second-order terms.  Here every backward op is itself a
Function over the same C-ABI kernels, and the Functions are closed under differentiation:

  ConvBwdP    (dx, dW, db) of the 3x3 stride-2 convolution with an explicit derivative (second order; one Function
              per conv layer instead of five, because the loop that uses it is bound by the host)
  MatmulP     C = op(A) op(B)            backward = two MatmulP          (any order)
  Im2colP / Col2imP   adjoint pair       backward = the other one        (any order)
  ColSumP / BcastRowsP  adjoint pair                                     (any order)
  MeanFwdP / MeanBwdP   adjoint pair                                     (any order)
  ActBwdP     dz = dy * act'(y)          backward: ActBwdP (dy), -2 y dy v (tanh, y)   (second order)
  BnActBwdP   batch-stat norm + scale/shift + ReLU backward; its derivative is a closed form on three kernels
              (`b200np_bn_act_bwd2`)     (second order)
  LossBwdP    d loss / d mu for the azimuth MSE loss (trainer/losses.py:59-61), constant Hessian 2/R   (second order)

The first-order Functions (mmaml.Conv3x3S2Fn, mmaml.BnActFn, engine.LinearFn, engine.AggregateFn, trainer.losses._LossFn)
switch to these inside ``backward`` when ``torch.is_grad_enabled()`` -- which autograd sets exactly when the caller asked
for ``create_graph=True`` -- and keep their fused single-pass kernels otherwise.  Orders beyond the second raise.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .lib import ACT_NONE, ACT_RELU, ACT_TANH


def _p(t, off=0):
    return t.data_ptr() + 4 * off


def _third_order(*_):
    raise NotImplementedError("the B200 path differentiates its backward once (second-order MAML); third-order "
                              "derivatives of this op are not implemented")


class MatmulP(Function):
    """C = op(A) @ op(B) for 2-D A, B (op = transpose when ta / tb) on the strided GEMM."""

    @staticmethod
    def forward(ctx, prec, A, B, ta, tb):
        A, B = A.contiguous(), B.contiguous()
        M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
        Kb, N = (B.shape[1], B.shape[0]) if tb else (B.shape[0], B.shape[1])
        assert K == Kb, (A.shape, B.shape, ta, tb)
        a_rs, a_cs = (1, A.shape[1]) if ta else (A.shape[1], 1)
        b_rs, b_cs = (1, B.shape[1]) if tb else (B.shape[1], 1)
        Cm = ops.empty((M, N), A)
        ops.gemm(_p(A), _p(B), _p(Cm), M, N, K, a_rs, a_cs, b_rs, b_cs, N, prec=prec)
        ctx.cfg = (prec, ta, tb)
        ctx.save_for_backward(A, B)
        return Cm

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        prec, ta, tb = ctx.cfg
        dA = dB = None
        if ctx.needs_input_grad[1]:   # d op(A) = dC op(B)^T
            dA = MatmulP.apply(prec, B, dC, tb, True) if ta else MatmulP.apply(prec, dC, B, False, not tb)
        if ctx.needs_input_grad[2]:   # d op(B) = op(A)^T dC
            dB = MatmulP.apply(prec, dC, A, True, ta) if tb else MatmulP.apply(prec, A, dC, not ta, False)
        return None, dA, dB, None, None


class Im2colP(Function):
    """x NHWC [N,H,W,C] -> col [N*(H/2)*(W/2), C*9] of the 3x3 stride-2 pad-1 convolution."""

    @staticmethod
    def forward(ctx, x):
        ctx.xshape = tuple(x.shape)
        return ops.im2col3x3s2(x.contiguous())

    @staticmethod
    def backward(ctx, dcol):
        return Col2imP.apply(dcol, ctx.xshape)


class Col2imP(Function):
    @staticmethod
    def forward(ctx, dcol, xshape):
        return ops.col2im3x3s2(dcol.contiguous(), xshape)

    @staticmethod
    def backward(ctx, v):
        return Im2colP.apply(v), None


class ColSumP(Function):
    """[M, C] -> [C]"""

    @staticmethod
    def forward(ctx, t):
        t = t.contiguous()
        ctx.M = t.shape[0]
        return ops.colsum(t, t.shape[0], t.shape[1], t.shape[1])

    @staticmethod
    def backward(ctx, v):
        return BcastRowsP.apply(v, ctx.M)


class BcastRowsP(Function):
    """[C] -> [M, C]"""

    @staticmethod
    def forward(ctx, v, M):
        return ops.repeat_rows(v.contiguous().view(1, -1), M)

    @staticmethod
    def backward(ctx, g):
        return ColSumP.apply(g), None


class MeanFwdP(Function):
    """[T, n, D] -> [T, D], mean over n."""

    @staticmethod
    def forward(ctx, feats):
        ctx.n = feats.shape[1]
        return ops.ctx_aggregate_fwd(feats.contiguous(), 0)[0]

    @staticmethod
    def backward(ctx, dout):
        return MeanBwdP.apply(dout, ctx.n)


class MeanBwdP(Function):
    """[T, D] -> [T, n, D], every row / n (the backward of the mean)."""

    @staticmethod
    def forward(ctx, dout, n):
        T, D = dout.shape
        return ops.ctx_aggregate_bwd(dout.contiguous(), None, T, n, D, 0)

    @staticmethod
    def backward(ctx, v):
        return MeanFwdP.apply(v), None


class _Mul3Once(Function):
    @staticmethod
    def forward(ctx, a, b, c, alpha):
        return ops.mul3(a.contiguous(), b.contiguous(), c.contiguous(), alpha)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        _third_order()


class ActBwdP(Function):
    """dz = dy * act'(y), y = act(z) the forward output (ReLU: y > 0; tanh: 1 - y^2)."""

    @staticmethod
    def forward(ctx, act, dy, y):
        dy, y = dy.contiguous(), y.contiguous()
        ctx.act = act
        ctx.save_for_backward(dy, y)
        return ops.act_bwd(dy, y, act)

    @staticmethod
    def backward(ctx, v):
        dy, y = ctx.saved_tensors
        d_dy = ActBwdP.apply(ctx.act, v, y) if ctx.needs_input_grad[1] else None
        d_y = None
        if ctx.act == ACT_TANH and ctx.needs_input_grad[2]:
            d_y = _Mul3Once.apply(y, dy, v, -2.0)        # d/dy [dy (1 - y^2)] = -2 y dy
        return None, d_dy, d_y


class BnActBwdP(Function):
    """(dx, dscale, dshift) of relu?(batch_norm(x) * (scale + plus_one) + shift); `y`, `mean`, `rstd` come from the
    forward and are constants here (the gate is piecewise constant; the closed form below accounts for the dependence
    of mean / rstd on x)."""

    @staticmethod
    def forward(ctx, x, dy, scale, y, mean, rstd, plus_one, relu):
        x, dy = x.contiguous(), dy.contiguous()
        scale = None if scale is None else scale.contiguous()
        dx, dscale, dshift = ops.bn_act_bwd(dy, y, x, mean, rstd, scale, plus_one, relu)
        ctx.cfg = (plus_one, relu)
        ctx.save_for_backward(x, dy, scale, y, mean, rstd)
        return dx, dscale, dshift

    @staticmethod
    @once_differentiable
    def backward(ctx, vx, vs, vt):
        x, dy, scale, y, mean, rstd = ctx.saved_tensors
        plus_one, relu = ctx.cfg
        need = ctx.needs_input_grad
        c = lambda t: None if t is None else t.contiguous()
        gx, gdy, gscale = ops.bn_act_bwd2(dy, y, x, mean, rstd, scale, plus_one, relu, c(vx), c(vs), c(vt),
                                          want_x=need[0], want_dy=need[1], want_scale=need[2] and scale is not None)
        return gx, gdy, gscale, None, None, None, None, None


class LossBwdP(Function):
    """d loss / d mu scaled by the incoming scalar gradient g."""

    @staticmethod
    def forward(ctx, kind, mu, y, g):
        mu, y = mu.contiguous(), y.contiguous()
        _, dmu = ops.loss_fwd_bwd(mu, y, kind)
        g1 = g.contiguous().view(1)
        ctx.kind, ctx.rows = kind, mu.numel() // mu.shape[-1]
        ctx.save_for_backward(dmu, g1)
        ctx.gshape = g.shape
        return ops.scale_by_device_scalar(dmu, g1)

    @staticmethod
    @once_differentiable
    def backward(ctx, v):
        if ctx.kind != 2:
            raise NotImplementedError("second-order gradients are implemented for the azimuth MSE loss "
                                      "(task 'shapenet_1d', trainer/losses.py:59-61) only")
        dmu, g1 = ctx.saved_tensors
        v = v.contiguous()
        d_mu = d_g = None
        if ctx.needs_input_grad[1]:      # Hessian of mean_r sum_k (t - mu)^2 is (2 / R) I
            d_mu = ops.mul3(ops.scale_by_device_scalar(v, g1), alpha=2.0 / ctx.rows)
        if ctx.needs_input_grad[3]:
            d_g = ops.reduce(ops.mul3(dmu, v).view(-1), 1).view(ctx.gshape)
        return None, d_mu, None, d_g


# ------------------------------------------------------------------------------------------------
# the differentiable backward passes of the first-order Functions
# ------------------------------------------------------------------------------------------------
class ConvBwdP(Function):
    """(dx, dW, db) of y = conv3x3_s2(x NHWC, W [Cout,Cin,3,3]) + b in ONE Function with an explicit derivative, instead of
    the five generic ops it is made of: the reference's meta-learning loop is bound by the host, and a Python
    `autograd.Function` costs ~17 us per application (`tools/bench_mmaml.py`).

    With X = im2col(x) [M,K], G = dy [M,Cout]:   dW = G^T X,   dx = col2im(G W),   db = colsum(G).
    For cotangents (Vx, Vw, Vb) and Vc = im2col(Vx):
        d/dG = X Vw^T + Vc W^T + 1 Vb^T        d/dx = col2im(G Vw)        d/dW = G^T Vc
    `col` is the im2col matrix the forward pass saved (X as a value; the graph goes through `x`)."""

    @staticmethod
    def forward(ctx, prec, x, w, dy, col, need_x):
        x, w, dy = x.contiguous(), w.contiguous(), dy.contiguous()
        Cout = w.shape[0]
        M, K = col.shape
        dw = ops.empty(tuple(w.shape), w)
        ops.gemm(_p(dy), _p(col), _p(dw), Cout, K, M, 1, Cout, K, 1, K, prec=prec)          # dW = G^T X
        db = ops.colsum(dy, M, Cout, Cout)
        dx = None
        if need_x:
            dcol = ops.empty((M, K), dy)
            ops.gemm(_p(dy), _p(w), _p(dcol), M, K, Cout, Cout, 1, K, 1, K, prec=prec)      # dcol = G W
            dx = ops.col2im3x3s2(dcol, tuple(x.shape))
        ctx.cfg = (prec, tuple(x.shape), need_x)
        ctx.save_for_backward(w, dy, col)
        ctx.set_materialize_grads(False)
        return dx, dw, db

    @staticmethod
    @once_differentiable
    def backward(ctx, vx, vw, vb):
        w, dy, col = ctx.saved_tensors
        prec, xshape, need_x = ctx.cfg
        Cout = w.shape[0]
        M, K = col.shape
        need = ctx.needs_input_grad       # (prec, x, w, dy, col, need_x)
        vc = ops.im2col3x3s2(vx.contiguous()) if vx is not None else None
        vw = None if vw is None else vw.contiguous()
        g_x = g_w = g_dy = None
        if need[3]:
            g_dy = ops.empty((M, Cout), dy)
            have = False
            if vw is not None:                                                               # X Vw^T (+ 1 Vb^T)
                ops.gemm(_p(col), _p(vw), _p(g_dy), M, Cout, K, K, 1, 1, K, Cout, prec=prec,
                         bias=None if vb is None else _p(vb.contiguous()))
                have = True
            if vc is not None:                                                               # + Vc W^T
                ops.gemm(_p(vc), _p(w), _p(g_dy), M, Cout, K, K, 1, 1, K, Cout, prec=prec, beta=1.0 if have else 0.0,
                         bias=None if (have or vb is None) else _p(vb.contiguous()))
                have = True
            if not have:
                g_dy = ops.repeat_rows(vb.contiguous().view(1, -1), M) if vb is not None else None
            if g_dy is not None:
                g_dy = g_dy.view(dy.shape)
        if need[1] and need_x and vw is not None:                                            # col2im(G Vw)
            t = ops.empty((M, K), dy)
            ops.gemm(_p(dy), _p(vw), _p(t), M, K, Cout, Cout, 1, K, 1, K, prec=prec)
            g_x = ops.col2im3x3s2(t, xshape)
        if need[2] and vc is not None:                                                       # G^T Vc
            g_w = ops.empty(tuple(w.shape), w)
            ops.gemm(_p(dy), _p(vc), _p(g_w), Cout, K, M, 1, Cout, K, 1, K, prec=prec)
        return None, g_x, g_w, g_dy, None, None


def conv3x3s2_backward(prec, x, w, dy, need_x, col=None):
    """-> (dx | None, dw, db) of y = conv3x3_s2(x NHWC, w [Cout,Cin,3,3]) + b, differentiable once more."""
    if col is None:
        col = ops.im2col3x3s2(x.detach().contiguous())
    return ConvBwdP.apply(prec, x, w, dy, col, need_x)


def linear_backward(act, prec, x, w, y, dy, need_x):
    """-> (dx | None, dw, db) of y = act(x w^T + b)."""
    N = w.shape[0]
    dy2 = dy.reshape(-1, N)
    dz = dy2 if act == ACT_NONE else ActBwdP.apply(act, dy2, y.reshape(-1, N))
    x2 = x.reshape(-1, x.shape[-1])
    dw = MatmulP.apply(prec, dz, x2, True, False)
    db = ColSumP.apply(dz)
    dx = MatmulP.apply(prec, dz, w, False, False).view_as(x) if need_x else None
    return dx, dw, db
