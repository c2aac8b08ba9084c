"""Second-order building blocks (b200np/second_order.py): every differentiable backward op, and the grad-mode branches
of the first-order Functions, against torch autograd differentiating the same mathematics twice in fp64.

Pattern: first = grad(f(inputs), inputs, cotangent, create_graph=True); L = sum_i <first_i, probe_i>; second =
grad(L, inputs) -- once through the B200 Functions on CUDA (fp32 kernels), once through plain torch ops in fp64."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float64) * scale


def double_grads(fn, inputs, cot, probes):
    """inputs: leaf tensors (requires_grad); -> (first-order grads, grads of sum <first_i, probe_i> w.r.t. inputs + cot)."""
    out = fn(*inputs)
    cot = cot.clone().requires_grad_()
    first = torch.autograd.grad(out, inputs, cot, create_graph=True, allow_unused=True)
    L = sum((f * p).sum() for f, p in zip(first, probes) if f is not None)
    second = torch.autograd.grad(L, list(inputs) + [cot], allow_unused=True)
    return [f.detach() for f in first], second


def compare(fn_b200, fn_ref, shapes, out_shape, tol=2e-5, seed=0):
    ins64 = [rnd(*s, seed=seed + i) for i, s in enumerate(shapes)]
    cot64 = rnd(*out_shape, seed=seed + 50)
    probes64 = [rnd(*s, seed=seed + 100 + i) for i, s in enumerate(shapes)]
    ref1, ref2 = double_grads(fn_ref, [t.clone().requires_grad_() for t in ins64], cot64, probes64)
    cu = lambda t: t.float().cuda()
    got1, got2 = double_grads(fn_b200, [cu(t).requires_grad_() for t in ins64], cu(cot64), [cu(p) for p in probes64])
    _check(got1, ref1, tol, "first")
    _check(got2, ref2, tol, "second")


def _check(got, ref, tol, what):
    """Per tensor rel-L2; tensors whose true gradient vanishes (a bias in front of a batch norm, the piecewise-constant
    ReLU gate) must be None or negligible against the largest gradient of the group."""
    big = max(float(b.norm()) for b in ref if b is not None)
    for i, (a, b) in enumerate(zip(got, ref)):
        if b is None or float(b.norm()) < 1e-9 * big:
            assert a is None or float(a.norm()) < 1e-4 * big, (what, i)
            continue
        assert a is not None, f"{what}-order gradient {i} missing"
        e = rel_l2(a.cpu().numpy(), b.numpy())
        assert e < tol, (what, i, e)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_matmul_primitive(ta, tb):
    from b200np import lib
    from b200np.second_order import MatmulP
    M, K, N = 37, 53, 29
    sa, sb = ((K, M) if ta else (M, K)), ((N, K) if tb else (K, N))
    op = lambda t, tr: t.t() if tr else t
    compare(lambda a, b: MatmulP.apply(lib.PREC_FP32_SIMT, a, b, ta, tb), lambda a, b: op(a, ta) @ op(b, tb), [sa, sb], (M, N))


def _conv_ref(x_nhwc, w, b):
    return F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, b, stride=2, padding=1).permute(0, 2, 3, 1)


@pytest.mark.parametrize("cin", [8, 3, 1])
@pytest.mark.parametrize("prec", ["fp32", "tf32x3"])
def test_conv_function_second_order(prec, cin):
    from b200np import lib
    from b200np.mmaml import Conv3x3S2Fn
    P = {"fp32": lib.PREC_FP32_SIMT, "tf32x3": lib.PREC_TF32X3}[prec]
    compare(lambda x, w, b: Conv3x3S2Fn.apply(P, x, w, b), _conv_ref, [(3, 16, 16, cin), (12, cin, 3, 3), (12,)], (3, 8, 8, 12))


@pytest.mark.parametrize("C", [1, 3, 8])
def test_im2col_col2im_are_adjoint(C):
    """<im2col(x), c> == <x, col2im(c)> and both match F.unfold / F.fold for any channel count."""
    from b200np import ops
    x = rnd(2, 8, 8, C, seed=1).float().cuda()
    c = rnd(2 * 4 * 4, C * 9, seed=2).float().cuda()
    col, back = ops.im2col3x3s2(x), ops.col2im3x3s2(c, tuple(x.shape))
    assert abs(float((col * c).sum()) - float((x * back).sum())) < 1e-4 * float((col * c).abs().sum())
    ref = F.unfold(x.permute(0, 3, 1, 2).double(), 3, padding=1, stride=2)          # [N, C*9, L], k = ci*9 + r*3 + s
    assert rel_l2(col.cpu().numpy(), ref.permute(0, 2, 1).reshape(-1, C * 9).cpu().numpy()) < 1e-6
    fold = F.fold(c.double().view(2, 16, C * 9).permute(0, 2, 1), (8, 8), 3, padding=1, stride=2).permute(0, 2, 3, 1)
    assert rel_l2(back.cpu().numpy(), fold.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("act", ["none", "relu", "tanh"])
def test_linear_function_second_order(act):
    from b200np import lib
    from b200np.engine import LinearFn
    A = {"none": lib.ACT_NONE, "relu": lib.ACT_RELU, "tanh": lib.ACT_TANH}[act]
    f = {"none": lambda z: z, "relu": torch.relu, "tanh": torch.tanh}[act]
    compare(lambda x, w, b: LinearFn.apply(A, lib.PREC_FP32_SIMT, x, None, w, b), lambda x, w, b: f(x @ w.t() + b),
            [(21, 40), (7, 40), (7,)], (21, 7))


def test_mean_aggregation_second_order():
    from b200np.engine import AggregateFn
    # linear op: the second-order gradients w.r.t. the input vanish, the one w.r.t. the cotangent is the mean again
    compare(lambda x: AggregateFn.apply(0, x) * AggregateFn.apply(0, x), lambda x: x.mean(1) * x.mean(1), [(4, 9, 16)], (4, 16))


def test_azimuth_loss_second_order():
    from trainer.losses import LossFunc
    lossf = LossFunc("mse", "shapenet_1d")
    y = rnd(33, 2, seed=9)
    yc = y.float().cuda()
    compare(lambda mu: lossf.calc_loss(mu, None, yc), lambda mu: torch.mean(torch.sum((y - mu) ** 2, dim=-1)), [(33, 2)], ())


@pytest.mark.parametrize("film", [True, False])
def test_bn_act_function_second_order(film):
    from b200np.mmaml import BnActFn

    def ref(x, sc, sh):
        mu, var = x.mean(0), x.var(0, unbiased=False)
        return torch.relu((x - mu) * (var + 1e-5).rsqrt() * (sc + (1.0 if film else 0.0)) + sh)

    compare(lambda x, sc, sh: BnActFn.apply(x, sc, sh, 1.0 if film else 0.0, True, 1e-5, None, None, 0.1), ref,
            [(300, 24), (24,), (24,)], (300, 24))


def test_two_layer_stack_second_order():
    """conv -> norm+FiLM+ReLU -> conv -> norm+ReLU -> mean -> dense tanh: the composition the MMAML nets use."""
    from b200np import lib
    from b200np.engine import AggregateFn, LinearFn
    from b200np.mmaml import BnActFn, Conv3x3S2Fn
    P = lib.PREC_FP32_SIMT

    def b200(x, w1, b1, g1, t1, w2, b2, wf, bf):
        h = BnActFn.apply(Conv3x3S2Fn.apply(P, x, w1, b1), g1, t1, 1.0, True, 1e-5, None, None, 0.1)
        h = BnActFn.apply(Conv3x3S2Fn.apply(P, h, w2, b2), None, None, 1.0, True, 1e-5, None, None, 0.1)   # plain norm: scale' = 0 + 1
        n = h.shape[0]
        feat = AggregateFn.apply(0, h.reshape(n, -1, h.shape[-1]))
        return LinearFn.apply(lib.ACT_TANH, P, feat, None, wf, bf)

    def bn(z):
        mu, var = z.mean((0, 1, 2)), z.var((0, 1, 2), unbiased=False)
        return (z - mu) * (var + 1e-5).rsqrt()

    def ref(x, w1, b1, g1, t1, w2, b2, wf, bf):
        h = torch.relu(bn(_conv_ref(x, w1, b1)) * (1 + g1) + t1)
        h = torch.relu(bn(_conv_ref(h, w2, b2)))
        return torch.tanh(h.mean((1, 2)) @ wf.t() + bf)

    compare(b200, ref, [(5, 16, 16, 3), (8, 3, 3, 3), (8,), (8,), (8,), (12, 8, 3, 3), (12,), (2, 12), (2,)], (5, 2), tol=1e-4)


def test_maml_step_through_the_stack():
    """One inner gradient step with create_graph=True, then the outer loss: the gradient of the outer loss w.r.t. the
    initial parameters and the FiLM inputs contains every second-order term of conv / norm / mean / dense / tanh / loss."""
    from b200np import lib
    from b200np.engine import AggregateFn, LinearFn
    from b200np.mmaml import BnActFn, Conv3x3S2Fn
    from trainer.losses import LossFunc
    P, lr = lib.PREC_FP32_SIMT, 0.1
    lossf = LossFunc("mse", "shapenet_1d")
    x64, xv64, y64, yv64 = rnd(6, 16, 16, 2, seed=1), rnd(6, 16, 16, 2, seed=2), rnd(6, 2, seed=3), rnd(6, 2, seed=4)

    def net_b200(x, w1, b1, g1, t1, w2, b2, wf, bf):
        h = BnActFn.apply(Conv3x3S2Fn.apply(P, x, w1, b1), g1, t1, 1.0, True, 1e-5, None, None, 0.1)
        h = BnActFn.apply(Conv3x3S2Fn.apply(P, h, w2, b2), g1.new_zeros(12), t1.new_zeros(12), 1.0, True, 1e-5, None, None, 0.1)
        feat = AggregateFn.apply(0, h.reshape(h.shape[0], -1, h.shape[-1]))
        return LinearFn.apply(lib.ACT_TANH, P, feat, None, wf, bf)

    def bn(z):
        mu, var = z.mean((0, 1, 2)), z.var((0, 1, 2), unbiased=False)
        return (z - mu) * (var + 1e-5).rsqrt()

    def net_ref(x, w1, b1, g1, t1, w2, b2, wf, bf):
        h = torch.relu(bn(_conv_ref(x, w1, b1)) * (1 + g1) + t1)
        h = torch.relu(bn(_conv_ref(h, w2, b2)))
        return torch.tanh(h.mean((1, 2)) @ wf.t() + bf)

    def outer(net, loss, x, y, xv, yv, w1, b1, g1, t1, w2, b2, wf, bf):
        params = [w1, b1, w2, b2, wf, bf]
        inner = loss(net(x, w1, b1, g1, t1, w2, b2, wf, bf), y)
        grads = torch.autograd.grad(inner, params, create_graph=True, allow_unused=True)
        w1, b1, w2, b2, wf, bf = [p if g is None else p - lr * g.clamp(-20, 20) for p, g in zip(params, grads)]
        return loss(net(xv, w1, b1, g1, t1, w2, b2, wf, bf), yv)

    shapes = [(8, 2, 3, 3), (8,), (8,), (8,), (12, 8, 3, 3), (12,), (2, 12), (2,)]
    ins64 = [rnd(*s, seed=20 + i, scale=0.5) for i, s in enumerate(shapes)]
    leaves = [t.clone().requires_grad_() for t in ins64]
    outer(net_ref, lambda p, y: torch.mean(torch.sum((y - p) ** 2, dim=-1)), x64, y64, xv64, yv64, *leaves).backward()
    cu = lambda t: t.float().cuda()
    leaves_c = [cu(t).requires_grad_() for t in ins64]
    outer(net_b200, lambda p, y: lossf.calc_loss(p, None, y), cu(x64), cu(y64), cu(xv64), cu(yv64), *leaves_c).backward()
    _check([t.grad for t in leaves_c], [t.grad for t in leaves], 1e-4, "outer")
