#!/usr/bin/env python
"""Precision calibration on the GPU box (diagnostic; writes gpurun_out/diag_precision.txt).

For one conv and for the whole ANPDistractor step it reports the relative-L2 error against an fp64
CPU evaluation of: our three precision modes, the oracle in fp32 on CPU (the reference's own noise
floor), and the oracle's torch ops on the GPU via cuDNN/cuBLAS in fp32 and TF32 (what the unmodified
reference would do on this box)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
from conftest import CASES, build_product_model, oracle_cfg  # noqa: E402
from oracle import np_oracle, synth  # noqa: E402

out = []


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out.append(s)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def conv_diag():
    from b200np import ops
    g = torch.Generator().manual_seed(0)
    N, H = 8, 32
    h = torch.rand(N, 64, H, H, generator=g, dtype=torch.float64)
    x = torch.rand(N, 64, 2 * H, 2 * H, generator=g, dtype=torch.float64)
    w2 = torch.randn(64, 64, 3, 3, generator=g, dtype=torch.float64) * 0.04
    ws = torch.randn(64, 64, 1, 1, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(64, generator=g, dtype=torch.float64) * 0.1
    ref = F.relu(F.conv2d(h, w2, b, padding=1) + F.conv2d(x, ws, b, stride=2))
    nh = lambda t: t.permute(0, 2, 3, 1).contiguous().float().cuda()
    wf2 = ops.pack_conv_weight(w2.float().cuda())
    wfs = ops.pack_conv_weight(ws.float().cuda())
    bg = b.float().cuda()
    for name, prec in (("fp32-simt", 0), ("tf32x3", 1), ("tf32", 2)):
        y = ops.conv_fwd(nh(h), wf2, bg, 1, 1, prec, skip=(nh(x), wfs, bg, 2))
        P(f"conv2+skip fwd  {name:10s} rel-L2 vs fp64: {rel(y.permute(0, 3, 1, 2), ref):.3e}")
    for tf in (False, True):
        torch.backends.cudnn.allow_tf32 = tf
        y = F.relu(F.conv2d(h.float().cuda(), w2.float().cuda(), bg, padding=1) +
                   F.conv2d(x.float().cuda(), ws.float().cuda(), bg, stride=2))
        P(f"conv2+skip fwd  cudnn tf32={tf!s:5s} rel-L2 vs fp64: {rel(y, ref):.3e}")
    y32 = F.relu(F.conv2d(h.float(), w2.float(), b.float(), padding=1) + F.conv2d(x.float(), ws.float(), b.float(), stride=2))
    P(f"conv2+skip fwd  cpu-fp32   rel-L2 vs fp64: {rel(y32, ref):.3e}")


def convops_diag():
    """dgrad / wgrad of one residual stage at a realistic size, every precision mode vs fp64."""
    from b200np import ops
    g = torch.Generator().manual_seed(3)
    N, H = 8, 32
    x = torch.rand(N, 64, 2 * H, 2 * H, generator=g, dtype=torch.float64).requires_grad_()
    w1 = (torch.randn(64, 64, 3, 3, generator=g, dtype=torch.float64) * 0.04).requires_grad_()
    w2 = (torch.randn(64, 64, 3, 3, generator=g, dtype=torch.float64) * 0.04).requires_grad_()
    ws = (torch.randn(64, 64, 1, 1, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
    h = F.relu(F.conv2d(x, w1, None, stride=2, padding=1) - 0.3)
    y = F.relu(F.conv2d(h, w2, None, padding=1) + F.conv2d(x, ws, None, stride=2) - 0.3)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    dz = (gy * (y > 0)).detach()
    hh = h.detach().requires_grad_()
    F.conv2d(hh, w2.detach(), None, padding=1).backward(dz)
    dh_ref = hh.grad * (h > 0)
    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().float().cuda()
    xg, hg, dzg = nh(x), nh(h), nh(dz)
    wd1 = ops.pack_conv_weight(w1.detach().float().cuda())
    wd2 = ops.pack_conv_weight(w2.detach().float().cuda())
    wds = ops.pack_conv_weight(ws.detach().float().cuda())
    for name, prec in (("fp32-simt", 0), ("tf32x3", 1), ("tf32", 2)):
        dw2, _, _ = ops.conv_wgrad(hg, dzg, 3, 1, prec)
        dws, _, _ = ops.conv_wgrad(xg, dzg, 1, 2, prec, want_db=False)
        dh = ops.conv_dgrad(dzg, wd2, hg.shape, 1, prec, mask_src=hg)
        dw1, _, _ = ops.conv_wgrad(xg, nh(dh_ref), 3, 2, prec)
        dx = ops.conv_dgrad(nh(dh_ref), wd1, xg.shape, 2, prec, mask_src=xg, skip=(dzg, wds, 2))
        P(f"stage bwd {name:10s} dgrad_s1 {rel(dh.permute(0, 3, 1, 2), dh_ref):.2e} dgrad_s2+skip {rel(dx.permute(0, 3, 1, 2), x.grad):.2e} "
          f"wgrad3x3s1 {rel(dw2, w2.grad):.2e} wgrad3x3s2 {rel(dw1, w1.grad):.2e} wgrad1x1s2 {rel(dws, ws.grad):.2e}")


def model_diag(case):
    from b200np import engine
    from trainer.losses import LossFunc
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    batch = synth.task_batch(task, T, nc, nt, seed=11)
    model, cfg = build_product_model(case, device="cuda")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    ocfg = oracle_cfg(cfg)

    def oracle_run(dtype, device):
        tr = np_oracle.OracleTrainer(method, ocfg, sd, dtype=dtype)
        if device != "cpu":
            tr.sd = {k: v.detach().to(device).requires_grad_(v.requires_grad) for k, v in tr.sd.items()}
            tr.params = np_oracle.params_with_grad(tr.sd)
        b = [torch.from_numpy(a).to(dtype).to(device) for a in batch]
        mu, loss = tr.forward_loss(*b)
        loss.backward()
        return mu.detach().cpu(), float(loss), {k: (None if v.grad is None else v.grad.detach().cpu()) for k, v in tr.params.items()}

    mu64, l64, g64 = oracle_run(torch.float64, "cpu")
    runs = {"oracle cpu fp32": oracle_run(torch.float32, "cpu")}
    for tf in (False, True):
        torch.backends.cudnn.allow_tf32 = tf
        torch.backends.cuda.matmul.allow_tf32 = tf
        runs[f"torch-gpu tf32={tf}"] = oracle_run(torch.float32, "cuda")
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    lossf = LossFunc("mse", task)
    for prec in ("fp32", "tf32x3", "tf32"):
        engine.set_precision(prec)
        m, _ = build_product_model(case, device="cuda")
        m = m.to("cuda")
        b = [torch.from_numpy(a).cuda() for a in batch]
        mu, _, _ = m(b[0], b[1], b[2])
        loss = lossf.calc_loss(mu, None, b[3])
        loss.backward()
        runs[f"b200np {prec}"] = (mu.detach().cpu(), float(loss), {k: (None if p.grad is None else p.grad.cpu()) for k, p in m.named_parameters()})
    P(f"== {case}: errors vs fp64 oracle (T={T}, nc={nc}, nt={nt})")
    for name, (mu, loss, grads) in runs.items():
        errs = {k: rel(grads[k], g) for k, g in g64.items() if g is not None}
        flat = torch.cat([grads[k].double().reshape(-1) for k in errs])
        flat64 = torch.cat([g64[k].double().reshape(-1) for k in errs])
        nonq = [v for k, v in errs.items() if "_W_q" not in k]
        wq = [v for k, v in errs.items() if "_W_q" in k]
        worst_k = max((k for k in errs if "_W_q" not in k), key=lambda k: errs[k])
        if name == "b200np tf32x3":
            for k in sorted(errs, key=lambda k: -errs[k])[:14]:
                P(f"      x3 {errs[k]:.2e}  {k}")
        P(f"{name:22s} mu {rel(mu, mu64):.2e} loss {abs(loss - l64) / abs(l64):.2e} | grads: global {rel(flat, flat64):.2e} "
          f"median {np.median(nonq):.2e} worst {max(nonq):.2e} ({worst_k})" + (f" | _W_q worst {max(wq):.2e}" if wq else ""))


def favor_diag():
    from b200np.engine import FavorAttentionFn
    T, H, nt, nc, d = 2, 8, 21, 15, 256
    M = int(d * np.log(d))
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    xq, xk, v, Pm, d_out = r(T, H, nt, d), r(T, H, nc, d), r(T, H, nc, d), r(M, d), r(T, H, nt, d)
    o64, (dq64, dk64, dv64) = np_oracle.favor_attention_fwd_bwd(xq, xk, v, Pm, d_out)

    def autograd(dtype, dev):
        q_, k_, v_ = (t.to(dtype).to(dev).requires_grad_() for t in (xq, xk, v))
        qp = np_oracle.softmax_kernel(q_, Pm.to(dtype).to(dev), True)
        kp = np_oracle.softmax_kernel(k_, Pm.to(dtype).to(dev), False)
        o = np_oracle.linear_attention(qp, kp, v_)
        o.backward(d_out.to(dtype).to(dev))
        return o, q_.grad, k_.grad, v_.grad
    for name, (o, dq, dk, dv) in (("ref-formulation cpu fp32", autograd(torch.float32, "cpu")),
                                  ("ref-formulation gpu fp32", autograd(torch.float32, "cuda"))):
        P(f"favor {name:26s} out {rel(o, o64):.2e} dq {rel(dq, dq64):.2e} dk {rel(dk, dk64):.2e} dv {rel(dv, dv64):.2e}")
    rows = lambda t: t.permute(0, 2, 1, 3).reshape(t.shape[0] * t.shape[2], -1).float().cuda().requires_grad_()
    qg, kg, vg = rows(xq), rows(xk), rows(v)
    o = FavorAttentionFn.apply(0, T, H, nt, nc, qg, kg, vg, Pm.float().cuda())
    o.backward(d_out.permute(0, 2, 3, 1).reshape(T * nt, d * H).float().cuda())
    back = lambda t, n: t.view(T, n, H, d).permute(0, 2, 1, 3)
    P(f"favor {'b200np fused':26s} out {rel(o.view(T, nt, d, H).permute(0, 3, 1, 2), o64):.2e} dq {rel(back(qg.grad, nt), dq64):.2e} "
      f"dk {rel(back(kg.grad, nc), dk64):.2e} dv {rel(back(vg.grad, nc), dv64):.2e}")


if __name__ == "__main__":
    torch.manual_seed(0)
    conv_diag()
    convops_diag()
    for c in sys.argv[1:] or ["cnp_distractor_max", "anp_distractor"]:
        model_diag(c)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "diag_precision.txt"), "w").write("\n".join(out) + "\n")
