import os, sys
os.environ["B200NP_MMAML"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
from b200np import engine, mmaml
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
for prec in ("fp32", "tf32x3"):
    engine.set_precision(prec)
    for (N, Cin, Cout, H) in ((15, 1, 32, 128), (15, 32, 64, 64), (15, 64, 128, 32), (15, 128, 256, 16)):
        g = torch.Generator().manual_seed(Cin)
        x = torch.rand(N, Cin, H, H, generator=g); w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1; b = torch.randn(Cout, generator=g) * 0.1
        dy = torch.randn(N, Cout, H // 2, H // 2, generator=g)
        xd, wd, bd = (t.double().cuda().requires_grad_(True) for t in (x, w, b))
        yd = F.conv2d(xd, wd, bd, stride=2, padding=1); yd.backward(dy.double().cuda())
        xg = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True); wg = w.cuda().requires_grad_(True); bg = b.cuda().requires_grad_(True)
        y = mmaml.Conv3x3S2Fn.apply(engine.PRECISION, xg, wg, bg); y.backward(dy.cuda().permute(0, 2, 3, 1).contiguous())
        print(prec, (N, Cin, Cout, H), "y", rel(y.permute(0, 3, 1, 2), yd), "dw", rel(wg.grad, wd.grad), "db", rel(bg.grad, bd.grad), "dx", rel(xg.grad.permute(0, 3, 1, 2), xd.grad))
    for (R, C) in ((15 * 64 * 64, 32), (15 * 8 * 8, 256)):
        g = torch.Generator().manual_seed(C)
        x = torch.randn(R, C, generator=g) * 2 + 1; sc = torch.randn(C, generator=g) * 0.3; sh = torch.randn(C, generator=g) * 0.3; dy = torch.randn(R, C, generator=g)
        xd, scd, shd = (t.double().cuda().requires_grad_(True) for t in (x, sc, sh))
        yd = F.relu(F.batch_norm(xd, None, None, training=True) * (1 + scd) + shd); yd.backward(dy.double().cuda())
        xg, scg, shg = (t.cuda().requires_grad_(True) for t in (x, sc, sh))
        y = mmaml.BnActFn.apply(xg, scg, shg, 1.0, True, 1e-5, None, None, 0.1); y.backward(dy.cuda())
        print(prec, "bn", (R, C), "y", rel(y, yd), "dx", rel(xg.grad, xd.grad), "dsc", rel(scg.grad, scd.grad), "dsh", rel(shg.grad, shd.grad))
# full nets, every tensor
from test_mmaml import _build_models, _oracle_run, N_IMG, SEED
from oracle import synth
engine.set_precision("fp32")
model, emb = _build_models(); rm, re_ = _build_models()
model = model.cuda(); emb.to("cuda")
cx, cy, _, _ = synth.task_batch("shapenet_1d", 1, N_IMG, 1, seed=SEED)
x = torch.from_numpy(cx[0]).cuda(); y = torch.from_numpy(cy[0]).cuda()
embs = emb(x); logits = model(x, embeddings=embs)
loss = torch.mean(torch.sum((y[..., :2] - logits) ** 2, dim=-1)); loss.backward()
_, l64, _, pm64, pe64 = _oracle_run(rm, re_, torch.float64)
for tag, m, p64 in (("model", model, pm64), ("emb", emb, pe64)):
    for k, p in m.named_parameters():
        print(tag, k, rel(p.grad, p64[k].grad), float(p64[k].grad.norm()))
