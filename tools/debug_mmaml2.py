import os, sys
os.environ["B200NP_MMAML"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
from b200np import engine, mmaml
from test_mmaml import _build_models, N_IMG, SEED
from oracle import synth
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-300))
engine.set_precision("fp32")
model, emb = _build_models(); model = model.cuda()
cx, cy, _, _ = synth.task_batch("shapenet_1d", 1, N_IMG, 1, seed=SEED)
x = torch.from_numpy(cx[0]).cuda()
g = torch.Generator().manual_seed(0)
embs = [torch.randn(1, 2 * c, generator=g).cuda() * 0.3 for c in (32, 64, 128, 256)]
P = dict(model.named_parameters())
# ours
h = x.reshape(15, 128, 128, 1); conv_out, bn_out = [], []
for i in range(1, 5):
    c = mmaml.Conv3x3S2Fn.apply(engine.PRECISION, h, P[f"features.layer{i}_conv.weight"], P[f"features.layer{i}_conv.bias"]); c.retain_grad(); conv_out.append(c)
    C = c.shape[-1]; e = embs[i - 1].reshape(-1)
    h = mmaml.BnActFn.apply(c, e[:C], e[C:], 1.0, True, 1e-5, None, None, 0.1); h.retain_grad(); bn_out.append(h)
MODE = os.environ.get("DBG_MODE", "mean")
if MODE == "mean":
    feat = h.reshape(15, 64, 256).mean(dim=1)                       # torch op: constant-over-the-map upstream gradient
    wtop = torch.randn(256, 2, generator=g).cuda()
    (torch.tanh(feat @ wtop)).pow(2).sum().backward()
    dfin = None
else:
    dfin = torch.randn(h.shape, generator=g).cuda()
    h.backward(dfin)
# torch fp64
hd = x.double(); co_d, bo_d = [], []
Pd = {k: v.detach().double().requires_grad_(True) for k, v in P.items()}
for i in range(1, 5):
    c = F.conv2d(hd, Pd[f"features.layer{i}_conv.weight"], Pd[f"features.layer{i}_conv.bias"], stride=2, padding=1); c.retain_grad(); co_d.append(c)
    C = c.shape[1]; e = embs[i - 1].double().reshape(-1)
    hd = F.relu(F.batch_norm(c, None, None, training=True) * (1 + e[:C].view(1, -1, 1, 1)) + e[C:].view(1, -1, 1, 1)); hd.retain_grad(); bo_d.append(hd)
if MODE == "mean":
    featd = hd.reshape(15, 256, 64).mean(dim=2)
    (torch.tanh(featd @ wtop.double())).pow(2).sum().backward()
else:
    hd.backward(dfin.permute(0, 3, 1, 2).double())
for i in range(4):
    print(f"layer{i+1}: conv_out {rel(conv_out[i].permute(0,3,1,2), co_d[i]):.2e} bn_out {rel(bn_out[i].permute(0,3,1,2), bo_d[i]):.2e} "
          f"d(bn_out) {rel(bn_out[i].grad.permute(0,3,1,2), bo_d[i].grad):.2e} d(conv_out) {rel(conv_out[i].grad.permute(0,3,1,2), co_d[i].grad):.2e} "
          f"dW {rel(P[f'features.layer{i+1}_conv.weight'].grad, Pd[f'features.layer{i+1}_conv.weight'].grad):.2e}")
    # gate disagreement
    a = (bn_out[i].permute(0,3,1,2) > 0); b = (bo_d[i] > 0)
    print("   gate mismatches", int((a != b).sum()), "of", a.numel())
