import os, sys
os.environ["B200NP_MMAML"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
from b200np import engine, mmaml
from b200np.lib import ACT_TANH
from test_mmaml import _build_models, N_IMG, SEED
from oracle import synth, mmaml_oracle
def rel(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-300))
engine.set_precision("fp32")
model, emb = _build_models(); model = model.cuda(); emb.to("cuda")
cx, cy, _, _ = synth.task_batch("shapenet_1d", 1, N_IMG, 1, seed=SEED)
x = torch.from_numpy(cx[0]).cuda(); y = torch.from_numpy(cy[0]).cuda()
P = dict(model.named_parameters())
embs = emb(x)
for e in embs: e.retain_grad()
h = x.reshape(15, 128, 128, 1); conv_out, bn_out = [], []
for i in range(1, 5):
    c = mmaml.Conv3x3S2Fn.apply(engine.PRECISION, h, P[f"features.layer{i}_conv.weight"], P[f"features.layer{i}_conv.bias"]); c.retain_grad(); conv_out.append(c)
    C = c.shape[-1]; e = embs[i - 1].reshape(-1)
    h = mmaml.BnActFn.apply(c, e[:C], e[C:2 * C], 1.0, True, 1e-5, None, None, 0.1); h.retain_grad(); bn_out.append(h)
feat = mmaml._mean_rows(h.reshape(15 * 64, 256), 15)
logits = engine.LinearFn.apply(ACT_TANH, engine.PRECISION, feat, None, P["classifier.fully_connected.weight"], P["classifier.fully_connected.bias"])
loss = torch.mean(torch.sum((y[..., :2] - logits) ** 2, dim=-1)); loss.backward()
# torch fp64 on the GPU
Pd = {k: v.detach().double().requires_grad_(True) for k, v in P.items()}
Ped = {k: v.detach().double().requires_grad_(True) for k, v in emb.named_parameters()}
embd, _ = mmaml_oracle.conv_embedding(Ped, x.double())
for e in embd: e.retain_grad()
hd = x.double(); co_d, bo_d = [], []
for i in range(1, 5):
    c = F.conv2d(hd, Pd[f"features.layer{i}_conv.weight"], Pd[f"features.layer{i}_conv.bias"], stride=2, padding=1); c.retain_grad(); co_d.append(c)
    C = c.shape[1]; e = embd[i - 1].reshape(-1)
    hd = F.relu(F.batch_norm(c, None, None, training=True) * (1 + e[:C].view(1, -1, 1, 1)) + e[C:].view(1, -1, 1, 1)); hd.retain_grad(); bo_d.append(hd)
featd = hd.reshape(15, 256, 64).mean(dim=2)
logd = torch.tanh(F.linear(featd, Pd["classifier.fully_connected.weight"], Pd["classifier.fully_connected.bias"]))
lossd = torch.mean(torch.sum((y.double()[..., :2] - logd) ** 2, dim=-1)); lossd.backward()
print("loss", float(loss), float(lossd), "logits", rel(logits, logd))
for i in range(4):
    a = (bn_out[i].permute(0,3,1,2) > 0); b = (bo_d[i] > 0)
    print(f"layer{i+1}: emb {rel(embs[i], embd[i]):.2e} conv_out {rel(conv_out[i].permute(0,3,1,2), co_d[i]):.2e} bn_out {rel(bn_out[i].permute(0,3,1,2), bo_d[i]):.2e} "
          f"d(bn_out) {rel(bn_out[i].grad.permute(0,3,1,2), bo_d[i].grad):.2e} d(conv_out) {rel(conv_out[i].grad.permute(0,3,1,2), co_d[i].grad):.2e} "
          f"d(emb) {rel(embs[i].grad, embd[i].grad):.2e} dW {rel(P[f'features.layer{i+1}_conv.weight'].grad, Pd[f'features.layer{i+1}_conv.weight'].grad):.2e} gate mismatches {int((a != b).sum())}")
