"""The CPU oracle (oracle/np_oracle.py) against golden vectors produced by the live reference."""
import numpy as np
import pytest
import torch

from conftest import CASES, build_product_model, fingerprint, oracle_cfg, rel_l2
from oracle import np_oracle, synth


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference_golden(case, golden):
    method, task, agg, img_agg, extra, T, nc, nt = CASES[case]
    model, cfg = build_product_model(case)
    tr = np_oracle.OracleTrainer(method, oracle_cfg(cfg), model.state_dict())
    cx, cy, tx, ty = (torch.from_numpy(a) for a in synth.task_batch(task, T, nc, nt, seed=11))
    inter = {}
    mu, loss = tr.forward_loss(cx, cy, tx, ty, inter)
    loss.backward()
    assert rel_l2(mu.detach().numpy(), golden[f"{case}/mu"]) < 2e-6
    assert abs(loss.item() - float(golden[f"{case}/loss"])) < 2e-6 * abs(float(golden[f"{case}/loss"]))
    grads = tr.grads()
    gkeys = list(golden[f"{case}/grad_keys"])
    assert sorted(k for k, g in grads.items() if g is not None) == sorted(gkeys)
    gfp = golden[f"{case}/grad_fp"]
    for k, ref in zip(gkeys, gfp):
        fp = fingerprint(grads[k])
        # l2 norm and the +-1 probe fingerprint; 1e-3 of the norm is the parity bar, the oracle
        # itself sits orders of magnitude below it except on the query-side attention weights
        # (noise floor of the reference itself)
        tol = 2e-2 if "_W_q" in k else 2e-4
        assert abs(fp[2] - ref[2]) <= tol * ref[2] + 1e-12, (k, fp, ref)
        assert abs(fp[3] - ref[3]) <= tol * ref[2] + 1e-12, (k, fp, ref)
    for k in golden.files:
        if k.startswith(f"{case}/grad/"):
            name = k[len(f"{case}/grad/"):]
            tol = 2e-2 if "_W_q" in name else 2e-4
            assert rel_l2(grads[name].numpy(), golden[k]) < tol, name
    # integer results are bit-exact
    if f"{case}/inter/pool_idx0" in golden.files:
        np.testing.assert_array_equal(inter["pool_idx"].numpy(), golden[f"{case}/inter/pool_idx0"])
    if f"{case}/inter/agg_idx" in golden.files:
        np.testing.assert_array_equal(inter["agg_idx"].numpy(), golden[f"{case}/inter/agg_idx"])
    if f"{case}/loss_test" in golden.files:
        lt = np_oracle.calc_loss(task, mu.detach(), ty, test=True).item()
        assert abs(lt - float(golden[f"{case}/loss_test"])) < 1e-4 * abs(float(golden[f"{case}/loss_test"]))


@pytest.mark.parametrize("task", ["distractor", "shapenet_3d", "shapenet_1d"])
def test_oracle_losses(task, golden):
    mu = torch.from_numpy(golden[f"loss/{task}/mu"]).requires_grad_(True)
    y = torch.from_numpy(golden[f"loss/{task}/y"])
    loss = np_oracle.calc_loss(task, mu, y)
    loss.backward()
    assert abs(loss.item() - float(golden[f"loss/{task}/loss"])) < 1e-6 * abs(float(golden[f"loss/{task}/loss"]))
    assert rel_l2(mu.grad.numpy(), golden[f"loss/{task}/dmu"]) < 1e-6
    if task == "shapenet_1d":
        lt = np_oracle.calc_loss(task, mu.detach(), y, test=True).item()
        assert abs(lt - float(golden[f"loss/{task}/loss_test"])) < 1e-5 * abs(float(golden[f"loss/{task}/loss_test"]))


def test_reassociated_attention_and_closed_form_backward():
    """A=q'k'^T re-association and the closed-form backward (incl. the
    row-argmax / global-argmax routing) equal autograd through the reference formulation."""
    torch.manual_seed(0)
    B, H, nt, nc, d = 3, 2, 5, 4, 64
    M = int(d * np.log(d))
    xq = torch.randn(B, H, nt, d, dtype=torch.float64, requires_grad=True)
    xk = torch.randn(B, H, nc, d, dtype=torch.float64, requires_grad=True)
    v = torch.randn(B, H, nc, d, dtype=torch.float64, requires_grad=True)
    P = torch.randn(M, d, dtype=torch.float64)
    qp = np_oracle.softmax_kernel(xq, P, True)
    kp = np_oracle.softmax_kernel(xk, P, False)
    out = np_oracle.linear_attention(qp, kp, v)
    out2 = np_oracle.linear_attention_reassoc(qp, kp, v)
    assert rel_l2(out2.detach().numpy(), out.detach().numpy()) < 1e-12
    d_out = torch.randn_like(out)
    out.backward(d_out)
    o3, (dxq, dxk, dv) = np_oracle.favor_attention_fwd_bwd(xq.detach(), xk.detach(), v.detach(), P, d_out)
    assert rel_l2(o3.numpy(), out.detach().numpy()) < 1e-12
    assert rel_l2(dxq.numpy(), xq.grad.numpy()) < 1e-10
    assert rel_l2(dxk.numpy(), xk.grad.numpy()) < 1e-10
    assert rel_l2(dv.numpy(), v.grad.numpy()) < 1e-10
