"""Host logic: the drop-in constructors reproduce the reference's state_dict layout and seeded
initialisation bit-for-bit (golden fingerprints come from the live reference,
tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import CASES, build_product_model, fingerprint


@pytest.mark.parametrize("case", sorted(CASES))
def test_state_dict_layout_and_seeded_init(case, golden):
    model, _ = build_product_model(case)
    sd = model.state_dict()
    assert list(sd.keys()) == list(golden[f"{case}/keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(golden[f"{case}/shapes"])
    fp = np.stack([fingerprint(v) for v in sd.values()])
    np.testing.assert_array_equal(fp, golden[f"{case}/init_fp"])


def test_strict_state_dict_roundtrip():
    a, _ = build_product_model("anp_distractor")
    b, _ = build_product_model("anp_distractor")
    for p in b.parameters():
        p.data.zero_()
    missing, unexpected = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing and not unexpected
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert (x == y).all(), k


def test_dead_fc_and_param_counts():
    # parameter counts of the reference classes
    expect = {"anp_distractor": 4221954, "cnp_distractor_max": 2118402, "anp_3d": 4225748,
              "cnp_3d_max": 2122196, "cnp_1d_max": 362142, "anp_1d": 488874}
    for case, n in expect.items():
        m, _ = build_product_model(case)
        assert sum(p.numel() for p in m.parameters()) == n, case
