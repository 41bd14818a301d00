"""GPU parity tests (-m gpu) of mi_gdv (validate.py:16-49): golden values produced by executing the
reference's own functions, the CPU oracle at a larger size, and the reference-shaped host call.
Tolerance: strict (hi/lo bf16 operands, fp32 distances, fp64 sums) 2e-5 relative on every term and on the GDV's own
scale |intra| (the GDV is a small difference of the terms); fast (bf16 operands) 2e-3."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "gdv_*.npz")))
TOL = {"strict": 2e-5, "fast": 2e-3}


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as g
    g.build()
    import mi_b200
    from oracle import gdv_oracle
    assert torch.cuda.is_available()
    return mi_b200, gdv_oracle


def _check(got, ref, tol):
    for k in ("intra_pos", "intra_neg", "inter"):
        assert abs(got[k] - float(ref[k])) <= tol * abs(float(ref[k])), k
    scale = max(abs(float(ref["intra_pos"])), abs(float(ref["inter"]))) / np.sqrt(2.0)
    assert abs(got["gdv"] - float(ref["gdv"])) <= tol * scale


@pytest.mark.parametrize("precision", ["strict", "fast"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_values_from_the_reference(env, path, precision):
    mi_b200, _ = env
    z = np.load(path)
    got = mi_b200.gdv_terms(list(z["pos"]), list(z["neg"]), precision=precision)     # lists of rows, as validate.py builds them
    _check(got, z, TOL[precision])


def test_oracle_parity_larger(env):
    mi_b200, gdv_oracle = env
    r = np.random.RandomState(3)
    pos = np.maximum(r.randn(1500, 768), 0).astype(np.float32) + 0.1
    neg = (np.maximum(r.randn(2300, 768), 0) + 0.05 * r.randn(1, 768)).astype(np.float32)
    ref = gdv_oracle.gdv_calculation(pos, neg)
    got = mi_b200.gdv_terms(torch.from_numpy(pos), torch.from_numpy(neg))
    _check(got, ref, TOL["strict"])
    assert isinstance(mi_b200.gdv_calculation(pos, neg), float)
    with pytest.raises(mi_b200.MIError):
        mi_b200.gdv_terms(pos[:1], neg)
