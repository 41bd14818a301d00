"""CPU tests (not gpu): the GDV oracle (oracle/gdv_oracle.py) against golden values produced by executing
the reference's own validate.py functions, and against the live reference when /root/reference is present."""
import glob
import os

import numpy as np
import pytest

from oracle import gdv_oracle, ref_loader

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "gdv_*.npz")))


def test_fixtures_exist():
    assert len(CASES) == 3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_golden(path):
    z = np.load(path)
    out = gdv_oracle.gdv_calculation(z["pos"], z["neg"])
    for k in ("intra_pos", "intra_neg", "inter", "gdv"):
        assert abs(out[k] - float(z[k])) <= 2e-6 * abs(float(z[k])) + 1e-12, k


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_oracle_matches_live_reference():
    ref = ref_loader.load_gdv()
    r = np.random.RandomState(1)
    pos, neg = r.randn(30, 20), r.randn(45, 20) + 0.5            # float64 inputs: exact agreement expected
    out = gdv_oracle.gdv_calculation(pos, neg)
    assert abs(out["gdv"] - ref.gdv_calculation(list(pos), list(neg))) < 1e-12
    zp = ref.z_scored_transform(source_tensor=pos)
    np.testing.assert_allclose(gdv_oracle.z_scored_transform(pos), zp, rtol=0, atol=1e-12)
    assert abs(out["intra_pos"] - ref.mean_intra_class_distance(zp)) < 1e-15
