"""CPU tests (not gpu): the concat-MLP critic oracle (oracle/mlp_oracle.py) against the golden
vectors produced by executing the reference's make_mlp + create_mi_pairs + estimators, and — when
/root/reference is present — against the live reference."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import matrix_oracle as mo
from oracle import mlp_oracle, ref_loader

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mlp_*.npz")))


def load_case(path):
    z = np.load(path)
    B, D, H1, H2 = (int(v) for v in z["dims"])
    X, Y = torch.from_numpy(z["X"]), torch.from_numpy(z["Y"])
    sid = [str(int(s)) for s in z["sid"]]
    if "W1" in z.files:
        p = {k: torch.from_numpy(z[k]) for k in mlp_oracle.PARAM_NAMES}
    else:                       # the shipped architecture: parameters are rebuilt from the seed
        p = mlp_oracle.init_params(D, H1, H2, int(z["seed"]), X.dtype)
        p["W2"] = p["W2"] * float(z["scales"][0])
        p["W3"] = p["W3"] * float(z["scales"][1])
    return z, X, Y, sid, p


def check_grads(z, out, seed, rtol):
    def rel(a, b):
        b = torch.as_tensor(b).double()
        return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-300))
    assert rel(out["dX"], z["dX"]) < rtol and rel(out["dY"], z["dY"]) < rtol
    for k in mlp_oracle.PARAM_NAMES:
        g = out["d" + k]
        if k == "b3":           # sum of dL/dlogits = 1 - 1: analytically zero, pure rounding noise
            assert float((g.double() - torch.as_tensor(z["db3"]).double()).abs().max()) < 1e-6
        elif "d" + k in z.files:
            assert rel(g, z["d" + k]) < rtol, k
        else:                   # digests of the large weight gradients
            from oracle.make_golden_mlp import digest_vectors
            r, l = digest_vectors(g.shape, seed + 100)
            assert rel(g.double() @ r, z["d" + k + "_r"]) < rtol, k
            assert rel(l @ g.double(), z["d" + k + "_l"]) < rtol, k


def test_fixtures_exist():
    assert len(CASES) == 4
    assert any("h1024x512_f32_shipped" in c for c in CASES)      # make_mlp(1536, [1024, 512]) as shipped


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_pair_and_matrix_form_match_reference_golden(path):
    z, X, Y, sid, p = load_case(path)
    est = str(z["estimator"])
    f32 = X.dtype == torch.float32
    tol = 2e-4 if f32 else 1e-10
    pf = mlp_oracle.mlp_loss_pair_form(X, Y, sid, p, est, dtype=X.dtype)
    assert pf["logits"].shape[0] == int(z["n_rows"])
    np.testing.assert_allclose(pf["logits"].numpy(), z["logits"], rtol=0, atol=2e-6 if f32 else 1e-12)
    assert tuple(pf["loss"].shape) == tuple(z["loss_shape"])                    # [1] (dv) / [] (infonce)
    assert abs(float(pf["loss"].sum()) - float(z["loss"].reshape(-1)[0])) < (2e-6 if f32 else 1e-12)
    check_grads(z, pf, int(z["seed"]), tol)
    mf = mlp_oracle.mlp_loss_matrix_form(X, Y, sid, p, est, dtype=torch.float64)
    assert abs(float(mf["loss"]) - float(z["loss"].reshape(-1)[0])) < (2e-6 if f32 else 1e-9)
    check_grads(z, mf, int(z["seed"]), tol)
    # the matrix S holds the reference's logits: positives on the diagonal, negatives in gap-major order
    idx = mo.negative_pair_index(sid)
    S = mf["S"]
    got = torch.cat([torch.diagonal(S), S[idx[:, 0], idx[:, 1]]])
    np.testing.assert_allclose(got.numpy(), z["logits"].reshape(-1), rtol=0, atol=2e-6 if f32 else 1e-12)


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_init_params_is_make_mlp_under_the_same_seed():
    ref = ref_loader.load()
    torch.manual_seed(24)
    net = ref.make_mlp(2 * 768, [1024, 512])                      # main_utils.py:77
    p = mlp_oracle.params_from_sequential(net)
    q = mlp_oracle.init_params(768, 1024, 512, 24)
    for k in mlp_oracle.PARAM_NAMES:
        assert torch.equal(p[k], q[k]), k


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_matrix_form_matches_live_reference():
    ref = ref_loader.load()
    B, D, H1, H2 = 10, 12, 48, 24
    g = torch.Generator().manual_seed(5)
    X = torch.relu(torch.randn(B, D, generator=g, dtype=torch.float64))
    Y = torch.tanh(torch.randn(B, D, generator=g, dtype=torch.float64))
    sid = [str(i) for i in range(B)]
    sid[4] = sid[3]
    torch.manual_seed(9)
    net = ref.make_mlp(2 * D, [H1, H2]).double()
    with torch.no_grad():
        net[4].weight.mul_(5.0)
    for est, fn in (("dv", ref.dv_bound_loss), ("infonce", ref.infonce_bound_loss)):
        net.zero_grad()
        Xl, Yl = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
        loss = fn(net(ref.create_mi_pairs(Xl, Yl, sid, torch.device("cpu"))), B, torch.device("cpu"))
        loss.sum().backward()
        mf = mlp_oracle.mlp_loss_matrix_form(X, Y, sid, mlp_oracle.params_from_sequential(net), est)
        assert abs(float(mf["loss"]) - float(loss.sum())) < 1e-6      # the reference's log N_neg is fp32
        assert float((mf["dX"] - Xl.grad).abs().max()) < 1e-12
        assert float((mf["dY"] - Yl.grad).abs().max()) < 1e-12
        assert float((mf["dW2"] - net[2].weight.grad).abs().max()) < 1e-12
        assert float((mf["dW1"] - net[0].weight.grad).abs().max()) < 1e-12
        assert float((mf["db3"] - net[4].bias.grad).abs().max()) < 1e-12
