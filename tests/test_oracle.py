"""Pins oracle.matrix_oracle: (1) against the committed golden vectors, which
were produced by executing the reference (oracle/make_golden.py); (2) against
the live reference when /root/reference is present; (3) closed-form gradients
against autograd; (4) the known answers of SURVEY.md 8c."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import matrix_oracle as mo
from oracle import ref_loader

CASES = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
               if "known_answers" not in p and not os.path.basename(p).startswith(("mlp_", "gdv_")))


def _load(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    X = torch.from_numpy(d["X"])
    Y = torch.from_numpy(d["Y"])
    W = torch.from_numpy(d["W"]) if "W" in d else None
    sid = [str(int(s)) for s in d["sid"]]
    return d, X, Y, W, sid


def test_fixtures_exist():
    assert len(CASES) >= 7


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_pair_form_matches_reference_vectors(path):
    d, X, Y, W, sid = _load(path)
    rows = mo.create_mi_pairs(X, Y, sid)
    assert rows.shape[0] == int(d["n_rows"])
    # ordering: row k pairs image row_i[k]
    ij = mo.negative_pair_index(sid)
    B = X.shape[0]
    assert np.array_equal(ij[:, 0].numpy().astype(np.int32), d["row_i"][B:])
    logits = mo.pair_logits_separable(rows, W, float(d["inv_tau"]))
    np.testing.assert_allclose(logits.numpy(), d["logits"], rtol=1e-5 if X.dtype == torch.float32 else 1e-12,
                               atol=1e-6 if X.dtype == torch.float32 else 1e-13)
    fn = mo.dv_bound_loss if str(d["estimator"]) == "dv" else mo.infonce_bound_loss
    loss = fn(logits, B)
    assert tuple(loss.shape) == tuple(d["loss_shape"])        # [1] for dv, [] for infonce
    np.testing.assert_allclose(loss.numpy(), d["loss"], rtol=2e-6 if X.dtype == torch.float32 else 1e-12)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_matrix_form_matches_reference_vectors(path):
    d, X, Y, W, sid = _load(path)
    est = str(d["estimator"])
    out = mo.critic_loss(X, Y, sid, W, float(d["inv_tau"]), est, dtype=torch.float64)
    f32 = X.dtype == torch.float32
    # fp64 matrix form vs reference (fp32 cases: the reference itself is fp32)
    np.testing.assert_allclose(out["loss"].numpy(), d["loss"].reshape(-1)[0], rtol=5e-6 if f32 else 1e-6,
                               atol=1e-6 if f32 else 3e-7)   # fp32 log(N_neg) in the reference (mi_critics.py:10)
    for k in ("dX", "dY", "dW"):
        if k in d:
            ref = d[k].astype(np.float64)
            err = np.abs(out[k].numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
            assert err < (2e-5 if f32 else 1e-12), (k, err)
    # N = B + N_neg and duplicates are excluded (main_utils.py:105)
    assert int(out["n_neg"]) == int(d["n_rows"]) - X.shape[0]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("est", ["dv", "infonce"])
@pytest.mark.parametrize("critic", ["dot", "bilinear"])
def test_live_reference_equivalence(est, critic):
    ref = ref_loader.load()
    assert ref.create_mi_pairs is not None, getattr(ref, "import_error", "")
    g = torch.Generator().manual_seed(5)
    B, D = 12, 20
    X = torch.relu(torch.randn(B, D, generator=g, dtype=torch.float64))
    Y = torch.tanh(torch.randn(B, D, generator=g, dtype=torch.float64))
    sid = [str(100 + i) for i in range(B)]
    sid[5] = sid[4]
    sid[11] = sid[0]
    W = torch.randn(D, D, generator=g, dtype=torch.float64) / D if critic == "bilinear" else None
    rows_ref = ref.create_mi_pairs(X, Y, sid, torch.device("cpu"))
    rows = mo.create_mi_pairs(X, Y, sid)
    assert torch.equal(rows_ref, rows)
    logits = mo.pair_logits_separable(rows, W, 0.5)
    fn_ref = ref.dv_bound_loss if est == "dv" else ref.infonce_bound_loss
    fn = mo.dv_bound_loss if est == "dv" else mo.infonce_bound_loss
    a = fn_ref(logits, B, torch.device("cpu"))
    b = fn(logits, B)
    assert a.shape == b.shape
    assert torch.equal(a, b)
    pf = mo.critic_loss_pair_form(X, Y, sid, W, 0.5, est)
    mf = mo.critic_loss(X, Y, sid, W, 0.5, est)
    assert abs(float(pf["loss"].reshape(-1)[0]) - float(mf["loss"])) < 3e-7
    for k in ("dX", "dY", "dW"):
        if k in pf:
            assert (pf[k] - mf[k]).abs().max() < 1e-13


@pytest.mark.parametrize("est", mo.ESTIMATORS)
@pytest.mark.parametrize("critic", ["dot", "bilinear"])
def test_closed_form_gradients_match_autograd(est, critic):
    g = torch.Generator().manual_seed(3)
    B, D = 14, 10
    X = torch.randn(B, D, generator=g, dtype=torch.float64).requires_grad_(True)
    Y = torch.randn(B, D, generator=g, dtype=torch.float64).requires_grad_(True)
    W = (torch.randn(D, D, generator=g, dtype=torch.float64) / 3).requires_grad_(True) if critic == "bilinear" else None
    sid = torch.arange(B)
    sid[3] = sid[2]
    sid[9] = sid[1]
    M = mo.negatives_mask(sid, sid)
    S = mo.score_matrix(X, Y, W, 0.7)
    loss = mo.estimator_from_scores(S, M, est)["loss"]
    params = [X, Y] + ([W] if W is not None else [])
    grads = torch.autograd.grad(loss, params)
    out = mo.critic_loss(X, Y, sid, W, 0.7, est)
    assert abs(float(out["loss"]) - float(loss)) < 1e-12
    assert (out["dX"] - grads[0]).abs().max() < 1e-12
    assert (out["dY"] - grads[1]).abs().max() < 1e-12
    if W is not None:
        assert (out["dW"] - grads[2]).abs().max() < 1e-12


def test_known_answers(golden_dir):
    z = np.load(os.path.join(golden_dir, "known_answers.npz"))
    c = torch.full((32 + 992, 1), 0.37)
    np.testing.assert_array_equal(mo.dv_bound_loss(c, 32).numpy(), z["const_dv"])
    np.testing.assert_array_equal(mo.infonce_bound_loss(c, 32).numpy(), z["const_infonce"])
    assert abs(float(z["const_dv"][0])) < 1e-6
    assert abs(float(z["const_infonce"]) - math.log(992)) < 1e-5
    l2 = torch.from_numpy(z["rand_logits"])
    dv = mo.dv_bound_loss(l2, 16)
    nce = mo.infonce_bound_loss(l2, 16)
    np.testing.assert_array_equal(dv.numpy(), z["rand_dv"])
    np.testing.assert_array_equal(nce.numpy(), z["rand_infonce"])
    assert dv.shape == (1,) and nce.shape == ()
    assert mo.dv_bound_loss(l2[:, 0], 16).shape == () and mo.infonce_bound_loss(l2[:, 0], 16).shape == ()
    # infonce_ref - dv == log(float32(N_neg)) (mi_critics.py:10,21)
    assert abs(float(nce) - float(dv[0]) - float(torch.log(torch.tensor(240).float()))) < 1e-6
    lg = l2.clone().requires_grad_(True)
    mo.dv_bound_loss(lg, 16).sum().backward()
    np.testing.assert_allclose(lg.grad.numpy(), z["rand_dv_dlogits"], rtol=1e-6, atol=1e-9)
    assert np.allclose(lg.grad[:16].numpy(), -1.0 / 16)
    assert abs(float(lg.grad[16:].sum()) - 1.0) < 1e-6
    # no negatives: reference yields nan / -inf
    assert np.isnan(z["noneg_dv"]).all() and np.isneginf(z["noneg_infonce"]).all()


def test_infonce_row_sym_properties():
    X, Y, sid, W = mo.synthetic_embeddings(48, 32, seed=9, dup_frac=0.1)
    row = mo.critic_loss(X, Y, sid, W, 1.0, "infonce_row")
    sym = mo.critic_loss(X, Y, sid, W, 1.0, "infonce_sym")
    # transposition symmetry: sym(X,Y,W) == sym(Y,X,W^T)
    sym_t = mo.critic_loss(Y, X, sid, W.t(), 1.0, "infonce_sym")
    assert abs(float(sym["loss"]) - float(sym_t["loss"])) < 1e-12
    assert (sym["dX"] - sym_t["dY"]).abs().max() < 1e-12
    # rows of the InfoNCE score gradient sum to zero
    M = mo.negatives_mask(sid, sid)
    S = mo.score_matrix(X.double(), Y.double(), W.double())
    G = mo.score_gradient(S, M, "infonce_row")
    assert G.sum(1).abs().max() < 1e-12
    assert float(row["loss"]) >= 0.0


@pytest.mark.parametrize("est", ["dv", "infonce", "infonce_row", "infonce_sym"])
@pytest.mark.parametrize("critic", ["dot", "bilinear"])
def test_chunked_oracle_equals_matrix_oracle(est, critic):
    """oracle.chunked_oracle (the row-chunked fp32 restatement the full-size GPU parity tests use) == the matrix oracle
    (itself pinned to the reference's golden vectors above), incl. duplicate study ids and a ragged last chunk."""
    from oracle import chunked_oracle as co
    B, D = 203, 40
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=21, dup_frac=0.15, bilinear=(critic == "bilinear"))
    inv_tau = 1.0 / math.sqrt(D) if critic == "dot" else 1.0
    ref = mo.critic_loss(X, Y, sid, W, inv_tau, est)
    got = co.critic_loss_chunked(X, Y, sid, W, inv_tau, est, chunk=64)
    assert abs(float(got["loss"]) - float(ref["loss"])) < 2e-6 * max(1.0, abs(float(ref["loss"])))
    assert float(got["n_neg"]) == float(ref["n_neg"])
    for k in ("dX", "dY") + (("dW",) if critic == "bilinear" else ()):
        err = float((got[k].double() - ref[k]).abs().max() / ref[k].abs().max())
        assert err < 2e-5, (k, err)                       # fp32 matmuls against the fp64 matrix form


def test_catloop_pair_construction_equals_vectorised():
    """The reference's growing-torch.cat pair construction (timed by bench.py for BASELINE config 1) == the oracle's gather."""
    X, Y, sid, _ = mo.synthetic_embeddings(12, 8, seed=3, dup_frac=0.3, bilinear=False)
    ids = [str(int(s)) for s in sid]
    assert torch.equal(mo.create_mi_pairs_catloop(X, Y, ids), mo.create_mi_pairs(X, Y, ids))
    a = mo.critic_loss_pair_form(X, Y, ids, None, 0.3, "dv", catloop=True)
    b = mo.critic_loss_pair_form(X, Y, ids, None, 0.3, "dv")
    assert torch.equal(a["loss"], b["loss"]) and torch.allclose(a["dX"], b["dX"], rtol=0, atol=1e-15)
