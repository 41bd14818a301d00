"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI / the reference-shaped
adapter, against (a) the golden vectors produced by the reference itself, (b) the CPU oracle on the
same seeded bf16-rounded inputs, (c) size-independent properties at the BASELINE size.

Tolerances (north star): strict ("fp32-accumulate") mode — loss 1e-4 relative, gradients 1e-3
relative max-norm; fast mode (single bf16 rounding of T and of the dS panel) — loss 2e-3,
gradients 1e-2."""
import ctypes
import glob
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

TOL = {"strict": (1e-4, 1e-3), "fast": (2e-3, 1e-2)}
GOLDEN = sorted(p for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))
                if "known_answers" not in p and not os.path.basename(p).startswith(("mlp_", "gdv_")))


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as g
    g.build()
    import mi_b200
    from mi_b200 import _lib, ops
    from oracle import matrix_oracle as mo
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    assert _lib.load().mi_device_check() == 0, "needs an sm_100 device"
    return mi_b200, ops, mo, torch.device("cuda:0")


def _rel(a, ref):
    return float((a.detach().cpu().double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


def _loss_rel(a, ref):
    # relative, with a floor so that losses that happen to sit near zero are judged on the scale of
    # the terms they are made of (|lse| + |pos| >= O(1))
    return abs(float(a) - float(ref)) / max(abs(float(ref)), 1.0)


def _run_adapter(mi_b200, dev, X, Y, W, study_id, inv_tau, est, precision):
    B, D = X.shape
    x = X.to(dev).float().requires_grad_(True)
    y = Y.to(dev).float().requires_grad_(True)
    critic = mi_b200.FusedCritic(D, "dot" if W is None else "bilinear", temperature=1.0 / inv_tau, precision=precision).to(dev)
    if W is not None:
        with torch.no_grad():
            critic.W.copy_(W.to(dev).float())
    pairs = mi_b200.create_mi_pairs(x, y, study_id, dev)            # main_utils.py:220-221
    out = critic(pairs)                                              # main_utils.py:222
    loss = mi_b200.select_estimator(est)(out, B, dev)                # main_utils.py:141-144, :224
    loss.sum().backward()                                            # main_utils.py:226
    torch.cuda.synchronize()
    return loss, x.grad, y.grad, (None if W is None else critic.W.grad)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors_from_the_reference(env, path):
    mi_b200, ops, mo, dev = env
    z = np.load(path)
    X, Y = torch.from_numpy(z["X"]), torch.from_numpy(z["Y"])
    W = torch.from_numpy(z["W"]) if "W" in z.files else None
    sid = [str(int(s)) for s in z["sid"]]
    est = str(z["estimator"])
    loss, dX, dY, dW = _run_adapter(mi_b200, dev, X, Y, W, sid, float(z["inv_tau"]), est, "strict")
    assert tuple(loss.shape) == tuple(z["loss_shape"])               # [1] for dv, [] for infonce
    lt, gt = TOL["strict"]
    assert _loss_rel(loss.sum().item(), z["loss"].reshape(-1)[0]) < lt
    assert _rel(dX, torch.from_numpy(z["dX"])) < gt
    assert _rel(dY, torch.from_numpy(z["dY"])) < gt
    if W is not None:
        assert _rel(dW, torch.from_numpy(z["dW"])) < gt


SWEEP = [
    # B, D, critic, estimator, dup_frac
    (32, 768, "dot", "dv", 0.0),            # BASELINE config 1 shape
    (33, 8, "dot", "dv", 0.3),              # ragged, tiny D
    (129, 72, "bilinear", "dv", 0.1),
    (257, 136, "bilinear", "infonce", 0.05),
    (300, 64, "dot", "infonce_row", 0.1),
    (512, 256, "bilinear", "infonce_row", 0.05),
    (600, 200, "dot", "infonce_sym", 0.1),
    (1024, 768, "bilinear", "infonce_sym", 0.05),   # BASELINE config 2 family (B reduced for the CPU oracle)
    (2048, 1024, "bilinear", "dv", 0.05),
]


@pytest.mark.parametrize("precision", ["strict", "fast"])
@pytest.mark.parametrize("B,D,critic,est,dup", SWEEP)
def test_oracle_parity(env, B, D, critic, est, dup, precision):
    mi_b200, ops, mo, dev = env
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=B + D, dup_frac=dup, bilinear=(critic == "bilinear"))
    Xb, Yb = X.bfloat16().float(), Y.bfloat16().float()
    Wb = None if W is None else W.bfloat16().float()
    inv_tau = 1.0 / math.sqrt(D) if critic == "dot" else 1.0
    ref = mo.critic_loss(Xb, Yb, sid, Wb, inv_tau, est)
    loss, dX, dY, dW = _run_adapter(mi_b200, dev, Xb, Yb, Wb, [int(s) for s in sid], inv_tau, est, precision)
    lt, gt = TOL[precision]
    assert _loss_rel(loss.sum().item(), ref["loss"]) < lt
    assert _rel(dX, ref["dX"]) < gt
    assert _rel(dY, ref["dY"]) < gt
    if W is not None:
        assert _rel(dW, ref["dW"]) < gt


def test_peaked_softmax_strict(env):
    """Scores multiplied by 8 (a nearly one-hot softmax): the case where rounding dS to a single
    bf16 would break the 1e-3 bound and the hi/lo split must hold it."""
    mi_b200, ops, mo, dev = env
    B, D = 384, 128
    X, Y, sid, _ = mo.synthetic_embeddings(B, D, seed=77, dup_frac=0.05, bilinear=False)
    Xb, Yb = X.bfloat16().float(), Y.bfloat16().float()
    inv_tau = 8.0 / math.sqrt(D)
    for est in ("dv", "infonce_sym"):
        ref = mo.critic_loss(Xb, Yb, sid, None, inv_tau, est)
        loss, dX, dY, _ = _run_adapter(mi_b200, dev, Xb, Yb, None, [int(s) for s in sid], inv_tau, est, "strict")
        assert _loss_rel(loss.sum().item(), ref["loss"]) < 1e-4
        assert _rel(dX, ref["dX"]) < 1e-3 and _rel(dY, ref["dY"]) < 1e-3


def test_edge_cases(env):
    mi_b200, ops, mo, dev = env
    D = 16
    x = torch.randn(8, D, device=dev)
    critic = mi_b200.FusedCritic(D, "dot").to(dev)
    # every study id equal -> no negatives: reference gives nan / -inf, the fused path refuses
    pairs = mi_b200.create_mi_pairs(x, x, ["7"] * 8, dev)
    with pytest.raises(ops.MIError):
        mi_b200.dv_bound_loss(critic(pairs), 8, dev)
    # B = 1 likewise
    with pytest.raises(ops.MIError):
        mi_b200.dv_bound_loss(critic(mi_b200.create_mi_pairs(x[:1], x[:1], ["1"], dev)), 1, dev)
    # D not a multiple of 8 is rejected by the ABI (TMA needs 16-byte row pitch)
    with pytest.raises(ops.MIError):
        c2 = mi_b200.FusedCritic(12, "dot").to(dev)
        mi_b200.dv_bound_loss(c2(mi_b200.create_mi_pairs(x[:, :12], x[:, :12], list(range(8)), dev)), 8, dev)
    # B = 2 with distinct ids: two negatives
    X, Y = torch.randn(2, D).bfloat16().float(), torch.randn(2, D).bfloat16().float()
    ref = mo.critic_loss(X, Y, torch.tensor([0, 1]), None, 0.25, "dv")
    loss, dX, dY, _ = _run_adapter(mi_b200, dev, X, Y, None, ["a", "b"], 0.25, "dv", "strict")
    assert _loss_rel(loss.item(), ref["loss"]) < 1e-4 and _rel(dX, ref["dX"]) < 1e-3
    # forward only (no_grad): loss still right, nothing saved
    with torch.no_grad():
        l2 = mi_b200.infonce_bound_loss(mi_b200.FusedCritic(D, "dot", temperature=4.0).to(dev)(
            mi_b200.create_mi_pairs(X.to(dev), Y.to(dev), ["a", "b"], dev)), 2, dev)
    ref2 = mo.critic_loss(X, Y, torch.tensor([0, 1]), None, 0.25, "infonce", grads=False)
    assert l2.shape == () and _loss_rel(l2.item(), ref2["loss"]) < 1e-4


def test_stage_ops_with_offsets(env):
    """The sharded usage: Bq != Bk, q_offset != 0 (rank r's row block against all columns)."""
    mi_b200, ops, mo, dev = env
    Bk, Bq, D, off = 1000, 300, 72, 500
    g = torch.Generator().manual_seed(1)
    K = (torch.randn(Bk, D, generator=g) / D ** 0.25).bfloat16()
    Q = (torch.randn(Bq, D, generator=g) / D ** 0.25).bfloat16()
    sid_k = torch.arange(Bk)
    sid_k[1::9] = sid_k[0::9][: len(sid_k[1::9])]
    sid_q = sid_k[off:off + Bq]
    S = (Q.double() @ K.double().t()) * 0.6
    M = sid_q[:, None] != sid_k[None, :]
    idx = torch.arange(Bq)
    diag = S[idx, idx + off]
    lse_neg = torch.logsumexp(torch.where(M, S, torch.full_like(S, -float("inf"))), 1)
    rows, scal = ops.score_stats(Q.to(dev), K.to(dev), sid_q.to(dev), sid_k.to(dev), off, 0.6)
    torch.cuda.synchronize()
    assert float((rows[:, 0].cpu().double() - lse_neg).abs().max()) < 1e-5
    assert torch.equal(rows[:, 1].cpu().double(), M.sum(1).double())
    assert float((rows[:, 2].cpu().double() - diag).abs().max()) < 1e-5
    assert abs(float(scal[0] + torch.log(scal[1])) - float(torch.logsumexp(lse_neg, 0))) < 1e-6
    assert float(scal[2]) == float(M.sum())


def test_sharded_composition_world1_equals_fused_call(env):
    """The stage-level composition used by the multi-GPU path (world 1 here) against the fused C call, in both
    its forms: single pass (Cauchy-Schwarz reference) and two pass (exact log-sum-exp references)."""
    mi_b200, ops, mo, dev = env
    from mi_b200 import dist as mdist
    B, D = 520, 136
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=4, dup_frac=0.1)
    Xd, Yd, Wd, sd = X.bfloat16().to(dev), Y.bfloat16().to(dev), W.bfloat16().to(dev), sid.to(dev)
    ref = {}
    for est in ("dv", "infonce", "infonce_row", "infonce_sym"):
        ref[est] = mo.critic_loss(X.bfloat16().float(), Y.bfloat16().float(), sid, W.bfloat16().float(), 1.0, est)
        for two_pass in (False, True):
            out, dX, dY, dW = mdist.sharded_critic_loss_fwd_bwd(Xd, Yd, Wd, sd, est, "strict", 1.0, True, two_pass=two_pass)
            lo, fX, fY, fW = ops.critic_loss_fwd_bwd(Xd, Yd, Wd, sd.to(torch.int32), est, "strict", 1.0, True, two_pass=two_pass)
            torch.cuda.synchronize()
            assert float(lo[7]) == 0.0 and abs(float(out["loss"]) - float(lo[0])) < 1e-6 * max(1.0, abs(float(lo[0])))
            for a, b in ((dX, fX), (dY, fY), (dW, fW)):
                assert _rel(a, b.cpu()) < 1e-4
            assert _loss_rel(out["loss"], ref[est]["loss"]) < 1e-4
            assert _rel(dX, ref[est]["dX"]) < 1e-3 and _rel(dY, ref[est]["dY"]) < 1e-3 and _rel(dW, ref[est]["dW"]) < 1e-3


def test_single_pass_reference_guard(env):
    """Rows whose Cauchy-Schwarz bound is far above all of their scores (here: an image embedding that is
    orthogonal to every text embedding but has a huge norm) must be reported (loss_out[7] > 0) and the adapter
    must fall back to the exact two-pass path."""
    mi_b200, ops, mo, dev = env
    B, D = 256, 64
    g = torch.Generator().manual_seed(3)
    Y = torch.zeros(B, D)
    Y[:, : D // 2] = torch.randn(B, D // 2, generator=g)
    X = torch.zeros(B, D)
    X[:, : D // 2] = torch.randn(B, D // 2, generator=g)
    X[7, :] = 0.0
    X[7, D // 2:] = 40.0                                  # |X_7| large, <X_7, Y_j> = 0 for all j: bound ~ 1e3, scores 0
    Xb, Yb = X.bfloat16().float(), Y.bfloat16().float()
    sid = torch.arange(B)
    out, *_ = ops.critic_loss_fwd_bwd(Xb.to(dev), Yb.to(dev), None, sid.to(dev), "dv", "strict", 1.0, True)
    assert float(out[7]) >= 1.0
    ref = mo.critic_loss(Xb, Yb, sid, None, 1.0, "dv")
    loss, dX, dY, _ = _run_adapter(mi_b200, dev, Xb, Yb, None, list(range(B)), 1.0, "dv", "strict")
    assert _loss_rel(loss.sum().item(), ref["loss"]) < 1e-4
    assert _rel(dX, ref["dX"]) < 1e-3 and _rel(dY, ref["dY"]) < 1e-3


def test_sharded_autograd_loss_world1(env):
    """mi_b200.sharded_mi_loss (the data-parallel entry used with DDP encoders) == the single-GPU adapter."""
    mi_b200, ops, mo, dev = env
    B, D = 320, 64
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=12, dup_frac=0.1)
    outs = []
    for sharded in (False, True):
        x = X.bfloat16().float().to(dev).requires_grad_(True)
        y = Y.bfloat16().float().to(dev).requires_grad_(True)
        critic = mi_b200.FusedCritic(D, "bilinear", precision="strict").to(dev)
        with torch.no_grad():
            critic.W.copy_(W.bfloat16().float())
        if sharded:
            loss = mi_b200.sharded_mi_loss(x, y, critic, sid * 7 + 3, "infonce_sym")
        else:
            loss = mi_b200.select_estimator("infonce_sym")(critic(mi_b200.create_mi_pairs(x, y, [int(s) for s in sid], dev)), B, dev)
        loss.backward()
        outs.append((loss.item(), x.grad.clone(), y.grad.clone(), critic.W.grad.clone()))
    assert abs(outs[0][0] - outs[1][0]) < 1e-6
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert _rel(a, b.cpu()) < 1e-5


def test_host_buffer_abi_matches_device_call(env):
    mi_b200, ops, mo, dev = env
    from mi_b200 import _lib
    lib = _lib.load()
    B, D = 384, 64
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=8, dup_frac=0.05)
    Xh, Yh, Wh = X.bfloat16().float().contiguous(), Y.bfloat16().float().contiguous(), W.bfloat16().float().contiguous()
    sh = sid.to(torch.int32).contiguous()
    n = lib.mi_critic_host_scratch_bytes(B, D, 1, 0, 1, 1)
    scratch = torch.empty(n, dtype=torch.uint8, device=dev)
    loss = torch.zeros(8, dtype=torch.float64)
    dX, dY, dW = torch.zeros(B, D), torch.zeros(B, D), torch.zeros(D, D)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lib.mi_critic_loss_fwd_bwd_host(p(Xh), p(Yh), p(Wh), p(sh), B, D, 1, 0, 1, 1.0, p(loss), p(dX), p(dY), p(dW),
                                         p(scratch), n, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0, lib.mi_status_string(st)
    ref = mo.critic_loss(Xh, Yh, sid, Wh, 1.0, "dv")
    assert _loss_rel(loss[0], ref["loss"]) < 1e-4
    assert _rel(dX, ref["dX"]) < 1e-3 and _rel(dY, ref["dY"]) < 1e-3 and _rel(dW, ref["dW"]) < 1e-3


@pytest.mark.parametrize("critic", ["dot", "bilinear"])
def test_host_buffer_abi_streams_panels(env, critic):
    """Several row panels: the host entry point feeds the image embeddings panel by panel from a second stream (cast and
    projection inside the pass, lambda from the first panel) and copies dY back from an in-pass event — the result must
    equal the device-pointer call on the same inputs."""
    mi_b200, ops, mo, dev = env
    from mi_b200 import _lib
    lib = _lib.load()
    B, D = 24576, 64                     # 2 panels of 18944 rows at D = 64
    bil = critic == "bilinear"
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=9, dup_frac=0.05, bilinear=bil)
    Xh, Yh = X.bfloat16().float().contiguous().pin_memory(), Y.bfloat16().float().contiguous().pin_memory()
    Wh = W.bfloat16().float().contiguous().pin_memory() if bil else None
    sh = sid.to(torch.int32).contiguous().pin_memory()
    inv_tau = 1.0 if bil else 1.0 / math.sqrt(D)
    for prec_name, prec in (("fast", 0), ("strict", 1)):
        n = lib.mi_critic_host_scratch_bytes(B, D, int(bil), 0, prec, 1)
        scratch = torch.empty(n, dtype=torch.uint8, device=dev)
        loss = torch.zeros(8, dtype=torch.float64).pin_memory()
        dX, dY = torch.zeros(B, D).pin_memory(), torch.zeros(B, D).pin_memory()
        dW = torch.zeros(D, D).pin_memory() if bil else None
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        st = lib.mi_critic_loss_fwd_bwd_host(p(Xh), p(Yh), p(Wh), p(sh), B, D, int(bil), 0, prec, inv_tau, p(loss), p(dX), p(dY),
                                             p(dW), p(scratch), n, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert st == 0, lib.mi_status_string(st)
        ref, rX, rY, rW = ops.critic_loss_fwd_bwd(Xh.to(dev), Yh.to(dev), None if Wh is None else Wh.to(dev), sh.to(dev), "dv",
                                                  prec_name, inv_tau, True)
        torch.cuda.synchronize()
        assert float(loss[7]) == 0.0 and float(ref[7]) == 0.0
        assert abs(float(loss[0]) - float(ref[0])) < 1e-6 * max(1.0, abs(float(ref[0])))
        tol = 1e-4 if prec_name == "strict" else 1e-2        # fast: a 1-ulp fp32 difference may flip a bf16 rounding of dT
        assert _rel(dX, rX.cpu()) < tol and _rel(dY, rY.cpu()) < tol
        if bil:
            assert _rel(dW, rW.cpu()) < tol


@pytest.mark.parametrize("critic", ["dot", "bilinear"])
def test_full_size_properties(env, critic):
    """B = 65536, D = 1024 (the BASELINE metric's size): the CPU oracle cannot form the 4.3e9 pairs,
    so parity is checked through (1) sampled rows / columns recomputed exactly on the CPU in fp64
    from the kernel's own global statistics, (2) Euler-type identities that tie the independently
    computed passes together: <dT,T> = <dY,Y> = <dW,W> (all equal sum(G * S))."""
    mi_b200, ops, mo, dev = env
    B, D = 65536, 1024
    X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=1234, dup_frac=0.05, bilinear=(critic == "bilinear"))
    Xb, Yb = X.bfloat16(), Y.bfloat16()
    Wb = None if W is None else W.bfloat16()
    inv_tau = 1.0 / math.sqrt(D) if critic == "dot" else 1.0
    out, dX, dY, dW = ops.critic_loss_fwd_bwd(Xb.to(dev), Yb.to(dev), None if Wb is None else Wb.to(dev), sid.to(dev),
                                              "dv", "fast", inv_tau, True)
    torch.cuda.synchronize()
    out = out.cpu()
    lse, n_neg, pos = float(out[2]), float(out[3]), float(out[1])
    # exact N_neg from the study ids alone
    _, counts = torch.unique(sid, return_counts=True)
    assert n_neg == float(B) * B - float((counts.double() ** 2).sum())
    assert abs(float(out[0]) - (lse - math.log(np.float32(n_neg)) - pos)) < 1e-9
    # (1) sampled rows: S_i,: in fp64 on the CPU, row statistics and dT_i from the global LSE
    Xd, Yd = Xb.double(), Yb.double()
    T = Xd if Wb is None else (Xb.float() @ Wb.float()).bfloat16().double()      # the path's own bf16 T
    rows = torch.tensor([0, 1, 4097, 33333, 65535])
    S = (T[rows] @ Yd.t()) * inv_tau
    M = sid[rows][:, None] != sid[None, :]
    G = torch.where(M, torch.exp(S - lse), torch.zeros_like(S))
    dT_ref = inv_tau * (G @ Yd - Yd[rows] / B)
    if Wb is None:
        assert _rel(dX[rows], dT_ref) < 1e-2
    cols = torch.tensor([5, 4096, 65000])
    Sc = (T @ Yd[cols].t()) * inv_tau
    Mc = sid[:, None] != sid[cols][None, :]
    Gc = torch.where(Mc, torch.exp(Sc - lse), torch.zeros_like(Sc))
    dY_ref = inv_tau * (Gc.t() @ T - T[cols] / B)
    assert _rel(dY[cols], dY_ref) < 1e-2
    pos_ref = float(((T * Yd).sum(1) * inv_tau).mean())
    assert abs(pos - pos_ref) < 1e-4 * max(1.0, abs(pos_ref))
    # (2) identities between the passes
    a = float((dY.double().cpu() * Yd).sum())
    if Wb is None:
        b = float((dX.double().cpu() * Xd).sum())
        assert abs(a - b) < 2e-3 * max(abs(a), abs(b), 1e-3)
    else:
        c = float((dW.double().cpu() * Wb.double()).sum())
        d = float((dX.double().cpu() * Xd).sum())
        assert abs(a - c) < 5e-3 * max(abs(a), abs(c), 1e-3)
        assert abs(a - d) < 5e-3 * max(abs(a), abs(d), 1e-3)


def test_cuda_graph_replay_equals_direct_call(env):
    """ops.GraphedCriticStep (launch-bound regime): the captured launch sequence reproduces the direct call bit for bit,
    also after the inputs change."""
    mi_b200, ops, mo, dev = env
    B, D = 1024, 256
    step = ops.GraphedCriticStep(B, D, bilinear=True, estimator="dv", precision="fast", inv_tau=1.0, device=dev)
    for seed in (1, 2):
        X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=seed, dup_frac=0.05, bilinear=True)
        Xd, Yd, Wd, sd = X.to(dev).bfloat16(), Y.to(dev).bfloat16(), W.to(dev).bfloat16(), sid.to(torch.int32).to(dev)
        got = [None if t is None else t.clone() for t in step(Xd, Yd, Wd, sd)]
        ref = ops.critic_loss_fwd_bwd(Xd, Yd, Wd, sd, "dv", "fast", 1.0, True)
        torch.cuda.synchronize()
        assert float(got[0][0]) == float(ref[0][0])
        for a, b in zip(got[1:], ref[1:]):
            assert torch.equal(a, b)
