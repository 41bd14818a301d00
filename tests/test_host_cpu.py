"""CPU-side checks (-m "not gpu"): the C-ABI library builds, loads and exports every symbol the
header declares; it fails loudly without a GPU; the host adapter mirrors the reference API; the
sharded path's collective logic is right under gloo with world_size 2."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="session", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from mi_b200 import _lib
    names = _header_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mi_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert _lib.load().mi_abi_version() == 5


def test_ctypes_prototypes_match_the_header_signatures():
    """Every ctypes prototype has exactly as many arguments as the declaration in include/mi_b200.h, pointer arguments are
    bound as pointers and 64-bit sizes as 64-bit integers (a mismatch would only show up as a crash on the GPU box)."""
    import ctypes
    from mi_b200 import _lib
    src = open(os.path.join(ROOT, "include", "mi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = dict(re.findall(r"\b(mi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(decls) == set(_lib.PROTOTYPES)
    for name, params in decls.items():
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [q.strip() for q in params.split(",")]
        _, argtypes = _lib.PROTOTYPES[name]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for q, t in zip(plist, argtypes):
            if "*" in q or q.startswith("mi_stream_t"):
                assert t is ctypes.c_void_p, (name, q)
            elif q.startswith(("int64_t", "size_t")):
                assert ctypes.sizeof(t) == 8 and t is not ctypes.c_void_p, (name, q)
            elif q.startswith("float"):
                assert t is ctypes.c_float, (name, q)
            elif q.startswith("int"):
                assert t is ctypes.c_int, (name, q)


def test_status_strings_and_planning_without_gpu():
    from mi_b200 import _lib
    lib = _lib.load()
    assert lib.mi_status_string(0) == b"ok"
    assert b"no CPU fallback" in lib.mi_status_string(-4)
    # workspace planning is host-only arithmetic
    small = lib.mi_critic_workspace_bytes(256, 64, 1, 0, 0, 1)
    big = lib.mi_critic_workspace_bytes(65536, 1024, 1, 3, 1, 1)
    assert 0 < small < big < (8 << 30)
    assert lib.mi_score_stats_workspace_bytes(1024, 4096, 768) > 0
    assert lib.mi_score_grad_workspace_bytes(1024, 4096, 768, 1) > lib.mi_score_grad_workspace_bytes(1024, 4096, 768, 0)
    assert lib.mi_critic_workspace_bytes(256, 60, 1, 0, 0, 1) == 0      # D % 8 != 0 is rejected


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import mi_b200
    from mi_b200 import _lib, ops
    assert _lib.load().mi_device_check() == -4
    x = torch.zeros(16, 8)
    with pytest.raises(ops.MIError):
        ops.gemm(x.bfloat16(), x.bfloat16())
    critic = mi_b200.FusedCritic(8, "dot")
    pairs = mi_b200.create_mi_pairs(x, x, [str(i) for i in range(16)])
    with pytest.raises(ops.MIError):
        mi_b200.dv_bound_loss(critic(pairs), 16, torch.device("cpu"))


def test_adapter_mirrors_reference_api(golden_dir):
    import mi_b200
    z = np.load(os.path.join(golden_dir, "known_answers.npz"))
    l2 = torch.from_numpy(z["rand_logits"])
    dv = mi_b200.dv_bound_loss(l2, 16, torch.device("cpu"))
    nce = mi_b200.infonce_bound_loss(l2, 16, torch.device("cpu"))
    np.testing.assert_array_equal(dv.numpy(), z["rand_dv"])
    np.testing.assert_array_equal(nce.numpy(), z["rand_infonce"])
    assert dv.shape == (1,) and nce.shape == ()
    assert mi_b200.select_estimator("dv") is mi_b200.dv_bound_loss
    assert mi_b200.select_estimator("infonce") is mi_b200.infonce_bound_loss
    with pytest.raises(ValueError):
        mi_b200.select_estimator("mine")
    c = mi_b200.FusedCritic(32, "bilinear")
    assert [tuple(p.shape) for p in c.parameters()] == [(32, 32)]
    assert list(mi_b200.FusedCritic(32, "dot").parameters()) == []
    X, Y = torch.randn(6, 32), torch.randn(6, 32)
    sid = ["5", "7", "5", "9", "7", "11"]
    pairs = mi_b200.create_mi_pairs(X, Y, sid, torch.device("cpu"))
    assert isinstance(pairs, mi_b200.PairBatch) and pairs.batch_size == 6
    assert pairs.sid.tolist() == [0, 1, 0, 2, 1, 3]
    handle = c(pairs)
    assert isinstance(handle, mi_b200.ScoreHandle)
    with pytest.raises(ValueError):
        mi_b200.dv_bound_loss(handle, 5, None)          # pos_size must be the batch size
    # compat path: explicit pair rows -> [N,1] logits of the same separable critic
    from oracle import matrix_oracle as mo
    rows = mo.create_mi_pairs(X, Y, sid)
    logits = c(rows)
    assert logits.shape == (rows.shape[0], 1)
    torch.testing.assert_close(logits, mo.pair_logits_separable(rows, c.W.detach(), c.inv_tau))


def test_compat_pair_tensor_matches_reference_order(golden_dir):
    """create_mi_pairs_tensor == the reference's pair tensor (row count, order, values) — checked against
    the golden vectors the reference produced, and differentiable like the original."""
    import glob
    import mi_b200
    from oracle import matrix_oracle as mo
    for path in sorted(p for p in glob.glob(os.path.join(golden_dir, "*dups.npz")) if not os.path.basename(p).startswith(("mlp_", "gdv_"))):
        z = np.load(path)
        X, Y = torch.from_numpy(z["X"]).requires_grad_(True), torch.from_numpy(z["Y"])
        sid = [str(int(s)) for s in z["sid"]]
        rows = mi_b200.create_mi_pairs_tensor(X, Y, sid, torch.device("cpu"))
        assert rows.shape[0] == int(z["n_rows"])
        assert torch.equal(rows.detach(), mo.create_mi_pairs(X.detach(), Y, sid))
        W = torch.from_numpy(z["W"]) if "W" in z.files else None
        logits = mo.pair_logits_separable(rows, W, float(z["inv_tau"]))
        np.testing.assert_allclose(logits.detach().numpy(), z["logits"], rtol=1e-12, atol=1e-13)
        fn = mi_b200.dv_bound_loss if str(z["estimator"]) == "dv" else mi_b200.infonce_bound_loss
        fn(logits, X.shape[0], None).sum().backward()
        assert float((X.grad - torch.from_numpy(z["dX"])).abs().max()) < 1e-12
    one = mi_b200.create_mi_pairs_tensor(torch.ones(1, 4), torch.ones(1, 4), ["a"])
    assert one.shape == (1, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dist_worker(rank, world, port, est, critic, ret, int32_ids=False, ref_stride=1, outlier=False, sym_rc=True):
    import torch.distributed as dist
    os.environ["MI_SYM_RC"] = "1" if sym_rc else "0"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_backend
    from mi_b200 import dist as mdist
    from oracle import matrix_oracle as mo
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        B, D = 24, 16
        X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=5, dup_frac=0.2, bilinear=(critic == "bilinear"))
        Xb, Yb = X.bfloat16().float(), Y.bfloat16().float()
        Wb = None if W is None else W.bfloat16().float()
        sid_raw = sid * 1000003 + 17                         # arbitrary int64 ids
        if int32_ids:                                        # exact int32 ids: ids + norm maxima travel in ONE all-gather
            sid_raw = sid.to(torch.int32)
        cpu_backend.REF_STRIDE = ref_stride                  # sampled softmax references (every ref_stride-th own column)
        if outlier:                                          # one score far above its row's SAMPLED columns (column 5 is odd)
            Yb[3] = -Yb[5]                                   # (its own positive pair far BELOW: the reference cannot lean on it)
            Xb[3] = (400.0 * Yb[5]).bfloat16().float()
        Bl = B // world
        sl = slice(rank * Bl, (rank + 1) * Bl)
        out, dX, dY, dW = mdist.sharded_critic_loss_fwd_bwd(Xb[sl], Yb[sl], Wb, sid_raw[sl], est, "strict", 0.5,
                                                           True, None, backend=cpu_backend)
        if ref_stride > 1:                                   # the guard verdict is the same on every rank
            assert float(out["guard"]) == (1.0 if outlier else 0.0), float(out["guard"])
        ref = mo.critic_loss(Xb, Yb, sid, Wb, 0.5, est)
        errs = [abs(float(out["loss"]) - float(ref["loss"])),
                float((dX.double() - ref["dX"][sl]).abs().max() / ref["dX"].abs().max()),
                float((dY.double() - ref["dY"][sl]).abs().max() / ref["dY"].abs().max())]
        if dW is not None:
            errs.append(float((dW.double() - ref["dW"]).abs().max() / ref["dW"].abs().max()))
        errs.append(abs(float(out["n_neg"]) - float(ref["n_neg"])))
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("est", ["dv", "infonce", "infonce_row", "infonce_sym"])
@pytest.mark.parametrize("critic", ["dot", "bilinear"])
def test_sharded_path_world2_gloo(est, critic):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_dist_worker, args=(2, port, est, critic, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        errs = ret[rank]
        assert errs[0] < 1e-6, (rank, errs)           # fp32 log N_neg and fp32 reference vectors
        assert max(errs[1:-1]) < 1e-6, (rank, errs)
        assert errs[-1] == 0


def test_sharded_path_world4_gloo_symmetric_one_statistics_pass():
    """Four ranks: every column's log-sum-exp is merged from four row-block partials (mi_score_stats_rc per rank)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dist_worker, args=(4, _free_port(), "infonce_sym", "bilinear", ret), nprocs=4, join=True)
    assert len(ret) == 4
    for rank in range(4):
        errs = ret[rank]
        assert errs[0] < 1e-6 and max(errs[1:-1]) < 1e-6 and errs[-1] == 0, (rank, errs)


def test_sharded_path_world2_gloo_symmetric_two_statistics_passes():
    """MI_SYM_RC=0: the older symmetric form (one statistics pass per direction, all-gather of T) stays available for A/B."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dist_worker, args=(2, _free_port(), "infonce_sym", "bilinear", ret, False, 1, False, False), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        errs = ret[rank]
        assert errs[0] < 1e-6 and max(errs[1:-1]) < 1e-6 and errs[-1] == 0, (rank, errs)


def test_sharded_path_world2_gloo_packed_int32_ids():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dist_worker, args=(2, _free_port(), "dv", "bilinear", ret, True), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        errs = ret[rank]
        assert errs[0] < 1e-6 and max(errs[1:-1]) < 1e-6 and errs[-1] == 0, (rank, errs)


@pytest.mark.parametrize("est", ["dv", "infonce_row"])
@pytest.mark.parametrize("outlier", [False, True])
def test_sharded_path_world2_gloo_sampled_references_and_guard(est, outlier):
    """Sampled references (stride 2 over the rank's own column block): same results as the exact path; a planted score that
    the sample misses by hundreds of nats overflows the row sum on ONE rank, the merged guard count reaches BOTH ranks and the step
    is repeated on the exact path — never a silently wrong result (dist.py ignored the flag)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dist_worker, args=(2, _free_port(), est, "dot", ret, False, 2, outlier), nprocs=2, join=True)
    assert len(ret) == 2
    for rank in range(2):
        errs = ret[rank]
        # (outlier: scores of several hundred, so the fp32 reference vectors of the exact path carry ~1e3 * 2^-24 relative)
        tol = 3e-4 if outlier else 1e-6
        assert errs[0] < 1000 * tol and max(errs[1:-1]) < tol and errs[-1] == 0, (rank, errs)


def test_merge_scalars_matches_single_block():
    from mi_b200 import dist as mdist
    g = torch.Generator().manual_seed(0)
    lse = torch.randn(40, generator=g, dtype=torch.float64) * 5
    parts = []
    for r in range(4):
        seg = lse[r * 10:(r + 1) * 10]
        m = seg.max()
        parts.append(torch.tensor([m, torch.exp(seg - m).sum(), 10.0, 1.0, 2.0, 0, 0, 0], dtype=torch.float64))
    out = mdist.merge_scalars(torch.stack(parts))
    assert abs(float(out["lse_neg"]) - float(torch.logsumexp(lse, 0))) < 1e-12
    assert float(out["n_neg"]) == 40 and float(out["diag_sum"]) == 4


def test_package_synthetic_generator_equals_the_oracles():
    """bench.py / examples use mi_b200.synthetic (the product never imports oracle/); the tests use the oracle's copy."""
    from mi_b200 import synthetic
    from oracle import matrix_oracle as mo
    for bil in (False, True):
        a = synthetic.synthetic_embeddings(96, 40, seed=5, dup_frac=0.1, bilinear=bil)
        b = mo.synthetic_embeddings(96, 40, seed=5, dup_frac=0.1, bilinear=bil)
        for x, y in zip(a, b):
            assert (x is None and y is None) or torch.equal(x, y)


@pytest.mark.parametrize("rr,B,H1", [(256, 4096, 1024), (1024, 1024, 1024), (48, 48, 1024), (11, 100, 192), (37, 300, 200),
                                     (1, 16, 40), (4096, 256, 320)])
def test_mlp_walk_schedule_covers_every_tile_exactly_once(rr, B, H1):
    """Sched::order 2 (fused dZ1 kernel of the MLP critic): the units' work items are an exact cover of
    {M blocks} x {N tiles}; inside a unit the N tile is fixed and the M blocks are `grp` apart (same text rows, next image
    row) — the property the epilogue's register-resident dC sums rely on; full-length walks are at least 32 tiles long unless the
    panel is shorter.  Planned on the host (mi_plan_mlp_walk), no device needed."""
    import ctypes
    from mi_b200 import _lib
    lib = _lib.load()
    rows_per_mblk = 256 if lib.mi_get_cta_group() == 2 else 128
    grp = -(-B // rows_per_mblk)
    n_nt = -(-H1 // 256)
    total = rr * grp * n_nt
    buf = (ctypes.c_int32 * (3 * total))()
    n = lib.mi_plan_mlp_walk(rr, B, H1, buf, total)
    assert n == total
    items = np.frombuffer(buf, dtype=np.int32).reshape(total, 3)
    seen = set()
    by_unit = {}
    for u, m, nt in items.tolist():
        assert 0 <= m < rr * grp and 0 <= nt < n_nt
        assert (m, nt) not in seen
        seen.add((m, nt))
        by_unit.setdefault(u, []).append((m, nt))
    assert len(seen) == total
    lens = []
    for u, its in by_unit.items():
        assert len({nt for _, nt in its}) == 1                       # one N tile per unit
        ms = [m for m, _ in its]
        assert all(b - a == grp for a, b in zip(ms, ms[1:]))        # consecutive image rows of the same text block
        lens.append(len(its))
    if rr >= 32:                                                     # the flush of the running sums is amortised over >= 32 tiles
        assert max(lens) >= 32


def test_split_k_plan_fills_whole_rounds():
    """choose_ksplit: never more splits than k_blocks / 4, at least min_ks, and never worse (in rounds x K blocks per unit)
    than the old ceil(pairs / tiles) rule that left a nearly empty second round (80 units on 74 CTA pairs)."""
    from mi_b200 import _lib
    lib = _lib.load()
    U = 74 if lib.mi_get_cta_group() == 2 else 148
    ceil = lambda a, b: -(-a // b)
    for tiles, kb in [(8, 8192), (16, 512), (12, 512), (1, 64), (3, 7), (40, 4096), (148, 33)]:
        ks = lib.mi_plan_ksplit(tiles, kb, 1)
        assert 1 <= ks <= max(1, kb // 4)
        cost = ceil(tiles * ks, U) * ceil(kb, ks)
        old = min(max(1, ceil(U, tiles)), max(1, kb // 4))
        assert cost <= ceil(tiles * old, U) * ceil(kb, old)
    ks = lib.mi_plan_ksplit(8, 8192, 1)                            # dW2 of the MLP critic (8 tiles, 2^20 pairs): whole rounds,
    assert ceil(8 * ks, U) * ceil(8192, ks) <= 1.03 * 8 * 8192 / U  # within 3 % of perfectly divisible work
    assert lib.mi_plan_ksplit(16, 1024, 64) >= 64                   # strict mode: chains of <= 16 K blocks
