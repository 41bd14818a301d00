"""TEST INFRASTRUCTURE: a CPU emulation of the stage ops of mi_b200.ops with identical semantics
(include/mi_b200.h), used ONLY to exercise the host-side sharding / collective logic of
mi_b200.dist under gloo without a GPU.  Never imported by the product."""
import torch


def as_bf16(x):
    return x.detach().double().contiguous()      # exact arithmetic: isolates the collective logic


def gemm(A, B, alpha=1.0, gamma=0.0, sub=None, out_dtype=torch.float32, out_split=False, a_t=False, b_t=False):
    A, B = A.double(), B.double()
    C = (A.t() if a_t else A) @ (B if b_t else B.t())
    if sub is not None:
        C = C - gamma * sub.double()
    return alpha * C                           # no rounding, whatever out_dtype / out_split ask for


def transpose(x):
    return x.t().contiguous()


def _scores(Q, K, sid_q, sid_k, q_offset, scale):
    S = (Q.double() @ K.double().t()) * scale
    M = sid_q[:, None] != sid_k[None, :]
    idx = torch.arange(Q.shape[0])
    return S, M, idx, idx + q_offset


def score_stats(Q, K, sid_q, sid_k, q_offset=0, scale=1.0):
    S, M, i, j = _scores(Q, K, sid_q, sid_k, q_offset, scale)
    diag = S[i, j]
    lse_neg = torch.logsumexp(torch.where(M, S, torch.full_like(S, float("-inf"))), 1)
    n_neg = M.sum(1).double()
    lse_all = torch.logaddexp(lse_neg, diag)
    rows = torch.stack([lse_neg, n_neg, diag, lse_all], 1)
    m = lse_neg.max()
    scal = torch.zeros(8, dtype=torch.float64)
    scal[0] = m
    scal[1] = torch.exp(lse_neg[torch.isfinite(lse_neg)] - m).sum()
    scal[2] = n_neg.sum()
    scal[3] = diag.sum()
    scal[4] = (lse_all - diag).sum()
    scal[5] = (n_neg == 0).sum()
    return rows, scal


def score_stats_rc(Q, K, sid_q, sid_k, q_offset=0, scale=1.0):
    """score_stats plus the block's column statistics: col_lse[c] = log-sum-exp over the block's negatives of column c."""
    rows, scal = score_stats(Q, K, sid_q, sid_k, q_offset, scale)
    S, M, _, _ = _scores(Q, K, sid_q, sid_k, q_offset, scale)
    col_lse = torch.logsumexp(torch.where(M, S, torch.full_like(S, float("-inf"))), 0)
    return rows, scal, col_lse.to(torch.float32)


def score_grad(Q, K, sid_q, sid_k, q_offset, scale, refq, wq, refk, wk, include_diag, precision,
               alpha, gamma, want_f32=True, want_bf16=False, out_split=False, want_k=False):
    S, M, i, j = _scores(Q, K, sid_q, sid_k, q_offset, scale)
    incl = M.clone()
    if include_diag:
        incl[i, j] = True
    G = torch.zeros_like(S)
    if refq is not None and wq > 0:
        G = G + wq * torch.exp(S - refq.double()[:, None])
    if refk is not None and wk > 0:
        G = G + wk * torch.exp(S - refk.double()[None, :])
    G = torch.where(incl, G, torch.zeros_like(G))
    Oq = alpha * (G @ K.double() - gamma * K.double()[j])
    Ok = None
    if want_k:
        Ok = G.t() @ Q.double()
        Ok[j] -= gamma * Q.double()
        Ok = alpha * Ok
    return (Oq if want_f32 else None), (Oq if want_bf16 else None), Ok


def row_norm_max(A):
    n = A.double().norm(dim=1)
    return n, n.max().reshape(1)


REF_STRIDE = 1      # tests set this to exercise the sampled references (the library picks its own stride)


def score_ref_sample(Q, K, sid_q, sid_k, q_offset, scale, include_diag, col0=0, n_cols=None, stride=0, subset=False):
    S, M, i, j = _scores(Q, K, sid_q, sid_k, q_offset, scale)
    n_cols = K.shape[0] - col0 if n_cols is None else n_cols
    stride = REF_STRIDE if stride < 1 else stride
    cols = torch.arange(col0, col0 + n_cols, stride)
    lse = torch.logsumexp(torch.where(M[:, cols], S[:, cols], torch.full_like(S[:, cols], float("-inf"))), 1)
    diag = S[i, j]
    ref = torch.logaddexp(lse, diag) if include_diag else torch.where(torch.isfinite(lse), lse, diag)
    margin = 48.0 if (stride > 1 or subset) else 0.0       # kRefMargin of the library
    ref = ref + margin
    return {"ref": ref, "diag": diag, "lam": (ref.max() - margin).reshape(1), "stride": stride}


def score_single_pass(Q, K, sid_q, sid_k, q_offset, scale, include_diag, precision, inv_bg, ref, lam, diag, want_k=True,
                      event_after_k=None, event_after_scal=None, event_k_ready=None, k_local_valid=False):
    S, M, i, j = _scores(Q, K, sid_q, sid_k, q_offset, scale)
    ref = ref.double()
    incl = M.clone()
    if include_diag:
        incl[i, j] = True
    P = torch.where(incl, torch.exp(S - ref[:, None]), torch.zeros_like(S))
    l = P.sum(1)
    n_neg = M.sum(1).double()
    n_incl = n_neg + (1.0 if include_diag else 0.0)
    bad = (n_incl > 0) & ~((l >= 1e-30) & (l <= 1e30))              # the fp32 window of the kernel's row sums
    if include_diag:
        lse_all = ref + torch.log(l)
        lse_neg = torch.logsumexp(torch.where(M, S, torch.full_like(S, float("-inf"))), 1)
        wrow = inv_bg / l
    else:
        lse_neg = ref + torch.log(l)
        lse_all = torch.logaddexp(lse_neg, diag.double())
        d = ref - lam.double().reshape(())
        wrow = torch.exp(d)
        bad = bad | ~(d <= 80.0)
    rows = torch.stack([lse_neg, n_neg, diag.double(), lse_all], 1)
    fin = torch.isfinite(lse_neg)
    m = lse_neg[fin].max() if fin.any() else torch.tensor(float("-inf"), dtype=torch.float64)
    scal = torch.zeros(8, dtype=torch.float64)
    scal[0] = m
    scal[1] = torch.exp(lse_neg[fin] - m).sum()
    scal[2] = n_neg.sum()
    scal[3] = diag.double().sum()
    scal[4] = (lse_all - diag.double()).sum()
    scal[5] = (n_neg == 0).sum()
    scal[6] = bad.sum()
    return {"rows": rows, "scal": scal, "oq_raw": P @ K.double(), "ok_raw": (P * wrow[:, None]).t() @ Q.double(),
            "ref": ref, "wrow": wrow, "lam": lam, "flag": bad.sum().to(torch.int32).reshape(1)}


def single_finalize_q(oq_raw, ref, wrow, lse, dv_like, alpha, gamma, kdiag, want_f32=True, want_bf16=False, out_split=False):
    c = torch.exp(ref.double() - lse.double()) if dv_like else wrow
    O = alpha * (c[:, None] * oq_raw - gamma * kdiag.double())
    return (O if want_f32 else None), (O if want_bf16 else None)


def single_finalize_k(ok, lam, lse, dv_like, alpha, gamma, qdiag):
    kappa = torch.exp(lam.double() - lse.double()) if dv_like else 1.0
    ok.copy_(alpha * (kappa * ok - gamma * qdiag.double()))
    return ok
