"""L1 oracle: CPU restatement of the reference MI critic / estimator path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Plain torch on the CPU,
fp64 by default.  Every function cites the reference lines it restates
(paths relative to ``/root/reference/mutual_info_img_txt/``).

Two forms are provided and proved equal in ``tests/``:

* pair form  — exactly what the reference executes: build the ``[B+N_neg, 2D]``
  pair tensor in the reference's order, apply a per-pair critic, feed the flat
  logits to ``dv_bound_loss`` / ``infonce_bound_loss``;
* matrix form — ``S = (X W) Y^T`` with a negatives mask, which is what the CUDA
  path computes and the only form that is feasible for large B.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

ESTIMATORS = ("dv", "infonce", "infonce_row", "infonce_sym")


# --------------------------------------------------------------------------
# pair form (reference order)
# --------------------------------------------------------------------------
def negative_pair_index(study_id: Sequence) -> torch.Tensor:
    """(i, j) of every negative pair in the reference's gap-major order.

    main_utils.py:99-108: ``for gap in range(B-1): for i in range(B):
    j = (i+gap+1) mod B; if study_id[i] != study_id[j]: append``.
    """
    B = len(study_id)
    out = []
    for gap in range(B - 1):
        for i in range(B):
            j = i + gap + 1
            if j >= B:
                j -= B
            if study_id[i] != study_id[j]:
                out.append((i, j))
    if not out:
        return torch.zeros((0, 2), dtype=torch.long)
    return torch.tensor(out, dtype=torch.long)


def create_mi_pairs(X: torch.Tensor, Y: torch.Tensor, study_id: Sequence) -> torch.Tensor:
    """main_utils.py:80-110 without the O(B^2) ``torch.cat`` loop: rows 0..B-1
    are ``[X_i ; Y_i]`` (main_utils.py:93), then one row ``[X_i ; Y_j]`` per
    negative pair in gap-major order (main_utils.py:99-108)."""
    pos = torch.cat((X, Y), 1)
    ij = negative_pair_index(study_id)
    if ij.shape[0] == 0:
        return pos
    neg = torch.cat((X[ij[:, 0]], Y[ij[:, 1]]), 1)
    return torch.cat((pos, neg), 0)


def create_mi_pairs_catloop(X: torch.Tensor, Y: torch.Tensor, study_id: Sequence) -> torch.Tensor:
    """main_utils.py:80-110 AS THE REFERENCE EXECUTES IT: the pair tensor grows by one ``torch.cat`` of the whole tensor per
    negative pair (main_utils.py:106-108), i.e. O(B^2) Python iterations and O(B^4 D) bytes copied.  Same result as
    ``create_mi_pairs`` above (tests prove it); kept so that ``bench.py`` can time BASELINE config 1 the way the reference
    spends its time (>= 87 % of the reference's step is this loop)."""
    B = len(study_id)
    pairs = torch.cat((X, Y), 1)                                   # :93 the B matched rows
    for gap in range(B - 1):                                       # :99
        for i in range(B):                                         # :100
            j = i + gap + 1
            if j >= B:
                j -= B
            if study_id[i] != study_id[j]:                         # :105
                row = torch.cat((X[i], Y[j])).reshape(1, -1)       # :106-107
                pairs = torch.cat((pairs, row), 0)                 # :108 re-copies everything built so far
    return pairs


def dv_bound_loss(logits: torch.Tensor, pos_size: int) -> torch.Tensor:
    """mi_critics.py:3-12.  ``log N_neg`` is evaluated in float32 on the host
    (mi_critics.py:10: ``torch.log(torch.tensor(n).float())``) — mirrored."""
    size = logits.shape[0]
    pos_energy = torch.mean(logits[:pos_size])
    lse = torch.logsumexp(logits[pos_size:], dim=0)
    log_n = torch.log(torch.tensor(size - pos_size).float())
    return lse - log_n.to(lse.dtype) - pos_energy


def infonce_bound_loss(logits: torch.Tensor, pos_size: int) -> torch.Tensor:
    """mi_critics.py:14-23.  The inner ``torch.mean`` (line 21) acts on a
    1-element tensor for ``[N,1]`` logits and is a no-op; for 1-D logits the
    logsumexp is 0-d and ``mean`` is the identity as well."""
    pos_energy = torch.mean(logits[:pos_size])
    lse = torch.logsumexp(logits[pos_size:], dim=0)
    return torch.mean(lse) - pos_energy


def pair_logits_separable(rows: torch.Tensor, W: Optional[torch.Tensor], inv_tau: float) -> torch.Tensor:
    """The separable critic evaluated on reference pair rows ``[x ; y]``:
    ``f = inv_tau * x^T W y`` (``W=None`` is the dot-product critic).  Occupies
    the ``self.mi_discriminator(mi_input)`` slot (main_utils.py:222) and returns
    ``[N,1]`` logits like ``make_mlp`` does (model.py:18-32)."""
    D = rows.shape[1] // 2
    x, y = rows[:, :D], rows[:, D:]
    t = x if W is None else x @ W
    return (t * y).sum(1, keepdim=True) * inv_tau


# --------------------------------------------------------------------------
# matrix form
# --------------------------------------------------------------------------
def dense_ids(study_id: Sequence) -> torch.Tensor:
    """Exact map study_id -> dense int64 (equal ids <=> equal codes)."""
    table: Dict = {}
    return torch.tensor([table.setdefault(s, len(table)) for s in study_id], dtype=torch.long)


def negatives_mask(sid_rows: torch.Tensor, sid_cols: torch.Tensor, row_offset: int = 0) -> torch.Tensor:
    """``M[i,j] = study_id[i] != study_id[j]`` (main_utils.py:105).  The
    diagonal never appears among the negatives because ``gap+1`` ranges over
    1..B-1 (main_utils.py:99-104); it has equal ids anyway."""
    M = sid_rows[:, None] != sid_cols[None, :]
    idx = torch.arange(sid_rows.shape[0]) + row_offset
    M[torch.arange(sid_rows.shape[0]), idx] = False
    return M


def score_matrix(X, Y, W=None, inv_tau: float = 1.0):
    T = X if W is None else X @ W
    return (T @ Y.t()) * inv_tau


def _masked_lse(S, M, dim=None):
    neg_inf = torch.full_like(S, float("-inf"))
    Sm = torch.where(M, S, neg_inf)
    if dim is None:
        return torch.logsumexp(Sm.reshape(-1), 0)
    return torch.logsumexp(Sm, dim)


def estimator_from_scores(S: torch.Tensor, M: torch.Tensor, estimator: str) -> Dict[str, torch.Tensor]:
    """All four estimators from a square score matrix and negatives mask.

    dv          : mi_critics.py:3-12 on the logits [diag ; S[M]]
    infonce     : mi_critics.py:14-23 (the reference's "infonce" = dv + log N)
    infonce_row : mean_i[ LSE_{j in {i} u M_i} S_ij - S_ii ]   (not in reference)
    infonce_sym : 1/2 (row + column version)                    (not in reference)
    """
    B = S.shape[0]
    diag = torch.diagonal(S)
    pos = diag.mean()
    n_neg = int(M.sum())
    out = {"pos_mean": pos, "n_neg": torch.tensor(float(n_neg), dtype=S.dtype)}
    if estimator in ("dv", "infonce"):
        lse = _masked_lse(S, M)
        out["lse_neg"] = lse
        if estimator == "dv":
            log_n = torch.log(torch.tensor(n_neg).float()).to(S.dtype)  # mi_critics.py:10
            out["loss"] = lse - log_n - pos
        else:
            out["loss"] = lse - pos
        return out
    R = M | torch.eye(B, dtype=torch.bool)
    row_lse = _masked_lse(S, R, dim=1)
    out["row_lse"] = row_lse
    loss_row = (row_lse - diag).mean()
    if estimator == "infonce_row":
        out["loss"] = loss_row
        return out
    if estimator == "infonce_sym":
        col_lse = _masked_lse(S, R, dim=0)
        out["col_lse"] = col_lse
        out["loss"] = 0.5 * (loss_row + (col_lse - diag).mean())
        return out
    raise ValueError(f"unknown estimator {estimator!r}")


def score_gradient(S: torch.Tensor, M: torch.Tensor, estimator: str) -> torch.Tensor:
    """Closed form of dL/dS — what autograd produces through
    mi_critics.py:7-12 (positives get -1/B, negatives softmax) and its row /
    column analogues."""
    B = S.shape[0]
    eye = torch.eye(B, dtype=S.dtype)
    if estimator in ("dv", "infonce"):
        lse = _masked_lse(S, M)
        return torch.where(M, torch.exp(S - lse), torch.zeros_like(S)) - eye / B
    R = M | torch.eye(B, dtype=torch.bool)
    row_lse = _masked_lse(S, R, dim=1)
    G = torch.where(R, torch.exp(S - row_lse[:, None]), torch.zeros_like(S)) / B
    if estimator == "infonce_row":
        return G - eye / B
    col_lse = _masked_lse(S, R, dim=0)
    Gc = torch.where(R, torch.exp(S - col_lse[None, :]), torch.zeros_like(S)) / B
    return 0.5 * (G + Gc) - eye / B


def critic_loss(X, Y, study_id, W=None, inv_tau: float = 1.0, estimator: str = "dv",
                dtype=torch.float64, grads: bool = True) -> Dict[str, torch.Tensor]:
    """Matrix-form loss (and closed-form gradients) on CPU in ``dtype``.

    Gradients: dT = G Y, dY = G^T T, dX = dT W^T, dW = X^T dT (bilinear) or
    dX = inv_tau * G Y (dot); see SURVEY.md 7.2, checked against autograd in
    tests/test_oracle.py.
    """
    X = X.detach().to("cpu", dtype)
    Y = Y.detach().to("cpu", dtype)
    Wd = None if W is None else W.detach().to("cpu", dtype)
    sid = study_id if torch.is_tensor(study_id) else dense_ids(study_id)
    sid = sid.to("cpu")
    M = negatives_mask(sid, sid)
    T = X if Wd is None else X @ Wd
    S = (T @ Y.t()) * inv_tau
    out = estimator_from_scores(S, M, estimator)
    if grads:
        G = score_gradient(S, M, estimator) * inv_tau
        dT = G @ Y
        out["dY"] = G.t() @ T
        if Wd is None:
            out["dX"] = dT
        else:
            out["dX"] = dT @ Wd.t()
            out["dW"] = X.t() @ dT
    return out


def critic_loss_pair_form(X, Y, study_id, W=None, inv_tau: float = 1.0, estimator: str = "dv",
                          dtype=torch.float64, catloop: bool = False):
    """The reference's three-call sequence (main_utils.py:220-226) with the
    separable critic in the discriminator slot, differentiated by autograd.
    ``catloop``: build the pair tensor with the reference's own growing-``torch.cat`` loop."""
    X = X.detach().to("cpu", dtype).requires_grad_(True)
    Y = Y.detach().to("cpu", dtype).requires_grad_(True)
    Wd = None if W is None else W.detach().to("cpu", dtype).requires_grad_(True)
    rows = (create_mi_pairs_catloop if catloop else create_mi_pairs)(X, Y, list(study_id))
    logits = pair_logits_separable(rows, Wd, inv_tau)
    fn = {"dv": dv_bound_loss, "infonce": infonce_bound_loss}[estimator]
    loss = fn(logits, len(study_id))
    params = [X, Y] + ([Wd] if Wd is not None else [])
    g = torch.autograd.grad(loss.sum(), params)
    out = {"loss": loss.detach(), "dX": g[0], "dY": g[1]}
    if Wd is not None:
        out["dW"] = g[2]
    return out


# --------------------------------------------------------------------------
# synthetic "CXR-shaped" embeddings — shared by tests and bench
# --------------------------------------------------------------------------
def synthetic_embeddings(B: int, D: int, seed: int = 1234, dup_frac: float = 0.0,
                         device="cpu", bilinear: bool = True):
    """Image embedding is post-ReLU (>= 0, model.py:355-365); text embedding is
    tanh-pooled in (-1, 1) (model.py:76-77).  Correlated positives so the
    diagonal is informative.  Returns fp32 tensors (callers round to bf16)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.relu(torch.randn(B, D, generator=g))
    P0 = torch.randn(D, D, generator=torch.Generator().manual_seed(99)) / math.sqrt(D)
    Y = torch.tanh(0.5 * (X @ P0 + torch.randn(B, D, generator=g)))
    sid = torch.arange(B, dtype=torch.long)
    if dup_frac > 0:
        n_dup = int(B * dup_frac)
        idx = torch.randperm(B - 1, generator=g)[:n_dup]
        sid[idx + 1] = sid[idx]
    W = None
    if bilinear:
        gw = torch.Generator().manual_seed(7)
        W = (torch.eye(D) + 0.1 * torch.randn(D, D, generator=gw) / math.sqrt(D)) / math.sqrt(D)
    return X.to(device), Y.to(device), sid.to(device), (None if W is None else W.to(device))
