"""L1 oracle, row-chunked: the matrix form of ``oracle/matrix_oracle.py`` for batches whose B x B score
matrix does not fit in memory (BASELINE configs 3 and 5: B = 65536 and beyond).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``) — a plain PyTorch restatement with the scores
materialised chunk by chunk (``chunk`` rows x B columns at a time), fp32 matmuls at full precision (TF32 off)
and fp64 reductions, on whatever device the inputs live on.  On the GPU box it is the "plain PyTorch fp32
reference of the same op" the full-size parity tests compare the CUDA path with; it shares nothing with the
product (no import of the package, no custom kernel).

Pinned: ``tests/test_oracle.py::test_chunked_oracle_equals_matrix_oracle`` proves it equal to
``matrix_oracle.critic_loss`` (itself pinned to the reference's golden vectors) for every estimator, with and
without duplicate study ids.  Formulas: SURVEY.md 7.2 / ``matrix_oracle.score_gradient``; reference lines
``mutual_info_img_txt/mi_critics.py:3-23`` (estimators), ``main_utils.py:99-108`` (negatives mask).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


def _lse_merge(m, s, vals, dim):
    """online (max, sum-exp) merge of a block of values along ``dim`` into the running pair (m, s) — fp64"""
    bm = vals.max(dim).values
    nm = torch.maximum(m, bm)
    safe = torch.where(torch.isfinite(nm), nm, torch.zeros_like(nm))
    s = s * torch.exp(torch.where(torch.isfinite(m), m - safe, torch.full_like(m, float("-inf")))) + \
        torch.exp(vals - safe.unsqueeze(dim)).sum(dim)
    return nm, s


def critic_loss_chunked(X: torch.Tensor, Y: torch.Tensor, sid: torch.Tensor, W: Optional[torch.Tensor] = None,
                        inv_tau: float = 1.0, estimator: str = "dv", chunk: int = 4096, grads: bool = True,
                        T: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Loss and (closed-form) gradients of the separable critic + estimator, scores formed ``chunk`` rows at a time.

    X, Y [B, D] fp32 (already rounded to whatever the tested path consumes), W [D, D] or None, sid [B] integer ids.
    ``T`` overrides the projection X W (the CUDA path's own bf16-rounded T, so that the comparison isolates the
    B^2-sized work).  Returns fp64 scalars and fp32 gradient matrices on the inputs' device."""
    allow = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = X.device
        B, D = X.shape
        Xf, Yf = X.float(), Y.float()
        Wf = None if W is None else W.float()
        Tf = (Xf if Wf is None else Xf @ Wf) if T is None else T.float()
        sid = sid.to(dev)
        f64 = dict(dtype=torch.float64, device=dev)
        ninf = float("-inf")
        dv_like = estimator in ("dv", "infonce")
        sym = estimator == "infonce_sym"
        diag = ((Tf.double() * Yf.double()).sum(1)) * inv_tau                 # S_ii, fp64
        # ---- pass 1: statistics
        g_m, g_s = torch.full((), ninf, **f64), torch.zeros((), **f64)          # global LSE over the negatives
        row_lse = torch.empty(B, **f64)                                          # LSE over {i} u negatives of row i
        c_m, c_s = torch.full((B,), ninf, **f64), torch.zeros(B, **f64)         # column LSE (symmetric form)
        n_neg = 0.0
        for r0 in range(0, B, chunk):
            r1 = min(B, r0 + chunk)
            S = (Tf[r0:r1] @ Yf.t()).double() * inv_tau
            M = sid[r0:r1, None] != sid[None, :]
            n_neg += float(M.sum())
            Sn = torch.where(M, S, torch.full_like(S, ninf))
            if dv_like:
                g_m, g_s = _lse_merge(g_m.reshape(1), g_s.reshape(1), Sn.reshape(1, -1), 1)
                g_m, g_s = g_m.reshape(()), g_s.reshape(())
            else:
                idx = torch.arange(r0, r1, device=dev)
                Sr = Sn.clone()
                Sr[idx - r0, idx] = S[idx - r0, idx]                            # the positive pair joins the softmax
                row_lse[r0:r1] = torch.logsumexp(Sr, 1)
                if sym:
                    c_m, c_s = _lse_merge(c_m, c_s, Sr, 0)
        out: Dict[str, torch.Tensor] = {"pos_mean": diag.mean(), "n_neg": torch.tensor(n_neg, **f64)}
        if dv_like:
            lse = g_m + torch.log(g_s)
            out["lse_neg"] = lse
            if estimator == "dv":      # mi_critics.py:10: log N_neg evaluated in fp32
                out["loss"] = lse - torch.log(torch.tensor(n_neg).float()).double().to(dev) - diag.mean()
            else:
                out["loss"] = lse - diag.mean()
        else:
            loss_row = (row_lse - diag).mean()
            out["row_lse"] = row_lse
            if sym:
                col_lse = c_m + torch.log(c_s)
                out["col_lse"] = col_lse
                out["loss"] = 0.5 * (loss_row + (col_lse - diag).mean())
            else:
                out["loss"] = loss_row
        if not grads:
            return out
        # ---- pass 2: G = dL/dS chunk by chunk; dT = inv_tau G Y, dY = inv_tau G^T T
        dT = torch.empty((B, D), dtype=torch.float32, device=dev)
        dY = torch.zeros((B, D), dtype=torch.float64, device=dev)
        for r0 in range(0, B, chunk):
            r1 = min(B, r0 + chunk)
            S = (Tf[r0:r1] @ Yf.t()).double() * inv_tau
            M = sid[r0:r1, None] != sid[None, :]
            idx = torch.arange(r0, r1, device=dev)
            if dv_like:
                G = torch.where(M, torch.exp(S - out["lse_neg"]), torch.zeros_like(S))
            else:
                R = M.clone()
                R[idx - r0, idx] = True
                G = torch.exp(S - row_lse[r0:r1, None]) / B
                if sym:
                    G = 0.5 * (G + torch.exp(S - out["col_lse"][None, :]) / B)
                G = torch.where(R, G, torch.zeros_like(G))
            G[idx - r0, idx] -= 1.0 / B
            G = (G * inv_tau).float()
            dT[r0:r1] = G @ Yf
            dY += (G.t() @ Tf[r0:r1]).double()
        out["dY"] = dY.float()
        if Wf is None:
            out["dX"] = dT
        else:
            out["dT"] = dT
            out["dX"] = dT @ Wf.t()
            out["dW"] = (Xf.double().t() @ dT.double()).float()
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32 = allow
