"""CPU oracle (TEST INFRASTRUCTURE) for the reference's own concat-MLP critic, SURVEY.md 8f-1.

The reference's critic is ``self.mi_discriminator = make_mlp(1536, [1024, 512])``
(main_utils.py:77; model.py:18-32): ``Linear(2D,H1) -> ReLU -> Linear(H1,H2) -> ReLU -> Linear(H2,1)``
applied to every row ``[x_i ; y_j]`` of the tensor ``create_mi_pairs`` builds (main_utils.py:80-110,
called at :220-222), followed by ``dv_bound_loss`` / ``infonce_bound_loss`` (mi_critics.py:3-23, :224)
and ``loss.backward()`` (:226).

Two restatements:

* ``mlp_loss_pair_form``  — the reference sequence itself on the oracle's pair tensor (row order of
  main_utils.py:99-108); O(B^2) rows of width 2D, small B only.
* ``mlp_loss_matrix_form`` — what the CUDA path computes.  Layer 1 is separable,
  ``W1 [x;y] + b1 = (W1x x_i + b1) + W1y y_j = A_i + C_j``, so the logit matrix is
  ``S[i,j] = w3 . relu(W2 relu(A_i + C_j) + b2) + b3`` and the estimators are those of
  ``matrix_oracle.estimator_from_scores`` on (S, negatives mask).  Row-chunked; gradients by autograd.

Pinned against the reference: ``tests/golden/mlp_*.npz`` are produced by EXECUTING the reference's
``make_mlp`` + ``create_mi_pairs`` + estimator functions (``oracle/make_golden_mlp.py``) and
``tests/test_oracle_mlp.py`` checks both restatements against them (and against the live reference
when /root/reference is present).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import matrix_oracle as mo

PARAM_NAMES = ("W1", "b1", "W2", "b2", "W3", "b3")      # nn.Sequential keys 0.weight 0.bias 2.weight 2.bias 4.weight 4.bias


def params_from_sequential(seq) -> Dict[str, torch.Tensor]:
    """The six tensors of a ``make_mlp(2D, [H1, H2])`` module (model.py:18-32), in order."""
    sd = seq.state_dict()
    keys = ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"]
    return {n: sd[k].detach().clone() for n, k in zip(PARAM_NAMES, keys)}


def init_params(D: int, H1: int, H2: int, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """nn.Linear default initialisation in the construction order of make_mlp (model.py:27-30) under
    ``torch.manual_seed(seed)`` — bit-identical to ``make_mlp(2*D, [H1, H2])`` built under the same seed."""
    torch.manual_seed(seed)
    l1 = torch.nn.Linear(2 * D, H1)
    l2 = torch.nn.Linear(H1, H2)
    l3 = torch.nn.Linear(H2, 1)
    vals = [l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias]
    return {n: v.detach().clone().to(dtype) for n, v in zip(PARAM_NAMES, vals)}


def mlp_logits_on_rows(rows: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    """nn.Sequential(Linear, ReLU, Linear, ReLU, Linear) on pair rows -> [N, 1] (model.py:27-32)."""
    h1 = torch.relu(rows @ p["W1"].t() + p["b1"])
    h2 = torch.relu(h1 @ p["W2"].t() + p["b2"])
    return h2 @ p["W3"].t() + p["b3"]


def mlp_score_matrix(X, Y, p: Dict[str, torch.Tensor], row_chunk: int = 64) -> torch.Tensor:
    """S[i,j] = mlp([x_i ; y_j]) for ALL (i, j), using the separable first layer."""
    D = X.shape[1]
    A = X @ p["W1"][:, :D].t() + p["b1"]                    # [B, H1]
    C = Y @ p["W1"][:, D:].t()                              # [B, H1]
    out = []
    for r0 in range(0, X.shape[0], row_chunk):
        h1 = torch.relu(A[r0:r0 + row_chunk, None, :] + C[None, :, :])          # [r, B, H1]
        h2 = torch.relu(h1 @ p["W2"].t() + p["b2"])                            # [r, B, H2]
        out.append((h2 @ p["W3"].t()).squeeze(-1) + p["b3"])                    # [r, B]
    return torch.cat(out, 0)


def _leafs(X, Y, p, dtype):
    X = X.detach().to(dtype).clone().requires_grad_(True)
    Y = Y.detach().to(dtype).clone().requires_grad_(True)
    q = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}
    return X, Y, q


def _pack(loss, X, Y, q, extra=None):
    out = {"loss": loss.detach(), "dX": X.grad, "dY": Y.grad}
    for k, v in q.items():
        out["d" + k] = v.grad if v.grad is not None else torch.zeros_like(v)
    if extra:
        out.update(extra)
    return out


def mlp_loss_matrix_form(X, Y, study_id: Sequence, p: Dict[str, torch.Tensor], estimator: str = "dv",
                         dtype=torch.float64, row_chunk: int = 64) -> Dict[str, torch.Tensor]:
    X, Y, q = _leafs(X, Y, p, dtype)
    sid = mo.dense_ids(study_id)
    S = mlp_score_matrix(X, Y, q, row_chunk)
    M = mo.negatives_mask(sid, sid)
    est = mo.estimator_from_scores(S, M, estimator)
    est["loss"].backward()
    return _pack(est["loss"], X, Y, q, {"S": S.detach(), "n_neg": est["n_neg"]})


def mlp_loss_pair_form(X, Y, study_id: Sequence, p: Dict[str, torch.Tensor], estimator: str = "dv",
                       dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """main_utils.py:220-226 with the oracle's own pair builder and estimator restatements."""
    X, Y, q = _leafs(X, Y, p, dtype)
    rows = mo.create_mi_pairs(X, Y, study_id)
    logits = mlp_logits_on_rows(rows, q)
    fn = mo.dv_bound_loss if estimator == "dv" else mo.infonce_bound_loss
    loss = fn(logits, len(study_id))
    loss.sum().backward()
    return _pack(loss, X, Y, q, {"logits": logits.detach()})
