"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (L0).

Run in the dev container only (needs /root/reference):

    python -m oracle.make_golden

For every case the reference's own ``create_mi_pairs`` (main_utils.py:80-110)
builds the pair rows, the separable critic is evaluated on those rows in the
``mi_discriminator`` slot (main_utils.py:222), and the reference's own
``dv_bound_loss`` / ``infonce_bound_loss`` (mi_critics.py:3-23) produce the
loss; gradients come from ``loss.backward()`` (main_utils.py:226).  The known
answers of SURVEY.md 8c (constant logits, shapes, N_neg = 0) are produced by
the same reference functions.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, B, D, critic, estimator, dup pattern, dtype, seed
    ("dv_dot_b32_d768_f32", 32, 768, "dot", "dv", None, torch.float32, 0),           # BASELINE config 1
    ("infonce_dot_b32_d768_f32", 32, 768, "dot", "infonce", None, torch.float32, 0),
    ("dv_bilinear_b16_d32_f64_dups", 16, 32, "bilinear", "dv", [(3, 4), (9, 10), (10, 11)], torch.float64, 1),
    ("infonce_bilinear_b16_d32_f64_dups", 16, 32, "bilinear", "infonce", [(3, 4), (9, 10), (10, 11)], torch.float64, 1),
    ("dv_bilinear_b24_d64_f64", 24, 64, "bilinear", "dv", None, torch.float64, 2),
    ("dv_dot_b8_d16_f64_dups", 8, 16, "dot", "dv", [(0, 7), (2, 5)], torch.float64, 3),
    ("dv_dot_b40_d48_f64_dups", 40, 48, "dot", "dv", [(0, 1), (1, 2), (2, 3), (20, 39)], torch.float64, 4),
]


def _inputs(B, D, critic, dups, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    # inputs are bf16-representable so the CUDA path (bf16 operands) sees IDENTICAL values
    X = torch.relu(torch.randn(B, D, generator=g)).bfloat16().to(dtype)
    Y = torch.tanh(0.5 * torch.randn(B, D, generator=g) + 0.3 * X.float()).bfloat16().to(dtype)
    sid = [str(50000000 + 7 * i) for i in range(B)]          # numeric strings (utils.py:16-18)
    for a, b in (dups or []):
        sid[b] = sid[a]
    if critic == "bilinear":
        W = ((torch.eye(D) + 0.1 * torch.randn(D, D, generator=g)) / D ** 0.5).bfloat16().to(dtype)
        inv_tau = 1.0
    else:
        W = None
        inv_tau = 1.0 / D ** 0.5
    return X, Y, sid, W, inv_tau


def run_reference(X, Y, sid, W, inv_tau, estimator):
    ref = ref_loader.load()
    X = X.clone().requires_grad_(True)
    Y = Y.clone().requires_grad_(True)
    Wp = None if W is None else W.clone().requires_grad_(True)
    rows = ref.create_mi_pairs(X, Y, sid, torch.device("cpu"))           # main_utils.py:220-221
    D = X.shape[1]
    t = rows[:, :D] if Wp is None else rows[:, :D] @ Wp
    logits = (t * rows[:, D:]).sum(1, keepdim=True) * inv_tau             # discriminator slot, [N,1]
    fn = ref.dv_bound_loss if estimator == "dv" else ref.infonce_bound_loss
    loss = fn(logits, len(sid), torch.device("cpu"))                      # main_utils.py:224
    loss.sum().backward()                                                 # main_utils.py:226
    return rows.detach(), logits.detach(), loss.detach(), X.grad, Y.grad, (None if Wp is None else Wp.grad)


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    for name, B, D, critic, est, dups, dtype, seed in CASES:
        X, Y, sid, W, inv_tau = _inputs(B, D, critic, dups, dtype, seed)
        rows, logits, loss, dX, dY, dW = run_reference(X, Y, sid, W, inv_tau, est)
        # ordering fingerprint: for every pair row, which (i, j) it is
        xi = (rows[:, :D].unsqueeze(1) == X.unsqueeze(0)).all(-1).float().argmax(1)
        payload = dict(
            X=X.numpy(), Y=Y.numpy(), sid=np.array([int(s) for s in sid], dtype=np.int64),
            inv_tau=np.float64(inv_tau), estimator=est, critic=critic,
            n_rows=np.int64(rows.shape[0]), logits=logits.numpy(), row_i=xi.numpy().astype(np.int32),
            loss=loss.numpy(), loss_shape=np.array(loss.shape, dtype=np.int64),
            dX=dX.numpy(), dY=dY.numpy(),
        )
        if W is not None:
            payload["W"] = W.numpy()
            payload["dW"] = dW.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
        print(f"{name}: rows={rows.shape[0]} loss={loss.reshape(-1)[0].item():.9f}")

    # known answers straight from the reference's estimator functions
    kat = {}
    dev = torch.device("cpu")
    c = torch.full((32 + 992, 1), 0.37)
    kat["const_dv"] = ref.dv_bound_loss(c, 32, dev).numpy()
    kat["const_infonce"] = ref.infonce_bound_loss(c, 32, dev).numpy()
    g = torch.Generator().manual_seed(11)
    l2 = torch.randn(16 + 240, 1, generator=g)
    kat["rand_logits"] = l2.numpy()
    kat["rand_dv"] = ref.dv_bound_loss(l2, 16, dev).numpy()
    kat["rand_infonce"] = ref.infonce_bound_loss(l2, 16, dev).numpy()
    kat["rand_dv_1d"] = ref.dv_bound_loss(l2[:, 0], 16, dev).numpy()
    kat["rand_infonce_1d"] = ref.infonce_bound_loss(l2[:, 0], 16, dev).numpy()
    lg = l2.clone().requires_grad_(True)
    ref.dv_bound_loss(lg, 16, dev).sum().backward()
    kat["rand_dv_dlogits"] = lg.grad.numpy()
    # all-same study id: no negatives (reference returns nan / -inf)
    e = torch.randn(4, 1, generator=g)
    kat["noneg_dv"] = ref.dv_bound_loss(e, 4, dev).numpy()
    kat["noneg_infonce"] = ref.infonce_bound_loss(e, 4, dev).numpy()
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), **kat)
    print("known answers:", {k: (v.shape, v.reshape(-1)[:1]) for k, v in kat.items() if "logits" not in k})


if __name__ == "__main__":
    main()
