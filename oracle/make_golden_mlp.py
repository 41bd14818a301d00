"""Generate tests/golden/mlp_*.npz by EXECUTING THE REFERENCE's own critic path (L0):

    rows   = MultiModalManager.create_mi_pairs(X, Y, study_id, device)    # main_utils.py:80-110, :220
    logits = make_mlp(2D, [H1, H2])(rows)                                  # model.py:18-32, main_utils.py:77, :222
    loss   = dv_bound_loss / infonce_bound_loss(logits, B, device)         # mi_critics.py:3-23, main_utils.py:224
    loss.backward()                                                        # main_utils.py:226

Run in the dev container only (needs /root/reference):   python -m oracle.make_golden_mlp

The last case is the critic exactly as shipped, ``make_mlp(1536, [1024, 512])`` in float32 with PyTorch's
default initialisation under ``torch.manual_seed``; its 2.1 M parameters are not stored — the test rebuilds
them with ``mlp_oracle.init_params`` (same constructor order, same seed) — and the two large weight
gradients are stored as random-projection digests (``dW @ r`` and ``l^T dW`` with seeded r, l).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import mlp_oracle, ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, B, D, H1, H2, estimator, dups, dtype, seed, (W2 scale, W3 scale), store_params
    ("mlp_dv_b8_d16_h64x32_f64_dups", 8, 16, 64, 32, "dv", [(0, 7), (2, 5)], torch.float64, 21, (1.0, 1.0), True),
    ("mlp_infonce_b12_d24_h128x64_f64_dups", 12, 24, 128, 64, "infonce", [(3, 4), (4, 5)], torch.float64, 22, (2.0, 4.0), True),
    ("mlp_dv_b20_d40_h192x96_f64", 20, 40, 192, 96, "dv", None, torch.float64, 23, (2.0, 4.0), True),
    ("mlp_dv_b16_d768_h1024x512_f32_shipped", 16, 768, 1024, 512, "dv", [(5, 6)], torch.float32, 24, (1.0, 1.0), False),
]


def digest_vectors(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape[1], generator=g, dtype=torch.float64), torch.randn(shape[0], generator=g, dtype=torch.float64)


def inputs(B, D, dups, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.relu(torch.randn(B, D, generator=g)).bfloat16().to(dtype)           # image emb: post-ReLU (model.py:355-365)
    Y = torch.tanh(0.5 * torch.randn(B, D, generator=g) + 0.3 * X.float()).bfloat16().to(dtype)   # text emb: tanh-pooled
    sid = [str(50000000 + 7 * i) for i in range(B)]
    for a, b in (dups or []):
        sid[b] = sid[a]
    return X, Y, sid


def build_reference_mlp(ref, D, H1, H2, seed, dtype, scales):
    torch.manual_seed(seed)
    net = ref.make_mlp(2 * D, [H1, H2]).to(dtype)                                  # model.py:18-32
    with torch.no_grad():
        net[2].weight.mul_(scales[0])
        net[4].weight.mul_(scales[1])
    return net


def main():
    assert ref_loader.available(), "needs /root/reference"
    ref = ref_loader.load()
    assert ref.make_mlp is not None and ref.create_mi_pairs is not None, getattr(ref, "import_error", "")
    os.makedirs(OUT, exist_ok=True)
    dev = torch.device("cpu")
    for name, B, D, H1, H2, est, dups, dtype, seed, scales, store in CASES:
        X, Y, sid = inputs(B, D, dups, dtype, seed)
        net = build_reference_mlp(ref, D, H1, H2, seed, dtype, scales)
        Xl, Yl = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
        rows = ref.create_mi_pairs(Xl, Yl, sid, dev)
        logits = net(rows)
        fn = ref.dv_bound_loss if est == "dv" else ref.infonce_bound_loss
        loss = fn(logits, B, dev)
        loss.sum().backward()
        p = mlp_oracle.params_from_sequential(net)
        grads = dict(zip(mlp_oracle.PARAM_NAMES, [net[0].weight.grad, net[0].bias.grad, net[2].weight.grad,
                                                  net[2].bias.grad, net[4].weight.grad, net[4].bias.grad]))
        payload = dict(X=X.numpy(), Y=Y.numpy(), sid=np.array([int(s) for s in sid], dtype=np.int64),
                       estimator=est, dims=np.array([B, D, H1, H2], dtype=np.int64), seed=np.int64(seed),
                       scales=np.array(scales, dtype=np.float64), n_rows=np.int64(rows.shape[0]),
                       logits=logits.detach().numpy(), loss=loss.detach().numpy(),
                       loss_shape=np.array(loss.shape, dtype=np.int64), dX=Xl.grad.numpy(), dY=Yl.grad.numpy())
        for k in mlp_oracle.PARAM_NAMES:
            gk = grads[k]
            if store:
                payload[k] = p[k].numpy()
                payload["d" + k] = gk.numpy()
            elif gk.dim() == 2 and gk.numel() > 4096:
                r, l = digest_vectors(gk.shape, seed + 100)
                payload["d" + k + "_r"] = (gk.double() @ r).numpy()
                payload["d" + k + "_l"] = (l @ gk.double()).numpy()
                payload["d" + k + "_absmax"] = np.float64(gk.abs().max())
            else:
                payload["d" + k] = gk.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
        print(f"{name}: rows={rows.shape[0]} loss={loss.reshape(-1)[0].item():.9f} "
              f"logit range [{logits.min().item():.4f}, {logits.max().item():.4f}]")


if __name__ == "__main__":
    main()
