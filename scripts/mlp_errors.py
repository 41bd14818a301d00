"""Prints the measured errors of the fused MLP-critic path against the CPU oracle (fp64) for both precisions.
python scripts/mlp_errors.py"""
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import ops, _lib
from oracle import matrix_oracle as mo, mlp_oracle

dev = torch.device("cuda:0")
NAMES = ("W1", "b1", "W2", "b2", "W3", "b3")
CASES = [(16, 768, 1024, 512, "dv", 1.0, 1.0, None), (64, 32, 128, 64, "dv", 2.0, 6.0, None), (96, 64, 256, 128, "infonce", 2.0, 6.0, None),
         (100, 40, 192, 96, "infonce_row", 2.0, 6.0, 3000), (256, 128, 1024, 512, "dv", 2.0, 6.0, 16384),
         (130, 72, 320, 264, "dv", 2.0, 6.0, 5000), (256, 768, 1024, 512, "dv", 1.0, 1.0, None)]
for B, D, H1, H2, est, s2, s3, panel in CASES:
    X, Y, sid, _ = mo.synthetic_embeddings(B, D, seed=B + D, dup_frac=0.05, bilinear=False)
    p = mlp_oracle.init_params(D, H1, H2, seed=H1 + H2)
    p["W2"] = p["W2"] * s2; p["W3"] = p["W3"] * s3
    t0 = time.time()
    ref = mlp_oracle.mlp_loss_matrix_form(X.float(), Y.float(), [int(s) for s in sid], p, est, dtype=torch.float64, row_chunk=16)
    t_ref = time.time() - t0
    params = tuple(p[k].to(dev).float() for k in NAMES)
    for prec in ("strict", "fast"):
        _lib.load().mi_set_mlp_panel_pairs(panel or 0)
        loss, S, g = ops.mlp_critic_loss_fwd_bwd(X.to(dev).float(), Y.to(dev).float(), params, torch.as_tensor(sid).int().to(dev), est, prec,
                                                 True, True)
        torch.cuda.synchronize()
        rel = lambda a, b: float((a.cpu().double().reshape(-1) - b.double().reshape(-1)).abs().max() / b.double().abs().max().clamp_min(1e-30))
        fro = lambda a, b: float((a.cpu().double().reshape(-1) - b.double().reshape(-1)).norm() / b.double().norm().clamp_min(1e-30))
        errs = {k: rel(g[k], ref[k]) for k in ("dX", "dY", "dW1", "db1", "dW2", "db2", "dW3")}
        errs.update({k + "_fro": fro(g[k], ref[k]) for k in ("dX", "dY", "dW1", "dW2")})
        print(f"B={B} D={D} H={H1}x{H2} {est} {prec}: loss {float(loss[0]):.6f} ref {float(ref['loss']):.6f} "
              f"abs {abs(float(loss[0]) - float(ref['loss'])):.2e} S maxabs {float((S.cpu().double() - ref['S']).abs().max()):.2e} "
              f"(|S| {float(ref['S'].abs().max()):.2f}) " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()) +
              f" db3 {float(g['db3']):.1e} (oracle {t_ref:.1f}s)", flush=True)
_lib.load().mi_set_mlp_panel_pairs(0)
