"""BASELINE config 5: estimator / shape sweep.  Prints a markdown table (pairs/s, F_alg/t) per
(B, D, critic, estimator, precision).  python scripts/sweep.py [quick]"""
import math
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import ops

dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
Bs = [256, 1024, 4096, 16384, 65536] + ([] if quick else [131072])
Ds = [128, 768, 1024] + ([] if quick else [256, 2048])
ESTS = ["dv", "infonce_sym"] if quick else ["dv", "infonce", "infonce_row", "infonce_sym"]


def timeit(fn, n):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("| B | D | critic | estimator | precision | ms/step | pairs/s | F_alg/t TFLOP/s |")
print("|---|---|---|---|---|---|---|---|")
for D in sorted(Ds):
    for B in Bs:
        if B * D > 131072 * 1024 * 2:
            continue
        g = torch.Generator().manual_seed(0)
        X = torch.relu(torch.randn(B, D, generator=g)).to(dev).bfloat16()
        Y = torch.tanh(torch.randn(B, D, generator=g)).to(dev).bfloat16()
        W = (torch.eye(D) / D ** 0.5).to(dev).bfloat16()
        sid = torch.arange(B, dtype=torch.int32, device=dev)
        for critic in ("bilinear", "dot"):
            for est in ESTS:
                for prec in (("fast",) if quick else ("fast", "strict")):
                    Wc = W if critic == "bilinear" else None
                    inv_tau = 1.0 if critic == "bilinear" else 1.0 / math.sqrt(D)
                    n = 20 if B <= 4096 else 5 if B <= 16384 else 2
                    out = (torch.empty(8, dtype=torch.float64, device=dev), torch.empty(B, D, device=dev),
                           torch.empty(B, D, device=dev), torch.empty(D, D, device=dev) if Wc is not None else None)
                    t = timeit(lambda: ops.critic_loss_fwd_bwd(X, Y, Wc, sid, est, prec, inv_tau, True, out=out), n)
                    falg = 6.0 * B * B * D + (6.0 * B * D * D if Wc is not None else 0)
                    print(f"| {B} | {D} | {critic} | {est} | {prec} | {t:.3f} | {B * B / t * 1e3:.3e} | {falg / t / 1e9:.0f} |", flush=True)
