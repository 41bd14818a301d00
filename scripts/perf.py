"""Stage timings on the GPU (CUDA events).  python scripts/perf.py [B] [D]"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import ops, _lib

lib = _lib.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cgs = [int(c) for c in os.environ.get("PERF_CG", "2,1").split(",")]


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


g = torch.Generator().manual_seed(0)
X = torch.relu(torch.randn(B, D, generator=g)).to(dev).bfloat16()
Y = torch.tanh(torch.randn(B, D, generator=g)).to(dev).bfloat16()
W = (torch.eye(D) / D ** 0.5 + 0.01 * torch.randn(D, D, generator=g) / D).to(dev).bfloat16()
sid = torch.arange(B, dtype=torch.int32, device=dev)

for cg in cgs:
    lib.mi_set_cta_group(cg)
    print(f"=== cta_group {cg}  B={B} D={D}", flush=True)
    n = 8192
    A = torch.randn(n, n, device=dev).bfloat16(); Bm = torch.randn(n, n, device=dev).bfloat16()
    t, _ = timeit(lambda: ops.gemm(A, Bm), 5, 2)
    t2, _ = timeit(lambda: torch.matmul(A, Bm.t()), 5, 2)
    print(f"gemm 8192^3: mine {t:.3f} ms = {2*n**3/t/1e9:.0f} TF/s ; cuBLAS {t2:.3f} ms = {2*n**3/t2/1e9:.0f} TF/s", flush=True)
    del A, Bm
    T = ops.gemm(X, ops.transpose(W), out_dtype=torch.bfloat16)
    t, _ = timeit(lambda: ops.score_stats(T, Y, sid, sid, 0, 1.0), 3, 1)
    print(f"stats: {t:.2f} ms = {2*B*B*D/t/1e9:.0f} TF/s", flush=True)
    rows, scal = ops.score_stats(T, Y, sid, sid, 0, 1.0)
    lse = float(scal[0] + torch.log(scal[1]))
    ref = torch.full((B,), lse, device=dev)
    for prec in ("fast", "strict"):
        t, _ = timeit(lambda: ops.score_grad(T, Y, sid, sid, 0, 1.0, ref, 1.0, None, 0.0, False, prec, 1.0, 1.0 / B, want_k=True), 2, 1)
        print(f"grad pass, both outputs ({prec}): {t:.2f} ms = {6*B*B*D/t/1e9:.0f} TF/s (6B^2D)", flush=True)
    for est in ("dv", "infonce_sym"):
        t, _ = timeit(lambda: ops.critic_loss_fwd_bwd(X, Y, W, sid, est, "fast", 1.0, True), 2, 1)
        falg = 6 * B * B * D + 6 * B * D * D
        print(f"critic bilinear {est} fast fwd+bwd: {t:.2f} ms ; pairs/s {B*B/t*1e3:.3e} ; F_alg/t = {falg/t/1e9:.0f} TF/s", flush=True)
    t, _ = timeit(lambda: ops.critic_loss_fwd_bwd(X, Y, W, sid, "dv", "fast", 1.0, False), 2, 1)
    print(f"critic bilinear dv forward only: {t:.2f} ms", flush=True)
print("launches", ops.launch_count())
