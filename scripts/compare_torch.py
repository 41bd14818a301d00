"""BASELINE config 2: fused op vs the 'reference torch path' on the same GPU (matrix form with S
materialised, cuBLAS bf16 GEMM + ATen logsumexp + autograd).  python scripts/compare_torch.py"""
import math
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import ops

dev = torch.device("cuda:0")


def torch_path(X, Y, W, sid, est):
    X = X.detach().requires_grad_(True)
    Y = Y.detach().requires_grad_(True)
    W = W.detach().requires_grad_(True)
    B = X.shape[0]
    S = ((X @ W) @ Y.t()).float()
    M = sid[:, None] != sid[None, :]
    eye = torch.eye(B, dtype=torch.bool, device=X.device)
    diag = torch.diagonal(S)
    ninf = torch.full_like(S, float("-inf"))
    if est == "dv":
        loss = torch.logsumexp(torch.where(M, S, ninf).reshape(-1), 0) - math.log(float(M.sum())) - diag.mean()
    else:
        R = M | eye
        Sm = torch.where(R, S, ninf)
        loss = 0.5 * ((torch.logsumexp(Sm, 1) - diag).mean() + (torch.logsumexp(Sm, 0) - diag).mean())
    loss.backward()
    return loss


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


print("| B | D | estimator | fused ms | fused, CUDA graph ms | torch (S materialised) ms | speed-up | torch peak MB |")
print("|---|---|---|---|---|---|---|---|")
for (B, D) in [(256, 768), (1024, 768), (4096, 768), (8192, 768), (16384, 1024), (32768, 1024)]:
    g = torch.Generator().manual_seed(0)
    X = torch.relu(torch.randn(B, D, generator=g)).to(dev).bfloat16()
    Y = torch.tanh(torch.randn(B, D, generator=g)).to(dev).bfloat16()
    W = (torch.eye(D) / D ** 0.5).to(dev).bfloat16()
    sid = torch.arange(B, dtype=torch.int32, device=dev)
    for est, name in (("dv", "dv"), ("infonce_sym", "infonce_sym")):
        t_f = timeit(lambda: ops.critic_loss_fwd_bwd(X, Y, W, sid, est, "fast", 1.0, True), 10)
        gstep = ops.GraphedCriticStep(B, D, True, est, "fast", 1.0, dev)
        t_g = timeit(lambda: gstep(X, Y, W, sid), 10)
        del gstep
        torch.cuda.reset_peak_memory_stats()
        try:
            t_t = timeit(lambda: torch_path(X, Y, W, sid, est), 3)
            mem = torch.cuda.max_memory_allocated() / 2 ** 20
            print(f"| {B} | {D} | {name} | {t_f:.3f} | {t_g:.3f} | {t_t:.3f} | {t_t / min(t_f, t_g):.1f}x | {mem:.0f} |", flush=True)
        except torch.OutOfMemoryError:
            print(f"| {B} | {D} | {name} | {t_f:.3f} | {t_g:.3f} | OOM | - | - |", flush=True)
            torch.cuda.empty_cache()
