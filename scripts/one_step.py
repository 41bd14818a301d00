"""One fused fwd+bwd step at the BASELINE size (for ncu launch lists / captures).
python scripts/one_step.py [B] [D] [estimator] [precision] [steps]"""
import math
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
est = sys.argv[3] if len(sys.argv) > 3 else "dv"
prec = sys.argv[4] if len(sys.argv) > 4 else "fast"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
X = torch.relu(torch.randn(B, D, generator=g)).to(dev).bfloat16()
Y = torch.tanh(torch.randn(B, D, generator=g)).to(dev).bfloat16()
W = (torch.eye(D) / D ** 0.5 + 0.01 * torch.randn(D, D, generator=g) / D).to(dev).bfloat16()
sid = torch.arange(B, dtype=torch.int32, device=dev)
for i in range(steps + 1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.critic_loss_fwd_bwd(X, Y, W, sid, est, prec, 1.0, True)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {i}: {e0.elapsed_time(e1):.2f} ms loss {float(out[0][0]):.6f}", flush=True)
