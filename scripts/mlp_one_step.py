"""One fused fwd+bwd step of the MLP-critic path (for ncu launch lists / captures).
python scripts/mlp_one_step.py [B] [precision] [steps]"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "fast"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
D, H1, H2 = 768, 1024, 512
g = torch.Generator().manual_seed(B)
X = torch.relu(torch.randn(B, D, generator=g)).to(dev)
Y = torch.tanh(torch.randn(B, D, generator=g)).to(dev)
sid = torch.arange(B, dtype=torch.int32, device=dev)
torch.manual_seed(1)
critic = mi_b200.FusedMLPCritic(D, (H1, H2)).to(dev)
params = (critic[0].weight, critic[0].bias, critic[2].weight, critic[2].bias, critic[4].weight, critic[4].bias)
for i in range(steps + 1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.mlp_critic_loss_fwd_bwd(X, Y, params, sid, "dv", prec, True, False)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {i}: {e0.elapsed_time(e1):.3f} ms loss {float(out[0][0]):.6f}", flush=True)
