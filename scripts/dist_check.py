"""Sharded (NCCL) path against the single-GPU fused call on the same inputs: every rank computes both and compares
its own rows.  torchrun --nproc-per-node N scripts/dist_check.py [B] [D]"""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import dist as mdist, ops
from mi_b200 import synthetic as mo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
Bl = B // world
off = rank * Bl
X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=7, dup_frac=0.05, bilinear=True)
Xd, Yd, Wd, sd = X.to(dev).bfloat16(), Y.to(dev).bfloat16(), W.to(dev).bfloat16(), sid.to(torch.int32).to(dev)
worst = {"fast": 0.0, "strict": 0.0}
for est in ("dv", "infonce_row", "infonce_sym"):
    for prec in ("fast", "strict"):
        out, dX, dY, dW = mdist.sharded_critic_loss_fwd_bwd(Xd[off:off + Bl], Yd[off:off + Bl], Wd, sd[off:off + Bl], est, prec, 1.0, True)
        ref, rX, rY, rW = ops.critic_loss_fwd_bwd(Xd, Yd, Wd, sd, est, prec, 1.0, True)
        torch.cuda.synchronize()
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
        e = (abs(float(out["loss"]) - float(ref[0])), rel(dX, rX[off:off + Bl]), rel(dY, rY[off:off + Bl]), rel(dW, rW))
        worst[prec] = max(worst[prec], *e[1:])
        if rank == 0:
            print(f"{est} {prec}: loss {float(out['loss']):.8f} vs {float(ref[0]):.8f}  |dloss| {e[0]:.1e}  dX {e[1]:.1e} dY {e[2]:.1e} dW {e[3]:.1e}", flush=True)
t = torch.tensor([worst["fast"], worst["strict"]], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    # the two paths round the same quantities to bf16 after fp32 sums taken in a different order: in fast mode a 1-ulp
    # fp32 difference can flip a bf16 rounding of dT (2^-9 of one element), so the bound is the fast-mode tolerance
    ok = float(t[0]) < 1e-2 and float(t[1]) < 1e-4
    print(f"worst relative gradient difference over ranks: fast {float(t[0]):.1e} (bound 1e-2), strict {float(t[1]):.1e} (bound 1e-4):",
          "OK" if ok else "MISMATCH")
dist.destroy_process_group()
