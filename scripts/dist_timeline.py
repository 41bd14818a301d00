"""Kernel timeline of ONE sharded step (rank 0) from torch.profiler — where the non-engine time of the multi-GPU
step goes.  torchrun --nproc-per-node N scripts/dist_timeline.py [B] [D] > profiles/..."""
import math
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import dist as mdist
from mi_b200 import synthetic as mo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
Bl = B // world
off = rank * Bl
X, Y, sid, W = mo.synthetic_embeddings(B, D, seed=1234, dup_frac=0.05, bilinear=True)
Xd, Yd, Wd = X[off:off + Bl].to(dev).bfloat16(), Y[off:off + Bl].to(dev).bfloat16(), W.to(dev).bfloat16()
sd = sid[off:off + Bl].to(torch.int32).to(dev)
step = lambda: mdist.sharded_critic_loss_fwd_bwd(Xd, Yd, Wd, sd, "dv", "fast", 1.0, True)
for _ in range(5):
    step()
dist.barrier(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    # the middle step: between the 1st and 2nd occurrence boundaries of the first kernel name of a step
    print(f"# world {world} B {B} D {D}: CUDA activities of 3 steps on rank 0 (us since first)")
    print("| start us | dur us | stream | name |")
    print("|---|---|---|---|")
    for e in evs:
        name = e.name[:70]
        print(f"| {e.time_range.start - t0:.0f} | {e.time_range.end - e.time_range.start:.0f} | {getattr(e, 'stream', '')} | {name} |")
dist.destroy_process_group()
