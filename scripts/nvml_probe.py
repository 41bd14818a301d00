"""Which NVML query disturbs a running multi-GPU step?  Runs the sharded critic step in a loop and issues ONE kind of
NVML query from a side thread every 40 ms on rank 0; prints the step-time distribution per query kind.
torchrun --nproc-per-node 2 scripts/nvml_probe.py"""
import os
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mi_b200  # noqa
from mi_b200 import dist as mdist, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
B, D = 65536, 1024
X, Y, sid, W = synthetic.synthetic_embeddings(B, D, seed=1, dup_frac=0.05, bilinear=True)
Bl = B // world
Xl, Yl = X[rank * Bl:(rank + 1) * Bl].to(dev).bfloat16(), Y[rank * Bl:(rank + 1) * Bl].to(dev).bfloat16()
Wd, sl = W.to(dev).bfloat16(), sid[rank * Bl:(rank + 1) * Bl].to(torch.int32).to(dev)
step = lambda: mdist.sharded_critic_loss_fwd_bwd(Xl, Yl, Wd, sl, "dv", "fast", 1.0 / D ** 0.5)
for _ in range(5):
    step()
torch.cuda.synchronize()
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(local)
queries = {
    "none": lambda: None,
    "clock_sm": lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM),
    "reasons": lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h),
    "power": lambda: n.nvmlDeviceGetPowerUsage(h),
    "all3": lambda: (n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), n.nvmlDeviceGetCurrentClocksEventReasons(h), n.nvmlDeviceGetPowerUsage(h)),
}
for name, q in queries.items():
    stop = False
    lat = []

    def poll():
        while not stop:
            t = time.perf_counter()
            q()
            lat.append(time.perf_counter() - t)
            time.sleep(0.04)
    th = threading.Thread(target=poll, daemon=True)
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0 and name != "none":
        th.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(61)]
    evs[0].record()
    for i in range(60):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    stop = True
    ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(60))
    if rank == 0:
        ql = sorted(lat)
        print(f"{name:9s} steps: median {ts[30]:.2f} p90 {ts[54]:.2f} max {ts[-1]:.2f} mean {sum(ts) / 60:.2f} ms | "
              f"query latency median {1e3 * ql[len(ql) // 2] if ql else 0:.2f} max {1e3 * ql[-1] if ql else 0:.2f} ms ({len(ql)} polls)", flush=True)
    time.sleep(0.2)
dist.destroy_process_group()
