#!/bin/bash
# 8-GPU A/B of the sharded-path knobs (run under: gpurun --gpus 8 --timeout 600 -- 'bash scripts/n8_experiments.sh').
# Writes gpurun_out/n8_ab.txt: ms/step and per-step times for each setting, then a rank-0 kernel timeline of the best one.
set -u
mkdir -p gpurun_out
out=gpurun_out/n8_ab.txt
: > "$out"
run() {   # name, env assignments...
  name=$1; shift
  env "$@" MI_BENCH_STEPTIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
      > "gpurun_out/bench_n8_$name.json" 2> "gpurun_out/bench_n8_$name.err"
  python - "$name" >> "$out" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_n8_{sys.argv[1]}.json"))
print(sys.argv[1], f"{d['ms_per_step']:.3f} ms/step", f"{d['value']:.4e} pairs/s", d["roofline"]["by_kernel"])
PY
  grep per-step "gpurun_out/bench_n8_$name.err" >> "$out"
}
run base MI_OWN_COLUMNS_FIRST=0
run own_first MI_OWN_COLUMNS_FIRST=1
run base_again MI_OWN_COLUMNS_FIRST=0
MI_OWN_COLUMNS_FIRST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    scripts/dist_timeline.py > gpurun_out/timeline_n8_own_first.md 2> gpurun_out/timeline_n8_own_first.err
cat "$out"
