#!/bin/bash
# 8-GPU A/B of the sharded path (run under: gpurun --gpus 8 --timeout 900 -- 'bash scripts/n8_experiments.sh').
# Writes gpurun_out/n8_ab.txt: ms/step and per-step times for each setting, then a rank-0 kernel timeline.
set -u
mkdir -p gpurun_out
out=gpurun_out/n8_ab.txt
: > "$out"
N=${N:-8}
run() {   # name, extra bench args, env assignments...
  name=$1; extra=$2; shift 2
  env "$@" MI_BENCH_STEPTIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $extra \
      > "gpurun_out/bench_n${N}_$name.json" 2> "gpurun_out/bench_n${N}_$name.err"
  python - "$name" "$N" >> "$out" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_n{sys.argv[2]}_{sys.argv[1]}.json"))
print(sys.argv[1], f"{d['ms_per_step']:.3f} ms/step", f"{d['value']:.4e} pairs/s", "e2e", d.get("e2e", {}).get("ms_per_step"),
      "guard", d["config"]["guard_rows"], d["roofline"]["by_kernel"])
PY
  grep per-step "gpurun_out/bench_n${N}_$name.err" >> "$out"
}
run c_step "" MI_SHARDED_IMPL=c
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/dist_timeline.py > gpurun_out/timeline_n${N}.md 2> gpurun_out/timeline_n${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    examples/train_step_config4.py > gpurun_out/config4_n${N}.log 2> gpurun_out/config4_n${N}.err
cat "$out"
