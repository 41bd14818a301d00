#!/bin/bash
# 8-GPU A/B of the overlap knobs of the sharded step (gpurun --gpus 8 --timeout 900 -- 'bash scripts/n8_knobs.sh')
set -u
mkdir -p gpurun_out
out=gpurun_out/n8_knobs.txt
: > "$out"
N=${N:-8}
run() {
  name=$1; shift
  env "$@" MI_BENCH_STEPTIMES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
      > "gpurun_out/knob_$name.json" 2> "gpurun_out/knob_$name.err"
  python - "$name" >> "$out" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/knob_{sys.argv[1]}.json"))
    print(sys.argv[1], f"{d['ms_per_step']:.3f} ms/step", d["roofline"]["by_kernel"]["gemm"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  grep per-step "gpurun_out/knob_$name.err" >> "$out"
}
run base MI_RS_RESERVE_SMS=0
run reserve16 MI_RS_RESERVE_SMS=16
run reserve32 MI_RS_RESERVE_SMS=32
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run base2 MI_RS_RESERVE_SMS=0
cat "$out"
