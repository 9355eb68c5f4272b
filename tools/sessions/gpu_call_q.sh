#!/bin/bash
# Round-2 GPU session Q (8 GPUs): BASELINE configs C3 / C4 / C5 as the driver launches them.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02_q_gpus.txt
run8() { # name, extra bench args...
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --no-cpu-baseline "$@" 2> $O/r02_q_bench_n8_$name.err | grep '^{"metric' > $O/r02_q_bench_n8_$name.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_q_bench_n8_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "gemm", d["roofline"]["dominant_kernel"]["gemm_ms_per_step"], d.get("ddp_check",{}).get("rel_l2_sync_vs_allreduce_mean"))
except Exception as e:
    print("$name failed", e)
PY
}
run8 vitl16 --steps 12 --warmup 3
run8 vith16 --config vith16 --steps 8 --warmup 3
run8 vith16_384 --config vith16_384 --steps 6 --warmup 3
