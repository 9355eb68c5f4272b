#!/bin/bash
# Round-2 GPU session S (8 GPUs): the N = 8 point of the scaling table, launched as the driver does (20 steps / 5 warm-ups).
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 8 --steps 20 --warmup 5 2> $O/r02_t_bench_n8.err | grep '^{"metric' > $O/r02_t_bench_n8.json
python - <<PY
import json
d=json.load(open("$O/r02_t_bench_n8.json"))
print("n8", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "gemm", d["roofline"]["dominant_kernel"]["gemm_ms_per_step"], d.get("ddp_check",{}).get("rel_l2_sync_vs_allreduce_mean"), d["clocks"])
PY
