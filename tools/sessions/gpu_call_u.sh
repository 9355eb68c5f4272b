#!/bin/bash
# Round-2 GPU session U (8 GPUs): NCCL algorithm / channel report at N = 8, and the SM-reservation hypothesis
# (persistent GEMM grids leave 8 SMs to NCCL, NCCL bounded to 8 CTAs).
mkdir -p gpurun_out
O=gpurun_out
run8() { local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 14 --warmup 5 --no-cpu-baseline 2> $O/r02_u_bench_n8_$name.err | grep '^{"metric' > $O/r02_u_bench_n8_$name.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_u_bench_n8_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "gemm", d["roofline"]["dominant_kernel"]["gemm_ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e)
PY
}
run8 default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV,TUNING
run8 reserve8 VJ_NCCL_MAX_CTAS=8 VJ_SM_RESERVE=8
grep -E "NVLS|nChannels|Channel [0-9]+/|algo|Algo|proto|Using network|comm 0x.* rank 0 " $O/r02_u_bench_n8_default.err | sort | uniq -c | sort -rn | head -14 | cut -c1-200
