#!/bin/bash
# Round-2 GPU session O (2 GPUs): NCCL gradient-equality test, N=2 bench with NCCL CTA budgets, N=1 on the same box.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02_o_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --timeout 500 > $O/r02_o_pytest_multi.log 2>&1
tail -4 $O/r02_o_pytest_multi.log
run2() { # name, env...
  local name=$1; shift
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline 2> $O/r02_o_bench_n2_$name.err | grep '^{"metric' > $O/r02_o_bench_n2_$name.json
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_o_bench_n2_$name.json"))
    print("$name", d["value"], d["ms_per_step"], "gemm", d["roofline"]["dominant_kernel"]["gemm_ms_per_step"], d.get("ddp_check"))
except Exception as e:
    print("$name failed", e)
PY
}
run2 default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV
run2 cta8 VJ_NCCL_MAX_CTAS=8
run2 cta4 VJ_NCCL_MAX_CTAS=4
run2 cta2 VJ_NCCL_MAX_CTAS=2
timeout 400 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2> $O/r02_o_bench_n1.err | grep '^{"metric' > $O/r02_o_bench_n1.json
python -c "
import json; d=json.load(open('$O/r02_o_bench_n1.json')); print('n1', d['value'], d['ms_per_step'], d['roofline']['dominant_kernel']['gemm_ms_per_step'])"
grep -E "NVLS|Channel|channels|nChannels|Connected|NET/|P2P" $O/r02_o_bench_n2_default.err | head -12
