#!/bin/bash
# Round-2 GPU session G: fwd4 after the register-pressure fix, polynomial-exp offload A/B, step-level effect.
mkdir -p gpurun_out
O=gpurun_out
cd tests/native
for poly in 0 2; do
  VJ_ATTN_POLY=$poly VJ_ATTN_FWD=4 timeout 200 ./test_attn fwd > ../../$O/r02_g_attn_fwd4_p${poly}_small.log 2>&1
  VJ_ATTN_POLY=$poly VJ_ATTN_FWD=4 timeout 200 ./test_attn fwdbig > ../../$O/r02_g_attn_fwd4_p${poly}_big.log 2>&1
  grep -E "FAIL|PASSED" ../../$O/r02_g_attn_fwd4_p${poly}_small.log | tail -3
  grep -E "PERF|FAIL|PASSED" ../../$O/r02_g_attn_fwd4_p${poly}_big.log | tail -8
done
cd ../..
VJ_ATTN_FWD=4 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_g_bench_fwd4.json 2> $O/r02_g_bench_fwd4.err
VJ_ATTN_POLY=2 VJ_ATTN_FWD=4 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_g_bench_fwd4_p2.json 2> $O/r02_g_bench_fwd4_p2.err
head -c 250 $O/r02_g_bench_fwd4.json; echo; head -c 250 $O/r02_g_bench_fwd4_p2.json
