#!/bin/bash
# Round-2 GPU session J: fwd4 as default (+ head dim 128 variant), bwd gen-1 ncu capture, full pytest, bench.
mkdir -p gpurun_out
O=gpurun_out
( cd tests/native
  timeout 200 ./test_attn fwd > ../../$O/r02_j_attn_fwd_small.log 2>&1
  timeout 200 ./test_attn fwdbig > ../../$O/r02_j_attn_fwd_big.log 2>&1
  VJ_ATTN_FWD=4 timeout 200 ./test_attn fwdbig > ../../$O/r02_j_attn_fwd4_hd128.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv -c 2 -o ../../$O/r02_prof_attn_bwd1 ./test_attn perf > ../../$O/r02_j_ncu_bwd1.log 2>&1 )
grep -E "FAIL|PASSED" $O/r02_j_attn_fwd_small.log | tail -3
grep -E "PERF|FAIL|PASSED" $O/r02_j_attn_fwd_big.log | tail -6
grep -E "PERF|FAIL|PASSED" $O/r02_j_attn_fwd4_hd128.log | tail -6
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_j_pytest.log 2>&1
tail -8 $O/r02_j_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r02_j_bench.err | grep '^{"metric' > $O/r02_j_bench.json
head -c 300 $O/r02_j_bench.json
