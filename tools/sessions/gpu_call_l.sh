#!/bin/bash
# Round-2 GPU session L: backward with P^T / dS^T through tensor memory (VJ_ATTN_BWD=3) vs gen 1; full pytest (fixed tests).
mkdir -p gpurun_out
O=gpurun_out
( cd tests/native
  export VJ_TEST_BADMAP=1
  VJ_ATTN_BWD=3 timeout 300 ./test_attn > ../../$O/r02_l_attn_bwd3_small.log 2>&1
  VJ_ATTN_BWD=3 timeout 300 ./test_attn bwdbig > ../../$O/r02_l_attn_bwd3_big.log 2>&1
  timeout 300 ./test_attn bwdbig > ../../$O/r02_l_attn_bwd1_big.log 2>&1
  VJ_ATTN_BWD=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv --launch-skip 8 -c 1 -o ../../$O/r02_prof_attn_bwd3_hd32 ./test_attn perf > ../../$O/r02_l_ncu_bwd3.log 2>&1 )
grep -E "PERF|FAIL|PASSED|badmap" $O/r02_l_attn_bwd3_small.log | cut -c1-250 | tail -8
grep -E "PERF|FAIL|PASSED|badmap" $O/r02_l_attn_bwd3_big.log | cut -c1-250 | tail -10
grep -E "PERF|FAIL|PASSED" $O/r02_l_attn_bwd1_big.log | tail -4
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_l_pytest.log 2>&1
tail -6 $O/r02_l_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r02_l_bench.err | grep '^{"metric' > $O/r02_l_bench.json
VJ_ATTN_BWD=3 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r02_l_bench_bwd3.err | grep '^{"metric' > $O/r02_l_bench_bwd3.json
head -c 260 $O/r02_l_bench.json; echo; head -c 260 $O/r02_l_bench_bwd3.json
