#!/bin/bash
# Round-2 GPU session E: fwd3 (8 softmax warps) bring-up, bwd2 memcheck, green baseline of the default (gen-1) path.
mkdir -p gpurun_out
O=gpurun_out
( cd tests/native
  VJ_ATTN_FWD=3 timeout 200 ./test_attn fwd > ../../$O/r02_e_attn_fwd3_small.log 2>&1
  VJ_ATTN_FWD=3 timeout 200 ./test_attn fwdbig > ../../$O/r02_e_attn_fwd3_big.log 2>&1
  timeout 200 ./test_attn fwdbig > ../../$O/r02_e_attn_fwd1_big.log 2>&1
  VJ_ATTN_BWD=2 timeout 300 compute-sanitizer --tool memcheck --print-limit 20 ./test_attn bwdbig > ../../$O/r02_e_bwd2_memcheck.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o ../../$O/r02_prof_attn_fwd1 ./test_attn perf > ../../$O/r02_e_ncu_fwd1.log 2>&1
  VJ_ATTN_FWD=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -c 1 -o ../../$O/r02_prof_attn_fwd3 ./test_attn perf > ../../$O/r02_e_ncu_fwd3.log 2>&1 )
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_e_pytest.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_e_bench.json 2> $O/r02_e_bench.err
VJ_ATTN_FWD=3 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_e_bench_fwd3.json 2> $O/r02_e_bench_fwd3.err
grep -E "PERF|PASS|FAIL" $O/r02_e_attn_fwd3_small.log | tail -8
grep -E "PERF|PASSED|FAIL" $O/r02_e_attn_fwd3_big.log | tail -8
grep -E "PERF" $O/r02_e_attn_fwd1_big.log | tail -4
grep -E "Invalid|at |ERROR SUMMARY" $O/r02_e_bwd2_memcheck.log | head -12
tail -4 $O/r02_e_pytest.log; head -c 300 $O/r02_e_bench.json; echo; head -c 300 $O/r02_e_bench_fwd3.json
