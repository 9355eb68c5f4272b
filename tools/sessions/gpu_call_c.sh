#!/bin/bash
# Round-2 GPU session C: re-validation after the MMA-issue-path rework of the second-generation attention kernels.
mkdir -p gpurun_out
O=gpurun_out
( cd tests/native
  timeout 300 ./test_gemm > ../../$O/r02_c_gemm.log 2>&1
  timeout 300 ./test_attn fwdbig > ../../$O/r02_c_attn_fwdbig.log 2>&1
  VJ_ATTN_PINGPONG=0 timeout 300 ./test_attn fwdbig > ../../$O/r02_c_attn_fwdbig_nopp.log 2>&1
  VJ_ATTN_POLY=2 timeout 300 ./test_attn fwdbig > ../../$O/r02_c_attn_fwdbig_poly2.log 2>&1
  VJ_ATTN_POLY=2 VJ_ATTN_PINGPONG=0 timeout 300 ./test_attn fwdbig > ../../$O/r02_c_attn_fwdbig_poly2_nopp.log 2>&1
  timeout 400 ./test_attn bwdbig > ../../$O/r02_c_attn_bwdbig.log 2>&1
  VJ_ATTN_PINGPONG=0 timeout 400 ./test_attn bwdbig > ../../$O/r02_c_attn_bwdbig_nopp.log 2>&1 )
VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rA > $O/r02_c_pytest_gen1.log 2>&1
ok=1
for f in r02_c_attn_fwdbig r02_c_attn_bwdbig; do grep -q "ALL PASSED" $O/$f.log || ok=0; done
echo "native_ok=$ok" > $O/r02_c_status.txt
if [ $ok = 1 ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rA > $O/r02_c_pytest_gen2.log 2>&1
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c_bench_gen2.json 2> $O/r02_c_bench_gen2.err
  VJ_ATTN_PINGPONG=0 timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c_bench_gen2_nopp.json 2> $O/r02_c_bench_gen2_nopp.err
fi
VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c_bench_gen1.json 2> $O/r02_c_bench_gen1.err
( cd tests/native
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -c 1 -o ../../$O/r02_prof_attn_fwd2 ./test_attn perf > ../../$O/r02_c_ncu_fwd2.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -c 1 -o ../../$O/r02_prof_attn_bwd2 ./test_attn perf > ../../$O/r02_c_ncu_bwd2.log 2>&1 )
tail -2 $O/r02_c_gemm.log; grep -E "PERF|PASSED|FAILED" $O/r02_c_attn_fwdbig.log | tail -5; grep -E "PERF|PASSED|FAILED" $O/r02_c_attn_bwdbig.log | tail -4
tail -3 $O/r02_c_pytest_gen1.log; tail -3 $O/r02_c_pytest_gen2.log 2>/dev/null; head -c 300 $O/r02_c_bench_gen2.json 2>/dev/null
