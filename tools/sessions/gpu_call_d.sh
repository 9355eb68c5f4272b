#!/bin/bash
# Round-2 GPU session D (re-entry): state of the tree after the container was re-created.
mkdir -p gpurun_out
O=gpurun_out
( cd tests/native
  timeout 200 ./test_gemm > ../../$O/r02_d_gemm.log 2>&1
  timeout 200 ./test_attn fwdbig > ../../$O/r02_d_attn_fwdbig.log 2>&1
  timeout 300 ./test_attn bwdbig > ../../$O/r02_d_attn_bwdbig.log 2>&1
  timeout 200 ./test_attn perf > ../../$O/r02_d_attn_perf.log 2>&1
  VJ_ATTN_PINGPONG=0 timeout 200 ./test_attn perf > ../../$O/r02_d_attn_perf_nopp.log 2>&1
  VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 timeout 200 ./test_attn perf > ../../$O/r02_d_attn_perf_gen1.log 2>&1 )
ok=1
for f in r02_d_attn_fwdbig r02_d_attn_bwdbig; do grep -q "ALL PASSED" $O/$f.log || ok=0; done
echo "native_ok=$ok" > $O/r02_d_status.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rA > $O/r02_d_pytest_gen2.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_d_bench_gen2.json 2> $O/r02_d_bench_gen2.err
VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_d_bench_gen1.json 2> $O/r02_d_bench_gen1.err
( cd tests/native
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -c 1 -o ../../$O/r02_prof_attn_fwd2 ./test_attn perf > ../../$O/r02_d_ncu_fwd2.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -c 1 -o ../../$O/r02_prof_attn_bwd2 ./test_attn perf > ../../$O/r02_d_ncu_bwd2.log 2>&1 )
tail -2 $O/r02_d_gemm.log; grep -E "PERF|PASSED|FAILED" $O/r02_d_attn_fwdbig.log | tail -5; grep -E "PERF|PASSED|FAILED" $O/r02_d_attn_bwdbig.log | tail -4
cat $O/r02_d_attn_perf.log | tail -12
tail -5 $O/r02_d_pytest_gen2.log; head -c 400 $O/r02_d_bench_gen2.json; echo; head -c 400 $O/r02_d_bench_gen1.json
