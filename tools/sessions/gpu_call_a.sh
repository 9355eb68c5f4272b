#!/bin/bash
# Round-2 GPU session A (run under gpurun from the repo root): safe-baseline tests, native bring-up of the new kernels,
# full test pass with the new kernels, the reference on the same B200, benches.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02_a_smi.txt; nproc >> $O/r02_a_smi.txt
# 1. first-generation kernels everywhere: validates the new TESTS (BASELINE shapes, reference-on-GPU, train entry, f2/f3/f4)
VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 VJ_GEMM_DIRECT=0 VJ_GEMM_STREAMK=0 timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rA 2>&1 | tail -120 > $O/r02_a_pytest_gen1.log
# 2. native bring-up of the new kernels
( cd tests/native
  timeout 300 ./test_gemm > ../../$O/r02_a_gemm_direct.log 2>&1
  VJ_GEMM_DIRECT=0 timeout 300 ./test_gemm perf > ../../$O/r02_a_gemm_tma_perf.log 2>&1
  VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 timeout 120 ./test_attn perf > ../../$O/r02_a_attn_perf_gen1.log 2>&1
  timeout 300 ./test_attn fwdbig > ../../$O/r02_a_attn_fwdbig.log 2>&1
  VJ_ATTN_PERSIST=0 timeout 300 ./test_attn fwdbig > ../../$O/r02_a_attn_fwdbig_np.log 2>&1
  VJ_ATTN_POLY=2 timeout 300 ./test_attn fwdbig > ../../$O/r02_a_attn_fwdbig_poly2.log 2>&1
  VJ_ATTN_POLY=3 timeout 300 ./test_attn fwdbig > ../../$O/r02_a_attn_fwdbig_poly3.log 2>&1
  timeout 400 ./test_attn bwdbig > ../../$O/r02_a_attn_bwdbig.log 2>&1
  VJ_ATTN_PERSIST=0 timeout 400 ./test_attn bwdbig > ../../$O/r02_a_attn_bwdbig_np.log 2>&1 )
# 3. the reference itself on this B200 (loggers off / on)
timeout 240 python tools/ref_gpu.py bench --config vitl16 --steps 8 --warmup 3 --loggers off --out $O/r02_ref_gpu_vitl16_off.json > $O/r02_a_ref_off.log 2>&1
timeout 240 python tools/ref_gpu.py bench --config vitl16 --steps 5 --warmup 2 --loggers on --out $O/r02_ref_gpu_vitl16_on.json > $O/r02_a_ref_on.log 2>&1
# 4. bench with the first-generation kernels (round-1 state + host changes) ...
VJ_ATTN_FWD=1 VJ_ATTN_BWD=1 VJ_GEMM_DIRECT=0 VJ_GEMM_STREAMK=0 timeout 500 python bench.py --steps 10 --warmup 3 > $O/r02_a_bench_gen1.json 2> $O/r02_a_bench_gen1.err
# 5. ... and, if the native tests passed, the full test suite and the bench with the new kernels
ok=1
for f in r02_a_gemm_direct r02_a_attn_fwdbig r02_a_attn_bwdbig; do grep -q "ALL PASSED" $O/$f.log || ok=0; done
echo "native_ok=$ok" > $O/r02_a_status.txt
if [ $ok = 1 ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rA 2>&1 | tail -120 > $O/r02_a_pytest_gen2.log
  timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_a_bench_gen2.json 2> $O/r02_a_bench_gen2.err
fi
tail -3 $O/r02_a_pytest_gen1.log; tail -2 $O/r02_a_gemm_direct.log; tail -2 $O/r02_a_attn_fwdbig.log; tail -2 $O/r02_a_attn_bwdbig.log
tail -c 300 $O/r02_a_ref_off.log; cat $O/r02_a_status.txt; tail -3 $O/r02_a_pytest_gen2.log 2>/dev/null
