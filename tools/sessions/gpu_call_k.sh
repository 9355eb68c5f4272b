#!/bin/bash
# Round-2 GPU session K: full pytest with fwd4 everywhere, ncu of the fused hd-32 backward, ViT-H bench lines.
mkdir -p gpurun_out
O=gpurun_out
( cd tests/native
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:dkv_kernel<32" -c 1 -o ../../$O/r02_prof_attn_bwd1_hd32 ./test_attn perf > ../../$O/r02_k_ncu_bwd1.log 2>&1 )
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_k_pytest.log 2>&1
tail -8 $O/r02_k_pytest.log
timeout 500 python bench.py --config vith16 --steps 8 --warmup 3 --no-cpu-baseline 2> $O/r02_k_bench_vith16.err | grep '^{"metric' > $O/r02_k_bench_vith16.json
head -c 300 $O/r02_k_bench_vith16.json; echo
timeout 600 python bench.py --config vith16_384 --steps 6 --warmup 3 --no-cpu-baseline 2> $O/r02_k_bench_vith16_384.err | grep '^{"metric' > $O/r02_k_bench_vith16_384.json
head -c 300 $O/r02_k_bench_vith16_384.json; echo; tail -3 $O/r02_k_bench_vith16_384.err
