#!/bin/bash
# Round-2 GPU session B: profiles (ncu launch list of one step, --set full of the attention kernels), every BASELINE
# config at N=1, the dynamic-mask and as-shipped (loggers on) variants.  Run under gpurun from the repo root.
mkdir -p gpurun_out
O=gpurun_out
# launch list of one bench step (cold-cache, serialised: SHARES only)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 --csv \
    --log-file $O/r02_step_launches.csv python bench.py --profile --steps 1 --warmup 1 > $O/r02_b_ncu_step.log 2>&1
python tools/summarize_launches.py $O/r02_step_launches.csv > $O/r02_step_launches_summary.txt 2>&1
# --set full of the new attention kernels (native test binary: short command, 1 GPU)
( cd tests/native
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -c 2 -o ../../$O/r02_prof_attn_fwd2 ./test_attn perf > ../../$O/r02_b_ncu_fwd2.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -c 1 -o ../../$O/r02_prof_attn_bwd2 ./test_attn perf > ../../$O/r02_b_ncu_bwd2.log 2>&1 )
# every BASELINE config on one GPU + variants
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02_bench_vitl16_n1.json 2> $O/r02_b_bench_vitl16.err
timeout 500 python bench.py --config vith16 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_vith16_n1.json 2> $O/r02_b_bench_vith16.err
timeout 500 python bench.py --config vith16_384 --steps 8 --warmup 3 --no-cpu-baseline > $O/r02_bench_vith16_384_n1.json 2> $O/r02_b_bench_vith16_384.err
timeout 400 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --dynamic-masks > $O/r02_bench_vitl16_dynamic.json 2> $O/r02_b_bench_dyn.err
timeout 400 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --loggers > $O/r02_bench_vitl16_loggers.json 2> $O/r02_b_bench_log.err
timeout 300 python bench.py --config tiny --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_tiny.json 2> $O/r02_b_bench_tiny.err
head -30 $O/r02_step_launches_summary.txt; for f in vitl16_n1 vith16_n1 vith16_384_n1 vitl16_dynamic vitl16_loggers tiny; do head -c 400 $O/r02_bench_$f.json; echo; done
