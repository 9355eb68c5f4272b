#!/bin/bash
# Round-2 GPU session N: ncu launch list of one step (current kernels), full pytest, default bench + reference arm.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_step_launches.csv python bench.py --profile --steps 1 --warmup 1 > $O/r02_n_profile_run.log 2>&1
tail -2 $O/r02_n_profile_run.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_n_pytest.log 2>&1
tail -4 $O/r02_n_pytest.log
timeout 600 python bench.py 2> $O/r02_n_bench.err | grep '^{"metric' > $O/r02_n_bench.json
head -c 200 $O/r02_n_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> $O/r02_n_bench_ref.err | grep '^{' > $O/r02_n_bench_ref.json
head -c 400 $O/r02_n_bench_ref.json; echo
timeout 400 python bench.py --dynamic-masks --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r02_n_bench_dyn.err | grep '^{"metric' > $O/r02_n_bench_dyn.json
timeout 400 python bench.py --loggers --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r02_n_bench_log.err | grep '^{"metric' > $O/r02_n_bench_log.json
head -c 200 $O/r02_n_bench_dyn.json; echo; head -c 200 $O/r02_n_bench_log.json; echo
