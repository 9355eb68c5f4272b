#!/bin/bash
# Round-2 GPU session P: validation of the tree before the 8-GPU run (new loss paths), Tiny config for the record.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_p_pytest.log 2>&1
tail -6 $O/r02_p_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r02_p_bench.err | grep '^{"metric' > $O/r02_p_bench.json
head -c 230 $O/r02_p_bench.json; echo
timeout 300 python bench.py --config tiny --steps 20 --warmup 5 --no-cpu-baseline 2> $O/r02_p_bench_tiny.err | grep '^{"metric' > $O/r02_p_bench_tiny.json
head -c 230 $O/r02_p_bench_tiny.json; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
