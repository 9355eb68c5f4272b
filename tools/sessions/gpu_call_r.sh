#!/bin/bash
# Round-2 GPU session R: final validation of the tree (full pytest, smoke, default bench) + ncu --set full of the dominant GEMM.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rA > $O/r02_r_pytest.log 2>&1
tail -3 $O/r02_r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py 2> $O/r02_r_bench.err | grep '^{"metric' > $O/r02_r_bench.json
head -c 230 $O/r02_r_bench.json; echo
( cd tests/native
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o ../../$O/r02_prof_gemm ./test_gemm perf:fc2_add_target > ../../$O/r02_r_ncu_gemm.log 2>&1 )
tail -2 $O/r02_r_ncu_gemm.log
