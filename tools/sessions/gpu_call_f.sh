#!/bin/bash
# Round-2 GPU session F: fwd4 (64-key tiles, 4 / 3 serial CTAs per SM) bring-up.
mkdir -p gpurun_out
O=gpurun_out
cd tests/native
for g in 4 5; do
  VJ_ATTN_FWD=$g timeout 200 ./test_attn fwd > ../../$O/r02_f_attn_fwd${g}_small.log 2>&1
  VJ_ATTN_FWD=$g timeout 200 ./test_attn fwdbig > ../../$O/r02_f_attn_fwd${g}_big.log 2>&1
  grep -E "PERF|FAIL|PASSED" ../../$O/r02_f_attn_fwd${g}_small.log | tail -4
  grep -E "PERF|FAIL|PASSED" ../../$O/r02_f_attn_fwd${g}_big.log | tail -8
done
VJ_ATTN_FWD=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd4 -c 1 -o ../../$O/r02_prof_attn_fwd4 ./test_attn perf > ../../$O/r02_f_ncu_fwd4.log 2>&1
