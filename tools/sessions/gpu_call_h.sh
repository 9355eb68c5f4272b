#!/bin/bash
# Round-2 GPU session H: bwd2 bring-up diagnostics.
mkdir -p gpurun_out
O=gpurun_out
cd tests/native
export VJ_TEST_BADMAP=1 VJ_ATTN_BWD=2
for dbg in 0 1; do
  for L in 128 64 256; do
    echo "== debug=$dbg L=$L"
    VJ_BWD2_DEBUG=$dbg timeout 100 ./test_attn bwdone $L 2>&1 | tail -12
  done
done > ../../$O/r02_h_bwd2_diag.log 2>&1
cat ../../$O/r02_h_bwd2_diag.log | cut -c1-600
