#!/bin/bash
# Round-2 GPU session I: bwd2 full native suite (hoisted and per-step descriptors), perf vs gen 1.
mkdir -p gpurun_out
O=gpurun_out
cd tests/native
export VJ_TEST_BADMAP=1
for dbg in 0 1; do
  VJ_ATTN_BWD=2 VJ_BWD2_DEBUG=$dbg timeout 300 ./test_attn bwdbig > ../../$O/r02_i_bwd2_dbg$dbg.log 2>&1
  echo "== debug=$dbg"; grep -E "PERF|PASS|FAIL|badmap" ../../$O/r02_i_bwd2_dbg$dbg.log | cut -c1-300 | tail -14
done
VJ_ATTN_BWD=2 VJ_ATTN_PINGPONG=0 timeout 300 ./test_attn bwdbig > ../../$O/r02_i_bwd2_nopp.log 2>&1
echo "== nopp"; grep -E "PERF|FAIL|PASSED" ../../$O/r02_i_bwd2_nopp.log | tail -5
timeout 300 ./test_attn bwdbig > ../../$O/r02_i_bwd1.log 2>&1
echo "== gen1"; grep -E "PERF|FAIL|PASSED" ../../$O/r02_i_bwd1.log | tail -5
