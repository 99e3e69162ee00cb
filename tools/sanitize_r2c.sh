#!/bin/bash
# compute-sanitizer over the tests of the kernels added in the third part of round 2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
OUT=gpurun_out/r2zz_sanitizer.txt
echo "# compute-sanitizer on B200, third part of round 2: memcheck over the tests of the chained / two-stream TAG launches (tile marks, narrow bulk copies), k_reduce_partials, FAConv kernels, GAT constant-memory slots and multi-head orchestration, k_mlp2_bwd_nh; racecheck over the chained-layer and two-stream tests" > $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider \
  -k "two_stream or chained or (tag_bwd_tensor and cigre14 and not reswitched) or gnn_dsse or (gat_dsse and cuda) or next_row" 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|passed|failed|Invalid|error" | head -20 >> $OUT
echo "rc_memcheck=${PIPESTATUS[0]}" >> $OUT
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider \
  -k "two_stream or chained" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard" | head -10 >> $OUT
echo "rc_racecheck=${PIPESTATUS[0]}" >> $OUT
cat $OUT
