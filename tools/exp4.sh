#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tag_fwd_tensor or tag_bwd_tensor or tie_aware or graphed_trainer or test_model_matches or large_graph or exact_weight or ragged or philox" 2>&1 | tail -15 > gpurun_out/exp4_tests.txt
cat gpurun_out/exp4_tests.txt
for v in 1 0; do
  DSS2_TC3=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp4_bench_tc3_$v.json 2> gpurun_out/exp4_bench_$v.err
done
DSS2_TILE_CAP=128 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp4_bench_tc3_cap128.json 2> gpurun_out/exp4_bench_cap128.err
python tools/show_bench.py gpurun_out/exp4_bench_tc3_1.json gpurun_out/exp4_bench_tc3_0.json gpurun_out/exp4_bench_tc3_cap128.json 2>&1 | grep -v "gw_ffma"
tail -n 3 gpurun_out/exp4_bench_*.err
