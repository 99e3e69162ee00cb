#!/bin/bash
# experiment: tile caps for the tcgen05 layer kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cap in 96 256; do
  DSS2_TILE_CAP=$cap timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tag_fwd_tensor or tag_bwd_tensor or tie_aware or graphed_trainer or test_model_matches" 2>&1 | tail -5 > gpurun_out/exp1_tests_cap$cap.txt
done
for cap in 96 128 256; do
  DSS2_TILE_CAP=$cap timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp1_bench_cap$cap.json 2> gpurun_out/exp1_bench_cap$cap.err
done
tail -3 gpurun_out/exp1_tests_cap*.txt
python tools/show_bench.py gpurun_out/exp1_bench_cap96.json gpurun_out/exp1_bench_cap128.json gpurun_out/exp1_bench_cap256.json 2>&1 | tail -60
