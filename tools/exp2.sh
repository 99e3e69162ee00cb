#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tag_fwd_tensor or tag_bwd_tensor or tie_aware or graphed_trainer or test_model_matches or large_graph" 2>&1 | tail -15 > gpurun_out/exp2_tests.txt
cat gpurun_out/exp2_tests.txt
for ta in 1 0; do
  DSS2_TC2_TA=$ta timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp2_bench_ta$ta.json 2> gpurun_out/exp2_bench_ta$ta.err
done
DSS2_TILE_CAP=128 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp2_bench_ta1_cap128.json 2> gpurun_out/exp2_bench_cap128.err
DSS2_TC2_RW3=1 DSS2_TILE_CAP=96 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp2_bench_ta1_rw3.json 2> gpurun_out/exp2_bench_rw3.err
DSS2_TC2_TA=0 DSS2_TC2_RW3=1 DSS2_TILE_CAP=96 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp2_bench_ta0_rw3.json 2> gpurun_out/exp2_bench_rw3b.err
python tools/show_bench.py gpurun_out/exp2_bench_ta0_rw3.json gpurun_out/exp2_bench_ta1.json gpurun_out/exp2_bench_ta0.json gpurun_out/exp2_bench_ta1_cap128.json gpurun_out/exp2_bench_ta1_rw3.json 2>&1 | grep -v "gw_ffma"
tail -3 gpurun_out/exp2_bench_*.err
CAPS=256 tools/stamps.sh
