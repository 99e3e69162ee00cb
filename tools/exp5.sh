#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tag_fwd_tensor or tag_bwd_tensor or tie_aware or graphed_trainer or test_model_matches or exact_weight or ragged or philox" 2>&1 | tail -15 > gpurun_out/exp5_tests.txt
cat gpurun_out/exp5_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp5_bench.json 2> gpurun_out/exp5_bench.err
DSS2_TILE_CAP=128 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp5_bench_cap128.json 2> gpurun_out/exp5_bench_cap128.err
python tools/show_bench.py gpurun_out/exp5_bench.json gpurun_out/exp5_bench_cap128.json 2>&1 | grep -v "gw_ffma"
tail -n 3 gpurun_out/exp5_bench*.err
CAPS=256 tools/stamps.sh
