#!/bin/bash
# ncu --set full of the EdgeAggregation kernels at the bench shape (tools/ea_bench.py); reports land in gpurun_out/
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-r2i}
for k in k_ea_row_fwd k_ea_row_bwd; do
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 3 -c 1 -f -o gpurun_out/${TAG}_$k python tools/ea_bench.py > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ea_launches.csv python tools/ea_bench.py > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
