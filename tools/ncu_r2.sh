#!/bin/bash
# round-2 ncu evidence: launch list of one eager step + --set full captures of the hot kernels (one GPU)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
python tools/ncu_launches.py gpurun_out/r2_launches.csv > gpurun_out/r2_launch_list_summary.txt
head -30 gpurun_out/r2_launch_list_summary.txt
# full captures: 2 launches of each hot kernel from the timed step (profile-from-start off = only the step between cudaProfilerStart/Stop)
for pat in k_tag_tc3 k_tag_gw k_edgeagg_fwd k_edgeagg_bwd k_wls; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$pat -s 4 -c 2 -f -o gpurun_out/r2_$pat \
      python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_$pat.log 2>&1
  ls -la gpurun_out/r2_$pat.ncu-rep
done
