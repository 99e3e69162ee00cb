#!/bin/bash
# ncu --set full of the TMA-fed TAG backward-to-input kernel (k_tag_tc3<BGX>: the step's launches 41.. of k_tag_tc3, after the 40 forward ones)
# and of the packer / loss kernels inside one eager bench step
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_tag_tc3" -s 48 -c 1 -f -o gpurun_out/r2q_k_tag_tc3_bgx \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2q_ncu_bgx.log 2>&1
ncu --set full --clock-control none --profile-from-start off -k "regex:k_pack_copy|k_wls" -c 3 -f -o gpurun_out/r2q_pack_wls \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline >> gpurun_out/r2q_ncu_bgx.log 2>&1
ls -la gpurun_out/r2q_*
