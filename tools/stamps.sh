#!/bin/bash
# builds the diagnostic (-DDSS2_STAMPS) library on the GPU box and prints the phase breakdown of the layer kernels
cd "$GRAFT_REPO_ROOT/deep-statistical-solver-for-distribution-system-state-estimation_b200/csrc"
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DDSS2_STAMPS -c tag_tc2.cu -o tag_tc2.o || exit 1
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DDSS2_STAMPS -c tag_tc3.cu -o tag_tc3.o || exit 1
nvcc -shared -gencode arch=compute_100a,code=sm_100a graph.o tag.o tag_tc.o tag_tc2.o tag_tc3.o edgeagg.o wls.o optim.o gat.o -o libdss2_b200.so || exit 1
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cap in ${CAPS:-256}; do
  echo "#### DSS2_TILE_CAP=$cap"
  DSS2_TILE_CAP=$cap python tools/stamps.py
done 2>&1 | tee gpurun_out/stamps.txt
