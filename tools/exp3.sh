#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for f in 0 1 2 3; do
  DSS2_TC2_FLAGS=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/exp3_bench_f$f.json 2> gpurun_out/exp3_f$f.err
done
python tools/show_bench.py gpurun_out/exp3_bench_f*.json 2>&1 | grep -v "gw_ffma"
cd deep-statistical-solver-for-distribution-system-state-estimation_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DDSS2_STAMPS -c tag_tc2.cu -o tag_tc2.o || exit 1
nvcc -shared -gencode arch=compute_100a,code=sm_100a graph.o tag.o tag_tc.o tag_tc2.o tag_tc3.o edgeagg.o wls.o optim.o gat.o -o libdss2_b200.so || exit 1
cd "$GRAFT_REPO_ROOT"
for f in 0 1 2 3; do echo "### flags $f"; DSS2_TC2_FLAGS=$f python tools/stamps.py; done 2>&1 | tee gpurun_out/stamps3.txt
