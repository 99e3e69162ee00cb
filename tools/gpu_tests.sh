#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} 2>&1 | tail -40 ) > gpurun_out/gpu_tests.txt 2>&1
cat gpurun_out/gpu_tests.txt
