#!/bin/bash
# 8-GPU: single vs bucketed all-reduce of the gradient (short runs)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${N:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/$2.json 2> gpurun_out/$2.err; }
DSS2_BUCKETED_ALLREDUCE=0 run 29531 b8_n${N}_single
DSS2_BUCKETED_ALLREDUCE=1 run 29532 b8_n${N}_bucketed
tail -n 2 gpurun_out/b8_*.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/b8_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d.get("n_gpus"), "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "dp", d.get("data_parallel_check"))
    except Exception as exc:
        print(f, "unreadable", exc)
PY
