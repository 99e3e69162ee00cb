#!/bin/bash
# 2-GPU: bucketed (per sub-net, under the backward) vs single all-reduce of the gradient; same-shards trajectory hash == single GPU
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${N:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; }
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bk_n1.json 2> gpurun_out/bk_n1.err
run 29521 bk_n${N}_bucketed --no-e2e
DSS2_BUCKETED_ALLREDUCE=0 run 29522 bk_n${N}_single --no-e2e
run 29523 bk_n${N}_bucketed_same --no-e2e --same-shards --steps 20
run 29524 bk_n${N}_e2e
tail -n 2 gpurun_out/bk_*.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bk_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d.get("n_gpus"), "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", (d.get("e2e") or {}).get("value"), "sha", d.get("loss_trajectory_sha256_16"), "dp", d.get("data_parallel_check"))
    except Exception as exc:
        print(f, "unreadable", exc)
PY
