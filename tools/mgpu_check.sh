#!/bin/bash
# 2-GPU checks: weak scaling line, replicas in sync, same-shards loss trajectory == single GPU, exact-global-batch mode
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${N:-2}
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/mg_n1.json 2> gpurun_out/mg_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg_n$N.json 2> gpurun_out/mg_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --same-shards --no-e2e > gpurun_out/mg_n${N}_same.json 2> gpurun_out/mg_n${N}_same.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 --exact-global-batch --no-e2e > gpurun_out/mg_n${N}_exact.json 2> gpurun_out/mg_n${N}_exact.err
if [ -z "$SKIP_REF" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/mg_n${N}_ref.json 2> gpurun_out/mg_n${N}_ref.err
fi
tail -n 3 gpurun_out/mg_*.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/mg_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d.get("n_gpus"), "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", (d.get("e2e") or {}).get("value"), "sha", d.get("loss_trajectory_sha256_16"), "dp", d.get("data_parallel_check"), d.get("data_parallel_mode"))
    except Exception as exc:
        print(f, "unreadable", exc)
PY
