#!/bin/bash
# the driver's round-end sequence on one box: GPU tests, smoke, both bench arms (short)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/full_tests.txt 2>&1
cat gpurun_out/full_tests.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6 | tee gpurun_out/full_smoke.txt
( time timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/full_bench.json ) 2> gpurun_out/full_bench.err
tail -n 5 gpurun_out/full_bench.err
python tools/show_bench.py gpurun_out/full_bench.json
( time timeout 900 python bench.py --impl reference --steps ${REF_STEPS:-3} --warmup 1 > gpurun_out/full_bench_ref.json ) 2> gpurun_out/full_bench_ref.err
tail -n 4 gpurun_out/full_bench_ref.err
python - <<'PY'
import json
for f in ("gpurun_out/full_bench.json", "gpurun_out/full_bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), "cpu", d.get("cpu_baseline"))
    except Exception as exc:
        print(f, "unreadable", exc)
a = json.loads(open("gpurun_out/full_bench.json").read().strip().splitlines()[-1])["config"]
b = json.loads(open("gpurun_out/full_bench_ref.json").read().strip().splitlines()[-1])["config"]
print("same_config:", a == b)
PY
