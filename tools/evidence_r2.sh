#!/bin/bash
# round-2 evidence: other bench configs, sanitizer over the new kernels, ncu captures of the backward-to-input and the loss
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
: > gpurun_out/r2_configs.jsonl
for cfg in "--config cigre14" "--config reswitched" "--config feeder10k" "--network gat" "--network gine" "--mode infer"; do
  timeout 600 python bench.py $cfg --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/r2_configs.jsonl 2> gpurun_out/r2_cfg.err || echo "FAILED $cfg: $(tail -n 2 gpurun_out/r2_cfg.err)"
done
python - <<'PY'
import json
for line in open("gpurun_out/r2_configs.jsonl"):
    d = json.loads(line)
    print(d["config"]["workload"][:90], "| value", round(d["value"]), "| ms", round(d["ms_per_step"], 3), "| e2e", round((d.get("e2e") or {}).get("value") or 0))
PY
# compute-sanitizer memcheck over the layer-kernel and trainer tests (the kernels touched in round 2), racecheck over the layer tests
( compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tag_fwd_tensor or tag_bwd_tensor or graphed_trainer or exact_weight or ragged or philox" 2>&1 | tail -8 ) > gpurun_out/r2_sanitizer.txt
( compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tag_bwd_tensor and ober_sub-7" 2>&1 | tail -8 ) >> gpurun_out/r2_sanitizer.txt
cat gpurun_out/r2_sanitizer.txt
for pat in "k_tag_tc3<1" k_wls; do
  name=$(echo $pat | tr -d '<')
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$pat" -c 2 -f -o gpurun_out/r2_$name \
      python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_$name.log 2>&1
  ls -la gpurun_out/r2_$name.ncu-rep
done
