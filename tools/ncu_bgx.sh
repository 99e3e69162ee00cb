#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_tag_tc3<1" -c 2 -f -o gpurun_out/r2_k_tag_tc3_bgx \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_bgx.log 2>&1
ls -la gpurun_out/r2_k_tag_tc3_bgx.ncu-rep
python bench.py --network gat --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('gat value', d['value'], 'e2e', d['e2e'])
    except Exception: pass
"
python - <<'PY'
# where does the GAT e2e step go?
import sys, os, time, torch
sys.path.insert(0, "deep-statistical-solver-for-distribution-system-state-estimation_b200")
import networks, data as d3
from dss2 import batching, synth
store = synth.synthetic_store(synth.load_grid("ober_sub"), 4096, seed=1234, device="cuda")
b = batching.pack_batch(store, torch.arange(4096, device="cuda"))
m = networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=8, edge_dim=6).cuda()
hx = b.x.cpu().pin_memory()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b.x.copy_(hx, non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = m(b.x[:, :8], b.edge_index, b.edge_attr[:, :6])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("copy %.2f ms  fwd %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), "edge_index version", b.edge_index._version, hasattr(b.edge_index, "_dss2_graph"))
PY
