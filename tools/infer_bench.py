"""Inference throughput (forward + get_pflow only) of SkipPFN on the bench workload, device-resident batch, eager launches.
usage (GPU box): python tools/infer_bench.py [B] [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch
import data as d3
import networks
from dss2 import _lib, batching, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
store = synth.synthetic_store(synth.load_grid("ober_sub"), B, seed=3, device="cuda")
batch = batching.pack_batch(store, torch.arange(B, device="cuda"))
model = networks.SkipPFN(dim_featn=8, dim_feate=6, dim_out=2, dim_hid=32, n_gnn_layers=8, K=2, dropout_rate=0.3, L=5).cuda().eval()
xs, xm = store.x_std.cuda(), store.x_mean.cuda()


def step():
    with torch.no_grad():
        out = model(batch.x[:, :8], batch.edge_index, batch.edge_attr[:, :6])
        est = torch.cat([out[:, 0:1] * xs[:1] + xm[:1], out[:, 1:] * (1. - batch.x[:, 9:10])], 1)       # dss2_run.py:183-184
        return d3.get_pflow(est, batch.edge_index, node_param=batch.x[:, 8:], edge_param=batch.edge_attr[:, 6:])[0:2]


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = _lib.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(K):
    step()
b.record()
b.synchronize()
ms = a.elapsed_time(b) / K
print(json.dumps({"metric": "inference scenarios/s (SkipPFN forward + get_pflow)", "workload": f"ober_sub B={B}", "ms_per_batch": ms,
                  "scenarios_per_s": B / ms * 1e3, "kernel_launches_per_batch": (_lib.launch_count() - n0) / K}))
