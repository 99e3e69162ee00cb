"""Train-step throughput of the as-shipped default model GAT_DSSE, or of GINE_DSSE (scope row 8f-1), through the drop-in modules on
CUDA tensors: forward, gsp_wls_edge, backward, torch.optim.Adamax.  usage (GPU box): python tools/gat_bench.py [B] [steps] [gat|gine]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import data as d3
import networks
from dss2 import _lib, batching, synth

REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
KIND = sys.argv[3] if len(sys.argv) > 3 else "gat"
store = synth.synthetic_store(synth.load_grid("ober_sub"), B, seed=3, device="cuda")
batch = batching.pack_batch(store, torch.arange(B, device="cuda"))
model = (networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=8, edge_dim=6) if KIND == "gat" else
         networks.GINE_DSSE(dim_feat=8, dim_dense=32, dim_out=2, num_layers=8, edge_dim=6)).cuda()
opt = torch.optim.Adamax(model.parameters(), lr=3e-3)
stats = [t.cuda() for t in (store.x_mean, store.x_std, store.edge_mean, store.edge_std)]


def step():
    opt.zero_grad(set_to_none=True)
    out = model(batch.x[:, :8], batch.edge_index, batch.edge_attr[:, :6])
    loss = d3.gsp_wls_edge(input=batch.x[:, :8], edge_input=batch.edge_attr[:, :6], output=out, x_mean=stats[0], x_std=stats[1],
                           edge_mean=stats[2], edge_std=stats[3], edge_index=batch.edge_index, reg_coefs=REG, num_samples=None,
                           node_param=batch.x[:, 8:], edge_param=batch.edge_attr[:, 6:])
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = _lib.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(K):
    loss = step()
b.record()
b.synchronize()
ms = a.elapsed_time(b) / K
print(json.dumps({"model": f"{type(model).__name__}(8,32,2,num_layers=8,edge_dim=6)", "workload": f"ober_sub B={B}", "ms_per_step": ms,
                  "scenarios_per_s": B / ms * 1e3, "kernel_launches_per_step": (_lib.launch_count() - n0) / K, "loss": float(loss)}))

# the same step on the host cores through the oracle (small sample)
import dss2_oracle as orc
torch.set_num_threads(os.cpu_count() or 1)
S = 128
cb = orc.collate([{k: v.cpu() for k, v in store.graph(i).items()} for i in range(S)])
p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.named_parameters()}
copt = torch.optim.Adamax(list(p.values()), lr=3e-3)
cst = [t.cpu() for t in stats]
t0 = None
for it in range(4):
    if it == 1:
        t0 = time.perf_counter()
    copt.zero_grad()
    fwd = orc.gat_dsse_forward if KIND == "gat" else orc.gine_dsse_forward
    o = fwd(p, cb["x"][:, :8], cb["edge_index"], cb["edge_attr"][:, :6], 8)
    l = orc.wls_loss(cb["x"], cb["edge_attr"], o, *cst, cb["edge_index"], REG)
    l.backward()
    copt.step()
dt = (time.perf_counter() - t0) / 3
print(json.dumps({"cpu_port": True, "cores": os.cpu_count(), "sample_scenarios": S, "ms_per_step": dt * 1e3, "scenarios_per_s": S / dt}))
