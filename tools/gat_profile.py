import os, sys
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch, data as d3, networks
from dss2 import batching, synth
REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
B = 4096
store = synth.synthetic_store(synth.load_grid("ober_sub"), B, seed=3, device="cuda")
batch = batching.pack_batch(store, torch.arange(B, device="cuda"))
model = networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=8, edge_dim=6).cuda()
stats = [t.cuda() for t in (store.x_mean, store.x_std, store.edge_mean, store.edge_std)]
def step():
    model.zero_grad(set_to_none=True)
    out = model(batch.x[:, :8], batch.edge_index, batch.edge_attr[:, :6])
    loss = d3.gsp_wls_edge(input=batch.x[:, :8], edge_input=batch.edge_attr[:, :6], output=out, x_mean=stats[0], x_std=stats[1], edge_mean=stats[2], edge_std=stats[3], edge_index=batch.edge_index, reg_coefs=REG, num_samples=None, node_param=batch.x[:, 8:], edge_param=batch.edge_attr[:, 6:])
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize(); torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
