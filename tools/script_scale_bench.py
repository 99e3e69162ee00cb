"""Per-step latency at the SCRIPT's own scale (dss2_run.py:34 batch_size = 64, CIGRE-14: 960 buses per batch, default GAT_DSSE and the
paper's SkipPFN) through the drop-in modules, the way the unmodified script calls them: (a) CPU tensors and CPU parameters (the script
never moves anything to a device: the drop-in stages per call), (b) CUDA tensors; and (c) the reference's own modules on the host cores
(oracle/_ref over the shim) for the same loop.  usage (GPU box): python tools/script_scale_bench.py [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200")
sys.path.insert(0, PKG)
import numpy as np
import torch
import data as d3
import networks
from dss2 import batching, dataset, synth

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
grid = synth.load_grid("cigre14")
store = synth.synthetic_store(grid, 128, seed=2)
graphs = [batching.Data(**{k: v.clone() for k, v in store.graph(i).items()}) for i in range(128)]
stats = [store.x_mean, store.x_std, store.edge_mean, store.edge_std]


def loop(model, loader, dev, loss_fn, steps):
    opt = torch.optim.Adamax(model.parameters(), lr=1e-3)
    st = [t.to(dev) for t in stats]
    it, t0, n = iter(loader), None, 0
    for s in range(steps + 5):
        try:
            b = next(it)
        except StopIteration:
            it = iter(loader)
            b = next(it)
        if s == 5:
            if dev != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
        b = b.to(dev) if dev != "cpu" else b
        opt.zero_grad()
        out = model(b.x[:, :8], b.edge_index, b.edge_attr[:, :6])
        loss = loss_fn(input=b.x[:, :8], edge_input=b.edge_attr[:, :6], output=out, x_mean=st[0], x_std=st[1], edge_mean=st[2], edge_std=st[3],
                       edge_index=b.edge_index, reg_coefs=REG, num_samples=b.num_graphs, node_param=b.x[:, 8:], edge_param=b.edge_attr[:, 6:])
        loss.backward()
        opt.step()
        float(loss)       # the script accumulates loss.item() (dss2_run.py:145)
    if dev != "cpu":
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


res = {}
for name, ctor in (("GAT_DSSE", lambda m: m.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=8, edge_dim=6)),
                   ("SkipPFN", lambda m: m.SkipPFN(8, 6, 2, 32, 8, 2, 0.3, 5))):
    loader = batching.DataLoader(graphs, batch_size=64, shuffle=False)
    res[name + " drop-in, CPU tensors (as the script)"] = loop(ctor(networks), loader, "cpu", d3.gsp_wls_edge, K)
    res[name + " drop-in, CUDA tensors"] = loop(ctor(networks).cuda(), loader, "cuda", d3.gsp_wls_edge, K)
ref_dir = os.path.join(ROOT, "oracle", "_ref")
if os.path.isdir(ref_dir):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_runner
    ref_networks, ref_data = ref_runner.load_reference(stub_laplacian=False)
    from torch_geometric.data import Data as PygData
    from torch_geometric.loader import DataLoader as PygLoader
    rgraphs = [PygData(**{k: v.clone() for k, v in store.graph(i).items()}) for i in range(128)]
    for name, ctor in (("GAT_DSSE", lambda m: m.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=8, edge_dim=6)),
                       ("SkipPFN", lambda m: m.SkipPFN(8, 6, 2, 32, 8, 2, 0.3, 5))):
        res[name + f" reference's own modules, {torch.get_num_threads()} host threads"] = loop(ctor(ref_networks), PygLoader(rgraphs, batch_size=64), "cpu",
                                                                                             ref_data.gsp_wls_edge, max(3, K // 6))
for k, v in res.items():
    print(f"{v:9.2f} ms/step  {k}")
print(json.dumps({"batch_size": 64, "grid": "cigre14", "ms_per_step": res}))
