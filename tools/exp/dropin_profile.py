import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch, data as d3, networks
from dss2 import batching, synth
REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
store = synth.synthetic_store(synth.load_grid("cigre14"), 128, seed=2)
graphs = [batching.Data(**{k: v.clone() for k, v in store.graph(i).items()}) for i in range(128)]
loader = batching.DataLoader(graphs, batch_size=64, shuffle=False)
st = [t.cuda() for t in (store.x_mean, store.x_std, store.edge_mean, store.edge_std)]
model = networks.SkipPFN(8, 6, 2, 32, 8, 2, 0.3, 5).cuda()
opt = torch.optim.Adamax(model.parameters(), lr=1e-3)
batches = [b.to("cuda") for b in loader]
def step(b):
    opt.zero_grad()
    out = model(b.x[:, :8], b.edge_index, b.edge_attr[:, :6])
    loss = d3.gsp_wls_edge(input=b.x[:, :8], edge_input=b.edge_attr[:, :6], output=out, x_mean=st[0], x_std=st[1], edge_mean=st[2], edge_std=st[3],
                           edge_index=b.edge_index, reg_coefs=REG, num_samples=b.num_graphs, node_param=b.x[:, 8:], edge_param=b.edge_attr[:, 6:])
    loss.backward(); opt.step(); return float(loss)
for i in range(6): step(batches[i % 2])
pr = cProfile.Profile(); pr.enable()
for i in range(20): step(batches[i % 2])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
