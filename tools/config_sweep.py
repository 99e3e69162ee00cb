"""Train-step throughput of the other BASELINE.json configs through the same GraphedTrainer the headline bench uses.
One JSON line per config (not the headline metric: that is bench.py).  usage (GPU box): python tools/config_sweep.py [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch
from dss2 import synth
from dss2.trainer import GraphedTrainer, default_spec

REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
CONFIGS = [
    ("configs[0] cigre14 (N=15) B=4096", lambda: synth.load_grid("cigre14"), 8192, 4096),
    ("configs[1] cigre14_reswitched (N=15) B=4096", lambda: synth.load_grid("cigre14_reswitched"), 8192, 4096),
    ("configs[2] ober_sub (N=70) B=4096", lambda: synth.load_grid("ober_sub"), 8192, 4096),
    ("configs[4] 10k-bus radial feeder (ober_sub x143 under one slack), B=32, large-graph path", lambda: synth.replicate_feeder(synth.load_grid("ober_sub"), 143), 64, 32),
    ("configs[4b] 346-bus feeder (ober_sub x5), B=1024, large-graph path", lambda: synth.replicate_feeder(synth.load_grid("ober_sub"), 5), 2048, 1024),
]
for name, mk, S, B in CONFIGS:
    grid = mk()
    store = synth.synthetic_store(grid, S, seed=7, device="cuda")
    tr = GraphedTrainer(store, B, spec=default_spec(), reg_coefs=REG, seed=0, use_cuda_graph=True).capture()
    gen = torch.Generator().manual_seed(1)
    ids = torch.randint(0, S, (K + 3, B), generator=gen).pin_memory()
    for i in range(3):
        tr.step(ids[i])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        tr.step(ids[3 + i])
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / K
    print(json.dumps({"config": name, "nodes_per_step": tr.nt, "edges_per_step": tr.et, "tiled": bool(tr.graph.c.num_tiles > 0),
                      "ms_per_step": ms, "scenarios_per_s": B / ms * 1e3, "bus_rows_per_s": tr.nt / ms * 1e3,
                      "launches_per_step": tr.launches_per_step, "loss": float(tr.loss)}), flush=True)
    del tr, store
    torch.cuda.empty_cache()
