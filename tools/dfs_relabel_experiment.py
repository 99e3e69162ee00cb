"""Hypothesis check: are the hop gathers of the TAG kernels bound by shared-memory bank conflicts caused by the bus numbering?
Relabel the Oberrhein feeder's buses in DFS order (neighbours get adjacent numbers) and time the same train step.
usage (GPU box): python tools/dfs_relabel_experiment.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import numpy as np
import torch
from dss2 import synth
from dss2.trainer import GraphedTrainer, default_spec

REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}


def dfs_order(grid):
    bus, edge = grid["bus_param"], grid["edge_param"]
    n = bus.shape[0]
    adj = [[] for _ in range(n)]
    for e in edge:
        if e[6] == 1.0:
            adj[int(e[0])].append(int(e[1]))
            adj[int(e[1])].append(int(e[0]))
    root = int(np.nonzero(bus[:, 1] == 1.0)[0][0])
    seen, order, stack = [False] * n, [], [root]
    while stack:
        u = stack.pop()
        if seen[u]:
            continue
        seen[u] = True
        order.append(u)
        for v in reversed(adj[u]):
            if not seen[v]:
                stack.append(v)
    order += [u for u in range(n) if not seen[u]]
    return order


def relabel(grid, order):
    new_id = np.empty(len(order), dtype=np.int64)
    new_id[np.array(order)] = np.arange(len(order))
    g = dict(grid)
    g["bus_param"] = grid["bus_param"][np.array(order)]
    e = grid["edge_param"].copy()
    e[:, 0] = new_id[grid["edge_param"][:, 0].astype(np.int64)]
    e[:, 1] = new_id[grid["edge_param"][:, 1].astype(np.int64)]
    g["edge_param"] = e
    g["meas_v"] = new_id[grid["meas_v"]]
    return g


def run(name, grid, B=4096, S=8192, K=10):
    store = synth.synthetic_store(grid, S, seed=7, device="cuda")
    tr = GraphedTrainer(store, B, spec=default_spec(), reg_coefs=REG, seed=0, use_cuda_graph=True).capture()
    ids = torch.randint(0, S, (K + 3, B), generator=torch.Generator().manual_seed(1)).pin_memory()
    for i in range(3):
        tr.step(ids[i])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        tr.step(ids[3 + i])
    b.record()
    b.synchronize()
    print(json.dumps({"grid": name, "ms_per_step": a.elapsed_time(b) / K, "scenarios_per_s": B * K / a.elapsed_time(b) * 1e3}), flush=True)


grid = synth.load_grid("ober_sub")
run("ober_sub (reference bus numbering)", grid)
run("ober_sub relabelled in DFS order", relabel(grid, dfs_order(grid)))
