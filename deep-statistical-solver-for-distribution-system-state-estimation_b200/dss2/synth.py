"""Grid definitions and the deterministic synthetic-scenario generator (host side, torch ops).

The reference ships solved scenarios only for CIGRE-14 (data/cigre14/{nodes,edges,labels}); the
scenario pickles of cigre14_reswitched and ober_sub are missing from the repository
(.MISSING_LARGE_BLOBS:1-6) and pandapower is not available, so the BASELINE configs on those grids
use self-consistent synthetic scenarios: draw a smooth voltage / angle profile on a
BFS tree from the slack bus, evaluate the reference's pi-model branch equations (data.py:370-376)
in float64 to get branch flows and bus injections, and hand the result to the same feature
pipeline as real data (`dss2.dataset.build_scenario_store`).  Arrays come out in the layout of the
reference pickles (NODE_COLS / EDGE_COLS) so both sources are interchangeable.
"""
import os

import numpy as np
import torch

from .dataset import build_scenario_store

_GRID_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grids")

# sensor placement of dss2_run.py:48-53 (bus ids / positions in the closed-edge list)
SENSORS = {
    "cigre": (np.array([0, 1, 12, 7, 11, 14]), np.array([0, 10])),
    "other": (np.array([35, 16, 52, 47, 6, 48, 59, 27, 37, 56]), np.array([40, 43, 11, 21, 54, 57])),
}


def load_grid(case):
    """case in {cigre14, cigre14_reswitched, ober_sub}: bus_param[N,3] (vn_kv, slack, zero_inj),
    edge_param[E_all,9] (from, to, G, B, Gs, Bs, closed, phase_shift, imax_or_sn), noise_param[6]."""
    z = np.load(os.path.join(_GRID_DIR, f"{case}.npz"))
    meas_v, meas_pflow = SENSORS["cigre" if "cigre" in case else "other"]
    return {"name": case, "bus_param": z["bus_param"], "edge_param": z["edge_param"],
            "noise_param": z["noise_param"], "meas_v": meas_v, "meas_pflow": meas_pflow}


def replicate_feeder(grid, copies):
    """Large radial test grid (BASELINE config 5): `copies` replicas of `grid`'s low-voltage feeder
    hung under ONE shared slack bus / transformer.  Bus 0 of the result is the shared slack; replica
    r maps original bus b != slack to 1 + r*(N-1) + rank(b)."""
    bus, edge = grid["bus_param"], grid["edge_param"]
    n = bus.shape[0]
    slack = int(np.nonzero(bus[:, 1] == 1.0)[0][0])
    others = [b for b in range(n) if b != slack]
    rank = {b: i for i, b in enumerate(others)}
    new_bus = [bus[slack]]
    new_edge = []
    for r in range(copies):
        base = 1 + r * (n - 1)
        for b in others:
            new_bus.append(bus[b])
        for e in edge:
            f, t = int(e[0]), int(e[1])
            row = e.copy()
            row[0] = 0 if f == slack else base + rank[f]
            row[1] = 0 if t == slack else base + rank[t]
            new_edge.append(row)
    return {"name": f"{grid['name']}_x{copies}", "bus_param": np.stack(new_bus), "edge_param": np.stack(new_edge),
            "noise_param": grid["noise_param"],
            "meas_v": np.concatenate([[0]] + [1 + r * (n - 1) + np.array([rank[b] for b in grid["meas_v"] if b != slack])
                                              for r in range(copies)]).astype(np.int64),
            "meas_pflow": grid["meas_pflow"]}


def _bfs_tree(num_nodes, frm, to, root):
    """Parent edge of every bus in a BFS from `root` over the closed edges: lists (order, parent, edge_id)."""
    adj = [[] for _ in range(num_nodes)]
    for e, (f, t) in enumerate(zip(frm, to)):
        adj[f].append((t, e))
        adj[t].append((f, e))
    seen = np.zeros(num_nodes, dtype=bool)
    seen[root] = True
    order, parent, via = [root], [-1], [-1]
    head = 0
    while head < len(order):
        u = order[head]
        head += 1
        for v, e in adj[u]:
            if not seen[v]:
                seen[v] = True
                order.append(v)
                parent.append(u)
                via.append(e)
    if not seen.all():
        raise ValueError("grid is not connected through its closed edges")
    return order, parent, via


def branch_flows(v, theta, frm, to, g, b, gs, bs, v_lv):
    """pi-model branch flows of data.py:370-376 (phase shift ignored as in data.py:362-363), float64,
    vectorised over scenarios: v, theta [S,N]; returns P_from, Q_from, P_to, Q_to [S,E]."""
    vi, vj = v[:, frm], v[:, to]
    d = theta[:, frm] - theta[:, to]
    c, s = torch.cos(d), torch.sin(d)
    base = v_lv ** 2
    p_f = (-vi * vj * (g * c + b * s) + (g + gs / 2) * vi ** 2) * base
    q_f = (vi * vj * (-g * s + b * c) - (b + bs / 2) * vi ** 2) * base
    p_t = (-vi * vj * (g * c - b * s) + (g + gs / 2) * vj ** 2) * base
    q_t = (vi * vj * (g * s + b * c) - (b + bs / 2) * vj ** 2) * base
    return p_f, q_f, p_t, q_t


def sample_scenarios(grid, num_scenarios, seed=1234, device="cpu", line_dv=(2e-4, 4e-4), line_dth=(2e-4, 3e-4),
                     trafo_dv=(2e-2, 5e-3), trafo_dth=(2e-2, 5e-3), v_slack=1.03):
    """Synthetic solved scenarios in pickle layout: nodes[S,N,7], edges[S,E_all,11], labels[S,N,2] (float64
    torch tensors on `device`).  Per tree edge the child bus is (mean, std)-normally below its parent in
    voltage magnitude and angle; open edges carry zero flow."""
    bus = torch.as_tensor(grid["bus_param"], dtype=torch.float64, device=device)
    edge = torch.as_tensor(grid["edge_param"], dtype=torch.float64, device=device)
    n = bus.shape[0]
    closed = grid["edge_param"][:, 6] == 1.0
    ce = grid["edge_param"][closed]
    frm_np, to_np = ce[:, 0].astype(np.int64), ce[:, 1].astype(np.int64)
    root = int(np.nonzero(grid["bus_param"][:, 1] == 1.0)[0][0])
    order, parent, via = _bfs_tree(n, frm_np, to_np, root)
    is_trafo = np.ceil(ce[:, 7]) > 0

    gen = torch.Generator(device=device).manual_seed(seed)
    S = num_scenarios
    dv = torch.randn(S, n, generator=gen, dtype=torch.float64, device=device)
    dth = torch.randn(S, n, generator=gen, dtype=torch.float64, device=device)
    v = torch.empty(S, n, dtype=torch.float64, device=device)
    th = torch.empty(S, n, dtype=torch.float64, device=device)
    v[:, root] = v_slack
    th[:, root] = 0.0
    for u, p, e in zip(order[1:], parent[1:], via[1:]):
        (mv, sv), (mt, st) = (trafo_dv, trafo_dth) if is_trafo[e] else (line_dv, line_dth)
        v[:, u] = v[:, p] - (mv + sv * dv[:, u])
        th[:, u] = th[:, p] - (mt + st * dth[:, u])

    cet = torch.as_tensor(ce, dtype=torch.float64, device=device)
    frm = torch.as_tensor(frm_np, device=device)
    to = torch.as_tensor(to_np, device=device)
    v_lv = bus[:, 0].min()
    p_f, q_f, p_t, q_t = branch_flows(v, th, frm, to, cet[:, 2], cet[:, 3], cet[:, 4], cet[:, 5], v_lv)
    # bus injections with the sign convention of data.py:428-429
    p_bus = torch.zeros(S, n, dtype=torch.float64, device=device)
    q_bus = torch.zeros(S, n, dtype=torch.float64, device=device)
    p_bus.index_add_(1, to, -p_t)
    p_bus.index_add_(1, frm, -p_f)
    q_bus.index_add_(1, to, -q_t)
    q_bus.index_add_(1, frm, -q_f)
    zero_inj = bus[:, 2] == 1.0
    p_bus[:, zero_inj] = 0.0      # declared zero-injection buses are reported as exactly zero
    q_bus[:, zero_inj] = 0.0

    nodes = torch.cat([bus.unsqueeze(0).expand(S, n, 3), v.unsqueeze(-1), th.unsqueeze(-1),
                       p_bus.unsqueeze(-1), q_bus.unsqueeze(-1)], dim=-1)
    e_all = edge.shape[0]
    flows = torch.zeros(S, e_all, 2, dtype=torch.float64, device=device)
    cidx = torch.as_tensor(np.nonzero(closed)[0], device=device)
    flows[:, cidx, 0] = p_f
    flows[:, cidx, 1] = q_f
    edges = torch.cat([edge.unsqueeze(0).expand(S, e_all, 9), flows], dim=-1)
    labels = torch.stack([v, th], dim=-1)
    return nodes, edges, labels


def synthetic_store(grid, num_scenarios, seed=1234, device="cpu"):
    """`sample_scenarios` + measurement noise from a torch generator + `build_scenario_store`."""
    nodes, edges, labels = sample_scenarios(grid, num_scenarios, seed, device)
    n = nodes.shape[1]
    e = int((grid["edge_param"][:, 6] == 1.0).sum())
    gen = torch.Generator(device=device).manual_seed(seed + 7919)
    zn = torch.randn(num_scenarios, n, 4, generator=gen, dtype=torch.float64, device=device)
    ze = torch.randn(num_scenarios, e, 2, generator=gen, dtype=torch.float64, device=device)
    return build_scenario_store(nodes, edges, labels, grid["noise_param"], grid["meas_v"], grid["meas_pflow"],
                                zn, ze, device=device)
