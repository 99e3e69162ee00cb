"""Scenario store and dataset builder (host side, torch tensor ops; device agnostic).

`build_scenario_store` restates the feature engineering of the reference `data_from_pickles`
(reference data.py:96-206) as vectorised tensor ops over all scenarios at once instead of the
reference's per-scenario python loop with O(n^2) `torch.cat` (data.py:174-176).  It produces the
same per-graph layout:

    x[N,11]          = (V, wV, theta, wTheta, P, wP, Q, wQ | vn_kv, bool_slack, bool_zero_inj)
    edge_attr[E,13]  = (Pf, wPf, Qf, wQf, G~, B~ | G, B, Gs, Bs, closed, phase_shift, imax_or_sn)
    edge_index[2,E]  = closed edges only, from -> to, int64, local bus numbering
    y[N,2]           = (vm_pu, va_rad)

with the first 8 / 6 columns z-scored over the non-zero entries of the whole set (data.py:179-190).
The measurement noise is passed in as standard-normal draws so that a caller can replay the
reference's `np.random` stream exactly (`reference_noise_stream`) or use a device generator.

Scenarios are kept scenario-major in one `ScenarioStore`; `dss2.batching.pack_batch` gathers a list
of scenario ids into a PyG-ordered disjoint-union batch with the CUDA packer.
"""
from dataclasses import dataclass

import numpy as np
import torch

NODE_COLS = ("vn_kv", "bool_slack", "bool_zero_inj", "vm_pu", "va_rad", "p_mw", "q_mvar")
EDGE_COLS = ("from_bus", "to_bus", "G", "B", "Gs", "Bs", "closed line", "phase shift", "imax or sn",
             "p_from_mw", "q_from_mvar")
NOISE_COLS = ("p_noise", "v_noise", "i_noise", "pm_noise", "sgen_noise", "zero_inj_coef")


@dataclass
class ScenarioStore:
    """Scenario-major storage.  Scenario s owns node rows node_off[s]:node_off[s+1] of `x`/`y` and
    edge rows edge_off[s]:edge_off[s+1] of `edge_attr` / columns of `edge_index` (local numbering).
    Uniform grids have node_off = arange(S+1)*N; ragged stores are allowed by the packer."""
    x: torch.Tensor            # [sum N, 11] f32
    edge_attr: torch.Tensor    # [sum E, 13] f32
    y: torch.Tensor            # [sum N, 2]  f32
    edge_index: torch.Tensor   # [2, sum E]  i64, local bus ids
    node_off: torch.Tensor     # [S+1] i64
    edge_off: torch.Tensor     # [S+1] i64
    x_mean: torch.Tensor       # [8]
    x_std: torch.Tensor        # [8]
    edge_mean: torch.Tensor    # [6]
    edge_std: torch.Tensor     # [6]
    max_nodes: int = 0         # largest graph (host-known, sizes the kernels' tiles)
    max_edges: int = 0

    @property
    def num_scenarios(self):
        return self.node_off.numel() - 1

    def to(self, device):
        kw = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()}
        return ScenarioStore(**kw)

    def graph(self, s):
        """Scenario s as a dict of tensors (views), the shape of one PyG `Data` object."""
        n0, n1 = int(self.node_off[s]), int(self.node_off[s + 1])
        e0, e1 = int(self.edge_off[s]), int(self.edge_off[s + 1])
        return {"x": self.x[n0:n1], "edge_attr": self.edge_attr[e0:e1], "y": self.y[n0:n1],
                "edge_index": self.edge_index[:, e0:e1]}


def reference_noise_stream(seed, num_scenarios, num_nodes, num_closed_edges):
    """Standard-normal draws in the order the reference consumes `np.random.normal`:
    per scenario one [N,4] block (data.py:131) then one [E,2] block (data.py:159).
    `np.random.normal(0, s)` is `0 + s * gauss()` on the legacy stream, so scaling these draws by
    |std| reproduces the reference's noise bit for bit after `np.random.seed(seed)`."""
    rs = np.random.RandomState(seed)
    zn = np.empty((num_scenarios, num_nodes, 4))
    ze = np.empty((num_scenarios, num_closed_edges, 2))
    for s in range(num_scenarios):
        zn[s] = rs.standard_normal((num_nodes, 4))
        ze[s] = rs.standard_normal((num_closed_edges, 2))
    return zn, ze


def _masked_zscore(t, num_feat):
    """data.py:179-190: mean/std over the non-zero entries of every column, z-score the masked
    entries, then restore the raw parameter columns (num_feat:)."""
    mask = t != 0.0
    cnt = mask.sum(dim=[0])
    mean = torch.nan_to_num((t * mask).sum(dim=[0]) / cnt)
    std = torch.nan_to_num(torch.sqrt((((t - mean) ** 2) * mask).sum(dim=[0]) / cnt))
    out = torch.nan_to_num((t - mean) * mask / std)
    out[:, num_feat:] = t[:, num_feat:]
    return out, mean, std


def _build_on_device(nodes, ce, labels, noise_param, meas_v, meas_pflow, zn, ze, device):
    """The CUDA builder (csrc/dataset.cu): one launch sequence over all scenarios; no host fallback."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    S, N, E = nodes.shape[0], nodes.shape[1], ce.shape[1]
    mv = torch.zeros(N, dtype=torch.uint8, device=device)
    mv[torch.as_tensor(np.asarray(meas_v), dtype=torch.long, device=device)] = 1
    mp = torch.zeros(max(E, 1), dtype=torch.uint8, device=device)
    mp[torch.as_tensor(np.asarray(meas_pflow), dtype=torch.long, device=device)] = 1
    x = torch.empty(S * N, 11, dtype=torch.float32, device=device)
    ea = torch.empty(S * E, 13, dtype=torch.float32, device=device)
    stats = torch.empty(28, dtype=torch.float32, device=device)
    ws = torch.empty(lib.dss2_build_scenarios_workspace_bytes(), dtype=torch.uint8, device=device)
    npar = (ctypes.c_double * 6)(*[float(v) for v in noise_param])
    nodes, ce11, zn, ze = nodes.contiguous(), ce[:, :, :11].contiguous(), zn.contiguous(), ze.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(lib.dss2_build_scenarios(_lib.ptr(nodes), _lib.ptr(ce11), _lib.ptr(zn), _lib.ptr(ze), _lib.ptr(mv), _lib.ptr(mp), npar, S, N, E,
                                            _lib.ptr(x), _lib.ptr(ea), _lib.ptr(stats), _lib.ptr(ws), ws.numel(), _lib.stream()), "dss2_build_scenarios")
    return x, ea, stats[0:8], stats[14:22], stats[8:14], stats[22:28]


def build_scenario_store(nodes, edges, labels, noise_param, meas_v, meas_pflow, noise_nodes, noise_edges,
                         num_nfeat=8, num_efeat=6, device="cpu", impl=None):
    """nodes[S,N,7] (NODE_COLS), edges[S,E_all,>=11] (EDGE_COLS first), labels[S,N,2]: float64 arrays
    shaped like the reference pickles; noise_param: the 6 NOISE_COLS values; meas_v / meas_pflow:
    measured bus ids / measured closed-edge positions (dss2_run.py:48-53); noise_nodes[S,N,4],
    noise_edges[S,E,2]: standard-normal draws (float64).  All scenarios must share one switching
    state (true for every grid of the reference).  On a CUDA device the hand-written builder kernels run
    (impl="kernel", csrc/dataset.cu); impl="torch" keeps the tensor-op restatement below (the CPU path, bit-exact
    with the reference's np.random stream, and the measured alternative on the device)."""
    f64 = torch.float64
    nodes = torch.as_tensor(nodes, dtype=f64, device=device)
    edges = torch.as_tensor(edges, dtype=f64, device=device)
    labels = torch.as_tensor(labels, dtype=f64, device=device)
    zn = torch.as_tensor(noise_nodes, dtype=f64, device=device)
    ze = torch.as_tensor(noise_edges, dtype=f64, device=device)
    S, N = nodes.shape[0], nodes.shape[1]
    npar = dict(zip(NOISE_COLS, [float(v) for v in noise_param]))

    closed = edges[0, :, 6] == 1.0
    if not bool((edges[:, :, 6] == edges[0, :, 6]).all()):
        raise ValueError("build_scenario_store: scenarios with different switching states are not supported")
    ce = edges[:, closed, :]                                   # data.py:144
    E = ce.shape[1]
    if impl is None:
        impl = "kernel" if torch.device(device).type == "cuda" else "torch"
    if impl == "kernel":
        if num_nfeat != 8 or num_efeat != 6:
            raise ValueError("build_scenario_store(impl='kernel'): the reference's 8 node / 6 edge feature columns")
        x_set, ea_set, x_mean, x_std, e_mean, e_std = _build_on_device(nodes, ce, labels, noise_param, meas_v, meas_pflow, zn, ze, device)
        ei_local = ce[0, :, 0:2].to(torch.long).t().contiguous()
        return ScenarioStore(
            x=x_set, edge_attr=ea_set, y=labels.to(torch.float32).reshape(S * N, 2).contiguous(),
            edge_index=ei_local.repeat(1, S).contiguous(),
            node_off=torch.arange(S + 1, dtype=torch.long, device=device) * N,
            edge_off=torch.arange(S + 1, dtype=torch.long, device=device) * E,
            x_mean=x_mean.clone(), x_std=x_std.clone(), edge_mean=e_mean.clone(), edge_std=e_std.clone(), max_nodes=N, max_edges=E)

    # ---- buses (data.py:121-141) ----
    slack = nodes[:, :, 1:2]
    zero_inj = nodes[:, :, 2:3]
    mask = torch.zeros(N, 4, dtype=f64, device=device)
    mask[:, 2:] = 1.0
    mask[torch.as_tensor(np.asarray(meas_v), dtype=torch.long, device=device), 0] = 1.0
    mean = nodes[:, :, 3:7] * mask
    slack_noise = torch.tensor([npar["v_noise"], npar["zero_inj_coef"], npar["p_noise"], npar["p_noise"]], dtype=f64, device=device)
    node_noise = torch.tensor([npar["v_noise"], npar["v_noise"], npar["pm_noise"], npar["pm_noise"]], dtype=f64, device=device)
    std = mean * (slack_noise * slack + node_noise * (1 - slack))
    xv = (mean + std.abs() * zn).to(torch.float32)             # data.py:131
    std = std.clone()
    std[:, :, 2:] += npar["zero_inj_coef"] * zero_inj          # data.py:133
    std[:, :, 1:2] += slack_noise[1] * slack                   # data.py:135
    cov = 1.0 / torch.maximum(std.to(torch.float32).abs(), torch.tensor(1e-6, dtype=torch.float32, device=device)) ** 2
    cov = cov * (cov < 1e12).to(torch.float32)                 # data.py:137-138
    node_par = nodes[:, :, 0:3].to(torch.float32)
    x_all = torch.stack([xv[..., 0], cov[..., 0], xv[..., 1], cov[..., 1], xv[..., 2], cov[..., 2],
                         xv[..., 3], cov[..., 3]], dim=-1)
    x_all = torch.cat([x_all, node_par], dim=-1).reshape(S * N, 11)

    # ---- closed edges (data.py:144-172) ----
    emask = torch.zeros(E, 2, dtype=f64, device=device)
    emask[torch.as_tensor(np.asarray(meas_pflow), dtype=torch.long, device=device)] = 1.0
    emean = ce[:, :, 9:11] * emask
    estd = emean * npar["p_noise"]
    ev = (emean + estd.abs() * ze).to(torch.float32)           # data.py:159
    ecov = 1.0 / torch.maximum(estd.to(torch.float32).abs(), torch.tensor(1e-5, dtype=torch.float32, device=device)) ** 2
    ecov = ecov * (ecov < 1e10).to(torch.float32)              # data.py:161-162
    imp = ce[:, :, 2:4].to(torch.float32)
    epar = ce[:, :, 2:9].to(torch.float32)
    ea_all = torch.stack([ev[..., 0], ecov[..., 0], ev[..., 1], ecov[..., 1], imp[..., 0], imp[..., 1]], dim=-1)
    ea_all = torch.cat([ea_all, epar], dim=-1).reshape(S * E, 13)

    x_set, x_mean, x_std = _masked_zscore(x_all, num_nfeat)
    ea_set, e_mean, e_std = _masked_zscore(ea_all, num_efeat)

    ei_local = ce[0, :, 0:2].to(torch.long).t().contiguous()   # data.py:153
    return ScenarioStore(
        x=x_set.contiguous(), edge_attr=ea_set.contiguous(),
        y=labels.to(torch.float32).reshape(S * N, 2).contiguous(),
        edge_index=ei_local.repeat(1, S).contiguous(),
        node_off=torch.arange(S + 1, dtype=torch.long, device=device) * N,
        edge_off=torch.arange(S + 1, dtype=torch.long, device=device) * E,
        x_mean=x_mean[:num_nfeat].clone(), x_std=x_std[:num_nfeat].clone(),
        edge_mean=e_mean[:num_efeat].clone(), edge_std=e_std[:num_efeat].clone(),
        max_nodes=N, max_edges=E,
    )
