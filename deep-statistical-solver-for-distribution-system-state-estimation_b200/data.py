"""Drop-in for the reference's `data` module (reference data.py), B200-native underneath.

  get_pflow        data.py:328-390   AC branch-flow equations               -> kernel `dss2_pflow`
  gsp_wls_edge     data.py:393-459   physics-informed WLS loss (+ autograd) -> fused fwd+bwd kernel `dss2_wls_fwd_bwd`
  data_from_pickles data.py:96-206   dataset construction                   -> vectorised builder (dss2.dataset)
  gsp_wls          data.py:462-522   older loss, broken in the reference (stale get_pflow call, :484): name only

Same signatures and return conventions.  Reference side effects kept: `gsp_wls_edge` zeroes the slack-bus angle
inside the caller's `output` tensor (data.py:412-413); `num_samples`, `mu_v`, `mu_theta` are accepted and ignored
(SURVEY.md appendix A).  The dead O(N^2) dense Laplacian of data.py:422-423 does not exist here.
No CPU fallback: without a CUDA device or the built library these functions raise.
"""
import pickle

import numpy as np
import torch

from dss2 import _lib, ops
from dss2.batching import Data
from dss2.dataset import EDGE_COLS, NODE_COLS, NOISE_COLS, build_scenario_store


def _joined(a, b):
    """[a | b] as one strided CUDA tensor: zero-copy when both are column slices of the same rows
    (dss2_run.py passes data.x[:, :8] and data.x[:, 8:]), else a concatenation."""
    if (a.device.type == "cuda" and a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
            and a.stride(1) == 1 and b.stride(1) == 1 and a.stride(0) == b.stride(0) and a.stride(0) >= a.size(1) + b.size(1)
            and b.data_ptr() == a.data_ptr() + 4 * a.size(1) and a.device == b.device):
        return a.as_strided((a.size(0), a.size(1) + b.size(1)), (a.stride(0), 1)), a.stride(0)
    t = torch.cat([a.float(), b.float()], dim=1)
    if t.device.type != "cuda":
        t = t.cuda(non_blocking=True)
    return t, t.stride(0)


def _vminmax(x_like, stride, col, n):
    out = torch.empty(2, dtype=torch.float32, device=x_like.device)
    _lib.check(_lib.load().dss2_col_minmax(_lib.ptr(x_like), stride, col, n, _lib.ptr(out), _lib.stream()), "dss2_col_minmax")
    return out


class _WlsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, xfull, xs, eafull, eas, edge_index, stats, coefs):
        ops.require_cuda()
        lib = _lib.load()
        dev = xfull.device
        nt = xfull.size(0)
        graph = ops.resolve_graph(edge_index, nt)
        if graph.c.undirected != 1:
            raise _lib.Dss2Error("gsp_wls_edge expects the one-way edge list of the reference's data (from_bus -> to_bus)")
        staged = output.device.type != "cuda" or not output.is_contiguous() or output.dtype != torch.float32
        out_c = output.detach().to(device=dev, dtype=torch.float32).contiguous() if staged else output
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grad = torch.empty(nt, 2, dtype=torch.float32, device=dev) if need_grad else None
        ws = graph.wls_workspace()
        with torch.cuda.device(dev):
            vmm = _vminmax(xfull, xs, 8, nt)
            rc = lib.dss2_wls_fwd_bwd(graph.ref, _lib.ptr(xfull), xs, _lib.ptr(eafull), eas, _lib.ptr(out_c), _lib.ptr(stats),
                                      coefs[0], coefs[1], coefs[2], coefs[3], _lib.ptr(vmm), 1, _lib.ptr(loss), None,
                                      _lib.ptr(grad), _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "dss2_wls_fwd_bwd")
        if staged:   # reproduce the in-place slack masking on the caller's tensor (data.py:412-413)
            output[:, 1].copy_(out_c[:, 1])
        ctx.mark_dirty(output)
        ctx.set_materialize_grads(False)     # an unused `output` result arrives as None in backward instead of a zero tensor
        ctx.grad = grad
        ctx.slack_keep = None
        if need_grad:
            ctx.slack_keep = (1.0 - xfull[:, 9]).detach()
        ctx.out_device = output.device
        return (loss if output.device.type == "cuda" else loss.to(output.device)), output

    @staticmethod
    def backward(ctx, g_loss, g_output):
        if g_loss is None:
            g = torch.zeros_like(ctx.grad)
        else:
            g = ctx.grad * g_loss.to(ctx.grad.device)
        if g_output is not None:   # gradient arriving through later uses of the (masked) output tensor
            go = g_output.to(g.device)
            g = g + torch.stack([go[:, 0], go[:, 1] * ctx.slack_keep], dim=1)
        return g.to(ctx.out_device), None, None, None, None, None, None, None


def gsp_wls_edge(input, edge_input, output, x_mean, x_std, edge_mean, edge_std, edge_index, reg_coefs, num_samples,
                 node_param, edge_param):
    """Physics-informed WLS loss of data.py:393-459; returns a 0-dim tensor differentiable w.r.t. `output`."""
    ops.require_cuda()
    xfull, xs = _joined(input, node_param)
    eafull, eas = _joined(edge_input, edge_param)
    if input.size(1) != 8 or node_param.size(1) != 3:
        raise _lib.Dss2Error("gsp_wls_edge expects 8 bus features + 3 bus parameters (data.py layout)")
    if edge_input.size(1) != 6 or edge_param.size(1) != 7:
        raise _lib.Dss2Error("gsp_wls_edge expects 6 branch features + 7 branch parameters (data.py layout)")
    stats = torch.cat([torch.as_tensor(t).float().reshape(-1).to(xfull.device) for t in (x_mean, x_std, edge_mean, edge_std)])
    coefs = (float(reg_coefs["lam_v"]), float(reg_coefs["lam_p"]), float(reg_coefs["lam_pf"]), float(reg_coefs["lam_reg"]))
    loss, _ = _WlsFunction.apply(output, xfull, xs, eafull, eas, edge_index, stats, coefs)
    return loss


class _PflowFunction(torch.autograd.Function):
    """get_pflow with its adjoint w.r.t. y (the reference's is plain autograd code, data.py:328-390)."""

    @staticmethod
    def forward(ctx, y, edge_index, node_param, edge_param, use_shift):
        lib = _lib.load()
        out_device = y.device
        yg, ys = ops.stage_rows(y.detach())
        npg, nps = ops.stage_rows(node_param)
        epg, eps = ops.stage_rows(edge_param)
        ei = edge_index if edge_index.device.type == "cuda" else edge_index.cuda(non_blocking=True)
        ei = ei.long().contiguous()
        et = ei.size(1)
        out8 = torch.empty(8, et, dtype=torch.float32, device=yg.device)
        with torch.cuda.device(yg.device):
            vmm = _vminmax(npg, nps, 0, npg.size(0))
            _lib.check(lib.dss2_pflow_ex(_lib.ptr(ei), et, _lib.ptr(yg), ys, _lib.ptr(epg), eps, _lib.ptr(vmm), int(use_shift), _lib.ptr(out8),
                                         _lib.stream()), "dss2_pflow")
        if ctx.needs_input_grad[0]:
            ctx.saved = (yg, ys, epg, eps, vmm, edge_index, int(use_shift), out_device, et)
        ctx.set_materialize_grads(False)
        if out_device.type != "cuda":
            out8 = out8.to(out_device)
        return tuple(out8[i] for i in range(8))

    @staticmethod
    def backward(ctx, *gouts):
        lib = _lib.load()
        yg, ys, epg, eps, vmm, edge_index, use_shift, out_device, et = ctx.saved
        graph = ops.resolve_graph(edge_index, yg.size(0))
        if graph.c.undirected != 1:
            raise _lib.Dss2Error("get_pflow backward expects the one-way edge list of the reference's data (from_bus -> to_bus)")
        g8 = torch.zeros(8, et, dtype=torch.float32, device=yg.device)
        for i, go in enumerate(gouts):
            if go is not None:
                g8[i].copy_(go.to(device=yg.device, dtype=torch.float32))
        gy = torch.empty(yg.size(0), 2, dtype=torch.float32, device=yg.device)
        with torch.cuda.device(yg.device):
            _lib.check(lib.dss2_pflow_bwd(graph.ref, _lib.ptr(yg), ys, _lib.ptr(epg), eps, _lib.ptr(vmm), use_shift, _lib.ptr(g8), _lib.ptr(gy),
                                          _lib.stream()), "dss2_pflow_bwd")
        return gy.to(out_device), None, None, None, None


def get_pflow(y, edge_index, node_param, edge_param, phase_shift=True):
    """data.py:328-390.  Returns (loading_lines, loading_trafo, P_from, Q_from, P_to, Q_to, I_from, I_to), each [Et], on the device of
    `y`, differentiable w.r.t. `y` like the reference's (one fused adjoint kernel; inside `gsp_wls_edge` the gradient is part of the
    loss kernel instead).  phase_shift=False subtracts the branch's phase shift from the angle difference (data.py:364-365)."""
    ops.require_cuda()
    if y.size(1) != 2:
        raise _lib.Dss2Error("get_pflow expects y = [V pu, theta rad] per bus")
    return _PflowFunction.apply(y, edge_index, node_param, edge_param, not phase_shift)


def gsp_wls(*args, **kwargs):
    raise NotImplementedError("gsp_wls is dead code in the reference (data.py:484 calls get_pflow with a stale signature); "
                              "use gsp_wls_edge")


def data_from_pickles(folder, num_nfeat, num_efeat, num_nmeas, num_emeas, meas_v, meas_pflow):
    """data.py:96-206: (list of graphs, x_mean[8], x_std[8], edge_mean[6], edge_std[6]) from the pickled pandas frames
    `folder + {nodes, edges, labels, noise_param}`.  Measurement noise is drawn from `np.random` in the reference's
    order, so `np.random.seed(s)` reproduces the reference's dataset bit for bit."""
    if num_nmeas != 4 or num_emeas != 2:
        raise ValueError("DSS2 data layout has 4 bus and 2 branch measurement types (dss2_run.py:44-45)")
    frames = {}
    for name in ("nodes", "edges", "labels", "noise_param"):
        with open(folder + name, "rb") as fh:
            frames[name] = pickle.load(fh)
    nodes = np.stack([df[list(NODE_COLS)].values for df in frames["nodes"]]).astype(np.float64)
    edges = np.stack([df[list(EDGE_COLS)].values for df in frames["edges"]]).astype(np.float64)
    labels = np.stack([df.values for df in frames["labels"]]).astype(np.float64)
    noise = frames["noise_param"][list(NOISE_COLS)].values.astype(np.float64).reshape(-1)
    S, N = nodes.shape[0], nodes.shape[1]
    E = int((edges[0, :, 6] == 1.0).sum())
    zn, ze = np.empty((S, N, 4)), np.empty((S, E, 2))
    for s in range(S):   # same consumption order as data.py:131 and :159
        zn[s] = np.random.standard_normal((N, 4))
        ze[s] = np.random.standard_normal((E, 2))
    st = build_scenario_store(nodes, edges, labels, noise, meas_v, meas_pflow, zn, ze, num_nfeat, num_efeat)
    cls = _graph_class()
    graphs = []
    for s in range(S):
        g = st.graph(s)
        d = cls(x=g["x"], edge_index=g["edge_index"], edge_attr=g["edge_attr"], y=g["y"])
        d.validate(raise_on_error=True)
        graphs.append(d)
    return graphs, st.x_mean, st.x_std, st.edge_mean, st.edge_std


def _graph_class():
    """The reference returns `torch_geometric.data.Data` objects (data.py:198) and the unmodified script batches them with
    `torch_geometric.loader.DataLoader` (dss2_run.py:18,68-69), whose collater only accepts PyG's own `BaseData` types.  So when
    torch_geometric is importable the graphs ARE PyG `Data`; only without it (this image has no torch_geometric) they are the
    structurally identical `dss2.batching.Data`, to be batched with `dss2.batching.DataLoader`."""
    try:
        from torch_geometric.data import Data as PygData
        return PygData
    except ImportError:
        return Data
