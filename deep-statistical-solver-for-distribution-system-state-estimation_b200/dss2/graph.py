"""Kernel-private structure of one batch (doubled-graph CSR, degree norm, tiling) and its cache.

The reference rebuilds the undirected edge list in each of the 5 sub-nets (networks.py:236-258, with a
host sync in `is_directed`) and PyG recomputes `gcn_norm` in each of the 40 TAGConv layers.  Here the
structure is built once per batch topology by `dss2_graph_build` and shared by every layer and by the
loss.  Batches produced by `dss2.batching` carry their structure with them (attribute `_dss2_graph` on
the edge_index tensor); foreign tensors get one built on first use.
"""
import ctypes

import torch

from . import _lib


class BatchGraph:
    """Owns the device workspace behind a `dss2_graph_t`."""

    def __init__(self, edge_index, num_nodes, ptr=None, undirect=-1, tile_cap=_lib.TILE_CAP):
        lib = _lib.load()
        if edge_index.device.type != "cuda":
            raise _lib.Dss2Error("BatchGraph needs CUDA tensors")
        if edge_index.dtype != torch.long or edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must be int64 [2, E]")
        self.edge_index = edge_index.contiguous()
        self.num_nodes = int(num_nodes)
        self.tile_cap = int(tile_cap)
        self.num_edges = int(edge_index.size(1))
        if ptr is None:
            ptr = discover_segments(self.edge_index, self.num_nodes)
        self.ptr = ptr.to(device=edge_index.device, dtype=torch.long).contiguous()
        self.num_graphs = self.ptr.numel() - 1
        nbytes = lib.dss2_graph_workspace_bytes(self.num_nodes, self.num_edges, self.num_graphs)
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=edge_index.device)
        self.c = _lib.GraphStruct()
        with torch.cuda.device(edge_index.device):
            rc = lib.dss2_graph_build(ctypes.byref(self.c), _lib.ptr(self.edge_index), self.num_edges, self.num_nodes,
                                      _lib.ptr(self.ptr), self.num_graphs, undirect, tile_cap, _lib.ptr(self.ws), nbytes,
                                      _lib.stream())
        _lib.check(rc, "dss2_graph_build")
        self._wls_ws = None
        self.scratch = None
        if self.c.num_tiles == 0:      # a graph exceeds a tile: the large-graph kernels need scratch
            n = lib.dss2_generic_scratch_bytes(self.num_nodes)
            self.scratch = torch.empty(n, dtype=torch.uint8, device=edge_index.device)
            self.c.scratch = self.scratch.data_ptr()
            self.c.scratch_bytes = n

    @property
    def ref(self):
        return ctypes.byref(self.c)

    @property
    def device(self):
        return self.edge_index.device

    def wls_workspace(self):
        if self._wls_ws is None:
            n = _lib.load().dss2_wls_workspace_bytes(self.ref)
            self._wls_ws = torch.zeros(n, dtype=torch.uint8, device=self.device)
        return self._wls_ws

    # views of the kernel-private arrays (tests / debugging)
    def arrays(self):
        def view(addr, n, dtype):
            if n == 0:
                return torch.empty(0, dtype=dtype, device=self.device)
            off = addr - self.ws.data_ptr()
            nbytes = n * torch.empty(0, dtype=dtype).element_size()
            return self.ws[off:off + nbytes].view(dtype)
        c = self.c
        return {
            "rowptr": view(c.rowptr, c.num_nodes + 1, torch.int32), "col": view(c.col, c.nnz, torch.int32),
            "eid": view(c.eid, c.nnz, torch.int32), "dis": view(c.dis, c.num_nodes, torch.float32),
            "eptr": view(c.eptr, c.num_graphs + 1, torch.int64), "w": view(c.w, c.nnz, torch.float32),
        }


def discover_segments(edge_index, num_nodes):
    """Node offsets of the independent segments of a batch when the caller did not supply `ptr`
    (the reference's model signature carries no batch vector, dss2_run.py:138).  A boundary b is a cut
    iff no edge spans it; plain tensor ops on the device, one host sync for the count."""
    dev = edge_index.device
    if edge_index.numel() == 0 or num_nodes == 0:
        return torch.tensor([0, num_nodes], dtype=torch.long, device=dev)
    lo = torch.minimum(edge_index[0], edge_index[1])
    hi = torch.maximum(edge_index[0], edge_index[1])
    span = torch.zeros(num_nodes + 1, dtype=torch.long, device=dev)
    span.index_add_(0, lo + 1, torch.ones_like(lo))
    span.index_add_(0, hi + 1, -torch.ones_like(hi))
    open_edges = torch.cumsum(span, 0)[:num_nodes]          # edges crossing the boundary before node b
    cuts = torch.nonzero(open_edges == 0).flatten()         # includes b = 0
    return torch.cat([cuts, torch.tensor([num_nodes], dtype=torch.long, device=dev)])


def graph_for(edge_index, num_nodes):
    """The cached BatchGraph attached to `edge_index`, building (and attaching) it if needed."""
    g = getattr(edge_index, "_dss2_graph", None)
    if g is not None and g.num_nodes == num_nodes and g.edge_index.data_ptr() == edge_index.data_ptr():
        return g
    g = BatchGraph(edge_index, num_nodes)
    try:
        edge_index._dss2_graph = g
    except AttributeError:
        pass
    return g
