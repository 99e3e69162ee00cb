"""`torch_geometric.nn.conv` subset (oracle shim, test infrastructure).

Implemented from PyG's published semantics: MessagePassing(aggr='add', flow='source_to_target'),
gcn_norm, TAGConv, GCN2Conv, and GATv2Conv (the as-shipped default model of
dss2_run.py:86).  The other conv classes named by reference networks.py:7 (FAConv,
GCNConv, ChebConv) exist only as names so that the reference module imports; GINEConv (networks.py:100) is restated too.
"""
import inspect

import torch
from torch import nn

from ..utils import add_self_loops, remove_self_loops, scatter, softmax


class MessagePassing(nn.Module):
    """PyG MessagePassing, the part EdgeAggregation (reference networks.py:159-209) and TAGConv use.

    propagate(edge_index, **kw): for every argument of `message()`: a name ending in `_j` is
    kw[name[:-2]].index_select(0, edge_index[0]) (source), `_i` is the same at edge_index[1]
    (target); other names are passed through; kwargs that `message()` does not name are dropped.
    Aggregation 'add' is scatter(msg, edge_index[1], dim=0, dim_size=N, reduce='sum')."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        if aggr not in ("add", "sum", "mean"):
            raise NotImplementedError(f"shim supports aggr add/mean, got {aggr}")
        if flow != "source_to_target":
            raise NotImplementedError("shim supports flow='source_to_target' only")
        self.aggr = "sum" if aggr == "add" else aggr
        self.flow = flow
        self.node_dim = node_dim
        self._msg_args = [p for p in inspect.signature(self.message).parameters]

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        num_nodes = None
        call = {}
        for name in self._msg_args:
            if name.endswith("_j") or name.endswith("_i"):
                base = kwargs[name[:-2]]
                if isinstance(base, (tuple, list)):      # PyG: a pair (source-side tensor, target-side tensor)
                    base = base[0] if name.endswith("_j") else base[1]
                if num_nodes is None:
                    num_nodes = base.size(self.node_dim)
                call[name] = base.index_select(self.node_dim, src if name.endswith("_j") else dst)
            else:
                call[name] = kwargs[name]
        if num_nodes is None:
            num_nodes = size[1] if size is not None else int(dst.max()) + 1
        msg = self.message(**call)
        out = scatter(msg, dst, dim=self.node_dim, dim_size=num_nodes, reduce=self.aggr)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False, add_self_loops=True,
             flow="source_to_target", dtype=None):
    """PyG `gcn_norm` for dense-index input.  TAGConv calls it with add_self_loops=False:
    w = 1; deg = scatter(w, col, N); dis = deg^-1/2 with inf -> 0; w = dis[row] * w * dis[col]."""
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    if add_self_loops:
        # PyG `add_remaining_self_loops(edge_index, edge_weight, fill_value = 2 if improved else 1, num_nodes)`: the non-loop edges keep
        # their order, then ONE loop per node is appended (an existing loop keeps its weight; the reference's data has none)
        keep = edge_index[0] != edge_index[1]
        loop_w = torch.full((num_nodes,), 2.0 if improved else 1.0, dtype=edge_weight.dtype, device=edge_index.device)
        if bool((~keep).any()):
            loop_w[edge_index[0][~keep]] = edge_weight[~keep]
        loop = torch.arange(num_nodes, device=edge_index.device)
        edge_index = torch.cat([edge_index[:, keep], torch.stack([loop, loop])], dim=1)
        edge_weight = torch.cat([edge_weight[keep], loop_w])
    row, col = edge_index[0], edge_index[1]
    idx = col if flow == "source_to_target" else row
    deg = scatter(edge_weight, idx, dim=0, dim_size=num_nodes, reduce="sum")
    deg_inv_sqrt = deg.pow_(-0.5)
    deg_inv_sqrt.masked_fill_(deg_inv_sqrt == float("inf"), 0)
    edge_weight = deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]
    return edge_index, edge_weight


class TAGConv(MessagePassing):
    """PyG TAGConv(in, out, K, bias=True, normalize=True):
    lins = ModuleList[(K+1) x Linear(in, out, bias=False)], bias zeros-initialised;
    forward: gcn_norm(no self loops); out = lins[0](x); for k=1..K: x = A_hat x; out = out + lins[k](x);
    out = out + bias.  Used at reference networks.py:230-234."""

    def __init__(self, in_channels, out_channels, K=3, bias=True, normalize=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.in_channels, self.out_channels, self.K, self.normalize = in_channels, out_channels, K, normalize
        self.lins = nn.ModuleList([nn.Linear(in_channels, out_channels, bias=False) for _ in range(K + 1)])
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = gcn_norm(edge_index, edge_weight, x.size(self.node_dim),
                                               improved=False, add_self_loops=False, flow=self.flow,
                                               dtype=x.dtype)
        out = self.lins[0](x)
        for lin in self.lins[1:]:
            x = self.propagate(edge_index, x=x, edge_weight=edge_weight)
            out = out + lin(x)
        if self.bias is not None:
            out = out + self.bias
        return out

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j


def _outside_hot_path(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"torch_geometric.nn.conv.{name}: not restated by the oracle shim "
                                      "(outside the hot path, SURVEY.md 8f-1)")
    _Stub.__name__ = name
    return _Stub


class GCN2Conv(MessagePassing):
    """PyG GCN2Conv(channels, alpha, theta=None, layer=None, shared_weights=True, cached=False, add_self_loops=True, normalize=True), as
    used at reference networks.py:44: weight1 [channels, channels] glorot (weight2 only when shared_weights=False);
    beta = log(theta / layer + 1) when both are given, else 1.
      forward(x, x_0, edge_index): gcn_norm(add_self_loops) (cached after the first call when cached=True);
      x = propagate(A_hat x); x.mul_(1 - alpha); x_0 = alpha * x_0[:N]; out = x.add_(x_0);
      out = addmm(out, out, weight1, beta = 1 - beta, alpha = beta)   (beta = 1: out @ weight1)."""

    def __init__(self, channels, alpha, theta=None, layer=None, shared_weights=True, cached=False, add_self_loops=True, normalize=True,
                 **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        import math
        self.channels, self.alpha, self.beta = channels, alpha, 1.0
        if theta is not None or layer is not None:
            assert theta is not None and layer is not None
            self.beta = math.log(theta / layer + 1)
        self.cached, self.normalize, self.add_self_loops_ = cached, normalize, add_self_loops
        self._cached_edge_index = None
        self.weight1 = nn.Parameter(torch.empty(channels, channels))
        if shared_weights:
            self.register_parameter("weight2", None)
        else:
            self.weight2 = nn.Parameter(torch.empty(channels, channels))
        nn.init.xavier_uniform_(self.weight1)
        if self.weight2 is not None:
            nn.init.xavier_uniform_(self.weight2)

    def forward(self, x, x_0, edge_index, edge_weight=None):
        if self.normalize:
            cache = self._cached_edge_index
            if cache is None:
                edge_index, edge_weight = gcn_norm(edge_index, edge_weight, x.size(self.node_dim), False, self.add_self_loops_, self.flow,
                                                   dtype=x.dtype)
                if self.cached:
                    self._cached_edge_index = (edge_index, edge_weight)
            else:
                edge_index, edge_weight = cache
        x = self.propagate(edge_index, x=x, edge_weight=edge_weight)
        x.mul_(1 - self.alpha)
        x_0 = self.alpha * x_0[:x.size(0)]
        if self.weight2 is None:
            out = x.add_(x_0)
            out = torch.addmm(out, out, self.weight1, beta=1. - self.beta, alpha=self.beta)
        else:
            out = torch.addmm(x, x, self.weight1, beta=1. - self.beta, alpha=self.beta)
            out = out + torch.addmm(x_0, x_0, self.weight2, beta=1. - self.beta, alpha=self.beta)
        return out

    def message(self, x_j, edge_weight):
        return x_j if edge_weight is None else edge_weight.view(-1, 1) * x_j


class FAConv(MessagePassing):
    """PyG FAConv(channels, eps=0.1, dropout=0.0, cached=False, add_self_loops=True, normalize=True), as used at reference
    networks.py:44-50: att_l, att_r = Linear(channels, 1, bias=False);
      forward(x, x_0, edge_index): gcn_norm(add_self_loops) (cached after the first call when cached=True);
      out = propagate(x=x, alpha=(att_l(x), att_r(x)), edge_weight);  out = out + eps * x_0  (when eps != 0);
      message: x_j * (dropout(tanh(alpha_j + alpha_i)) * edge_weight)  with alpha_j = att_l(x)[source], alpha_i = att_r(x)[target]."""

    def __init__(self, channels, eps=0.1, dropout=0.0, cached=False, add_self_loops=True, normalize=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.channels, self.eps, self.dropout = channels, eps, dropout
        self.cached, self.normalize, self.add_self_loops_ = cached, normalize, add_self_loops
        self._cached_edge_index = None
        self.att_l = torch.nn.Linear(channels, 1, bias=False)
        self.att_r = torch.nn.Linear(channels, 1, bias=False)

    def forward(self, x, x_0, edge_index, edge_weight=None):
        if self.normalize:
            assert edge_weight is None
            cache = self._cached_edge_index
            if cache is None:
                edge_index, edge_weight = gcn_norm(edge_index, None, x.size(self.node_dim), False, self.add_self_loops_, self.flow, dtype=x.dtype)
                if self.cached:
                    self._cached_edge_index = (edge_index, edge_weight)
            else:
                edge_index, edge_weight = cache
        else:
            assert edge_weight is not None
        out = self.propagate(edge_index, x=x, alpha=(self.att_l(x), self.att_r(x)), edge_weight=edge_weight)
        if self.eps != 0.0:
            out = out + self.eps * x_0
        return out

    def message(self, x_j, alpha_j, alpha_i, edge_weight):
        alpha = (alpha_j + alpha_i).squeeze(-1).tanh()
        alpha = torch.nn.functional.dropout(alpha, p=self.dropout, training=self.training)
        return x_j * (alpha * edge_weight).view(-1, 1)


class GINEConv(MessagePassing):
    """PyG GINEConv(nn, eps=0., train_eps=False, edge_dim=None), as used at reference networks.py:100:
      out_i = nn( (1 + eps) * x_i + sum_{j -> i} relu(x_j + lin(a_ji)) ), lin = Linear(edge_dim, in_channels of nn) (with bias),
      eps a buffer (train_eps=False) or a parameter, initialised to `eps`."""

    def __init__(self, nn, eps=0.0, train_eps=False, edge_dim=None, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.empty(1))
        else:
            self.register_buffer("eps", torch.empty(1))
        if edge_dim is not None:
            first = nn[0] if isinstance(nn, torch.nn.Sequential) else nn
            in_channels = first.in_features if hasattr(first, "in_features") else first.in_channels
            self.lin = torch.nn.Linear(edge_dim, in_channels)
        else:
            self.lin = None
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index, edge_attr=None):
        out = self.propagate(edge_index, x=x, edge_attr=edge_attr)
        out = out + (1 + self.eps) * x
        return self.nn(out)

    def message(self, x_j, edge_attr):
        if self.lin is not None:
            edge_attr = self.lin(edge_attr)
        return (x_j + edge_attr).relu()

GCNConv = _outside_hot_path("GCNConv")
ChebConv = _outside_hot_path("ChebConv")


class GATv2Conv(MessagePassing):
    """PyG GATv2Conv(in, out, heads=1, concat=True, negative_slope=0.2, dropout=0., add_self_loops=True, edge_dim=None,
    fill_value='mean', bias=True, share_weights=False), as used at reference networks.py:146:
      lin_l, lin_r = Linear(in, H*C, bias=bias) (glorot weights, zero bias); att [1, H, C] glorot; lin_edge = Linear(edge_dim, H*C,
      bias=False); bias [H*C] zeros.
      forward: x_l = lin_l(x), x_r = lin_r(x); remove_self_loops + add_self_loops(fill_value) on (edge_index, edge_attr);
      per edge (j -> i): s = x_r[i] + x_l[j] + lin_edge(a); alpha = softmax_i(sum_c att * leaky_relu(s, slope)); dropout(alpha);
      out[i] = sum_j alpha_ij * x_l[j]; concat heads; + bias."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True, share_weights=False, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(node_dim=0, **kwargs)
        if share_weights:
            raise NotImplementedError("shim: share_weights=False only")
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope, self.dropout, self.add_self_loops_, self.edge_dim, self.fill_value = negative_slope, dropout, add_self_loops, edge_dim, fill_value
        self.lin_l = nn.Linear(in_channels, heads * out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, heads * out_channels, bias=bias)
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = nn.Linear(edge_dim, heads * out_channels, bias=False) if edge_dim is not None else None
        if bias and concat:
            self.bias = nn.Parameter(torch.empty(heads * out_channels))
        elif bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in (self.lin_l, self.lin_r, self.lin_edge):
            if lin is not None:
                nn.init.xavier_uniform_(lin.weight)
                if lin.bias is not None:
                    nn.init.zeros_(lin.bias)
        nn.init.xavier_uniform_(self.att)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_attr=None):
        H, C = self.heads, self.out_channels
        x_l = self.lin_l(x).view(-1, H, C)
        x_r = self.lin_r(x).view(-1, H, C)
        if self.add_self_loops_:
            n = x_l.size(0)
            edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
            edge_index, edge_attr = add_self_loops(edge_index, edge_attr, fill_value=self.fill_value, num_nodes=n)
        src, dst = edge_index[0], edge_index[1]
        s = x_r.index_select(0, dst) + x_l.index_select(0, src)
        if edge_attr is not None:
            if edge_attr.dim() == 1:
                edge_attr = edge_attr.view(-1, 1)
            s = s + self.lin_edge(edge_attr).view(-1, H, C)
        s = torch.nn.functional.leaky_relu(s, self.negative_slope)
        alpha = (s * self.att).sum(dim=-1)
        alpha = softmax(alpha, dst, None, x_l.size(0))
        alpha = torch.nn.functional.dropout(alpha, p=self.dropout, training=self.training)
        out = scatter(x_l.index_select(0, src) * alpha.unsqueeze(-1), dst, dim=0, dim_size=x_l.size(0), reduce="sum")
        out = out.view(-1, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        return out
