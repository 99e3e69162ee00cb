"""`torch_geometric.utils` subset: scatter, degree, get_laplacian, and the self-loop / softmax helpers GATv2Conv uses
(oracle shim, test infrastructure)."""
import torch


def _expand_index(index, src, dim):
    # PyG `utils._scatter.broadcast`: index is 1-D along `dim`, broadcast to src's shape.
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """PyG `utils.scatter` (>=2.3): `src.new_zeros(size).scatter_add_(dim, broadcast(index), src)`
    for reduce in {sum, add}; mean divides by the clamped count.  Used at reference data.py:428-429."""
    dim = dim if dim >= 0 else src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.shape)
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _expand_index(index, src, dim), src)
    if reduce == "mean":
        count = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _expand_index(index, src, dim), src)
        shape = [1] * src.dim()
        shape[dim] = -1
        return out / count.view(shape)
    if reduce == "max":
        out = src.new_full(size, float("-inf")).scatter_reduce_(dim, _expand_index(index, src, dim), src, reduce="amax", include_self=True)
        return out.masked_fill(out == float("-inf"), 0)      # PyG: empty segments give 0
    raise NotImplementedError(reduce)


def remove_self_loops(edge_index, edge_attr=None):
    """PyG `utils.remove_self_loops`: drop edges with row == col (and their attributes)."""
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    """PyG `utils.add_self_loops`: append (i, i) for i < N after the existing edges.  For `fill_value='mean'` the loop attribute of node
    i is `scatter(edge_attr, edge_index[1], dim=0, dim_size=N, reduce='mean')[i]`: the mean over the edges that point AT i, 0 when
    there is none."""
    n = num_nodes if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    loop = torch.arange(n, device=edge_index.device)
    loop_index = torch.stack([loop, loop])
    if edge_attr is not None:
        if fill_value == "mean" or fill_value == "add" or fill_value == "sum":
            loop_attr = scatter(edge_attr, edge_index[1], dim=0, dim_size=n, reduce="mean" if fill_value == "mean" else "sum")
        elif fill_value is None:
            loop_attr = edge_attr.new_ones((n,) + tuple(edge_attr.shape[1:]))
        else:
            loop_attr = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), float(fill_value))
        edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
    return torch.cat([edge_index, loop_index], dim=1), edge_attr


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """PyG `utils.softmax` over segments given by `index`: subtract the (detached) segment max, exp, divide by segment sum + 1e-16."""
    n = num_nodes if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    src_max = scatter(src.detach(), index, dim, dim_size=n, reduce="max")
    out = (src - src_max.index_select(dim, index)).exp()
    out_sum = scatter(out, index, dim, dim_size=n, reduce="sum") + 1e-16
    return out / out_sum.index_select(dim, index)


def degree(index, num_nodes=None, dtype=None):
    """PyG `utils.degree`: zeros(N).scatter_add_(0, index, ones).  Used at reference networks.py:197."""
    if num_nodes is None:
        num_nodes = int(index.max()) + 1
    out = torch.zeros((num_nodes,), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    """PyG `utils.get_laplacian`, normalization=None: L = D - A as (edge_index, edge_weight) with
    self loops removed first and the degree appended as N self-loop entries.  The only call site
    (reference data.py:422) discards the result."""
    keep = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, keep]
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    else:
        edge_weight = edge_weight[keep]
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    row = edge_index[0]
    deg = scatter(edge_weight, row, 0, dim_size=num_nodes, reduce="sum")
    loop = torch.arange(num_nodes, device=edge_index.device)
    edge_index = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
    edge_weight = torch.cat([-edge_weight, deg], dim=0)
    return edge_index, edge_weight
