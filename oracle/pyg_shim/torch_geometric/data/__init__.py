"""`torch_geometric.data` subset (oracle shim, test infrastructure): Data, Batch and the two names
reference data.py:4 imports without using (InMemoryDataset, download_url)."""
import torch


class Data:
    """PyG `Data(x, edge_index, edge_attr, y)`; num_nodes is inferred from x.size(0)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def validate(self, raise_on_error=True):
        """PyG `Data.validate`: edge_index is [2, E] int64 with entries in [0, num_nodes)."""
        ok = True
        ei = self.edge_index
        if ei is not None:
            ok = ei.dim() == 2 and ei.size(0) == 2 and ei.dtype == torch.long
            if ok and ei.numel() > 0:
                ok = int(ei.min()) >= 0 and int(ei.max()) < self.num_nodes
        if not ok and raise_on_error:
            raise ValueError("invalid edge_index")
        return ok

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    """PyG `Batch.from_data_list`: x / edge_attr / y concatenated on dim 0,
    edge_index concatenated on dim 1 with the running node offset added, graph-major, original
    edge order kept; `batch` = repeat_interleave(arange(B), num_nodes_g); `ptr` = cumulative nodes."""

    @classmethod
    def from_data_list(cls, data_list):
        offsets, total = [], 0
        for d in data_list:
            offsets.append(total)
            total += d.num_nodes
        out = cls(
            x=torch.cat([d.x for d in data_list], dim=0),
            edge_index=torch.cat([d.edge_index + o for d, o in zip(data_list, offsets)], dim=1),
            edge_attr=torch.cat([d.edge_attr for d in data_list], dim=0),
            y=torch.cat([d.y for d in data_list], dim=0),
        )
        out.batch = torch.cat([torch.full((d.num_nodes,), i, dtype=torch.long) for i, d in enumerate(data_list)])
        out.ptr = torch.tensor(offsets + [total], dtype=torch.long)
        out.num_graphs = len(data_list)
        return out


class InMemoryDataset:  # imported, never used (reference data.py:4)
    pass


def download_url(*a, **k):  # imported, never used (reference data.py:4)
    raise NotImplementedError
