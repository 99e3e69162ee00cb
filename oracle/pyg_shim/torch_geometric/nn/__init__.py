"""`torch_geometric.nn` subset (oracle shim, test infrastructure): the conv classes and a
`Sequential` name (reference networks.py:4 uses it only for the models outside the hot path)."""
from . import conv
from .conv import (GCN2Conv, FAConv, TAGConv, GINEConv, MessagePassing, GCNConv, ChebConv,
                   GATv2Conv, gcn_norm)


def Sequential(*args, **kwargs):
    raise NotImplementedError("torch_geometric.nn.Sequential: not restated by the oracle shim "
                              "(outside the hot path, SURVEY.md 8f-1)")
