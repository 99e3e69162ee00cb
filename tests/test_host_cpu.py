"""CPU tests of the host-side logic: C-ABI export table, model spec vs. reference parameter names, flat layout, the
loud no-fallback behaviour, and world_size-2 gloo checks of the data-parallel plumbing (shard ids, flat all-reduce)."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT, golden_model
from dss2 import _lib, ops
import networks


def test_library_exports_every_declared_symbol():
    """Every function declared in include/dss2_b200.h is exported by the built library and bound in _lib (no compute calls)."""
    header = open(os.path.join(ROOT, "include", "dss2_b200.h")).read()
    declared = set(re.findall(r"\b(dss2_[a-z0-9_]+)\s*\(", header)) - {"dss2_graph"}
    assert os.path.exists(_lib.LIB_PATH), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.exported_symbols())
    assert _lib.load(require_cuda=False).dss2_version() >= 100


@pytest.mark.parametrize("tag", ["skippfn_cigre", "pfn_small_cigre", "mpn_cigre", "skipmpn_cigre"])
def test_spec_matches_reference_state_dict(tag):
    """Parameter names / shapes / order of the drop-in modules and of PFNSpec equal the reference run's state_dict."""
    ctor, kind, sd, _, _, _ = golden_model(tag)
    model = getattr(networks, kind)(**ctor)
    got = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    assert got == [(n, tuple(v.shape)) for n, v in sd.items() if n in dict(got)] and len(got) == len(sd)
    spec = model._spec()
    assert spec.param_names() == got
    table, size = spec.layout()
    for (name, shape), (off, n) in zip(got, table.values()):
        assert off % 4 == 0 and n == torch.Size(shape).numel()
    # the K+1 matrices of one TAGConv are contiguous in the flat buffer
    pre = spec.prefix_fmt.format(s=0)
    o0, n0 = table[pre + "convs.0.lins.0.weight"]
    assert table[pre + "convs.0.lins.1.weight"][0] == o0 + n0
    model.load_state_dict(sd, strict=True)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    model = networks.MPN(8, 6, 2, 32, 2, 2, 0.0)
    with pytest.raises(_lib.Dss2Error):
        model(torch.zeros(4, 8), torch.tensor([[0, 1], [1, 2]]), torch.zeros(2, 6))
    import data
    with pytest.raises(_lib.Dss2Error):
        data.get_pflow(torch.zeros(3, 2), torch.tensor([[0], [1]]), torch.zeros(3, 3), torch.zeros(1, 7))
    gnn = networks.gnn_dsse(dim_feat=8, dim_dense=32, dim_out=2, num_layers=3)
    with pytest.raises(_lib.Dss2Error):
        gnn(torch.zeros(4, 8), torch.tensor([[0, 1], [1, 2]]))
    with pytest.raises(NotImplementedError):
        networks.gnn_dsse(dim_feat=8, dim_dense=32, dim_out=2, num_layers=3, model='fagcn', dropout=0.1)
    gine = networks.GINE_DSSE(dim_feat=8, dim_dense=32, dim_out=2, num_layers=3, edge_dim=6)
    with pytest.raises(_lib.Dss2Error):
        gine(torch.zeros(4, 8), torch.tensor([[0, 1], [1, 2]]), torch.zeros(2, 6))
    trainable = networks.GINE_DSSE(dim_feat=8, dim_dense=32, dim_out=2, num_layers=3, edge_dim=6, eps=0.2, train_eps=True)
    assert "model.module_0.eps" in dict(trainable.named_parameters()) and float(trainable.model.module_2.eps) == pytest.approx(0.2)
    with pytest.raises(NotImplementedError):
        networks.GINE_DSSE(dim_feat=8, dim_dense=32, dim_out=2, num_layers=3, edge_dim=6, nonlin='relu')
    gat = networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=3, edge_dim=6)
    with pytest.raises(_lib.Dss2Error):
        gat(torch.zeros(4, 8), torch.tensor([[0, 1], [1, 2]]), torch.zeros(2, 6))
    with pytest.raises(NotImplementedError):      # the reference's own stack cannot run heads > 1 with concat=True
        networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=2, num_layers=3, edge_dim=6)
    multi = networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=2, concat=False, num_layers=3, edge_dim=6)
    assert tuple(multi.model.module_0.lin_l.weight.shape) == (16, 8) and tuple(multi.model.module_0.att.shape) == (1, 2, 8)


def test_gat_dsse_state_dict_names_match_reference_layout():
    """Parameter names / shapes of GAT_DSSE equal the reference's (PyG Sequential naming), and the flat layout covers all of them."""
    from dss2 import gat
    import dss2_oracle as orc
    m = networks.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=1, num_layers=8, edge_dim=6)
    sd = orc.init_gat_state_dict(seed=1)
    assert set(sd) == set(m.state_dict()) and all(tuple(sd[k].shape) == tuple(v.shape) for k, v in m.state_dict().items())
    table, size = gat.GATSpec(8, 32, 2, 8, 6).layout()
    assert set(table) == set(dict(m.named_parameters()))
    spans = sorted(table.values())
    assert all(a[0] + a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][0] + spans[-1][1] <= size


def test_gine_dsse_state_dict_names_match_reference():
    """GINE_DSSE exposes the reference's state_dict keys, including the shared Linear's aliases `model.module_{2l}.nn.*` and the
    `eps` buffers, and reports the shared parameters once."""
    from conftest import load_golden
    from dss2 import gine
    z = load_golden("golden_model_gine_cigre.npz")
    m = networks.GINE_DSSE(dim_feat=8, dim_dense=32, dim_out=2, num_layers=int(z["num_layers"]), edge_dim=6)
    assert sorted(m.state_dict().keys()) == sorted(z["state_dict_keys"].tolist())
    names = [n for n, _ in m.named_parameters()]
    assert names.count("nn.weight") == 1 and not any(".nn." in n for n in names)
    table, _ = gine.GINESpec(8, 32, 2, int(z["num_layers"]), 6).layout()
    assert set(table) == set(names)


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(_lib.Dss2Error):
        ops.validate_spec(ops.PFNSpec(fn=8, fe=6, dim_out=2, n_layers=2, K=2, L=1, p_drop=0.0, skip=(False,), prefix_fmt="", hid=64))
    with pytest.raises(_lib.Dss2Error):
        ops.validate_spec(ops.PFNSpec(fn=8, fe=6, dim_out=2, n_layers=2, K=2, L=1, p_drop=0.0, skip=(True,), prefix_fmt=""))


def _dp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the exchange of the data-parallel step: one sum all-reduce of the flat gradient, then scale by 1/world (trainer.py)
    spec = ops.PFNSpec(fn=8, fe=6, dim_out=2, n_layers=8, K=2, L=5, p_drop=0.3, skip=(True,) * 4 + (False,), prefix_fmt="mpns.{s}.")
    _, size = spec.layout()
    g = torch.full((size,), float(rank + 1))
    dist.all_reduce(g)
    mean = g / world
    # shard ownership: contiguous scenario ranges, disjoint and covering
    total = 1000
    lo, hi = rank * total // world, (rank + 1) * total // world
    owned = torch.zeros(total)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    out[rank] = (float(mean[0]), float(mean[-1]), bool((owned == 1).all()), size)
    dist.destroy_process_group()


def test_data_parallel_plumbing_gloo_world2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, 29611, out), nprocs=world, join=True)
        res = dict(out)
    for r in range(world):
        first, last, covered, size = res[r]
        assert first == 1.5 and last == 1.5 and covered and size >= 120898


def test_gnn_dsse_state_dict_names_match_reference_layout():
    """Parameter names / shapes of gnn_dsse ('gcn2', 'tagcn', 'fagcn') equal the reference's (PyG Sequential naming; the goldens hold the
    reference's own named_parameters()), and the flat layout covers all of them."""
    import dss2_oracle as orc
    from conftest import golden_gat
    from dss2 import gnn
    for tag in ("gnn_gcn2_cigre", "gnn_tagcn_cigre", "gnn_fagcn_cigre"):
        nl, sd, _, z = golden_gat(tag)
        m = networks.gnn_dsse(dim_feat=8, dim_dense=32, dim_out=2, num_layers=nl, K=int(z["K"]), model=str(z["model"]))
        ours = {k: tuple(v.shape) for k, v in m.named_parameters()}
        assert ours == {k: tuple(v.shape) for k, v in sd.items()}
        m.load_state_dict(sd, strict=True)
        table, size = gnn.GNNSpec(str(z["model"]), 8, 32, 2, nl, K=int(z["K"])).layout()
        assert set(table) == set(ours) and max(o + n for o, n in table.values()) <= size
