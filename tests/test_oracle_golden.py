"""CPU tests: pin the oracle (oracle/dss2_oracle.py) and the host-side dataset builder against
 (i) pandapower-solved columns of the reference's CIGRE-14 pickles (physics KAT) and
 (ii) outputs of the reference's own code run over the PyG shim (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import REG_COEFS, assert_fp32_parity, golden_model, load_golden, split_masks
import dss2_oracle as orc
from dss2 import dataset, synth

CIGRE_MEAS_V, CIGRE_MEAS_PF = np.array([0, 1, 12, 7, 11, 14]), np.array([0, 10])


def _fixture():
    return load_golden("cigre14_scenarios.npz")


def test_get_pflow_matches_pandapower_fp64():
    """Feeding pandapower's solved (vm_pu, va_rad) to the branch equations reproduces pandapower's own
    branch results (fp64): P/Q <= 1e-11 MW, currents <= 1e-7 kA, line loading <= 1e-7, bus P <= 1e-7 MW."""
    fx = _fixture()
    for s in (0, 17, 127):
        nodes, edges, labels = fx["nodes"][s], fx["edges"][s], fx["labels"][s]
        closed = edges[:, 6] == 1.0
        ce = torch.from_numpy(edges[closed])
        ei = ce[:, 0:2].long().t().contiguous()
        out = orc.get_pflow(torch.from_numpy(labels), ei, torch.from_numpy(nodes[:, 0:3]), ce[:, 2:9])
        ll, lt, pf, qf, pt, qt, i_f, i_t = [o.numpy() for o in out]
        cols = ce.numpy()
        assert np.abs(pf - cols[:, 9]).max() < 1e-11 and np.abs(qf - cols[:, 10]).max() < 1e-11
        assert np.abs(pt - cols[:, 11]).max() < 1e-11 and np.abs(qt - cols[:, 12]).max() < 1e-11
        assert np.abs(i_f - cols[:, 15]).max() < 1e-7 and np.abs(i_t - cols[:, 16]).max() < 1e-7
        is_line = cols[:, 7] == 0.0
        assert np.abs(ll[is_line] - cols[is_line, 17] / 100.0).max() < 1e-7
        # reference quirk (data.py:388): its transformer loading is pandapower's divided by sqrt(3)
        assert np.abs(lt[~is_line] * np.sqrt(3.0) - cols[~is_line, 17] / 100.0).max() < 1e-6
        n = nodes.shape[0]
        p_bus = -orc.segment_sum(out[4], ei[1], n) - orc.segment_sum(out[2], ei[0], n)
        assert np.abs(p_bus.numpy() - nodes[:, 5]).max() < 1e-7


def test_dataset_builder_bit_exact_vs_reference():
    """dss2.dataset.build_scenario_store == reference data_from_pickles (same np.random stream), bit for bit."""
    fx, gd = _fixture(), load_golden("golden_dataset_cigre14.npz")
    S = fx["nodes"].shape[0]
    grid = synth.load_grid("cigre14")
    zn, ze = dataset.reference_noise_stream(0, S, 15, 14)
    st = dataset.build_scenario_store(fx["nodes"], fx["edges"], fx["labels"], grid["noise_param"], CIGRE_MEAS_V,
                                      CIGRE_MEAS_PF, zn, ze)
    assert np.array_equal(st.x.numpy(), gd["x"])
    assert np.array_equal(st.edge_attr.numpy(), gd["edge_attr"])
    assert np.array_equal(st.y.numpy(), gd["y"])
    assert np.array_equal(st.edge_index[:, :14].numpy(), gd["edge_index"])
    for ours, key in ((st.x_mean, "x_mean"), (st.x_std, "x_std"), (st.edge_mean, "edge_mean"), (st.edge_std, "edge_std")):
        assert np.array_equal(ours.numpy(), gd[key]), key
    # structural facts recorded in SURVEY.md 8c
    assert st.edge_index[:, :14].tolist() == [[1, 2, 3, 4, 5, 7, 8, 9, 10, 3, 12, 13, 0, 0],
                                               [2, 3, 4, 5, 6, 8, 9, 10, 11, 8, 13, 14, 1, 12]]


def _store():
    fx = _fixture()
    grid = synth.load_grid("cigre14")
    zn, ze = dataset.reference_noise_stream(0, fx["nodes"].shape[0], 15, 14)
    return dataset.build_scenario_store(fx["nodes"], fx["edges"], fx["labels"], grid["noise_param"], CIGRE_MEAS_V,
                                        CIGRE_MEAS_PF, zn, ze)


def test_collate_bit_exact():
    gd = load_golden("golden_dataset_cigre14.npz")
    st = _store()
    b = orc.collate([st.graph(int(i)) for i in gd["pick"]])
    for k in ("x", "edge_index", "edge_attr", "y", "batch", "ptr"):
        assert np.array_equal(b[k].numpy(), gd["pick_" + k]), k


def test_wls_loss_known_answers():
    """gsp_wls_edge KATs: output = truth and output = 0 on the first 64 graphs (reference values)."""
    gd = load_golden("golden_dataset_cigre14.npz")
    st = _store()
    b = orc.collate([st.graph(i) for i in range(64)])
    truth = torch.stack([(b["y"][:, 0] - st.x_mean[0]) / st.x_std[0], b["y"][:, 1]], 1)
    for out, key in ((truth, "kat_loss_truth"), (torch.zeros_like(truth), "kat_loss_zeros")):
        loss = orc.wls_loss(b["x"], b["edge_attr"], out, st.x_mean, st.x_std, st.edge_mean, st.edge_std,
                            b["edge_index"], REG_COEFS)
        assert abs(loss.item() - float(gd[key])) <= 2e-6 * abs(float(gd[key])), key
    assert abs(float(gd["kat_loss_truth"]) - 3.93080e-05) < 1e-9


def test_wls_loss_and_grad_all_penalties_active():
    z = load_golden("golden_loss_ober_wild.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    out = t["output"].clone().requires_grad_(True)
    loss = orc.wls_loss(t["x"], t["edge_attr"], out, t["x_mean"], t["x_std"], t["edge_mean"], t["edge_std"],
                        t["edge_index"], REG_COEFS)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    scale = float(np.abs(z["grad_out"]).max())
    assert float((out.grad - t["grad_out"]).abs().max()) <= 1e-5 * scale


def _oracle_run(kind, ctor, sd, x, ea, ei, masks, stats, grad_out, dtype):
    sd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    x, ea = x.to(dtype), ea.to(dtype)
    p = ctor["dropout_rate"]
    if kind in ("MPN", "SkipMPN"):
        out = orc.mpn_forward(sd, "", x[:, :8], ei, ea[:, :6], p, skip=(kind == "SkipMPN"), masks=masks)
    else:
        out = orc.pfn_forward(sd, x[:, :8], ei, ea[:, :6], p, skip=(kind == "SkipPFN"), masks=split_masks(masks, ctor))
    if ctor["dim_out"] == 2:
        loss = orc.wls_loss(x, ea, out, *[s.to(dtype) for s in stats], ei, REG_COEFS)
    else:
        loss = (out * grad_out.to(dtype)).sum()
    loss.backward()
    return out.detach(), loss.detach(), {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("tag", ["skippfn_cigre", "pfn_small_cigre", "mpn_cigre", "skipmpn_cigre", "skippfn_ober"])
def test_model_forward_backward_vs_reference(tag):
    """Oracle model + loss + autograd == reference networks.py/data.py run (same weights, same dropout masks).
    fp64 oracle is the arbiter (conftest.assert_fp32_parity)."""
    ctor, kind, sd, grads, masks, z = golden_model(tag)
    x, ea, ei = torch.from_numpy(z["x"]), torch.from_numpy(z["edge_attr"]), torch.from_numpy(z["edge_index"])
    st = [torch.from_numpy(z[k]) for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    go = torch.from_numpy(z["grad_out"])
    out32, loss32, g32 = _oracle_run(kind, ctor, sd, x, ea, ei, masks, st, go, torch.float32)
    out64, loss64, g64 = _oracle_run(kind, ctor, sd, x, ea, ei, masks, st, go, torch.float64)
    assert_fp32_parity(out32, z["out"], out64, "out")
    if "loss" in z.files:
        assert_fp32_parity(loss32, z["loss"], loss64, "loss")
    for name, g in grads.items():
        assert_fp32_parity(g32[name], g, g64[name], name)


@pytest.mark.parametrize("tag", ["gat_cigre", "gat_ober", "gat_noloop_tanh_cigre", "gat_relu_cigre", "gat_heads2_cigre", "gat_heads3_tanh_cigre"])
def test_gat_dsse_forward_backward_vs_reference(tag):
    """Oracle GAT_DSSE (7 GATv2 layers + 2 Linear) + loss + autograd == the reference's GAT_DSSE run over the shim (same weights)."""
    from conftest import golden_gat, oracle_gat_run
    nl, sd, grads, z = golden_gat(tag)
    out32, loss32, g32 = oracle_gat_run(orc, nl, sd, z, torch.float32)
    out64, loss64, g64 = oracle_gat_run(orc, nl, sd, z, torch.float64)
    # pinning direction: the REFERENCE's recorded run must sit within fp32 noise of the fp64 oracle (noise = the oracle's own fp32 run)
    assert_fp32_parity(z["out"], out32, out64, "out")
    assert_fp32_parity(z["loss"], loss32, loss64, "loss")
    for name, g in grads.items():
        assert_fp32_parity(g, g32[name], g64[name], name)


@pytest.mark.parametrize("tag", ["gine_cigre", "gine_ober", "gine_traineps_cigre"])
def test_gine_dsse_forward_backward_vs_reference(tag):
    """Oracle GINE_DSSE (7 GINEConv layers sharing one Linear + 2 Linear) + loss + autograd == the reference's GINE_DSSE run over the shim."""
    from conftest import golden_gat, oracle_gine_run
    nl, sd, grads, z = golden_gat(tag)
    out32, loss32, g32 = oracle_gine_run(orc, nl, sd, z, torch.float32)
    out64, loss64, g64 = oracle_gine_run(orc, nl, sd, z, torch.float64)
    # pinning direction (see the GAT test above)
    assert_fp32_parity(z["out"], out32, out64, "out")
    assert_fp32_parity(z["loss"], loss32, loss64, "loss")
    for name, g in grads.items():
        assert_fp32_parity(g, g32[name], g64[name], name)


@pytest.mark.parametrize("tag", ["gnn_gcn2_cigre", "gnn_tagcn_cigre", "gnn_gcn2_ober", "gnn_fagcn_cigre", "gnn_fagcn_ober"])
def test_gnn_dsse_forward_backward_vs_reference(tag):
    """Oracle gnn_dsse (GCN2Conv / TAGConv stacks on the one-way edge list + 2 Linear) + loss + autograd == the reference's gnn_dsse run
    over the shim (same weights): the reference's recorded run must sit within fp32 noise of the fp64 oracle."""
    from conftest import golden_gat, oracle_gnn_run
    _, sd, grads, z = golden_gat(tag)
    out32, loss32, g32 = oracle_gnn_run(orc, z, sd, torch.float32)
    out64, loss64, g64 = oracle_gnn_run(orc, z, sd, torch.float64)
    assert_fp32_parity(z["out"], out32, out64, "out")
    assert_fp32_parity(z["loss"], loss32, loss64, "loss")
    for name, g in grads.items():
        assert_fp32_parity(g, g32[name], g64[name], name)


def test_load_samplers_match_the_reference_bit_for_bit():
    """oracle mc_uniform / mc_normal / sample_profiles against the reference's own loadsampling functions (golden_loadsampling.npz,
    generated by importing /root/reference/loadsampling.py): identical np.random stream, identical float64 arithmetic -> bit-exact."""
    z = load_golden("golden_loadsampling.npz")
    sd = [int(v) for v in z["seeds"]]
    np.random.seed(sd[0])
    assert np.array_equal(orc.mc_uniform(z["lb"], z["ub"], 5), z["uni"])
    np.random.seed(sd[1])
    assert np.array_equal(orc.mc_normal(z["mu"], z["sig"], 5), z["nor"])
    np.random.seed(sd[2])
    assert np.array_equal(orc.mc_uniform(0.4, 1.7, 6), z["uni_s"])
    np.random.seed(sd[3])
    assert np.array_equal(orc.mc_normal(0.9, 0.2, 6), z["nor_s"])
    from dss2 import sampling
    iters, err = int(z["iters"]), float(z["pm_error"])
    np.random.seed(sd[4])
    assert np.array_equal(orc.sample_profiles(z["p_mw"], z["hh_mask"], z["ind_mask"], sampling.HOUSEHOLD, sampling.INDUSTRY, iters, "normal", err / 2),
                          z["mc_normal"])
    np.random.seed(sd[5])
    assert np.array_equal(orc.sample_profiles(z["p_mw"], z["hh_mask"], z["ind_mask"], sampling.HOUSEHOLD, sampling.INDUSTRY, iters, "uniform", err),
                          z["mc_uniform"])
