"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's unmodified `data.py` and `networks.py` with `oracle/pyg_shim` standing
in for torch_geometric (absent from the image), feeds them seeded inputs and stores inputs and
outputs as .npz.  Tests never import the reference; they compare the oracle (CPU tests) and the
CUDA path (gpu tests) with these files.

Files written
  golden_dataset_cigre14.npz  reference data_from_pickles on the first 128 CIGRE-14 scenarios after
                              np.random.seed(0); loss KATs of gsp_wls_edge on the first 64 graphs
  golden_model_<tag>.npz      reference model forward + gsp_wls_edge + backward: inputs, recorded
                              dropout masks, output (before / after the in-place slack masking), loss,
                              all parameter gradients and the gradient w.r.t. the model output
  golden_loadsampling.npz     reference loadsampling.samplermontecarlo / samplermontecarlo_normal (vector and scalar arguments) after
                              np.random.seed, and the load-profile step of toy_network.py:106-129 re-typed on top of them (toy_network
                              itself imports pandapower, which the image does not have)
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200")
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import data as ref_data          # noqa: E402  (reference)
import networks as ref_net       # noqa: E402  (reference)
from torch_geometric.data import Batch, Data  # noqa: E402  (shim)
import dss2_oracle as orc        # noqa: E402

sys.path.insert(0, PKG)
from dss2 import synth           # noqa: E402

REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}  # dss2_run.py:104-112
CIGRE_MEAS_V, CIGRE_MEAS_PF = np.array([0, 1, 12, 7, 11, 14]), np.array([0, 10])


def ref_loss(b, out, stats):
    xm, xs, em, es = stats
    return ref_data.gsp_wls_edge(input=b.x[:, :8], edge_input=b.edge_attr[:, :6], output=out, x_mean=xm, x_std=xs,
                                 edge_mean=em, edge_std=es, edge_index=b.edge_index, reg_coefs=REG,
                                 num_samples=b.batch[-1] + 1, node_param=b.x[:, 8:], edge_param=b.edge_attr[:, 6:])


class RecordingDropout(torch.nn.Module):
    """Same arithmetic as torch's CPU dropout (x * (bernoulli(1-p) / (1-p))) but keeps the mask."""
    record = []

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p = p

    def forward(self, x):
        noise = F.dropout(torch.ones_like(x), self.p, True)
        RecordingDropout.record.append((noise != 0).numpy().copy())
        return x * noise


def run_model(tag, kind, ctor, batch, stats, sd_seed, torch_seed, dtype=torch.float32):
    sd = orc.init_state_dict(kind=kind, dim_featn=ctor["dim_featn"], dim_feate=ctor["dim_feate"], dim_out=ctor["dim_out"],
                             dim_hid=ctor["dim_hid"], n_gnn_layers=ctor["n_gnn_layers"], K=ctor["K"],
                             L=ctor.get("L", 1), seed=sd_seed)
    # non-zero TAG biases so that the bias path is exercised
    g = torch.Generator().manual_seed(sd_seed + 1)
    for k in sd:
        if "convs." in k and k.endswith(".bias"):
            sd[k] = (torch.rand(sd[k].shape, generator=g) - 0.5) * 0.2
    model = getattr(ref_net, kind)(**ctor)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(dtype)
    real = torch.nn.Dropout
    torch.nn.Dropout = RecordingDropout
    RecordingDropout.record = []
    try:
        torch.manual_seed(torch_seed)
        model.train()
        x = batch.x.to(dtype)
        ea = batch.edge_attr.to(dtype)
        out = model(x[:, :8], batch.edge_index, ea[:, :6])
    finally:
        torch.nn.Dropout = real
    res = {"x": batch.x.numpy(), "edge_index": batch.edge_index.numpy(), "edge_attr": batch.edge_attr.numpy(),
           "ptr": batch.ptr.numpy(), "out": out.detach().numpy().copy(),
           "masks": np.packbits(np.stack(RecordingDropout.record).reshape(-1)) if RecordingDropout.record else np.zeros(0, np.uint8),
           "num_masks": len(RecordingDropout.record),
           "kind": kind, "ctor": np.array(sorted(ctor.items()), dtype=object).astype(str), "sd_seed": sd_seed,
           "x_mean": stats[0].numpy(), "x_std": stats[1].numpy(), "edge_mean": stats[2].numpy(), "edge_std": stats[3].numpy()}
    if ctor["dim_out"] == 2:
        got = {}
        out.register_hook(lambda g_: got.__setitem__("g", g_.clone()))   # gradient w.r.t. the pre-masking output
        bb = Batch(x=x, edge_index=batch.edge_index, edge_attr=ea, y=batch.y)
        bb.batch = batch.batch
        loss = ref_loss(bb, out, [s.to(dtype) for s in stats])
        loss.backward()
        res["loss"] = np.array(loss.item())
        res["out_after_loss"] = out.detach().numpy().copy()
        res["grad_out"] = got["g"].numpy().copy()
        slack = batch.x[:, 9] == 1.0
        assert float(np.abs(res["grad_out"][slack.numpy(), 1]).max()) == 0.0
    else:
        # models that do not end in (V, theta): a fixed linear functional of the output as the "loss"
        gw = torch.linspace(-1.0, 1.0, out.numel(), dtype=dtype).reshape(out.shape)
        (out * gw).sum().backward()
        res["grad_out"] = gw.numpy()
    for name, p in model.named_parameters():
        res["grad." + name] = p.grad.detach().numpy().copy()
        res["param." + name] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"golden_model_{tag}.npz"), **res)
    print(tag, "out", tuple(out.shape), "loss", res.get("loss"), "masks", res["num_masks"])


def run_gat(tag, batch, stats, sd_seed, num_layers=8, **opts):
    """Reference GAT_DSSE (dss2_run.py:86 hyper-parameters; `opts`: concat / slope / self_loops / nonlin) forward + gsp_wls_edge + backward."""
    heads = opts.pop("heads", 1)
    sd = orc.init_gat_state_dict(num_layers=num_layers, seed=sd_seed, heads=heads)
    model = ref_net.GAT_DSSE(dim_feat=8, dim_dense=32, dim_out=2, heads=heads, num_layers=num_layers, edge_dim=6, **opts)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    out = model(batch.x[:, :8], batch.edge_index, batch.edge_attr[:, :6])
    res = {"x": batch.x.numpy(), "edge_index": batch.edge_index.numpy(), "edge_attr": batch.edge_attr.numpy(), "ptr": batch.ptr.numpy(),
           "out": out.detach().numpy().copy(), "num_layers": num_layers, "sd_seed": sd_seed,
           "x_mean": stats[0].numpy(), "x_std": stats[1].numpy(), "edge_mean": stats[2].numpy(), "edge_std": stats[3].numpy()}
    if opts:
        res.update(opt_nonlin=np.array(opts.get("nonlin", "leaky_relu")), opt_slope=np.array(opts.get("slope", 0.2)),
                   opt_self_loops=np.array(opts.get("self_loops", True)), opt_concat=np.array(opts.get("concat", True)),
                   opt_heads=np.array(heads))
    got = {}
    out.register_hook(lambda g_: got.__setitem__("g", g_.clone()))
    loss = ref_loss(batch, out, stats)
    loss.backward()
    res["loss"] = np.array(loss.item())
    res["out_after_loss"] = out.detach().numpy().copy()
    res["grad_out"] = got["g"].numpy().copy()
    for name, p in model.named_parameters():
        res["grad." + name] = p.grad.detach().numpy().copy()
        res["param." + name] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"golden_model_{tag}.npz"), **res)
    print(tag, "out", tuple(out.shape), "loss", res["loss"])


def run_gnn(tag, model, batch, stats, sd_seed, num_layers=4, K=3):
    """Reference gnn_dsse (networks.py:11-69; model='gcn2', 'fagcn' or 'tagcn', defaults otherwise; cached=False so that the run does not depend
    on call history) forward + gsp_wls_edge + backward on `batch`.  forward(x, edge_index): the one-way edge list as the script has it."""
    sd = orc.init_gnn_state_dict(model=model, num_layers=num_layers, K=K, seed=sd_seed)
    net = ref_net.gnn_dsse(dim_feat=8, dim_dense=32, dim_out=2, num_layers=num_layers, K=K, cached=False, model=model)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.train()
    out = net(batch.x[:, :8], batch.edge_index)
    res = {"x": batch.x.numpy(), "edge_index": batch.edge_index.numpy(), "edge_attr": batch.edge_attr.numpy(), "ptr": batch.ptr.numpy(),
           "out": out.detach().numpy().copy(), "num_layers": num_layers, "K": K, "model": model, "sd_seed": sd_seed,
           "x_mean": stats[0].numpy(), "x_std": stats[1].numpy(), "edge_mean": stats[2].numpy(), "edge_std": stats[3].numpy()}
    got = {}
    out.register_hook(lambda g_: got.__setitem__("g", g_.clone()))
    loss = ref_loss(batch, out, stats)
    loss.backward()
    res["loss"] = np.array(loss.item())
    res["grad_out"] = got["g"].numpy().copy()
    for name, p in net.named_parameters():
        res["grad." + name] = p.grad.detach().numpy().copy()
        res["param." + name] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"golden_model_{tag}.npz"), **res)
    print(tag, "out", tuple(out.shape), "loss", res["loss"])


def run_gine(tag, batch, stats, sd_seed, num_layers=8, **opts):
    """Reference GINE_DSSE forward + gsp_wls_edge + backward on `batch` (parameters stored under their named_parameters() names).
    opts: eps / train_eps (trainable eps parameters are drawn per layer so that every layer has its own value)."""
    sd = orc.init_gine_state_dict(num_layers=num_layers, seed=sd_seed)
    model = ref_net.GINE_DSSE(dim_feat=8, dim_dense=32, dim_out=2, num_layers=num_layers, edge_dim=6, **opts)
    if opts.get("train_eps"):
        gen = torch.Generator().manual_seed(sd_seed + 1000)
        for l in range(num_layers - 1):
            sd[f"model.module_{2 * l}.eps"] = (torch.rand(1, generator=gen) - 0.5) * 0.6
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(sd[name])
    model.train()
    out = model(batch.x[:, :8], batch.edge_index, batch.edge_attr[:, :6])
    res = {"x": batch.x.numpy(), "edge_index": batch.edge_index.numpy(), "edge_attr": batch.edge_attr.numpy(), "ptr": batch.ptr.numpy(),
           "out": out.detach().numpy().copy(), "num_layers": num_layers, "sd_seed": sd_seed,
           "x_mean": stats[0].numpy(), "x_std": stats[1].numpy(), "edge_mean": stats[2].numpy(), "edge_std": stats[3].numpy(),
           "state_dict_keys": np.array(list(model.state_dict().keys()))}
    got = {}
    out.register_hook(lambda g_: got.__setitem__("g", g_.clone()))
    loss = ref_loss(batch, out, stats)
    loss.backward()
    res["loss"] = np.array(loss.item())
    res["grad_out"] = got["g"].numpy().copy()
    for name, p in model.named_parameters():
        res["grad." + name] = p.grad.detach().numpy().copy()
        res["param." + name] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"golden_model_{tag}.npz"), **res)
    print(tag, "out", tuple(out.shape), "loss", res["loss"])


def run_loadsampling():
    """The reference's own samplers (pure numpy: importable) on seeded inputs; the profile step of toy_network.py:100-129 needs pandapower
    objects, so those few numpy lines are re-typed here verbatim over a synthetic load table and fed to the reference's samplers."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_loadsampling", "/root/reference/loadsampling.py")   # not the drop-in of the same name
    ref_ls = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ls)
    rng = np.random.RandomState(7)
    lb = rng.uniform(0.1, 2.0, 9)
    ub = lb * (1.0 + rng.uniform(0.1, 0.9, 9))
    mu = rng.uniform(-1.0, 3.0, 9)
    sig = np.abs(mu) * 0.15
    np.random.seed(21)
    uni = ref_ls.samplermontecarlo(lb, ub, 5)
    np.random.seed(22)
    nor = ref_ls.samplermontecarlo_normal(mu, sig, 5)
    np.random.seed(23)
    uni_s = ref_ls.samplermontecarlo(0.4, 1.7, 6)
    np.random.seed(24)
    nor_s = ref_ls.samplermontecarlo_normal(0.9, 0.2, 6)
    # toy_network.py:83-86, 100-129 with LOAD_DIST = 'normal' (the default) and 'uniform', ITERATIONS = 3
    household = np.array([0.25, 0.2, 0.2, 0.2, 0.2, 0.25, 0.4, 0.65, 0.65, 0.65, 0.7, 0.6, 0.7, 0.65, 0.55, 0.5, 0.45, 0.6, 0.8, 0.9, 0.8, 0.7, 0.55, 0.3])
    industry = np.array([0.35, 0.35, 0.3, 0.3, 0.4, 0.5, 0.6, 0.9, 1., 1., 1., 0.9, 0.85, 0.85, 0.85, 0.85, 0.8, 0.55, 0.5, 0.45, 0.4, 0.4, 0.35, 0.35])
    p_mw = rng.uniform(0.02, 0.6, 7)
    hh_mask = np.array([1., 1., 0., 1., 0., 0., 1.])      # load_r_mask + load_lv_mask
    ind_mask = 1.0 - hh_mask                              # load_ind_mask + load_mv_mask
    load_p = np.stack([hh_mask * (p_mw * household[i]) + ind_mask * (p_mw * industry[i]) for i in range(24)], axis=1)   # :106
    unroll = np.reshape(load_p, load_p.size)                                                                             # :111
    iters, pm_error = 3, 0.3
    np.random.seed(25)
    mc_n = np.reshape(ref_ls.samplermontecarlo_normal(unroll, unroll * (pm_error / 2), iters), [load_p.shape[0], load_p.shape[1] * iters])   # :122,129
    np.random.seed(26)
    mc_u = np.reshape(ref_ls.samplermontecarlo(unroll * (1 - pm_error), unroll * (1 + pm_error), iters), [load_p.shape[0], load_p.shape[1] * iters])
    np.savez_compressed(os.path.join(HERE, "golden_loadsampling.npz"), lb=lb, ub=ub, mu=mu, sig=sig, uni=uni, nor=nor, uni_s=uni_s, nor_s=nor_s,
                        p_mw=p_mw, hh_mask=hh_mask, ind_mask=ind_mask, mc_normal=mc_n, mc_uniform=mc_u, iters=np.array(iters),
                        pm_error=np.array(pm_error), seeds=np.array([21, 22, 23, 24, 25, 26]))
    print("loadsampling:", uni.shape, nor.shape, uni_s.shape, nor_s.shape, mc_n.shape)


def cigre_dataset():
    """Reference data_from_pickles on the same 128 scenarios the fixture holds, after np.random.seed(0)."""
    fix = np.load(os.path.join(HERE, "cigre14_scenarios.npz"), allow_pickle=False)
    S = fix["nodes"].shape[0]
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("nodes", "edges", "labels"):
            with open(f"/root/reference/data/cigre14/{name}", "rb") as fh:
                full = pickle.load(fh)
            with open(os.path.join(tmp, name), "wb") as fh:
                pickle.dump(full[:S], fh)
        with open("/root/reference/data/cigre14/noise_param", "rb") as fh:
            noise = pickle.load(fh)
        with open(os.path.join(tmp, "noise_param"), "wb") as fh:
            pickle.dump(noise, fh)
        np.random.seed(0)
        ds, xm, xs, em, es = ref_data.data_from_pickles(tmp + "/", 8, 6, 4, 2, CIGRE_MEAS_V, CIGRE_MEAS_PF)
    return ds, (xm, xs, em, es)


def main():
    run_loadsampling()
    torch.set_num_threads(1)
    fix = np.load(os.path.join(HERE, "cigre14_scenarios.npz"), allow_pickle=False)
    S = fix["nodes"].shape[0]

    # ---- reference data_from_pickles on the same 128 scenarios the fixture holds ----
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("nodes", "edges", "labels"):
            with open(f"/root/reference/data/cigre14/{name}", "rb") as fh:
                full = pickle.load(fh)
            with open(os.path.join(tmp, name), "wb") as fh:
                pickle.dump(full[:S], fh)
        with open("/root/reference/data/cigre14/noise_param", "rb") as fh:
            noise = pickle.load(fh)
        with open(os.path.join(tmp, "noise_param"), "wb") as fh:
            pickle.dump(noise, fh)
        np.random.seed(0)
        ds, xm, xs, em, es = ref_data.data_from_pickles(tmp + "/", 8, 6, 4, 2, CIGRE_MEAS_V, CIGRE_MEAS_PF)
    stats = (xm, xs, em, es)
    b64 = Batch.from_data_list(ds[:64])
    truth = torch.stack([(b64.y[:, 0] - xm[0]) / xs[0], b64.y[:, 1]], 1)
    kat_truth = ref_loss(b64, truth.clone(), stats).item()
    kat_zeros = ref_loss(b64, torch.zeros_like(truth), stats).item()
    pick = [3, 0, 7, 1, 127]
    bp = Batch.from_data_list([ds[i] for i in pick])
    np.savez_compressed(
        os.path.join(HERE, "golden_dataset_cigre14.npz"),
        x=torch.cat([d.x for d in ds]).numpy(), edge_attr=torch.cat([d.edge_attr for d in ds]).numpy(),
        y=torch.cat([d.y for d in ds]).numpy(), edge_index=ds[0].edge_index.numpy(),
        x_mean=xm.numpy(), x_std=xs.numpy(), edge_mean=em.numpy(), edge_std=es.numpy(),
        kat_loss_truth=np.array(kat_truth), kat_loss_zeros=np.array(kat_zeros),
        pick=np.array(pick), pick_x=bp.x.numpy(), pick_edge_index=bp.edge_index.numpy(),
        pick_edge_attr=bp.edge_attr.numpy(), pick_y=bp.y.numpy(), pick_batch=bp.batch.numpy(), pick_ptr=bp.ptr.numpy())
    print("dataset: KAT truth", kat_truth, "zeros", kat_zeros)

    default = dict(dim_featn=8, dim_feate=6, dim_out=2, dim_hid=32, n_gnn_layers=8, K=2, dropout_rate=0.3, L=5)
    small = dict(dim_featn=8, dim_feate=6, dim_out=2, dim_hid=32, n_gnn_layers=3, K=2, dropout_rate=0.0, L=2)
    mpn = dict(dim_featn=8, dim_feate=6, dim_out=2, dim_hid=32, n_gnn_layers=4, K=2, dropout_rate=0.3)
    skipmpn = dict(dim_featn=8, dim_feate=6, dim_out=8, dim_hid=32, n_gnn_layers=3, K=2, dropout_rate=0.3)

    b4 = Batch.from_data_list(ds[10:14])
    run_model("skippfn_cigre", "SkipPFN", default, b4, stats, sd_seed=1, torch_seed=11)
    run_model("pfn_small_cigre", "PFN", small, Batch.from_data_list(ds[20:23]), stats, sd_seed=2, torch_seed=12)
    run_model("mpn_cigre", "MPN", mpn, Batch.from_data_list(ds[30:35]), stats, sd_seed=3, torch_seed=13)
    run_model("skipmpn_cigre", "SkipMPN", skipmpn, Batch.from_data_list(ds[40:42]), stats, sd_seed=4, torch_seed=14)

    run_gat("gat_cigre", Batch.from_data_list(ds[50:56]), stats, sd_seed=6)
    # the constructor's other runnable settings: no self loops + tanh + averaged single head + another attention slope; ReLU
    run_gat("gat_noloop_tanh_cigre", Batch.from_data_list(ds[56:60]), stats, sd_seed=31, num_layers=4, self_loops=False, nonlin="tanh",
            concat=False, slope=0.1)
    run_gat("gat_relu_cigre", Batch.from_data_list(ds[90:93]), stats, sd_seed=32, num_layers=3, nonlin="relu")
    # multi-head: only concat=False can run in the reference (concat=True changes the channel count between layers)
    run_gat("gat_heads2_cigre", Batch.from_data_list(ds[100:105]), stats, sd_seed=34, num_layers=4, heads=2, concat=False)
    run_gat("gat_heads3_tanh_cigre", Batch.from_data_list(ds[105:108]), stats, sd_seed=35, num_layers=3, heads=3, concat=False, nonlin="tanh",
            self_loops=False)
    run_gine("gine_cigre", Batch.from_data_list(ds[60:65]), stats, sd_seed=8)
    run_gine("gine_traineps_cigre", Batch.from_data_list(ds[65:69]), stats, sd_seed=33, num_layers=4, eps=0.1, train_eps=True)
    run_gnn("gnn_gcn2_cigre", "gcn2", Batch.from_data_list(ds[70:76]), stats, sd_seed=21)
    run_gnn("gnn_tagcn_cigre", "tagcn", Batch.from_data_list(ds[80:85]), stats, sd_seed=22)
    run_gnn("gnn_fagcn_cigre", "fagcn", Batch.from_data_list(ds[94:100]), stats, sd_seed=24)

    # ---- Oberrhein: synthetic scenarios from the product generator, reference model + loss on top ----
    grid = synth.load_grid("ober_sub")
    store = synth.synthetic_store(grid, 16, seed=1234)
    graphs = [Data(**store.graph(s)) for s in range(3)]
    ob = Batch.from_data_list(graphs)
    ostats = (store.x_mean, store.x_std, store.edge_mean, store.edge_std)
    run_model("skippfn_ober", "SkipPFN", dict(default, n_gnn_layers=4, L=2), ob, ostats, sd_seed=5, torch_seed=15)
    run_gat("gat_ober", ob, ostats, sd_seed=7)
    run_gine("gine_ober", ob, ostats, sd_seed=9)
    run_gnn("gnn_gcn2_ober", "gcn2", ob, ostats, sd_seed=23, num_layers=8)
    run_gnn("gnn_fagcn_ober", "fagcn", ob, ostats, sd_seed=25, num_layers=5)
    # a far-from-solution output so that all three soft-constraint penalties are active
    torch.manual_seed(99)
    wild = torch.stack([torch.randn(ob.x.shape[0]) * 3.0, torch.randn(ob.x.shape[0]) * 0.8], 1).requires_grad_(True)
    loss = ref_loss(ob, wild * 1.0, ostats)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "golden_loss_ober_wild.npz"), x=ob.x.numpy(), edge_index=ob.edge_index.numpy(),
                        edge_attr=ob.edge_attr.numpy(), output=wild.detach().numpy(), loss=np.array(loss.item()),
                        grad_out=wild.grad.numpy(), x_mean=ostats[0].numpy(), x_std=ostats[1].numpy(),
                        edge_mean=ostats[2].numpy(), edge_std=ostats[3].numpy())
    print("ober wild loss", loss.item())


if __name__ == "__main__":
    main()
