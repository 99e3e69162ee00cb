"""BASELINE configs[0] and [1] end to end: does the GPU path TRAIN like the reference?

configs[0]  CIGRE MV 14-bus train + test on the reference's own scenarios (the 128-scenario fixture of its data/cigre14 pickles, run
            through the reference's feature pipeline bit-exactly): SkipPFN(8,6,2,32,8,2,.,5) + gsp_wls_edge + Adamax(3e-3) for EPOCHS
            epochs through the throughput tier (GraphedTrainer: packer -> forward -> fused loss -> backward -> flat Adamax, one CUDA
            graph per step) against the same loop on the CPU oracle: per-step loss curve, then the script's validation metrics
            (dss2_run.py:183-209) of both trained models on the held-out scenarios.
configs[1]  the trained weights evaluated on synthetic scenarios of the re-switched (meshed) CIGRE topology: generalisation eval.

Dropout is switched off in both runs (the reference draws its masks from torch's global generator; with p = 0 the two runs are
comparable step by step).  Adamax divides every gradient entry by a running max of its magnitude, so fp32 rounding differences in
near-zero gradient entries turn into +-lr parameter differences: two correct fp32 implementations drift apart slowly.  The
tolerances below are that drift, as measured on B200 (recorded in gpurun_out/TRAINING.json -> profiles/TRAINING_r02.json).
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import REG_COEFS, ROOT, load_golden, permute_edges

import dss2_oracle as orc

EPOCHS = 30
BATCH = 56          # 112 training scenarios = 2 steps per epoch, 16 held out


def _cigre_store():
    from dss2 import dataset, synth
    fx = load_golden("cigre14_scenarios.npz")
    grid = synth.load_grid("cigre14")
    zn, ze = dataset.reference_noise_stream(0, fx["nodes"].shape[0], 15, 14)
    return dataset.build_scenario_store(fx["nodes"], fx["edges"], fx["labels"], grid["noise_param"], grid["meas_v"], grid["meas_pflow"], zn, ze)


def _script_metrics(out, b, x_mean, x_std):
    """dss2_run.py:183-206 for one batch on the CPU oracle (torchmetrics' MeanAbsoluteError = mean |a - b|)."""
    import torch.nn.functional as F
    out = torch.concat([out[:, 0:1] * x_std[:1] + x_mean[:1], out[:, 1:]], axis=1)
    out[:, 1:] *= (1. - b["x"][:, 9:10])
    y = b["y"]
    mae = lambda a, c: float((a - c).abs().mean())
    tl, tt = orc.get_pflow(y, b["edge_index"], b["x"][:, 8:], b["edge_attr"][:, 6:])[0:2]
    ol, ot = orc.get_pflow(out, b["edge_index"], b["x"][:, 8:], b["edge_attr"][:, 6:])[0:2]
    tl2, ol2, tt2, ot2 = tl[tl.nonzero()], ol[tl.nonzero()], tt[tt.nonzero()], ot[tt.nonzero()]
    return {"rmse_v": float(torch.sqrt(F.mse_loss(out[:, :1], y[:, :1]))), "mae_v": mae(out[:, :1], y[:, :1]),
            "rmse_th": float(torch.sqrt(F.mse_loss(out[:, 1:], y[:, 1:]))), "mae_th": mae(out[:, 1:], y[:, 1:]),
            "rmse_loading": float(torch.sqrt(F.mse_loss(ol2, tl2))), "mae_loading": mae(ol2, tl2),
            "rmse_loading_trafos": float(torch.sqrt(F.mse_loss(ot2, tt2))), "mae_loading_trafos": mae(ot2, tt2)}


@pytest.mark.gpu
def test_cigre14_training_curve_and_reswitched_eval_match_the_oracle():
    from dss2 import batching, metrics, synth
    from dss2.trainer import GraphedTrainer, default_spec
    import networks
    store = _cigre_store()
    stats = {"x_mean": store.x_mean, "x_std": store.x_std, "edge_mean": store.edge_mean, "edge_std": store.edge_std}
    sd0 = orc.init_state_dict("SkipPFN", seed=4)
    spec = default_spec(p_drop=0.0)
    tr = GraphedTrainer(store.to("cuda"), BATCH, spec=spec, reg_coefs=REG_COEFS, lr=3e-3, seed=0, init_state_dict=sd0, use_cuda_graph=True).capture()
    ref = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt_state = {}
    batches = [orc.collate([store.graph(i) for i in range(lo, lo + BATCH)]) for lo in (0, BATCH)]
    # a SECOND oracle run on the same batches with their edge lists in another order: mathematically the same training run, other
    # fp32 rounding.  How far two runs of the reference itself drift apart is the yardstick for our distance to the reference.
    ref2 = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt_state2 = {}
    batches2 = [permute_edges(b_, 7 + i_) for i_, b_ in enumerate(batches)]
    ids = [torch.arange(lo, lo + BATCH, device="cuda") for lo in (0, BATCH)]
    torch.set_num_threads(os.cpu_count() or 1)
    ours, theirs, theirs2 = [], [], []
    for epoch in range(EPOCHS):
        for k in range(2):
            ours.append(float(tr.step(ids[k]).item()))
            theirs.append(float(orc.train_step(ref, batches[k], stats, REG_COEFS, 0.0, opt_state)))
            theirs2.append(float(orc.train_step(ref2, batches2[k], stats, REG_COEFS, 0.0, opt_state2)))
    ours, theirs, theirs2 = np.array(ours), np.array(theirs), np.array(theirs2)
    rel = np.abs(ours - theirs) / np.abs(theirs)
    rel_ref = np.abs(theirs2 - theirs) / np.abs(theirs)       # reference vs reference (other summation order)

    # ---- configs[0]: held-out metrics of both trained models ----
    test_ids = list(range(2 * BATCH, store.num_scenarios))
    model = networks.SkipPFN(8, 6, 2, 32, 8, 2, 0.0, 5)
    model.load_state_dict({k: v.cpu() for k, v in tr.state_dict().items()}, strict=True)
    model = model.cuda().eval()
    frozen = {k: v.detach() for k, v in ref.items()}
    frozen2 = {k: v.detach() for k, v in ref2.items()}

    def eval_both(st, sel):
        b_gpu = batching.pack_batch(st.to("cuda"), sel)
        with torch.no_grad():
            out = model(b_gpu.x[:, :8], b_gpu.edge_index, b_gpu.edge_attr[:, :6])
        m_ours = metrics.evaluate_batch(out, b_gpu.y, b_gpu.x, b_gpu.edge_index, b_gpu.edge_attr, st.x_mean, st.x_std)
        b = orc.collate([st.graph(i) for i in sel])
        with torch.no_grad():
            o = orc.pfn_forward(frozen, b["x"][:, :8], b["edge_index"], b["edge_attr"][:, :6], 0.0, skip=True)
            o2 = orc.pfn_forward(frozen2, b["x"][:, :8], b["edge_index"], b["edge_attr"][:, :6], 0.0, skip=True)
        return m_ours, _script_metrics(o, b, st.x_mean, st.x_std), _script_metrics(o2, b, st.x_mean, st.x_std)

    held_ours, held_theirs, held_theirs2 = eval_both(store, test_ids)
    # untrained model on the same held-out scenarios: what "it trains" is measured against
    b0 = orc.collate([store.graph(i) for i in test_ids])
    with torch.no_grad():
        o0 = orc.pfn_forward(sd0, b0["x"][:, :8], b0["edge_index"], b0["edge_attr"][:, :6], 0.0, skip=True)
    held_init = _script_metrics(o0, b0, store.x_mean, store.x_std)

    # ---- configs[1]: generalisation to the re-switched topology ----
    rs = synth.synthetic_store(synth.load_grid("cigre14_reswitched"), 64, seed=21)
    rs_ours, rs_theirs, rs_theirs2 = eval_both(rs, list(range(64)))

    record = {"epochs": EPOCHS, "steps": len(ours), "batch": BATCH, "loss_ours": ours.tolist(), "loss_oracle": theirs.tolist(),
              "loss_oracle_other_edge_order": theirs2.tolist(), "rel_diff_per_step": rel.tolist(),
              "rel_diff_per_step_oracle_vs_oracle_other_edge_order": rel_ref.tolist(), "held_out_init": held_init, "held_out_ours": held_ours,
              "held_out_oracle": held_theirs, "held_out_oracle_other_edge_order": held_theirs2, "reswitched_ours": rs_ours,
              "reswitched_oracle": rs_theirs, "reswitched_oracle_other_edge_order": rs_theirs2}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "TRAINING.json"), "w") as f:
        json.dump(record, f, indent=1)

    # the loss curves: the first step agrees to rounding.  Afterwards the curves separate: measured on B200 two fp32 runs of the reference
    # itself (edge order permuted) are up to 3 % apart over these 60 steps, ours up to 12 % from the reference at isolated late steps
    # where the loss (~0.01) is dominated by step-to-step noise - the 3xTF32 transforms carry ~2^-21 per operand against fp32's
    # 2^-24, and Adamax turns the rounding of near-zero gradient entries into +-lr parameter steps.  Bounds: 20 % anywhere, 5 % median.
    assert rel[0] <= 1e-4, rel[:3]
    assert rel.max() <= 0.2 and float(np.median(rel)) <= 5e-2, (int(rel.argmax()), float(rel.max()), float(np.median(rel)), float(rel_ref.max()))
    # it trains: the loss falls by orders of magnitude, and the held-out voltage error ends below the untrained model's
    assert ours[-1] < 1e-2 * ours[0] and theirs[-1] < 1e-2 * theirs[0], (ours[0], ours[-1], theirs[0], theirs[-1])
    assert held_ours["rmse_v"] < held_init["rmse_v"] and held_theirs["rmse_v"] < held_init["rmse_v"]
    # both trained models agree on every validation metric of the script, on the held-out scenarios and on the re-switched grid:
    # within 3x the spread of the two reference runs, or 6 %
    for name, a, b, b2 in (("held-out", held_ours, held_theirs, held_theirs2), ("reswitched", rs_ours, rs_theirs, rs_theirs2)):
        for k in b:
            tol = max(3.0 * abs(b2[k] - b[k]), 6e-2 * abs(b[k]), 1e-7)
            assert abs(a[k] - b[k]) <= tol, (name, k, a[k], b[k], b2[k])
