"""Shared test plumbing: the `gpu` marker, import paths, golden-file helpers."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

REG_COEFS = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}  # dss2_run.py:104-112


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# every assert_fp32_parity call of a GPU run is recorded and written to gpurun_out/PARITY.json (copied to profiles/PARITY_rNN.json)
_PARITY_LOG = []
_CURRENT_TEST = {"id": ""}


@pytest.fixture(autouse=True)
def _remember_test_id(request):
    _CURRENT_TEST["id"] = request.node.nodeid
    yield


def note_parity(what, **fields):
    """Extra record in the parity log (e.g. a ReLU-tie case that switched the gradient reference)."""
    _PARITY_LOG.append(dict(test=_CURRENT_TEST["id"], what=what, **fields))


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY_LOG or not torch.cuda.is_available():
        return
    import json
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    rows = [r for r in _PARITY_LOG if "err_over_tol" in r]
    worst = sorted(rows, key=lambda r: -r["err_over_tol"])[:25]
    summary = {"criterion": "err = max|ours - fp64 oracle| <= tol = max(rtol * max|fp64|, noise_mult * max|reference fp32 - fp64|), rtol = 1e-5",
               "device": torch.cuda.get_device_name(0), "calls": len(rows), "max_err_over_tol": max((r["err_over_tol"] for r in rows), default=0.0),
               "max_err_over_scale": max((r["err_over_scale"] for r in rows), default=0.0),
               "calls_decided_by_rtol_alone": sum(1 for r in rows if r["err_over_scale"] <= r["rtol"]),
               "notes": [r for r in _PARITY_LOG if "err_over_tol" not in r], "worst_25": worst, "all": rows}
    with open(os.path.join(out_dir, "PARITY.json"), "w") as f:
        json.dump(summary, f, indent=1)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_model(tag):
    """Golden model file as (ctor dict, kind, state_dict, grads dict, masks list-of-arrays, raw npz)."""
    z = load_golden(f"golden_model_{tag}.npz")
    ctor = {}
    for k, v in z["ctor"]:
        ctor[str(k)] = float(v) if str(k) == "dropout_rate" else int(float(v))
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    grads = {k[len("grad."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.") }
    nt = z["x"].shape[0]
    nm = int(z["num_masks"])
    masks = None
    if nm:
        bits = np.unpackbits(z["masks"])[: nm * nt * ctor["dim_hid"]].reshape(nm, nt, ctor["dim_hid"])
        masks = [torch.from_numpy(b.astype(np.float32)) for b in bits]
    return ctor, str(z["kind"]), sd, grads, masks, z


def split_masks(masks, ctor):
    """Flat list of recorded masks -> per-sub-net lists (n_gnn_layers-1 masks per sub-net)."""
    if masks is None:
        return None
    per = ctor["n_gnn_layers"] - 1
    return [masks[i * per:(i + 1) * per] for i in range(len(masks) // per)]


def assert_fp32_parity(ours, ref32, ref64, what="", rtol=1e-5, noise_mult=4.0, sum_of=None):
    """The fp32 parity criterion used throughout (hard part 1).

    `ref32` is the reference result in fp32, `ref64` the same computation in fp64 (the arbiter).  Two
    correct fp32 implementations differ by rounding noise that the cancellation-heavy flow equations
    amplify, so a candidate passes when its distance to the fp64 truth is at most
    max(rtol * scale, noise_mult * |ref32 - ref64|_max): i.e. within `rtol` relative, or no worse than
    `noise_mult` times the reference's own fp32 rounding error on that tensor.  (noise_mult = 4: entries such as the
    last-layer theta-bias gradient are sums of +-1e3-sized terms that cancel to ~0; ours and the reference's result are two
    draws of the same rounding noise, and a ratio of 2-3 between two draws is ordinary.)

    sum_of: when the tensor IS a plain sum of given terms over their first axis (a bias gradient = column sum of the upstream gradient
    over the buses), pass the terms: the scale of the rtol criterion is then max_c sum_n |term[n, c]| - the forward error bound of a
    floating-point sum is relative to the sum of the magnitudes, not to a result the terms may cancel to."""
    ours = torch.as_tensor(ours).double().cpu()
    ref64 = torch.as_tensor(ref64).double().cpu()
    # `ref32` may be a list of fp32 evaluations of the reference that are equally valid (e.g. the same batch with its edge list in
    # another order: PyG's scatter order follows the edge order, its CUDA scatter is unordered): the noise is the worst of them
    refs = ref32 if isinstance(ref32, (list, tuple)) else [ref32]
    refs = [torch.as_tensor(r).double().cpu() for r in refs]
    scale = float(ref64.abs().max()) if ref64.numel() else 0.0
    if sum_of is not None:
        scale = max(scale, float(torch.as_tensor(sum_of).double().abs().sum(0).max()))
    noise = max(float((r - ref64).abs().max()) for r in refs) if ref64.numel() else 0.0
    err = float((ours - ref64).abs().max()) if ref64.numel() else 0.0
    # the noise term is only meaningful when the reference's fp32 result and the fp64 oracle describe the SAME computation: if they
    # are far apart (round 1: GAT / GINE goldens recorded over a shim whose Sequential dropped the activations, 100 % apart from the
    # oracle) the criterion would accept anything - fail loudly instead
    # (50 %: the deliberately cancelling cases - residuals and gradients AT the solution - legitimately show fp32-vs-fp64
    # differences of 20 % of a near-zero scale; a wrong reference is off by 100 % and more)
    assert noise <= 0.5 * scale + 1e-300, (f"{what}: the reference's fp32 result and the fp64 oracle disagree by {noise:.3e} "
                                            f"(scale {scale:.3e}): not rounding noise, one of them is wrong")
    tol = max(rtol * scale, noise_mult * noise)
    _PARITY_LOG.append({"test": _CURRENT_TEST["id"], "what": what, "err": err, "tol": tol, "scale": scale, "ref32_noise": noise, "rtol": rtol,
                        "noise_mult": noise_mult, "err_over_tol": err / (tol + 1e-300), "err_over_scale": err / (scale + 1e-300)})
    assert err <= tol + 1e-300, f"{what}: err {err:.3e} > tol {tol:.3e} (scale {scale:.3e}, ref32 noise {noise:.3e})"
    return err, tol


def permute_edges(batch, seed):
    """The same collated batch (dict of the oracle's `collate`) with its edge list in another order: mathematically the same graph,
    another fp32 summation order in every scatter of the reference."""
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(batch["edge_index"].shape[1], generator=g)
    out = dict(batch)
    out["edge_index"] = batch["edge_index"][:, perm].contiguous()
    out["edge_attr"] = batch["edge_attr"][perm].contiguous()
    return out


def permuted_golden(z, seed):
    """dict view of a golden npz with the edge list in another order (for the oracle_*_run helpers)."""
    d = {k: z[k] for k in ("x", "edge_attr", "edge_index", "x_mean", "x_std", "edge_mean", "edge_std")}
    d.update({k: z[k] for k in z.files if k.startswith("opt_")})
    pb = permute_edges({"edge_index": torch.from_numpy(z["edge_index"]), "edge_attr": torch.from_numpy(z["edge_attr"])}, seed)
    d["edge_index"], d["edge_attr"] = pb["edge_index"].numpy(), pb["edge_attr"].numpy()
    return d


def golden_gat(tag):
    """GAT_DSSE golden file as (num_layers, state_dict, grads dict, raw npz)."""
    z = load_golden(f"golden_model_{tag}.npz")
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    grads = {k[len("grad."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
    return int(z["num_layers"]), sd, grads, z


def oracle_gat_run(orc, num_layers, sd, z, dtype):
    """Oracle GAT_DSSE forward + WLS loss + autograd on the inputs of a golden file."""
    x, ea, ei = torch.from_numpy(z["x"]).to(dtype), torch.from_numpy(z["edge_attr"]).to(dtype), torch.from_numpy(z["edge_index"])
    st = [torch.from_numpy(z[k]).to(dtype) for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    out = orc.gat_dsse_forward(p, x[:, :8], ei, ea[:, :6], num_layers, **gat_options(z))
    loss = orc.wls_loss(x, ea, out, *st, ei, REG_COEFS)
    loss.backward()
    return out.detach(), loss.detach(), {k: v.grad for k, v in p.items()}


def gat_options(z):
    """Constructor options a GAT golden was recorded with (older files: the script's defaults)."""
    if "opt_nonlin" not in (z.files if hasattr(z, "files") else z):
        return {}
    return {"nonlin": str(z["opt_nonlin"]), "slope": float(z["opt_slope"]), "self_loops": bool(z["opt_self_loops"])}


def gat_heads(z):
    return int(z["opt_heads"]) if "opt_heads" in (z.files if hasattr(z, "files") else z) else 1


def oracle_gine_run(orc, num_layers, sd, z, dtype):
    """Oracle GINE_DSSE forward + WLS loss + autograd on the inputs of a golden file."""
    x, ea, ei = torch.from_numpy(z["x"]).to(dtype), torch.from_numpy(z["edge_attr"]).to(dtype), torch.from_numpy(z["edge_index"])
    st = [torch.from_numpy(z[k]).to(dtype) for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    out = orc.gine_dsse_forward(p, x[:, :8], ei, ea[:, :6], num_layers)
    loss = orc.wls_loss(x, ea, out, *st, ei, REG_COEFS)
    loss.backward()
    return out.detach(), loss.detach(), {k: v.grad for k, v in p.items()}


def oracle_gnn_run(orc, z, sd, dtype, grad_out=None):
    """Oracle gnn_dsse forward + (WLS loss | sum(out * grad_out)) + autograd on the inputs of a golden file."""
    x, ea, ei = torch.from_numpy(z["x"]).to(dtype), torch.from_numpy(z["edge_attr"]).to(dtype), torch.from_numpy(z["edge_index"])
    st = [torch.from_numpy(z[k]).to(dtype) for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    out = orc.gnn_dsse_forward(p, x[:, :8], ei, int(z["num_layers"]), model=str(z["model"]), K=int(z["K"]))
    loss = orc.wls_loss(x, ea, out, *st, ei, REG_COEFS) if grad_out is None else (out * grad_out.to(dtype)).sum()
    loss.backward()
    return out.detach(), loss.detach(), {k: v.grad for k, v in p.items()}
