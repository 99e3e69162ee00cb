"""CPU test of the loss kernel's math core: csrc/wls_math.cuh is compiled for the host (g++,
-ffp-contract=off) behind a tiny test harness and compared with the oracle's forward and autograd on
reference-generated vectors.  This pins the hand-derived adjoint without a GPU;
the CUDA kernels that own tiling and reductions around the same functions are tested under -m gpu."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import REG_COEFS, ROOT, assert_fp32_parity, golden_model, load_golden
import dss2_oracle as orc

HERE = os.path.join(ROOT, "tests", "host_math")


@pytest.fixture(scope="module")
def host_lib():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libwls_host.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "wls_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.wls_host_loss.restype = ctypes.c_int
    return lib


def _run(lib, x, ea, ei, out, stats, want_grad=True):
    x = np.ascontiguousarray(x, np.float32)
    ea = np.ascontiguousarray(ea, np.float32)
    ei = np.ascontiguousarray(ei, np.int64)
    out = np.ascontiguousarray(out, np.float32)
    st = np.concatenate([np.asarray(s, np.float32) for s in stats]).astype(np.float32)
    coefs = np.array([REG_COEFS["lam_v"], REG_COEFS["lam_p"], REG_COEFS["lam_pf"], REG_COEFS["lam_reg"]], np.float32)
    nt, et = x.shape[0], ea.shape[0]
    loss = np.zeros(1, np.float32)
    grad = np.zeros((nt, 2), np.float32)
    pf = np.zeros((8, et), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.wls_host_loss(p(x), p(ea), p(ei), p(out), p(st), p(coefs), ctypes.c_int64(nt), ctypes.c_int64(et), p(loss),
                           p(grad) if want_grad else None, p(pf))
    assert rc == 0
    return float(loss[0]), grad, pf


def _oracle(x, ea, ei, out, stats, dtype):
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    o = t(out).clone().requires_grad_(True)
    loss = orc.wls_loss(t(x), t(ea), o, *[t(s) for s in stats], torch.as_tensor(np.asarray(ei)), REG_COEFS)
    loss.backward()
    return loss.detach(), o.grad


def test_all_penalties_active(host_lib):
    z = load_golden("golden_loss_ober_wild.npz")
    stats = [z[k] for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    loss, grad, _ = _run(host_lib, z["x"], z["edge_attr"], z["edge_index"], z["output"], stats)
    l64, g64 = _oracle(z["x"], z["edge_attr"], z["edge_index"], z["output"], stats, torch.float64)
    assert_fp32_parity(loss, z["loss"], l64, "loss")
    assert_fp32_parity(grad, z["grad_out"], g64, "grad_out")


@pytest.mark.parametrize("tag", ["skippfn_cigre", "mpn_cigre", "skippfn_ober"])
def test_model_outputs(host_lib, tag):
    _, _, _, _, _, z = golden_model(tag)
    stats = [z[k] for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    loss, grad, _ = _run(host_lib, z["x"], z["edge_attr"], z["edge_index"], z["out"], stats)
    l64, g64 = _oracle(z["x"], z["edge_attr"], z["edge_index"], z["out"], stats, torch.float64)
    assert_fp32_parity(loss, z["loss"], l64, "loss")
    assert_fp32_parity(grad, z["grad_out"], g64, "grad_out")


def test_near_solution_and_pflow(host_lib):
    """Output = normalised truth: tiny residuals (max cancellation); also the 8 get_pflow outputs."""
    gd = load_golden("golden_dataset_cigre14.npz")
    stats = [gd[k] for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    n, e = 15 * 8, 14 * 8
    x, ea, y = gd["x"][:n], gd["edge_attr"][:e], gd["y"][:n]
    ei = np.concatenate([gd["edge_index"] + 15 * g for g in range(8)], axis=1)
    out = np.stack([(y[:, 0] - stats[0][0]) / stats[1][0], y[:, 1]], 1).astype(np.float32)
    loss, grad, pf = _run(host_lib, x, ea, ei, out, stats)
    l32, g32 = _oracle(x, ea, ei, out, stats, torch.float32)
    l64, g64 = _oracle(x, ea, ei, out, stats, torch.float64)
    assert_fp32_parity(loss, l32, l64, "loss")
    assert_fp32_parity(grad, g32, g64, "grad_out")
    state = torch.stack([torch.from_numpy(out[:, 0]) * float(stats[1][0]) + float(stats[0][0]), torch.from_numpy(out[:, 1])], 1)
    ref = orc.get_pflow(state, torch.from_numpy(ei), torch.from_numpy(x[:, 8:]), torch.from_numpy(ea[:, 6:]))
    ref64 = orc.get_pflow(state.double(), torch.from_numpy(ei), torch.from_numpy(x[:, 8:]).double(), torch.from_numpy(ea[:, 6:]).double())
    for q in range(8):
        assert_fp32_parity(pf[q], ref[q], ref64[q], f"pflow[{q}]")


@pytest.mark.parametrize("phase_shift", [True, False])
def test_pflow_adjoint_vs_oracle_autograd(host_lib, phase_shift):
    """get_pflow is plain autograd code in the reference (data.py:328-390): the adjoint of all eight outputs w.r.t. y = (V, theta),
    `wls_pflow_backward`, against the oracle's autograd; phase_shift=False takes the branch's shift column (data.py:364-365).  The
    Oberrhein batch has a transformer (trafo_pos = 3 quirk) and both loading columns active."""
    z = load_golden("golden_loss_ober_wild.npz")
    x, ea, ei = z["x"], z["edge_attr"], z["edge_index"]
    nt, et = x.shape[0], ea.shape[0]
    rng = np.random.default_rng(5)
    y = np.stack([1.0 + 0.05 * rng.standard_normal(nt), 0.2 * rng.standard_normal(nt)], 1).astype(np.float32)
    gout = rng.standard_normal((8, et)).astype(np.float32)
    node_param = np.ascontiguousarray(x[:, 8:], np.float32)
    edge_param = np.ascontiguousarray(ea[:, 6:], np.float32)
    if not phase_shift:
        assert np.abs(edge_param[:, 5]).max() > 0
    out8 = np.zeros((8, et), np.float32)
    gy = np.zeros((nt, 2), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    eic = np.ascontiguousarray(ei, np.int64)
    rc = host_lib.wls_host_pflow(p(y), p(node_param), ctypes.c_int64(node_param.shape[1]), p(edge_param), p(eic), ctypes.c_int64(nt),
                                 ctypes.c_int64(et), ctypes.c_int(0 if phase_shift else 1), p(gout), p(out8), p(gy))
    assert rc == 0

    def oracle(dtype):
        yt = torch.from_numpy(y).to(dtype).requires_grad_(True)
        outs = orc.get_pflow(yt, torch.from_numpy(eic), torch.from_numpy(node_param).to(dtype), torch.from_numpy(edge_param).to(dtype),
                             phase_shift=phase_shift)
        sum((o * torch.from_numpy(gout[q]).to(dtype)).sum() for q, o in enumerate(outs)).backward()
        return [o.detach() for o in outs], yt.grad

    o32, g32 = oracle(torch.float32)
    o64, g64 = oracle(torch.float64)
    for q in range(8):
        assert_fp32_parity(out8[q], o32[q], o64[q], f"pflow[{q}]")
    assert_fp32_parity(gy, g32, g64, "grad_y")
