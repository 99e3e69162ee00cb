"""Validation metrics of the reference script on the device: one kernel per batch instead of the ~40 eager ops,
two `get_pflow` calls, `nonzero()` filtering and torchmetrics objects of dss2_run.py:183-209."""
import math

import torch

from . import _lib
from .ops import require_cuda, stage_rows

_WS = {}


def evaluate_batch(out, y, x, edge_index, edge_attr, x_mean, x_std):
    """Metrics of ONE validation batch, i.e. the terms dss2_run.py:186-205 adds up per batch (the script then averages them over
    the loader).  out [Nt,2] = model output (normalised V, raw theta; NOT modified), y [Nt,2] labels, x [Nt,11] and edge_attr [Et,13]
    the full feature tensors of the batch, x_mean / x_std the dataset statistics.  Returns a dict of python floats:
    rmse_v, mae_v, rmse_th, mae_th, rmse_loading, mae_loading, rmse_loading_trafos, mae_loading_trafos, prop_std_v, prop_std_th."""
    require_cuda()
    lib = _lib.load()
    og, os_ = stage_rows(out.detach())
    yg, ys = stage_rows(y)
    xg, xs = stage_rows(x)
    eg, es = stage_rows(edge_attr)
    if xg.size(1) < 11 or eg.size(1) < 13:
        raise _lib.Dss2Error("evaluate_batch needs the full x [Nt,11] and edge_attr [Et,13] tensors of the batch")
    ei = (edge_index if edge_index.device.type == "cuda" else edge_index.cuda(non_blocking=True)).long().contiguous()
    dev = xg.device
    with torch.cuda.device(dev):
        ws = _WS.get(dev.index)
        if ws is None:
            ws = _WS[dev.index] = torch.zeros(lib.dss2_eval_workspace_bytes(), dtype=torch.uint8, device=dev)
        vmm = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.dss2_col_minmax(_lib.ptr(xg), xs, 8, xg.size(0), _lib.ptr(vmm), _lib.stream()), "dss2_col_minmax")
        sums = torch.empty(19, dtype=torch.float64, device=dev)
        _lib.check(lib.dss2_eval_metrics(_lib.ptr(ei), xg.size(0), ei.size(1), _lib.ptr(xg), xs, _lib.ptr(eg), es, _lib.ptr(og), os_,
                                         _lib.ptr(yg), ys, float(x_mean[0]), float(x_std[0]), _lib.ptr(vmm), _lib.ptr(sums),
                                         _lib.ptr(ws), ws.numel(), _lib.stream()), "dss2_eval_metrics")
    s = sums.cpu().tolist()
    n = float(xg.size(0))

    def std(sx, sxx):   # torch.std: unbiased
        return math.sqrt(max((sxx - sx * sx / n) / (n - 1.0), 0.0)) if n > 1 else float("nan")

    def ratio(a, b):
        return a / b if b != 0.0 else float("nan")

    return {
        "rmse_v": math.sqrt(s[0] / n), "mae_v": s[1] / n, "rmse_th": math.sqrt(s[2] / n), "mae_th": s[3] / n,
        "rmse_loading": math.sqrt(ratio(s[13], s[12])), "mae_loading": ratio(s[14], s[12]),
        "rmse_loading_trafos": math.sqrt(ratio(s[16], s[15])), "mae_loading_trafos": ratio(s[17], s[15]),
        "prop_std_v": 100.0 * ratio(std(s[4], s[5]), std(s[8], s[9])), "prop_std_th": 100.0 * ratio(std(s[6], s[7]), std(s[10], s[11])),
    }
