import os, sys, subprocess, json
import numpy as np, torch
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/deep-statistical-solver-for-distribution-system-state-estimation_b200")
import conftest
from conftest import golden_model, REG_COEFS, split_masks
import networks, data
import test_gpu_parity as T
tag = "skippfn_ober"
ctor, kind, sd, grads, masks, z = golden_model(tag)
res = {}
for impl in ("warp", "row"):
    os.environ["DSS2_EA_IMPL"] = impl
    model = getattr(networks, kind)(**ctor); model.load_state_dict(sd, strict=True); model = model.cuda()
    x, ea, ei = [torch.from_numpy(z[k]).cuda() for k in ("x", "edge_attr", "edge_index")]
    st = [torch.from_numpy(z[k]) for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
    model._dss2_masks = split_masks(masks, ctor); model.train()
    out = model(x[:, :8], ei, ea[:, :6])
    loss = data.gsp_wls_edge(input=x[:, :8], edge_input=ea[:, :6], output=out, x_mean=st[0], x_std=st[1], edge_mean=st[2], edge_std=st[3], edge_index=ei,
                             reg_coefs=REG_COEFS, num_samples=None, node_param=x[:, 8:], edge_param=ea[:, 6:])
    loss.backward()
    res[impl] = {n: p.grad.detach().cpu().double() for n, p in model.named_parameters()}
    res[impl]["__out"] = out.detach().cpu().double()
go = torch.from_numpy(z["grad_out"])
_, _, g64 = T._oracle_model(kind, ctor, sd, x.cpu(), ea.cpu(), ei.cpu(), masks, st, go, torch.float64)
print("out diff warp/row", float((res["warp"]["__out"] - res["row"]["__out"]).abs().max()))
for n in g64:
    sc = float(g64[n].abs().max())
    ew, er = float((res["warp"][n] - g64[n]).abs().max()) / sc, float((res["row"][n] - g64[n]).abs().max()) / sc
    if max(ew, er) > 3.5e-6 or "mpns.1.convs.3" in n or "edge_aggr" in n:
        print(f"{n:45s} warp {ew:.2e} row {er:.2e}  |warp-row| {float((res['warp'][n]-res['row'][n]).abs().max())/sc:.2e}")
n = "mpns.1.convs.3.bias"
print((res["row"][n] - g64[n]) / float(g64[n].abs().max()))
print((res["warp"][n] - g64[n]) / float(g64[n].abs().max()))
