import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from conftest import golden_model, split_masks, REG_COEFS
import networks, data
from dss2 import ops
ctor, kind, sd, grads, masks, z = golden_model("skippfn_cigre")
x, ea, ei = [torch.from_numpy(z[k]).cuda() for k in ("x", "edge_attr", "edge_index")]
st = [torch.from_numpy(z[k]) for k in ("x_mean", "x_std", "edge_mean", "edge_std")]
res = {}
for impl in ("ffma", "tc"):
    ops.TAG_FWD_IMPL = impl
    model = networks.SkipPFN(**ctor); model.load_state_dict(sd); model = model.cuda()
    model._dss2_masks = split_masks(masks, ctor)
    out = model(x[:, :8], ei, ea[:, :6]); ob = out.detach().clone()
    loss = data.gsp_wls_edge(input=x[:, :8], edge_input=ea[:, :6], output=out, x_mean=st[0], x_std=st[1], edge_mean=st[2], edge_std=st[3],
                             edge_index=ei, reg_coefs=REG_COEFS, num_samples=None, node_param=x[:, 8:], edge_param=ea[:, 6:])
    loss.backward()
    res[impl] = (ob, loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()})
print("out diff", float((res["tc"][0] - res["ffma"][0]).abs().max()), "loss", res["tc"][1], res["ffma"][1])
worst = []
for n in res["tc"][2]:
    a, b = res["tc"][2][n], res["ffma"][2][n]
    worst.append((float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30), n))
for w, n in sorted(worst, reverse=True)[:8]: print(f"{w:.3e} {n}")
for w, n in sorted(worst)[:3]: print(f"{w:.3e} {n}")
