import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from conftest import golden_model, split_masks
import networks
from dss2 import ops
ctor, kind, sd, grads, masks, z = golden_model("skippfn_cigre")
x, ea, ei = [torch.from_numpy(z[k]).cuda() for k in ("x", "edge_attr", "edge_index")]
model = networks.SkipPFN(**ctor); model.load_state_dict(sd); model = model.cuda()
runner, pack = model._machinery()
flat = pack.gather(dict(model.named_parameters()))
graph = ops.resolve_graph(ei, x.size(0))
m = [[t.cuda().to(torch.uint8).contiguous() for t in sub] for sub in split_masks(masks, ctor)]
out = {}
for impl in ("ffma", "tc"):
    ops.TAG_FWD_IMPL = impl
    bufs = runner.alloc(x.size(0), x.device, need_grad=False)
    bufs["acts"].fill_(float("nan")); bufs["bits"].fill_(-7)
    runner.forward(graph, x, 11, ea, 13, flat, bufs, drop_mode=2, masks=m)
    torch.cuda.synchronize()
    out[impl] = (bufs["acts"].clone(), bufs["bits"].clone(), [o.clone() for o in bufs["outs"]])
A, B = out["ffma"], out["tc"]
for s in range(5):
    for l in range(8):
        da = float((A[0][s, l] - B[0][s, l]).abs().max())
        nb = int((A[1][s, l] != B[1][s, l]).sum()) if l < 7 else -1
        nanb = int(torch.isnan(B[0][s, l]).sum())
        print(f"s{s} l{l}: max|dact| {da:.2e}  bits differ {nb}  nan {nanb}  bits==-7: {int((B[1][s,l]==-7).sum()) if l<7 else -1}")
    print(f"  outs[{s}] diff {float((A[2][s]-B[2][s]).abs().max()):.2e}")
