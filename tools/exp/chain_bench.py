"""8 consecutive TAG forward layers (one sub-net) at the bench shape, ordinary launches vs per-tile chained launches, on a side stream."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
from dss2 import _lib, batching, synth, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
store = synth.synthetic_store(synth.load_grid("ober_sub"), B, seed=1, device="cuda")
b = batching.pack_batch(store, torch.arange(B, device="cuda"))
g = ops.resolve_graph(b.edge_index, b.x.size(0))
nt, L = b.x.size(0), 8
acts = torch.randn(L + 1, nt, 32, device="cuda") * 0.1
bits = torch.zeros(L, (nt + 63) // 64 * 64, dtype=torch.int32, device="cuda")
w = torch.randn(L, 3, 32, 32, device="cuda") * 0.1
bias = torch.zeros(L, 32, device="cuda")
marks = torch.zeros(L, nt + 256, dtype=torch.int32, device="cuda")
rng = torch.tensor([12345, 7], dtype=torch.long, device="cuda")
side = torch.cuda.Stream()

def run(chain):
    st = _lib.stream()
    for l in range(L):
        args = (g.ref, _lib.ptr(acts[l]), _lib.ptr(w[l]), _lib.ptr(bias[l]), 32, 2, 1, 0.3, 1, _lib.ptr(rng), l, None, None, 0, _lib.ptr(acts[l + 1]), _lib.ptr(bits[l]))
        if chain:
            _lib.check(lib.dss2_tag_fwd_tc2_chain(*args, _lib.ptr(marks[l]) if l < L - 1 else None, _lib.ptr(marks[l - 1]) if l > 0 else None, st), "chain")
        else:
            _lib.check(lib.dss2_tag_fwd_tc2(*args, st), "plain")

with torch.cuda.stream(side):
    for chain in (0, 1, 0, 1):
        for _ in range(3):
            run(chain); rng[1] += 1
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run(chain); rng[1] += 1
        e1.record(); torch.cuda.synchronize()
        out = acts[L].clone()
        print(f"chain={chain}: {e0.elapsed_time(e1) / 10 * 1e3 / L:.2f} us per layer launch ({L} layers back to back)", float(out.abs().sum()))
# debug timeline (temporary build): per launch CTA start / end times
torch.cuda.synchronize()
import numpy as np
with torch.cuda.stream(side):
    run(1); torch.cuda.synchronize()
m = marks.cpu().numpy().astype(np.uint32)
t00 = None
for l in range(L - 1):
    d = m[l, 8192:8192 + 296].reshape(148, 2).astype(np.int64)
    if t00 is None: t00 = d[:, 0].min()
    print(f"layer {l}: CTA starts {d[:,0].min()-t00:7d}..{d[:,0].max()-t00:7d} ns  ends {d[:,1].min()-t00:7d}..{d[:,1].max()-t00:7d} ns")
for l in (0, 3):
    d = m[l, 8192:8192 + 330].astype(np.int64)
    s0 = d[0]
    print(f"layer {l} CTA0: start 0, setup done {d[300]-s0}, tiles (input ready, tile done): " + ", ".join(f"({d[301+2*i]-s0},{d[302+2*i]-s0})" for i in range(10)), " end", d[1]-s0)
