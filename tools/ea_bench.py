"""EdgeAggregation forward / backward kernel times at the bench shape (Oberrhein B=4096), CUDA events, L2 flushed between launches.
Usage: python tools/ea_bench.py [B]   (DSS2_EA_IMPL=warp selects the round-1 warp-per-row kernels)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
from dss2 import _lib, batching, synth, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
store = synth.synthetic_store(synth.load_grid("ober_sub"), B, seed=1, device="cuda")
b = batching.pack_batch(store, torch.arange(B, device="cuda"))
g = ops.resolve_graph(b.edge_index, b.x.size(0))
nt = b.x.size(0)
torch.manual_seed(0)
w1, b1, w2, b2 = [torch.randn(*s, device="cuda") * 0.2 for s in ((32, 22), (32,), (32, 32), (32,))]
out = torch.empty(nt, 32, device="cuda")
gout = torch.randn(nt, 32, device="cuda")
gx = torch.empty(nt, 8, device="cuda")
npart = lib.dss2_num_partials()
count = 32 * 22 + 32 + 1024 + 32
part = torch.zeros(npart, count, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = _lib.stream()

def fwd():
    _lib.check(lib.dss2_edgeagg_fwd(g.ref, _lib.ptr(b.x), 11, 8, _lib.ptr(b.edge_attr), 13, 6, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(out), st), "fwd")

def bwd():
    _lib.check(lib.dss2_edgeagg_bwd(g.ref, _lib.ptr(b.x), 11, 8, _lib.ptr(b.edge_attr), 13, 6, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(gout),
                                    None, 0, _lib.ptr(gx), _lib.ptr(part), count, st), "bwd")

import ctypes
PA = ctypes.c_void_p * 1
def upload():
    _lib.check(lib.dss2_edgeagg_upload(0, 1, PA(_lib.ptr(w1)), PA(_lib.ptr(b1)), PA(_lib.ptr(w2)), PA(_lib.ptr(b2)), 8, 6, st), "upload")

def fwd_slot():
    _lib.check(lib.dss2_edgeagg_fwd_slot(g.ref, _lib.ptr(b.x), 11, 8, _lib.ptr(b.edge_attr), 13, 6, 0, _lib.ptr(out), st), "fwd_slot")

def bwd_slot():
    _lib.check(lib.dss2_edgeagg_bwd_slot(g.ref, _lib.ptr(b.x), 11, 8, _lib.ptr(b.edge_attr), 13, 6, 0, _lib.ptr(gout), None, 0, _lib.ptr(gx), _lib.ptr(part), count, st), "bwd_slot")

def timeit(f, n=10):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

print(f"impl={os.environ.get('DSS2_EA_IMPL', 'row')} B={B} Nt={nt} tiles={g.c.num_tiles} rows/tile<={g.c.max_tile_nodes}")
print(f"edgeagg fwd {timeit(fwd):.1f} us   bwd {timeit(bwd):.1f} us   (incl. the weight upload nodes)")
if os.environ.get("DSS2_EA_IMPL", "row")[0] != "w" and lib.dss2_edgeagg_slots_ok(g.ref, 11, 13, 6):
    o1, gx1, p1 = out.clone(), gx.clone(), part.sum(0)
    upload()
    print(f"prepared weights: upload {timeit(upload):.1f} us   fwd {timeit(fwd_slot):.1f} us   bwd {timeit(bwd_slot):.1f} us")
    print("slot API == pointer API:", torch.equal(o1, out), torch.equal(gx1, gx), torch.equal(p1, part.sum(0)))
