"""Time the EdgeAggregation tile kernels on the bench workload for several CTA sizes (DSS2_EA_{FWD,BWD}_THREADS).
usage (GPU box): python tools/ea_sweep.py [B]"""
import ctypes, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch
from dss2 import _lib, synth
from dss2.trainer import GraphedTrainer, default_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
store = synth.synthetic_store(synth.load_grid("ober_sub"), 8192, seed=1, device="cuda")
tr = GraphedTrainer(store, B, spec=default_spec(), reg_coefs={"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2},
                    seed=0, use_cuda_graph=False)
tr._enqueue(with_optimizer=False)
torch.cuda.synchronize()
run, sp, bufs, g = tr.runner, tr.spec, tr.bufs, tr.graph.ref
P = _lib.ptr
pre = sp.prefix_fmt.format(s=1)
names = [pre + "edge_aggr.edge_aggr.0.weight", pre + "edge_aggr.edge_aggr.0.bias", pre + "edge_aggr.edge_aggr.2.weight", pre + "edge_aggr.edge_aggr.2.bias"]
w = [run._p(tr.flat, n) for n in names]
xin, ea = bufs["outs"][0], tr.batch["edge_attr"]
out, gy, gprev = bufs["acts"][1, 0], bufs["g32"][0], bufs["gsub"][0]
part = ctypes.c_void_p(bufs["partials"].data_ptr() + 4 * run.table[names[0]][0])
st = _lib.stream


def fwd():
    _lib.check(lib.dss2_edgeagg_fwd(g, P(xin), sp.fn, sp.fn, P(ea), 13, sp.fe, *w, P(out), st()), "fwd")


def bwd():
    _lib.check(lib.dss2_edgeagg_bwd(g, P(xin), sp.fn, sp.fn, P(ea), 13, sp.fe, *w, P(gy), None, 0, P(gprev), part, run.flat_size, st()), "bwd")


def timeit(fn, reps=20):
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for _ in range(3):
        fn()
    d = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        d.append(a.elapsed_time(b) * 1e3)
    return statistics.mean(d)


ref = {}
for nt in (256, 384, 512):
    os.environ["DSS2_EA_FWD_THREADS"] = str(nt)
    os.environ["DSS2_EA_BWD_THREADS"] = str(nt)
    tf, tb = timeit(fwd), timeit(bwd)
    torch.cuda.synchronize()
    res = (out.clone(), gprev.clone(), bufs["partials"].sum(0).clone())
    if not ref:
        ref = res
    dev = [float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(res, ref)]
    print(f"threads={nt}: fwd {tf:7.1f} us  bwd {tb:7.1f} us   max rel dev vs 256-thread run: out {dev[0]:.1e} gx {dev[1]:.1e} partial sums {dev[2]:.1e}", flush=True)
