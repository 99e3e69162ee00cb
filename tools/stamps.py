"""Diagnostic: per-phase clock64 totals of one worker thread / the MMA issuer of CTA 0 of the tcgen05 layer kernels.
Needs the library built with -DDSS2_STAMPS (tools/stamps.sh does that on the GPU box)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200"))
import torch  # noqa: E402
from dss2 import _lib, synth  # noqa: E402
from dss2.trainer import GraphedTrainer, default_spec  # noqa: E402

NAMES = ["tile top / wait input", "store L0", "publish L0", "hop 1", "wait buffer 1", "store L1", "publish L1", "hop 2", "wait buffer 2", "store L2",
         "publish L2", "prefetch/rng", "wait MMAs", "TMEM ld", "epilogue"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda", 0)
    store = synth.synthetic_store(synth.load_grid("ober_sub"), B, seed=1, device=dev)
    tr = GraphedTrainer(store, B, spec=default_spec(), seed=0, use_cuda_graph=False)
    lib, P = _lib.load(), _lib.ptr
    getter = lib.dss2_tc2_stamps if os.environ.get("DSS2_TC3") == "0" else lib.dss2_tc3_stamps
    getter.restype = ctypes.c_int
    getter.argtypes = [ctypes.c_void_p]
    run, bufs, sp = tr.runner, tr.bufs, tr.spec
    tr._enqueue(with_optimizer=False)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    getter(buf)
    name_w, name_b = "mpns.0.convs.3.lins.0.weight", "mpns.0.convs.3.bias"
    wp, bp = run._p(tr.flat, name_w), run._p(tr.flat, name_b)
    g = tr.graph
    print("tiles", g.c.num_tiles, "graphs/tile", g.c.graphs_per_tile, "max rows", g.c.max_tile_nodes)
    x_l, y_l, bits_l = bufs["acts"][0, 3], bufs["acts"][0, 4], bufs["bits"][0, 3]
    gy_l, gx_l, lvl = bufs["g32"][0], bufs["g32"][1], bufs["lvl"]
    reps = 5

    def show(title, tiles_cta0):
        torch.cuda.synchronize()
        getter(buf)
        v = [int(buf[i]) for i in range(32)]
        tot = sum(v[:15])
        n = reps * tiles_cta0
        print(f"== {title}: worker thread 0 of CTA 0, cycles per tile (total {tot / n:.0f})")
        for i, nm in enumerate(NAMES):
            print(f"   {nm:22s} {v[i] / n:9.0f}")
        print("   issuer: " + "  ".join(f"L{k}: wait {v[16 + 2 * k] / n:.0f} issue {v[17 + 2 * k] / n:.0f}" for k in range(3)))

    grid_cap = {96: 3, 128: 2}.get(0, 1)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rows = g.c.max_tile_nodes
    per_sm = 3 if rows <= 96 else (2 if rows <= 128 else 1)
    grid = min(g.c.num_tiles, per_sm * sms)
    tiles0 = (g.c.num_tiles - 1) // grid + 1
    for _ in range(reps):
        _lib.check(lib.dss2_tag_fwd_tc2(g.ref, P(x_l), wp, bp, 32, sp.K, 1, sp.p_drop, 1, P(tr.step_state), 3, None, None, 0, P(y_l),
                                        P(bits_l), _lib.stream()), "fwd")
    show("FWD", tiles0)
    for _ in range(reps):
        _lib.check(lib.dss2_tag_bwd_tc2_gx(g.ref, wp, 32, sp.K, 1, sp.p_drop, P(bits_l), P(gy_l), P(gx_l), P(lvl), lvl.numel() * 4,
                                           _lib.stream()), "gx")
    show("BGX", tiles0)


if __name__ == "__main__":
    main()
