#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the DSS2 hot path on B200.

Metric (BASELINE.json): train scenarios/s = scenarios pushed through one full training step
(batch packer -> SkipPFN(8,6,2,32,8,2,0.3,5) forward with in-kernel dropout -> fused branch-flow/WLS loss forward+backward ->
GNN backward (recompute) -> gradient reduction [-> NCCL all-reduce] -> flat Adamax) per second, whole job.
Workload: BASELINE config 3, Oberrhein feeder (N=70 buses, E=69 closed branches), B=4096 scenarios per GPU per step,
synthetic scenarios from the deterministic generator (dss2.synth; the reference's Oberrhein scenario pickles are missing
from its repository), random-init weights.  Weak scaling: every rank owns its own shard and batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line on rank 0 (see the task's contract): value = device-resident throughput, e2e = through the host-buffer
API (pinned host scenarios -> H2D -> step -> D2H loss) inside the timed region, roofline = the dominant kernel (TAG-layer
backward) timed live with CUDA events against MEASURED_PEAKS.json, cpu_baseline = the oracle port of the reference path on
the host cores (bounded sample), clocks = SM clocks / throttle reasons sampled during the timed region.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200")
sys.path.insert(0, PKG)

CASE = "ober_sub"
REG = {"mu_v": 1e-1, "mu_theta": 1e-1, "lam_v": 1e-4, "lam_p": 1e-8, "lam_pf": 1e-6, "lam_reg": 1e2}
WORKLOAD = "ober_sub (N=70,E=69) B=4096/GPU SkipPFN(8,6,2,32,8,2,0.3,5) train step: pack+fwd+WLS loss+bwd+Adamax"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--scenarios", type=int, default=16384, help="synthetic scenarios resident per GPU")
    ap.add_argument("--cpu-sample", type=int, default=256, help="graphs per step of the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--same-shards", action="store_true", help="every rank uses rank 0's scenarios and ids (multi-GPU == single-GPU check)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (reference networks.py/data.py semantics in plain torch, oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_graphs, steps, warmup, seed=1234):
    """scenarios/s of fwd + gsp_wls_edge + bwd + Adamax on the host cores for `sample_graphs` Oberrhein scenarios per step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dss2_oracle as orc
    from dss2 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    store = synth.synthetic_store(synth.load_grid(CASE), sample_graphs, seed=seed)
    batch = orc.collate([store.graph(i) for i in range(sample_graphs)])
    stats = {"x_mean": store.x_mean, "x_std": store.x_std, "edge_mean": store.edge_mean, "edge_std": store.edge_std}
    sd = {k: v.requires_grad_(True) for k, v in orc.init_state_dict("SkipPFN", seed=0).items()}
    opt_state = {}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.train_step(sd, batch, stats, REG, 0.3, opt_state)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return sample_graphs * len(times) / total, cores, 1000.0 * total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, cores, ms = cpu_reference_rate(args.cpu_sample, args.steps, max(args.warmup, 1))
    sample = f"{args.cpu_sample} Oberrhein scenarios per step (same model/loss/optimizer), oracle port of reference networks.py+data.py"
    line = {
        "impl": "reference", "metric": "train scenarios/s", "value": rate, "unit": "scenarios/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "cpu_sample_graphs_per_step": args.cpu_sample},
        "cpu_baseline": {"value": rate, "unit": "scenarios/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "scenarios/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, threading.Event(), [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as exc:   # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dss2 import _lib, synth
    from dss2.dataset import ScenarioStore
    from dss2.trainer import GraphedTrainer, default_spec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    K, W, B = args.steps, max(args.warmup, 3), args.batch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    grid = synth.load_grid(CASE)
    shard = 0 if args.same_shards else rank
    store = synth.synthetic_store(grid, args.scenarios, seed=1234 + shard, device=dev)
    n, e = store.max_nodes, store.max_edges
    trainer = GraphedTrainer(store, B, spec=default_spec(), reg_coefs=REG, seed=0, process_group=pg, world_size=world,
                             use_cuda_graph=not args.no_graph).capture()
    gen = torch.Generator().manual_seed(99 + shard)
    ids_host = torch.randint(0, args.scenarios, (W + K, B), generator=gen).pin_memory()

    # ---- device-resident throughput: scenarios already in HBM, ids copied per step (32 KB) ----
    for i in range(W):
        trainer.step(ids_host[i])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()      # lets `ncu --profile-from-start off` see exactly the timed steps
    ev0.record()
    for i in range(K):
        trainer.step(ids_host[W + i])
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    sampler.stop_flag.set()
    sampler.join()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    loss_end = float(trainer.loss.item())
    value = world * B * K / (ms_total / 1e3)

    # ---- layer kernels timed live on the launching stream (hidden layer 32 -> 32, K = 2, the shapes of 35 of the 40 TAG layers) ----
    lib, P = _lib.load(), _lib.ptr
    from dss2 import ops as _ops
    sp, run, bufs = trainer.spec, trainer.runner, trainer.bufs
    nt, et = trainer.nt, trainer.et
    name_w, name_b = "mpns.0.convs.3.lins.0.weight", "mpns.0.convs.3.bias"
    w_off, b_off = run.table[name_w][0], run.table[name_b][0]
    part_w = ctypes.c_void_p(bufs["partials"].data_ptr() + 4 * w_off)
    gref, st_ = trainer.graph.ref, _lib.stream
    x_l, y_l, bits_l = bufs["acts"][0, 3], bufs["acts"][0, 4], bufs["bits"][0, 3]
    gy_l, gx_l, lvl = bufs["g32"][0], bufs["g32"][1], bufs["lvl"]
    wp, bp = run._p(trainer.flat, name_w), run._p(trainer.flat, name_b)
    tc2 = _ops.TAG_IMPL == "tc2" and bool(lib.dss2_tag_tc2_supported(gref, sp.K))
    kernels = {}
    if tc2:
        kernels["k_tag_tc2<BGX> (TAG backward-to-input, tcgen05)"] = (lambda: _lib.check(lib.dss2_tag_bwd_tc2_gx(
            gref, wp, 32, sp.K, 1, sp.p_drop, P(bits_l), P(gy_l), P(gx_l), P(lvl), lvl.numel() * 4, st_()), "gx"),
            nt * (128 + 128 + 4 + 24), "grad_y + sign word + ELL topology in, grad_x out; excludes the 256 B/node hop-level spill it writes for k_tag_gw")
        kernels["k_tag_tc2<FWD> (TAG forward, tcgen05)"] = (lambda: _lib.check(lib.dss2_tag_fwd_tc2(
            gref, P(x_l), wp, bp, 32, sp.K, 1, sp.p_drop, 1, P(trainer.step_state), 3, None, None, 0, P(y_l), P(bits_l), st_()), "fwd"),
            nt * (128 + 128 + 4 + 24), "x + ELL topology in, y + sign word out")
        in_step = " [in the step]"
        kernels["k_tag_gw (TAG weight gradients, tcgen05 MN-major, TMA ring)" + (in_step if _ops.GW_IMPL == "tc" else " [alternative]")] = (
            lambda: _lib.check(lib.dss2_tag_bwd_tc2_gw(
                nt, P(x_l), 32, sp.K, 1, sp.p_drop, P(bits_l), P(gy_l), part_w, run.flat_size, b_off - w_off, P(lvl), lvl.numel() * 4, st_()), "gw"),
            nt * (128 + 128 + 4), "x + grad_y + sign word in; excludes the 256 B/node hop levels it re-reads")
        kernels["k_tag_gw_ffma (TAG weight gradients, exact fp32 FFMA, TMA ring)" + (in_step if _ops.GW_IMPL != "tc" else " [alternative]")] = (
            lambda: _lib.check(lib.dss2_tag_gw_ffma(
                nt, P(x_l), 32, sp.K, 1, sp.p_drop, P(bits_l), P(gy_l), part_w, run.flat_size, b_off - w_off, P(lvl), lvl.numel() * 4, st_()), "gwf"),
            nt * (128 + 128 + 4), "x + grad_y + sign word in; excludes the 256 B/node hop levels it re-reads")
    else:
        kernels["k_tag_bwd<2,32> (TAG backward, CUDA cores)"] = (lambda: _lib.check(lib.dss2_tag_bwd(
            gref, P(x_l), wp, 32, sp.K, 1, sp.p_drop, P(bits_l), P(gy_l), P(gx_l), part_w, run.flat_size, b_off - w_off, st_()), "bwd"),
            nt * (3 * 128 + 4) + 4 * (nt + 1) + 16 * et, "x, grad_y in, grad_x out, sign word, CSR")
        kernels["k_tag_fwd<2> (TAG forward, CUDA cores)"] = (lambda: _lib.check(lib.dss2_tag_fwd(
            gref, P(x_l), wp, bp, 32, sp.K, 1, sp.p_drop, 1, P(trainer.step_state), 3, None, None, 0, P(y_l), P(bits_l), st_()), "fwd"),
            nt * (2 * 128 + 4) + 4 * (nt + 1) + 16 * et, "x in, y out, sign word, CSR")

    # EdgeAggregation of sub-net 1 (input = the previous sub-net's 8-wide output, gradient to the input and skip path needed)
    pre1 = "mpns.1.edge_aggr.edge_aggr."
    ea_w = [run._p(trainer.flat, pre1 + n_) for n_ in ("0.weight", "0.bias", "2.weight", "2.bias")]
    ea_part = ctypes.c_void_p(bufs["partials"].data_ptr() + 4 * run.table[pre1 + "0.weight"][0])
    x_in, eattr = bufs["outs"][0], trainer.batch["edge_attr"]
    ea_fwd_bytes = 32 * nt + 52 * et + 16 * et + 128 * nt
    kernels["k_edgeagg_fwd (EdgeAggregation forward)"] = (lambda: _lib.check(lib.dss2_edgeagg_fwd(
        gref, P(x_in), sp.fn, sp.fn, P(eattr), 13, sp.fe, *ea_w, P(bufs["acts"][1, 0]), st_()), "ea_fwd"),
        ea_fwd_bytes, "SURVEY 8(d): x row 32 + out row 128 per bus, edge_attr row 52 + edge_index 16 per branch")
    kernels["k_edgeagg_bwd (EdgeAggregation backward: grad_x + parameter gradients)"] = (lambda: _lib.check(lib.dss2_edgeagg_bwd(
        gref, P(x_in), sp.fn, sp.fn, P(eattr), 13, sp.fe, *ea_w, P(gy_l), P(bufs["gsub"][0]), sp.fn, P(bufs["gsub"][1]), ea_part,
        run.flat_size, st_()), "ea_bwd"),
        ea_fwd_bytes + 32 * nt, "SURVEY 8(d): forward bytes with grad_out in place of out, + grad_x row 32 per bus")

    # the fused loss (2 kernels: reduction pass + gradient pass), on the trainer's own buffers
    kernels["k_wls<false>+k_wls<true> (branch flows + WLS loss, forward and backward)"] = (lambda: _lib.check(lib.dss2_wls_fwd_bwd(
        gref, P(trainer.batch["x"]), 11, P(trainer.batch["edge_attr"]), 13, P(bufs["outs"][-1]), P(trainer.stats), REG["lam_v"], REG["lam_p"],
        REG["lam_pf"], REG["lam_reg"], P(trainer.batch["vminmax"]), 1, P(trainer.loss), None, P(trainer.grad_out), P(trainer.wls_ws),
        trainer.wls_ws.numel(), st_()), "wls"), 2 * (60 * nt + 68 * et), "x row 44 + out 8 + grad_out 8 per bus, edge_attr row 52 + edge_index 16 per branch, two passes")

    def time_kernel(fn, reps=30):
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
        durs = []
        for _ in range(3):
            fn()
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            durs.append(a.elapsed_time(b))
        return statistics.mean(durs) * 1e-3

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath))
    timed = []
    for kname, (fn, nbytes, what) in kernels.items():
        t = time_kernel(fn)
        timed.append({"kernel": kname, "us_per_launch": t * 1e6, "algorithmic_bytes_per_launch": nbytes, "bytes_counted": what,
                      "achieved": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / peak, "traffic": traffic_tab.get(kname.split(" ")[0])})
    dom = max((r for r in timed if "[alternative]" not in r["kernel"]), key=lambda r: r["us_per_launch"])
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": dom["traffic"], "peak_source": peak_src, "us_per_launch": dom["us_per_launch"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "bytes_counted": dom["bytes_counted"],
                "note": "layer kernels are bound by the per-tile dependency chain (shared-memory gathers, barriers, tcgen05 issue), not by HBM: "
                        "see DESIGN.md section 4 and profiles/",
                "all_layer_kernels": timed}
    # whole-step view with SURVEY.md 8(d)'s per-layer algorithmic bytes (topology counted as 0: per-topology template)
    hid, fn = 32, sp.fn
    step_bytes = 2 * (60 * nt + 68 * et)                                     # loss, two passes
    for s_ in range(sp.L):
        cin0 = 44 if s_ == 0 else 4 * fn                                     # EdgeAggregation reads the 11-wide x rows in sub-net 0
        step_bytes += (cin0 * nt + 52 * et + 16 * et + 4 * hid * nt) + (cin0 * nt + 52 * et + 16 * et + 4 * hid * nt + (4 * fn * nt if s_ else 0))
        for l_ in range(sp.n_layers):
            cout_ = (sp.dim_out if s_ == sp.L - 1 else fn) if l_ == sp.n_layers - 1 else hid
            step_bytes += 4 * nt * (hid + cout_) + 4 * nt * (2 * hid + cout_)  # TAG forward + recompute-style backward
    roofline["step"] = {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms_total / K / 1e3) / 1e9, "unit": "GB/s",
                        "frac": step_bytes / (ms_total / K / 1e3) / 1e9 / peak,
                        "definition": "SURVEY.md 8(d): loss 2x(60 Nt + 68 Et); EdgeAggregation fwd+bwd; TAG 4 Nt (Cin+Cout) fwd + 4 Nt (2 Cin+Cout) bwd per layer"}

    # ---- end to end through the host-buffer API: pinned host scenarios -> H2D -> step -> D2H loss, all inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # Two slots of B scenarios in ONE staging store: step i trains on slot i % 2 (selected by the scenario ids the packer gathers)
        # while a copy stream lands step i+1's scenarios in the other slot - the input pipeline a user of the API would write.
        slots = 2 if args.scenarios >= 2 * B else 1
        SB = slots * B
        stage = ScenarioStore(x=torch.empty(SB * n, 11, device=dev), edge_attr=torch.empty(SB * e, 13, device=dev),
                              y=torch.zeros(SB * n, 2, device=dev), edge_index=store.edge_index[:, :SB * e].contiguous(),
                              node_off=store.node_off[:SB + 1].contiguous(), edge_off=store.edge_off[:SB + 1].contiguous(),
                              x_mean=store.x_mean, x_std=store.x_std, edge_mean=store.edge_mean, edge_std=store.edge_std,
                              max_nodes=n, max_edges=e)
        stage.x.copy_(store.x[:SB * n])
        stage.edge_attr.copy_(store.edge_attr[:SB * e])
        t2 = GraphedTrainer(stage, B, spec=default_spec(), reg_coefs=REG, seed=0, process_group=pg, world_size=world,
                            use_cuda_graph=not args.no_graph).capture()
        nbuf = 4
        host_x = [store.x[i * B * n:(i + 1) * B * n].cpu().pin_memory() for i in range(nbuf)]
        host_ea = [store.edge_attr[i * B * e:(i + 1) * B * e].cpu().pin_memory() for i in range(nbuf)]
        host_loss = torch.zeros(K + W, dtype=torch.float32).pin_memory()
        slot_ids = [torch.arange(B, device=dev) + sl * B for sl in range(slots)]
        copy_stream = torch.cuda.Stream(device=dev)
        slot_free = [torch.cuda.Event() for _ in range(slots)]     # the last step that read the slot has finished
        slot_ready = [torch.cuda.Event() for _ in range(slots)]    # the slot's H2D copies have landed

        def feed(i):   # host -> device copy of step i's scenarios
            sl = i % slots
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(slot_free[sl])
                stage.x[sl * B * n:(sl + 1) * B * n].copy_(host_x[i % nbuf], non_blocking=True)
                stage.edge_attr[sl * B * e:(sl + 1) * B * e].copy_(host_ea[i % nbuf], non_blocking=True)
                slot_ready[sl].record(copy_stream)

        def e2e_step(i):
            sl = i % slots
            if slots == 1:
                feed(i)
            else:
                feed(i + 1)                                  # overlaps this step's compute
            torch.cuda.current_stream().wait_event(slot_ready[sl])
            t2.step(slot_ids[sl])
            slot_free[sl].record()
            host_loss[i:i + 1].copy_(t2.loss.reshape(1), non_blocking=True)

        if slots == 2:
            feed(0)
        for i in range(W):
            e2e_step(i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            e2e_step(W + i)
        b.record()
        barrier()
        ms_e2e = max_over_ranks(a.elapsed_time(b))
        e2e = {"value": world * B * K / (ms_e2e / 1e3), "unit": "scenarios/s", "h2d_bytes_per_step": (B * n * 11 + B * e * 13) * 4,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K, "api": "GraphedTrainer.step on a staging store fed from pinned host memory; " +
               ("two slots, the H2D copy of step i+1 runs on a copy stream while step i computes" if slots == 2 else "one slot, copies serialised"),
               "h2d_copies_in_timed_region": K}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, ms = cpu_reference_rate(args.cpu_sample, 3, 1)
        cpu = {"value": rate, "unit": "scenarios/s", "cores": cores, "kind": "port", "ms_per_step": ms,
               "sample": f"{args.cpu_sample} Oberrhein scenarios per step x 3 timed steps (fwd+loss+bwd+Adamax), oracle port of the "
                         "reference's networks.py/data.py semantics in eager torch on the host cores; dead O(N^2) Laplacian omitted"}

    if rank == 0:
        line = {
            "metric": "train scenarios/s", "value": value, "unit": "scenarios/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "nodes_per_step_per_gpu": nt, "edges_per_step_per_gpu": et,
                       "resident_scenarios_per_gpu": args.scenarios, "cuda_graph": not args.no_graph, "tag_fwd_impl": __import__("dss2.ops", fromlist=["x"]).TAG_IMPL,
                       "l2": "per-step working set (saved activations 1.47 GB + 110 MB scenario store) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world}" if world > 1 else "single"},
            "e2e": e2e, "gpu_launches": int(trainer.launches_per_step) * K, "launches_per_step": int(trainer.launches_per_step),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": sampler.result(), "final_loss": loss_end,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # a process group that has collectives captured in live CUDA graphs can hang in destroy_process_group(): leave together, hard
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
