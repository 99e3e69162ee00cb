"""Summarise an .ncu-rep: headline metrics, stall reasons, and the executed-SASS opcode histogram per node row.
usage: python tools/ncu_ops.py <report.ncu-rep> [rows]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else 286720.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, d = rr[0], dict(zip(rr[0], rr[2]))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in d:
        print(f"{w:72s} {d[w]}")
st = []
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("stalls/issue:", ", ".join(f"{k}={v:.2f}" for v, k in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(sr) if "Source" in r and "Instructions Executed" in r]
h = sr[hi[0]]
si, ei = h.index("Source"), h.index("Instructions Executed")
sec = sr[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(sr))]
agg, tot = collections.Counter(), 0
for r in sec:
    try:
        n = int(r[ei])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
    op = ".".join((m.group(2) if m else r[si][:20]).split(".")[:2])
    agg[op] += n
    tot += n
print(f"warp-instructions per row: {tot / rows_n:.1f}")
print("  ".join(f"{op}={n / rows_n:.1f}" for op, n in agg.most_common(28)))
