"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel name.
usage: python tools/ncu_launches.py launches.csv > profiles/<round>_launch_list_summary.txt"""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    name = re.sub(r"^void\s+|<unnamed>::|\(.*$", "", r[ki]).replace("(int)", "")
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off: one eager training step (bench.py --steps 1 --no-graph)")
print(f"# source: {sys.argv[1]}; durations are cold-cache and serialised: compare SHARES")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:58s} launches={n:4d} total_us={t:10.1f} avg_us={t / n:8.1f} share={100 * t / tot:5.1f}%")
print(f"TOTAL_us={tot:.1f}")
