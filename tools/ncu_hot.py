"""Hot source lines of the first kernel in an .ncu-rep: stall samples per CUDA source line (needs -lineinfo and --import-source on).
usage: python tools/ncu_hot.py <report.ncu-rep> [top] [kernel-name substring]"""
import collections, csv, io, subprocess, sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = sys.argv[3] if len(sys.argv) > 3 else None
take = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
files, cur, hdr = collections.OrderedDict(), None, None
seen_kernel = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        take = only is None or only in r[1]
        continue
    if not take:
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur is None or len(r) < len(hdr):
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)          # two "Source" columns: keep the CUDA one (first)
    try:
        smp = int(d["# Samples"])
    except (KeyError, ValueError):
        continue
    if smp == 0 or not d.get("Line No", "").isdigit():
        continue
    key = (cur.split("/")[-1], int(d["Line No"]))
    e = files.setdefault(key, {"samples": 0, "src": d["Source"].strip()[:110], "stalls": collections.Counter(), "inst": 0})
    e["samples"] += smp
    try:
        e["inst"] += int(d["Instructions Executed"])
    except ValueError:
        pass
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                e["stalls"][k[6:]] += int(v)
            except ValueError:
                pass
tot = sum(e["samples"] for e in files.values())
print(f"total samples {tot}")
for (f, ln), e in sorted(files.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ", ".join(f"{k}={v}" for k, v in e["stalls"].most_common(3))
    print(f"{100 * e['samples'] / tot:5.1f}%  {f}:{ln:<4d} inst={e['inst']:<9d} [{st}]  {e['src']}")
