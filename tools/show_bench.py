"""Print the headline fields of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:
        print(path, "unreadable:", exc)
        continue
    e2e = d.get("e2e") or {}
    print(f"{path}: n_gpus={d.get('n_gpus')} value={d.get('value'):.0f} ms/step={d.get('ms_per_step'):.3f} e2e={e2e.get('value')} "
          f"loss={d.get('final_loss')} cfg={ {k: d['config'].get(k) for k in ('tag_fwd_impl',)} }")
    for r in (d.get("roofline") or {}).get("all_layer_kernels", []):
        print(f"    {r['kernel'][:58]:58s} {r['us_per_launch']:7.1f} us  frac={r['frac']:.3f}")
    if d.get("cpu_baseline"):
        print("    cpu_baseline", d["cpu_baseline"]["value"], "cores", d["cpu_baseline"]["cores"])
