"""Developer helper: print the interesting numbers of a bench.py JSON line."""
import json
import sys


def brief(x):
    s = {k: round(x[k], 2) for k in ("value", "ms_per_step") if k in x}
    for k in ("e2e", "e2e_rgbaf", "cpu_baseline", "psnr", "roofline", "scaling_strong"):
        if k in x:
            s[k] = {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in x[k].items()
                    if kk in ("value", "gpu", "reference", "delta_db", "ok", "frac", "achieved", "ms_per_step", "e2e")}
    return s


for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e)
        continue
    print(f, "n_gpus", d.get("n_gpus"), d["config"]["workload"])
    print("  ", brief(d))
    for s in d.get("secondary", []):
        print("  ", s["config"]["workload"])
        print("     ", brief(s))
    print("  clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
