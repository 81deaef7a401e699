"""Not a test: print a readable summary of a bench.py JSON line (python tools/show_bench.py FILE)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("headline", {k: d[k] for k in ("value", "ms_per_step", "launches_per_step", "n_gpus")}, "e2e",
      d["e2e"]["value"])
if d.get("roofline"):
    print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "frac_l2_warm", "frac_cold",
                                                         "traffic")})
print("cpu", d.get("cpu_baseline"), "eager", d.get("gpu_eager_baseline"), "dp", d.get("dp"))
for k in d.get("kernels") or []:
    print(f"  {k['kernel'][:74]:74s} n={k['launches']:3d} warm={k['us_per_launch']:7.1f} "
          f"cold={k['us_per_launch_cold'] or 0:7.1f} share={k['share']:.3f} {k['bound']:6s} "
          f"frac={k['frac']:.2f} warm={k['frac_l2_warm']:.2f}")
for c in d.get("configs") or []:
    print("=====", c.get("name"), c.get("error") or "")
    if "error" in c:
        continue
    if "batches" in c:
        for k, v in c["batches"].items():
            print("  ", k, {x: (round(v[x], 3) if isinstance(v[x], float) else v[x]) for x in v})
        continue
    print("  ", {k: c[k] for k in ("value", "ms_per_step", "launches_per_step", "global_batch",
                                   "kernel_time_sum_ms", "scaling", "n_gpus") if k in c},
          "e2e", round(c["e2e"]["value"], 1))
    if c.get("cpu_baseline"):
        print("   cpu", round(c["cpu_baseline"]["value"], 1), "eager", c.get("gpu_eager_baseline"))
    for k in c.get("kernels") or []:
        print(f"     {k['kernel'][:86]:86s} n={k['launches']:3d} warm={k['us_per_launch']:7.1f} "
              f"cold={k['us_per_launch_cold'] or 0:7.1f} share={k['share']:.3f} {k['bound']:6s} "
              f"frac={k['frac']:.3f}")
