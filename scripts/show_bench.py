import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], d["clocks"], "cpu", d["cpu_baseline"]["value"])
for k, v in d["roofline"]["kernels"].items():
    print(f"{k:55s} {v['ms']*1e3:8.2f} us  {v['gbs']:8.1f} GB/s  frac {v['frac_of_peak']:.3f} share {v['share_of_step']:.3f}")
