import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "steps", "repeats", "n_gpus") if k in d}, "e2e", d["e2e"]["value"], d["clocks"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
print("parallelism:", d["config"]["parallelism"][:90])
r = d["roofline"]
print("roofline:", r["kernel"][:50], "frac %.3f" % r["frac"], "launch_ms %.4f" % r["launch_ms"], "| adjoint frac %.3f launch_ms %.4f" % (r["adjoint"]["frac"], r["adjoint"]["launch_ms"]))
for k, v in r["kernels"].items():
    print(f"{k[:55]:55s} cold {v['ms_cold']*1e3:8.2f} us hot {v['ms_hot']*1e3:8.2f} us  frac cold {v['frac_of_peak']:.3f} hot {v['frac_of_peak_hot']:.3f} share {v['share_of_step_hot']:.3f}")
s = d.get("secondary") or {}
for k, v in s.items():
    print(k, json.dumps(v)[:600])
