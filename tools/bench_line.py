#!/usr/bin/env python3
"""Run bench.py with the given arguments and print a one-line digest of its JSON line."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", *sys.argv[1:]], capture_output=True, text=True)
lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not lines:
    print("bench failed:", out.stdout[-2000:], out.stderr[-3000:]); sys.exit(1)
d = json.loads(lines[-1]); r = d.get("roofline", {}); inc = d.get("incremental") or {}
print("%s n=%d value %.1f e2e %.1f ms/step %.4f partials %.4f ms (%.0f launches) frac %.3f %.0f GF/s mat %.4f root %.4f inc %.1f us (C replay %.1f us) logL %.6f cpu %s" % (
    d["config"]["workload"], d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("partials_ms_per_step", 0), r.get("launches_per_step", 0),
    r.get("frac", 0), r.get("gflops", 0), r.get("other_kernels_ms_per_step", {}).get("transition_matrices", 0), r.get("other_kernels_ms_per_step", {}).get("root", 0),
    inc.get("us_per_eval", 0), (inc.get("c_abi_replay") or {}).get("us_per_eval", 0), d["logL"], (d.get("cpu_baseline") or {}).get("value")))
