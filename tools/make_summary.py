#!/usr/bin/env python3
"""profiles/<R>_summary.md: one table over every bench line of the round, the previous round's figure next to it.
   python tools/make_summary.py r02 r01"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, PREV = (sys.argv[1:] + ["r02", "r01"])[:2]
P = os.path.join(ROOT, "profiles")


def line(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


def fmt(x, nd=1):
    return "—" if x is None else (f"{x:,.{nd}f}" if isinstance(x, float) else str(x))


names = [("cfg2", "cfg 2: gtr_g4_1000x10k (headline)"), ("cfg2_steps20", "cfg 2, `--steps 20` (the driver's command)"),
         ("cfg2_rescaled", "cfg 2, rescaled every evaluation"), ("makona_like", "cfg 4: makona_like_1610x6k"), ("hky1441", "hky_1441x593"),
         ("benchmark1_xml", "benchmark1.xml (1441 taxa × 987 patterns)"), ("benchmark2_xml", "benchmark2.xml (62 taxa × 10869 patterns)"),
         ("codon", "cfg 3: codon_mg94_500x5k (61 states)"), ("codon_g4", "codon_mg94_500x5k_g4"), ("aa20", "aa20_g4_500x5k (20 states)")]
rows = []
for key, label in names:
    d, o = line(os.path.join(P, f"{R}_bench_{key}.json")), line(os.path.join(P, f"{PREV}_bench_{key}.json"))
    if d is None:
        continue
    r = d.get("roofline") or {}
    cpu = d.get("cpu_baseline") or {}
    inc = d.get("incremental") or {}
    rep = inc.get("c_abi_replay") or {}
    rows.append("| {} | {} | {} | {} | {} | {} {} | {} | {} | {} | {} / {} | {} |".format(
        label, fmt(d["value"]), fmt(o["value"]) if o else "—", fmt(d["e2e"]["value"]), fmt(o["e2e"]["value"]) if o else "—",
        fmt(r.get("partials_ms_per_step"), 4), f"({r.get('launches_per_step', 0):.0f})", r.get("bound", "—"), fmt(r.get("frac"), 3),
        fmt(r.get("dram_frac"), 3) if r.get("dram_frac") is not None else "—",
        fmt(inc.get("us_per_eval")), fmt(rep.get("us_per_eval")), fmt(cpu.get("value")) + (f" ({cpu.get('cores')} thr)" if cpu else "")))

with open(os.path.join(P, f"{R}_summary.md"), "w") as f:
    f.write(f"# {R} — bench lines on one B200 (`tools/collect_round_evidence.sh {R}`; raw lines: `profiles/{R}_bench_*.json`)\n\n")
    f.write("`value` = evaluations/s, data resident, median 500-step block (DESIGN.md §5); `e2e` = synchronous public calls with host buffers;\n"
            "`partials` = CUDA-event time of one step's `updatePartials` launches (count in brackets); `frac` = algorithmic bytes (S ≤ 20) or flops\n"
            "(S > 20) per second ÷ the measured peak (`MEASURED_PEAKS.json` HBM copy 6562.6 GB/s; fp64 tensor pipe 37.1 TFLOP/s from\n"
            f"`profiles/{R}_fp64_peaks.json`); `dram` = ncu DRAM bytes of the same launches (`profiles/{R}_traffic.json`) ÷ live time ÷ HBM peak;\n"
            "incremental = one branch changed (≈13 ops): µs per evaluation through Python ctypes / through the C replay driver;\n"
            "CPU = `oracle/beagle_cpu.c` on the box's host cores (kind \"port\").\n\n")
    f.write(f"| workload | value | {PREV} | e2e | {PREV} | partials ms (launches) | bound | frac | dram | incremental µs (py / C) | CPU port evals/s |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
    f.write("\n".join(rows) + "\n")
    ref = line(os.path.join(P, f"{R}_bench_cfg2_reference.json"))
    if ref:
        cb = ref.get("cpu_baseline") or {}
        f.write(f"\n`--impl reference` arm (cfg 2): {ref['value']:.1f} evaluations/s on {cb.get('cores')} threads ({cb.get('kind')}); "
                f"median step {ref['ms_per_step']:.2f} ms.\n")
    pk = line(os.path.join(P, f"{R}_fp64_peaks.json"))
    if pk:
        f.write(f"\nMeasured peaks on this box (`tools/fp64_peaks.cu`): DFMA {pk['dfma_tflops']} TFLOP/s, DMMA m8n8k4 {pk['dmma_m8n8k4_tflops']} TFLOP/s, "
                f"256-bit streaming write {pk['write_gbs']} GB/s, copy (read+write) {pk['copy_gbs']} GB/s.\n")
    for extra, title in ((f"{R}_gpu_tests.txt", "GPU test suite"), (f"{R}_sanitizer.txt", "compute-sanitizer")):
        path = os.path.join(P, extra)
        if os.path.exists(path):
            f.write(f"\n{title} (`profiles/{extra}`): " + " / ".join(l.strip() for l in open(path).read().strip().splitlines()[-2:]) + "\n")
print(open(os.path.join(P, f"{R}_summary.md")).read())
