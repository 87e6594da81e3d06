#!/bin/bash
# One GPU-box pass that regenerates everything under profiles/ for a round:  bash tools/collect_round_evidence.sh r01
R=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/${R}_gpu_tests.txt; cat gpurun_out/${R}_gpu_tests.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1        # alignment cache
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk4 -s 30 -c 4 -o gpurun_out/${R}_walk4_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
B200_WALK_VARIANT=2 B200_TENSOR_R=4 ncu --set full --clock-control none --import-source on -k regex:k_walk4t -s 30 -c 1 -o gpurun_out/${R}_walk4t_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python bench.py --workload codon_mg94_500x5k --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk_mma -s 8 -c 1 -o gpurun_out/${R}_walk_mma_codon_full \
    python bench.py --workload codon_mg94_500x5k --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python bench.py 2>&1 | tail -1 > gpurun_out/${R}_bench_cfg2.json
python bench.py --impl reference --steps 30 --warmup 3 2>&1 | tail -1 > gpurun_out/${R}_bench_cfg2_reference.json
python bench.py --workload codon_mg94_500x5k --steps 200 --warmup 5 --cpu-budget 8 2>&1 | tail -1 > gpurun_out/${R}_bench_codon.json
python bench.py --workload hky_1441x593 --steps 1000 --warmup 10 --cpu-budget 4 2>&1 | tail -1 > gpurun_out/${R}_bench_hky1441.json
python bench.py --workload makona_like_1610x6k --steps 500 --warmup 10 --cpu-budget 6 2>&1 | tail -1 > gpurun_out/${R}_bench_makona_like.json
python bench.py --workload benchmark1_xml --steps 1000 --warmup 10 --cpu-budget 4 2>&1 | tail -1 > gpurun_out/${R}_bench_benchmark1_xml.json
python bench.py --workload benchmark2_xml --steps 1000 --warmup 10 --cpu-budget 4 2>&1 | tail -1 > gpurun_out/${R}_bench_benchmark2_xml.json
python bench.py --workload gtr_g4_1000x10k_rescaled --steps 500 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${R}_bench_cfg2_rescaled.json
python bench.py --workload codon_mg94_500x5k_g4 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${R}_bench_codon_g4.json
python tools/bench_gradient.py > /dev/null 2>&1
WORKLOAD=codon_mg94_500x5k STEPS=5 python tools/bench_gradient.py > /dev/null 2>&1
STEPS=100 python tools/bench_partitions.py 2>&1 | tail -1 > gpurun_out/${R}_bench_partitions.json
bash tools/sanitize.sh > gpurun_out/${R}_sanitizer.txt 2>&1; cat gpurun_out/${R}_sanitizer.txt
for f in gpurun_out/${R}_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], "UNREADABLE", e); sys.exit(0)
r = d.get("roofline", {}); c = d.get("cpu_baseline") or {}
print(sys.argv[1].split("/")[-1], "value %.1f e2e %.1f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
      "partials_ms", r.get("partials_ms_per_step"), "frac", r.get("frac"), "gflops", r.get("gflops"),
      "cpu %s (%s thr)" % (c.get("value"), c.get("cores")), "inc_us", (d.get("incremental") or {}).get("us_per_eval"))
PY
done
