#!/bin/bash
# One digest line per workload (GPU box): bash tools/bench_all.sh [out-file] [extra bench args]
OUT=${1:-gpurun_out/bench_all.txt}; shift
mkdir -p gpurun_out; : > $OUT
for w in gtr_g4_1000x10k gtr_g4_1000x10k_rescaled makona_like_1610x6k hky_1441x593 benchmark1_xml benchmark2_xml codon_mg94_500x5k aa20_g4_500x5k; do
  steps=300; case $w in codon*|aa20*) steps=100;; hky*|benchmark*) steps=1000;; esac
  python tools/bench_line.py --workload $w --steps $steps --warmup 10 --no-cpu-baseline "$@" >> $OUT 2>&1
done
cat $OUT
