#!/bin/bash
# Kernel-variant sweep for the 4-state walk (run on the GPU box): one bench line per setting.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1   # builds the alignment cache
for cfg in "0 1 128 12" "1 1 32 12" "1 1 64 12" "1 1 128 12" "1 1 256 12" "1 0 128 12"; do
  set -- $cfg
  echo "variant=$1 reorder=$2 block=$3 depth=$4"
  B200_WALK_VARIANT=$1 B200_REORDER=$2 B200_WALK_BLOCK=$3 B200_STACK_DEPTH=$4 \
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('  value %.1f evals/s  e2e %.1f  ms/step %.4f  walk %.4f ms  %.0f GB/s (frac %.3f)  %.1f GF/s  mat %.4f root %.4f logL %.6f' % (
            d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['gflops'],
            d['roofline']['other_kernels_ms_per_step']['transition_matrices'], d['roofline']['other_kernels_ms_per_step']['root'], d['logL']))
    else:
        sys.stdout.write(l)
"
done 2>&1 | tee gpurun_out/sweep_walk.txt
