#!/bin/bash
# Variant sweep of the eigen-form 4-state walk (walk4e.cu) against the matrix-form kernel, one digest line per setting.
#   usage (GPU box):  bash tools/sweep_walk4e.sh [workload] [out-file]
W=${1:-gtr_g4_1000x10k}
OUT=${2:-gpurun_out/r02_sweep_walk4e.txt}
mkdir -p gpurun_out
python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1     # alignment cache
: > $OUT
run() {   # label, env assignments...
  local label=$1; shift
  echo -n "$label | " >> $OUT
  env "$@" python tools/bench_line.py --workload $W --steps ${STEPS:-200} --warmup 10 --no-cpu-baseline --no-extras >> $OUT 2>&1
}
run "matrix form k_walk4 (round 1)    " B200_EIGEN_WALK=0
run "eigen async-staged R4 minb4      " B200_TIP_MODE=3
run "eigen async-staged R4 minb5      " B200_TIP_MODE=3 B200_WALK_MINB=5
run "eigen async-staged R4 minb6      " B200_TIP_MODE=3 B200_WALK_MINB=6
run "eigen async-staged R4 minb3      " B200_TIP_MODE=3 B200_WALK_MINB=3
run "eigen async-staged R2 minb4      " B200_TIP_MODE=3 B200_WALK_R=2
run "eigen async-staged R2 minb6      " B200_TIP_MODE=3 B200_WALK_R=2 B200_WALK_MINB=6
run "eigen async-staged R4 minb5 ovs2 " B200_TIP_MODE=3 B200_WALK_MINB=5 B200_PHASE_OVERSUB=2
run "eigen async-staged R4 minb5 nola " B200_TIP_MODE=3 B200_WALK_MINB=5 B200_LOOKAHEAD=0
run "eigen table        R4 minb5      " B200_TIP_MODE=2 B200_WALK_MINB=5
run "eigen contraction  R4 minb4      " B200_TIP_MODE=0
cat $OUT
