mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
run() { local label=$1; shift; echo -n "$label | "; env "$@" python tools/bench_line.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras; }
( run "default                " B200_NOP=1
  run "thin phases: table      " B200_THIN_TIP_MODE=2
  run "thin: table + lookahead " B200_THIN_TIP_MODE=2 B200_LOOKAHEAD=2
  run "thin: contraction       " B200_THIN_TIP_MODE=0 B200_LOOKAHEAD=2
  run "phase T=40              " B200_PHASE_T=40
  run "phase T=90              " B200_PHASE_T=90
  run "phase T=125             " B200_PHASE_T=125
  run "phase T=200             " B200_PHASE_T=200
  run "thin R1 off             " B200_THIN_R1=0
  run "phase small=48          " B200_PHASE_SMALL=48
  run "phase small=8           " B200_PHASE_SMALL=8
) > gpurun_out/r02l_sweep.txt 2>&1; cat gpurun_out/r02l_sweep.txt
