mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02k_gpu_tests.txt; cat gpurun_out/r02k_gpu_tests.txt
bash tools/bench_all.sh gpurun_out/r02k_bench_all.txt --no-extras
( echo -n "codon thin MT=2 (B200_THIN_R1=0) | "; B200_THIN_R1=0 python tools/bench_line.py --workload codon_mg94_500x5k --steps 100 --warmup 10 --no-cpu-baseline --no-extras
  echo -n "hky R rule old (B200_WALK_R=1) | "; B200_WALK_R=1 python tools/bench_line.py --workload hky_1441x593 --steps 1000 --warmup 10 --no-cpu-baseline --no-extras
) > gpurun_out/r02k_sweep.txt 2>&1; cat gpurun_out/r02k_sweep.txt
