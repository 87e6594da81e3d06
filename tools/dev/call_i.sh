mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_preorder_oracle.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02i_gpu_tests.txt; cat gpurun_out/r02i_gpu_tests.txt
bash tools/bench_all.sh gpurun_out/r02i_bench_all.txt
for mb in 3 4 5; do echo -n "walk4p minb$mb | "; B200_WALK_MINB=$mb python tools/bench_line.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras; done > gpurun_out/r02i_sweep.txt 2>&1; cat gpurun_out/r02i_sweep.txt
ncu --set full --clock-control none --import-source on -k regex:k_walk4p -s 30 -c 3 -o gpurun_out/r02i_walk4p_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk_mma -s 8 -c 1 -o gpurun_out/r02i_walk_mma_codon_full python bench.py --workload codon_mg94_500x5k --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
