mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02n_gpu_tests.txt; cat gpurun_out/r02n_gpu_tests.txt
for w in gtr_g4_1000x10k gtr_g4_1000x10k_rescaled makona_like_1610x6k hky_1441x593 benchmark1_xml benchmark2_xml; do
  steps=300; case $w in hky*|benchmark*) steps=1000;; esac
  python tools/bench_line.py --workload $w --steps $steps --warmup 10 --no-cpu-baseline --no-extras
done > gpurun_out/r02n_bench.txt 2>&1; cat gpurun_out/r02n_bench.txt
