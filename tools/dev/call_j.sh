mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02j_gpu_tests.txt; cat gpurun_out/r02j_gpu_tests.txt
bash tools/bench_all.sh gpurun_out/r02j_bench_all.txt
( echo -n "cfg2 lookahead=2 | "; B200_LOOKAHEAD=2 python tools/bench_line.py --steps 200 --warmup 10 --no-cpu-baseline --no-extras
  echo -n "makona R=4 | "; B200_WALK_R=4 python tools/bench_line.py --workload makona_like_1610x6k --steps 200 --warmup 10 --no-cpu-baseline --no-extras
  echo -n "makona R=4 lookahead=2 | "; B200_WALK_R=4 B200_LOOKAHEAD=2 python tools/bench_line.py --workload makona_like_1610x6k --steps 200 --warmup 10 --no-cpu-baseline --no-extras
  echo -n "makona lookahead=2 | "; B200_LOOKAHEAD=2 python tools/bench_line.py --workload makona_like_1610x6k --steps 200 --warmup 10 --no-cpu-baseline --no-extras
  echo -n "cfg2 nofuse | "; B200_FUSE=0 python tools/bench_line.py --steps 200 --warmup 10 --no-cpu-baseline
) > gpurun_out/r02j_sweep.txt 2>&1; cat gpurun_out/r02j_sweep.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02j_bench_cfg2_steps20.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_cfg2_steps20.json').read())
print('steps20', d['value'], d['e2e']['value'], d['e2e'].get('c_abi_replay'), d['incremental'].get('c_abi_replay'))
PY
