mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity2.py -m gpu -q -x -k "sharded or reduce_group or two_instances or distinct" 2>&1 | tail -3 > gpurun_out/r02o_gpu_tests_2gpu.txt; cat gpurun_out/r02o_gpu_tests_2gpu.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02o_bench_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02o_bench_n2.json 2> gpurun_out/r02o_bench_n2.err
STEPS=200 python tools/bench_sharded.py > gpurun_out/r02o_sharded_makona.json 2> gpurun_out/r02o_sharded.err
python - <<'PY'
import json
for f in ('gpurun_out/r02o_bench_n1.json','gpurun_out/r02o_bench_n2.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); s=d.get('strong_scaling') or {}
        print(f, d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d.get('block_ms_p10'), d.get('block_ms_p50'), d.get('block_ms_p90'),
              {k:(round(v['joint_evals_per_s'],1), round(v.get('e2e_joint_evals_per_s',0),1)) for k,v in s.items() if isinstance(v,dict)})
    except Exception as e: print(f, 'ERR', e)
print(open('gpurun_out/r02o_sharded_makona.json').read()[:1500])
PY
