mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02p_bench_n4.json 2> gpurun_out/r02p_bench_n4.err
python - <<'PY'
import json
f='gpurun_out/r02p_bench_n4.json'
d=json.loads([l for l in open(f) if l.startswith('{')][-1]); s=d.get('strong_scaling') or {}
print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), d.get('block_ms_p10'), d.get('block_ms_p50'), d.get('block_ms_p90'),
      {k:(round(v['joint_evals_per_s'],1), round(v.get('e2e_joint_evals_per_s',0),1)) for k,v in s.items() if isinstance(v,dict)})
PY
