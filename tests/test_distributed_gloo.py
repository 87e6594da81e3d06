"""N > 1 host logic on CPU: two gloo ranks, each owning one contiguous pattern block
(Patterns.java:142-169), one all-reduce of one double -> equals the unsharded value.  The engine on
each rank is the C restatement (checker use: there is no GPU here)."""
import os
import subprocess
import sys
import textwrap

import helpers as H

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import torch.distributed as dist
    import helpers as H
    from beast_mcmc_b200 import build
    from harness import sharding, treedatalikelihood as tdl
    from oracle import cpu
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, pats, model, site = H.synthetic_case(30, 501, 4, seed=8)
    sh = sharding.ShardedTreeDataLikelihood(tree, pats, model, site, cpu.factory(threads=2), rank, world,
                                            rescalingScheme=tdl.PartialsRescalingScheme.NONE)
    joint = sh.getLogLikelihood()
    whole = tdl.TreeDataLikelihood(tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, cpu.factory(threads=2),
                                   rescalingScheme=tdl.PartialsRescalingScheme.NONE), tree).getLogLikelihood()
    sizes = [pats.subSet(k, world).patternCount for k in range(world)]
    assert sum(sizes) == pats.patternCount and sizes[0] - sizes[-1] in (0, 1), sizes
    assert abs(joint - whole) <= 1e-11 * abs(whole), (joint, whole)
    print(f"rank {{rank}} shard {{sh.shard.patternCount}} joint {{joint:.8f}} whole {{whole:.8f}} OK")
    dist.destroy_process_group()
""")


def test_two_rank_pattern_sharding(tmp_path):
    from beast_mcmc_b200 import build
    build.build_oracle()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=H.ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o, o


def test_reference_arm_under_torchrun_prints_one_line():
    """The driver launches `bench.py --impl reference --gpus N` the same way as the GPU arm (torchrun, N ranks): rank 0 alone
    times the CPU restatement and prints the JSON line, the other ranks exit 0 without work.  No GPU involved."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--workload", "hky_1441x593", "--steps", "2", "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["metric"] == "tree log-likelihood evals/sec" and d["config"]["workload"] == "hky_1441x593"
