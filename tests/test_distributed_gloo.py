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
    from beast_mcmc_b200 import sharding, treedatalikelihood as tdl, build
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
