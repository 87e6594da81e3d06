"""Pattern sharding across ranks: the engine-side analogue of BEAST's `-beagle_instances N`
(TreeDataLikelihoodParser.java:205-229 builds one delegate per contiguous pattern block,
Patterns.java:142-169; CompoundLikelihood.java:198-241 sums them).  One process per GPU: every rank
evaluates its own block and the per-shard log-likelihoods are summed with a single all-reduce of
one double (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable

from .evomodel import GammaSiteRateModel, Patterns, SubstitutionModel, Tree
from .treedatalikelihood import BeagleDataLikelihoodDelegate, TreeDataLikelihood


class ShardedTreeDataLikelihood:
    def __init__(self, tree: Tree, patterns: Patterns, model: SubstitutionModel, site: GammaSiteRateModel,
                 beagleFactory: Callable, rank: int, world: int, resourceList=None, **delegateArgs):
        self.rank, self.world = rank, world
        self.shard = patterns.subSet(rank, world) if world > 1 else patterns
        self.delegate = BeagleDataLikelihoodDelegate(tree, self.shard, model, site, beagleFactory,
                                                     resourceList=resourceList, **delegateArgs)
        self.likelihood = TreeDataLikelihood(self.delegate, tree)

    def localLogLikelihood(self) -> float:
        return self.likelihood.getLogLikelihood()

    def getLogLikelihood(self, device=None) -> float:
        """Joint log-likelihood: local shard value, then ONE sum all-reduce of a single double."""
        local = self.localLogLikelihood()
        if self.world == 1:
            return local
        import torch
        import torch.distributed as dist
        t = torch.tensor([local], dtype=torch.float64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
