"""Test/bench harness (NOT part of the product): Python re-enactments of the reference's Java callers
of the BEAGLE boundary, used to drive the CUDA engine, the oracle and the CPU port with the identical
call sequences BEAST issues.  Nothing under beast-mcmc_b200/ imports this package.

  evomodel.py              producers of the inputs BEAST hands to BEAGLE (eigen systems, rates, trees, patterns)
  treedatalikelihood.py    BufferIndexHelper, HomogenousSubstitutionModelDelegate, BeagleDataLikelihoodDelegate,
                           TreeDataLikelihood / LikelihoodTreeTraversal, pre-order gradient delegates
  multipartition.py        MultiPartitionDataLikelihoodDelegate (the *ByPartition route)
  beagletreelikelihood.py  the older BeagleTreeLikelihood front-end (benchmark1/2.xml use it)
  sharding.py              -beagle_instances pattern split across ranks (gloo/NCCL sum of one double)
"""
