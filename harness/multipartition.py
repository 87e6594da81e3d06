"""Caller re-enactment of the multi-partition route (test/bench harness, not the product).

Mirrors src/dr/evomodel/treedatalikelihood/MultiPartitionDataLikelihoodDelegate.java of the reference:

  * constructor: buffer accounting and instance creation ............ :170-473, set-up of tips / weights / partition map :520-559
  * calculateLikelihood ............................................ :744-1207
      - per-partition rescale state machine ........................ :746-797
      - eigen systems and indexed category rates ................... :799-846
      - updateTransitionMatricesWithMultipleModels ................. :848-895
      - per-partition flipping of the (shared-index) partials ...... :897-911
      - 9-int operation tuples {dest, scaleWrite, scaleRead, c1, m1, c2, m2, partition, cumulative = NONE} .. :914-995
      - resetScaleFactorsByPartition / accumulateScaleFactorsByPartition ........................ :1005-1022
      - calculateRootLogLikelihoodsByPartition over the updated partitions ...................... :1046-1083
      - underflow handling per partition ............................ :1110-1190
  * storeState / restoreState ....................................... :1253-1305

All partitions address the SAME partials / scale buffer indices (every BufferIndexHelper is buffer set 0, :253-262): a
buffer spans the concatenated patterns of all partitions and each partition keeps its own flip state for its window.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import numpy as np

from .evomodel import GammaSiteRateModel, Patterns, SubstitutionModel, Tree
from .treedatalikelihood import (FLAG_EIGEN_COMPLEX, FLAG_FRAMEWORK_CPU, FLAG_PROCESSOR_GPU, NONE, BufferIndexHelper,
                                 HomogenousSubstitutionModelDelegate, LikelihoodRescalingException,
                                 LikelihoodUnderflowException, PartialsRescalingScheme)

PARTITION_OPERATION_TUPLE_SIZE = 9
RESCALE_FREQUENCY = 100          # MPDLD:149
RESCALE_TIMES = 1


class DelegateTypeException(Exception):
    """MPDLD:468-470 -- the instance reports FRAMEWORK_CPU: BEAST falls back to one delegate per partition."""


class MultiPartitionDataLikelihoodDelegate:
    def __init__(self, tree: Tree, patternLists: Sequence[Patterns], substitutionModels: Sequence[SubstitutionModel],
                 siteRateModels: Sequence[GammaSiteRateModel], beagleFactory: Callable, useAmbiguities: bool = False,
                 rescalingScheme: str = PartialsRescalingScheme.DEFAULT, delayRescalingUntilUnderflow: bool = True,
                 resourceList=None, preferenceFlags: int = 0, requirementFlags: int = 0,
                 rescalingFrequency: int = RESCALE_FREQUENCY, stateSetFn=None):
        self.patternLists = list(patternLists)
        self.stateCount = self.patternLists[0].stateCount
        self.partitionCount = len(self.patternLists)
        self.patternCounts = [p.patternCount for p in self.patternLists]
        self.totalPatternCount = sum(self.patternCounts)
        P = self.partitionCount
        self.useScaleFactors = [False] * P
        self.recomputeScaleFactors = [False] * P
        self.everUnderflowed = [False] * P
        self.flip = [True] * P
        self.updatePartition = [False] * P
        self.partitionWasUpdated = [False] * P
        self.updateAllPartitions = True
        self.cachedLogLikelihoodsByPartition = [0.0] * P
        self.storedCachedLogLikelihoodsByPartition = [0.0] * P
        assert len(substitutionModels) in (1, P) and len(siteRateModels) in (1, P)
        self.branchModels = list(substitutionModels)
        self.siteRateModels = list(siteRateModels)
        self.categoryCount = self.siteRateModels[0].getCategoryCount()
        self.nodeCount, self.tipCount = tree.nodeCount, tree.tipCount
        self.internalNodeCount = self.nodeCount - self.tipCount
        self.branchUpdateIndices = np.zeros(self.nodeCount, dtype=np.int32)
        self.branchLengths = np.zeros(self.nodeCount)
        self.scaleBufferIndices = [[0] * self.internalNodeCount for _ in range(P)]
        self.storedScaleBufferIndices = [[0] * self.internalNodeCount for _ in range(P)]
        self.operations = np.zeros(self.internalNodeCount * PARTITION_OPERATION_TUPLE_SIZE * P, dtype=np.int32)
        self.rescalingCount = [0] * P
        self.rescalingCountInner = [0] * P
        self.firstRescaleAttempt = True
        self.initialEvaluation = True

        compactPartialsCount = 0 if useAmbiguities else self.tipCount
        self.partialBufferHelper = [BufferIndexHelper(self.nodeCount, self.tipCount) for _ in range(P)]
        self.scaleBufferHelper = [BufferIndexHelper(self.internalNodeCount + 1, 0) for _ in range(P)]
        self.categoryRateBufferHelper = [BufferIndexHelper(1, 0, i) for i in range(P)]
        self.evolutionaryProcessDelegates = [HomogenousSubstitutionModelDelegate(tree, m, k)
                                             for k, m in enumerate(self.branchModels)]
        eigenBufferCount = sum(d.getEigenBufferCount() for d in self.evolutionaryProcessDelegates)
        matrixBufferCount = sum(d.getMatrixBufferCount() for d in self.evolutionaryProcessDelegates)

        self.rescalingScheme = rescalingScheme
        self.delayRescalingUntilUnderflow = delayRescalingUntilUnderflow
        if resourceList is not None and resourceList[0] > 0:
            preferenceFlags |= FLAG_PROCESSOR_GPU
        if self.rescalingScheme == PartialsRescalingScheme.DEFAULT:
            self.rescalingScheme = PartialsRescalingScheme.DYNAMIC
        if self.rescalingScheme == PartialsRescalingScheme.DELAYED:
            self.delayRescalingUntilUnderflow = True
            self.rescalingScheme = PartialsRescalingScheme.ALWAYS
        if self.rescalingScheme == PartialsRescalingScheme.AUTO:      # not supported for multi-partition instances (:330-333,533-537)
            self.rescalingScheme = PartialsRescalingScheme.DYNAMIC
        self.rescalingFrequency = rescalingFrequency
        if self.evolutionaryProcessDelegates[0].canReturnComplexDiagonalization():
            requirementFlags |= FLAG_EIGEN_COMPLEX

        self.beagle = beagleFactory(self.tipCount, self.partialBufferHelper[0].getBufferCount(), compactPartialsCount,
                                    self.stateCount, self.totalPatternCount, eigenBufferCount, matrixBufferCount,
                                    self.categoryCount, self.scaleBufferHelper[0].getBufferCount(), resourceList,
                                    preferenceFlags, requirementFlags)
        details = self.beagle.getDetails()
        if details is not None and (details.flags & FLAG_FRAMEWORK_CPU):
            raise DelegateTypeException()

        self.patternPartitions = np.concatenate([np.full(n, j, dtype=np.int32) for j, n in enumerate(self.patternCounts)])
        self.patternWeights = np.concatenate([np.asarray(p.weights, dtype=np.float64) for p in self.patternLists])
        for i in range(self.tipCount):
            if useAmbiguities:
                rows = [np.array([stateSetFn(int(s)) for s in p.states[i]], dtype=np.float64) for p in self.patternLists]
                part = np.concatenate(rows, axis=0)
                self.beagle.setPartials(i, np.ascontiguousarray(np.broadcast_to(
                    part, (self.categoryCount,) + part.shape)).reshape(-1))
            else:
                self.beagle.setTipStates(i, np.ascontiguousarray(
                    np.concatenate([p.states[i] for p in self.patternLists]), dtype=np.int32))
        self.beagle.setPatternWeights(self.patternWeights)
        self.beagle.setPatternPartitions(self.partitionCount, self.patternPartitions)
        self.updateSubstitutionModels = [True] * len(self.branchModels)
        self.updateSiteRateModels = [True] * len(self.siteRateModels)

    def getOptimalTraversalType(self) -> str:
        return "REVERSE_LEVEL_ORDER"                                  # MPDLD:480-483

    def makeDirty(self) -> None:
        self.updateSiteRateModels = [True] * len(self.siteRateModels)
        self.updateSubstitutionModels = [True] * len(self.branchModels)

    def calculateLikelihood(self, branchOperations: List[Tuple[int, float]],
                            nodeOperations: List[Tuple[int, int, int]], rootNodeNumber: int) -> float:
        S = PartialsRescalingScheme
        P = self.partitionCount
        throwRescaling = False
        if not self.initialEvaluation:
            for i in range(P):
                if not self.delayRescalingUntilUnderflow or self.everUnderflowed[i]:
                    if self.rescalingScheme in (S.ALWAYS, S.DELAYED):
                        self.useScaleFactors[i] = True
                        self.recomputeScaleFactors[i] = True
                    elif self.rescalingScheme == S.DYNAMIC:
                        self.useScaleFactors[i] = True
                        if self.rescalingCount[i] > self.rescalingFrequency:
                            self.rescalingCount[i] = 0
                            self.rescalingCountInner[i] = 0
                        if self.rescalingCountInner[i] < RESCALE_TIMES:
                            self.recomputeScaleFactors[i] = True
                            self.updatePartition[i] = True
                            self.rescalingCountInner[i] += 1
                            throwRescaling = True
            if throwRescaling:
                raise LikelihoodRescalingException()

        beagle = self.beagle
        for k, epd in enumerate(self.evolutionaryProcessDelegates):
            if self.updateSubstitutionModels[k]:
                epd.updateSubstitutionModels(beagle, self.flip[k])
                self.updatePartition[k] = True
                self.updateAllPartitions = False
        for k, site in enumerate(self.siteRateModels):
            if self.updateSiteRateModels[k]:
                rates = site.getCategoryRates()
                if rates is None:
                    self.updateSubstitutionModels = [False] * len(self.branchModels)
                    self.updateSiteRateModels = [False] * len(self.siteRateModels)
                    return -math.inf
                if self.flip[k]:
                    self.categoryRateBufferHelper[k].flipOffset(0)
                beagle.setCategoryRatesWithIndex(self.categoryRateBufferHelper[k].getOffsetIndex(0),
                                                 np.ascontiguousarray(rates, dtype=np.float64))
                self.updatePartition[k] = True
                self.updateAllPartitions = False

        branchUpdateCount = 0
        for branchNumber, branchLength in branchOperations:
            self.branchUpdateIndices[branchUpdateCount] = branchNumber
            self.branchLengths[branchUpdateCount] = branchLength
            branchUpdateCount += 1
        if branchUpdateCount > 0:
            eig, rate, prob, lens = [], [], [], []
            for partition, epd in enumerate(self.evolutionaryProcessDelegates):
                if self.updatePartition[partition] or self.updateAllPartitions:
                    if self.flip[partition]:
                        for i in range(branchUpdateCount):      # flipTransitionMatrices
                            epd.matrixBufferHelper.flipOffset(int(self.branchUpdateIndices[i]))
                    for i in range(branchUpdateCount):
                        eig.append(epd.getEigenIndex(0))
                        rate.append(self.categoryRateBufferHelper[partition].getOffsetIndex(0))
                        prob.append(epd.getMatrixIndex(int(self.branchUpdateIndices[i])))
                        lens.append(self.branchLengths[i])
            beagle.updateTransitionMatricesWithMultipleModels(
                np.asarray(eig, dtype=np.int32), np.asarray(rate, dtype=np.int32), np.asarray(prob, dtype=np.int32),
                None, None, np.asarray(lens, dtype=np.float64), len(eig))

        for i in range(P):
            if (self.updatePartition[i] or self.updateAllPartitions) and self.flip[i]:
                for nodeNum, _, _ in nodeOperations:
                    self.partialBufferHelper[i].flipOffset(nodeNum)

        ops = self.operations
        k = 0
        operationCount = 0
        mapPartition = P // len(self.evolutionaryProcessDelegates)
        for nodeNum, leftChild, rightChild in nodeOperations:
            writeScale, readScale = [NONE] * P, [NONE] * P
            for i in range(P):
                if (self.updatePartition[i] or self.updateAllPartitions) and self.useScaleFactors[i]:
                    n = nodeNum - self.tipCount
                    if self.recomputeScaleFactors[i]:
                        self.scaleBufferHelper[i].flipOffset(n)
                        self.scaleBufferIndices[i][n] = self.scaleBufferHelper[i].getOffsetIndex(n)
                        writeScale[i] = self.scaleBufferIndices[i][n]
                    else:
                        readScale[i] = self.scaleBufferIndices[i][n]
            for i in range(P):
                if self.updatePartition[i] or self.updateAllPartitions:
                    epd = self.evolutionaryProcessDelegates[i // mapPartition]
                    h = self.partialBufferHelper[i]
                    ops[k:k + 9] = (h.getOffsetIndex(nodeNum), writeScale[i], readScale[i],
                                    h.getOffsetIndex(leftChild), epd.getMatrixIndex(leftChild),
                                    h.getOffsetIndex(rightChild), epd.getMatrixIndex(rightChild), i, NONE)
                    k += PARTITION_OPERATION_TUPLE_SIZE
                    operationCount += 1
        beagle.updatePartialsByPartition(ops, operationCount)

        cumulativeScaleIndices = [NONE] * P
        for i in range(P):
            if self.useScaleFactors[i]:
                if self.recomputeScaleFactors[i] and (self.updatePartition[i] or self.updateAllPartitions):
                    self.scaleBufferHelper[i].flipOffset(self.internalNodeCount)
                    cumulativeScaleIndices[i] = self.scaleBufferHelper[i].getOffsetIndex(self.internalNodeCount)
                    beagle.resetScaleFactorsByPartition(cumulativeScaleIndices[i], i)
                    beagle.accumulateScaleFactorsByPartition(np.asarray(self.scaleBufferIndices[i], dtype=np.int32),
                                                             self.internalNodeCount, cumulativeScaleIndices[i], i)
                else:
                    cumulativeScaleIndices[i] = self.scaleBufferHelper[i].getOffsetIndex(self.internalNodeCount)

        for i, site in enumerate(self.siteRateModels):
            beagle.setCategoryWeights(i, np.ascontiguousarray(site.getCategoryProportions(), dtype=np.float64))
            beagle.setStateFrequencies(i, np.ascontiguousarray(
                self.evolutionaryProcessDelegates[i].getRootStateFrequencies(), dtype=np.float64))

        partitionIndices, rootIndices, wIdx, fIdx, cumIdx = [], [], [], [], []
        for i in range(P):
            if self.updatePartition[i] or self.updateAllPartitions:
                partitionIndices.append(i)
                rootIndices.append(self.partialBufferHelper[i].getOffsetIndex(rootNodeNumber))
                wIdx.append(i % len(self.siteRateModels))
                fIdx.append(i % len(self.siteRateModels))
                cumIdx.append(cumulativeScaleIndices[i])
        updatedPartitionCount = len(partitionIndices)
        sumByPartition = np.zeros(P)
        sumLogLikelihoods = np.zeros(1)
        as_i = lambda v: np.asarray(v, dtype=np.int32)
        beagle.calculateRootLogLikelihoodsByPartition(as_i(rootIndices), as_i(wIdx), as_i(fIdx), as_i(cumIdx),
                                                      as_i(partitionIndices), updatedPartitionCount, 1,
                                                      sumByPartition, sumLogLikelihoods)
        for i in range(updatedPartitionCount):
            self.cachedLogLikelihoodsByPartition[partitionIndices[i]] = float(sumByPartition[i])
            self.updatePartition[partitionIndices[i]] = False
            self.recomputeScaleFactors[partitionIndices[i]] = False
            self.partitionWasUpdated[partitionIndices[i]] = True
        tmpLogL = float(sumLogLikelihoods[0])
        self.updateSubstitutionModels = [False] * len(self.branchModels)
        self.updateSiteRateModels = [False] * len(self.siteRateModels)
        self.updateAllPartitions = True

        bad = lambda v: math.isnan(v) or math.isinf(v)
        if bad(tmpLogL):
            for i in range(updatedPartitionCount):
                if bad(float(sumByPartition[i])):
                    self.everUnderflowed[partitionIndices[i]] = True
            if self.firstRescaleAttempt:
                for i in range(updatedPartitionCount):
                    if (self.delayRescalingUntilUnderflow or self.rescalingScheme == S.DELAYED) and \
                            bad(float(sumByPartition[i])):
                        q = partitionIndices[i]
                        self.useScaleFactors[q] = True
                        self.recomputeScaleFactors[q] = True
                        self.updatePartition[q] = True
                        self.flip[q] = False         # overwrite the underflowed buffers on the retry
                        self.updateAllPartitions = False
                self.firstRescaleAttempt = False
                raise LikelihoodUnderflowException()
            return -math.inf
        for i in range(updatedPartitionCount):
            q = partitionIndices[i]
            if self.partitionWasUpdated[q]:
                if (not self.delayRescalingUntilUnderflow or self.everUnderflowed[q]) and \
                        self.rescalingScheme == S.DYNAMIC and not self.initialEvaluation:
                    self.rescalingCount[q] += 1
                self.partitionWasUpdated[q] = False
            self.recomputeScaleFactors[q] = False
            self.flip[q] = True
        self.firstRescaleAttempt = True
        self.initialEvaluation = False
        return float(sum(self.cachedLogLikelihoodsByPartition))

    def getSiteLogLikelihoods(self) -> np.ndarray:
        out = np.zeros(self.totalPatternCount)
        self.beagle.getSiteLogLikelihoods(out)
        return out

    def getPartials(self, partition: int, number: int) -> np.ndarray:
        """[C][totalPatterns][S] of the buffer partition `partition` currently maps node `number` to."""
        out = np.zeros(self.totalPatternCount * self.stateCount * self.categoryCount)
        self.beagle.getPartials(self.partialBufferHelper[partition].getOffsetIndex(number), NONE, out)
        return out

    def storeState(self) -> None:
        for i in range(self.partitionCount):
            self.partialBufferHelper[i].storeState()
            self.categoryRateBufferHelper[i].storeState()
        for epd in self.evolutionaryProcessDelegates:
            epd.storeState()
        for i in range(self.partitionCount):
            if self.useScaleFactors[i]:
                self.scaleBufferHelper[i].storeState()
                self.storedScaleBufferIndices[i] = list(self.scaleBufferIndices[i])
            self.flip[i] = True
        self.storedCachedLogLikelihoodsByPartition = list(self.cachedLogLikelihoodsByPartition)

    def restoreState(self) -> None:
        for i in range(self.partitionCount):
            self.partialBufferHelper[i].restoreState()
            self.categoryRateBufferHelper[i].restoreState()
        for epd in self.evolutionaryProcessDelegates:
            epd.restoreState()
        for i in range(self.partitionCount):
            if self.useScaleFactors[i]:
                self.scaleBufferHelper[i].restoreState()
                self.scaleBufferIndices[i], self.storedScaleBufferIndices[i] = \
                    self.storedScaleBufferIndices[i], self.scaleBufferIndices[i]
        self.cachedLogLikelihoodsByPartition, self.storedCachedLogLikelihoodsByPartition = \
            self.storedCachedLogLikelihoodsByPartition, self.cachedLogLikelihoodsByPartition

    def finalize(self) -> None:
        self.beagle.finalize()
