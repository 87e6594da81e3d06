"""Caller re-enactment of the older front-end (test/bench harness, not the product).

Mirrors src/dr/evomodel/treelikelihood/BeagleTreeLikelihood.java -- the likelihood examples/Benchmarks/benchmark1.xml
and benchmark2.xml instantiate (<treeLikelihood>):

  * buffer accounting / instance creation ................. :186-190, :420-433
  * calculateLogLikelihood ................................ :863-1130  (rescale decision :884-912, tip-states models
                                                            :917-930 [setTipPartials / setTipStates per dirty tip], the
                                                            do-while underflow retry INSIDE the call :994-1113,
                                                            ascertainment correction from getSiteLogLikelihoods :1050-1057)
  * traverse (recursive post-order, flips while it walks) . :1202-1320
  * storeState / restoreState ............................. :816-851
  * SubstitutionModelDelegate (homogeneous case) .......... treelikelihood/SubstitutionModelDelegate.java:72-112,272-287,395-407
  * AscertainedSitePatterns.getAscertainmentCorrection .... src/dr/evolution/alignment/AscertainedSitePatterns.java:174-195
"""
from __future__ import annotations

import math
import sys
from typing import Callable, Optional, Sequence

import numpy as np

from .evomodel import GammaSiteRateModel, Patterns, SubstitutionModel, Tree
from .treedatalikelihood import (FLAG_EIGEN_COMPLEX, FLAG_PROCESSOR_GPU, NONE, OPERATION_TUPLE_SIZE, BufferIndexHelper,
                                 PartialsRescalingScheme)

RESCALE_FREQUENCY = 10000
RESCALE_TIMES = 1


class SubstitutionModelDelegate:
    """One substitution model on every branch, no matrix convolution (extraBufferCount = 0)."""

    def __init__(self, tree: Tree, substitutionModel: SubstitutionModel):
        self.substitutionModel = substitutionModel
        self.eigenBufferHelper = BufferIndexHelper(1, 0)
        self.matrixBufferHelper = BufferIndexHelper(tree.nodeCount, 0)

    def getEigenBufferCount(self) -> int:
        return self.eigenBufferHelper.getBufferCount()

    def getMatrixBufferCount(self) -> int:
        return self.matrixBufferHelper.getBufferCount() + 1          # + the reserve buffer (:109)

    def getMatrixIndex(self, branchIndex: int) -> int:
        return self.matrixBufferHelper.getOffsetIndex(branchIndex)

    def flipMatrixBuffer(self, branchIndex: int) -> None:
        self.matrixBufferHelper.flipOffset(branchIndex)

    def updateSubstitutionModels(self, beagle) -> None:
        self.eigenBufferHelper.flipOffset(0)
        ed = self.substitutionModel.getEigenDecomposition()
        beagle.setEigenDecomposition(self.eigenBufferHelper.getOffsetIndex(0), ed.Evec, ed.Ievc, ed.Eval)

    def updateTransitionMatrices(self, beagle, branchIndices, edgeLengths, updateCount: int) -> None:
        prob = np.array([self.matrixBufferHelper.getOffsetIndex(int(branchIndices[i])) for i in range(updateCount)],
                        dtype=np.int32)
        beagle.updateTransitionMatrices(self.eigenBufferHelper.getOffsetIndex(0), prob, None, None,
                                        np.ascontiguousarray(edgeLengths[:updateCount]), updateCount)

    def getRootStateFrequencies(self) -> np.ndarray:
        return self.substitutionModel.getFrequencies()

    def storeState(self) -> None:
        self.eigenBufferHelper.storeState()
        self.matrixBufferHelper.storeState()

    def restoreState(self) -> None:
        self.eigenBufferHelper.restoreState()
        self.matrixBufferHelper.restoreState()


class TipPartialsModel:
    """A TipStatesModel of Type.PARTIALS (tip-error models): getTipPartials(tip) -> [P][S]."""

    def __init__(self, partials: Sequence[np.ndarray]):
        self.partials = [np.ascontiguousarray(p, dtype=np.float64) for p in partials]

    def getTipPartials(self, index: int) -> np.ndarray:
        return self.partials[index]


class BeagleTreeLikelihood:
    def __init__(self, patternList: Patterns, tree: Tree, substitutionModel: SubstitutionModel,
                 siteRateModel: GammaSiteRateModel, beagleFactory: Callable, tipStatesModel: Optional[TipPartialsModel] = None,
                 rescalingScheme: str = PartialsRescalingScheme.DEFAULT, delayRescalingUntilUnderflow: bool = True,
                 resourceList=None, preferenceFlags: int = 0, requirementFlags: int = 0,
                 ascertainedExclude: Optional[Sequence[int]] = None, rescalingFrequency: int = RESCALE_FREQUENCY):
        self.patternList, self.treeModel = patternList, tree
        self.siteRateModel = siteRateModel
        self.patternCount, self.stateCount = patternList.patternCount, patternList.stateCount
        self.categoryCount = siteRateModel.getCategoryCount()
        self.nodeCount, self.tipCount = tree.nodeCount, tree.tipCount
        self.internalNodeCount = self.nodeCount - self.tipCount
        self.partialBufferHelper = BufferIndexHelper(self.nodeCount, self.tipCount)
        self.scaleBufferHelper = BufferIndexHelper(self.internalNodeCount + 1, 0)
        self.substitutionModelDelegate = SubstitutionModelDelegate(tree, substitutionModel)
        self.tipStatesModel = tipStatesModel
        self.ascertainedExclude = None if ascertainedExclude is None else np.asarray(ascertainedExclude, dtype=np.int64)

        self.rescalingScheme = rescalingScheme
        self.delayRescalingUntilUnderflow = delayRescalingUntilUnderflow
        if resourceList is not None and resourceList[0] > 0:
            preferenceFlags |= FLAG_PROCESSOR_GPU
        if self.rescalingScheme == PartialsRescalingScheme.DEFAULT:
            self.rescalingScheme = PartialsRescalingScheme.DYNAMIC
        if self.rescalingScheme == PartialsRescalingScheme.DELAYED:
            self.delayRescalingUntilUnderflow = True
            self.rescalingScheme = PartialsRescalingScheme.ALWAYS
        self.rescalingFrequency = rescalingFrequency
        if substitutionModel.canReturnComplexDiagonalization():
            requirementFlags |= FLAG_EIGEN_COMPLEX
        compact = 0 if tipStatesModel is not None else self.tipCount      # PARTIALS tip models take partials buffers (:286-294)
        self.beagle = beagleFactory(self.tipCount, self.partialBufferHelper.getBufferCount(), compact, self.stateCount,
                                    self.patternCount, self.substitutionModelDelegate.getEigenBufferCount(),
                                    self.substitutionModelDelegate.getMatrixBufferCount(), self.categoryCount,
                                    self.scaleBufferHelper.getBufferCount(), resourceList, preferenceFlags, requirementFlags)
        if tipStatesModel is None:
            for i in range(self.tipCount):
                self.beagle.setTipStates(i, np.ascontiguousarray(patternList.states[i], dtype=np.int32))
        self.patternWeights = np.ascontiguousarray(patternList.weights, dtype=np.float64)
        self.beagle.setPatternWeights(self.patternWeights)

        self.updateNode = np.ones(self.nodeCount, dtype=bool)
        self.updateSubstitutionModel = True
        self.updateSiteModel = True
        self.useScaleFactors = False
        self.recomputeScaleFactors = False
        self.everUnderflowed = False
        self.rescalingCount = 0
        self.rescalingCountInner = 0
        self.branchUpdateIndices = np.zeros(self.nodeCount, dtype=np.int32)
        self.branchLengths = np.zeros(self.nodeCount)
        self.scaleBufferIndices = [0] * self.internalNodeCount
        self.storedScaleBufferIndices = [0] * self.internalNodeCount
        self.operations = np.zeros(self.internalNodeCount * OPERATION_TUPLE_SIZE, dtype=np.int32)
        self.patternLogLikelihoods = np.zeros(self.patternCount)
        self.likelihoodKnown = False
        self.logLikelihood = 0.0

    # AbstractTreeLikelihood
    def updateAllNodes(self) -> None:
        self.updateNode[:] = True
        self.likelihoodKnown = False

    def updateNodeAndChildren(self, node: int) -> None:
        self.updateNode[node] = True
        for c in self.treeModel.child[node]:
            if c >= 0:
                self.updateNode[c] = True
        self.likelihoodKnown = False

    def makeDirty(self) -> None:
        self.updateSiteModel = True
        self.updateSubstitutionModel = True
        self.updateAllNodes()

    def getLogLikelihood(self) -> float:
        if not self.likelihoodKnown:
            self.logLikelihood = self.calculateLogLikelihood()
            self.likelihoodKnown = True
        return self.logLikelihood

    # ---- :863-1130
    def calculateLogLikelihood(self) -> float:
        S = PartialsRescalingScheme
        beagle = self.beagle
        self.recomputeScaleFactors = False
        if not self.delayRescalingUntilUnderflow or self.everUnderflowed:
            if self.rescalingScheme in (S.ALWAYS, S.DELAYED):
                self.useScaleFactors = True
                self.recomputeScaleFactors = True
            elif self.rescalingScheme == S.DYNAMIC:
                self.useScaleFactors = True
                if self.rescalingCount > self.rescalingFrequency:
                    self.rescalingCount = 0
                    self.rescalingCountInner = 0
                if self.rescalingCountInner < RESCALE_TIMES:
                    self.recomputeScaleFactors = True
                    self.updateNode[:] = True
                    self.rescalingCountInner += 1
                self.rescalingCount += 1

        if self.tipStatesModel is not None:
            for index in range(self.tipCount):
                if self.updateNode[index]:
                    beagle.setTipPartials(index, self.tipStatesModel.getTipPartials(index).reshape(-1))

        self.branchUpdateCount = 0
        self.operationCount = 0
        root = self.treeModel.root
        self._traverse(root, True)

        if self.updateSubstitutionModel:
            self.substitutionModelDelegate.updateSubstitutionModels(beagle)
        if self.updateSiteModel:
            rates = self.siteRateModel.getCategoryRates()
            if rates is None:
                return -math.inf
            beagle.setCategoryRates(np.ascontiguousarray(rates, dtype=np.float64))
        if self.branchUpdateCount > 0:
            self.substitutionModelDelegate.updateTransitionMatrices(beagle, self.branchUpdateIndices, self.branchLengths,
                                                                    self.branchUpdateCount)
        firstRescaleAttempt = True
        while True:
            beagle.updatePartials(self.operations, self.operationCount, NONE)
            rootIndex = self.partialBufferHelper.getOffsetIndex(root)
            cumulateScaleBufferIndex = NONE
            if self.useScaleFactors:
                if self.recomputeScaleFactors:
                    self.scaleBufferHelper.flipOffset(self.internalNodeCount)
                    cumulateScaleBufferIndex = self.scaleBufferHelper.getOffsetIndex(self.internalNodeCount)
                    beagle.resetScaleFactors(cumulateScaleBufferIndex)
                    beagle.accumulateScaleFactors(np.asarray(self.scaleBufferIndices, dtype=np.int32),
                                                  self.internalNodeCount, cumulateScaleBufferIndex)
                else:
                    cumulateScaleBufferIndex = self.scaleBufferHelper.getOffsetIndex(self.internalNodeCount)
            beagle.setCategoryWeights(0, np.ascontiguousarray(self.siteRateModel.getCategoryProportions(), dtype=np.float64))
            beagle.setStateFrequencies(0, np.ascontiguousarray(self.substitutionModelDelegate.getRootStateFrequencies(),
                                                               dtype=np.float64))
            sumLogLikelihoods = np.zeros(1)
            beagle.calculateRootLogLikelihoods(np.array([rootIndex], dtype=np.int32), np.zeros(1, dtype=np.int32),
                                               np.zeros(1, dtype=np.int32),
                                               np.array([cumulateScaleBufferIndex], dtype=np.int32), 1, sumLogLikelihoods)
            logL = float(sumLogLikelihoods[0])
            beagle.getSiteLogLikelihoods(self.patternLogLikelihoods)
            if self.ascertainedExclude is not None:
                logL = self._ascertainmentCorrected()
            if math.isnan(logL) or math.isinf(logL):
                self.everUnderflowed = True
                logL = -math.inf
                if firstRescaleAttempt and (self.delayRescalingUntilUnderflow or self.rescalingScheme == S.DELAYED):
                    self.useScaleFactors = True
                    self.recomputeScaleFactors = True
                    self.branchUpdateCount = 0
                    self.updateNode[:] = True
                    self.operationCount = 0
                    self._traverse(root, False)      # same destinations: overwrite the underflowed attempt
                    firstRescaleAttempt = False
                    continue
            break
        self.updateNode[:] = False
        self.updateSubstitutionModel = False
        self.updateSiteModel = False
        return logL

    def _ascertainmentCorrected(self) -> float:
        """AscertainedSitePatterns.getAscertainmentCorrection with only excluded patterns (the common use: the invariant
        or unobservable patterns are appended to the pattern list with weight 0 and conditioned out)."""
        excludeProb = float(np.exp(self.patternLogLikelihoods[self.ascertainedExclude]).sum())
        correction = math.log(1.0 - excludeProb)
        logL = 0.0
        for i in range(self.patternCount):
            logL += (self.patternLogLikelihoods[i] - correction) * self.patternWeights[i]
        return logL

    # ---- :1202-1320
    def _traverse(self, node: int, flip: bool) -> bool:
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * self.nodeCount + 100))
        tree = self.treeModel
        update = False
        parent = int(tree.parent[node])
        if parent >= 0 and self.updateNode[node]:
            branchLength = tree.branchLength(node)
            if flip:
                self.substitutionModelDelegate.flipMatrixBuffer(node)
            self.branchUpdateIndices[self.branchUpdateCount] = node
            self.branchLengths[self.branchUpdateCount] = branchLength
            self.branchUpdateCount += 1
            update = True
        if not tree.isExternal(node):
            child1, child2 = int(tree.child[node][0]), int(tree.child[node][1])
            update1 = self._traverse(child1, flip)
            update2 = self._traverse(child2, flip)
            if update1 or update2:
                x = self.operationCount * OPERATION_TUPLE_SIZE
                if flip:
                    self.partialBufferHelper.flipOffset(node)
                ops = self.operations
                ops[x] = self.partialBufferHelper.getOffsetIndex(node)
                if self.useScaleFactors:
                    n = node - self.tipCount
                    if self.recomputeScaleFactors:
                        self.scaleBufferHelper.flipOffset(n)
                        self.scaleBufferIndices[n] = self.scaleBufferHelper.getOffsetIndex(n)
                        ops[x + 1], ops[x + 2] = self.scaleBufferIndices[n], NONE
                    else:
                        ops[x + 1], ops[x + 2] = NONE, self.scaleBufferIndices[n]
                else:
                    ops[x + 1], ops[x + 2] = NONE, NONE
                ops[x + 3] = self.partialBufferHelper.getOffsetIndex(child1)
                ops[x + 4] = self.substitutionModelDelegate.getMatrixIndex(child1)
                ops[x + 5] = self.partialBufferHelper.getOffsetIndex(child2)
                ops[x + 6] = self.substitutionModelDelegate.getMatrixIndex(child2)
                self.operationCount += 1
                update = True
        return update

    def getPartials(self, number: int) -> np.ndarray:
        out = np.zeros(self.patternCount * self.stateCount * self.categoryCount)
        self.beagle.getPartials(self.partialBufferHelper.getOffsetIndex(number), NONE, out)
        return out

    def storeState(self) -> None:
        self.partialBufferHelper.storeState()
        self.substitutionModelDelegate.storeState()
        if self.useScaleFactors:
            self.scaleBufferHelper.storeState()
            self.storedScaleBufferIndices = list(self.scaleBufferIndices)
        self._storedLogL, self._storedKnown = self.logLikelihood, self.likelihoodKnown

    def restoreState(self) -> None:
        self.updateSiteModel = True
        self.partialBufferHelper.restoreState()
        self.substitutionModelDelegate.restoreState()
        if self.useScaleFactors:
            self.scaleBufferHelper.restoreState()
            self.scaleBufferIndices, self.storedScaleBufferIndices = self.storedScaleBufferIndices, self.scaleBufferIndices
        self.logLikelihood, self.likelihoodKnown = self._storedLogL, self._storedKnown

    def finalize(self) -> None:
        self.beagle.finalize()
