"""Caller re-enactment: the host-side mirror of the reference's L2/L3 likelihood front-end.

BEAST itself cannot run here (no JVM), so the drop-in claim is carried by re-issuing the exact
BEAGLE call sequence the reference issues, with the same buffer-index flipping, store/restore
and rescaling state machine, against any object that implements the ``beagle.Beagle`` method
set (the ctypes binding in ``beagle.py`` over our C ABI, or the numpy oracle in tests).

Mirrored reference code (same names, same argument meaning):

  * BufferIndexHelper ........................ src/dr/evomodel/treedatalikelihood/BufferIndexHelper.java:39-115
  * HomogenousSubstitutionModelDelegate ...... .../HomogenousSubstitutionModelDelegate.java:80-287
  * BeagleDataLikelihoodDelegate ............. .../BeagleDataLikelihoodDelegate.java:114-570 (ctor),
                                               :734-1018 (calculateLikelihood), :1085-1136 (store/restore)
  * LikelihoodTreeTraversal / TreeTraversal .. .../LikelihoodTreeTraversal.java:49-204, TreeTraversal.java:66-124
  * TreeDataLikelihood ....................... .../TreeDataLikelihood.java:145-181,330-368
"""
from __future__ import annotations

import math
from typing import Callable, List, Tuple

import numpy as np

from .evomodel import GammaSiteRateModel, Patterns, SubstitutionModel, Tree

NONE = -1
OPERATION_TUPLE_SIZE = 7

# beagle.BeagleFlag masks (decoded from lib/beagle.jar; see include/libhmsbeagle_b200.h)
FLAG_PRECISION_DOUBLE = 1 << 1
FLAG_EIGEN_COMPLEX = 1 << 5
FLAG_SCALING_AUTO = 1 << 7
FLAG_PROCESSOR_GPU = 1 << 16
FLAG_FRAMEWORK_CPU = 1 << 27
FLAG_THREADING_CPP = 1 << 30


class LikelihoodException(Exception):
    pass


class LikelihoodUnderflowException(LikelihoodException):
    pass


class LikelihoodRescalingException(LikelihoodException):
    pass


class BufferIndexHelper:
    """Double buffering of per-node buffers: index i lives at i or i + doubleBufferCount."""

    def __init__(self, maxIndexValue: int, minIndexValue: int, bufferSetNumber: int = 0):
        self.minIndexValue = minIndexValue
        self.doubleBufferCount = maxIndexValue - minIndexValue
        self.indexOffsets = [0] * self.doubleBufferCount
        self.storedIndexOffsets = [0] * self.doubleBufferCount
        self.indexOffsetsFlipped = [False] * self.doubleBufferCount
        self.constantOffset = bufferSetNumber * self.getBufferCount()

    def getBufferCount(self) -> int:
        return 2 * self.doubleBufferCount + self.minIndexValue

    def flipOffset(self, i: int) -> None:
        k = i - self.minIndexValue
        assert k >= 0
        if not self.indexOffsetsFlipped[k]:      # only flip once before reject / accept
            self.indexOffsets[k] = self.doubleBufferCount - self.indexOffsets[k]
            self.indexOffsetsFlipped[k] = True

    def getOffsetIndex(self, i: int) -> int:
        if i < self.minIndexValue:
            return i + self.constantOffset
        return self.indexOffsets[i - self.minIndexValue] + i + self.constantOffset

    def isSafeUpdate(self, i: int) -> bool:
        k = i - self.minIndexValue
        return self.storedIndexOffsets[k] != self.indexOffsets[k]

    def storeState(self) -> None:
        self.indexOffsetsFlipped = [False] * self.doubleBufferCount
        self.storedIndexOffsets = list(self.indexOffsets)

    def restoreState(self) -> None:
        self.indexOffsets, self.storedIndexOffsets = self.storedIndexOffsets, self.indexOffsets
        self.indexOffsetsFlipped = [False] * self.doubleBufferCount


class HomogenousSubstitutionModelDelegate:
    """One substitution model on every branch: 2 eigen slots, 2 matrices per node."""

    def __init__(self, tree: Tree, substitutionModel: SubstitutionModel, partitionNumber: int = 0):
        self.substitutionModel = substitutionModel
        self.nodeCount = tree.nodeCount
        self.eigenBufferHelper = BufferIndexHelper(1, 0, partitionNumber)
        self.matrixBufferHelper = BufferIndexHelper(self.nodeCount, 0, partitionNumber)

    def canReturnComplexDiagonalization(self) -> bool:
        return self.substitutionModel.canReturnComplexDiagonalization()

    def getEigenBufferCount(self) -> int:
        return self.eigenBufferHelper.getBufferCount()

    def getMatrixBufferCount(self) -> int:
        return self.matrixBufferHelper.getBufferCount()

    def getEigenIndex(self, bufferIndex: int) -> int:
        return self.eigenBufferHelper.getOffsetIndex(bufferIndex)

    def getMatrixIndex(self, branchIndex: int) -> int:
        return self.matrixBufferHelper.getOffsetIndex(branchIndex)

    def getRootStateFrequencies(self) -> np.ndarray:
        return self.substitutionModel.getFrequencies()

    def updateSubstitutionModels(self, beagle, flip: bool) -> None:
        if flip:
            self.eigenBufferHelper.flipOffset(0)
        ed = self.substitutionModel.getEigenDecomposition()
        beagle.setEigenDecomposition(self.eigenBufferHelper.getOffsetIndex(0), ed.Evec, ed.Ievc, ed.Eval)

    def updateTransitionMatrices(self, beagle, branchIndices, edgeLengths, updateCount: int, flip: bool) -> None:
        probabilityIndices = np.empty(updateCount, dtype=np.int32)
        for i in range(updateCount):
            if flip:
                self.matrixBufferHelper.flipOffset(int(branchIndices[i]))
            probabilityIndices[i] = self.matrixBufferHelper.getOffsetIndex(int(branchIndices[i]))
        beagle.updateTransitionMatrices(self.eigenBufferHelper.getOffsetIndex(0), probabilityIndices,
                                        None, None, edgeLengths, updateCount)

    def storeState(self) -> None:
        self.eigenBufferHelper.storeState()
        self.matrixBufferHelper.storeState()

    def restoreState(self) -> None:
        self.eigenBufferHelper.restoreState()
        self.matrixBufferHelper.restoreState()


class PartialsRescalingScheme:
    DEFAULT = "default"
    NONE = "none"
    DYNAMIC = "dynamic"
    ALWAYS = "always"
    DELAYED = "delayed"
    AUTO = "auto"


RESCALE_FREQUENCY = 10000
RESCALE_TIMES = 1


class BeagleDataLikelihoodDelegate:
    """Re-enactment of BeagleDataLikelihoodDelegate.java for the single-partition path.

    ``beagleFactory(tipCount, partialsBufferCount, compactBufferCount, stateCount, patternCount,
    eigenBufferCount, matrixBufferCount, categoryCount, scaleBufferCount, resourceList,
    preferenceFlags, requirementFlags)`` returns an object with the ``beagle.Beagle`` methods and
    a ``getDetails()`` carrying ``.flags`` (BeagleFactory.loadBeagleInstance, BDLD:439-452).
    """

    def __init__(self, tree: Tree, patternList: Patterns, substitutionModel: SubstitutionModel,
                 siteRateModel: GammaSiteRateModel, beagleFactory: Callable, useAmbiguities: bool = False,
                 rescalingScheme: str = PartialsRescalingScheme.DEFAULT,
                 delayRescalingUntilUnderflow: bool = True, resourceList=None,
                 preferenceFlags: int = 0, requirementFlags: int = 0,
                 rescalingFrequency: int = RESCALE_FREQUENCY, stateSetFn=None, usePreOrder: bool = False):
        self.patternList = patternList
        self.patternCount = patternList.patternCount
        self.stateCount = patternList.stateCount
        self.siteRateModel = siteRateModel
        self.categoryCount = siteRateModel.getCategoryCount()
        self.nodeCount = tree.nodeCount
        self.tipCount = tree.tipCount
        self.internalNodeCount = self.nodeCount - self.tipCount

        self.branchUpdateIndices = np.zeros(self.nodeCount, dtype=np.int32)
        self.branchLengths = np.zeros(self.nodeCount)
        self.scaleBufferIndices = [0] * self.internalNodeCount
        self.storedScaleBufferIndices = [0] * self.internalNodeCount
        self.operations = np.zeros(self.internalNodeCount * OPERATION_TUPLE_SIZE, dtype=np.int32)

        self.firstRescaleAttempt = True
        self.isRestored = False
        self.useAmbiguities = useAmbiguities

        compactPartialsCount = 0 if useAmbiguities else self.tipCount
        self.partialBufferHelper = BufferIndexHelper(self.nodeCount, self.tipCount)
        self.scaleBufferHelper = BufferIndexHelper(self.internalNodeCount + 1, 0)
        self.evolutionaryProcessDelegate = HomogenousSubstitutionModelDelegate(tree, substitutionModel)

        numPartials = self.partialBufferHelper.getBufferCount()
        numScaleBuffers = self.scaleBufferHelper.getBufferCount()
        numMatrices = self.evolutionaryProcessDelegate.getMatrixBufferCount()
        self.usePreOrder = usePreOrder
        if usePreOrder:        # BDLD:239-244: one pre-order partial per node, cached infinitesimal matrices
            numPartials += self.nodeCount
            numScaleBuffers += self.nodeCount - 1
            numMatrices += 2 * self.evolutionaryProcessDelegate.getEigenBufferCount()

        self.rescalingScheme = rescalingScheme
        self.delayRescalingUntilUnderflow = delayRescalingUntilUnderflow
        if resourceList is not None and resourceList[0] > 0:
            preferenceFlags |= FLAG_PROCESSOR_GPU
        if self.rescalingScheme == PartialsRescalingScheme.DEFAULT:
            self.rescalingScheme = PartialsRescalingScheme.DYNAMIC
        if self.rescalingScheme == PartialsRescalingScheme.DELAYED:
            self.delayRescalingUntilUnderflow = True
            self.rescalingScheme = PartialsRescalingScheme.ALWAYS
        self.useAutoScaling = False
        if self.rescalingScheme == PartialsRescalingScheme.AUTO:
            preferenceFlags |= FLAG_SCALING_AUTO
            self.useAutoScaling = True
        self.rescalingFrequency = rescalingFrequency
        if self.evolutionaryProcessDelegate.canReturnComplexDiagonalization():
            requirementFlags |= FLAG_EIGEN_COMPLEX

        self.beagle = beagleFactory(self.tipCount, numPartials, compactPartialsCount, self.stateCount,
                                    self.patternCount, self.evolutionaryProcessDelegate.getEigenBufferCount(),
                                    numMatrices, self.categoryCount, numScaleBuffers, resourceList,
                                    preferenceFlags, requirementFlags)
        details = self.beagle.getDetails()
        self.instanceFlags = details.flags if details is not None else FLAG_FRAMEWORK_CPU
        if self.useAutoScaling and not (self.instanceFlags & FLAG_SCALING_AUTO):
            # "Auto rescaling not supported in BEAGLE, using dynamic" (BDLD:538-544)
            self.rescalingScheme = PartialsRescalingScheme.DYNAMIC
            self.useAutoScaling = False

        for i in range(self.tipCount):
            if useAmbiguities:
                self._setPartials(i, stateSetFn)
            else:
                self.beagle.setTipStates(i, np.ascontiguousarray(patternList.states[i], dtype=np.int32))
        self.patternWeights = np.ascontiguousarray(patternList.weights, dtype=np.float64)
        self.beagle.setPatternWeights(self.patternWeights)

        self.everUnderflowed = False
        self.updateSubstitutionModel = True
        self.updateSiteModel = True
        self.updateRootFrequency = True
        self.useScaleFactors = False
        self.recomputeScaleFactors = False
        self.rescalingCount = 0
        self.rescalingCountInner = 0
        self.initialEvaluation = True
        self.underflowHandling = 0
        self.flip = True
        self.totalMatrixUpdateCount = 0
        self.totalPartialsUpdateCount = 0
        self.totalEvaluationCount = 0

    # BDLD:638-681 -- tip partials from ambiguity state sets, replicated per category
    def _setPartials(self, tip: int, stateSetFn) -> None:
        P, S, C = self.patternCount, self.stateCount, self.categoryCount
        part = np.zeros((P, S))
        for p in range(P):
            part[p] = stateSetFn(int(self.patternList.states[tip, p]))
        self.beagle.setPartials(tip, np.ascontiguousarray(np.broadcast_to(part, (C, P, S))).reshape(-1))

    def getOptimalTraversalType(self) -> str:
        return "POST_ORDER" if (self.instanceFlags & FLAG_FRAMEWORK_CPU) else "REVERSE_LEVEL_ORDER"

    def makeDirty(self) -> None:
        self.updateSiteModel = True
        self.updateSubstitutionModel = True
        self.updateRootFrequency = True

    # ---- BDLD:734-1018 -----------------------------------------------------------------------
    def calculateLikelihood(self, branchOperations: List[Tuple[int, float]],
                            nodeOperations: List[Tuple[int, int, int]], rootNodeNumber: int) -> float:
        S = PartialsRescalingScheme
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
                    self.rescalingCountInner += 1
                    raise LikelihoodRescalingException()
                if self.initialEvaluation:
                    if self.underflowHandling < 1:
                        self.underflowHandling += 1
                    elif self.underflowHandling == 1:
                        self.recomputeScaleFactors = True
                        self.underflowHandling += 1
                        self.initialEvaluation = False
                self.rescalingCount += 1

        branchUpdateCount = 0
        for branchNumber, branchLength in branchOperations:
            self.branchUpdateIndices[branchUpdateCount] = branchNumber
            self.branchLengths[branchUpdateCount] = branchLength
            branchUpdateCount += 1

        beagle = self.beagle
        epd = self.evolutionaryProcessDelegate
        if self.updateSubstitutionModel:
            epd.updateSubstitutionModels(beagle, self.flip)
        if self.updateSiteModel:
            categoryRates = self.siteRateModel.getCategoryRates()
            if categoryRates is None:
                return -math.inf
            beagle.setCategoryRates(np.ascontiguousarray(categoryRates, dtype=np.float64))
            beagle.setCategoryWeights(0, np.ascontiguousarray(self.siteRateModel.getCategoryProportions(),
                                                              dtype=np.float64))
        if self.updateRootFrequency:
            beagle.setStateFrequencies(0, np.ascontiguousarray(epd.getRootStateFrequencies(), dtype=np.float64))
        if branchUpdateCount > 0:
            epd.updateTransitionMatrices(beagle, self.branchUpdateIndices, self.branchLengths,
                                         branchUpdateCount, self.flip)
        self.totalMatrixUpdateCount += branchUpdateCount

        if self.flip:
            for nodeNum, _, _ in nodeOperations:
                self.partialBufferHelper.flipOffset(nodeNum)

        operationCount = len(nodeOperations)
        ops = self.operations
        k = 0
        for nodeNum, leftChild, rightChild in nodeOperations:
            ops[k] = self.partialBufferHelper.getOffsetIndex(nodeNum)
            if self.useScaleFactors:
                n = nodeNum - self.tipCount
                if self.recomputeScaleFactors:
                    self.scaleBufferHelper.flipOffset(n)
                    self.scaleBufferIndices[n] = self.scaleBufferHelper.getOffsetIndex(n)
                    ops[k + 1] = self.scaleBufferIndices[n]
                    ops[k + 2] = NONE
                else:
                    ops[k + 1] = NONE
                    ops[k + 2] = self.scaleBufferIndices[n]
            else:
                if self.useAutoScaling:
                    self.scaleBufferIndices[nodeNum - self.tipCount] = self.partialBufferHelper.getOffsetIndex(nodeNum)
                ops[k + 1] = NONE
                ops[k + 2] = NONE
            ops[k + 3] = self.partialBufferHelper.getOffsetIndex(leftChild)
            ops[k + 4] = epd.getMatrixIndex(leftChild)
            ops[k + 5] = self.partialBufferHelper.getOffsetIndex(rightChild)
            ops[k + 6] = epd.getMatrixIndex(rightChild)
            k += OPERATION_TUPLE_SIZE

        beagle.updatePartials(ops, operationCount, NONE)
        self.totalEvaluationCount += 1
        self.totalPartialsUpdateCount += operationCount

        rootIndex = self.partialBufferHelper.getOffsetIndex(rootNodeNumber)
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
        elif self.useAutoScaling:
            beagle.accumulateScaleFactors(np.asarray(self.scaleBufferIndices, dtype=np.int32),
                                          self.internalNodeCount, NONE)

        sumLogLikelihoods = np.zeros(1)
        beagle.calculateRootLogLikelihoods(np.array([rootIndex], dtype=np.int32), np.array([0], dtype=np.int32),
                                           np.array([0], dtype=np.int32),
                                           np.array([cumulateScaleBufferIndex], dtype=np.int32), 1,
                                           sumLogLikelihoods)
        logL = float(sumLogLikelihoods[0])

        if math.isnan(logL) or math.isinf(logL):
            self.everUnderflowed = True
            logL = -math.inf
            if self.firstRescaleAttempt and (self.delayRescalingUntilUnderflow or
                                             self.rescalingScheme == S.DELAYED):
                self.useScaleFactors = True
                self.recomputeScaleFactors = True
                self.firstRescaleAttempt = False      # only try to rescale once
                self.rescalingCount -= 1
            # turn off double-buffer flipping so the next call overwrites the underflowed buffers
            self.flip = False
            self.underflowHandling = 0
            raise LikelihoodUnderflowException()
        else:
            self.firstRescaleAttempt = True
            self.recomputeScaleFactors = False
            self.flip = True

        self.updateSubstitutionModel = False
        self.updateSiteModel = False
        self.updateRootFrequency = False
        return logL

    def getPartialBufferCount(self) -> int:
        return self.partialBufferHelper.getBufferCount()

    def getPartialBufferIndex(self, nodeNumber: int) -> int:
        return self.partialBufferHelper.getOffsetIndex(nodeNumber)

    def getSiteLogLikelihoods(self) -> np.ndarray:
        out = np.zeros(self.patternCount)
        self.beagle.getSiteLogLikelihoods(out)
        return out

    def getPartials(self, number: int) -> np.ndarray:
        out = np.zeros(self.patternCount * self.stateCount * self.categoryCount)
        self.beagle.getPartials(self.partialBufferHelper.getOffsetIndex(number), NONE, out)
        return out

    def storeState(self) -> None:
        self.partialBufferHelper.storeState()
        self.evolutionaryProcessDelegate.storeState()
        if self.useScaleFactors or self.useAutoScaling:
            self.scaleBufferHelper.storeState()
            self.storedScaleBufferIndices = list(self.scaleBufferIndices)
        self.flip = True
        self.isRestored = False

    def restoreState(self) -> None:
        self.updateSiteModel = True
        self.updateRootFrequency = True
        self.partialBufferHelper.restoreState()
        self.evolutionaryProcessDelegate.restoreState()
        if self.useScaleFactors or self.useAutoScaling:
            self.scaleBufferHelper.restoreState()
            self.scaleBufferIndices, self.storedScaleBufferIndices = \
                self.storedScaleBufferIndices, self.scaleBufferIndices
        self.isRestored = True

    def finalize(self) -> None:
        self.beagle.finalize()


MAX_UNDERFLOWS_BEFORE_ERROR = 100


class TreeDataLikelihood:
    """TreeDataLikelihood + LikelihoodTreeTraversal: dirty-node tracking, op-list building in
    post-order or reverse-level-order, and the underflow retry loop (TDL:330-368)."""

    def __init__(self, likelihoodDelegate: BeagleDataLikelihoodDelegate, tree: Tree):
        self.likelihoodDelegate = likelihoodDelegate
        self.tree = tree
        self.traversalType = likelihoodDelegate.getOptimalTraversalType()
        self.updateNode = np.ones(tree.nodeCount, dtype=bool)
        self.likelihoodKnown = False
        self.logLikelihood = 0.0
        self.branchOperations: List[Tuple[int, float]] = []
        self.nodeOperations: List[Tuple[int, int, int]] = []

    # -- TreeTraversal.java:66-105
    def updateAllNodes(self) -> None:
        self.updateNode[:] = True
        self.likelihoodKnown = False

    def updateNodeAndChildren(self, node: int) -> None:
        self.updateNode[node] = True
        for c in self.tree.child[node]:
            if c >= 0:
                self.updateNode[c] = True
        self.likelihoodKnown = False

    def makeDirty(self) -> None:
        self.likelihoodDelegate.makeDirty()
        self.updateAllNodes()

    # -- LikelihoodTreeTraversal.java:49-204
    def _dispatch(self) -> None:
        self.branchOperations = []
        self.nodeOperations = []
        tree = self.tree
        if self.traversalType == "POST_ORDER":
            self._postOrder()
        else:
            levels = {}
            self._levelOrder(levels)
            for key in sorted(levels.keys(), reverse=True):
                self.nodeOperations.extend(levels[key])

    def _postOrder(self) -> None:
        tree = self.tree
        # iterative version of traversePostOrder (Python recursion depth on comb-like trees)
        result = {}
        stack = [(tree.root, 0)]
        while stack:
            node, stage = stack.pop()
            if stage == 0:
                upd = False
                if tree.parent[node] >= 0 and self.updateNode[node]:
                    self.branchOperations.append((node, tree.branchLength(node)))
                    upd = True
                result[node] = upd
                if not tree.isExternal(node):
                    stack.append((node, 1))
                    stack.append((int(tree.child[node][1]), 0))
                    stack.append((int(tree.child[node][0]), 0))
            else:
                c1, c2 = int(tree.child[node][0]), int(tree.child[node][1])
                if result[c1] or result[c2]:
                    self.nodeOperations.append((node, c1, c2))
                    result[node] = True
        # NB: the recursive reference emits branch ops in pre-order visit order; so does this.

    def _levelOrder(self, levels) -> None:
        tree = self.tree
        result = {}
        stack = [(tree.root, 0, 0)]
        while stack:
            node, level, stage = stack.pop()
            if stage == 0:
                upd = False
                if tree.parent[node] >= 0 and self.updateNode[node]:
                    self.branchOperations.append((node, tree.branchLength(node)))
                    upd = True
                result[node] = upd
                if not tree.isExternal(node):
                    stack.append((node, level, 1))
                    stack.append((int(tree.child[node][1]), level + 1, 0))
                    stack.append((int(tree.child[node][0]), level + 1, 0))
            else:
                c1, c2 = int(tree.child[node][0]), int(tree.child[node][1])
                if result[c1] or result[c2]:
                    levels.setdefault(level, []).append((node, c1, c2))
                    result[node] = True

    # -- TreeDataLikelihood.java:145-181,330-368
    def getLogLikelihood(self) -> float:
        if not self.likelihoodKnown:
            self.logLikelihood = self._calculateLogLikelihood()
            self.updateNode[:] = False
            self.likelihoodKnown = True
        return self.logLikelihood

    def _calculateLogLikelihood(self) -> float:
        logL = -math.inf
        done = False
        underflowCount = 0
        while not done and underflowCount < MAX_UNDERFLOWS_BEFORE_ERROR:
            self._dispatch()
            try:
                logL = self.likelihoodDelegate.calculateLikelihood(self.branchOperations, self.nodeOperations,
                                                                   self.tree.root)
                done = True
            except LikelihoodException:
                self.updateNode[:] = True
                underflowCount += 1
        return logL

    def storeState(self) -> None:
        self.likelihoodDelegate.storeState()
        self._storedLogL = self.logLikelihood
        self._storedKnown = self.likelihoodKnown

    def restoreState(self) -> None:
        self.likelihoodDelegate.restoreState()
        self.logLikelihood = self._storedLogL
        self.likelihoodKnown = self._storedKnown


class DiscreteTraitBranchRateDelegate:
    """Re-enactment of the pre-order / branch-gradient route:
    preorder/AbstractBeagleGradientDelegate.java:108-149,206-233 (simulateRoot, vectorizeNodeOperations,
    updatePrePartials), preorder/AbstractBeagleBranchGradientDelegate.java:57-96 (calculateEdgeDifferentials),
    discrete/DiscreteTraitBranchRateDelegate.java:49-89 (rate-scaled infinitesimal matrix),
    SimulationTreeTraversal.java:78-122 (pre-order op list) and
    HomogenousSubstitutionModelDelegate.java:140-147 (differential matrix buffer index).
    getGradient() returns d logL / d (branch length) for every non-root node, in node-number order."""

    def __init__(self, tree: Tree, likelihoodDelegate: BeagleDataLikelihoodDelegate, substitutionModel: SubstitutionModel):
        assert likelihoodDelegate.usePreOrder
        self.tree = tree
        self.likelihoodDelegate = likelihoodDelegate
        self.beagle = likelihoodDelegate.beagle
        self.substitutionModel = substitutionModel
        self.siteRateModel = likelihoodDelegate.siteRateModel
        self.preOrderPartialOffset = likelihoodDelegate.getPartialBufferCount()
        epd = likelihoodDelegate.evolutionaryProcessDelegate
        self.firstDerivativeMatrixIndex = epd.getMatrixBufferCount() + epd.getEigenIndex(0)

    def getPreOrderPartialIndex(self, node: int) -> int:
        return self.preOrderPartialOffset + node

    def _preOrderOperations(self):
        tree, ops = self.tree, []
        stack = [(tree.root, -1, -1)]
        while stack:
            node, parent, sibling = stack.pop()
            if parent >= 0:
                ops.append((parent, node, sibling))
            if not tree.isExternal(node):
                c1, c2 = int(tree.child[node][0]), int(tree.child[node][1])
                stack.append((c2, node, c1))
                stack.append((c1, node, c2))
        return ops

    def simulate(self) -> None:
        d, b, tree = self.likelihoodDelegate, self.beagle, self.tree
        P, S, C = d.patternCount, d.stateCount, d.categoryCount
        freqs = np.asarray(d.evolutionaryProcessDelegate.getRootStateFrequencies(), dtype=np.float64)
        b.setPartials(self.getPreOrderPartialIndex(tree.root), np.tile(freqs, P * C))        # simulateRoot
        epd = d.evolutionaryProcessDelegate
        ops = []
        for parent, node, sibling in self._preOrderOperations():
            ops += [self.getPreOrderPartialIndex(node), NONE, NONE, self.getPreOrderPartialIndex(parent),
                    epd.getMatrixIndex(node), d.getPartialBufferIndex(sibling), epd.getMatrixIndex(sibling)]
        b.updatePrePartials(np.asarray(ops, dtype=np.int32), len(ops) // 7, NONE)

    def cacheDifferentialMassMatrix(self) -> None:
        q = self.substitutionModel.infinitesimalMatrix().reshape(-1)
        rates = self.siteRateModel.getCategoryRates()
        scaled = np.concatenate([q * r for r in rates])
        self.beagle.setDifferentialMatrix(self.firstDerivativeMatrixIndex, scaled)

    def getGradient(self) -> np.ndarray:
        tree, d = self.tree, self.likelihoodDelegate
        self.simulate()
        self.cacheDifferentialMassMatrix()
        nodes = [n for n in range(tree.nodeCount) if n != tree.root]
        post = np.asarray([d.getPartialBufferIndex(n) for n in nodes], dtype=np.int32)
        pre = np.asarray([self.getPreOrderPartialIndex(n) for n in nodes], dtype=np.int32)
        der = np.full(len(nodes), self.firstDerivativeMatrixIndex, dtype=np.int32)
        first, firstSquared = np.zeros(len(nodes)), np.zeros(len(nodes))
        self.beagle.calculateEdgeDifferentials(post, pre, der, np.zeros(1, dtype=np.int32), len(nodes), None,
                                               first, firstSquared)
        return first


class SubstitutionModelCrossProductDelegate(DiscreteTraitBranchRateDelegate):
    """discrete/SubstitutionModelCrossProductDelegate.java:85-181 (coverWholeTree + getNodeDerivatives) for a single
    substitution model: the S x S cross-product differentials of the whole tree, the input of
    AbstractLogAdditiveSubstitutionModelGradient.java:239-270."""

    def getBranchLength(self, node: int) -> float:
        return self.tree.branchLength(node)

    def getCrossProducts(self) -> np.ndarray:
        tree, d = self.tree, self.likelihoodDelegate
        self.simulate()
        nodes = [n for n in range(tree.nodeCount) if n != tree.root]
        post = np.asarray([d.getPartialBufferIndex(n) for n in nodes], dtype=np.int32)
        pre = np.asarray([self.getPreOrderPartialIndex(n) for n in nodes], dtype=np.int32)
        lengths = np.asarray([self.getBranchLength(n) for n in nodes], dtype=np.float64)
        first = np.zeros(d.stateCount * d.stateCount)
        zero = np.zeros(1, dtype=np.int32)
        self.beagle.calculateCrossProductDifferentials(post, pre, zero, zero, lengths, len(nodes), first, None)
        return first.reshape(d.stateCount, d.stateCount)
