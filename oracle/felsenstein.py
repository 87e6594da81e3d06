"""ORACLE -- test infrastructure only. Never imported by the product path.

A plain numpy (fp64) restatement of the tree-likelihood arithmetic that
beast-dev/beast-mcmc delegates to the (un-vendored) BEAGLE library, exposed
through the same method names as the ``beagle.Beagle`` Java interface
(lib/beagle.jar, 42 methods) so that the caller re-enactment in
``beast_mcmc_b200.treedatalikelihood`` can drive either this oracle or the
CUDA engine with identical call sequences.

The arithmetic lives in beagle-dev/beagle-lib (branch ``v4_release``, pinned
only in the reference's CI: .github/workflows/ci.yml:19), which is absent from
/root/reference.  The restatement therefore follows the reference's own
in-tree statements of the same algorithm:

  * pruning kernels ............ src/dr/oldevomodel/treelikelihood/GeneralLikelihoodCore.java:52-203
  * category integration ....... GeneralLikelihoodCore.java:358-387
  * root log-likelihood ........ GeneralLikelihoodCore.java:395-408
  * per-pattern max rescaling .. src/dr/oldevomodel/treelikelihood/AbstractLikelihoodCore.java:406-442
                                 (BEAGLE rescales unconditionally when asked; no 1e-100 threshold)
  * scale accumulation ......... AbstractLikelihoodCore.java:451-459
  * P(t) from an eigen system .. src/dr/evomodel/substmodel/BaseSubstitutionModel.java:206-241 (abs() convention)
  * complex-pair eigen systems . src/dr/evomodel/substmodel/ComplexColtEigenSystem.java:71-139
  * op-tuple semantics ......... src/dr/evomodel/treedatalikelihood/BeagleDataLikelihoodDelegate.java:857-937

Pinned (tests/test_oracle_golden.py) against the ten log-likelihoods of
src/test/dr/evomodel/treedatalikelihood/TreeDataLikelihoodTest.java:131-314 and
the BEAGLE tiny-test value -1574.63623 (lib/beagle.jar BeagleFactory.main).
Scaled / by-partition / large-state paths have no literal pins in the
reference: for those the oracle is self-consistency-pinned only.
"""
from __future__ import annotations

import numpy as np

NONE = -1
OPERATION_TUPLE_SIZE = 7
PARTITION_OPERATION_TUPLE_SIZE = 9

# beagle.BeagleFlag masks used by the oracle (decoded from lib/beagle.jar)
FLAG_SCALERS_RAW = 1 << 9
FLAG_SCALERS_LOG = 1 << 10
FLAG_EIGEN_COMPLEX = 1 << 5


class OracleBeagle:
    """numpy re-statement of one BEAGLE instance (see module docstring)."""

    def __init__(self, tipCount, partialsBufferCount, compactBufferCount, stateCount,
                 patternCount, eigenBufferCount, matrixBufferCount, categoryCount,
                 scaleBufferCount, resourceList=None, preferenceFlags=0, requirementFlags=0):
        self.tipCount = tipCount
        self.S = stateCount
        self.P = patternCount
        self.C = categoryCount
        self.flags = preferenceFlags | requirementFlags
        self.log_scalers = bool(self.flags & FLAG_SCALERS_LOG)
        nbuf = partialsBufferCount + compactBufferCount
        self.partials = [None] * nbuf          # each [C][P][S]
        self.tipStates = [None] * nbuf         # each int[P] (compact)
        self.eigen = [None] * eigenBufferCount
        self.matrices = [None] * matrixBufferCount    # each [C][S][S]
        self.scale = [np.zeros(patternCount) for _ in range(scaleBufferCount)]
        # per-node buffers hold raw factors (logs under SCALERS_LOG); cumulative buffers -- anything touched by
        # reset / accumulate / remove / an in-list cumulative index -- always hold logs
        self.scaleIsLog = [self.log_scalers] * scaleBufferCount
        self.categoryRates = {0: np.ones(categoryCount)}
        self.categoryWeights = {}
        self.frequencies = {}
        self.patternWeights = np.ones(patternCount)
        self.patternPartitions = None
        self.partitionCount = 1
        self.siteLogL = np.zeros(patternCount)

    # ---- data upload -----------------------------------------------------------------
    def finalize(self):
        pass

    def setPatternWeights(self, w):
        self.patternWeights = np.array(w, dtype=np.float64).copy()

    def setPatternPartitions(self, partitionCount, patternPartitions):
        self.partitionCount = partitionCount
        self.patternPartitions = np.array(patternPartitions, dtype=np.int64).copy()

    def setTipStates(self, tipIndex, states):
        s = np.array(states, dtype=np.int64).copy()
        assert s.shape == (self.P,)
        self.tipStates[tipIndex] = s
        self.partials[tipIndex] = None

    def getTipStates(self, tipIndex, out):
        out[:] = self.tipStates[tipIndex]

    def setTipPartials(self, tipIndex, inPartials):
        # [P][S] replicated over categories by the engine
        p = np.array(inPartials, dtype=np.float64).reshape(self.P, self.S)
        self.partials[tipIndex] = np.broadcast_to(p, (self.C, self.P, self.S)).copy()
        self.tipStates[tipIndex] = None

    def setPartials(self, bufferIndex, inPartials):
        self.partials[bufferIndex] = np.array(inPartials, dtype=np.float64).reshape(
            self.C, self.P, self.S).copy()
        self.tipStates[bufferIndex] = None

    def getPartials(self, bufferIndex, scaleIndex, out):
        p = self.partials[bufferIndex]
        if scaleIndex != NONE:
            f = self.scale[scaleIndex]
            f = np.exp(f)          # cumulative buffers are always in log form
            p = p * f[None, :, None]
        out[:] = p.reshape(-1)

    def setEigenDecomposition(self, eigenIndex, evec, ievc, evals):
        S = self.S
        self.eigen[eigenIndex] = (np.array(evec, dtype=np.float64).reshape(S, S).copy(),
                                  np.array(ievc, dtype=np.float64).reshape(S, S).copy(),
                                  np.array(evals, dtype=np.float64).copy())

    def setStateFrequencies(self, idx, freqs):
        self.frequencies[idx] = np.array(freqs, dtype=np.float64).copy()

    def setCategoryWeights(self, idx, w):
        self.categoryWeights[idx] = np.array(w, dtype=np.float64).copy()

    def setCategoryRates(self, rates):
        self.categoryRates[0] = np.array(rates, dtype=np.float64).copy()

    def setCategoryRatesWithIndex(self, idx, rates):
        self.categoryRates[idx] = np.array(rates, dtype=np.float64).copy()

    def setTransitionMatrix(self, matrixIndex, inMatrix, paddedValue=0.0):
        self.matrices[matrixIndex] = np.array(inMatrix, dtype=np.float64).reshape(
            self.C, self.S, self.S).copy()

    def getTransitionMatrix(self, matrixIndex, out):
        out[:] = self.matrices[matrixIndex].reshape(-1)

    # ---- P(t) ---------------------------------------------------------------------------
    def _transition(self, eigenIndex, t, rates):
        """BaseSubstitutionModel.java:206-241 / ComplexColtEigenSystem.java:71-139."""
        evec, ievc, evals = self.eigen[eigenIndex]
        S = self.S
        out = np.empty((len(rates), S, S))
        complex_form = evals.shape[0] == 2 * S
        for c, r in enumerate(rates):
            d = t * r
            iexp = np.empty((S, S))
            i = 0
            while i < S:
                if not complex_form or evals[S + i] == 0.0:
                    iexp[i, :] = ievc[i, :] * np.exp(d * evals[i])
                    i += 1
                else:
                    b = evals[S + i]
                    expat = np.exp(d * evals[i])
                    ec = expat * np.cos(d * b)
                    es = expat * np.sin(d * b)
                    iexp[i, :] = ec * ievc[i, :] + es * ievc[i + 1, :]
                    iexp[i + 1, :] = ec * ievc[i + 1, :] - es * ievc[i, :]
                    i += 2
            # explicit k-ordered accumulation, as the Java loop does
            m = np.zeros((S, S))
            for k in range(S):
                m += evec[:, k][:, None] * iexp[k, :][None, :]
            out[c] = np.abs(m)
        return out

    def updateTransitionMatrices(self, eigenIndex, probabilityIndices, firstDerivativeIndices,
                                 secondDerivativeIndices, edgeLengths, count):
        rates = self.categoryRates[0]
        for k in range(count):
            self.matrices[probabilityIndices[k]] = self._transition(eigenIndex, edgeLengths[k], rates)

    def updateTransitionMatricesWithMultipleModels(self, eigenIndices, categoryRateIndices,
                                                   probabilityIndices, firstDerivativeIndices,
                                                   secondDerivativeIndices, edgeLengths, count):
        for k in range(count):
            rates = self.categoryRates[categoryRateIndices[k]]
            self.matrices[probabilityIndices[k]] = self._transition(eigenIndices[k], edgeLengths[k], rates)

    # ---- pruning ------------------------------------------------------------------------
    def _child_term(self, bufIndex, matIndex, sel):
        """sum_j M[c,i,j] * child[c,p,j]   (GeneralLikelihoodCore.java:171-203) or, for a
        compact tip, M[c,i,state_p] with 1 for state >= S (GeneralLikelihoodCore.java:52-107)."""
        M = self.matrices[matIndex]                    # [C][S][S]
        if self.tipStates[bufIndex] is not None:
            st = self.tipStates[bufIndex][sel]
            known = st < self.S
            safe = np.where(known, st, 0)
            term = np.transpose(M[:, :, safe], (0, 2, 1))        # [C][P][S(i)]
            return np.where(known[None, :, None], term, 1.0)
        x = self.partials[bufIndex][:, sel, :]         # [C][P][S]
        out = np.zeros_like(x)
        for j in range(self.S):                        # j-ordered accumulation like the Java loop
            out += M[:, None, :, j] * x[:, :, j][:, :, None]
        return out

    def _rescale(self, dest, sel, writeIdx, cumIdx):
        """AbstractLikelihoodCore.java:406-442, unconditional (BEAGLE semantics)."""
        d = self.partials[dest][:, sel, :]
        m = d.max(axis=(0, 2))
        m = np.where(m == 0.0, 1.0, m)
        self.partials[dest][:, sel, :] = d / m[None, :, None]
        logm = np.log(m)
        self.scale[writeIdx][sel] = logm if self.log_scalers else m
        self.scaleIsLog[writeIdx] = self.log_scalers
        if cumIdx != NONE:
            self.scale[cumIdx][sel] += logm
            self.scaleIsLog[cumIdx] = True

    def _update_one(self, op, sel, cumIdx):
        dest, sw, sr, c1, m1, c2, m2 = op[:7]
        if self.partials[dest] is None:
            self.partials[dest] = np.zeros((self.C, self.P, self.S))
            self.tipStates[dest] = None
        self.partials[dest][:, sel, :] = self._child_term(c1, m1, sel) * self._child_term(c2, m2, sel)
        if sw >= 0:
            self._rescale(dest, sel, sw, cumIdx)
        elif sr >= 0:
            f = self.scale[sr][sel]
            f = np.exp(f) if self.log_scalers else f
            self.partials[dest][:, sel, :] /= f[None, :, None]

    def updatePartials(self, operations, operationCount, cumulativeScaleIndex):
        ops = np.asarray(operations, dtype=np.int64).reshape(-1)
        sel = slice(None)
        for k in range(operationCount):
            self._update_one(ops[7 * k: 7 * k + 7], sel, cumulativeScaleIndex)

    def updatePartialsByPartition(self, operations, operationCount):
        ops = np.asarray(operations, dtype=np.int64).reshape(-1)
        for k in range(operationCount):
            op = ops[9 * k: 9 * k + 9]
            sel = np.nonzero(self.patternPartitions == op[7])[0]
            self._update_one(op, sel, op[8])

    def waitForPartials(self, destinationPartials, count):
        pass

    # ---- pre-order partials and edge derivatives ---------------------------------
    # Semantics recovered from the reference's call sites: op tuple = {pre[node], NONE, NONE, pre[parent],
    # matrix(node), post[sibling], matrix(sibling)} (preorder/AbstractBeagleGradientDelegate.java:206-220),
    # root pre-partial = frequencies (:139-149), and the defining identity the reference's own debug code
    # states (preorder/AbstractBeagleBranchGradientDelegate.java:97-150):
    #   site likelihood = sum_c w_c sum_j pre[node][c,p,j] * post[node][c,p,j]   for EVERY node.
    def setDifferentialMatrix(self, matrixIndex, inMatrix):
        self.setTransitionMatrix(matrixIndex, inMatrix)

    def transposeTransitionMatrices(self, inputIndices, resultIndices, count):
        for k in range(count):
            self.matrices[resultIndices[k]] = np.transpose(self.matrices[inputIndices[k]], (0, 2, 1)).copy()

    def convolveTransitionMatrices(self, firstIndices, secondIndices, resultIndices, count):
        # SubstitutionModelDelegate.java:303-470: P_result = P_first x P_second per category
        for k in range(count):
            self.matrices[resultIndices[k]] = np.einsum("cik,ckj->cij", self.matrices[firstIndices[k]],
                                                        self.matrices[secondIndices[k]])

    def addTransitionMatrices(self, firstIndices, secondIndices, resultIndices, count):
        for k in range(count):
            self.matrices[resultIndices[k]] = self.matrices[firstIndices[k]] + self.matrices[secondIndices[k]]

    def _pre_one(self, op, sel, cumIdx):
        dest, sw, sr, parentPre, m1, sib, m2 = op[:7]
        q = self.partials[parentPre][:, sel, :] * self._child_term(sib, m2, sel)    # at the parent: [C][P][S(i)]
        M1 = self.matrices[m1]                                                      # [C][i][j]
        out = np.zeros_like(q)
        for i in range(self.S):                                                     # down the branch: sum_i q_i M[i][j]
            out += q[:, :, i][:, :, None] * M1[:, None, i, :]
        if self.partials[dest] is None:
            self.partials[dest] = np.zeros((self.C, self.P, self.S))
        self.partials[dest][:, sel, :] = out
        self.tipStates[dest] = None
        if sw >= 0:
            self._rescale(dest, sel, sw, cumIdx)
        elif sr >= 0:
            f = self.scale[sr][sel]
            f = np.exp(f) if self.log_scalers else f
            self.partials[dest][:, sel, :] /= f[None, :, None]

    def updatePrePartials(self, operations, operationCount, cumulativeScaleIndex):
        ops = np.asarray(operations, dtype=np.int64).reshape(-1)
        for k in range(operationCount):
            self._pre_one(ops[7 * k: 7 * k + 7], slice(None), cumulativeScaleIndex)

    def updatePrePartialsByPartition(self, operations, operationCount):
        """9-int tuples like updatePartialsByPartition; the op applies to its partition's pattern window only."""
        ops = np.asarray(operations, dtype=np.int64).reshape(-1)
        for k in range(operationCount):
            op = ops[9 * k: 9 * k + 9]
            self._pre_one(op, np.nonzero(self.patternPartitions == op[7])[0], op[8])

    def _post_as_partials(self, idx):
        if self.tipStates[idx] is not None:
            st = self.tipStates[idx]
            out = np.ones((self.C, self.P, self.S))
            known = st < self.S
            onehot = np.zeros((self.P, self.S))
            onehot[np.nonzero(known)[0], st[known]] = 1.0
            out[:, known, :] = onehot[known][None, :, :]
            return out
        return self.partials[idx]

    def calculateEdgeDifferentials(self, postBufferIndices, preBufferIndices, derivativeMatrixIndices,
                                   categoryWeightsIndices, count, outDerivatives, outSumDerivatives,
                                   outSumSquaredDerivatives):
        w = self.categoryWeights[categoryWeightsIndices[0]]
        for e in range(count):
            post = self._post_as_partials(postBufferIndices[e])
            pre = self.partials[preBufferIndices[e]]
            D = self.matrices[derivativeMatrixIndices[e]]                       # [C][j][k]
            num = np.zeros(self.P)
            den = np.zeros(self.P)
            for c in range(self.C):
                Dpost = np.zeros((self.P, self.S))
                for k in range(self.S):
                    Dpost += D[c][None, :, k] * post[c][:, k][:, None]
                num += w[c] * (pre[c] * Dpost).sum(axis=1)
                den += w[c] * (pre[c] * post[c]).sum(axis=1)
            d = num / den
            if outDerivatives is not None:
                outDerivatives[e * self.P:(e + 1) * self.P] = d
            if outSumDerivatives is not None:
                outSumDerivatives[e] = float(np.dot(self.patternWeights, d))
            if outSumSquaredDerivatives is not None:
                outSumSquaredDerivatives[e] = float(np.dot(self.patternWeights, d * d))

    def calculateCrossProductDifferentials(self, postBufferIndices, preBufferIndices, categoryRatesIndices,
                                           categoryWeightsIndices, edgeLengths, count, outSumDerivatives,
                                           outSumSquaredDerivatives):
        """Call site: SubstitutionModelCrossProductDelegate.java:158-176.  The arithmetic is beagle-lib's (4.0.x,
        not vendored in the reference tree): per branch, pattern and category the outer product of the pre-order and
        post-order partials at the child end of the branch, weighted by weight_c * rate_c * t and divided by the
        pattern likelihood, summed into an S x S array that is ADDED to ``outSumDerivatives``.  Pinned here by the
        identity sum_ij out[ij] Q[ij] = d logL / d log(branch-length factor) (tests/test_preorder_oracle.py), which is
        how AbstractLogAdditiveSubstitutionModelGradient.java:220-227 consumes it."""
        assert outSumSquaredDerivatives is None
        w = self.categoryWeights[categoryWeightsIndices[0]]
        r = self.categoryRates[categoryRatesIndices[0]]
        acc = np.zeros((self.S, self.S))
        for e in range(count):
            post = self._post_as_partials(postBufferIndices[e])
            pre = self.partials[preBufferIndices[e]]
            den = np.zeros(self.P)
            num = np.zeros((self.P, self.S, self.S))
            for c in range(self.C):
                den += w[c] * (pre[c] * post[c]).sum(axis=1)
                num += w[c] * r[c] * pre[c][:, :, None] * post[c][:, None, :]
            acc += edgeLengths[e] * np.einsum("p,pij->ij", self.patternWeights / den, num)
        outSumDerivatives[:self.S * self.S] += acc.reshape(-1)

    # ---- scale factors ------------------------------------------------------------------
    def _logf(self, idx):
        return self.scale[idx] if self.scaleIsLog[idx] else np.log(self.scale[idx])

    def accumulateScaleFactors(self, scaleIndices, count, cumulativeScaleIndex):
        self.scaleIsLog[cumulativeScaleIndex] = True
        for k in range(count):
            self.scale[cumulativeScaleIndex] += self._logf(scaleIndices[k])

    def removeScaleFactors(self, scaleIndices, count, cumulativeScaleIndex):
        self.scaleIsLog[cumulativeScaleIndex] = True
        for k in range(count):
            self.scale[cumulativeScaleIndex] -= self._logf(scaleIndices[k])

    def accumulateScaleFactorsByPartition(self, scaleIndices, count, cumulativeScaleIndex, partitionIndex):
        self.scaleIsLog[cumulativeScaleIndex] = True
        sel = self.patternPartitions == partitionIndex
        for k in range(count):
            self.scale[cumulativeScaleIndex][sel] += self._logf(scaleIndices[k])[sel]

    def removeScaleFactorsByPartition(self, scaleIndices, count, cumulativeScaleIndex, partitionIndex):
        self.scaleIsLog[cumulativeScaleIndex] = True
        sel = self.patternPartitions == partitionIndex
        for k in range(count):
            self.scale[cumulativeScaleIndex][sel] -= self._logf(scaleIndices[k])[sel]

    def resetScaleFactors(self, cumulativeScaleIndex):
        self.scaleIsLog[cumulativeScaleIndex] = True
        self.scale[cumulativeScaleIndex][:] = 0.0

    def resetScaleFactorsByPartition(self, cumulativeScaleIndex, partitionIndex):
        self.scaleIsLog[cumulativeScaleIndex] = True
        self.scale[cumulativeScaleIndex][self.patternPartitions == partitionIndex] = 0.0

    def copyScaleFactors(self, destScalingIndex, srcScalingIndex):
        self.scale[destScalingIndex][:] = self.scale[srcScalingIndex]
        self.scaleIsLog[destScalingIndex] = self.scaleIsLog[srcScalingIndex]

    def getLogScaleFactors(self, scaleIndex, out):
        out[:] = self._logf(scaleIndex)

    def getScaleFactors(self, scaleIndex, out):
        out[:] = self.scale[scaleIndex]

    # ---- root ---------------------------------------------------------------------------
    def _site(self, rootIdx, wIdx, fIdx, cumIdx, sel=slice(None)):
        """GeneralLikelihoodCore.java:358-408: integrate over categories, then frequencies."""
        root = self.partials[rootIdx][:, sel, :]
        w = self.categoryWeights[wIdx]
        integrated = root[0] * w[0]
        for l in range(1, self.C):
            integrated = integrated + root[l] * w[l]
        freqs = self.frequencies[fIdx]
        s = np.zeros(integrated.shape[0])
        for i in range(self.S):
            s += freqs[i] * integrated[:, i]
        with np.errstate(divide="ignore", invalid="ignore"):
            site = np.log(s)
        if cumIdx != NONE:
            site = site + self.scale[cumIdx][sel]
        return site

    def calculateRootLogLikelihoods(self, bufferIndices, categoryWeightsIndices, stateFrequenciesIndices,
                                    cumulativeScaleIndices, count, outSumLogLikelihood):
        assert count == 1, "oracle restates the count==1 call BEAST issues (BDLD:934-935)"
        site = self._site(bufferIndices[0], categoryWeightsIndices[0], stateFrequenciesIndices[0],
                          cumulativeScaleIndices[0])
        self.siteLogL = site
        total = 0.0
        for p in range(self.P):          # strictly ordered sum (the reference sums sequentially)
            total += self.patternWeights[p] * site[p]
        outSumLogLikelihood[0] = total

    def calculateRootLogLikelihoodsByPartition(self, bufferIndices, categoryWeightsIndices,
                                               stateFrequenciesIndices, cumulativeScaleIndices,
                                               partitionIndices, partitionCount, count,
                                               outSumLogLikelihoodByPartition, outSumLogLikelihood):
        assert count == 1
        total = 0.0
        for k in range(partitionCount):
            part = partitionIndices[k]
            sel = np.nonzero(self.patternPartitions == part)[0]
            site = self._site(bufferIndices[k], categoryWeightsIndices[k], stateFrequenciesIndices[k],
                              cumulativeScaleIndices[k], sel)
            self.siteLogL[sel] = site
            s = float(np.dot(self.patternWeights[sel], site))
            outSumLogLikelihoodByPartition[k] = s
            total += s
        outSumLogLikelihood[0] = total

    def getSiteLogLikelihoods(self, out):
        out[:] = self.siteLogL
