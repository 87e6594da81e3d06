"""ORACLE -- test infrastructure only.  ctypes face of oracle/beagle_cpu.c (the C restatement of
the hot path used as checker and as bench.py's cpu_baseline / --impl reference arm), exposing the
same ``beagle.Beagle`` method names as the numpy oracle so the caller re-enactment can drive it."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NONE = -1


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_cpu.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with __graft_entry__.build())")
        lib = C.CDLL(path)
        lib.oc_create.restype = C.c_void_p
        lib.oc_create.argtypes = [C.c_int] * 10
        lib.oc_calculate_root_log_likelihoods.restype = C.c_double
        _LIB = lib
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Details:
    def __init__(self, flags):
        self.flags = flags


class OracleCpuBeagle:
    def __init__(self, tipCount, partialsBufferCount, compactBufferCount, stateCount, patternCount,
                 eigenBufferCount, matrixBufferCount, categoryCount, scaleBufferCount, resourceList=None,
                 preferenceFlags=0, requirementFlags=0, threads=1, reportFlags=1 << 27):
        self.lib = load()
        self.S, self.P, self.C = stateCount, patternCount, categoryCount
        logs = 1 if ((preferenceFlags | requirementFlags) & (1 << 10)) else 0
        self.h = C.c_void_p(self.lib.oc_create(tipCount, partialsBufferCount + compactBufferCount, stateCount,
                                               patternCount, eigenBufferCount, matrixBufferCount, categoryCount,
                                               scaleBufferCount, threads, logs))
        self._details = _Details(reportFlags)

    def getDetails(self):
        return self._details

    def finalize(self):
        if self.h:
            self.lib.oc_free(self.h)
            self.h = None

    def setPatternWeights(self, w):
        self.lib.oc_set_pattern_weights(self.h, _p(_d(w)))

    def setTipStates(self, tip, states):
        self.lib.oc_set_tip_states(self.h, tip, _p(_i(states)))

    def setPartials(self, idx, x):
        self.lib.oc_set_partials(self.h, idx, _p(_d(x)))

    def getPartials(self, idx, scaleIndex, out):
        assert scaleIndex == NONE
        self.lib.oc_get_partials(self.h, idx, _p(out))

    def setEigenDecomposition(self, idx, evec, ievc, evals):
        self.lib.oc_set_eigen(self.h, idx, _p(_d(evec)), _p(_d(ievc)), _p(_d(evals)))

    def setStateFrequencies(self, idx, f):
        assert idx == 0
        self.lib.oc_set_state_frequencies(self.h, _p(_d(f)))

    def setCategoryWeights(self, idx, w):
        assert idx == 0
        self.lib.oc_set_category_weights(self.h, _p(_d(w)))

    def setCategoryRates(self, r):
        self.lib.oc_set_category_rates(self.h, _p(_d(r)))

    def updateTransitionMatrices(self, eigenIndex, probabilityIndices, d1, d2, edgeLengths, count):
        self.lib.oc_update_transition_matrices(self.h, eigenIndex, _p(_i(probabilityIndices)), _p(_d(edgeLengths)), count)

    def updatePartials(self, operations, operationCount, cumulativeScaleIndex):
        self.lib.oc_update_partials(self.h, _p(_i(operations)), operationCount, cumulativeScaleIndex)

    def resetScaleFactors(self, cum):
        self.lib.oc_reset_scale_factors(self.h, cum)

    def accumulateScaleFactors(self, scaleIndices, count, cum):
        self.lib.oc_accumulate_scale_factors(self.h, _p(_i(scaleIndices)), count, cum)

    def getLogScaleFactors(self, idx, out):
        self.lib.oc_get_log_scale_factors(self.h, idx, _p(out))

    def calculateRootLogLikelihoods(self, bufferIndices, wIdx, fIdx, cumIdx, count, out):
        assert count == 1
        out[0] = self.lib.oc_calculate_root_log_likelihoods(self.h, int(bufferIndices[0]), int(cumIdx[0]))

    def getSiteLogLikelihoods(self, out):
        self.lib.oc_get_site_log_likelihoods(self.h, _p(out))


def factory(threads=1, reportFlags=1 << 27):
    def make(*args):
        return OracleCpuBeagle(*args, threads=threads, reportFlags=reportFlags)
    return make
