/* cdriver.c -- bench infrastructure: replays a PREPARED sequence of BEAGLE C-ABI calls from C and times every evaluation
 * with clock_gettime, i.e. what the JVM's JNI thread does (beagle.jar -> libhmsbeagle-jni -> libhmsbeagle), without a
 * Python interpreter between the calls.  bench.py prepares the arrays; this file only issues
 *   beagleUpdateTransitionMatrices -> beagleUpdatePartials -> beagleCalculateRootLogLikelihoods
 * per evaluation (BeagleDataLikelihoodDelegate.calculateLikelihood, BDLD:734-1018) against include/libhmsbeagle_b200.h.
 * Build: gcc -O2 -fPIC -shared -I include -o harness/libcdriver.so harness/cdriver.c -L beast-mcmc_b200/csrc -lhmsbeagle */
#include <time.h>
#include "libhmsbeagle_b200.h"

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* evaluation e uses operations [opOffsets[e], opOffsets[e+1]) of `ops` (7 ints each), matrices
 * [matOffsets[e], matOffsets[e+1]) of probIdx / lengths, root buffer rootIdx[e].  seconds[round * nEvals + e] = wall time of
 * the three calls; returns the first non-zero BEAGLE return code (0 = success), *lastLogL = the last value. */
int cdriver_replay(int instance, int nEvals, int rounds, const int* opOffsets, const int* ops, const int* matOffsets,
                   const int* probIdx, const double* lengths, const int* rootIdx, int eigenIndex, int cumulativeScaleIndex,
                   double* seconds, double* lastLogL) {
    const int zero = 0, cum = cumulativeScaleIndex;
    double out = 0.0;
    for (int r = 0; r < rounds; ++r) {
        for (int e = 0; e < nEvals; ++e) {
            const double t0 = now_s();
            const int m0 = matOffsets[e], m1 = matOffsets[e + 1];
            int rc = beagleUpdateTransitionMatrices(instance, eigenIndex, probIdx + m0, 0, 0, lengths + m0, m1 - m0);
            if (rc != 0) return rc;
            rc = beagleUpdatePartials(instance, (const BeagleOperation*)(ops + 7 * (long)opOffsets[e]),
                                      opOffsets[e + 1] - opOffsets[e], BEAGLE_OP_NONE);
            if (rc != 0) return rc;
            rc = beagleCalculateRootLogLikelihoods(instance, rootIdx + e, &zero, &zero, &cum, 1, &out);
            if (rc != 0) return rc;
            seconds[(long)r * nEvals + e] = now_s() - t0;
        }
    }
    *lastLogL = out;
    return 0;
}

/* One FULL evaluation per step, as BeagleDataLikelihoodDelegate.calculateLikelihood issues it when model and tree are dirty
 * (BDLD:734-1018): eigen system, category rates / weights, frequencies up, every branch's matrix, the whole operation list,
 * (rescaling: reset + accumulate), root -- host buffers in, the log-likelihood out.  Step k uses buffer parity k & 1 (BEAST's
 * flip): ops / probIdx / scaleIdx hold both parities back to back.  seconds[k] = wall time of step k. */
int cdriver_full_evaluations(int instance, int steps, int stateCount, int nOps, const int* ops2 /* [2][nOps*7] */, int nMats,
                             const int* probIdx2 /* [2][nMats] */, const double* lengths, const int* rootIdx2, const double* evec,
                             const double* ievc, const double* eval, const double* rates, const double* weights,
                             const double* freqs, int scaling, const int* scaleIdx2 /* [2][nOps] */, const int* cumIdx2,
                             double* seconds, double* lastLogL) {
    const int zero = 0, none = BEAGLE_OP_NONE;
    double out = 0.0;
    (void)stateCount;
    for (int k = 0; k < steps; ++k) {
        const int p = k & 1;
        const double t0 = now_s();
        int rc = beagleSetEigenDecomposition(instance, p, evec, ievc, eval);
        if (rc == 0) rc = beagleSetCategoryRates(instance, rates);
        if (rc == 0) rc = beagleSetCategoryWeights(instance, 0, weights);
        if (rc == 0) rc = beagleSetStateFrequencies(instance, 0, freqs);
        if (rc == 0) rc = beagleUpdateTransitionMatrices(instance, p, probIdx2 + (long)p * nMats, 0, 0, lengths, nMats);
        if (rc == 0) rc = beagleUpdatePartials(instance, (const BeagleOperation*)(ops2 + (long)p * nOps * 7), nOps, BEAGLE_OP_NONE);
        const int* cum = &none;
        if (rc == 0 && scaling) {
            rc = beagleResetScaleFactors(instance, cumIdx2[p]);
            if (rc == 0) rc = beagleAccumulateScaleFactors(instance, scaleIdx2 + (long)p * nOps, nOps, cumIdx2[p]);
            cum = cumIdx2 + p;
        }
        if (rc == 0) rc = beagleCalculateRootLogLikelihoods(instance, rootIdx2 + p, &zero, &zero, cum, 1, &out);
        if (rc != 0) return rc;
        seconds[k] = now_s() - t0;
    }
    *lastLogL = out;
    return 0;
}
