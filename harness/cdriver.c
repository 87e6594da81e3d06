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
