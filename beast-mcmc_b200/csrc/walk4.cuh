// walk4.cuh -- device helpers and launch arguments shared by the 4-state walk kernels (kernels.cu, walk4e.cu).
#pragma once
#include "engine.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
// read-only path for data produced by an EARLIER launch (matrices)
__device__ __forceinline__ void ldg256_nc(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void stg256(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

struct WalkArgs {
    const Op4* ops;
    const int4* subs;          // [subtree] = (first op, one-past-last op, first pattern, one-past-last pattern)
    double* partials;          // slab base
    size_t stride;             // elements per slot
    const uint8_t* states;     // [tip][Ppad]
    const double* mats;        // [matrix][4][CP][4]
    double* scale;             // [buffer][Ppad]
    int S, C, Ppad, logScalers;
    size_t matStride;          // elements per matrix buffer
    int matMmaOffset;          // offset of the [c][i][j] + [c][j][i] copies inside a matrix buffer
    // eigen-form walk (walk4e.cu): per-branch spectra exp(lambda_k r_c t) [matrix][CP][4] in HBM, and the ONE eigen system
    // of the list by value -- kernel parameters live in the constant bank, so V / V^-1 cost neither registers nor LSU
    // traffic (DFMA takes them as uniform-register operands)
    const double* evecs;
    double V[16];              // Evec[i][k], rows/columns >= S zero
    double Vi[16];             // Ievc[k][j]
};

__device__ __forceinline__ Op4 loadOp(const Op4* p) {
    const int4* q = reinterpret_cast<const int4*>(p);
    int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    Op4 o;
    o.dest = a.x; o.c1 = a.y; o.c2 = a.z; o.m1 = a.w;
    o.m2 = b.x; o.sw = b.y; o.sr = b.z; o.cum = b.w;
    o.pBegin = c.x; o.pEnd = c.y; o.slots = (unsigned)c.z; o.pad_ = c.w;
    o.pfA = d.x; o.pfB = d.y; o.pfM1 = d.z; o.pfM2 = d.w;
    return o;
}

__device__ __forceinline__ void prefetchL1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

// non-volatile: a read-only load the scheduler may hoist freely
__device__ __forceinline__ void ldg256_ro(const double* p, double (&v)[4]) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
        : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

}  // namespace b200
