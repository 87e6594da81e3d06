// walk4e.cu -- the 4-state (nucleotide) walk in EIGEN FORM: the default updatePartials kernel of S <= 4 instances.
//
// What bounded k_walk4 (profiles/r01_ncu_raw.txt, source page of the same capture): the L1/LSU data pipe at 70 % of peak
// wavefronts, 60 % of them matrix traffic -- every thread needs its category's 4x4 matrix in registers (128 B per child)
// and a compact-tip child costs one 32-B matrix column per PATTERN -- while HBM only saw the mandatory destination writes.
//
// Here no transition matrix is read at all.  With one real eigen system per list (what HomogenousSubstitutionModelDelegate
// hands over, HSMD:228-266)
//        P_c(t) x = V ( e ⊙ (V^-1 x) ),     e_k = exp(lambda_k r_c t)
// so a branch is 4 doubles per category ("spectrum", written by k_transition next to the matrices) and V, V^-1 are the same
// for every op of the launch: they travel BY VALUE in the kernel parameters, i.e. in the constant bank, and reach the FP64
// pipe as uniform-register operands (SASS: LDCU.128 + DFMA R, R, UR, R) -- zero registers, zero LSU wavefronts.
//   internal child : u = V^-1 x (16 FMA), w = e ⊙ u (4), y = |V w| (16)        -- 32 B of spectrum instead of 128 B of matrix
//   compact tip    : u_k = V^-1[k][s] (register selects on the state byte), then as above; gap/unknown -> y = 1
// |.| mirrors the reference's abs() on P(t) entries (BaseSubstitutionModel.java:236): for a tip the value IS |P[i][s]|
// with the reference's own summation order; partials stay non-negative.  The FP64 pipe goes from 12 % to ~40 % busy; the
// LSU data pipe keeps only what is irreducible: destination stores, the child cells that are not forwarded in registers,
// spectra, op records.  Lists the form does not cover (matrices set directly or convolved, complex pairs, several eigen
// systems in one list) run on k_walk4; pre-order lists keep their own kernel.  Results agree with the matrix form to
// rounding (tests/test_gpu_parity.py::test_walk_variants_agree, 1e-13).
//
// ALIGNED = every op spans whole 32-pattern groups [0, Ppad): no per-pattern predicates at all (padded columns hold
// harmless finite values and are never read back).  Partition windows take the predicated instance.
#include "engine.h"
#include "walk4.cuh"

namespace b200 {

namespace {

// TIP = 0: a compact-tip child goes through the same contraction with a one-hot x (no memory traffic, no extra code path);
// TIP = 1: its value is column s of the stored P matrix (one 32-B load per pattern, no arithmetic) -- the LSU / FP64
//          trade-off is measured, not guessed (B200_TIP_MODE)
//          TIP = 2: the columns of P for the five possible tip symbols (4 states + gap) of all categories are staged once per
//          (op, child) in a per-warp shared-memory table by a handful of lanes; every pattern then picks its 32-B column
//          with two LDS.128 -- no FP64 work, no dependent global load (the default)
__device__ __forceinline__ double absBits(double v) {       // |v| on the integer pipe (the FP64 pipe is the busy one here)
    return __hiloint2double(__double2hiint(v) & 0x7fffffff, __double2loint(v));
}

template <int CP, int R, bool ALIGNED, bool FIRST, int TIP>
__device__ __forceinline__ void childTermE(const WalkArgs& A, const double (&Vi)[16], int child, int matIdx, bool fromRegisters,
                                           int cc, size_t off0, int p0, bool catValid, int pBegin, int pEnd, double (&d)[R][4],
                                           double* tab) {
    constexpr int G = 32 / CP;
    const int S = A.S;
    const bool tip = child < 0;
    if (TIP == 2 && tip) {
        const int lane = threadIdx.x & 31;
        const uint8_t* t = A.states + (size_t)(-child - 1) * A.Ppad;
        int s[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = p0 + r * G;
            s[r] = (ALIGNED || (catValid && p >= pBegin && p < pEnd)) ? (int)__ldg(t + p) : 4;
        }
        // table [CP][5][4]: column s of P_c (the tip's contribution when it shows state s), s = 4 (and s >= S): all ones
        const double* m = A.mats + (size_t)matIdx * A.matStride;
        __syncwarp();                                              // the previous table of this slot is no longer read
#pragma unroll
        for (int q = lane; q < 5 * CP; q += 32) {
            const int c = q / 5, sym = q - 5 * c;
            double v[4];
            if (sym < S && sym < 4 && c < A.C) ldg256_ro(m + (sym * CP + c) * 4, v);
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = (i < S) ? 1.0 : 0.0;
            }
            double2* dst = reinterpret_cast<double2*>(tab + q * 4);
            dst[0] = make_double2(v[0], v[1]);
            dst[1] = make_double2(v[2], v[3]);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double2* col = reinterpret_cast<const double2*>(tab + (cc * 5 + min(s[r], 4)) * 4);
            const double2 lo = col[0], hi = col[1];
            if (FIRST) { d[r][0] = lo.x; d[r][1] = lo.y; d[r][2] = hi.x; d[r][3] = hi.y; }
            else { d[r][0] *= lo.x; d[r][1] *= lo.y; d[r][2] *= hi.x; d[r][3] *= hi.y; }
        }
        return;
    }
    if (TIP == 1 && tip) {
        const uint8_t* t = A.states + (size_t)(-child - 1) * A.Ppad;
        const double* m = A.mats + (size_t)matIdx * A.matStride + cc * 4;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = p0 + r * G;
            const int s = (ALIGNED || (catValid && p >= pBegin && p < pEnd)) ? (int)__ldg(t + p) : S;
            double v[4];
            if (s < S) ldg256_ro(m + 4 * CP * s, v);
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = (i < S) ? 1.0 : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) d[r][i] = FIRST ? v[i] : d[r][i] * v[i];
        }
        return;
    }
    double e[4];
    ldg256_ro(A.evecs + ((size_t)matIdx * CP + cc) * 4, e);
    const uint8_t* t = A.states + (size_t)(tip ? -child - 1 : 0) * A.Ppad;
    const double* xg = A.partials + (size_t)(tip ? 0 : child) * A.stride + off0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int p = p0 + r * G;
        const bool live = ALIGNED || (catValid && p >= pBegin && p < pEnd);
        double x[4], u[4], y[4];
        int s = 0;
        if (tip) {
            s = live ? (int)__ldg(t + p) : S;
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = (s == j) ? 1.0 : 0.0;
        } else if (fromRegisters) {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = d[r][i];
        } else if (live) {
            ldg256(xg + (size_t)r * G * 4, x);
        } else {
            x[0] = x[1] = x[2] = x[3] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            u[k] = (Vi[4 * k] * x[0] + Vi[4 * k + 1] * x[1] + Vi[4 * k + 2] * x[2] + Vi[4 * k + 3] * x[3]) * e[k];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            y[i] = absBits(A.V[4 * i] * u[0] + A.V[4 * i + 1] * u[1] + A.V[4 * i + 2] * u[2] + A.V[4 * i + 3] * u[3]);
        if (tip && s >= S) {                                       // gap / unknown: every state is compatible
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = (i < S) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) d[r][i] = FIRST ? y[i] : d[r][i] * y[i];
    }
}

template <int CP, int R, bool ALIGNED, int MINB, int TIP>
__global__ void __launch_bounds__(128, MINB)
k_walk4e(const WalkArgs A) {
    constexpr int G = 32 / CP;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int c = lane / G;
    const int4 range = __ldg(A.subs + blockIdx.y);
    const int p0 = range.z + warp * (G * R) + (lane % G);          // patterns p0 + r*G
    if (range.z + warp * (G * R) >= range.w) return;               // whole warp outside this subtree's pattern window
    const bool catValid = c < A.C;
    const int cc = catValid ? c : 0;
    const size_t off0 = ((size_t)cc * A.Ppad + p0) * 4;
    const int last = range.y - 1;
    // per-warp tip tables (TIP == 2): [child slot][CP][5][4] doubles
    __shared__ __align__(16) double tipTab[TIP == 2 ? 4 * 2 * CP * 20 : 2];
    double* tab1 = tipTab + (TIP == 2 ? ((threadIdx.x >> 5) * 2) * CP * 20 : 0);
    double* tab2 = tab1 + (TIP == 2 ? CP * 20 : 0);

    // V^-1 in vector registers (the tip selects and the first contraction read it), V stays in the constant bank: both in
    // uniform registers do not fit (64 > 63) and ptxas would spill.  The asm keeps ptxas from folding the copy back.
    double Vi[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) asm volatile("mov.f64 %0, %1;" : "=d"(Vi[q]) : "d"(A.Vi[q]));
    Op4 cur = loadOp(A.ops + range.x);
    double d[R][4];                                                // survives the loop: op k+1 may take it as its first child
#pragma unroll
    for (int r = 0; r < R; ++r) d[r][0] = d[r][1] = d[r][2] = d[r][3] = 0.0;
    for (int k = range.x; k <= last; ++k) {
        Op4 nxt;
        if (R == 1) nxt = loadOp(A.ops + min(k + 1, last));
        else if (lane == 0) prefetchL1(A.ops + min(k + 2, last));
        // look-ahead: what the NEXT op reads from memory (never this op's destination) starts its trip to L1 now
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int pf = w == 0 ? cur.pfA : cur.pfB;
            if (pf == 0) continue;
            if (pf & 1) {
                const uint8_t* t = A.states + (size_t)(pf >> 1) * A.Ppad + p0;
                if ((lane % G) == 0 && c == 0) prefetchL1(t);                  // G*R consecutive bytes: one line
            } else if (catValid) {
                const double* xg = A.partials + (size_t)((pf >> 1) - 1) * A.stride + off0;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (ALIGNED || p0 + r * G < A.Ppad) prefetchL1(xg + (size_t)r * G * 4);
            }
        }
        if (lane < 2) {
            const int mi = lane == 0 ? cur.pfM1 : cur.pfM2;
            if (mi >= 0) prefetchL1(A.evecs + (size_t)mi * CP * 4);            // all categories of a branch: one 128-B line
        } else if (TIP != 0 && lane >= 8 && lane < 16) {
            const int mi = lane < 12 ? cur.pfM1 : cur.pfM2;                     // matrix rows, should a child be a compact tip
            if (mi >= 0) prefetchL1(A.mats + (size_t)mi * A.matStride + (lane & 3) * 4 * CP);
        }
        childTermE<CP, R, ALIGNED, true, TIP>(A, Vi, cur.c1, cur.m1, (cur.pad_ & 2) != 0, cc, off0, p0, catValid, cur.pBegin, cur.pEnd, d, tab1);
        childTermE<CP, R, ALIGNED, false, TIP>(A, Vi, cur.c2, cur.m2, false, cc, off0, p0, catValid, cur.pBegin, cur.pEnd, d, tab2);
        if (R != 1) nxt = loadOp(A.ops + min(k + 1, last));
        double* dg = A.partials + (size_t)cur.dest * A.stride + off0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = p0 + r * G;
            const bool active = ALIGNED ? catValid : (catValid && p >= cur.pBegin && p < cur.pEnd);
            // ---- rescaling (AbstractLikelihoodCore.java:406-442, unconditional as in BEAGLE) -----
            if (cur.sw >= 0) {
                double m = active ? fmax(fmax(d[r][0], d[r][1]), fmax(d[r][2], d[r][3])) : 0.0;
#pragma unroll
                for (int sh = G; sh < 32; sh <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, sh));
                if (m == 0.0) m = 1.0;
                const double inv = 1.0 / m;
#pragma unroll
                for (int i = 0; i < 4; ++i) d[r][i] *= inv;
                if (active && c == 0) A.scale[(size_t)cur.sw * A.Ppad + p] = A.logScalers ? log(m) : m;
            } else if (cur.sr >= 0) {
                double f = active ? A.scale[(size_t)cur.sr * A.Ppad + p] : 1.0;
                if (A.logScalers) f = exp(f);
                const double inv = 1.0 / f;
#pragma unroll
                for (int i = 0; i < 4; ++i) d[r][i] *= inv;
            }
            if (active) stg256(dg + (size_t)r * G * 4, d[r]);
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_walk4p -- the same arithmetic with PER-WARP ASYNCHRONOUS OPERAND STAGING (the default for aligned lists, CP <= 8).
//
// What bounded k_walk4e after the matrices were gone (profiles/r02_walk4e_summary.md): every pipe below 50 %, but 38 % of the
// stall samples "long scoreboard" on first uses of small per-op operands -- the op record, the tips' state bytes, the
// spectra, the P columns of tip children -- i.e. a chain of dependent L1/L2 latencies per op with 16 warps per SM to hide it.
// Those operands are tiny, warp-uniform and known one op ahead, so each warp runs a two-deep cp.async (LDGSTS) pipeline into
// its own slice of shared memory: while op k computes, record k+2 and ALL small operands of op k+1 travel global -> shared
// without touching a register; op k+1 finds them with shared-memory latency.  Per op and warp: three LDGSTS instructions.
//   ring[4]            op records (64 B), fetched two ops ahead
//   stage[k & 1]       mat[child][2][5*CP][2] P block of a tip child: the contiguous [j][CP][i] block in HBM travels as 16-byte
//                                            pieces, one per lane, and every piece lands where the READS are conflict-free:
//                                            two half tables (states 0-1 / 2-3) of 16-byte entries indexed c*4 + j, so that
//                                            the 8 lanes of a quarter warp (same category, 8 patterns) hit 4 distinct
//                                            bank groups whatever their states -- with the HBM order kept ([j][c]: 128 B
//                                            between states) they collided up to 4-way and 63 % of the kernel's shared-memory
//                                            wavefronts were bank conflicts (profiles/r02_walk4e_summary.md).  Entries
//                                            4*CP + c = the gap column (written once).  A pattern picks its column with two LDS.128
//                      ev[child][CP][4]      spectrum of an internal child
//                      st[child][G*R]        state bytes of a tip child for this warp's patterns
// Child partials that are not forwarded in registers keep the look-ahead L1 prefetch.  Aligned lists only (every op spans
// [0, Ppad)), thin R = 1 phases included; pattern windows (by-partition lists) use k_walk4e.
__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gmemSrc) {
    const unsigned sAddr = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(sAddr), "l"(gmemSrc) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cpAsyncSmall(void* smemDst, const void* gmemSrc) {      // 4, 8 or 16 bytes
    const unsigned sAddr = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(sAddr), "l"(gmemSrc), "n"(BYTES) : "memory");
}

template <int CP, int R>
struct WarpStage {
    static constexpr int G = 32 / CP, NP = G * R;
    double mat[2][5 * CP * 4];
    double ev[2][CP * 4];
    alignas(16) unsigned char st[2][NP < 16 ? 16 : NP];
};

template <int CP, int R, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_walk4p(const WalkArgs A) {
    constexpr int G = 32 / CP, NP = G * R;
    constexpr int PIECE = NP < 16 ? NP : 16;                       // state bytes travel in 4-, 8- or 16-byte pieces
    static_assert(NP % PIECE == 0 && (PIECE == 4 || PIECE == 8 || PIECE == 16), "state-byte staging");
    __shared__ __align__(16) WarpStage<CP, R> stages[4][2];
    __shared__ __align__(16) Op4 rings[4][4];
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));            // volatile: never rematerialised as an S2R inside the loop
    const int wib = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int c = lane / G;
    const int4 range = __ldg(A.subs + blockIdx.y);
    const int pBase = range.z + warp * NP;                         // first pattern of this warp
    if (pBase >= range.w) return;
    const int p0 = pBase + (lane % G);                             // patterns p0 + r*G
    const bool catValid = c < A.C;
    const int cc = catValid ? c : 0;
    const size_t off0 = ((size_t)cc * A.Ppad + p0) * 4;
    const int last = range.y - 1, S = A.S;
    WarpStage<CP, R>* stage = stages[wib];
    Op4* ring = rings[wib];

    // the gap column (row 4 of every table), once
    for (int q = lane; q < 4 * CP * 4; q += 32) {
        const int tbl = q / (CP * 4), e = q % (CP * 4), gc = e >> 2, i = e & 3;
        stage[tbl >> 1].mat[tbl & 1][(i >> 1) * 10 * CP + 2 * (4 * CP + gc) + (i & 1)] = (i < S) ? 1.0 : 0.0;
    }
    // what op j reads beyond partials goes to stage j & 1; reads record j from the ring (it has arrived) ONCE -- the fields
    // the compute part needs travel on in registers (4 shared-memory reads per op instead of 20)
    struct Rec { int dest, c1, c2, sw, sr, flags, pfA, pfB; };
    auto issueOperands = [&](int j) -> Rec {
        const int4 rec = *reinterpret_cast<const int4*>(&ring[j & 3]);          // dest, c1, c2, m1
        const int4 rec2 = *(reinterpret_cast<const int4*>(&ring[j & 3]) + 1);   // m2, sw, sr, cum
        const int flags = ring[j & 3].pad_;
        const int2 pf = *reinterpret_cast<const int2*>(&ring[j & 3].pfA);
        const int m2 = rec2.x;
        WarpStage<CP, R>& sg = stage[j & 1];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int child = ch == 0 ? rec.y : rec.z, m = ch == 0 ? rec.w : m2;
            if (child < 0) {
                const double* src = A.mats + (size_t)m * A.matStride;          // [j][CP][i]: 4 * CP * 4 doubles, contiguous
#pragma unroll
                for (int q = lane; q < CP * 8; q += 32)                         // piece q = (j, c, half) -> half table, entry c*4 + j
                    cpAsync16(&sg.mat[ch][(q & 1) * 10 * CP + 2 * (((q >> 1) % CP) * 4 + q / (2 * CP))], src + 2 * q);
            } else if (lane < CP * 2) {
                cpAsync16(&sg.ev[ch][2 * lane], A.evecs + (size_t)m * CP * 4 + 2 * lane);
            }
        }
        // one more instruction: state bytes of the tip children (lanes 4 ..) and, on lanes 0-3, record j + 1
        constexpr int SL = NP / PIECE;
        if (lane < 4) {
            if (j + 1 <= last) cpAsync16(reinterpret_cast<char*>(&ring[(j + 1) & 3]) + 16 * lane,
                                         reinterpret_cast<const char*>(A.ops + j + 1) + 16 * lane);
        } else if (lane < 4 + 2 * SL) {
            const int ch = (lane - 4) / SL, piece = (lane - 4) % SL;
            const int child = ch == 0 ? rec.y : rec.z;
            if (child < 0)
                cpAsyncSmall<PIECE>(&sg.st[ch][PIECE * piece], A.states + (size_t)(-child - 1) * A.Ppad + pBase + PIECE * piece);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return Rec{rec.x, rec.y, rec.z, rec2.y, rec2.z, flags, pf.x, pf.y};
    };
    // prologue: records k0 (and k0+1 through issueOperands), then the operands of k0
    if (lane < 4) cpAsync16(reinterpret_cast<char*>(&ring[range.x & 3]) + 16 * lane,
                            reinterpret_cast<const char*>(A.ops + range.x) + 16 * lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    Rec nxt = issueOperands(range.x);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    double Vi[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) asm volatile("mov.f64 %0, %1;" : "=d"(Vi[q]) : "d"(A.Vi[q]));
    double d[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) d[r][0] = d[r][1] = d[r][2] = d[r][3] = 0.0;

    for (int k = range.x; k <= last; ++k) {
        const Rec cur = nxt;
        if (k + 1 <= last) nxt = issueOperands(k + 1);             // travels while op k computes
        const WarpStage<CP, R>& sg = stage[k & 1];
        // look-ahead for the child partials of op k+1 that are not forwarded (never this op's destination)
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int pf = w == 0 ? cur.pfA : cur.pfB;
            if (pf == 0 || (pf & 1) || !catValid) continue;
            const double* xg = A.partials + (size_t)((pf >> 1) - 1) * A.stride + off0;
#pragma unroll
            for (int r = 0; r < R; ++r) prefetchL1(xg + (size_t)r * G * 4);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int child = ch == 0 ? cur.c1 : cur.c2;
            if (child < 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int sym = sg.st[ch][(lane % G) + r * G];
                    const int e2 = 2 * (sym < S ? cc * 4 + sym : 4 * CP + cc);
                    const double2 lo = *reinterpret_cast<const double2*>(&sg.mat[ch][e2]);
                    const double2 hi = *reinterpret_cast<const double2*>(&sg.mat[ch][10 * CP + e2]);
                    if (ch == 0) { d[r][0] = lo.x; d[r][1] = lo.y; d[r][2] = hi.x; d[r][3] = hi.y; }
                    else { d[r][0] *= lo.x; d[r][1] *= lo.y; d[r][2] *= hi.x; d[r][3] *= hi.y; }
                }
            } else {
                const double2* ep = reinterpret_cast<const double2*>(&sg.ev[ch][cc * 4]);
                const double2 e01 = ep[0], e23 = ep[1];
                const double e[4] = {e01.x, e01.y, e23.x, e23.y};
                const bool fromRegisters = ch == 0 && (cur.flags & 2) != 0;
                const double* xg = A.partials + (size_t)child * A.stride + off0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double x[4], u[4], y[4];
                    if (fromRegisters) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = d[r][i];
                    } else {
                        ldg256(xg + (size_t)r * G * 4, x);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        u[q] = (Vi[4 * q] * x[0] + Vi[4 * q + 1] * x[1] + Vi[4 * q + 2] * x[2] + Vi[4 * q + 3] * x[3]) * e[q];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        y[i] = absBits(A.V[4 * i] * u[0] + A.V[4 * i + 1] * u[1] + A.V[4 * i + 2] * u[2] + A.V[4 * i + 3] * u[3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d[r][i] = ch == 0 ? y[i] : d[r][i] * y[i];
                }
            }
        }
        double* dg = A.partials + (size_t)cur.dest * A.stride + off0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int p = p0 + r * G;
            if (cur.sw >= 0) {                 // rescaling (AbstractLikelihoodCore.java:406-442, unconditional as in BEAGLE)
                double m = catValid ? fmax(fmax(d[r][0], d[r][1]), fmax(d[r][2], d[r][3])) : 0.0;
#pragma unroll
                for (int sh = G; sh < 32; sh <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, sh));
                if (m == 0.0) m = 1.0;
                const double inv = 1.0 / m;
#pragma unroll
                for (int i = 0; i < 4; ++i) d[r][i] *= inv;
                if (c == 0) A.scale[(size_t)cur.sw * A.Ppad + p] = A.logScalers ? log(m) : m;
            } else if (cur.sr >= 0) {
                double f = A.scale[(size_t)cur.sr * A.Ppad + p];
                if (A.logScalers) f = exp(f);
                const double inv = 1.0 / f;
#pragma unroll
                for (int i = 0; i < 4; ++i) d[r][i] *= inv;
            }
            if (catValid) stg256(dg + (size_t)r * G * 4, d[r]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // op k+1's operands and record k+2 have landed
        __syncwarp();
    }
}

template <int CP, int R, int MINB>
cudaError_t launchP(Instance* in, const WalkArgs& A, dim3 grid) {
    k_walk4p<CP, R, MINB><<<grid, 128, 0, in->stream>>>(A);
    return cudaGetLastError();
}

template <int CP, int R, bool ALIGNED, int MINB, int TIP>
cudaError_t launchK(Instance* in, const WalkArgs& A, dim3 grid) {
    k_walk4e<CP, R, ALIGNED, MINB, TIP><<<grid, 128, 0, in->stream>>>(A);
    return cudaGetLastError();
}

// shipped configuration per category count: R in {1, 4} x {aligned, windows}, tips through the shared-memory column table
// (CP <= 8) or the contraction; CP = 4 (the Gamma-4 workloads the metric is quoted on) additionally carries the tuning space
// behind B200_WALK_R / B200_WALK_MINB / B200_TIP_MODE
template <int CP, int R>
cudaError_t launchR(Instance* in, WalkArgs& A, int nSubs, int maxWindow, bool aligned) {
    constexpr int G = 32 / CP;
    constexpr int TIPD = CP <= 8 ? 2 : 0;
    const int warps = (maxWindow + G * R - 1) / (G * R);
    dim3 grid((warps + 3) / 4, nSubs);
    // predicate-free only when a warp's G*R patterns can never straddle the end of the padded pattern axis
    if (!aligned || in->Ppad % (G * R) != 0) return launchK<CP, R, false, 4, TIPD>(in, A, grid);
    if constexpr (CP <= 8 && G * R >= 4) {
        if ((R == 1 ? in->thinTipMode : in->tipMode) == 3) {     // per-warp asynchronous operand staging (k_walk4p)
            if constexpr (CP == 4) {
                // measured (profiles/r02_sweep_cfg2.txt): a launch bound of 3 blocks lets ptxas keep 120 registers without a
                // spill and 4 blocks still fit -- the fastest setting unless B200_WALK_MINB says otherwise
                const int minb = in->walkMinBlocksSet ? in->walkMinBlocks : 3;
                if (minb >= 6) return launchP<CP, R, 6>(in, A, grid);
                if (minb == 5) return launchP<CP, R, 5>(in, A, grid);
                if (minb == 3) return launchP<CP, R, 3>(in, A, grid);
            }
            return launchP<CP, R, 4>(in, A, grid);
        }
    }
    if constexpr (CP == 4 && R >= 2) {
        const int minb = in->walkMinBlocks, tip = in->tipMode;
        if (tip == 0) {
            if (minb >= 5) return launchK<CP, R, true, 5, 0>(in, A, grid);
            if (minb == 3) return launchK<CP, R, true, 3, 0>(in, A, grid);
            return launchK<CP, R, true, 4, 0>(in, A, grid);
        }
        if (tip == 1) {
            if (minb >= 5) return launchK<CP, R, true, 5, 1>(in, A, grid);
            return launchK<CP, R, true, 4, 1>(in, A, grid);
        }
        if (minb >= 6) return launchK<CP, R, true, 6, 2>(in, A, grid);
        if (minb == 5) return launchK<CP, R, true, 5, 2>(in, A, grid);
        if (minb == 3) return launchK<CP, R, true, 3, 2>(in, A, grid);
    }
    return launchK<CP, R, true, 4, TIPD>(in, A, grid);
}

template <int CP>
cudaError_t launchCP(Instance* in, WalkArgs& A, int nSubs, int maxWindow, bool aligned) {
    // a thin phase (few walks in flight) is latency-bound: one pattern group per thread gives the most warps per op
    const long walks = (long)nSubs * ((maxWindow + (32 / CP) * in->walkR - 1) / ((32 / CP) * in->walkR));
    if ((in->thinR1 && walks < (long)in->smCount * 8) || in->walkR == 1) return launchR<CP, 1>(in, A, nSubs, maxWindow, aligned);
    if constexpr (CP == 4) {
        if (in->walkR == 8) return launchR<CP, 8>(in, A, nSubs, maxWindow, aligned);
        if (in->walkR == 2) return launchR<CP, 2>(in, A, nSubs, maxWindow, aligned);
    }
    return launchR<CP, 4>(in, A, nSubs, maxWindow, aligned);
}

}  // namespace

// eigen: [V (16, row-major Evec[i][k]) | V^-1 (16, Ievc[k][j])], padded to 4 x 4 with zeros
cudaError_t launchWalk4E(Instance* in, const Op4* dOps, const int4* dSubs, int nSubs, int maxWindow, bool aligned,
                         const double* eigen) {
    if (nSubs <= 0) return cudaSuccess;
    WalkArgs A;
    A.ops = dOps; A.subs = dSubs; A.partials = in->partialsBase; A.stride = in->partialsElems;
    A.states = in->states8Base; A.mats = in->dMat; A.scale = in->dScale;
    A.S = in->S; A.C = in->C; A.Ppad = in->Ppad; A.logScalers = in->logScalers ? 1 : 0;
    A.matStride = in->matStride; A.matMmaOffset = 16 * in->matCP;
    A.evecs = in->dEvec;
    for (int q = 0; q < 16; ++q) { A.V[q] = eigen[q]; A.Vi[q] = eigen[16 + q]; }
    switch (in->matCP) {
#ifndef B200_W4E_QUICK
        case 1: return launchCP<1>(in, A, nSubs, maxWindow, aligned);
        case 2: return launchCP<2>(in, A, nSubs, maxWindow, aligned);
        case 8: return launchCP<8>(in, A, nSubs, maxWindow, aligned);
        case 16: return launchCP<16>(in, A, nSubs, maxWindow, aligned);
        case 32: return launchCP<32>(in, A, nSubs, maxWindow, aligned);
#endif
        default: return launchCP<4>(in, A, nSubs, maxWindow, aligned);
    }
}

cudaError_t updateWalk4EGraph(cudaGraphExec_t exec, const std::vector<cudaGraphNode_t>& kernelNodes, const double* eigen) {
    for (cudaGraphNode_t node : kernelNodes) {
        cudaKernelNodeParams kp;
        cudaError_t e = cudaGraphKernelNodeGetParams(node, &kp);
        if (e != cudaSuccess) return e;
        if (kp.kernelParams == nullptr || kp.kernelParams[0] == nullptr) return cudaErrorInvalidValue;
        WalkArgs A = *static_cast<const WalkArgs*>(kp.kernelParams[0]);       // every eigen-form walk kernel takes ONE WalkArgs
        for (int q = 0; q < 16; ++q) { A.V[q] = eigen[q]; A.Vi[q] = eigen[16 + q]; }
        void* args[1] = {&A};
        kp.kernelParams = args;
        kp.extra = nullptr;
        e = cudaGraphExecKernelNodeSetParams(exec, node, &kp);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace b200
