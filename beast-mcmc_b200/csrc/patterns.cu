// Site-pattern compression on the GPU: the step BEFORE the likelihood path.
//
// What it replaces: SitePatterns.addPatterns / addPattern (src/dr/evolution/alignment/SitePatterns.java:226-372) with
// CompressionType.UNIQUE_ONLY -- every alignment column is compared with the patterns found so far (exact equality,
// comparePatterns(.., false)), equal columns are merged (weight += site weight), new ones are appended.  The result is
// therefore: unique columns in FIRST-OCCURRENCE order, weights = multiplicities, sitePatternIndices[site] = index of the
// site's pattern.  The Java is O(sites x patterns x taxa); here it is a hash-table build:
//
//   k_hash_insert : thread per site: 64-bit hash of the column (reads coalesced across sites), open-addressing insert with
//                   atomicCAS on the key, atomicMin of the first site that carries the key
//   k_verify      : thread per site: the column equals the column of its table entry's first site (a hash collision would be
//                   reported as an error, never returned as a wrong answer)
//   k_flag/k_scan : flag first-occurrence sites, exclusive prefix sum over sites => pattern index in first-occurrence order
//   k_emit        : pattern index per site, integer multiplicities (integer atomics: order-independent), pattern columns
//
// HBM-bound byte/integer work: the alignment is read twice (hash, verify) plus once per unique column (emit).
#include "engine.h"

#include <cstdint>
#include <vector>

namespace b200 {

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}

__global__ void __launch_bounds__(256)
k_hash_insert(const int* __restrict__ states, int taxa, int sites, unsigned long long* __restrict__ keys,
              int* __restrict__ first, unsigned int tableMask, int* __restrict__ slotOfSite) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= sites) return;
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int t = 0; t < taxa; ++t) h = mix64(h, (uint64_t)(uint32_t)states[(size_t)t * sites + s]);
    if (h == 0) h = 1;                                   // 0 marks an empty slot
    unsigned int slot = (unsigned int)(h >> 17) & tableMask;
    while (true) {
        const unsigned long long old = atomicCAS(&keys[slot], 0ull, (unsigned long long)h);
        if (old == 0ull || old == (unsigned long long)h) break;
        slot = (slot + 1) & tableMask;
    }
    atomicMin(&first[slot], s);
    slotOfSite[s] = (int)slot;
}

__global__ void __launch_bounds__(256)
k_verify(const int* __restrict__ states, int taxa, int sites, const int* __restrict__ first,
         const int* __restrict__ slotOfSite, int* __restrict__ mismatch) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= sites) return;
    const int f = first[slotOfSite[s]];
    if (f == s) return;
    for (int t = 0; t < taxa; ++t)
        if (states[(size_t)t * sites + s] != states[(size_t)t * sites + f]) { atomicAdd(mismatch, 1); return; }
}

__global__ void __launch_bounds__(256)
k_flag(int sites, const int* __restrict__ first, const int* __restrict__ slotOfSite, int* __restrict__ flag) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s < sites) flag[s] = first[slotOfSite[s]] == s ? 1 : 0;
}

// exclusive prefix sum of flag[0..n) by ONE block (n is a site count: at most a few million); total -> *outTotal
__global__ void __launch_bounds__(1024)
k_scan(const int* __restrict__ flag, int n, int* __restrict__ prefix, int* __restrict__ outTotal) {
    __shared__ int warpSums[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? flag[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warpSums[w] = x;
        __syncthreads();
        if (w == 0) {
            int ws = warpSums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            warpSums[lane] = ws;                         // inclusive over warps
        }
        __syncthreads();
        const int before = carry + (w > 0 ? warpSums[w - 1] : 0) + x - v;
        if (i < n) prefix[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) *outTotal = carry;
}

__global__ void __launch_bounds__(256)
k_emit_index(int sites, const int* __restrict__ first, const int* __restrict__ slotOfSite,
             const int* __restrict__ prefix, int* __restrict__ patternOfSite, int* __restrict__ counts) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= sites) return;
    const int p = prefix[first[slotOfSite[s]]];
    patternOfSite[s] = p;
    atomicAdd(&counts[p], 1);
}

// grid (site blocks, taxa): copy the columns of first-occurrence sites into [taxon][pattern]
__global__ void __launch_bounds__(256)
k_emit_columns(const int* __restrict__ states, int sites, const int* __restrict__ flag, const int* __restrict__ prefix,
               int patternCount, int* __restrict__ outPatterns) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= sites || !flag[s]) return;
    const int t = blockIdx.y;
    outPatterns[(size_t)t * patternCount + prefix[s]] = states[(size_t)t * sites + s];
}

struct DeviceBuffers {
    std::vector<void*> ptrs;
    ~DeviceBuffers() { for (void* p : ptrs) cudaFree(p); }
    template <typename T> cudaError_t alloc(T** p, size_t n) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * (n ? n : 1));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

}  // namespace

// returns 0, or a negative BEAGLE error code
int compressSitePatterns(int device, int taxa, int sites, const int* hStates, int* hPatternOfSite, int* hPatterns,
                         double* hWeights, int* hPatternCount) {
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cudaGetLastError(); return e_ == cudaErrorMemoryAllocation ? -2 : -1; } } while (0)
    *hPatternCount = 0;
    if (sites == 0) return 0;
    CK(cudaSetDevice(device));
    unsigned int table = 1024;
    while (table < 2u * (unsigned int)sites) table <<= 1;
    DeviceBuffers buf;
    int *dStates, *first, *slotOfSite, *flag, *prefix, *patternOfSite, *counts, *misc, *dPatterns;
    unsigned long long* keys;
    CK(buf.alloc(&dStates, (size_t)taxa * sites));
    CK(buf.alloc(&keys, table));
    CK(buf.alloc(&first, table));
    CK(buf.alloc(&slotOfSite, sites));
    CK(buf.alloc(&flag, sites));
    CK(buf.alloc(&prefix, sites));
    CK(buf.alloc(&patternOfSite, sites));
    CK(buf.alloc(&counts, sites));
    CK(buf.alloc(&misc, 2));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    CK(cudaMemcpyAsync(dStates, hStates, sizeof(int) * (size_t)taxa * sites, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * table, st));
    CK(cudaMemsetAsync(first, 0x7f, sizeof(int) * table, st));              // 0x7f7f7f7f > any site index
    CK(cudaMemsetAsync(counts, 0, sizeof(int) * sites, st));
    CK(cudaMemsetAsync(misc, 0, sizeof(int) * 2, st));
    const int blocks = (sites + 255) / 256;
    k_hash_insert<<<blocks, 256, 0, st>>>(dStates, taxa, sites, keys, first, table - 1, slotOfSite);
    k_verify<<<blocks, 256, 0, st>>>(dStates, taxa, sites, first, slotOfSite, misc);
    k_flag<<<blocks, 256, 0, st>>>(sites, first, slotOfSite, flag);
    k_scan<<<1, 1024, 0, st>>>(flag, sites, prefix, misc + 1);
    k_emit_index<<<blocks, 256, 0, st>>>(sites, first, slotOfSite, prefix, patternOfSite, counts);
    CK(cudaGetLastError());
    int h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, misc, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h[0] != 0) return -1;                     // 64-bit hash collision between different columns: refuse, never guess
    const int patterns = h[1];
    CK(buf.alloc(&dPatterns, (size_t)taxa * patterns));
    k_emit_columns<<<dim3(blocks, taxa), 256, 0, st>>>(dStates, sites, flag, prefix, patterns, dPatterns);
    CK(cudaGetLastError());
    std::vector<int> hc(patterns);
    CK(cudaMemcpyAsync(hPatternOfSite, patternOfSite, sizeof(int) * sites, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hPatterns, dPatterns, sizeof(int) * (size_t)taxa * patterns, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hc.data(), counts, sizeof(int) * patterns, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int p = 0; p < patterns; ++p) hWeights[p] = (double)hc[p];
    *hPatternCount = patterns;
    return 0;
#undef CK
}

}  // namespace b200
