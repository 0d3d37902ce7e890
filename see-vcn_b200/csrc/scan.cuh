// Single-pass device-wide prefix sums (decoupled look-back), hand-written for the voxelizers: no CUB on the path.
//
// Protocol: every tile owns one 64-bit status word, written with ONE store so flag and value can never be seen torn:
//   bits 63..62  flag   0 = nothing yet, 1 = AGGREGATE (sum of this tile only), 2 = PREFIX (inclusive prefix up to this tile)
//   bits 61..0   value  (callers pack two 31-bit counters when they need two sums)
// A tile publishes its aggregate as soon as it is known, then walks back over its predecessors 32 at a time (one per
// lane) summing aggregates until it meets a PREFIX, and finally publishes its own inclusive prefix.  Tiles must be
// STARTED in index order (take the index from an atomic ticket) so that every predecessor a tile waits for is already
// running: forward progress then needs no co-residency assumption.  The status words must be zero before the launch.
#pragma once
#include "common.cuh"

namespace seevcn_scan {

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// Whole warp, uniform arguments.  Publishes `aggregate` for tile `tile`, returns the EXCLUSIVE prefix of the tile
// (sum of the aggregates of tiles 0..tile-1) to every lane and publishes the inclusive prefix.
__device__ __forceinline__ unsigned long long lookback_publish(unsigned long long* status, int tile, unsigned long long aggregate) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_status(status, kFlagPrefix | aggregate);
        return 0ull;
    }
    if (lane == 0) st_status(status + tile, kFlagAgg | aggregate);
    unsigned long long excl = 0ull;
    int j = tile - 1;
    while (true) {
        const int idx = j - lane;
        unsigned long long s = idx >= 0 ? ld_status(status + idx) : kFlagPrefix;
        while (__any_sync(0xffffffffu, (s >> 62) == 0ull)) {
            if ((s >> 62) == 0ull) s = ld_status(status + idx);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
        if (pm) {
            const int first = __ffs(pm) - 1;                       // nearest predecessor that already knows its prefix
            excl += warp_sum_u64(lane <= first ? (s & kValueMask) : 0ull);
            break;
        }
        excl += warp_sum_u64(s & kValueMask);
        j -= 32;
    }
    if (lane == 0) st_status(status + tile, kFlagPrefix | ((excl + aggregate) & kValueMask));
    return excl;
}

// Split form for callers that have useful work between the two steps: publish the aggregate now ...
__device__ __forceinline__ void publish_aggregate(unsigned long long* status, int tile, unsigned long long aggregate) {
    if ((threadIdx.x & 31) == 0) st_status(status + tile, (tile == 0 ? kFlagPrefix : kFlagAgg) | aggregate);
}
// ... and resolve the exclusive prefix later (publishes the inclusive prefix).
__device__ __forceinline__ unsigned long long resolve_prefix(unsigned long long* status, int tile, unsigned long long aggregate) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) return 0ull;
    unsigned long long excl = 0ull;
    int j = tile - 1;
    while (true) {
        const int idx = j - lane;
        unsigned long long s = idx >= 0 ? ld_status(status + idx) : kFlagPrefix;
        while (__any_sync(0xffffffffu, (s >> 62) == 0ull)) {
            if ((s >> 62) == 0ull) s = ld_status(status + idx);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
        if (pm) {
            const int first = __ffs(pm) - 1;
            excl += warp_sum_u64(lane <= first ? (s & kValueMask) : 0ull);
            break;
        }
        excl += warp_sum_u64(s & kValueMask);
        j -= 32;
    }
    if (lane == 0) st_status(status + tile, kFlagPrefix | ((excl + aggregate) & kValueMask));
    return excl;
}

// ---- exclusive scan of n uint32 values, in place or out of place ------------------------------------------------
// grid = ceil(n / kScanTile) CTAs of kScanThreads; status: one word per CTA (zeroed), ticket: one int (zeroed).
// total (may be NULL) receives the sum of all n values.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
exclusive_scan_u32_kernel(int n, const unsigned* __restrict__ in, unsigned* __restrict__ out,
                          unsigned long long* __restrict__ status, int* __restrict__ ticket, unsigned* __restrict__ total) {
    __shared__ int s_tile;
    __shared__ unsigned s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const int i0 = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned sum = 0;
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) { v[e] = i0 + e < n ? in[i0 + e] : 0u; sum += v[e]; }
    unsigned inc = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, s); if ((threadIdx.x & 31) >= s) inc += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned w = threadIdx.x < kScanThreads / 32 ? s_warp[threadIdx.x] : 0u;
        unsigned winc = w;
#pragma unroll
        for (int s = 1; s < kScanThreads / 32; s <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, winc, s); if (threadIdx.x >= s) winc += t; }
        if (threadIdx.x < kScanThreads / 32) s_warp[threadIdx.x] = winc - w;
        const unsigned agg = __shfl_sync(0xffffffffu, winc, kScanThreads / 32 - 1);
        const unsigned long long excl = lookback_publish(status, tile, agg);
        if (threadIdx.x == 0) {
            s_excl = excl;
            if (total && (long long)(tile + 1) * kScanTile >= n) *total = (unsigned)(excl + agg);
        }
    }
    __syncthreads();
    unsigned run = (unsigned)s_excl + s_warp[threadIdx.x >> 5] + inc - sum;
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) {
        if (i0 + e < n) out[i0 + e] = run;
        run += v[e];
    }
}

}  // namespace seevcn_scan
