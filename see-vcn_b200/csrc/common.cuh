// Shared helpers for the seevcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/seevcn_b200.h"

// Multiprocessors of the CURRENT device (148 on B200: 2 dies x 74), queried once per device (abi.cu).
int seevcn_num_sms();

void seevcn_set_error(const char* fmt, ...);

#define SEEVCN_REQUIRE(cond, ...)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            seevcn_set_error(__VA_ARGS__);                          \
            return SEEVCN_E_INVALID;                                \
        }                                                           \
    } while (0)

#define SEEVCN_CUDA_CHECK(expr)                                                        \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            seevcn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                             cudaGetErrorString(_e));                                  \
            return SEEVCN_E_CUDA;                                                      \
        }                                                                              \
    } while (0)

void seevcn_count_launch();
#define SEEVCN_LAUNCH_CHECK()                      \
    do {                                           \
        seevcn_count_launch();                     \
        SEEVCN_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

// Optional CUDA-event timing of a launch group (see abi.cu; off unless seevcn_prof_enable(1)).
struct SeevcnProfScope {
    SeevcnProfScope(const char* name, cudaStream_t st);
    ~SeevcnProfScope();
    cudaStream_t st_;
    long slot_;
};
#define SEEVCN_PROF(name, st) SeevcnProfScope _seevcn_prof_scope(name, st)

static inline cudaStream_t as_stream(seevcn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
static inline T div_up(T a, T b) { return (a + b - 1) / b; }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Streaming (read-once) 128-bit load that does not allocate in L1.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Stage `nfl` floats starting at src into smem so that s[mis + j] = src[j], using aligned
// 128-bit loads for the body.  Returns mis (0..3).
__device__ __forceinline__ int stage_floats(float* s, const float* __restrict__ src, int nfl) {
    const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    const float* base = src - mis;                  // 16 B aligned
    const int total = mis + nfl;
    const int nvec = (total + 3) >> 2;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const int j0 = v * 4;
        if (j0 >= mis && j0 + 4 <= total) {
            float4 q = ld_stream_f4(reinterpret_cast<const float4*>(base + j0));
            *reinterpret_cast<float4*>(s + j0) = q;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                if (j >= mis && j < total) s[j] = base[j];
            }
        }
    }
    return mis;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ---- shared by the resampling kernels (crop.cu, isolate.cu) ----
// Pseudo-random permutation of [0, n) evaluated point-wise: a 4-round Feistel network on
// m = ceil(log2 n) bits (made even) with cycle walking.  perm(j) for j = 0..n_points-1 are the first
// n_points entries of one permutation — exactly what ResamplePoints draws
// (np.random.permutation(len)[:n], data_transforms.py:258-260), without a host RNG or a sort.
__host__ __device__ inline unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ inline unsigned feistel_perm(unsigned j, unsigned n, unsigned key) {
    unsigned bits = 2;
    while ((1u << bits) < n) bits += 2;            // even number of bits >= log2 n
    const unsigned half = bits >> 1, hmask = (1u << half) - 1u;
    unsigned v = j;
    do {
        unsigned l = v >> half, r = v & hmask;
#pragma unroll
        for (unsigned round = 0; round < 4; ++round) {
            const unsigned f = mix32(r ^ (key + 0x9e3779b9u * (round + 1))) & hmask;
            const unsigned nl = r; r = l ^ f; l = nl;
        }
        v = (l << half) | r;
    } while (v >= n);                               // cycle walking keeps it a bijection on [0, n)
    return v;
}
