// Stage 3 — furthest point sampling.
//
// Replaces detector3d/pcdet/ops/pointnet2/pointnet2_batch/src/sampling_gpu.cu:100-260.
//
// The reference re-reads xyz (12 B) and temp (4+4 B) for every point from global memory on
// each of the M-1 rounds and reduces with a 10-barrier shared-memory tree.  Here, per object
// (one CTA):
//  * xyz is staged ONCE into shared memory as SoA (3*N floats, N <= 18,900 fits 227 KB);
//  * the running min-distance array lives in registers (<= 16 points per thread);
//  * the block argmax is a warp-shuffle reduction + one barrier per round (per-warp
//    partials are double-buffered in smem and every warp redundantly reduces them).
// HBM traffic drops from M*N*16 B to the compulsory 12*N + 4*M B per object; what is left
// is the serial chain of M-1 rounds (on-chip latency bound).
//
// Bit-exactness: same block size rule (cuda_utils.h:10-14), same strided point->thread
// map, the distance expression with the contraction nvcc 12.9 emits for the reference
// (fma(dz,dz, fma(dy,dy, dx*dx))), fminf, strict '>' inside a thread and the reference tree's
// tie rule across threads (smallest bit-reversed thread id, see better()), so indices match even
// on clouds with duplicate points.
#include <cmath>
#include "common.cuh"

namespace {

// ref: opt_n_threads, pointnet2_batch/src/cuda_utils.h:10-14
int ref_block_size(int n) {
    const int pow_2 = (int)(std::log(static_cast<double>(n)) / std::log(2.0));
    int bs = 1 << pow_2;
    if (bs > 1024) bs = 1024;
    if (bs < 1) bs = 1;
    return bs;
}

struct Cand { float v; int i; };

// The reference's shared-memory tree (__update, sampling_gpu.cu:93-98, strides bs/2..1) keeps the
// LOWER position on equal values; after the level with stride s position p holds the winner of the
// threads == p (mod s), so among equal maxima the thread with the smallest bit-reversed id wins.
// The owning thread of candidate k is k & (bs-1), which makes the rule a pure function of (v, i):
// any reduction order gives the reference's answer.
__device__ __forceinline__ Cand better(Cand a, Cand b, unsigned mask) {
    if (b.v > a.v) return b;
    if (b.v == a.v && __brev((unsigned)b.i & mask) < __brev((unsigned)a.i & mask)) return b;
    return a;
}

__device__ __forceinline__ Cand warp_argmax(Cand c, unsigned mask) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        Cand o;
        o.v = __shfl_xor_sync(0xffffffffu, c.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, c.i, off);
        c = better(c, o, mask);
    }
    return c;   // valid in every lane
}

// PPT: points per thread (their min-distances live in registers).  SMEM_XYZ: xyz staged in shared memory.
// The coordinates of the first REG of a thread's points are kept in registers as well: at PPT = 16 a round reads
// 3 x 16,384 floats from shared memory (196 KB = ~1,500 cycles of shared-memory bandwidth, the bulk of a round);
// holding half of them in the register file (64 registers per thread at 1024 threads) halves that.
template <int PPT, bool SMEM_XYZ>
__global__ void __launch_bounds__(1024)
fps_kernel(int n, int m, int bs, const float* __restrict__ dataset, float* __restrict__ temp, int* __restrict__ idxs) {
    extern __shared__ float s_dyn[];
    __shared__ float s_v[2][32];
    __shared__ int s_i[2][32];
    if (m <= 0) return;
    // bs = the reference's logical block size (point->thread map, tie rule); blockDim >= 32
    const int tid = threadIdx.x;
    const int nwarps = (bs + 31) >> 5;
    dataset += (size_t)blockIdx.x * n * 3;
    idxs += (size_t)blockIdx.x * m;
    if (temp) temp += (size_t)blockIdx.x * n;

    float* sx = s_dyn; float* sy = s_dyn + n; float* sz = s_dyn + 2 * n;
    if (SMEM_XYZ) {
        for (int f = tid; f < 3 * n; f += blockDim.x) {
            const float v = dataset[f];
            const int k = f / 3, c = f - 3 * k;
            s_dyn[c * n + k] = v;
        }
        __syncthreads();
    }
    constexpr int REG = (!SMEM_XYZ || PPT <= 4) ? PPT : (PPT <= 8 ? 6 : 8);
    float dist[PPT];
    float px[REG], py[REG], pz[REG];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        dist[j] = 1e10f;   // pointnet2_utils.py:26
        const int k = tid + j * bs;
        if (j < REG) {
            const bool ok = k < n;
            px[j] = ok ? (SMEM_XYZ ? sx[k] : dataset[k * 3 + 0]) : 0.f;
            py[j] = ok ? (SMEM_XYZ ? sy[k] : dataset[k * 3 + 1]) : 0.f;
            pz[j] = ok ? (SMEM_XYZ ? sz[k] : dataset[k * 3 + 2]) : 0.f;
        }
    }
    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int r = 1; r < m; ++r) {
        float x1, y1, z1;
        if (SMEM_XYZ) { x1 = sx[old]; y1 = sy[old]; z1 = sz[old]; }
        else { x1 = dataset[old * 3 + 0]; y1 = dataset[old * 3 + 1]; z1 = dataset[old * 3 + 2]; }
        Cand c; c.v = tid < bs ? -1.f : -2.f; c.i = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = tid + j * bs;
            if (k < n && tid < bs) {
                float x2, y2, z2;
                if (j < REG) { x2 = px[j]; y2 = py[j]; z2 = pz[j]; }
                else { x2 = sx[k]; y2 = sy[k]; z2 = sz[k]; }
                const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const float d2 = fminf(d, dist[j]);
                dist[j] = d2;
                if (d2 > c.v) { c.v = d2; c.i = k; }
            }
        }
        c = warp_argmax(c, (unsigned)bs - 1u);
        const int buf = r & 1;
        if (lane_id() == 0) { s_v[buf][warp_id()] = c.v; s_i[buf][warp_id()] = c.i; }
        __syncthreads();
        Cand w; w.v = -2.f; w.i = 0;
        if (lane_id() < nwarps) { w.v = s_v[buf][lane_id()]; w.i = s_i[buf][lane_id()]; }
        w = warp_argmax(w, (unsigned)bs - 1u);
        old = w.i;
        if (tid == 0) idxs[r] = old;
    }
    if (temp) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = tid + j * bs;
            if (k < n && tid < bs) temp[k] = dist[j];
        }
    }
}

// Fast path for n == PPT * bs exactly (every thread owns PPT valid points: 16,384 -> PPT 16, 2,048 -> 2, 1,024 -> 1).
// The kernel above spends ~35 instructions per point and round on bounds predicates, index tracking and the tie rule of
// the reference's tree; here a round is, per point, 3 subtractions, 3 multiply-adds, one min and one max:
//  * the thread keeps only the running MAX of its updated distances; the point that holds it (the lowest j on ties, as a
//    strict '>' scan finds it) is looked up afterwards among PPT registers;
//  * the block argmax reduces ONE 64-bit key per thread: distance bits (non-negative floats order like unsigned ints) in
//    the high word, ~bit_reverse(thread id) in the top bits of the low word — on equal distances the thread with the
//    smallest bit-reversed id wins, which is exactly what the reference's shared-memory tree leaves (better() above) —
//    and j in its low four bits, from which the winner's index tid + j * bs is rebuilt.
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, off);
        k = o > k ? o : k;
    }
    return k;
}

template <int PPT>
__global__ void __launch_bounds__(1024)
fps_kernel_full(int n, int m, const float* __restrict__ dataset, float* __restrict__ temp, int* __restrict__ idxs) {
    extern __shared__ float s_dyn[];
    __shared__ unsigned long long s_k[2][32];
    if (m <= 0) return;
    const int tid = threadIdx.x, bs = blockDim.x;
    const int nwarps = bs >> 5;
    dataset += (size_t)blockIdx.x * n * 3;
    idxs += (size_t)blockIdx.x * m;
    if (temp) temp += (size_t)blockIdx.x * n;
    float* sx = s_dyn; float* sy = s_dyn + n; float* sz = s_dyn + 2 * n;
    for (int f = tid; f < 3 * n; f += bs) {
        const float v = dataset[f];
        const int k = f / 3, c = f - 3 * k;
        s_dyn[c * n + k] = v;
    }
    __syncthreads();
    constexpr int REG = PPT <= 4 ? PPT : (PPT <= 8 ? 6 : 8);     // coordinates held in registers; the rest from shared memory
    float dist[PPT], px[REG], py[REG], pz[REG];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        dist[j] = 1e10f;   // pointnet2_utils.py:26
        if (j < REG) { px[j] = sx[tid + j * bs]; py[j] = sy[tid + j * bs]; pz[j] = sz[tid + j * bs]; }
    }
    const float* mx = sx + tid; const float* my = sy + tid; const float* mz = sz + tid;
    const unsigned low_tid = ~__brev((unsigned)tid) & 0xffc00000u;
    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int r = 1; r < m; ++r) {
        const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
        float best = -1.f;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            float x2, y2, z2;
            if (j < REG) { x2 = px[j]; y2 = py[j]; z2 = pz[j]; }
            else { x2 = mx[j * bs]; y2 = my[j * bs]; z2 = mz[j * bs]; }
            const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float d2 = fminf(d, dist[j]);
            dist[j] = d2;
            best = fmaxf(best, d2);
        }
        unsigned jb = 0;
#pragma unroll
        for (int j = PPT - 1; j >= 0; --j) jb = dist[j] == best ? (unsigned)j : jb;     // lowest j holding the maximum
        unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (low_tid | jb);
        key = warp_max_u64(key);
        const int buf = r & 1;
        if ((tid & 31) == 0) s_k[buf][tid >> 5] = key;
        __syncthreads();
        unsigned long long w = (tid & 31) < nwarps ? s_k[buf][tid & 31] : 0ull;
        w = warp_max_u64(w);
        const unsigned low = (unsigned)w;
        old = (int)__brev(~low & 0xffc00000u) + (int)(low & 0xfu) * bs;
        if (tid == 0) idxs[r] = old;
    }
    if (temp) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) temp[tid + j * bs] = dist[j];
    }
}

// Any N: min-distance array in global `temp` (L2 resident), xyz from global.
__global__ void __launch_bounds__(1024)
fps_kernel_global(int n, int m, int bs, const float* __restrict__ dataset, float* __restrict__ temp, int* __restrict__ idxs) {
    __shared__ float s_v[2][32];
    __shared__ int s_i[2][32];
    if (m <= 0) return;
    const int tid = threadIdx.x;
    const int nwarps = (bs + 31) >> 5;
    dataset += (size_t)blockIdx.x * n * 3;
    idxs += (size_t)blockIdx.x * m;
    temp += (size_t)blockIdx.x * n;
    for (int k = tid; k < n; k += blockDim.x) temp[k] = 1e10f;
    __syncthreads();
    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int r = 1; r < m; ++r) {
        const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
        Cand c; c.v = tid < bs ? -1.f : -2.f; c.i = 0;
        for (int k = tid; k < n && tid < bs; k += bs) {
            const float dx = __fsub_rn(dataset[k * 3 + 0], x1), dy = __fsub_rn(dataset[k * 3 + 1], y1),
                        dz = __fsub_rn(dataset[k * 3 + 2], z1);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > c.v) { c.v = d2; c.i = k; }
        }
        c = warp_argmax(c, (unsigned)bs - 1u);
        const int buf = r & 1;
        if (lane_id() == 0) { s_v[buf][warp_id()] = c.v; s_i[buf][warp_id()] = c.i; }
        __syncthreads();
        Cand w; w.v = -2.f; w.i = 0;
        if (lane_id() < nwarps) { w.v = s_v[buf][lane_id()]; w.i = s_i[buf][lane_id()]; }
        w = warp_argmax(w, (unsigned)bs - 1u);
        old = w.i;
        if (tid == 0) idxs[r] = old;
    }
}

template <int PPT>
int launch_fps(int b, int n, int m, int bs, const float* dataset, float* temp, int* idxs, cudaStream_t st) {
    const size_t smem = (size_t)3 * n * sizeof(float);
    if (n == PPT * bs && bs >= 32) {       // every thread owns exactly PPT points: the lean kernel
        auto fast = fps_kernel_full<PPT>;
        if (smem > 40 * 1024) SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SEEVCN_PROF("fps", st);
        fast<<<b, bs, smem, st>>>(n, m, dataset, temp, idxs);
        SEEVCN_LAUNCH_CHECK();
        return SEEVCN_OK;
    }
    auto kern = fps_kernel<PPT, true>;
    if (smem > 40 * 1024)   // dynamic + 512 B static must stay under the 48 KB default
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SEEVCN_PROF("fps", st);
    kern<<<b, bs < 32 ? 32 : bs, smem, st>>>(n, m, bs, dataset, temp, idxs);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

}  // namespace

extern "C" int seevcn_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp, int* idxs,
                                              seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n >= 0 && m >= 0, "fps: negative size");
    if (b == 0 || m == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(n >= 1, "fps: n must be >= 1 when m > 0");
    SEEVCN_REQUIRE(dataset && idxs, "fps: null pointer");
    const int bs = ref_block_size(n);
    const int ppt = div_up(n, bs);
    cudaStream_t st = as_stream(stream);
    const size_t smem = (size_t)3 * n * sizeof(float);
    if (ppt <= 16 && smem <= 226 * 1024) {
        if (ppt <= 1) return launch_fps<1>(b, n, m, bs, dataset, temp, idxs, st);
        if (ppt <= 2) return launch_fps<2>(b, n, m, bs, dataset, temp, idxs, st);
        if (ppt <= 4) return launch_fps<4>(b, n, m, bs, dataset, temp, idxs, st);
        if (ppt <= 8) return launch_fps<8>(b, n, m, bs, dataset, temp, idxs, st);
        return launch_fps<16>(b, n, m, bs, dataset, temp, idxs, st);
    }
    SEEVCN_REQUIRE(temp != nullptr, "fps: n=%d needs the temp (B,N) scratch buffer", n);
    SEEVCN_PROF("fps", st);
    fps_kernel_global<<<b, bs < 32 ? 32 : bs, 0, st>>>(n, m, bs, dataset, temp, idxs);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
