// Stage 4b — k nearest neighbours and the kNN surface selection built on it.
//
// Replaces the host-side scipy cKDTree loop of
// see/surface_completion/models/vcn/utils/sampling.py:8-41 (partial_with_KDTree) and its
// torch twin :43-67 (dist.topk(k, largest=False)); batch driver :69-80.
//
// Brute force on chip: reference points are staged in shared memory (SoA), every thread
// owns one query and keeps its k best (ascending, ties -> lower index first) in registers.
// The surface-select kernel runs one CTA per object, de-duplicates the queries with a
// shared-memory hash set (the union of neighbour sets is unchanged by duplicate queries —
// the reference's np.unique(partial) at sampling.py:31 is the same optimisation), marks the
// union in a bitmask, and emits complete[sorted(S)] cyclically — no host round trip.
#include "common.cuh"

namespace {

template <int KMAX>
struct TopK {
    float d[KMAX];
    int i[KMAX];
    // The k live entries occupy slots [KMAX-k, KMAX) so that every register index is static;
    // slots below hold -inf sentinels that nothing can displace.
    __device__ __forceinline__ void init(int k) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            d[j] = __int_as_float(j < KMAX - k ? 0xff800000 : 0x7f800000);
            i[j] = -1;
        }
    }
    // keeps ascending order; an equal distance goes AFTER existing entries (lower index first)
    __device__ __forceinline__ void push(float cd, int ci) {
        if (cd < d[KMAX - 1]) {
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                if (cd < d[j]) {
                    const float td = d[j]; d[j] = cd; cd = td;
                    const int ti = i[j]; i[j] = ci; ci = ti;
                }
            }
        }
    }
};

__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float rx, float ry, float rz) {
    const float dx = __fsub_rn(rx, qx), dy = __fsub_rn(ry, qy), dz = __fsub_rn(rz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

template <int KMAX>
__device__ __forceinline__ void scan_refs(TopK<KMAX>& tk, float qx, float qy, float qz,
                                          const float* sx, const float* sy, const float* sz, int cnt, int base) {
    for (int r = 0; r < cnt; ++r) {
        const float d = sqdist(qx, qy, qz, sx[r], sy[r], sz[r]);
        tk.push(d, base + r);
    }
}

constexpr int kKnnThreads = 128;
constexpr int kRefTile = 2048;

// grid (ceil(Q/128), B)
template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(int r, int q, int k, const float* __restrict__ ref_pts, const float* __restrict__ query,
           float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ float sx[kRefTile], sy[kRefTile], sz[kRefTile];
    const int b = blockIdx.y;
    const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
    const float* rp = ref_pts + (size_t)b * r * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (qi < q) {
        const float* qp = query + ((size_t)b * q + qi) * 3;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
    }
    TopK<KMAX> tk; tk.init(k);
    for (int r0 = 0; r0 < r; r0 += kRefTile) {
        const int cnt = min(kRefTile, r - r0);
        __syncthreads();
        for (int f = threadIdx.x; f < cnt * 3; f += kKnnThreads) {
            const float v = rp[(size_t)r0 * 3 + f];
            const int p = f / 3, c = f - 3 * p;
            (c == 0 ? sx : c == 1 ? sy : sz)[p] = v;
        }
        __syncthreads();
        if (qi < q) scan_refs<KMAX>(tk, qx, qy, qz, sx, sy, sz, cnt, r0);
    }
    if (qi < q) {
        int* io = idx + ((size_t)b * q + qi) * k;
        float* dop = dist ? dist + ((size_t)b * q + qi) * k : nullptr;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            const int o = j - (KMAX - k);
            if (o >= 0) {
                io[o] = tk.i[j];
                if (dop) dop[o] = sqrtf(tk.d[j]);
            }
        }
    }
}

__device__ __forceinline__ unsigned hash3(float x, float y, float z) {
    unsigned h = __float_as_uint(x) * 0x9E3779B1u;
    h ^= __float_as_uint(y) * 0x85EBCA77u + (h << 6) + (h >> 2);
    h ^= __float_as_uint(z) * 0xC2B2AE3Du + (h << 6) + (h >> 2);
    return h ^ (h >> 15);
}

constexpr int kSelThreads = 256;

// One CTA per object.  Dynamic smem: complete SoA (3R floats) | partial SoA (3Np floats) |
// hash table (H ints) | bitmask (R/32 words) | word prefix (R/32 + 1 ints)
template <int KMAX>
__global__ void __launch_bounds__(kSelThreads)
knn_surface_select_kernel(int np, int r, int k, int surface_pts, int hash_size,
                          const float* __restrict__ partial, const float* __restrict__ complete,
                          float* __restrict__ out, int* __restrict__ sel_count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int nwords = (r + 31) >> 5;
    float* cx = reinterpret_cast<float*>(s_raw);
    float* cy = cx + r; float* cz = cy + r;
    float* qx = cz + r; float* qy = qx + np; float* qz = qy + np;
    int* tab = reinterpret_cast<int*>(qz + np);
    unsigned* mask = reinterpret_cast<unsigned*>(tab + hash_size);
    int* prefix = reinterpret_cast<int*>(mask + nwords);
    __shared__ int s_total;

    const int b = blockIdx.x;
    const float* cp = complete + (size_t)b * r * 3;
    const float* pp = partial + (size_t)b * np * 3;
    for (int f = threadIdx.x; f < r * 3; f += kSelThreads) {
        const float v = cp[f]; const int p = f / 3, c = f - 3 * p;
        (c == 0 ? cx : c == 1 ? cy : cz)[p] = v;
    }
    for (int f = threadIdx.x; f < np * 3; f += kSelThreads) {
        const float v = pp[f]; const int p = f / 3, c = f - 3 * p;
        (c == 0 ? qx : c == 1 ? qy : qz)[p] = v;
    }
    for (int i = threadIdx.x; i < hash_size; i += kSelThreads) tab[i] = -1;
    for (int i = threadIdx.x; i < nwords; i += kSelThreads) mask[i] = 0u;
    __syncthreads();

    for (int qi = threadIdx.x; qi < np; qi += kSelThreads) {
        const float x = qx[qi], y = qy[qi], z = qz[qi];
        // hash-set insert: the first thread to claim a slot for these exact coordinates runs the query
        unsigned h = hash3(x, y, z) & (hash_size - 1);
        bool unique = false;
        while (true) {
            const int prev = atomicCAS(&tab[h], -1, qi);
            if (prev == -1) { unique = true; break; }
            if (qx[prev] == x && qy[prev] == y && qz[prev] == z) break;   // duplicate query
            h = (h + 1) & (hash_size - 1);
        }
        if (!unique) continue;
        TopK<KMAX> tk; tk.init(k);
        scan_refs<KMAX>(tk, x, y, z, cx, cy, cz, r, 0);
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (tk.i[j] >= 0) atomicOr(&mask[tk.i[j] >> 5], 1u << (tk.i[j] & 31));
    }
    __syncthreads();
    // exclusive prefix of popcounts over mask words (nwords <= 512): one warp, serial chunks
    if (threadIdx.x < 32) {
        int run = 0;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
            const int w = w0 + threadIdx.x;
            const int c = w < nwords ? __popc(mask[w]) : 0;
            int inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, off);
                if ((int)threadIdx.x >= off) inc += t;
            }
            if (w < nwords) prefix[w] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) { prefix[nwords] = run; s_total = run; sel_count[b] = run; }
    }
    __syncthreads();
    const int total = s_total;
    float* o = out + (size_t)b * surface_pts * 3;
    for (int j = threadIdx.x; j < surface_pts; j += kSelThreads) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (total > 0) {
            const int rank = j % total;
            int lo = 0, hi = nwords;           // last w with prefix[w] <= rank
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= rank) lo = mid; else hi = mid; }
            const int bit = __fns(mask[lo], 0, rank - prefix[lo] + 1);
            const int src = (lo << 5) + bit;
            x = cx[src]; y = cy[src]; z = cz[src];
        }
        o[j * 3 + 0] = x; o[j * 3 + 1] = y; o[j * 3 + 2] = z;
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

extern "C" int seevcn_knn(int b, int r, int q, int k, const float* ref_pts, const float* query, float* dist,
                          int* idx, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && r >= 0 && q >= 0, "knn: negative size");
    SEEVCN_REQUIRE(k >= 1 && k <= 64, "knn: k=%d outside [1,64]", k);
    if (b == 0 || q == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(k <= r, "knn: k=%d > number of reference points %d", k, r);
    SEEVCN_REQUIRE(ref_pts && query && idx, "knn: null pointer");
    SEEVCN_REQUIRE(b <= 65535, "knn: b > 65535");
    dim3 grid(div_up(q, kKnnThreads), b);
    cudaStream_t st = as_stream(stream);
    if (k <= 16) knn_kernel<16><<<grid, kKnnThreads, 0, st>>>(r, q, k, ref_pts, query, dist, idx);
    else if (k <= 32) knn_kernel<32><<<grid, kKnnThreads, 0, st>>>(r, q, k, ref_pts, query, dist, idx);
    else knn_kernel<64><<<grid, kKnnThreads, 0, st>>>(r, q, k, ref_pts, query, dist, idx);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

template <int KMAX>
static int launch_select(int b, int np, int r, int k, int surface_pts, const float* partial, const float* complete,
                         float* out, int* sel_count, cudaStream_t st) {
    const int hash_size = next_pow2(2 * np);
    const int nwords = (r + 31) / 32;
    const size_t smem = (size_t)(3 * r + 3 * np) * 4 + (size_t)hash_size * 4 + (size_t)nwords * 4 + (size_t)(nwords + 1) * 4;
    SEEVCN_REQUIRE(smem <= 227 * 1024, "knn_surface_select: r=%d np=%d needs %zu B of shared memory (> 227 KB)", r, np, smem);
    auto kern = knn_surface_select_kernel<KMAX>;
    if (smem > 40 * 1024)
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<b, kSelThreads, smem, st>>>(np, r, k, surface_pts, hash_size, partial, complete, out, sel_count);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_knn_surface_select(int b, int n_partial, int r, int k, int surface_pts, const float* partial,
                                         const float* complete, float* out, int* sel_count,
                                         seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n_partial >= 0 && r >= 0 && surface_pts >= 0, "knn_surface_select: negative size");
    SEEVCN_REQUIRE(k >= 1 && k <= 64, "knn_surface_select: k=%d outside [1,64]", k);
    if (b == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(k <= r, "knn_surface_select: k=%d > r=%d", k, r);
    SEEVCN_REQUIRE(partial && complete && out && sel_count, "knn_surface_select: null pointer");
    cudaStream_t st = as_stream(stream);
    if (k <= 16) return launch_select<16>(b, n_partial, r, k, surface_pts, partial, complete, out, sel_count, st);
    if (k <= 32) return launch_select<32>(b, n_partial, r, k, surface_pts, partial, complete, out, sel_count, st);
    return launch_select<64>(b, n_partial, r, k, surface_pts, partial, complete, out, sel_count, st);
}
