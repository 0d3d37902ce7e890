// Stage 4b — k nearest neighbours and the kNN surface selection built on it.
//
// Replaces the host-side scipy cKDTree loop of
// see/surface_completion/models/vcn/utils/sampling.py:8-41 (partial_with_KDTree) and its
// torch twin :43-67 (dist.topk(k, largest=False)); batch driver :69-80.
//
// Brute force on chip.  Reference points are staged in shared memory as float4; every thread owns one
// query.  The k best of a thread live in shared memory as 64-bit keys (distance bits << 32 | index:
// non-negative float bits order like unsigned ints, so one u64 compare orders by distance and then by
// LOWER index, the tie order of a stable scan) in an UNSORTED list with a tracked maximum ("replace the
// max").  To keep warps convergent, candidates that beat the current k-th distance are first appended
// to a small per-thread buffer; the warp merges buffers into lists together, when any lane's buffer
// runs full.  The fast path per (query, reference) pair is 1 LDS.128 + 6 FP32 ops + compare.
//
// The surface-select kernel runs one CTA per object: it de-duplicates the queries with a shared-memory
// hash set (duplicate queries cannot change a union — the reference's np.unique(partial) at
// sampling.py:31 is the same optimisation; resampled clouds are mostly duplicates), scans only the
// unique ones, marks the union of neighbour sets in a bitmask and emits complete[sorted(S)] cyclically.
#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr u64 kInfKey = ~0ull;
constexpr int kBuf = 16;   // per-thread candidate buffer (flushed when more than 8 are pending)

__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float4 r) {
    const float dx = __fsub_rn(r.x, qx), dy = __fsub_rn(r.y, qy), dz = __fsub_rn(r.z, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Per-thread top-k state.  list / buf are this thread's columns of [slot][T] arrays in shared memory.
template <int T>
struct TopK {
    u64* list; u64* buf;
    u64 thr; float thr_f; int maxpos, nbuf, k;
    __device__ __forceinline__ void init(u64* list_, u64* buf_, int k_) {
        list = list_; buf = buf_; k = k_;
        for (int j = 0; j < k; ++j) list[j * T] = kInfKey;
        thr = kInfKey; thr_f = __int_as_float(0x7f800000); maxpos = 0; nbuf = 0;
    }
    __device__ __forceinline__ void offer(float d, int idx) {
        if (d < thr_f) { buf[nbuf * T] = ((u64)__float_as_uint(d) << 32) | (unsigned)idx; ++nbuf; }
    }
    __device__ __forceinline__ void flush() {
        for (int e = 0; e < nbuf; ++e) {
            const u64 key = buf[e * T];
            if (key < thr) {
                list[maxpos * T] = key;
                u64 m = 0; int mp = 0;
                for (int j = 0; j < k; ++j) { const u64 v = list[j * T]; if (v > m) { m = v; mp = j; } }
                thr = m; maxpos = mp;
                thr_f = thr == kInfKey ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(thr >> 32));
            }
        }
        nbuf = 0;
    }
};

// Scan `cnt` staged references (global indices base..base+cnt) for this thread's query.  Whole warps call
// this together; `active` lanes own a query.
template <int T>
__device__ __forceinline__ void scan_refs(TopK<T>& tk, bool active, float qx, float qy, float qz, const float4* refs,
                                          int cnt, int base) {
    for (int r0 = 0; r0 < cnt; r0 += 8) {
        if (active) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u;
                if (r < cnt) tk.offer(sqdist(qx, qy, qz, refs[r]), base + r);
            }
        }
        if (__any_sync(0xffffffffu, tk.nbuf > kBuf - 8)) tk.flush();
    }
    tk.flush();
}

constexpr int kKnnThreads = 128;
constexpr int kRefTile = 1024;

// grid (ceil(Q/128), B); dynamic smem: list (k x T u64) | buf (kBuf x T u64)
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(int r, int q, int k, const float* __restrict__ ref_pts, const float* __restrict__ query,
           float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ float4 s_ref[kRefTile];
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* s_list = reinterpret_cast<u64*>(s_raw);
    u64* s_buf = s_list + (size_t)k * kKnnThreads;
    const int b = blockIdx.y;
    const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
    const float* rp = ref_pts + (size_t)b * r * 3;
    const bool active = qi < q;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float* qp = query + ((size_t)b * q + qi) * 3;
        qx = qp[0]; qy = qp[1]; qz = qp[2];
    }
    TopK<kKnnThreads> tk;
    tk.init(s_list + threadIdx.x, s_buf + threadIdx.x, k);
    for (int r0 = 0; r0 < r; r0 += kRefTile) {
        const int cnt = min(kRefTile, r - r0);
        __syncthreads();
        for (int p = threadIdx.x; p < cnt; p += kKnnThreads) {
            const float* s = rp + (size_t)(r0 + p) * 3;
            s_ref[p] = make_float4(s[0], s[1], s[2], 0.f);
        }
        __syncthreads();
        scan_refs<kKnnThreads>(tk, active, qx, qy, qz, s_ref, cnt, r0);
    }
    if (active) {
        // ascending order: insertion sort of this thread's k keys
        for (int a = 1; a < k; ++a) {
            const u64 key = tk.list[a * kKnnThreads];
            int j = a - 1;
            while (j >= 0 && tk.list[j * kKnnThreads] > key) { tk.list[(j + 1) * kKnnThreads] = tk.list[j * kKnnThreads]; --j; }
            tk.list[(j + 1) * kKnnThreads] = key;
        }
        int* io = idx + ((size_t)b * q + qi) * k;
        float* dop = dist ? dist + ((size_t)b * q + qi) * k : nullptr;
        for (int j = 0; j < k; ++j) {
            const u64 key = tk.list[j * kKnnThreads];
            io[j] = (int)(unsigned)key;
            if (dop) dop[j] = sqrtf(__uint_as_float((unsigned)(key >> 32)));
        }
    }
}

__device__ __forceinline__ unsigned hash3(float x, float y, float z) {
    unsigned h = __float_as_uint(x) * 0x9E3779B1u;
    h ^= __float_as_uint(y) * 0x85EBCA77u + (h << 6) + (h >> 2);
    h ^= __float_as_uint(z) * 0xC2B2AE3Du + (h << 6) + (h >> 2);
    return h ^ (h >> 15);
}

// Surface selection: grid (chunks, objects), T threads per CTA.  Every CTA of an object stages the completed
// cloud, de-duplicates the object's queries DETERMINISTICALLY (representative = lowest index of equal
// coordinates, unique list in index order) and takes the chunk-th slice of T unique queries: objects with
// 1024 distinct queries spread over 4 CTAs, objects with few are done in one and the others exit at once —
// without this split the kernel time is set by the largest object while half the SMs idle.  Neighbour sets
// are OR-ed into a global bit mask; the last CTA of an object to arrive (global counter) emits the output.
// Dynamic smem:
//   complete float4[R] | list k*T u64 | buf kBuf*T u64 | partial SoA 3*Np f32 | hash table H i32 |
//   unique list Np i32 | bitmask nw u32 | prefix (nw+1) i32
template <int T>
__global__ void __launch_bounds__(T)
knn_surface_select_kernel(int np, int r, int k, int surface_pts, int hash_size,
                          const float* __restrict__ partial, const float* __restrict__ complete,
                          float* __restrict__ out, int* __restrict__ sel_count,
                          unsigned* __restrict__ g_mask, unsigned* __restrict__ g_arrive) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int nwords = (r + 31) >> 5;
    float4* cref = reinterpret_cast<float4*>(s_raw);
    u64* s_list = reinterpret_cast<u64*>(cref + r);
    u64* s_buf = s_list + (size_t)k * T;
    float* qx = reinterpret_cast<float*>(s_buf + (size_t)kBuf * T);
    float* qy = qx + np; float* qz = qy + np;
    int* tab = reinterpret_cast<int*>(qz + np);
    int* uq = tab + hash_size;
    unsigned* mask = reinterpret_cast<unsigned*>(uq + np);
    int* prefix = reinterpret_cast<int*>(mask + nwords);
    __shared__ int s_total, s_run, s_last;
    __shared__ int s_warp_cnt[T / 32];

    const int b = blockIdx.y, chunk = blockIdx.x;
    const float* cp = complete + (size_t)b * r * 3;
    const float* pp = partial + (size_t)b * np * 3;
    for (int p = threadIdx.x; p < r; p += T) cref[p] = make_float4(cp[p * 3 + 0], cp[p * 3 + 1], cp[p * 3 + 2], 0.f);
    for (int f = threadIdx.x; f < np * 3; f += T) {
        const float v = pp[f]; const int p = f / 3, c = f - 3 * p;
        (c == 0 ? qx : c == 1 ? qy : qz)[p] = v;
    }
    for (int i = threadIdx.x; i < hash_size; i += T) tab[i] = -1;
    for (int i = threadIdx.x; i < nwords; i += T) mask[i] = 0u;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();

    // hash-set insert keyed by the exact coordinates; a slot ends up holding the LOWEST query index of its key
    for (int qi = threadIdx.x; qi < np; qi += T) {
        const float x = qx[qi], y = qy[qi], z = qz[qi];
        unsigned h = hash3(x, y, z) & (hash_size - 1);
        while (true) {
            const int prev = atomicCAS(&tab[h], -1, qi);
            if (prev == -1) break;
            if (qx[prev] == x && qy[prev] == y && qz[prev] == z) { atomicMin(&tab[h], qi); break; }   // duplicate query
            h = (h + 1) & (hash_size - 1);
        }
    }
    __syncthreads();
    // unique list in ascending index order (block scan over the representative flags)
    for (int q0 = 0; q0 < np; q0 += T) {
        const int qi = q0 + threadIdx.x;
        bool rep = false;
        if (qi < np) {
            const float x = qx[qi], y = qy[qi], z = qz[qi];
            unsigned h = hash3(x, y, z) & (hash_size - 1);
            while (true) {
                const int cur = tab[h];
                if (qx[cur] == x && qy[cur] == y && qz[cur] == z) { rep = cur == qi; break; }
                h = (h + 1) & (hash_size - 1);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, rep);
        if (lane_id() == 0) s_warp_cnt[warp_id()] = __popc(bal);
        __syncthreads();
        if (warp_id() == 0) {
            const int c = lane_id() < T / 32 ? s_warp_cnt[lane_id()] : 0;
            int inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane_id() >= off) inc += t;
            }
            if (lane_id() < T / 32) s_warp_cnt[lane_id()] = s_run + inc - c;
            __syncwarp();
            if (lane_id() == 31) s_run += inc;
        }
        __syncthreads();
        if (rep) uq[s_warp_cnt[warp_id()] + __popc(bal & ((1u << lane_id()) - 1))] = qi;
        __syncthreads();
    }
    const int nuniq = s_run;

    // this CTA's slice of the unique queries
    const int u = chunk * T + threadIdx.x;
    if (chunk * T + (int)(threadIdx.x & ~31u) < nuniq) {            // warp-uniform
        const bool active = u < nuniq;
        float x = 0.f, y = 0.f, z = 0.f;
        if (active) { const int qi = uq[u]; x = qx[qi]; y = qy[qi]; z = qz[qi]; }
        TopK<T> tk;
        tk.init(s_list + threadIdx.x, s_buf + threadIdx.x, k);
        scan_refs<T>(tk, active, x, y, z, cref, r, 0);
        if (active)
            for (int j = 0; j < k; ++j) {
                const u64 key = tk.list[j * T];
                if (key != kInfKey) { const unsigned i = (unsigned)key; atomicOr(&mask[i >> 5], 1u << (i & 31)); }
            }
    }
    __syncthreads();
    unsigned* gm = g_mask + (size_t)b * nwords;
    for (int w = threadIdx.x; w < nwords; w += T)
        if (mask[w]) atomicOr(&gm[w], mask[w]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&g_arrive[b], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int w = threadIdx.x; w < nwords; w += T) mask[w] = __ldcg(&gm[w]);
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
    for (int f = threadIdx.x; f < surface_pts * 3; f += T) {
        float v = 0.f;
        if (total > 0) {
            const int j = f / 3, c = f - 3 * j;
            const int rank = j % total;
            int lo = 0, hi = nwords;           // last w with prefix[w] <= rank
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= rank) lo = mid; else hi = mid; }
            const int bit = __fns(mask[lo], 0, rank - prefix[lo] + 1);
            const float4 s = cref[(lo << 5) + bit];
            v = c == 0 ? s.x : c == 1 ? s.y : s.z;
        }
        o[f] = v;
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

size_t select_smem(int T, int np, int r, int k, int hash_size) {
    const int nwords = (r + 31) / 32;
    return (size_t)r * 16 + (size_t)(k + kBuf) * T * 8 + (size_t)np * 12 + (size_t)hash_size * 4 + (size_t)np * 4 +
           (size_t)nwords * 4 + (size_t)(nwords + 1) * 4 + 16;
}

template <int T>
int launch_select(int b, int np, int r, int k, int surface_pts, int hash_size, size_t smem, const float* partial,
                  const float* complete, float* out, int* sel_count, unsigned* g_mask, unsigned* g_arrive, cudaStream_t st) {
    auto kern = knn_surface_select_kernel<T>;
    if (smem > 40 * 1024)
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = np > 0 ? div_up(np, T) : 1;
    kern<<<dim3(chunks, b), T, smem, st>>>(np, r, k, surface_pts, hash_size, partial, complete, out, sel_count, g_mask, g_arrive);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

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
    const size_t smem = (size_t)(k + kBuf) * kKnnThreads * 8;
    if (smem > 30 * 1024)
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<<<grid, kKnnThreads, smem, as_stream(stream)>>>(r, q, k, ref_pts, query, dist, idx);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" size_t seevcn_knn_surface_select_workspace_bytes(int b, int r) {
    return ((size_t)(b > 0 ? b : 0) * ((size_t)((r > 0 ? r : 0) + 31) / 32 + 1)) * 4 + 256;
}

extern "C" int seevcn_knn_surface_select(int b, int n_partial, int r, int k, int surface_pts, const float* partial,
                                         const float* complete, float* out, int* sel_count,
                                         void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n_partial >= 0 && r >= 0 && surface_pts >= 0, "knn_surface_select: negative size");
    SEEVCN_REQUIRE(k >= 1 && k <= 64, "knn_surface_select: k=%d outside [1,64]", k);
    if (b == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(k <= r, "knn_surface_select: k=%d > r=%d", k, r);
    SEEVCN_REQUIRE(b <= 65535, "knn_surface_select: b > 65535");
    SEEVCN_REQUIRE(partial && complete && out && sel_count && workspace, "knn_surface_select: null pointer");
    if (workspace_bytes < seevcn_knn_surface_select_workspace_bytes(b, r)) {
        seevcn_set_error("knn_surface_select: workspace too small");
        return SEEVCN_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int nwords = (r + 31) / 32;
    unsigned* g_mask = static_cast<unsigned*>(workspace);
    unsigned* g_arrive = g_mask + (size_t)b * nwords;
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(workspace, 0, ((size_t)b * (nwords + 1)) * 4, st));
    const int hash_size = next_pow2(2 * (n_partial > 0 ? n_partial : 1));
    // widest block whose per-thread lists still fit next to the staged clouds
    const size_t lim = 227 * 1024;
    if (select_smem(256, n_partial, r, k, hash_size) <= lim / 2)   // two CTAs per SM
        return launch_select<256>(b, n_partial, r, k, surface_pts, hash_size, select_smem(256, n_partial, r, k, hash_size),
                                  partial, complete, out, sel_count, g_mask, g_arrive, st);
    if (select_smem(128, n_partial, r, k, hash_size) <= lim)
        return launch_select<128>(b, n_partial, r, k, surface_pts, hash_size, select_smem(128, n_partial, r, k, hash_size),
                                  partial, complete, out, sel_count, g_mask, g_arrive, st);
    const size_t need = select_smem(32, n_partial, r, k, hash_size);
    SEEVCN_REQUIRE(need <= lim, "knn_surface_select: r=%d np=%d k=%d needs %zu B of shared memory (> 227 KB)", r, n_partial, k, need);
    return launch_select<32>(b, n_partial, r, k, surface_pts, hash_size, need, partial, complete, out, sel_count, g_mask, g_arrive, st);
}
