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
        if (d <= thr_f) { buf[nbuf * T] = ((u64)__float_as_uint(d) << 32) | (unsigned)idx; ++nbuf; }
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

// float -> unsigned with the same order (-0 < +0)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float axis_of(float4 v, int a) { return a == 0 ? v.x : a == 1 ? v.y : v.z; }

// ascending bitonic sort of s[0..len), len a power of two, by the whole CTA (ends with a barrier)
template <int T>
__device__ void bitonic_sort(u64* s, int len) {
    for (int k = 2; k <= len; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (len >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const u64 a = s[i], b = s[p];
                if ((a > b) == ((i & k) == 0)) { s[i] = b; s[p] = a; }
            }
            __syncthreads();
        }
}

// Surface selection = union over the object's queries of their k nearest completed points.  Two kernels.
//
// (1) knn_prepare_kernel, one CTA per object: picks the axis along which the completed cloud is longest,
// sorts the completed points along it (bitonic sort of (coordinate, index) keys in shared memory) and writes
// them as float4 (x, y, z, original index) to the workspace; de-duplicates the object's queries with a
// shared-memory hash set keyed by the exact coordinates (duplicate queries cannot change a union — the
// reference's np.unique(partial) at sampling.py:31 is the same optimisation; resampled clouds are mostly
// duplicates), sorts the unique ones along the same axis and writes them out, with their number.
//
// (2) knn_sweep_select_kernel, grid (query chunks, objects): one thread per unique query.  The thread finds
// its position in the sorted cloud by binary search and sweeps outwards on both sides; a side stops when
// the squared coordinate difference alone exceeds the k-th best distance so far (every further point on
// that side is farther still: fl(da*da) <= fl(dx*dx + dy*dy + dz*dz) under round-to-nearest).  A car-sized
// cloud of 1024 points with k = 20 visits ~1/5 of the pairs of the brute-force scan, and because queries
// are sorted along the sweep axis the lanes of a warp walk neighbouring addresses for similar trip counts.
// Neighbour sets are OR-ed into a global bit mask; the last CTA of an object to arrive (global counter)
// emits complete[sorted(S)] cyclically.  The result is exactly the brute-force one (ties: lower index).
//
// prepare smem: keys max(rp2, qp2) u64 | partial SoA 3*np f32 | hash table H i32
template <int T>
__global__ void __launch_bounds__(T)
knn_prepare_kernel(int np, int r, int rp2, int qp2, int hash_size, const float* __restrict__ partial,
                   const float* __restrict__ complete, float4* __restrict__ ws_refs, float4* __restrict__ ws_q,
                   int* __restrict__ ws_meta) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* keys = reinterpret_cast<u64*>(s_raw);
    float* qx = reinterpret_cast<float*>(keys + max(rp2, qp2));
    float* qy = qx + np; float* qz = qy + np;
    int* tab = reinterpret_cast<int*>(qz + np);
    __shared__ float s_lo[3][T / 32], s_hi[3][T / 32];
    __shared__ int s_axis, s_run;

    const int b = blockIdx.x;
    const float* cp = complete + (size_t)b * r * 3;
    const float* pp = partial + (size_t)b * np * 3;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int p = threadIdx.x; p < r; p += T)
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float v = cp[p * 3 + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], off));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], off));
        }
        if (lane_id() == 0) { s_lo[c][warp_id()] = lo[c]; s_hi[c][warp_id()] = hi[c]; }
    }
    for (int f = threadIdx.x; f < np * 3; f += T) {
        const float v = pp[f]; const int p = f / 3, c = f - 3 * p;
        (c == 0 ? qx : c == 1 ? qy : qz)[p] = v;
    }
    for (int i = threadIdx.x; i < hash_size; i += T) tab[i] = -1;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        float ext[3];
        for (int c = 0; c < 3; ++c) {
            float l = inf, h = -inf;
            for (int w = 0; w < T / 32; ++w) { l = fminf(l, s_lo[c][w]); h = fmaxf(h, s_hi[c][w]); }
            ext[c] = h - l;
        }
        int a = 0;
        if (ext[1] > ext[a]) a = 1;
        if (ext[2] > ext[a]) a = 2;
        s_axis = a;
    }
    __syncthreads();
    const int axis = s_axis;

    // completed cloud, sorted along the axis
    for (int p = threadIdx.x; p < rp2; p += T)
        keys[p] = p < r ? ((u64)f2ord(cp[p * 3 + axis]) << 32) | (unsigned)p : kInfKey;
    __syncthreads();
    bitonic_sort<T>(keys, rp2);
    float4* wr = ws_refs + (size_t)b * r;
    for (int p = threadIdx.x; p < r; p += T) {
        const int i = (int)(unsigned)keys[p];
        wr[p] = make_float4(cp[i * 3 + 0], cp[i * 3 + 1], cp[i * 3 + 2], __int_as_float(i));
    }
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
    // representatives (lowest index of equal coordinates) -> sort keys; the sort makes their order deterministic
    for (int qi = threadIdx.x; qi < np; qi += T) {
        const float x = qx[qi], y = qy[qi], z = qz[qi];
        unsigned h = hash3(x, y, z) & (hash_size - 1);
        while (true) {
            const int cur = tab[h];
            if (qx[cur] == x && qy[cur] == y && qz[cur] == z) {
                if (cur == qi) keys[atomicAdd(&s_run, 1)] = ((u64)f2ord(axis == 0 ? x : axis == 1 ? y : z) << 32) | (unsigned)qi;
                break;
            }
            h = (h + 1) & (hash_size - 1);
        }
    }
    __syncthreads();
    const int nuniq = s_run;
    for (int p = nuniq + threadIdx.x; p < qp2; p += T) keys[p] = kInfKey;
    __syncthreads();
    bitonic_sort<T>(keys, qp2);
    float4* wq = ws_q + (size_t)b * np;
    for (int p = threadIdx.x; p < nuniq; p += T) {
        const int qi = (int)(unsigned)keys[p];
        wq[p] = make_float4(qx[qi], qy[qi], qz[qi], 0.f);
    }
    if (threadIdx.x == 0) { ws_meta[2 * b] = nuniq; ws_meta[2 * b + 1] = axis; }
}

// select smem: [sorted cloud float4[R] if kSmemRefs] | list k*T u64 | buf kBuf*T u64 | bitmask nw u32 | prefix (nw+1) i32
template <int T, bool kSmemRefs>
__global__ void __launch_bounds__(T)
knn_sweep_select_kernel(int np, int r, int k, int surface_pts, const float* __restrict__ complete,
                        const float4* __restrict__ ws_refs, const float4* __restrict__ ws_q,
                        const int* __restrict__ ws_meta, float* __restrict__ out, int* __restrict__ sel_count,
                        unsigned* __restrict__ g_mask, unsigned* __restrict__ g_arrive) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int nwords = (r + 31) >> 5;
    float4* cref = reinterpret_cast<float4*>(s_raw);
    u64* s_list = reinterpret_cast<u64*>(cref + (kSmemRefs ? r : 0));
    u64* s_buf = s_list + (size_t)k * T;
    unsigned* mask = reinterpret_cast<unsigned*>(s_buf + (size_t)kBuf * T);
    int* prefix = reinterpret_cast<int*>(mask + nwords);
    __shared__ int s_total, s_last;

    const int b = blockIdx.y, chunk = blockIdx.x;
    const int nuniq = ws_meta[2 * b], axis = ws_meta[2 * b + 1];
    const float4* gref = ws_refs + (size_t)b * r;
    for (int i = threadIdx.x; i < nwords; i += T) mask[i] = 0u;
    if (chunk * T < nuniq) {                                          // CTA-uniform
        const float4* refs = gref;
        if (kSmemRefs) {
            for (int p = threadIdx.x; p < r; p += T) cref[p] = gref[p];
            refs = cref;
        }
        __syncthreads();
        const int u = chunk * T + threadIdx.x;
        if (chunk * T + (int)(threadIdx.x & ~31u) < nuniq) {            // warp-uniform
            const bool active = u < nuniq;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) q = ws_q[(size_t)b * np + u];
            const float qa = axis_of(q, axis);
            int lo = 0, hi = r;                                         // first sorted position with coordinate >= qa
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (axis_of(refs[mid], axis) < qa) lo = mid + 1; else hi = mid;
            }
            int pr = lo, pl = lo - 1;
            bool go_r = active && pr < r, go_l = active && pl >= 0;
            TopK<T> tk;
            tk.init(s_list + threadIdx.x, s_buf + threadIdx.x, k);
            int it = 0;
            while (__any_sync(0xffffffffu, go_r || go_l)) {
                if (go_r) {
                    const float4 c = refs[pr];
                    const float da = __fsub_rn(axis_of(c, axis), qa);
                    if (__fmul_rn(da, da) > tk.thr_f) go_r = false;
                    else { tk.offer(sqdist(q.x, q.y, q.z, c), __float_as_int(c.w)); go_r = ++pr < r; }
                }
                if (go_l) {
                    const float4 c = refs[pl];
                    const float da = __fsub_rn(axis_of(c, axis), qa);
                    if (__fmul_rn(da, da) > tk.thr_f) go_l = false;
                    else { tk.offer(sqdist(q.x, q.y, q.z, c), __float_as_int(c.w)); go_l = --pl >= 0; }
                }
                if ((++it & 3) == 0 && __any_sync(0xffffffffu, tk.nbuf > kBuf - 8)) tk.flush();
            }
            tk.flush();
            if (active)
                for (int j = 0; j < k; ++j) {
                    const u64 key = tk.list[j * T];
                    if (key != kInfKey) { const unsigned i = (unsigned)key; atomicOr(&mask[i >> 5], 1u << (i & 31)); }
                }
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
    const float* cp = complete + (size_t)b * r * 3;
    float* o = out + (size_t)b * surface_pts * 3;
    for (int f = threadIdx.x; f < surface_pts * 3; f += T) {
        float v = 0.f;
        if (total > 0) {
            const int j = f / 3, c = f - 3 * j;
            const int rank = j % total;
            int lo = 0, hi = nwords;           // last w with prefix[w] <= rank
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= rank) lo = mid; else hi = mid; }
            const int bit = __fns(mask[lo], 0, rank - prefix[lo] + 1);
            v = __ldg(&cp[((lo << 5) + bit) * 3 + c]);
        }
        o[f] = v;
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

size_t select_smem(int T, bool smem_refs, int r, int k) {
    const int nwords = (r + 31) / 32;
    return (smem_refs ? (size_t)r * 16 : 0) + (size_t)(k + kBuf) * T * 8 + (size_t)nwords * 4 + (size_t)(nwords + 1) * 4 + 16;
}

template <int T, bool kSmemRefs>
int launch_select(int b, int np, int r, int k, int surface_pts, size_t smem, const float* complete, const float4* ws_refs,
                  const float4* ws_q, const int* ws_meta, float* out, int* sel_count, unsigned* g_mask, unsigned* g_arrive,
                  cudaStream_t st) {
    auto kern = knn_sweep_select_kernel<T, kSmemRefs>;
    if (smem > 40 * 1024)
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = np > 0 ? div_up(np, T) : 1;
    SEEVCN_PROF("knn_sweep_select_kernel", st);
    kern<<<dim3(chunks, b), T, smem, st>>>(np, r, k, surface_pts, complete, ws_refs, ws_q, ws_meta, out, sel_count, g_mask, g_arrive);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

struct SelectWs { size_t mask, arrive, meta, refs, q, total; };
SelectWs select_ws(int b, int np, int r) {
    const size_t B = b > 0 ? b : 0, R = r > 0 ? r : 0, N = np > 0 ? np : 0;
    SelectWs w;
    w.mask = 0;
    w.arrive = w.mask + B * ((R + 31) / 32) * 4;
    w.meta = w.arrive + B * 4;
    w.refs = align_up(w.meta + B * 8, 256);
    w.q = w.refs + B * R * 16;
    w.total = w.q + B * N * 16 + 256;
    return w;
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

extern "C" size_t seevcn_knn_surface_select_workspace_bytes(int b, int n_partial, int r) {
    return select_ws(b, n_partial, r).total;
}

extern "C" int seevcn_knn_surface_select(int b, int n_partial, int r, int k, int surface_pts, const float* partial,
                                         const float* complete, float* out, int* sel_count,
                                         void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n_partial >= 0 && r >= 0 && surface_pts >= 0, "knn_surface_select: negative size");
    SEEVCN_REQUIRE(k >= 1 && k <= 64, "knn_surface_select: k=%d outside [1,64]", k);
    if (b == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(k <= r, "knn_surface_select: k=%d > r=%d", k, r);
    SEEVCN_REQUIRE(b <= 65535, "knn_surface_select: b > 65535");
    SEEVCN_REQUIRE(r <= 16384 && n_partial <= 4096, "knn_surface_select: r=%d > 16384 or n_partial=%d > 4096", r, n_partial);
    SEEVCN_REQUIRE(partial && complete && out && sel_count && workspace, "knn_surface_select: null pointer");
    const SelectWs w = select_ws(b, n_partial, r);
    if (workspace_bytes < w.total) {
        seevcn_set_error("knn_surface_select: workspace too small");
        return SEEVCN_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    unsigned char* base = static_cast<unsigned char*>(workspace);
    SEEVCN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "knn_surface_select: workspace must be 16-byte aligned");
    unsigned* g_mask = reinterpret_cast<unsigned*>(base + w.mask);
    unsigned* g_arrive = reinterpret_cast<unsigned*>(base + w.arrive);
    int* ws_meta = reinterpret_cast<int*>(base + w.meta);
    float4* ws_refs = reinterpret_cast<float4*>(base + w.refs);
    float4* ws_q = reinterpret_cast<float4*>(base + w.q);
    SEEVCN_PROF("knn_surface_select", st);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(base, 0, w.meta, st));
    {
        constexpr int TP = 512;
        const int rp2 = next_pow2(r), qp2 = next_pow2(n_partial > 0 ? n_partial : 1);
        const int hash_size = next_pow2(2 * (n_partial > 0 ? n_partial : 1));
        const size_t smem = (size_t)(rp2 > qp2 ? rp2 : qp2) * 8 + (size_t)n_partial * 12 + (size_t)hash_size * 4;
        auto kern = knn_prepare_kernel<TP>;
        if (smem > 40 * 1024)
            SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SEEVCN_PROF("knn_prepare_kernel", st);
        kern<<<b, TP, smem, st>>>(n_partial, r, rp2, qp2, hash_size, partial, complete, ws_refs, ws_q, ws_meta);
        SEEVCN_LAUNCH_CHECK();
    }
    const size_t lim = 227 * 1024;
    const bool fits = r <= 4096;
    if (fits && select_smem(256, true, r, k) <= lim / 2)   // two CTAs per SM
        return launch_select<256, true>(b, n_partial, r, k, surface_pts, select_smem(256, true, r, k), complete, ws_refs, ws_q,
                                        ws_meta, out, sel_count, g_mask, g_arrive, st);
    if (fits && select_smem(128, true, r, k) <= lim)
        return launch_select<128, true>(b, n_partial, r, k, surface_pts, select_smem(128, true, r, k), complete, ws_refs, ws_q,
                                        ws_meta, out, sel_count, g_mask, g_arrive, st);
    return launch_select<128, false>(b, n_partial, r, k, surface_pts, select_smem(128, false, r, k), complete, ws_refs, ws_q,
                                     ws_meta, out, sel_count, g_mask, g_arrive, st);
}
