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

// Per-thread top-k state.  list / buf are this thread's columns of [slot][T] arrays in shared memory.  list is a 4-ary
// MAX-heap of k keys (root = the k-th best so far, kInfKey while fewer than k were seen): an accepted candidate replaces
// the root and sifts down, <= log4(k) levels (two for k = 20) of four loads and one store.  The first k candidates
// are appended and the heap is built once (Floyd), so filling costs no sifts.
template <int T>
struct TopK {
    u64* list; u64* buf;
    u64 thr; float thr_f; int nbuf, k, count;
    __device__ __forceinline__ void init(u64* list_, u64* buf_, int k_) {
        list = list_; buf = buf_; k = k_;
        for (int j = 0; j < k; ++j) list[j * T] = kInfKey;
        thr = kInfKey; thr_f = __int_as_float(0x7f800000); nbuf = 0; count = 0;
    }
    __device__ __forceinline__ void offer(float d, int idx) {
        if (d <= thr_f) { buf[nbuf * T] = ((u64)__float_as_uint(d) << 32) | (unsigned)idx; ++nbuf; }
    }
    // put key into the hole at node i and let it sink
    __device__ __forceinline__ void sift_down(int i, u64 key) {
        while (true) {
            const int c = 4 * i + 1;
            if (c >= k) break;
            u64 m = list[c * T];
            int mc = c;
#pragma unroll
            for (int t = 1; t < 4; ++t)
                if (c + t < k) { const u64 v = list[(c + t) * T]; if (v > m) { m = v; mc = c + t; } }
            if (m <= key) break;
            list[i * T] = m;
            i = mc;
        }
        list[i * T] = key;
    }
    __device__ __forceinline__ void set_thr() {
        thr = list[0];
        thr_f = thr == kInfKey ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(thr >> 32));
    }
    __device__ __forceinline__ void insert(u64 key) {   // key < thr
        if (count < k) {                                // filling: append, heapify once when the k-th arrives
            list[count * T] = key;
            if (++count == k) {
                for (int i = (k - 2) / 4; i >= 0; --i) sift_down(i, list[i * T]);
                set_thr();
            }
            return;
        }
        sift_down(0, key);
        set_thr();
    }
    __device__ __forceinline__ void flush() {
        for (int e = 0; e < nbuf; ++e) {
            const u64 key = buf[e * T];
            if (key < thr) insert(key);
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

// ascending bitonic sort of s[0..len), len a power of two, by the whole CTA (ends with a barrier).  Pair t of a stage
// with stride j <= 32 lies in the aligned 64-element chunk that the 32 consecutive pairs of one warp cover, so those
// stages (45 of the 55 for 1024 keys) only need a warp barrier; a CTA barrier separates them from the wide stages.
template <int T>
__device__ void bitonic_sort(u64* s, int len) {
    bool wide_prev = true;                       // callers arrive behind a CTA barrier
    for (int k = 2; k <= len; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool wide = j > 32;
            if (wide && !wide_prev) __syncthreads();
            for (int t = threadIdx.x; t < (len >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const u64 a = s[i], b = s[p];
                if ((a > b) == ((i & k) == 0)) { s[i] = b; s[p] = a; }
            }
            if (wide) __syncthreads(); else __syncwarp();
            wide_prev = wide;
        }
    __syncthreads();
}

// Surface selection = union over the object's queries of their k nearest completed points.  Three kernels.
//
// (1) knn_prepare_kernel, one CTA per object: sorts the completed points by the Morton code of their position in
// the cloud's bounding box (bitonic sort of (code, index) keys in shared memory), writes them as float4
// (x, y, z, original index) and, for every block of 32 consecutive sorted points — a compact patch of the cloud —
// its bounding box; sorts the object's queries by Morton code too and drops the copies of a point, which sit in
// one run of equal codes (duplicate queries cannot change a union — the reference's np.unique(partial) at
// sampling.py:31 is the same optimisation; resampled clouds are mostly duplicates), so that the 32 queries of a
// warp are distinct neighbours in space.
//
// (2) knn_scan_kernel, grid (query chunks, objects): one thread per unique query, warps independent (no CTA
// barrier).  A warp walks the blocks in lockstep — every lane reads the same point, one broadcast load per
// point — and skips a block when, for every lane, the squared distance from the query to the block's box
// exceeds the lane's k-th best distance so far.  The box distance is evaluated with the operation order of
// the point distance, so in fp32 it never exceeds the distance of a point inside the box and the result is
// exactly the brute-force one (ties: lower index).  Blocks are visited nearest first: the warp sorts them once by
// the smallest box distance over its 32 queries, so the heaps fill with a tight k-th distance at once and the walk ends at the first entry beyond
// every lane's k-th distance: a far-away query only touches the blocks of the cap of the cloud that faces it, a
// query inside the cloud only the patches around it.  Boxes are fetched two entries ahead and a needed block's
// points (one coalesced 512 B read) one entry ahead, while the current one is scanned from the warp's
// shared-memory tile.  Each warp ORs its neighbour sets into a private shared-memory mask and then, one word per
// lane, into the object's global bit mask.
//
// (3) knn_emit_kernel, one CTA per object: complete[sorted(S)] repeated cyclically, and |S|.
//
__device__ __forceinline__ unsigned spread10(unsigned v) {   // bit i -> bit 3i
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// lo/inv: box origin and 1023 / extent (0 for a flat axis).  Any monotone-per-axis code would do: the order only
// decides which points share a block, never the result.
__device__ __forceinline__ unsigned morton30(float x, float y, float z, const float* lo, const float* inv) {
    const int ix = min(1023, max(0, (int)((x - lo[0]) * inv[0])));
    const int iy = min(1023, max(0, (int)((y - lo[1]) * inv[1])));
    const int iz = min(1023, max(0, (int)((z - lo[2]) * inv[2])));
    return spread10((unsigned)ix) | (spread10((unsigned)iy) << 1) | (spread10((unsigned)iz) << 2);
}

// prepare smem: keys max(rp2, qp2) u64 | partial SoA 3*np f32
template <int T>
__global__ void __launch_bounds__(T)
knn_prepare_kernel(int np, int r, int rp2, int qp2, const float* __restrict__ partial,
                   const float* __restrict__ complete, float4* __restrict__ ws_refs, float4* __restrict__ ws_box,
                   float4* __restrict__ ws_q, int* __restrict__ ws_meta) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    u64* keys = reinterpret_cast<u64*>(s_raw);
    float* qx = reinterpret_cast<float*>(keys + max(rp2, qp2));
    float* qy = qx + np; float* qz = qy + np;
    __shared__ float s_lo[6][T / 32], s_hi[6][T / 32];   // 0..2 completed cloud, 3..5 queries
    __shared__ float s_org[6], s_inv[6];
    __shared__ int s_cnt[T / 32];

    const int b = blockIdx.x;
    const float* cp = complete + (size_t)b * r * 3;
    const float* pp = partial + (size_t)b * np * 3;
    const float inf = __int_as_float(0x7f800000);
    float lo[6] = {inf, inf, inf, inf, inf, inf}, hi[6] = {-inf, -inf, -inf, -inf, -inf, -inf};
    for (int p = threadIdx.x; p < r; p += T)
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float v = cp[p * 3 + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    for (int p = threadIdx.x; p < np; p += T) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = pp[p * 3 + c];
            (c == 0 ? qx : c == 1 ? qy : qz)[p] = v;
            lo[3 + c] = fminf(lo[3 + c], v); hi[3 + c] = fmaxf(hi[3 + c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], off));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], off));
        }
        if (lane_id() == 0) { s_lo[c][warp_id()] = lo[c]; s_hi[c][warp_id()] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int c = threadIdx.x;
        float l = inf, h = -inf;
        for (int w = 0; w < T / 32; ++w) { l = fminf(l, s_lo[c][w]); h = fmaxf(h, s_hi[c][w]); }
        const float ext = h - l;
        s_org[c] = l;
        s_inv[c] = (ext > 0.f && ext < inf) ? 1023.f / ext : 0.f;
    }
    __syncthreads();

    // completed cloud in Morton order + the box of every 32 consecutive points
    for (int p = threadIdx.x; p < rp2; p += T)
        keys[p] = p < r ? ((u64)morton30(cp[p * 3], cp[p * 3 + 1], cp[p * 3 + 2], s_org, s_inv) << 32) | (unsigned)p : kInfKey;
    __syncthreads();
    bitonic_sort<T>(keys, rp2);
    float4* wr = ws_refs + (size_t)b * r;
    const int nblk = (r + 31) >> 5;
    float4* wb = ws_box + (size_t)b * nblk * 2;
    for (int p0 = warp_id() * 32; p0 < r; p0 += T) {                    // one block per warp pass
        const int p = p0 + lane_id();
        float bl[3] = {inf, inf, inf}, bh[3] = {-inf, -inf, -inf};
        if (p < r) {
            const int i = (int)(unsigned)keys[p];
            const float x = cp[i * 3 + 0], y = cp[i * 3 + 1], z = cp[i * 3 + 2];
            wr[p] = make_float4(x, y, z, __int_as_float(i));
            bl[0] = bh[0] = x; bl[1] = bh[1] = y; bl[2] = bh[2] = z;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                bl[c] = fminf(bl[c], __shfl_xor_sync(0xffffffffu, bl[c], off));
                bh[c] = fmaxf(bh[c], __shfl_xor_sync(0xffffffffu, bh[c], off));
            }
        if (lane_id() == 0) {
            wb[(p0 >> 5) * 2] = make_float4(bl[0], bl[1], bl[2], 0.f);
            wb[(p0 >> 5) * 2 + 1] = make_float4(bh[0], bh[1], bh[2], 0.f);
        }
    }
    __syncthreads();

    // Queries: sort ALL of them by (Morton code, index).  Equal coordinates have equal codes, so the copies of a point
    // sit in one run of equal codes; the first of them in that run (lowest index) represents it.  Representatives are
    // compacted in sorted order straight into the workspace — no hash table, no atomics, no second sort.
    for (int p = threadIdx.x; p < qp2; p += T)
        keys[p] = p < np ? ((u64)morton30(qx[p], qy[p], qz[p], s_org + 3, s_inv + 3) << 32) | (unsigned)p : kInfKey;
    __syncthreads();
    bitonic_sort<T>(keys, qp2);
    float4* wq = ws_q + (size_t)b * np;
    int base = 0;
    for (int p0 = 0; p0 < np; p0 += T) {
        const int p = p0 + threadIdx.x;
        bool rep = false;
        float x = 0.f, y = 0.f, z = 0.f;
        if (p < np) {
            const u64 key = keys[p];
            const unsigned code = (unsigned)(key >> 32);
            const int qi = (int)(unsigned)key;
            x = qx[qi]; y = qy[qi]; z = qz[qi];
            rep = true;
            for (int j = p - 1; j >= 0 && (unsigned)(keys[j] >> 32) == code; --j) {
                const int qj = (int)(unsigned)keys[j];
                if (qx[qj] == x && qy[qj] == y && qz[qj] == z) { rep = false; break; }   // an earlier copy represents it
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, rep);
        if (lane_id() == 0) s_cnt[warp_id()] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < T / 32; ++w) { const int c = s_cnt[w]; if (w < warp_id()) before += c; total += c; }
        if (rep) wq[base + before + __popc(m & ((1u << lane_id()) - 1))] = make_float4(x, y, z, 0.f);
        base += total;
        __syncthreads();
    }
    const int nuniq = base;
    if (threadIdx.x == 0) { ws_meta[2 * b] = nuniq; ws_meta[2 * b + 1] = nblk; }
}

constexpr int kScanThreads = 128;
constexpr int kScanBuf = 8;   // candidate buffer of the block scan (flushed when more than 4 are pending)

// squared distance from q to the box [lo, hi], in the operation order of sqdist(): for a point r inside the box
// |fl(r.x - q.x)| >= ex etc. (rounding is monotone), hence box_sqdist <= sqdist(q, r) in fp32 as well.
__device__ __forceinline__ float box_sqdist(float qx, float qy, float qz, float4 lo, float4 hi) {
    const float ex = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
    const float ey = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
    const float ez = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
    return __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
}

// every lane offers the cnt points of one block staged in the warp's shared-memory tile (broadcast loads)
__device__ __forceinline__ void scan_block(TopK<32>& tk, bool active, float qx, float qy, float qz,
                                           const float4* pts, int cnt) {
    if (cnt == 32) {
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 4) {
            if (active) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { const float4 c = pts[r0 + u]; tk.offer(sqdist(qx, qy, qz, c), __float_as_int(c.w)); }
            }
            if (__any_sync(0xffffffffu, tk.nbuf > kScanBuf - 4)) tk.flush();
        }
    } else {
        for (int r0 = 0; r0 < cnt; r0 += 4) {
            if (active) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (r0 + u < cnt) { const float4 c = pts[r0 + u]; tk.offer(sqdist(qx, qy, qz, c), __float_as_int(c.w)); }
            }
            if (__any_sync(0xffffffffu, tk.nbuf > kScanBuf - 4)) tk.flush();
        }
    }
    tk.flush();
}

// per warp in shared memory: heap k x 32 u64 | buffer kScanBuf x 32 u64 | block tile 32 float4 | union mask nwords u32 |
// walk order pow2(nwords) u32
__host__ __device__ inline size_t scan_warp_smem(int k, int r) {
    const size_t nwords = (size_t)(r + 31) >> 5;
    size_t np2 = 1;
    while (np2 < nwords) np2 <<= 1;
    return ((size_t)(k + kScanBuf) * 32 * 8 + 512 + (nwords + np2) * 4 + 15) & ~(size_t)15;
}

// ascending bitonic sort of s[0..len), len a power of two, by one warp
__device__ __forceinline__ void warp_bitonic_sort_u32(unsigned* s, int len) {
    for (int k = 2; k <= len; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane_id(); t < (len >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const unsigned a = s[i], b = s[p];
                if ((a > b) == ((i & k) == 0)) { s[i] = b; s[p] = a; }
            }
            __syncwarp();
        }
}

__global__ void __launch_bounds__(kScanThreads)
knn_scan_kernel(int np, int r, int k, const float4* __restrict__ ws_refs, const float4* __restrict__ ws_box,
                const float4* __restrict__ ws_q, const int* __restrict__ ws_meta, unsigned* __restrict__ g_mask) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int nwords = (r + 31) >> 5;            // = number of blocks
    const int b = blockIdx.y;
    const int nuniq = ws_meta[2 * b];
    const int u0 = (blockIdx.x * (kScanThreads / 32) + warp_id()) * 32;
    if (u0 >= nuniq) return;                                            // warp-uniform
    unsigned char* mine = s_raw + scan_warp_smem(k, r) * warp_id();
    u64* s_list = reinterpret_cast<u64*>(mine);
    u64* s_buf = s_list + (size_t)k * 32;
    float4* tile = reinterpret_cast<float4*>(s_buf + (size_t)kScanBuf * 32);
    unsigned* mask = reinterpret_cast<unsigned*>(tile + 32);
    unsigned* order = mask + nwords;
    const unsigned full = 0xffffffffu;
    for (int i = lane_id(); i < nwords; i += 32) mask[i] = 0u;

    const float4* refs = ws_refs + (size_t)b * r;
    const float4* box = ws_box + (size_t)b * nwords * 2;
    const int u = u0 + lane_id();
    const bool active = u < nuniq;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) q = ws_q[(size_t)b * np + u];
    TopK<32> tk;
    tk.init(s_list + lane_id(), s_buf + lane_id(), k);

    // Walk order: blocks by the smallest box distance over the warp's queries, nearest first (non-negative float bits
    // order like unsigned), so the walk may stop at the first entry beyond every lane's k-th distance.  Entry =
    // distance bits with the low `ib` bits replaced by the block number (rounds the key DOWN, which keeps the stop
    // test conservative; the order only decides how fast the heaps tighten, never the result).
    const int ib = nwords > 1 ? 32 - __clz(nwords - 1) : 0;
    const int np2 = 1 << ib;
    const unsigned imask = (unsigned)np2 - 1u;
#pragma unroll 4
    for (int blk = 0; blk < nwords; ++blk) {
        const float bd = box_sqdist(q.x, q.y, q.z, __ldg(&box[2 * blk]), __ldg(&box[2 * blk + 1]));
        const unsigned m = __reduce_min_sync(full, active ? __float_as_uint(bd) : 0xffffffffu);
        if (lane_id() == 0) order[blk] = (m & ~imask) | (unsigned)blk;
    }
    for (int j = nwords + lane_id(); j < np2; j += 32) order[j] = 0xffffffffu;
    __syncwarp();
    warp_bitonic_sort_u32(order, np2);

    // Entry i's box is loaded two iterations ahead and its points one iteration ahead (when some lane still needs the
    // block by the k-th distances as they stand then: they only shrink, so a needed block is never left out).
    auto entry_blk = [&](int i) -> int { return i < nwords ? (int)(order[i] & imask) : -1; };
    auto load_box = [&](int blk, float4& lo, float4& hi) {
        if (blk >= 0) { lo = __ldg(&box[2 * blk]); hi = __ldg(&box[2 * blk + 1]); }
    };
    auto needed = [&](float4 lo, float4 hi) -> bool {
        return __any_sync(full, active && box_sqdist(q.x, q.y, q.z, lo, hi) <= tk.thr_f);
    };
    auto load_pts = [&](int blk) -> float4 {
        const int p = blk * 32 + lane_id();
        return p < r ? __ldg(&refs[p]) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int blk0 = entry_blk(0), blk1 = entry_blk(1);
    float4 lo0 = zero4, hi0 = zero4, lo1 = zero4, hi1 = zero4;
    load_box(blk0, lo0, hi0);
    load_box(blk1, lo1, hi1);
    bool need0 = needed(lo0, hi0);                // heaps are empty: true whenever the warp has a query
    float4 pts0 = need0 ? load_pts(blk0) : zero4;
    unsigned tmax = 0x7f800000u;
    for (int i = 0; i < nwords; ++i) {
        if ((order[i] & ~imask) > tmax) break;    // sorted: every later entry is farther still
        const int blk2 = entry_blk(i + 2);
        float4 lo2 = zero4, hi2 = zero4;
        load_box(blk2, lo2, hi2);
        const bool need1 = blk1 >= 0 && needed(lo1, hi1);
        const float4 pts1 = need1 ? load_pts(blk1) : zero4;
        if (need0 && needed(lo0, hi0)) {
            tile[lane_id()] = pts0;
            __syncwarp();
            scan_block(tk, active, q.x, q.y, q.z, tile, min(32, r - blk0 * 32));
            __syncwarp();
            tmax = __reduce_max_sync(full, active ? __float_as_uint(tk.thr_f) : 0u);
        }
        blk0 = blk1; lo0 = lo1; hi0 = hi1; need0 = need1; pts0 = pts1;
        blk1 = blk2; lo1 = lo2; hi1 = hi2;
    }
    if (active)
        for (int j = 0; j < k; ++j) {
            const u64 key = tk.list[j * 32];
            if (key != kInfKey) { const unsigned i = (unsigned)key; atomicOr(&mask[i >> 5], 1u << (i & 31)); }
        }
    __syncwarp();
    unsigned* gm = g_mask + (size_t)b * nwords;
    for (int w = lane_id(); w < nwords; w += 32)
        if (mask[w]) atomicOr(&gm[w], mask[w]);
}

// One CTA per object: complete[sorted(S)] repeated cyclically, S = the bits of the object's union mask.
// smem: mask nwords u32 | prefix (nwords + 1) i32
template <int T>
__global__ void __launch_bounds__(T)
knn_emit_kernel(int r, int surface_pts, const float* __restrict__ complete, const unsigned* __restrict__ g_mask,
                float* __restrict__ out, int* __restrict__ sel_count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int nwords = (r + 31) >> 5;
    unsigned* mask = reinterpret_cast<unsigned*>(s_raw);
    int* prefix = reinterpret_cast<int*>(mask + nwords);
    __shared__ int s_total;
    const int b = blockIdx.x;
    const unsigned* gm = g_mask + (size_t)b * nwords;
    for (int w = threadIdx.x; w < nwords; w += T) mask[w] = gm[w];
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

struct SelectWs { size_t mask, arrive, meta, refs, box, q, total; };
SelectWs select_ws(int b, int np, int r) {
    const size_t B = b > 0 ? b : 0, R = r > 0 ? r : 0, N = np > 0 ? np : 0;
    SelectWs w;
    w.mask = 0;
    w.arrive = w.mask + B * ((R + 31) / 32) * 4;
    w.meta = w.arrive + B * 4;
    w.refs = align_up(w.meta + B * 8, 256);
    w.box = w.refs + B * R * 16;
    w.q = w.box + B * ((R + 31) / 32) * 32;
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
    int* ws_meta = reinterpret_cast<int*>(base + w.meta);
    float4* ws_refs = reinterpret_cast<float4*>(base + w.refs);
    float4* ws_box = reinterpret_cast<float4*>(base + w.box);
    float4* ws_q = reinterpret_cast<float4*>(base + w.q);
    SEEVCN_PROF("knn_surface_select", st);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(base, 0, w.meta, st));
    {
        constexpr int TP = 512;
        const int rp2 = next_pow2(r), qp2 = next_pow2(n_partial > 0 ? n_partial : 1);
        const size_t smem = (size_t)(rp2 > qp2 ? rp2 : qp2) * 8 + (size_t)n_partial * 12;
        auto kern = knn_prepare_kernel<TP>;
        if (smem > 40 * 1024)
            SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SEEVCN_PROF("knn_prepare_kernel", st);
        kern<<<b, TP, smem, st>>>(n_partial, r, rp2, qp2, partial, complete, ws_refs, ws_box, ws_q, ws_meta);
        SEEVCN_LAUNCH_CHECK();
    }
    {
        const int nwords = (r + 31) / 32;
        const size_t smem = (size_t)(kScanThreads / 32) * scan_warp_smem(k, r);
        if (smem > 40 * 1024)
            SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(knn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int chunks = n_partial > 0 ? div_up(n_partial, kScanThreads) : 1;
        {
            SEEVCN_PROF("knn_scan_kernel", st);
            knn_scan_kernel<<<dim3(chunks, b), kScanThreads, smem, st>>>(n_partial, r, k, ws_refs, ws_box, ws_q, ws_meta, g_mask);
            SEEVCN_LAUNCH_CHECK();
        }
        SEEVCN_PROF("knn_emit_kernel", st);
        knn_emit_kernel<256><<<b, 256, (size_t)(2 * nwords + 1) * 4, st>>>(r, surface_pts, complete, g_mask, out, sel_count);
        SEEVCN_LAUNCH_CHECK();
    }
    return SEEVCN_OK;
}
