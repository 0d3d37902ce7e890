// Largest-cluster filter — the step right after the kNN surface selection in VCN.inference.
//
// Replaces get_largest_cluster / get_largest_cluster_batch
// (see/surface_completion/models/vcn/utils/sampling.py:83-109), which runs open3d's cluster_dbscan
// per object on the host (eps = CLUSTER_EPS, min_points = 2, models/VCN.py:95-98).
//
// With min_points <= 2 DBSCAN is exactly "connected components of the eps-graph": a point is core iff
// its eps-ball holds min_points points counting itself, any neighbour of a core point is core by
// symmetry, and isolated points (min_points = 2) are noise, i.e. components of size 1.
//
// One CTA per object, everything on chip (24 B of shared memory per point, so several CTAs per SM):
//   1. every unordered pair {i, j} is tested ONCE: thread t owns rows t and n-1-t (n-1 tests per
//      thread, balanced) and scans j > i four at a time with 128-bit shared-memory loads.  The test is
//      fp32 with a guard band; only pairs whose fp32 distance lies within 1e-5 relative of eps^2 are
//      re-evaluated in float64 with separately rounded operations (open3d's KD-tree works on float64
//      copies of the points, strict `dist < radius^2`), so the decision is the float64 one everywhere.
//   2. adjacent pairs are merged in a lock-free union-find in shared memory (CAS-link the larger root
//      under the smaller one, path halving), so a component's root is its smallest point index
//      = open3d's label order.
//   3. component sizes, largest (ties -> first label, np.argmax), members in ascending row order,
//      tiled cyclically to total_pts rows (np.tile(...)[:total_pts]).
// PARITY UNPINNED: open3d is not vendored (see/surface_completion/setup.py:25 pins 0.14.1).
#include "common.cuh"

namespace {

constexpr int kMaxN = 8192;      // 24 B/point of shared memory
constexpr int kThreads = 512;

__device__ __forceinline__ int uf_find(volatile int* par, int x) {
    int p = par[x];
    while (p != x) {
        const int g = par[p];
        if (g != p) par[x] = g;   // path halving: only ever rewrites a non-root with a (possibly stale) ancestor
        x = p; p = g;
    }
    return x;
}

__device__ __forceinline__ void uf_union(volatile int* par, int a, int b) {
    while (true) {
        a = uf_find(par, a); b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // a > b: link the larger root under the smaller
        if (atomicCAS(const_cast<int*>(par) + a, a, b) == a) return;
    }
}

__device__ __forceinline__ bool exact_adjacent(const float* sx, const float* sy, const float* sz, int i, int j, double eps2) {
    const double dx = __dsub_rn((double)sx[i], (double)sx[j]), dy = __dsub_rn((double)sy[i], (double)sy[j]),
                 dz = __dsub_rn((double)sz[i], (double)sz[j]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) < eps2;
}

__global__ void __launch_bounds__(kThreads)
largest_cluster_kernel(int n, int total_pts, double eps2, float lo, float hi, int min_points,
                       const float* __restrict__ pts, const int* __restrict__ period,
                       float* __restrict__ out, int* __restrict__ out_count, int* __restrict__ out_distinct) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int b = blockIdx.x, t = threadIdx.x;
    // rows repeat with period m (pts[r] == pts[r % m], the tiling of the kNN surface selection): cluster the
    // m distinct rows, row u standing for w(u) = #{r < n : r % m == u} rows.  No hint: m = n, w = 1.
    int m = n;
    if (period) { m = period[b]; m = m < 1 ? (n > 0 ? 1 : 0) : (m > n ? n : m); }
    const int m4 = (m + 3) & ~3;
    float* sx = reinterpret_cast<float*>(s_raw);
    float* sy = sx + m4;
    float* sz = sy + m4;
    int* par = reinterpret_cast<int*>(sz + m4);
    int* size = par + m4;
    int* list = size + m4;
    __shared__ int s_best, s_bestsize, s_run;
    __shared__ int s_warp_cnt[kThreads / 32];

    const float* p = pts + (size_t)b * n * 3;
    for (int f = t; f < m * 3; f += kThreads) {
        const float v = p[f]; const int k = f / 3, c = f - 3 * k;
        (c == 0 ? sx : c == 1 ? sy : sz)[k] = v;
    }
    for (int i = t; i < m4; i += kThreads) {
        par[i] = i; size[i] = 0;
        if (i >= m) { sx[i] = 1e30f; sy[i] = 1e30f; sz[i] = 1e30f; }   // padding: infinitely far from everything
    }
    if (t == 0) { s_best = -1; s_bestsize = 0; s_run = 0; }
    __syncthreads();
    volatile int* vpar = par;

    // 1 + 2. pair tests (fp32 screen, float64 in the guard band) + union-find.  `root` caches the row's
    // root: once a component has formed, a neighbour costs one compare of its parent against it.
    auto scan_row = [&](int i) {
        const float xi = sx[i], yi = sy[i], zi = sz[i];
        int root = uf_find(vpar, i);
        for (int j0 = (i + 1) & ~3; j0 < m; j0 += 4) {
            const float4 X = *reinterpret_cast<const float4*>(sx + j0);
            const float4 Y = *reinterpret_cast<const float4*>(sy + j0);
            const float4 Z = *reinterpret_cast<const float4*>(sz + j0);
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
            float d2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = xi - xs[u], dy = yi - ys[u], dz = zi - zs[u];
                d2[u] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            }
            const float dmin = fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3]));
            if (dmin < hi) {
                const int4 P = *reinterpret_cast<const int4*>(par + j0);   // may be stale: only a shortcut
                const int ps[4] = {P.x, P.y, P.z, P.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u;
                    if (ps[u] != root && j > i && d2[u] < hi && (d2[u] < lo || exact_adjacent(sx, sy, sz, i, j, eps2))) {
                        uf_union(vpar, root, j);
                        root = uf_find(vpar, i);
                    }
                }
            }
        }
    };
    for (int i = t; 2 * i < m; i += kThreads) {
        scan_row(i);
        const int i2 = m - 1 - i;
        if (i2 != i) scan_row(i2);
    }
    __syncthreads();

    // 3. weighted sizes -> best root (largest, ties -> smallest root) -> member list in ascending order
    for (int i = t; i < m; i += kThreads) atomicAdd(&size[uf_find(vpar, i)], (n - i + m - 1) / m);
    __syncthreads();
    // from here on par[] is only read (uf_find may still halve paths: benign)
    for (int i = t; i < m; i += kThreads)
        if (vpar[i] == i && size[i] >= min_points) atomicMax(&s_bestsize, size[i]);
    __syncthreads();
    for (int i = t; i < m; i += kThreads)
        if (vpar[i] == i && size[i] >= min_points && size[i] == s_bestsize)
            atomicMin(reinterpret_cast<unsigned*>(&s_best), (unsigned)i);
    __syncthreads();
    const int best = s_best;
    for (int i0 = 0; i0 < m; i0 += kThreads) {
        const int i = i0 + t;
        const bool member = best >= 0 && i < m && uf_find(vpar, i) == best;
        const unsigned bal = __ballot_sync(0xffffffffu, member);
        if (lane_id() == 0) s_warp_cnt[warp_id()] = __popc(bal);
        __syncthreads();
        if (warp_id() == 0) {
            const int c = lane_id() < kThreads / 32 ? s_warp_cnt[lane_id()] : 0;
            int inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane_id() >= off) inc += v;
            }
            if (lane_id() < kThreads / 32) s_warp_cnt[lane_id()] = s_run + inc - c;
            __syncwarp();
            if (lane_id() == 31) s_run += inc;
        }
        __syncthreads();
        if (member) list[s_warp_cnt[warp_id()] + __popc(bal & ((1u << lane_id()) - 1))] = i;
        __syncthreads();
    }
    // member ROWS in ascending order: period q holds rows q*m + list[k] (k ascending); the last, partial
    // period keeps a prefix of the list, so rank -> (q, k) = (rank / c, rank % c) for every rank < count
    const int c = best >= 0 ? s_run : 0;
    const int count = best >= 0 ? s_bestsize : 0;
    if (t == 0) {
        out_count[b] = count;                      // member ROWS (a tiled row counts every time it appears): np.bincount's size
        if (out_distinct) out_distinct[b] = c;     // members among the clustered (distinct) rows = leading rows of `out` that differ
    }
    float* o = out + (size_t)b * total_pts * 3;
    for (int f = t; f < total_pts * 3; f += kThreads) {
        float v = 0.f;
        if (count > 0) {
            const int j = f / 3, ch = f - 3 * j;
            const int s = list[(j % count) % c];
            v = (ch == 0 ? sx : ch == 1 ? sy : sz)[s];
        }
        o[f] = v;
    }
}

}  // namespace

static int largest_cluster_launch(int b, int n, int total_pts, double eps, int min_points, const float* pts,
                                  const int* period, float* out, int* out_count, int* out_distinct, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n >= 0 && total_pts >= 0, "largest_cluster: negative size");
    SEEVCN_REQUIRE(n <= kMaxN, "largest_cluster: n=%d > %d points per object", n, kMaxN);
    SEEVCN_REQUIRE(min_points >= 1 && min_points <= 2,
                   "largest_cluster: min_points=%d; only 1 or 2 (connected components) are supported", min_points);
    if (b == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(pts && out && out_count, "largest_cluster: null pointer");
    const int n4 = (n + 3) & ~3;
    const size_t smem = (size_t)n4 * 24 + 16;
    // per device and cheap: set on every launch
    SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(largest_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxN * 24 + 16));
    const double e2 = eps * eps;
    // fp32 screen: below `lo` certainly adjacent, at or above `hi` certainly not (fp32 error of the
    // distance <= ~4 ulp = 2.4e-7 relative); in between the float64 test decides
    const float lo = (float)(e2 * (1.0 - 1e-5)), hi = (float)(e2 * (1.0 + 1e-5));
    SEEVCN_PROF("largest_cluster", as_stream(stream));
    largest_cluster_kernel<<<b, kThreads, smem, as_stream(stream)>>>(n, total_pts, e2, lo, hi, min_points, pts, period, out,
                                                                    out_count, out_distinct);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_largest_cluster(int b, int n, int total_pts, double eps, int min_points, const float* pts,
                                      float* out, int* out_count, int* out_distinct, seevcn_stream_t stream) {
    return largest_cluster_launch(b, n, total_pts, eps, min_points, pts, nullptr, out, out_count, out_distinct, stream);
}

extern "C" int seevcn_largest_cluster_periodic(int b, int n, int total_pts, double eps, int min_points, const float* pts,
                                               const int* period, float* out, int* out_count, int* out_distinct, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(period || b == 0, "largest_cluster_periodic: null period");
    SEEVCN_REQUIRE(eps > 0.0, "largest_cluster_periodic: eps must be > 0 (duplicate rows are merged)");
    return largest_cluster_launch(b, n, total_pts, eps, min_points, pts, period, out, out_count, out_distinct, stream);
}
