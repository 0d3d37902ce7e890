// Largest-cluster filter — the step right after the kNN surface selection in VCN.inference.
//
// Replaces get_largest_cluster / get_largest_cluster_batch
// (see/surface_completion/models/vcn/utils/sampling.py:83-109), which runs open3d's cluster_dbscan
// per object on the host (eps = CLUSTER_EPS, min_points = 2, models/VCN.py:95-98).
//
// With min_points <= 2 DBSCAN is exactly "connected components of the eps-graph": a point is core iff
// its eps-ball holds min_points points counting itself, any neighbour of a core point is core by
// symmetry, and isolated points (min_points = 2) are noise.  Per object (one CTA, n <= 1024 points):
//   1. adjacency as an n x n bit matrix in shared memory (128 KB at n = 1024); squared distances in
//      float64 with separately rounded operations, strict `<` against eps^2 — open3d's KD-tree works on
//      float64 copies of the points (nanoflann radius search, `dist < radius^2`)
//   2. min-label propagation over the bit rows + pointer jumping until nothing changes
//      (labels converge to the smallest point index of each component = open3d's label order)
//   3. component sizes, largest (ties -> first label, np.argmax), members in ascending row order,
//      tiled cyclically to total_pts rows (np.tile(...)[:total_pts]).
// PARITY UNPINNED: open3d is not vendored (see/surface_completion/setup.py:25 pins 0.14.1).
#include "common.cuh"

namespace {

constexpr int kMaxN = 1024;
constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads, 1)
largest_cluster_kernel(int n, int total_pts, double eps2, int min_points, const float* __restrict__ pts,
                       float* __restrict__ out, int* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int nw = (n + 31) >> 5;
    float* sx = reinterpret_cast<float*>(s_raw);
    float* sy = sx + kMaxN;
    float* sz = sy + kMaxN;
    int* lab = reinterpret_cast<int*>(sz + kMaxN);
    int* size = lab + kMaxN;
    int* list = size + kMaxN;
    unsigned* adj = reinterpret_cast<unsigned*>(list + kMaxN);   // n rows x nw words
    __shared__ int s_best, s_bestsize, s_members;
    __shared__ int s_warp_cnt[32];

    const int b = blockIdx.x, i = threadIdx.x;
    const float* p = pts + (size_t)b * n * 3;
    for (int f = i; f < n * 3; f += kThreads) {
        const float v = p[f]; const int k = f / 3, c = f - 3 * k;
        (c == 0 ? sx : c == 1 ? sy : sz)[k] = v;
    }
    if (i == 0) { s_best = -1; s_bestsize = 0; }
    __syncthreads();

    // 1. adjacency row i
    int deg = 0;
    if (i < n) {
        const double xi = sx[i], yi = sy[i], zi = sz[i];
        for (int w = 0; w < nw; ++w) {
            unsigned bits = 0;
            const int jend = min(32, n - w * 32);
            for (int t = 0; t < jend; ++t) {
                const int j = w * 32 + t;
                const double dx = __dsub_rn(xi, (double)sx[j]), dy = __dsub_rn(yi, (double)sy[j]), dz = __dsub_rn(zi, (double)sz[j]);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (d2 < eps2) bits |= 1u << t;
            }
            adj[(size_t)i * nw + w] = bits;
            deg += __popc(bits);
        }
    }
    const bool core = i < n && deg >= min_points;   // deg counts the point itself
    if (i < kMaxN) { lab[i] = core ? i : 0x7fffffff; size[i] = 0; }
    __syncthreads();

    // 2. min-label propagation + pointer jumping
    while (true) {
        bool changed = false;
        if (core) {
            int m = lab[i];
            for (int w = 0; w < nw; ++w) {
                unsigned bits = adj[(size_t)i * nw + w];
                while (bits) {
                    const int t = __ffs(bits) - 1;
                    bits &= bits - 1;
                    m = min(m, lab[w * 32 + t]);
                }
            }
            if (m < lab[i]) { lab[i] = m; changed = true; }
        }
        __syncthreads();
        if (core) {
            int l = lab[i];
            for (int hop = 0; hop < 4; ++hop) l = lab[l];   // monotone: racing readers only see smaller labels
            if (l < lab[i]) { lab[i] = l; changed = true; }
        }
        if (!__syncthreads_or(changed)) break;
    }

    // 3. sizes -> best root (largest, ties -> smallest root) -> member list in ascending order
    if (core) atomicAdd(&size[lab[i]], 1);
    __syncthreads();
    if (core && lab[i] == i) {
        const int packed_self = size[i];
        atomicMax(&s_bestsize, packed_self);
    }
    __syncthreads();
    if (core && lab[i] == i && size[i] == s_bestsize) atomicMin(reinterpret_cast<unsigned*>(&s_best), (unsigned)i);
    __syncthreads();
    const int best = s_best;
    const bool member = core && best >= 0 && lab[i] == best;
    const unsigned bal = __ballot_sync(0xffffffffu, member);
    if (lane_id() == 0) s_warp_cnt[warp_id()] = __popc(bal);
    __syncthreads();
    if (warp_id() == 0) {
        int c = s_warp_cnt[lane_id()];
        int inc = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane_id() >= off) inc += t;
        }
        s_warp_cnt[lane_id()] = inc - c;
        if (lane_id() == 31) s_members = inc;
    }
    __syncthreads();
    if (member) list[s_warp_cnt[warp_id()] + __popc(bal & ((1u << lane_id()) - 1))] = i;
    __syncthreads();
    const int m = best >= 0 ? s_members : 0;
    if (i == 0) out_count[b] = m;
    float* o = out + (size_t)b * total_pts * 3;
    for (int j = i; j < total_pts; j += kThreads) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (m > 0) { const int s = list[j % m]; x = sx[s]; y = sy[s]; z = sz[s]; }
        o[j * 3 + 0] = x; o[j * 3 + 1] = y; o[j * 3 + 2] = z;
    }
}

}  // namespace

extern "C" int seevcn_largest_cluster(int b, int n, int total_pts, double eps, int min_points, const float* pts,
                                      float* out, int* out_count, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n >= 0 && total_pts >= 0, "largest_cluster: negative size");
    SEEVCN_REQUIRE(n <= kMaxN, "largest_cluster: n=%d > %d points per object", n, kMaxN);
    SEEVCN_REQUIRE(min_points >= 1 && min_points <= 2,
                   "largest_cluster: min_points=%d; only 1 or 2 (connected components) are supported", min_points);
    if (b == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(pts && out && out_count, "largest_cluster: null pointer");
    const int nw = (n + 31) / 32;
    const size_t smem = (size_t)kMaxN * 4 * 6 + (size_t)n * nw * 4 + 16;
    static bool attr = false;
    if (!attr) {
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(largest_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const double e = eps;
    largest_cluster_kernel<<<b, kThreads, smem, as_stream(stream)>>>(n, total_pts, e * e, min_points, pts, out, out_count);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
