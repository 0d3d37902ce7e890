// Mask-based isolation — the target-domain (no ground-truth boxes) front end of the path.
//
// Replaces, for one frame and one camera,
//   map_pointcloud_to_image   see/surface_completion/datasets/custom_dataset/custom_dataset_objects.py:141-193
//                             (LiDAR -> camera frame, pre-distortion cull, pinhole / equidistant distortion, intrinsics,
//                              field-of-view cull, np.round to pixels; all in float64 numpy)
//   get_pts_in_mask           see/surface_completion/datasets/shared_utils.py:36-106
//                             (per instance: mask[v, u] lookup of every in-view point -> the instance's points)
//   SEE_VCN.isolate_det_pts   see/surface_completion/SEE_VCN.py:144-181
//                             (per instance: eps from the range of its centre, open3d cluster_dbscan(eps, min_points = 3),
//                              largest cluster, kept when it holds more than min_cluster points)
// which the reference runs on the host with numpy / pycocotools / open3d.  The instance masks themselves (polygon ->
// binary mask, pycocotools annToMask) are the caller's input: I binary images (I, H, W) uint8.
//
// DBSCAN with any min_points, without open3d's sequential expansion (PARITY UNPINNED: open3d is not vendored):
//   core(i)   <=> #{j : d(i,j) < eps} >= min_points, the point itself included (float64 distances, strict <)
//   clusters  =   connected components of the core points under core-core adjacency; open3d discovers clusters in
//                 the order of their lowest-index core point, so cluster ids order like component roots (= min index)
//   border b  ->  a non-core point adjacent to a core point joins the FIRST cluster that reaches it = the adjacent
//                 cluster discovered first = the smallest root among its core neighbours
//   noise     =   everything else
// One CTA per instance, everything on chip: degrees (a thread per row), union-find over core-core pairs (each
// unordered pair once), border labels, sizes, largest (ties: first cluster, np.argmax), members in ascending order.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- projection --
struct CamModel {
    double ext[12];      // lidar2cam[:3, :], row major
    double fx, fy, cx, cy;
    double d[5];         // distortion coefficients (pinhole: k1 k2 p1 p2 k3 as distcoeff[0..4]; equidistant: [0..3])
    double pre_limit;    // arctan(IMG_W / IMG_H)
    int width, height, equidistant;
};

// ref: custom_dataset_objects.py:156-187.  uv (N,2) int32 = np.round(u), np.round(v) (half to even) or (-1,-1);
// fov (N) uint8; depth (N) f32 or NULL.
__global__ void __launch_bounds__(256)
project_kernel(int n, const float* __restrict__ pts, CamModel cm, int2* __restrict__ uv, unsigned char* __restrict__ fov,
               float* __restrict__ depth) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
    const double X = fma(cm.ext[0], x, fma(cm.ext[1], y, fma(cm.ext[2], z, cm.ext[3])));
    const double Y = fma(cm.ext[4], x, fma(cm.ext[5], y, fma(cm.ext[6], z, cm.ext[7])));
    const double Z = fma(cm.ext[8], x, fma(cm.ext[9], y, fma(cm.ext[10], z, cm.ext[11])));
    bool ok = Z > 0.0;
    const double tx = X / Z, ty = Y / Z;
    ok = ok && fabs(tx) < cm.pre_limit;
    double u = 0.0, v = 0.0;
    if (ok) {
        const double r2 = tx * tx + ty * ty;
        if (cm.equidistant) {
            const double r1 = sqrt(r2), a0 = atan(r1), a2 = a0 * a0;
            const double a1 = a0 * (1.0 + cm.d[0] * a2 + cm.d[1] * (a2 * a2) + cm.d[2] * (a2 * a2 * a2) + cm.d[3] * (a2 * a2 * a2 * a2));
            u = (a1 / r1) * tx; v = (a1 / r1) * ty;
        } else {
            const double td = 1.0 + cm.d[0] * r2 + cm.d[1] * (r2 * r2) + cm.d[4] * (r2 * r2 * r2);
            u = tx * td + 2.0 * cm.d[2] * tx * ty + cm.d[3] * (r2 + 2.0 * tx * tx);
            v = ty * td + cm.d[2] * (r2 + 2.0 * ty * ty) + 2.0 * cm.d[3] * tx * ty;
        }
        u = cm.fx * u + cm.cx; v = cm.fy * v + cm.cy;
        ok = u > 0.0 && u < (double)(cm.width - 1) && v > 0.0 && v < (double)(cm.height - 1);
    }
    uv[p] = ok ? make_int2((int)rint(u), (int)rint(v)) : make_int2(-1, -1);
    fov[p] = ok ? 1 : 0;
    if (depth) depth[p] = ok ? (float)Z : 0.f;
}

// ---------------------------------------------------------------------------------------------- mask lookup --
// One CTA per instance: the in-view points whose pixel is set in the instance's mask, ascending point index
// (what boolean indexing gives, shared_utils.py:77-80).  lists (I, n) int32, counts (I).
__global__ void __launch_bounds__(256)
mask_lookup_kernel(int n, int width, int height, const int2* __restrict__ uv, const unsigned char* __restrict__ fov,
                   const unsigned char* __restrict__ masks, int* __restrict__ lists, int* __restrict__ counts) {
    __shared__ int s_w[8];
    __shared__ int s_base;
    const int inst = blockIdx.x;
    const unsigned char* m = masks + (size_t)inst * width * height;
    int* out = lists + (size_t)inst * n;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int p0 = 0; p0 < n; p0 += 256) {
        const int p = p0 + threadIdx.x;
        bool in = false;
        if (p < n && fov[p]) { const int2 q = uv[p]; in = m[(size_t)q.y * width + q.x] != 0; }
        const unsigned b = __ballot_sync(0xffffffffu, in);
        if (lane_id() == 0) s_w[warp_id()] = __popc(b);
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < 8; ++w) { const int c = s_w[w]; if (w < warp_id()) before += c; tot += c; }
        if (in) out[s_base + before + __popc(b & ((1u << lane_id()) - 1))] = p;
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[inst] = s_base;
}

// ------------------------------------------------------------------------------------------------- DBSCAN --
constexpr int kDbMaxN = 9600;    // 24 B of shared memory per point (225 KB); larger instances use the global workspace
constexpr int kDbThreads = 512;

struct EpsRule {                 // eps = fixed, or clip(scaling * |centre| * tan(vres deg), min, max) (SEE_VCN.py:167-170)
    int adaptive;
    double fixed, vres_deg, scaling, min_eps, max_eps;
};

__device__ __forceinline__ int uf_find(volatile int* par, int x) {
    int p = par[x];
    while (p != x) {
        const int g = par[p];
        if (g != p) par[x] = g;
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void uf_union(volatile int* par, int a, int b) {
    while (true) {
        a = uf_find(par, a); b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }          // link the larger root under the smaller
        if (atomicCAS(const_cast<int*>(par) + a, a, b) == a) return;
    }
}

__device__ __forceinline__ bool adjacent(const float* sx, const float* sy, const float* sz, int i, int j, float lo, float hi,
                                         double eps2) {
    const float dx = sx[i] - sx[j], dy = sy[i] - sy[j], dz = sz[i] - sz[j];
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (d2 >= hi) return false;
    if (d2 < lo) return true;
    const double ex = __dsub_rn((double)sx[i], (double)sx[j]), ey = __dsub_rn((double)sy[i], (double)sy[j]),
                 ez = __dsub_rn((double)sz[i], (double)sz[j]);
    return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)) < eps2;
}

// lists (I, stride) point indices into pts (P,3), counts (I).  out_lists (I, stride): members of the largest cluster,
// ascending; out_counts (I): their number, 0 when the instance has <= min_instance_pts points, no cluster, or a largest
// cluster of <= min_instance_pts points (SEE_VCN.py:155,176).  out_eps (I) float64 or NULL.
__global__ void __launch_bounds__(kDbThreads)
dbscan_largest_kernel(int stride, const float* __restrict__ pts, const int* __restrict__ lists, const int* __restrict__ counts,
                      EpsRule rule, int min_points, int min_instance_pts, int* __restrict__ out_lists,
                      int* __restrict__ out_counts, double* __restrict__ out_eps, unsigned char* __restrict__ gws,
                      unsigned long long gws_bytes, unsigned long long* __restrict__ gws_bump) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ unsigned long long s_goff;
    __shared__ double s_red[3][kDbThreads / 32];
    __shared__ double s_eps2;
    __shared__ float s_lo, s_hi;
    __shared__ int s_best, s_bestsize, s_run, s_wcnt[kDbThreads / 32];
    const int inst = blockIdx.x, t = threadIdx.x;
    const int n = counts[inst];
    const int* lst = lists + (size_t)inst * stride;
    int* olst = out_lists + (size_t)inst * stride;
    if (n <= min_instance_pts) {                                // too few points
        if (t == 0) { out_counts[inst] = 0; if (out_eps) out_eps[inst] = 0.0; }
        return;
    }
    unsigned char* base = s_raw;
    if (n > kDbMaxN) {                                          // does not fit on chip: a slice of the global workspace
        const unsigned long long need = ((unsigned long long)n * 24 + 255) & ~255ull;
        if (t == 0) s_goff = gws ? atomicAdd(gws_bump, need) : ~0ull;
        __syncthreads();
        if (!gws || s_goff + need > gws_bytes) {
            if (t == 0) { out_counts[inst] = -1; if (out_eps) out_eps[inst] = 0.0; }   // caller retries with more workspace
            return;
        }
        base = gws + s_goff;
    }
    float* sx = reinterpret_cast<float*>(base);
    float* sy = sx + n; float* sz = sy + n;
    int* par = reinterpret_cast<int*>(sz + n);
    int* size = par;              // cluster sizes reuse the union-find array once the labels are final
    int* label = par + n;         // root of the point's cluster, -1 = noise
    int* deg = label + n;
    int* member = deg;            // the member list reuses the degrees once the labels are final
    double sum[3] = {0.0, 0.0, 0.0};
    for (int i = t; i < n; i += kDbThreads) {
        const int p = lst[i];
        const float x = pts[p * 3], y = pts[p * 3 + 1], z = pts[p * 3 + 2];
        sx[i] = x; sy[i] = y; sz[i] = z;
        par[i] = i; label[i] = -1;
        sum[0] += x; sum[1] += y; sum[2] += z;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int s = 16; s > 0; s >>= 1) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], s);
        if (lane_id() == 0) s_red[c][warp_id()] = sum[c];
    }
    if (t == 0) { s_best = -1; s_bestsize = 0; s_run = 0; }
    __syncthreads();
    if (t == 0) {
        double eps = rule.fixed;
        if (rule.adaptive) {
            double c[3] = {0.0, 0.0, 0.0};
            for (int w = 0; w < kDbThreads / 32; ++w) { c[0] += s_red[0][w]; c[1] += s_red[1][w]; c[2] += s_red[2][w]; }
            c[0] /= n; c[1] /= n; c[2] /= n;
            const double dist = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            const double ring = dist * tan(rule.vres_deg * 3.141592653589793 / 180.0);
            eps = fmin(fmax(rule.scaling * ring, rule.min_eps), rule.max_eps);
        }
        s_eps2 = eps * eps;
        s_lo = (float)(eps * eps * (1.0 - 1e-5)); s_hi = (float)(eps * eps * (1.0 + 1e-5));
        if (out_eps) out_eps[inst] = eps;
    }
    __syncthreads();
    const double eps2 = s_eps2;
    const float lo = s_lo, hi = s_hi;
    volatile int* vpar = par;
    // 1. degrees: a thread per row, the point itself included
    for (int i = t; i < n; i += kDbThreads) {
        int d = 0;
        for (int j = 0; j < n; ++j) d += adjacent(sx, sy, sz, i, j, lo, hi, eps2) ? 1 : 0;
        deg[i] = d;
    }
    __syncthreads();
    // 2. core-core components: every unordered pair once (rows i and n-1-i per thread: balanced)
    auto scan_row = [&](int i) {
        if (deg[i] < min_points) return;
        for (int j = i + 1; j < n; ++j)
            if (deg[j] >= min_points && adjacent(sx, sy, sz, i, j, lo, hi, eps2)) uf_union(vpar, i, j);
    };
    for (int i = t; 2 * i < n; i += kDbThreads) {
        scan_row(i);
        const int i2 = n - 1 - i;
        if (i2 != i) scan_row(i2);
    }
    __syncthreads();
    // 3. labels: core -> its root; border -> the smallest root among its core neighbours; else noise
    for (int i = t; i < n; i += kDbThreads) {
        int l = -1;
        if (deg[i] >= min_points) l = uf_find(vpar, i);
        else {
            int best = 0x7fffffff;
            for (int j = 0; j < n; ++j)
                if (deg[j] >= min_points && adjacent(sx, sy, sz, i, j, lo, hi, eps2)) best = min(best, uf_find(vpar, j));
            if (best != 0x7fffffff) l = best;
        }
        label[i] = l;
    }
    __syncthreads();
    for (int i = t; i < n; i += kDbThreads) size[i] = 0;       // par is dead from here on
    __syncthreads();
    for (int i = t; i < n; i += kDbThreads)
        if (label[i] >= 0) atomicAdd(&size[label[i]], 1);
    __syncthreads();
    for (int i = t; i < n; i += kDbThreads)
        if (size[i] > 0) atomicMax(&s_bestsize, size[i]);
    __syncthreads();
    for (int i = t; i < n; i += kDbThreads)
        if (size[i] > 0 && size[i] == s_bestsize) atomicMin(reinterpret_cast<unsigned*>(&s_best), (unsigned)i);
    __syncthreads();
    const int best = s_best;
    // 4. members in ascending order
    for (int i0 = 0; i0 < n; i0 += kDbThreads) {
        const int i = i0 + t;
        const bool mem = best >= 0 && i < n && label[i] == best;
        const unsigned bal = __ballot_sync(0xffffffffu, mem);
        if (lane_id() == 0) s_wcnt[warp_id()] = __popc(bal);
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < kDbThreads / 32; ++w) { const int c = s_wcnt[w]; if (w < warp_id()) before += c; tot += c; }
        if (mem) member[s_run + before + __popc(bal & ((1u << lane_id()) - 1))] = lst[i];
        __syncthreads();
        if (t == 0) s_run += tot;
        __syncthreads();
    }
    const int cnt = s_run > min_instance_pts ? s_run : 0;
    for (int i = t; i < cnt; i += kDbThreads) olst[i] = member[i];
    if (t == 0) out_counts[inst] = cnt;
}

// ref: ResamplePoints (data_transforms.py:247-262) on the isolated instances: object o = list row obj_inst[o].
// The draw is seevcn_resample_perm(j, reps * count, seed, instance) (common.cuh), reproducible on the host.
__global__ void __launch_bounds__(256)
resample_lists_kernel(int n_points, int stride, unsigned seed, const float* __restrict__ pts, const int* __restrict__ lists,
                      const int* __restrict__ counts, const int* __restrict__ obj_inst, float* __restrict__ out) {
    const int o = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_points) return;
    const int inst = obj_inst[o];
    const int cnt = counts[inst];
    float* dst = out + ((size_t)o * n_points + j) * 3;
    if (cnt <= 0) { dst[0] = dst[1] = dst[2] = 0.f; return; }
    const unsigned reps = (unsigned)((n_points + cnt - 1) / cnt);
    const unsigned src = feistel_perm((unsigned)j, reps * (unsigned)cnt, mix32(seed ^ mix32((unsigned)inst))) % (unsigned)cnt;
    const int p = lists[(size_t)inst * stride + src];
    dst[0] = pts[p * 3]; dst[1] = pts[p * 3 + 1]; dst[2] = pts[p * 3 + 2];
}

}  // namespace

extern "C" int seevcn_project_points(int num_points, const float* pts, const double* lidar2cam, const double* intrinsic,
                                     const double* distcoeff, int equidistant, int img_w, int img_h, int* uv,
                                     unsigned char* fov, float* depth, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_points >= 0 && img_w > 1 && img_h > 1, "project_points: bad sizes");
    if (num_points == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(pts && lidar2cam && intrinsic && distcoeff && uv && fov, "project_points: null pointer");
    CamModel cm;
    for (int i = 0; i < 12; ++i) cm.ext[i] = lidar2cam[i];
    cm.fx = intrinsic[0]; cm.cx = intrinsic[2]; cm.fy = intrinsic[4]; cm.cy = intrinsic[5];   // 3x3 row major
    for (int i = 0; i < 5; ++i) cm.d[i] = distcoeff[i];
    cm.pre_limit = atan((double)img_w / (double)img_h);
    cm.width = img_w; cm.height = img_h; cm.equidistant = equidistant ? 1 : 0;
    SEEVCN_PROF("project_points", as_stream(stream));
    project_kernel<<<div_up(num_points, 256), 256, 0, as_stream(stream)>>>(num_points, pts, cm, reinterpret_cast<int2*>(uv), fov, depth);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_points_in_masks(int num_points, int num_inst, int img_w, int img_h, const int* uv, const unsigned char* fov,
                                      const unsigned char* masks, int* lists, int* counts, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_points >= 0 && num_inst >= 0 && img_w > 0 && img_h > 0, "points_in_masks: bad sizes");
    if (num_inst == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(counts && (num_points == 0 || (uv && fov && masks && lists)), "points_in_masks: null pointer");
    SEEVCN_PROF("points_in_masks", as_stream(stream));
    mask_lookup_kernel<<<num_inst, 256, 0, as_stream(stream)>>>(num_points, img_w, img_h, reinterpret_cast<const int2*>(uv), fov, masks,
                                                               lists, counts);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_dbscan_largest(int num_inst, int stride, const float* pts, const int* lists, const int* counts,
                                     int adaptive, double eps, double vres_deg, double eps_scaling, double min_eps, double max_eps,
                                     int min_points, int min_instance_pts, int* out_lists, int* out_counts, double* out_eps,
                                     void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_inst >= 0 && stride >= 0 && min_points >= 1, "dbscan_largest: bad sizes");
    if (num_inst == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(pts && lists && counts && out_lists && out_counts, "dbscan_largest: null pointer");
    SEEVCN_REQUIRE(adaptive || eps > 0.0, "dbscan_largest: eps must be > 0");
    EpsRule rule{adaptive ? 1 : 0, eps, vres_deg, eps_scaling, min_eps, max_eps};
    const size_t smem = (size_t)std::min(stride, kDbMaxN) * 24 + 16;
    SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(dbscan_largest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kDbMaxN * 24 + 16)));
    // workspace: [bump counter, 256 B] [slices for the instances that do not fit in shared memory]
    unsigned char* gws = nullptr; unsigned long long* bump = nullptr; size_t gbytes = 0;
    if (workspace && workspace_bytes > 256) {
        bump = static_cast<unsigned long long*>(workspace);
        gws = static_cast<unsigned char*>(workspace) + 256; gbytes = workspace_bytes - 256;
        SEEVCN_CUDA_CHECK(cudaMemsetAsync(bump, 0, 8, as_stream(stream)));
    }
    SEEVCN_PROF("dbscan_largest", as_stream(stream));
    dbscan_largest_kernel<<<num_inst, kDbThreads, smem, as_stream(stream)>>>(stride, pts, lists, counts, rule, min_points,
                                                                            min_instance_pts, out_lists, out_counts, out_eps, gws,
                                                                            gbytes, bump);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_resample_lists(int num_obj, int n_points, int stride, unsigned seed, const float* pts, const int* lists,
                                     const int* counts, const int* obj_inst, float* out, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_obj >= 0 && n_points >= 0 && stride >= 0, "resample_lists: bad sizes");
    if (num_obj == 0 || n_points == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(pts && lists && counts && obj_inst && out, "resample_lists: null pointer");
    SEEVCN_REQUIRE(num_obj <= 65535, "resample_lists: more than 65535 objects per call");
    resample_lists_kernel<<<dim3(div_up(n_points, 256), num_obj), 256, 0, as_stream(stream)>>>(n_points, stride, seed, pts, lists,
                                                                                                counts, obj_inst, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
