// Stage 1 — points in rotated boxes (crop), with optional per-box compaction.
//
// Replaces detector3d/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:313-359
// (points_in_boxes_kernel / launcher) and roiaware_pool3d.cpp:121-168 (points_in_boxes_cpu).
//
// Design (B200): HBM-bound, 16 B/point (12 in + 4 out).
//  * each CTA stages a tile of 1024 points (12 KB) into shared memory with 128-bit
//    streaming loads (aligned body, scalar head/tail), then reads it back at stride 3
//    floats (conflict-free: 3 is coprime to 32 banks);
//  * per-box constants (centre, cos/sin(-heading), half-extent thresholds) are computed
//    once per CTA into shared memory instead of once per (point, box) pair as in the
//    reference (roiaware_pool3d_kernel.cu:16-20 recomputes cos/sin per pair);
//  * each thread tests 4 points per box read (box constants are warp-broadcast LDS.128);
//  * results are written coalesced.
//
// Bit-exactness with the reference kernel as nvcc 12.9 compiles it for sm_100a: the
// reference expression tree, including the FMA contraction ptxas picks
// (local_x = fma(sx, cosa, rn(sy * -sina)), local_y = fma(sy, cosa, rn(sx * sina)))
// and the double-precision comparisons, is reproduced with explicit intrinsics.  The
// double comparisons `(double)|l| < (double)d/2.0 + (double)MARGIN` are folded exactly
// into float thresholds: for float f and double h, f < h  <=>  f < round_up_to_float(h),
// and f > h <=> f > round_down_to_float(h).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kTilePts = kThreads * kPtsPerThread;   // 1024 points = 12 KB
constexpr int kBoxChunk = 256;                       // boxes staged per pass
constexpr int kMaxCropBoxes = 1024;                  // compaction path: per-warp histograms in smem

struct __align__(16) BoxConst {
    float cx, cy, cz, hz;        // centre, z half extent (rounded down)
    float cosa, nsina, tx, ty;   // cos(-rz), -sin(-rz), x/y thresholds (rounded up)
};

// ref: check_pt_in_box3d + lidar_to_local_coords, roiaware_pool3d_kernel.cu:16-36
// FUSED selects the GPU kernel's contraction; !FUSED the CPU twin's separately rounded
// products (g++ x86-64 baseline has no FMA), roiaware_pool3d.cpp:121-140.
template <bool FUSED>
__device__ __forceinline__ bool pt_in_box(const BoxConst& b, float x, float y, float z) {
    if (fabsf(__fsub_rn(z, b.cz)) > b.hz) return false;
    const float sx = __fsub_rn(x, b.cx), sy = __fsub_rn(y, b.cy);
    float lx, ly;
    if (FUSED) {
        lx = __fmaf_rn(sx, b.cosa, __fmul_rn(sy, b.nsina));
        ly = __fmaf_rn(sy, b.cosa, __fmul_rn(sx, -b.nsina));
    } else {
        lx = __fadd_rn(__fmul_rn(sx, b.cosa), __fmul_rn(sy, b.nsina));
        ly = __fadd_rn(__fmul_rn(sx, -b.nsina), __fmul_rn(sy, b.cosa));
    }
    return (fabsf(lx) < b.tx) & (fabsf(ly) < b.ty);
}

__device__ __forceinline__ void load_box_const(BoxConst& o, const float* __restrict__ box,
                                               const float* __restrict__ trig, float margin) {
    const float cx = box[0], cy = box[1], cz = box[2];
    const float dx = box[3], dy = box[4], dz = box[5], rz = box[6];
    float cosa, sina;
    if (trig) { cosa = trig[0]; sina = trig[1]; }          // host libm values (CPU-twin parity)
    else      { cosa = cosf(-rz); sina = sinf(-rz); }      // same libdevice calls as the reference
    o.cx = cx; o.cy = cy; o.cz = cz;
    o.hz = __double2float_rd((double)dz * 0.5);
    o.cosa = cosa; o.nsina = -sina;
    o.tx = __double2float_ru((double)dx * 0.5 + (double)margin);
    o.ty = __double2float_ru((double)dy * 0.5 + (double)margin);
}

// grid (ceil(P / 1024), B).  HIST: also emit per-(tile, box) counts for the compaction pass.
template <bool HIST>
__global__ void __launch_bounds__(kThreads)
points_in_boxes_kernel(int boxes_num, int pts_num, const float* __restrict__ boxes,
                       const float* __restrict__ pts, int* __restrict__ box_idx_of_points,
                       int* __restrict__ tile_counts /* (B, ntiles, T) */) {
    __shared__ __align__(16) float s_pts[kTilePts * 3 + 8];
    __shared__ BoxConst s_box[kBoxChunk];
    __shared__ float s_rxy[kBoxChunk];
    extern __shared__ int s_hist[];   // HIST: boxes_num ints

    const int b = blockIdx.y;
    const int tile = blockIdx.x;
    const int p0 = tile * kTilePts;
    const int npts = min(kTilePts, pts_num - p0);
    const float* src = pts + ((size_t)b * pts_num + p0) * 3;
    const int mis = stage_floats(s_pts, src, npts * 3);
    if (HIST) for (int k = threadIdx.x; k < boxes_num; k += kThreads) s_hist[k] = 0;

    int res[kPtsPerThread];
    float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
#pragma unroll
    for (int j = 0; j < kPtsPerThread; ++j) res[j] = -1;
    // a warp owns 128 CONSECUTIVE points (point j of lane l = warp*128 + j*32 + l): in sensor order that is a short
    // arc of one beam, so its bounding box misses almost every object box and the exact test is skipped for those
    float wlo[3], whi[3];
    const int wbase = warp_id() * (32 * kPtsPerThread) + lane_id();

    const float* fboxes = boxes + (size_t)b * boxes_num * 7;
    for (int k0 = 0; k0 < boxes_num; k0 += kBoxChunk) {
        const int nb = min(kBoxChunk, boxes_num - k0);
        __syncthreads();   // staging done (first pass) / previous chunk consumed
        if (threadIdx.x < nb) {
            BoxConst bc;
            load_box_const(bc, fboxes + (size_t)(k0 + threadIdx.x) * 7, nullptr, 1e-5f);
            s_box[threadIdx.x] = bc;
            // a point that passes the test lies within rxy of the centre in xy and hz in z (the rotation preserves
            // length up to fp32 rounding: 1e-4 relative + 1e-4 absolute slack is orders of magnitude above it)
            s_rxy[threadIdx.x] = sqrtf(bc.tx * bc.tx + bc.ty * bc.ty) * 1.0001f + 1e-4f;
        }
        __syncthreads();
        if (k0 == 0) {
            const float inf = __int_as_float(0x7f800000);
            wlo[0] = wlo[1] = wlo[2] = inf; whi[0] = whi[1] = whi[2] = -inf;
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j) {
                const int i = wbase + j * 32;
                const bool ok = i < npts;
                px[j] = ok ? s_pts[mis + 3 * i + 0] : 0.f;
                py[j] = ok ? s_pts[mis + 3 * i + 1] : 0.f;
                pz[j] = ok ? s_pts[mis + 3 * i + 2] : 0.f;
                if (ok) {
                    wlo[0] = fminf(wlo[0], px[j]); whi[0] = fmaxf(whi[0], px[j]);
                    wlo[1] = fminf(wlo[1], py[j]); whi[1] = fmaxf(whi[1], py[j]);
                    wlo[2] = fminf(wlo[2], pz[j]); whi[2] = fmaxf(whi[2], pz[j]);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    wlo[c] = fminf(wlo[c], __shfl_xor_sync(0xffffffffu, wlo[c], off));
                    whi[c] = fmaxf(whi[c], __shfl_xor_sync(0xffffffffu, whi[c], off));
                }
        }
        for (int k = 0; k < nb; ++k) {
            const BoxConst bc = s_box[k];
            const float rxy = s_rxy[k], rz = bc.hz * 1.0001f + 1e-4f;
            // warp-uniform; written so that NaN boxes are never skipped (they fall through to the exact test)
            if (wlo[0] > bc.cx + rxy || whi[0] < bc.cx - rxy || wlo[1] > bc.cy + rxy || whi[1] < bc.cy - rxy ||
                wlo[2] > bc.cz + rz || whi[2] < bc.cz - rz)
                continue;
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j)
                if (res[j] < 0 && pt_in_box<true>(bc, px[j], py[j], pz[j])) res[j] = k0 + k;
        }
    }
    int* out = box_idx_of_points + (size_t)b * pts_num + p0;
#pragma unroll
    for (int j = 0; j < kPtsPerThread; ++j) {
        const int i = wbase + j * 32;
        if (i < npts) {
            out[i] = res[j];
            if (HIST && res[j] >= 0) atomicAdd(&s_hist[res[j]], 1);
        }
    }
    if (HIST) {
        __syncthreads();
        int* tc = tile_counts + ((size_t)b * gridDim.x + tile) * boxes_num;
        for (int k = threadIdx.x; k < boxes_num; k += kThreads) tc[k] = s_hist[k];
    }
}

// Dense (T,P) 0/1 matrix with the CPU twin's MARGIN = 1e-2 (roiaware_pool3d.cpp:133).
// grid (ceil(P/1024)).  Write-bound: 4*T bytes out per point.
__global__ void __launch_bounds__(kThreads)
points_in_boxes_dense_kernel(int boxes_num, int pts_num, const float* __restrict__ boxes,
                             const float* __restrict__ trig, const float* __restrict__ pts,
                             int* __restrict__ pts_indices) {
    __shared__ __align__(16) float s_pts[kTilePts * 3 + 8];
    __shared__ BoxConst s_box[kBoxChunk];
    const int p0 = blockIdx.x * kTilePts;
    const int npts = min(kTilePts, pts_num - p0);
    const int mis = stage_floats(s_pts, pts + (size_t)p0 * 3, npts * 3);
    float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
    for (int k0 = 0; k0 < boxes_num; k0 += kBoxChunk) {
        const int nb = min(kBoxChunk, boxes_num - k0);
        __syncthreads();
        if (threadIdx.x < nb)
            load_box_const(s_box[threadIdx.x], boxes + (size_t)(k0 + threadIdx.x) * 7,
                           trig ? trig + (size_t)(k0 + threadIdx.x) * 2 : nullptr, 1e-2f);
        __syncthreads();
        if (k0 == 0) {
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j) {
                const int i = threadIdx.x + j * kThreads;
                const bool ok = i < npts;
                px[j] = ok ? s_pts[mis + 3 * i + 0] : 0.f;
                py[j] = ok ? s_pts[mis + 3 * i + 1] : 0.f;
                pz[j] = ok ? s_pts[mis + 3 * i + 2] : 0.f;
            }
        }
        for (int k = 0; k < nb; ++k) {
            const BoxConst bc = s_box[k];
            int* out = pts_indices + (size_t)(k0 + k) * pts_num + p0;
#pragma unroll
            for (int j = 0; j < kPtsPerThread; ++j) {
                const int i = threadIdx.x + j * kThreads;
                if (i < npts) out[i] = pt_in_box<false>(bc, px[j], py[j], pz[j]) ? 1 : 0;
            }
        }
    }
}

// Compaction pass B: per frame, exclusive scan of tile_counts over tiles (in place) and of
// the per-box totals over boxes.  grid B, block 256: one warp per box, lanes stride over tiles with a
// shuffle scan; then warp 0 scans the box totals.
__device__ __forceinline__ int warp_excl_scan(int v, int& total) {
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane_id() >= off) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

__global__ void __launch_bounds__(1024)
crop_scan_kernel(int boxes_num, int ntiles, int* __restrict__ tile_counts,
                 int* __restrict__ box_counts, int* __restrict__ box_offsets) {
    extern __shared__ int s_tot[];   // boxes_num
    const int b = blockIdx.x;
    const int nwarps = blockDim.x >> 5;
    int* tc = tile_counts + (size_t)b * ntiles * boxes_num;
    for (int k = warp_id(); k < boxes_num; k += nwarps) {
        int run = 0;
        for (int t0 = 0; t0 < ntiles; t0 += 256) {          // 8 independent loads in flight per lane, then the scans
            int c[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * 32 + lane_id();
                c[u] = t < ntiles ? tc[(size_t)t * boxes_num + k] : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * 32 + lane_id();
                int tot;
                const int ex = warp_excl_scan(c[u], tot);
                if (t < ntiles) tc[(size_t)t * boxes_num + k] = run + ex;
                run += tot;
            }
        }
        if (lane_id() == 0) { s_tot[k] = run; box_counts[(size_t)b * boxes_num + k] = run; }
    }
    __syncthreads();
    if (warp_id() == 0) {
        int run = 0;
        for (int k0 = 0; k0 < boxes_num; k0 += 32) {
            const int k = k0 + lane_id();
            const int c = k < boxes_num ? s_tot[k] : 0;
            int tot;
            const int ex = warp_excl_scan(c, tot);
            if (k < boxes_num) box_offsets[(size_t)b * boxes_num + k] = run + ex;
            run += tot;
        }
    }
}

// Compaction pass C: stable scatter of point indices into per-box lists.  Each warp owns a
// contiguous 128-point slice of the tile; ranks inside a 32-point row come from
// __match_any_sync (the multi-key form of ballot compaction).
__global__ void __launch_bounds__(kThreads)
crop_scatter_kernel(int boxes_num, int pts_num, const int* __restrict__ box_idx_of_points,
                    const int* __restrict__ tile_offsets, const int* __restrict__ box_offsets,
                    int* __restrict__ box_points) {
    extern __shared__ int s_w[];   // (8 warps, boxes_num) running write positions
    constexpr int kWarps = kThreads / 32;
    constexpr int kRows = kTilePts / kWarps / 32;   // 4 rows of 32 points per warp
    const int b = blockIdx.y, tile = blockIdx.x;
    const int p0 = tile * kTilePts;
    const int w = warp_id(), l = lane_id();
    for (int i = threadIdx.x; i < kWarps * boxes_num; i += kThreads) s_w[i] = 0;
    __syncthreads();
    int key[kRows];
    const int* in = box_idx_of_points + (size_t)b * pts_num;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int p = p0 + (w * kRows + r) * 32 + l;
        key[r] = p < pts_num ? in[p] : -1;
        const unsigned m = __match_any_sync(0xffffffffu, key[r]);
        if (key[r] >= 0 && l == __ffs(m) - 1) s_w[w * boxes_num + key[r]] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps + tile offset + box offset
    const int* toff = tile_offsets + ((size_t)b * gridDim.x + tile) * boxes_num;
    for (int k = threadIdx.x; k < boxes_num; k += kThreads) {
        int run = toff[k] + box_offsets[(size_t)b * boxes_num + k];
        for (int ww = 0; ww < kWarps; ++ww) {
            const int c = s_w[ww * boxes_num + k];
            s_w[ww * boxes_num + k] = run;
            run += c;
        }
    }
    __syncthreads();
    int* out = box_points + (size_t)b * pts_num;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int p = p0 + (w * kRows + r) * 32 + l;
        const unsigned m = __match_any_sync(0xffffffffu, key[r]);
        int base = 0;
        if (key[r] >= 0) base = s_w[w * boxes_num + key[r]];
        __syncwarp();
        if (key[r] >= 0) {
            out[base + __popc(m & ((1u << l) - 1))] = p;
            if (l == __ffs(m) - 1) s_w[w * boxes_num + key[r]] = base + __popc(m);
        }
        __syncwarp();
    }
}

// ref: ResamplePoints, data_transforms.py:247-262 (tile then permute-truncate; the tiled
// element j is original point j % count).  One warp per 32 output points.
__global__ void resample_gather_kernel(int num_obj, int n_points, int boxes_num, int pts_num,
                                       const float* __restrict__ pts, const int* __restrict__ box_counts,
                                       const int* __restrict__ box_offsets, const int* __restrict__ box_points,
                                       const int* __restrict__ obj_frame, const int* __restrict__ obj_box,
                                       const int* __restrict__ choice, float* __restrict__ out) {
    const int o = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_points) return;
    const int f = obj_frame[o], k = obj_box[o];
    const int cnt = box_counts[(size_t)f * boxes_num + k];
    float x = 0.f, y = 0.f, z = 0.f;
    if (cnt > 0) {
        const int off = box_offsets[(size_t)f * boxes_num + k];
        const int c = choice[(size_t)o * n_points + j] % cnt;
        const int p = box_points[(size_t)f * pts_num + off + c];
        const float* s = pts + ((size_t)f * pts_num + p) * 3;
        x = s[0]; y = s[1]; z = s[2];
    }
    float* d = out + ((size_t)o * n_points + j) * 3;
    d[0] = x; d[1] = y; d[2] = z;
}


__global__ void resample_gather_rng_kernel(int num_obj, int n_points, int boxes_num, int pts_num, unsigned seed,
                                           const float* __restrict__ pts, const int* __restrict__ box_counts,
                                           const int* __restrict__ box_offsets, const int* __restrict__ box_points,
                                           const int* __restrict__ obj_frame, const int* __restrict__ obj_box,
                                           float* __restrict__ out) {
    const int o = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_points) return;
    const int f = obj_frame[o], k = obj_box[o];
    const int cnt = box_counts[(size_t)f * boxes_num + k];
    float x = 0.f, y = 0.f, z = 0.f;
    if (cnt > 0) {
        const int off = box_offsets[(size_t)f * boxes_num + k];
        const unsigned reps = (unsigned)((n_points + cnt - 1) / cnt);
        const unsigned c = feistel_perm((unsigned)j, reps * (unsigned)cnt, mix32(seed ^ mix32((unsigned)(f * boxes_num + k))));
        const int p = box_points[(size_t)f * pts_num + off + (int)(c % (unsigned)cnt)];
        const float* s = pts + ((size_t)f * pts_num + p) * 3;
        x = s[0]; y = s[1]; z = s[2];
    }
    float* d = out + ((size_t)o * n_points + j) * 3;
    d[0] = x; d[1] = y; d[2] = z;
}

}  // namespace

// Host-visible copy of the permutation so tests / the oracle can reproduce the draw.
extern "C" unsigned seevcn_resample_perm(unsigned j, unsigned n, unsigned seed, unsigned frame_box) {
    return feistel_perm(j, n, mix32(seed ^ mix32(frame_box)));
}

extern "C" int seevcn_resample_gather_rng(int num_obj, int n_points, int boxes_num, int pts_num, unsigned seed,
                                          const float* pts, const int* box_counts, const int* box_offsets,
                                          const int* box_points, const int* obj_frame, const int* obj_box, float* out,
                                          seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_obj >= 0 && n_points >= 0, "resample_gather_rng: negative size");
    if (num_obj == 0 || n_points == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(num_obj <= 65535, "resample_gather_rng: num_obj > 65535 per call");
    SEEVCN_REQUIRE(pts && box_counts && box_offsets && box_points && obj_frame && obj_box && out,
                   "resample_gather_rng: null pointer");
    dim3 grid(div_up(n_points, 256), num_obj);
    resample_gather_rng_kernel<<<grid, 256, 0, as_stream(stream)>>>(num_obj, n_points, boxes_num, pts_num, seed, pts,
                                                                    box_counts, box_offsets, box_points, obj_frame,
                                                                    obj_box, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_points_in_boxes(int batch_size, int boxes_num, int pts_num, const float* boxes,
                                      const float* pts, int* box_idx_of_points, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(batch_size >= 0 && boxes_num >= 0 && pts_num >= 0, "points_in_boxes: negative size");
    if (batch_size == 0 || pts_num == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(pts && box_idx_of_points && (boxes || boxes_num == 0), "points_in_boxes: null pointer");
    SEEVCN_REQUIRE(batch_size <= 65535, "points_in_boxes: batch_size > 65535");
    dim3 grid(div_up(pts_num, kTilePts), batch_size);
    points_in_boxes_kernel<false><<<grid, kThreads, 0, as_stream(stream)>>>(boxes_num, pts_num, boxes, pts,
                                                                            box_idx_of_points, nullptr);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

// trig: optional (T,2) device array of host-computed {cosf(-rz), sinf(-rz)} so the result
// is bit-identical to the reference's CPU build (glibc cosf/sinf); exported separately.
extern "C" int seevcn_points_in_boxes_dense_trig(int boxes_num, int pts_num, const float* boxes,
                                                 const float* trig, const float* pts, int* pts_indices,
                                                 seevcn_stream_t stream) {
    SEEVCN_REQUIRE(boxes_num >= 0 && pts_num >= 0, "points_in_boxes_dense: negative size");
    if (boxes_num == 0 || pts_num == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(boxes && pts && pts_indices, "points_in_boxes_dense: null pointer");
    points_in_boxes_dense_kernel<<<div_up(pts_num, kTilePts), kThreads, 0, as_stream(stream)>>>(
        boxes_num, pts_num, boxes, trig, pts, pts_indices);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_points_in_boxes_dense(int boxes_num, int pts_num, const float* boxes, const float* pts,
                                            int* pts_indices, seevcn_stream_t stream) {
    return seevcn_points_in_boxes_dense_trig(boxes_num, pts_num, boxes, nullptr, pts, pts_indices, stream);
}

extern "C" size_t seevcn_crop_workspace_bytes(int batch_size, int boxes_num, int pts_num) {
    if (batch_size <= 0 || boxes_num <= 0 || pts_num <= 0) return 16;
    return align_up((size_t)batch_size * div_up(pts_num, kTilePts) * boxes_num * sizeof(int), 256);
}

extern "C" int seevcn_crop_points_in_boxes(int batch_size, int boxes_num, int pts_num, const float* boxes,
                                           const float* pts, int* box_idx_of_points, int* box_counts,
                                           int* box_offsets, int* box_points, void* workspace,
                                           size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(batch_size >= 0 && boxes_num >= 0 && pts_num >= 0, "crop: negative size");
    if (batch_size == 0 || boxes_num == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(boxes_num <= kMaxCropBoxes, "crop: boxes_num %d > %d", boxes_num, kMaxCropBoxes);
    SEEVCN_REQUIRE(batch_size <= 65535, "crop: batch_size > 65535");
    SEEVCN_REQUIRE(box_counts && box_offsets, "crop: null pointer");
    cudaStream_t st = as_stream(stream);
    if (pts_num == 0) {
        SEEVCN_CUDA_CHECK(cudaMemsetAsync(box_counts, 0, (size_t)batch_size * boxes_num * sizeof(int), st));
        SEEVCN_CUDA_CHECK(cudaMemsetAsync(box_offsets, 0, (size_t)batch_size * boxes_num * sizeof(int), st));
        return SEEVCN_OK;
    }
    SEEVCN_REQUIRE(boxes && pts && box_idx_of_points && box_points && workspace, "crop: null pointer");
    if (workspace_bytes < seevcn_crop_workspace_bytes(batch_size, boxes_num, pts_num)) {
        seevcn_set_error("crop: workspace too small");
        return SEEVCN_E_WORKSPACE;
    }
    int* tile_counts = static_cast<int*>(workspace);
    const int ntiles = div_up(pts_num, kTilePts);
    dim3 grid(ntiles, batch_size);
    SEEVCN_PROF("crop", st);
    {
        SEEVCN_PROF("points_in_boxes_kernel", st);
        points_in_boxes_kernel<true><<<grid, kThreads, boxes_num * sizeof(int), st>>>(
            boxes_num, pts_num, boxes, pts, box_idx_of_points, tile_counts);
    }
    SEEVCN_LAUNCH_CHECK();
    crop_scan_kernel<<<batch_size, 1024, boxes_num * sizeof(int), st>>>(boxes_num, ntiles, tile_counts,
                                                                       box_counts, box_offsets);
    SEEVCN_LAUNCH_CHECK();
    crop_scatter_kernel<<<grid, kThreads, (kThreads / 32) * boxes_num * sizeof(int), st>>>(
        boxes_num, pts_num, box_idx_of_points, tile_counts, box_offsets, box_points);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

namespace {
// single CTA: ordered compaction of the (frame, box) pairs with enough points
__global__ void __launch_bounds__(1024)
select_objects_kernel(int n, int boxes_num, const int* __restrict__ counts, int min_pts, int* __restrict__ obj_frame,
                      int* __restrict__ obj_box, int* __restrict__ num_obj) {
    __shared__ int s_warp[33];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool ok = i < n && counts[i] >= min_pts;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane_id() == 0) s_warp[warp_id()] = __popc(m);
        __syncthreads();
        if (threadIdx.x < 32) {
            const int c = s_warp[threadIdx.x];
            int inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, off); if ((int)threadIdx.x >= off) inc += t; }
            s_warp[threadIdx.x] = inc - c;
            if (threadIdx.x == 31) s_warp[32] = inc;
        }
        __syncthreads();
        if (ok) {
            const int o = s_base + s_warp[warp_id()] + __popc(m & ((1u << lane_id()) - 1));
            obj_frame[o] = i / boxes_num;
            obj_box[o] = i % boxes_num;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *num_obj = s_base;
}
}  // namespace

extern "C" int seevcn_select_objects(int batch, int boxes_num, const int* box_counts, int min_pts, int* obj_frame,
                                     int* obj_box, int* num_obj, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(batch >= 0 && boxes_num >= 0, "select_objects: negative size");
    SEEVCN_REQUIRE((long long)batch * boxes_num < (1ll << 31), "select_objects: too many boxes");
    SEEVCN_REQUIRE(num_obj, "select_objects: null pointer");
    const int n = batch * boxes_num;
    SEEVCN_REQUIRE(n == 0 || (box_counts && obj_frame && obj_box), "select_objects: null pointer");
    select_objects_kernel<<<1, 1024, 0, as_stream(stream)>>>(n, boxes_num > 0 ? boxes_num : 1, box_counts, min_pts, obj_frame,
                                                             obj_box, num_obj);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_resample_gather(int num_obj, int n_points, int boxes_num, int pts_num, const float* pts,
                                      const int* box_counts, const int* box_offsets, const int* box_points,
                                      const int* obj_frame, const int* obj_box, const int* choice, float* out,
                                      seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_obj >= 0 && n_points >= 0, "resample_gather: negative size");
    if (num_obj == 0 || n_points == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(num_obj <= 65535, "resample_gather: num_obj > 65535 per call");
    SEEVCN_REQUIRE(pts && box_counts && box_offsets && box_points && obj_frame && obj_box && choice && out,
                   "resample_gather: null pointer");
    dim3 grid(div_up(n_points, 256), num_obj);
    resample_gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(num_obj, n_points, boxes_num, pts_num, pts,
                                                                box_counts, box_offsets, box_points, obj_frame,
                                                                obj_box, choice, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
