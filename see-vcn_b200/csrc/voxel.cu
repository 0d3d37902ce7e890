// Stage 6 — voxelization: MeanVFE, dynamic (hashed scatter-mean) and hard (spconv-style).
//
// Replaces
//   MeanVFE.forward          detector3d/pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31
//   DynamicMeanVFE.forward   detector3d/pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:37-76
//     (torch.unique = device-wide sort + torch_scatter.scatter_mean + ~15 elementwise launches)
//   VoxelGeneratorWrapper    detector3d/pcdet/datasets/processor/data_processor.py:15-60
//     (single-thread CPU loop inside third-party spconv)
//
// Dynamic voxelization here is sort-free: one pass inserts every in-range point into an
// open-addressing hash table keyed by the reference's merge key
// b*XYZ + x*YZ + y*Z + z (64-bit) and accumulates [sum features, count] with one vector
// reduction per point; a second pass compacts the occupied slots into
// voxel_coords / voxel_features / voxel_counts.  HBM-bound: 4*(1+C) B/point in,
// (16 + 4*C + 4) B/voxel out.
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace {

constexpr unsigned long long kEmpty = ~0ull;
constexpr int kMaxFeat = 8;   // point features (xyz + up to 5 extras) held in registers

struct VoxGeom {
    float lo[3], vs[3];
    int g[3];
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

// ref: dynamic_mean_vfe.py:53-54 — floor((xyz - range_lo) / voxel_size).int(), fp32 IEEE ops
__device__ __forceinline__ bool voxel_coord(const VoxGeom& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    const float fx = floorf(__fdiv_rn(__fsub_rn(x, g.lo[0]), g.vs[0]));
    const float fy = floorf(__fdiv_rn(__fsub_rn(y, g.lo[1]), g.vs[1]));
    const float fz = floorf(__fdiv_rn(__fsub_rn(z, g.lo[2]), g.vs[2]));
    // comparisons on the float value avoid undefined float->int conversions for far-away points
    if (!(fx >= 0.f && fx < (float)g.g[0] && fy >= 0.f && fy < (float)g.g[1] && fz >= 0.f && fz < (float)g.g[2]))
        return false;
    cx = (int)fx; cy = (int)fy; cz = (int)fz;
    return true;
}

// acc layout per slot: [sum_0 .. sum_{C-1}, count] padded to ACCW floats (4 or 8) so one
// slot is one 16/32-byte sector and C=3 uses a single red.global.add.v4.f32.
template <int ACCW>
__device__ __forceinline__ void dynvox_add(const VoxGeom& g, long long b, const float (&f)[kMaxFeat], unsigned long long nslots,
                                           unsigned long long* __restrict__ keys, float* __restrict__ acc) {
    int cx, cy, cz;
    if (!voxel_coord(g, f[0], f[1], f[2], cx, cy, cz)) return;
    const unsigned long long key =
        (unsigned long long)(((b * g.g[0] + cx) * g.g[1] + cy) * (long long)g.g[2] + cz);
    unsigned long long slot = __umul64hi(mix64(key), nslots);   // fast range reduction: any table size, no modulo
    while (true) {
        const unsigned long long prev = atomicCAS(&keys[slot], kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        if (++slot == nslots) slot = 0;
    }
    float* a = acc + slot * ACCW;
    if (ACCW == 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(1.f) : "memory");
    } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a + 4), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(1.f) : "memory");
    }
}

template <int ACCW>
__global__ void __launch_bounds__(256)
dynvox_insert_kernel(int n, int c, const float* __restrict__ points, VoxGeom g, unsigned long long nslots,
                     unsigned long long* __restrict__ keys, float* __restrict__ acc) {
    const int stride = 1 + c;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float* row = points + (size_t)p * stride;
        float f[kMaxFeat];
        const float bf = row[0];
#pragma unroll
        for (int j = 0; j < kMaxFeat; ++j) f[j] = j < c ? row[1 + j] : 0.f;
        dynvox_add<ACCW>(g, (long long)(int)bf /* points[:,0].int() */, f, nslots, keys, acc);
    }
}

// The frame pipeline's rows without the concatenated (N,4) matrix: frame f's raw points carry batch index f,
// object o's completed points carry obj_frame[o].
__global__ void __launch_bounds__(256)
dynvox_insert_frames_kernel(int n_frame_pts, int pts_per_frame, const float* __restrict__ frame_pts,
                            const unsigned char* __restrict__ frame_keep, int n_obj_pts,
                            int pts_per_obj, const float* __restrict__ obj_pts, const int* __restrict__ obj_frame,
                            const int* __restrict__ obj_count,
                            VoxGeom g, unsigned long long nslots, unsigned long long* __restrict__ keys,
                            float* __restrict__ acc) {
    const int n = n_frame_pts + n_obj_pts;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const float* row;
        long long b;
        if (p < n_frame_pts) {
            if (frame_keep && !frame_keep[p]) continue;          // replaced by a completed cloud (splice step)
            row = frame_pts + (size_t)p * 3; b = p / pts_per_frame;
        } else {
            const int q = p - n_frame_pts, o = q / pts_per_obj;
            if (obj_count && q - o * pts_per_obj >= obj_count[o]) continue;   // cyclic repeats of the object's distinct rows
            row = obj_pts + (size_t)q * 3; b = obj_frame[o];
        }
        float f[kMaxFeat];
#pragma unroll
        for (int j = 0; j < kMaxFeat; ++j) f[j] = j < 3 ? row[j] : 0.f;
        dynvox_add<4>(g, b, f, nslots, keys, acc);
    }
}

// Compacts occupied slots.  Row order = slot order within a warp-aggregated claim.
template <int ACCW>
__global__ void __launch_bounds__(256)
dynvox_finalize_kernel(unsigned long long nslots, int c, VoxGeom g, const unsigned long long* __restrict__ keys,
                       const float* __restrict__ acc, int max_voxels, int* __restrict__ voxel_coords,
                       float* __restrict__ voxel_features, int* __restrict__ voxel_counts,
                       unsigned long long* __restrict__ out_keys, int key32, int* __restrict__ num_voxels) {
    __shared__ int s_wcnt[8];
    __shared__ int s_base;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (nslots + nthreads - 1) / nthreads;
    for (unsigned long long it = 0; it < rounds; ++it) {
        const unsigned long long s = it * nthreads + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        unsigned long long key = kEmpty;
        if (s < nslots) key = keys[s];
        const bool occ = key != kEmpty;
        // one global atomic per CTA per round: warp ballots -> smem prefix over the 8 warps
        const unsigned m = __ballot_sync(0xffffffffu, occ);
        __syncthreads();                       // previous round's s_base / s_wcnt consumed
        if (lane_id() == 0) s_wcnt[warp_id()] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) { const int c = s_wcnt[w]; s_wcnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(num_voxels, tot) : 0;
        }
        __syncthreads();
        if (!occ) continue;
        const int row = s_base + s_wcnt[warp_id()] + __popc(m & ((1u << lane_id()) - 1));
        if (row >= max_voxels) continue;
        const float* a = acc + s * ACCW;
        const float cnt = a[ACCW - 1];
        const long long z = (long long)(key % (unsigned long long)g.g[2]);
        const long long y = (long long)((key / (unsigned long long)g.g[2]) % (unsigned long long)g.g[1]);
        const long long x = (long long)((key / ((unsigned long long)g.g[2] * g.g[1])) % (unsigned long long)g.g[0]);
        const long long b = (long long)(key / ((unsigned long long)g.g[2] * g.g[1] * g.g[0]));
        reinterpret_cast<int4*>(voxel_coords)[row] = make_int4((int)b, (int)z, (int)y, (int)x);   // [b,z,y,x]
        for (int j = 0; j < c; ++j) voxel_features[(size_t)row * c + j] = __fdiv_rn(a[j], cnt);
        voxel_counts[row] = (int)cnt;
        if (out_keys) {   // sort keys: 32-bit when the whole grid x batch fits (halves the radix sort's key traffic)
            if (key32) reinterpret_cast<unsigned*>(out_keys)[row] = (unsigned)key;
            else out_keys[row] = key;
        }
    }
}

__global__ void iota_kernel(int n, int* v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void dynvox_permute_kernel(int max_voxels, int c, const int* __restrict__ num_voxels,
                                      const int* __restrict__ order, const int4* __restrict__ coords_in,
                                      const float* __restrict__ feat_in, const int* __restrict__ cnt_in,
                                      int4* __restrict__ coords_out, float* __restrict__ feat_out,
                                      int* __restrict__ cnt_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = min(*num_voxels, max_voxels);
    if (i >= m) return;
    const int s = order[i];
    coords_out[i] = coords_in[s];
    for (int j = 0; j < c; ++j) feat_out[(size_t)i * c + j] = feat_in[(size_t)s * c + j];
    cnt_out[i] = cnt_in[s];
}

// ref: mean_vfe.py:23-29.  One thread per (voxel, feature).
__global__ void __launch_bounds__(256)
mean_vfe_kernel(int m, int t, int c, const float* __restrict__ voxels, const float* __restrict__ num_points,
                float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)m * c) return;
    const int v = (int)(e / c), j = (int)(e - (long long)v * c);
    const float* src = voxels + (size_t)v * t * c + j;
    float s = 0.f;
    for (int i = 0; i < t; ++i) s = __fadd_rn(s, src[(size_t)i * c]);   // torch.sum over dim 1, sequential for t<=~32
    const float norm = fmaxf(num_points[v], 1.0f);                     // clamp_min(1.0)
    out[e] = __fdiv_rn(s, norm);
}


// ------------------------------------------------------------------ hard voxelization --
// spconv semantics without its serial loop:
//  1. hash-insert every in-grid point; first[slot] = atomicMin(point index)  -> the point that opens the voxel
//  2. exclusive scan of "opens a voxel" flags over points                     -> voxel id in first-seen order
//  3. ids >= max_voxels are dropped (their points are skipped, like the reference loop's `continue`)
//  4. every point runs a T-stage atomicMin chain on its voxel's slot list: stage s keeps the
//     minimum it has seen and forwards the loser, so the list ends as the T smallest point
//     indices in ascending order, independent of execution order (deterministic)
//  5. gather rows.
__global__ void __launch_bounds__(256)
hardvox_insert_kernel(int n, int c, const float* __restrict__ points, VoxGeom g, unsigned long long hmask,
                      unsigned long long* __restrict__ keys, int* __restrict__ first, int* __restrict__ slot_of_point) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float* row = points + (size_t)p * c;
    int cx, cy, cz;
    if (!voxel_coord(g, row[0], row[1], row[2], cx, cy, cz)) { slot_of_point[p] = -1; return; }
    const unsigned long long key = (unsigned long long)(((long long)cx * g.g[1] + cy) * (long long)g.g[2] + cz);
    unsigned long long slot = mix64(key) & hmask;
    while (true) {
        const unsigned long long prev = atomicCAS(&keys[slot], kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        slot = (slot + 1) & hmask;
    }
    atomicMin(&first[slot], p);
    slot_of_point[p] = (int)slot;
}

__global__ void __launch_bounds__(256)
hardvox_flag_kernel(int n, const int* __restrict__ first, const int* __restrict__ slot_of_point, int* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = slot_of_point[p];
    flag[p] = (s >= 0 && first[s] == p) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
hardvox_assign_kernel(int n, int max_voxels, VoxGeom g, const unsigned long long* __restrict__ keys,
                      const int* __restrict__ slot_of_point, const int* __restrict__ flag, const int* __restrict__ rank,
                      int* __restrict__ vid_of_slot, int* __restrict__ coordinates, int* __restrict__ num_voxels) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (p == n - 1) *num_voxels = min(rank[p] + flag[p], max_voxels);
    if (!flag[p]) return;
    const int s = slot_of_point[p];
    const int vid = rank[p];
    if (vid >= max_voxels) return;   // vid_of_slot stays -1
    vid_of_slot[s] = vid;
    const unsigned long long key = keys[s];
    const int z = (int)(key % (unsigned long long)g.g[2]);
    const int y = (int)((key / (unsigned long long)g.g[2]) % (unsigned long long)g.g[1]);
    const int x = (int)(key / ((unsigned long long)g.g[2] * g.g[1]));
    coordinates[vid * 3 + 0] = z; coordinates[vid * 3 + 1] = y; coordinates[vid * 3 + 2] = x;
}

__global__ void __launch_bounds__(256)
hardvox_pick_kernel(int n, int t, const int* __restrict__ slot_of_point, const int* __restrict__ vid_of_slot,
                    int* __restrict__ sel, int* __restrict__ cnt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = slot_of_point[p];
    if (s < 0) return;
    const int vid = vid_of_slot[s];
    if (vid < 0) return;
    atomicAdd(&cnt[vid], 1);
    int v = p;
    for (int j = 0; j < t; ++j) {
        const int old = atomicMin(&sel[(size_t)vid * t + j], v);
        v = max(old, v);
        if (v >= 0x7f000000) break;   // forwarding the empty sentinel
    }
}

__global__ void __launch_bounds__(256)
hardvox_gather_kernel(int max_voxels, int t, int c, const float* __restrict__ points, const int* __restrict__ num_voxels,
                      const int* __restrict__ sel, const int* __restrict__ cnt, float* __restrict__ voxels,
                      int* __restrict__ num_points_per_voxel) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = *num_voxels;
    if (e >= (long long)m * t) return;
    const int vid = (int)(e / t), j = (int)(e - (long long)vid * t);
    const int k = min(cnt[vid], t);
    if (j == 0) num_points_per_voxel[vid] = k;
    float* dst = voxels + (size_t)e * c;
    if (j < k) {
        const float* src = points + (size_t)sel[e] * c;
        for (int f = 0; f < c; ++f) dst[f] = src[f];
    } else {
        for (int f = 0; f < c; ++f) dst[f] = 0.f;
    }
}

struct HardWs {
    unsigned long long nslots;
    size_t off_keys, off_first, off_vid, off_slot, off_flag, off_rank, off_sel, off_cnt, off_cub, cub_bytes, total;
};

HardWs hard_layout(int n, int t, int max_voxels) {
    HardWs w{};
    unsigned long long h = 1024;
    while (h < 2ull * (unsigned long long)(n > 0 ? n : 1)) h <<= 1;
    w.nslots = h;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t np = (size_t)(n > 0 ? n : 1), mv = (size_t)(max_voxels > 0 ? max_voxels : 1), tt = (size_t)(t > 0 ? t : 1);
    w.off_keys = take(h * 8);
    w.off_first = take(h * 4);
    w.off_vid = take(h * 4);
    w.off_slot = take(np * 4);
    w.off_flag = take(np * 4);
    w.off_rank = take(np * 4);
    w.off_sel = take(mv * tt * 4);
    w.off_cnt = take(mv * 4);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, (int)np);
    size_t cub_bytes32 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes32, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)mv);
    if (cub_bytes32 > cub_bytes) cub_bytes = cub_bytes32;
    w.cub_bytes = cub_bytes;
    w.off_cub = take(cub_bytes);
    w.total = o;
    return w;
}

struct DynWs {
    unsigned long long nslots;
    int accw;
    size_t off_keys, off_acc, off_tmp_coords, off_tmp_feat, off_tmp_cnt, off_sort_keys_in, off_sort_keys_out,
        off_order_in, off_order_out, off_cub, cub_bytes, total;
};

DynWs dyn_layout(int n, int c, int max_voxels) {
    DynWs w{};
    // 1.5 slots per point (load factor <= 2/3 even if every point opens its own voxel; ~0.2 on LiDAR frames).  Not a power
    // of two: the slot comes from a multiply-high, so the table, its memsets and the finalize scan are not rounded up 2x.
    unsigned long long h = ((unsigned long long)(n > 0 ? n : 1) * 3 / 2 + 1024 + 255) & ~255ull;
    w.nslots = h;
    w.accw = (c + 1 <= 4) ? 4 : 8;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    w.off_keys = take(h * 8);
    w.off_acc = take(h * w.accw * 4);
    const size_t mv = (size_t)(max_voxels > 0 ? max_voxels : 1);
    w.off_tmp_coords = take(mv * 16);
    w.off_tmp_feat = take(mv * c * 4);
    w.off_tmp_cnt = take(mv * 4);
    w.off_sort_keys_in = take(mv * 8);
    w.off_sort_keys_out = take(mv * 8);
    w.off_order_in = take(mv * 4);
    w.off_order_out = take(mv * 4);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)mv);
    size_t cub_bytes32 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes32, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)mv);
    if (cub_bytes32 > cub_bytes) cub_bytes = cub_bytes32;
    w.cub_bytes = cub_bytes;
    w.off_cub = take(cub_bytes);
    w.total = o;
    return w;
}

}  // namespace

extern "C" int seevcn_mean_vfe(int num_voxels, int max_points, int num_features, const float* voxels,
                               const float* voxel_num_points, float* voxel_features, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_voxels >= 0 && max_points >= 0 && num_features >= 0, "mean_vfe: negative size");
    if (num_voxels == 0 || num_features == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(voxels && voxel_num_points && voxel_features, "mean_vfe: null pointer");
    const long long total = (long long)num_voxels * num_features;
    mean_vfe_kernel<<<(unsigned)div_up(total, 256ll), 256, 0, as_stream(stream)>>>(
        num_voxels, max_points, num_features, voxels, voxel_num_points, voxel_features);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" size_t seevcn_dynamic_voxelize_workspace_bytes(int num_points, int num_features, int max_voxels) {
    return dyn_layout(num_points, num_features, max_voxels).total;
}

namespace {
struct DynSrc {   // either one (N,1+C) matrix, or frames + object clouds
    const float* points;
    int n_frame_pts, pts_per_frame; const float* frame_pts; const unsigned char* frame_keep;
    int n_obj_pts, pts_per_obj; const float* obj_pts; const int* obj_frame; const int* obj_count;
};

int dynvox_run(const char* what, int num_points, int num_features, const DynSrc& src, const float* pc_range,
               const float* voxel_size, const int* grid_size, int max_voxels, int sorted, int batch_hint,
               int* voxel_coords, float* voxel_features, int* voxel_counts, int* num_voxels, void* workspace,
               size_t workspace_bytes, cudaStream_t st) {
    SEEVCN_PROF("dynamic_voxelize", st);
    const DynWs w = dyn_layout(num_points, num_features, max_voxels);
    if (workspace_bytes < w.total) {
        seevcn_set_error("%s: workspace %zu < %zu", what, workspace_bytes, w.total);
        return SEEVCN_E_WORKSPACE;
    }
    VoxGeom g;
    for (int i = 0; i < 3; ++i) {
        g.lo[i] = pc_range[i]; g.vs[i] = voxel_size[i]; g.g[i] = grid_size[i];
        SEEVCN_REQUIRE(grid_size[i] > 0 && voxel_size[i] > 0.f, "%s: bad grid", what);
    }
    char* ws = static_cast<char*>(workspace);
    auto* keys = reinterpret_cast<unsigned long long*>(ws + w.off_keys);
    auto* acc = reinterpret_cast<float*>(ws + w.off_acc);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(keys, 0xff, w.nslots * 8, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(acc, 0, w.nslots * w.accw * 4, st));
    const int grid_ins = (int)std::min<long long>(div_up((long long)num_points, 256ll), SEEVCN_NUM_SMS * 16);
    const int grid_fin = (int)std::min<unsigned long long>(div_up(w.nslots, 256ull), SEEVCN_NUM_SMS * 16);
    int* o_coords = voxel_coords; float* o_feat = voxel_features; int* o_cnt = voxel_counts;
    unsigned long long* o_keys = nullptr;
    if (sorted) {
        o_coords = reinterpret_cast<int*>(ws + w.off_tmp_coords);
        o_feat = reinterpret_cast<float*>(ws + w.off_tmp_feat);
        o_cnt = reinterpret_cast<int*>(ws + w.off_tmp_cnt);
        o_keys = reinterpret_cast<unsigned long long*>(ws + w.off_sort_keys_in);
        SEEVCN_CUDA_CHECK(cudaMemsetAsync(o_keys, 0xff, (size_t)max_voxels * 8, st));
    }
    const double span = (double)batch_hint * g.g[0] * g.g[1] * g.g[2];   // keys are < span
    const int key32 = sorted && span <= 4294967295.0 ? 1 : 0;
    if (!src.points) {
        dynvox_insert_frames_kernel<<<grid_ins, 256, 0, st>>>(src.n_frame_pts, src.pts_per_frame, src.frame_pts, src.frame_keep,
                                                             src.n_obj_pts, src.pts_per_obj, src.obj_pts, src.obj_frame, src.obj_count,
                                                             g, w.nslots, keys, acc);
        SEEVCN_LAUNCH_CHECK();
        dynvox_finalize_kernel<4><<<grid_fin, 256, 0, st>>>(w.nslots, num_features, g, keys, acc, max_voxels, o_coords,
                                                           o_feat, o_cnt, o_keys, key32, num_voxels);
    } else if (w.accw == 4) {
        dynvox_insert_kernel<4><<<grid_ins, 256, 0, st>>>(num_points, num_features, src.points, g, w.nslots, keys, acc);
        SEEVCN_LAUNCH_CHECK();
        dynvox_finalize_kernel<4><<<grid_fin, 256, 0, st>>>(w.nslots, num_features, g, keys, acc, max_voxels, o_coords,
                                                           o_feat, o_cnt, o_keys, key32, num_voxels);
    } else {
        dynvox_insert_kernel<8><<<grid_ins, 256, 0, st>>>(num_points, num_features, src.points, g, w.nslots, keys, acc);
        SEEVCN_LAUNCH_CHECK();
        dynvox_finalize_kernel<8><<<grid_fin, 256, 0, st>>>(w.nslots, num_features, g, keys, acc, max_voxels, o_coords,
                                                           o_feat, o_cnt, o_keys, key32, num_voxels);
    }
    SEEVCN_LAUNCH_CHECK();
    if (sorted) {
        auto* keys_out = reinterpret_cast<unsigned long long*>(ws + w.off_sort_keys_out);
        int* order_in = reinterpret_cast<int*>(ws + w.off_order_in);
        int* order_out = reinterpret_cast<int*>(ws + w.off_order_out);
        iota_kernel<<<div_up(max_voxels, 256), 256, 0, st>>>(max_voxels, order_in);
        SEEVCN_LAUNCH_CHECK();
        size_t cub_bytes = w.cub_bytes;
        // keys are < 2^kb except the all-ones padding of unused rows, which any bit range keeps last
        int kb = 1;
        while (kb < 64 && (double)(1ull << kb) < span) ++kb;
        if (key32) {
            SEEVCN_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, reinterpret_cast<const unsigned*>(o_keys),
                                                              reinterpret_cast<unsigned*>(keys_out), order_in, order_out,
                                                              max_voxels, 0, kb < 32 ? kb + 1 : 32, st));
        } else {
            SEEVCN_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + w.off_cub, cub_bytes, o_keys, keys_out, order_in,
                                                              order_out, max_voxels, 0, kb < 64 ? kb + 1 : 64, st));
        }
        dynvox_permute_kernel<<<div_up(max_voxels, 256), 256, 0, st>>>(
            max_voxels, num_features, num_voxels, order_out, reinterpret_cast<const int4*>(o_coords), o_feat, o_cnt,
            reinterpret_cast<int4*>(voxel_coords), voxel_features, voxel_counts);
        SEEVCN_LAUNCH_CHECK();
    }
    return SEEVCN_OK;
}
}  // namespace

extern "C" int seevcn_dynamic_voxelize(int num_points, int num_features, const float* points, const float* pc_range,
                                       const float* voxel_size, const int* grid_size, int max_voxels, int sorted,
                                       int batch_hint, int* voxel_coords, float* voxel_features, int* voxel_counts, int* num_voxels,
                                       void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_points >= 0 && max_voxels >= 0, "dynamic_voxelize: negative size");
    if (batch_hint <= 0) batch_hint = 1 << 20;   // unknown batch size: sort on (almost) all key bits
    SEEVCN_REQUIRE(num_features >= 3 && num_features < kMaxFeat, "dynamic_voxelize: num_features=%d outside [3,%d]",
                   num_features, kMaxFeat - 1);
    SEEVCN_REQUIRE(pc_range && voxel_size && grid_size && num_voxels, "dynamic_voxelize: null pointer");
    cudaStream_t st = as_stream(stream);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(num_voxels, 0, sizeof(int), st));
    if (num_points == 0 || max_voxels == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && voxel_coords && voxel_features && voxel_counts && workspace,
                   "dynamic_voxelize: null pointer");
    DynSrc src{};
    src.points = points;
    return dynvox_run("dynamic_voxelize", num_points, num_features, src, pc_range, voxel_size, grid_size, max_voxels, sorted,
                      batch_hint, voxel_coords, voxel_features, voxel_counts, num_voxels, workspace, workspace_bytes, st);
}

extern "C" int seevcn_dynamic_voxelize_spliced(int num_frames, int pts_per_frame, const float* frame_pts,
                                               const unsigned char* frame_keep, int num_obj,
                                               int pts_per_obj, const float* obj_pts, const int* obj_frame, const int* obj_count,
                                              const float* pc_range, const float* voxel_size, const int* grid_size,
                                              int max_voxels, int sorted, int* voxel_coords, float* voxel_features,
                                              int* voxel_counts, int* num_voxels, void* workspace, size_t workspace_bytes,
                                              seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_frames >= 0 && pts_per_frame >= 0 && num_obj >= 0 && pts_per_obj >= 0 && max_voxels >= 0,
                   "dynamic_voxelize_frames: negative size");
    const long long n_frame = (long long)num_frames * pts_per_frame, n_obj = (long long)num_obj * pts_per_obj;
    SEEVCN_REQUIRE(n_frame + n_obj < (1ll << 31), "dynamic_voxelize_frames: more than 2^31 points");
    SEEVCN_REQUIRE(pc_range && voxel_size && grid_size && num_voxels, "dynamic_voxelize_frames: null pointer");
    cudaStream_t st = as_stream(stream);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(num_voxels, 0, sizeof(int), st));
    if (n_frame + n_obj == 0 || max_voxels == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE((n_frame == 0 || frame_pts) && (n_obj == 0 || (obj_pts && obj_frame)) && voxel_coords && voxel_features &&
                   voxel_counts && workspace, "dynamic_voxelize_frames: null pointer");
    DynSrc src{};
    src.n_frame_pts = (int)n_frame; src.pts_per_frame = pts_per_frame > 0 ? pts_per_frame : 1; src.frame_pts = frame_pts; src.frame_keep = frame_keep;
    src.n_obj_pts = (int)n_obj; src.pts_per_obj = pts_per_obj > 0 ? pts_per_obj : 1; src.obj_pts = obj_pts; src.obj_frame = obj_frame; src.obj_count = obj_count;
    return dynvox_run("dynamic_voxelize_frames", (int)(n_frame + n_obj), 3, src, pc_range, voxel_size, grid_size, max_voxels,
                      sorted, num_frames > 0 ? num_frames : 1, voxel_coords, voxel_features, voxel_counts, num_voxels, workspace,
                      workspace_bytes, st);
}

extern "C" int seevcn_dynamic_voxelize_frames(int num_frames, int pts_per_frame, const float* frame_pts, int num_obj,
                                              int pts_per_obj, const float* obj_pts, const int* obj_frame,
                                              const float* pc_range, const float* voxel_size, const int* grid_size,
                                              int max_voxels, int sorted, int* voxel_coords, float* voxel_features,
                                              int* voxel_counts, int* num_voxels, void* workspace, size_t workspace_bytes,
                                              seevcn_stream_t stream) {
    return seevcn_dynamic_voxelize_spliced(num_frames, pts_per_frame, frame_pts, nullptr, num_obj, pts_per_obj, obj_pts, obj_frame,
                                           nullptr, pc_range, voxel_size, grid_size, max_voxels, sorted, voxel_coords,
                                           voxel_features, voxel_counts, num_voxels, workspace, workspace_bytes, stream);
}

extern "C" size_t seevcn_hard_voxelize_workspace_bytes(int num_points, int max_points, int max_voxels) {
    return hard_layout(num_points, max_points, max_voxels).total;
}

extern "C" int seevcn_hard_voxelize(int num_points, int num_features, const float* points, const float* pc_range,
                                    const float* voxel_size, const int* grid_size, int max_points, int max_voxels,
                                    float* voxels, int* coordinates, int* num_points_per_voxel, int* num_voxels,
                                    void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_points >= 0 && max_points >= 1 && max_voxels >= 0, "hard_voxelize: bad sizes");
    SEEVCN_REQUIRE(num_features >= 3, "hard_voxelize: num_features must be >= 3");
    SEEVCN_REQUIRE(pc_range && voxel_size && grid_size && num_voxels, "hard_voxelize: null pointer");
    cudaStream_t st = as_stream(stream);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(num_voxels, 0, sizeof(int), st));
    if (num_points == 0 || max_voxels == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && voxels && coordinates && num_points_per_voxel && workspace, "hard_voxelize: null pointer");
    const HardWs w = hard_layout(num_points, max_points, max_voxels);
    if (workspace_bytes < w.total) {
        seevcn_set_error("hard_voxelize: workspace %zu < %zu", workspace_bytes, w.total);
        return SEEVCN_E_WORKSPACE;
    }
    VoxGeom g;
    for (int i = 0; i < 3; ++i) {
        g.lo[i] = pc_range[i]; g.vs[i] = voxel_size[i]; g.g[i] = grid_size[i];
        SEEVCN_REQUIRE(grid_size[i] > 0 && voxel_size[i] > 0.f, "hard_voxelize: bad grid");
    }
    char* ws = static_cast<char*>(workspace);
    auto* keys = reinterpret_cast<unsigned long long*>(ws + w.off_keys);
    int* first = reinterpret_cast<int*>(ws + w.off_first);
    int* vid = reinterpret_cast<int*>(ws + w.off_vid);
    int* slot = reinterpret_cast<int*>(ws + w.off_slot);
    int* flag = reinterpret_cast<int*>(ws + w.off_flag);
    int* rank = reinterpret_cast<int*>(ws + w.off_rank);
    int* sel = reinterpret_cast<int*>(ws + w.off_sel);
    int* cnt = reinterpret_cast<int*>(ws + w.off_cnt);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(keys, 0xff, w.nslots * 8, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(first, 0x7f, w.nslots * 4, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(vid, 0xff, w.nslots * 4, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(sel, 0x7f, (size_t)max_voxels * max_points * 4, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(cnt, 0, (size_t)max_voxels * 4, st));
    const int gp = div_up(num_points, 256);
    hardvox_insert_kernel<<<gp, 256, 0, st>>>(num_points, num_features, points, g, w.nslots - 1, keys, first, slot);
    SEEVCN_LAUNCH_CHECK();
    hardvox_flag_kernel<<<gp, 256, 0, st>>>(num_points, first, slot, flag);
    SEEVCN_LAUNCH_CHECK();
    size_t cub_bytes = w.cub_bytes;
    SEEVCN_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws + w.off_cub, cub_bytes, flag, rank, num_points, st));
    hardvox_assign_kernel<<<gp, 256, 0, st>>>(num_points, max_voxels, g, keys, slot, flag, rank, vid, coordinates, num_voxels);
    SEEVCN_LAUNCH_CHECK();
    hardvox_pick_kernel<<<gp, 256, 0, st>>>(num_points, max_points, slot, vid, sel, cnt);
    SEEVCN_LAUNCH_CHECK();
    const long long tot = (long long)max_voxels * max_points;
    hardvox_gather_kernel<<<(unsigned)div_up(tot, 256ll), 256, 0, st>>>(max_voxels, max_points, num_features, points,
                                                                        num_voxels, sel, cnt, voxels, num_points_per_voxel);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
