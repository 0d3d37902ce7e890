// Stage 6 — voxelization: MeanVFE, dynamic (bucketed scatter-mean) and hard (spconv-style).
//
// Replaces
//   MeanVFE.forward          detector3d/pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31
//   DynamicMeanVFE.forward   detector3d/pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:37-76
//     (torch.unique = device-wide sort + torch_scatter.scatter_mean + ~15 elementwise launches)
//   VoxelGeneratorWrapper    detector3d/pcdet/datasets/processor/data_processor.py:15-60
//     (single-thread CPU loop inside third-party spconv)
//
// Dynamic voxelization (dynvox_*): the reference's merge key  b*XYZ + x*YZ + y*Z + z  (64-bit here) is split into
// bucket = key >> LOGW and cell = key & (W-1).  Buckets are contiguous key ranges, so rows grouped by bucket and ordered
// by cell inside a bucket ARE in torch.unique order — no sort pass, no hash table, no capacity guess:
//   1. count    one pass over the points: key, warp-aggregated atomicAdd on the bucket histogram, the returned rank is kept
//   2. scan     single-pass look-back scan of the histogram -> bucket offsets + the list of non-empty buckets
//   3. scatter  second pass over the points: 16-byte record {cell, fixed-point offsets inside the voxel} to offset + rank
//   4. reduce   a warp per non-empty bucket: occupancy bitmap of its W cells in shared memory -> popcount prefix =
//               the voxel's row inside the bucket; integer accumulation in shared memory; rows written at the bucket's
//               global row offset, which comes from a second look-back chain over the buckets
// The mean is accumulated as 2^-30-voxel fixed-point offsets from the voxel's origin in integers, so it does not depend
// on the order the points arrive in (bit-reproducible run to run) and is the float64 mean rounded once (the reference's
// scatter_mean adds absolute fp32 coordinates with atomics: ~1e-6 relative noise of its own).  Features beyond xyz are
// plain fp32 atomic sums.
// HBM-bound: 4*(1+C) B/point in, (16 + 4*C + 4) B/voxel out; the records (16 B/point) stay in L2 at frame-batch sizes.
#include <algorithm>
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int kMaxFeat = 8;   // point features (xyz + up to 4 extras) held in registers

struct VoxGeom {
    float lo[3], vs[3];
    int g[3];
};

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

// ------------------------------------------------------------------ dynamic voxelization --
struct DynSrc {   // either one (N,1+C) matrix, or frames + object clouds
    const float* points; int c;
    int n_frame_pts, pts_per_frame; const float* frame_pts; const unsigned char* frame_keep;
    int n_obj_pts, pts_per_obj; const float* obj_pts; const int* obj_frame; const int* obj_count;
    int nbatch;
};

// Row p of the virtual [batch_idx, features...] matrix.  false: the row does not take part (spliced out, cyclic repeat
// of an object's distinct rows, batch index outside [0, nbatch)).
__device__ __forceinline__ bool dyn_load(const DynSrc& s, int p, int& b, float (&f)[kMaxFeat]) {
    const float* row;
    int nf;
    if (s.points) {
        row = s.points + (size_t)p * (1 + s.c);
        const float bf = row[0];
        if (!(bf >= 0.f && bf < (float)s.nbatch)) return false;
        b = (int)bf;   // points[:,0].int()
        ++row; nf = s.c;
    } else if (p < s.n_frame_pts) {
        if (s.frame_keep && !s.frame_keep[p]) return false;          // replaced by a completed cloud (splice step)
        row = s.frame_pts + (size_t)p * 3; b = p / s.pts_per_frame; nf = 3;
    } else {
        const int q = p - s.n_frame_pts, o = q / s.pts_per_obj;
        if (s.obj_count && q - o * s.pts_per_obj >= s.obj_count[o]) return false;   // only the object's distinct rows
        b = s.obj_frame[o];
        if (b < 0 || b >= s.nbatch) return false;
        row = s.obj_pts + (size_t)q * 3; nf = 3;
    }
#pragma unroll
    for (int j = 0; j < kMaxFeat; ++j) f[j] = j < nf ? row[j] : 0.f;
    return true;
}

__device__ __forceinline__ unsigned long long dyn_key(const VoxGeom& g, int b, int cx, int cy, int cz) {
    return (unsigned long long)((((long long)b * g.g[0] + cx) * g.g[1] + cy) * (long long)g.g[2] + cz);
}

// 1. count: hist[bucket] += 1 per in-grid point; rank[p] = the point's arrival number inside its bucket (~0: not voxelized).
// Lanes of a warp that hit the same bucket (consecutive returns of a beam do) share one atomic.
__global__ void __launch_bounds__(256)
dynvox_count_kernel(DynSrc s, VoxGeom g, int n, int logw, unsigned* __restrict__ hist, unsigned* __restrict__ rank) {
    const int nthreads = gridDim.x * blockDim.x;
    const int rounds = (n + nthreads - 1) / nthreads;
    const int lane = lane_id();
    for (int it = 0; it < rounds; ++it) {
        const int p = it * nthreads + blockIdx.x * blockDim.x + threadIdx.x;
        float f[kMaxFeat];
        int b = 0, cx, cy, cz;
        bool valid = p < n && dyn_load(s, p, b, f);
        valid = valid && voxel_coord(g, f[0], f[1], f[2], cx, cy, cz);
        const unsigned bucket = valid ? (unsigned)(dyn_key(g, b, cx, cy, cz) >> logw) : 0xffffffffu;
        const unsigned m = __match_any_sync(0xffffffffu, bucket);
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if (valid && lane == leader) base = atomicAdd(&hist[bucket], (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (p < n) rank[p] = valid ? base + (unsigned)__popc(m & ((1u << lane) - 1u)) : 0xffffffffu;
    }
}

// 2. scan: hist (counts) -> exclusive offsets in place; tasks[i] = {bucket, start, count, 0} for the i-th non-empty bucket;
// meta[1] = number of non-empty buckets, meta[2] = number of voxelized points.  One look-back chain carries both sums
// (31 bits each).  meta[0] = tile ticket.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
dynvox_scan_kernel(int nb, unsigned* __restrict__ hist, unsigned long long* __restrict__ status, int* __restrict__ meta,
                   int4* __restrict__ tasks) {
    __shared__ int s_tile;
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(&meta[0], 1);
    __syncthreads();
    const int tile = s_tile;
    const int i0 = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned long long sum = 0;
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) {
        v[e] = i0 + e < nb ? hist[i0 + e] : 0u;
        sum += (unsigned long long)v[e] + ((unsigned long long)(v[e] != 0u) << 31);
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, sft);
        if (lane_id() >= sft) inc += t;
    }
    if (lane_id() == 31) s_warp[warp_id()] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned long long w = threadIdx.x < kScanThreads / 32 ? s_warp[threadIdx.x] : 0ull;
        unsigned long long winc = w;
#pragma unroll
        for (int sft = 1; sft < kScanThreads / 32; sft <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, sft);
            if ((int)threadIdx.x >= sft) winc += t;
        }
        if (threadIdx.x < kScanThreads / 32) s_warp[threadIdx.x] = winc - w;
        const unsigned long long agg = __shfl_sync(0xffffffffu, winc, kScanThreads / 32 - 1);
        const unsigned long long excl = seevcn_scan::lookback_publish(status, tile, agg);
        if (threadIdx.x == 0) {
            s_excl = excl;
            if ((long long)(tile + 1) * kScanTile >= nb) {
                const unsigned long long tot = excl + agg;
                meta[1] = (int)(tot >> 31);
                meta[2] = (int)(tot & 0x7fffffffull);
            }
        }
    }
    __syncthreads();
    const unsigned long long ex = s_excl + s_warp[warp_id()] + inc - sum;
    unsigned run = (unsigned)(ex & 0x7fffffffull);
    unsigned ne = (unsigned)(ex >> 31);
#pragma unroll
    for (int e = 0; e < kScanItems; ++e) {
        if (i0 + e < nb) {
            hist[i0 + e] = run;
            if (v[e]) tasks[ne++] = make_int4(i0 + e, (int)run, (int)v[e], 0);
            run += v[e];
        }
    }
}

// Fixed point of a coordinate inside its voxel: u = (p - origin) / voxel_size in [0,1) up to fp32 rounding of the floor;
// stored as (u + 0.25) * 2^30 so it is a positive 31-bit integer.
constexpr double kFixOne = 1073741824.0;   // 2^30
constexpr double kFixBias = 0.25;

struct VoxGeomD { double lo[3], vs[3], inv_vs[3]; };

// 3. scatter: record {cell, qx, qy, qz} (+ {f3..f6} when WIDE) to offsets[bucket] + rank.
template <bool WIDE>
__global__ void __launch_bounds__(256)
dynvox_scatter_kernel(DynSrc s, VoxGeom g, VoxGeomD gd, int n, int logw, const unsigned* __restrict__ offsets,
                      const unsigned* __restrict__ rank, uint4* __restrict__ records) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const unsigned r = rank[p];
        if (r == 0xffffffffu) continue;
        float f[kMaxFeat];
        int b = 0, c[3];
        dyn_load(s, p, b, f);
        voxel_coord(g, f[0], f[1], f[2], c[0], c[1], c[2]);
        const unsigned long long key = dyn_key(g, b, c[0], c[1], c[2]);
        const unsigned bucket = (unsigned)(key >> logw);
        unsigned q[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double origin = gd.lo[j] + (double)c[j] * gd.vs[j];
            double u = ((double)f[j] - origin) * gd.inv_vs[j];
            u = fmin(fmax(u, -kFixBias), 1.0 + kFixBias);
            q[j] = (unsigned)__double2ll_rn((u + kFixBias) * kFixOne);
        }
        const size_t pos = (size_t)offsets[bucket] + r;
        const uint4 rec = make_uint4((unsigned)(key & ((1ull << logw) - 1ull)), q[0], q[1], q[2]);
        if (WIDE) {
            records[2 * pos] = rec;
            records[2 * pos + 1] = make_uint4(__float_as_uint(f[3]), __float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]));
        } else {
            records[pos] = rec;
        }
    }
}

// 4. reduce.  Shared memory per warp: bitmap[W/32] u32 | prefix[W/32] u16 | acc[kChunk][ACCW] u32.
// acc row: {x lo, x hi, y lo, y hi, z lo, z hi, count, cell} (+ 4 fp32 sums when WIDE).
constexpr int kChunk = 64;   // voxel rows of a bucket accumulated per pass over its records

__device__ __forceinline__ void add_u64_split(unsigned* lo_hi, unsigned long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(lo_hi, lo);
    const unsigned carry = (old + lo) < old ? 1u : 0u;
    if (hi + carry) atomicAdd(lo_hi + 1, hi + carry);
}

template <bool WIDE>
__global__ void __launch_bounds__(256)
dynvox_reduce_kernel(int logw, int c, VoxGeom g, VoxGeomD gd, const int4* __restrict__ tasks, int* __restrict__ meta,
                     const uint4* __restrict__ records, unsigned long long* __restrict__ status, int max_voxels,
                     int* __restrict__ voxel_coords, float* __restrict__ voxel_features, int* __restrict__ voxel_counts,
                     int* __restrict__ num_voxels) {
    constexpr int ACCW = WIDE ? 12 : 8;
    constexpr int RS = WIDE ? 2 : 1;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int words = 1 << (logw - 5);
    const int wpl = words >> 5;                                        // bitmap words per lane (logw >= 10)
    const size_t per_warp = (size_t)words * 6 + (size_t)kChunk * ACCW * 4;
    unsigned char* mine = s_raw + per_warp * warp_id();
    unsigned* bm = reinterpret_cast<unsigned*>(mine);
    unsigned* acc = reinterpret_cast<unsigned*>(mine + (size_t)words * 4);
    unsigned short* pre = reinterpret_cast<unsigned short*>(mine + (size_t)words * 4 + (size_t)kChunk * ACCW * 4);
    const int lane = lane_id();
    const int ntasks = meta[1];
    const unsigned long long yz = (unsigned long long)g.g[1] * g.g[2];
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&meta[3], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;
        const int4 d = tasks[t];
        const int bucket = d.x, n = d.z;
        const uint4* rec = records + (size_t)d.y * RS;

        // occupancy bitmap of the bucket's cells -> row of a voxel inside the bucket = popcount prefix
        for (int w = lane; w < words; w += 32) bm[w] = 0u;
        __syncwarp();
        for (int r = lane; r < n; r += 32) {
            const unsigned cell = rec[(size_t)r * RS].x;
            atomicOr(&bm[cell >> 5], 1u << (cell & 31));
        }
        __syncwarp();
        int cnt = 0;
        for (int k = 0; k < wpl; ++k) cnt += __popc(bm[lane * wpl + k]);
        int inc = cnt;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, sft); if (lane >= sft) inc += v; }
        const int nv = __shfl_sync(0xffffffffu, inc, 31);
        int run = inc - cnt;
        for (int k = 0; k < wpl; ++k) { pre[lane * wpl + k] = (unsigned short)run; run += __popc(bm[lane * wpl + k]); }
        seevcn_scan::publish_aggregate(status, t, (unsigned long long)nv);
        __syncwarp();

        long long base = -1;
        for (int c0 = 0; c0 < nv; c0 += kChunk) {
            for (int w = lane; w < kChunk * ACCW; w += 32) acc[w] = 0u;
            __syncwarp();
            for (int r0 = 0; r0 < n; r0 += 32) {
                const int r = r0 + lane;
                uint4 q = make_uint4(0, 0, 0, 0);
                if (r < n) q = rec[(size_t)r * RS];
                const unsigned cell = q.x;
                const int vr = (int)pre[cell >> 5] + __popc(bm[cell >> 5] & ((1u << (cell & 31)) - 1u)) - c0;
                const bool in = r < n && vr >= 0 && vr < kChunk;
                const unsigned mk = in ? (unsigned)vr : (0x80000000u | (unsigned)lane);
                const unsigned m = __match_any_sync(0xffffffffu, mk);
                unsigned* a = acc + (in ? vr : 0) * ACCW;
                if (!__any_sync(0xffffffffu, __popc(m) > 4)) {
                    if (in) {
                        add_u64_split(a + 0, q.y); add_u64_split(a + 2, q.z); add_u64_split(a + 4, q.w);
                        atomicAdd(a + 6, 1u);
                        a[7] = cell;
                    }
                } else {   // many points of the warp in one voxel (dense blobs): one atomic per voxel instead of one per point
                    unsigned long long sum[3];
                    const unsigned qq[3] = {q.y, q.z, q.w};
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const unsigned lo = __reduce_add_sync(m, qq[j] & 0xffffu);
                        const unsigned hi = __reduce_add_sync(m, qq[j] >> 16);
                        sum[j] = ((unsigned long long)hi << 16) + lo;
                    }
                    if (in && lane == __ffs(m) - 1) {
                        add_u64_split(a + 0, sum[0]); add_u64_split(a + 2, sum[1]); add_u64_split(a + 4, sum[2]);
                        atomicAdd(a + 6, (unsigned)__popc(m));
                        a[7] = cell;
                    }
                }
                if (WIDE && in) {
                    const uint4 e = rec[(size_t)r * RS + 1];
                    atomicAdd(reinterpret_cast<float*>(a + 8), __uint_as_float(e.x));
                    atomicAdd(reinterpret_cast<float*>(a + 9), __uint_as_float(e.y));
                    atomicAdd(reinterpret_cast<float*>(a + 10), __uint_as_float(e.z));
                    atomicAdd(reinterpret_cast<float*>(a + 11), __uint_as_float(e.w));
                }
            }
            __syncwarp();
            if (base < 0) base = (long long)seevcn_scan::resolve_prefix(status, t, (unsigned long long)nv);
            const int nrow = min(kChunk, nv - c0);
            for (int v = lane; v < nrow; v += 32) {
                const long long row = base + c0 + v;
                if (row >= max_voxels) continue;
                const unsigned* a = acc + v * ACCW;
                const unsigned pc = a[6];
                const unsigned long long key = ((unsigned long long)(unsigned)bucket << logw) | a[7];
                int cc[3], bb;
                if ((key >> 32) == 0ull && (yz >> 32) == 0ull) {   // 32-bit decode when it fits (the usual case)
                    const unsigned k32 = (unsigned)key, yz32 = (unsigned)yz;
                    const unsigned bx = k32 / yz32, rem = k32 - bx * yz32;
                    cc[2] = (int)(rem % (unsigned)g.g[2]); cc[1] = (int)(rem / (unsigned)g.g[2]);
                    cc[0] = (int)(bx % (unsigned)g.g[0]); bb = (int)(bx / (unsigned)g.g[0]);
                } else {
                    const unsigned long long bx = key / yz, rem = key - bx * yz;
                    cc[2] = (int)(rem % (unsigned long long)g.g[2]); cc[1] = (int)(rem / (unsigned long long)g.g[2]);
                    cc[0] = (int)(bx % (unsigned long long)g.g[0]); bb = (int)(bx / (unsigned long long)g.g[0]);
                }
                reinterpret_cast<int4*>(voxel_coords)[row] = make_int4(bb, cc[2], cc[1], cc[0]);   // [b,z,y,x]
                const double inv = 1.0 / ((double)pc * kFixOne);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const unsigned long long sum = ((unsigned long long)a[2 * j + 1] << 32) | a[2 * j];
                    const double u = (double)sum * inv - kFixBias;
                    voxel_features[(size_t)row * c + j] = (float)(gd.lo[j] + ((double)cc[j] + u) * gd.vs[j]);
                }
                if (WIDE)
                    for (int j = 3; j < c; ++j)
                        voxel_features[(size_t)row * c + j] = __fdiv_rn(__uint_as_float(a[8 + j - 3]), (float)pc);
                voxel_counts[row] = (int)pc;
            }
            __syncwarp();
        }
        if (t == ntasks - 1 && lane == 0) *num_voxels = (int)(base + nv);
    }
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
constexpr unsigned long long kEmpty = ~0ull;
__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

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
    int scan_tiles;
    size_t off_keys, off_first, off_vid, off_slot, off_flag, off_rank, off_sel, off_cnt, off_scan, scan_bytes, total;
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
    w.scan_tiles = (int)div_up(np, (size_t)seevcn_scan::kScanTile);
    w.scan_bytes = 8 * (size_t)w.scan_tiles + 64;          // look-back status words + the tile ticket
    w.off_scan = take(w.scan_bytes);
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


namespace {

struct DynWs {
    int logw, nb, scan_tiles, max_tasks, wide;
    size_t off_meta, off_hist, off_scan_status, off_red_status, zero_bytes, off_tasks, off_rank, off_records, total;
};

// Bucket width: W = 2^logw consecutive keys per bucket.  16384 keys = 410 y-columns of the Waymo grid at one x: a LiDAR
// frame puts ~40 points into a non-empty bucket, so a warp has work for every lane and the histogram stays small.
int dyn_logw(double span) {
    int logw = 14;
    if (const char* e = getenv("SEEVCN_VOX_LOGW")) {   // tuning knob (10..16)
        const int v = atoi(e);
        if (v >= 10 && v <= 16) logw = v;
    }
    while (span / (double)(1ull << logw) >= 1073741824.0 && logw < 16) ++logw;
    return logw;
}

bool dyn_layout(long long n, int c, int batch, const int* grid, DynWs* out) {
    DynWs w{};
    const double span = (double)(batch > 0 ? batch : 1) * grid[0] * grid[1] * grid[2];
    w.logw = dyn_logw(span);
    const double nbd = span / (double)(1ull << w.logw) + 1.0;
    if (nbd >= 1073741824.0 || n >= (1ll << 31)) return false;
    w.nb = (int)nbd;
    w.scan_tiles = div_up(w.nb, kScanTile);
    w.max_tasks = (int)std::min<long long>(w.nb, n > 0 ? n : 1);
    w.wide = c > 3 ? 1 : 0;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    w.off_meta = take(64);
    w.off_hist = take(4 * (size_t)w.nb);
    w.off_scan_status = take(8 * (size_t)w.scan_tiles);
    w.off_red_status = take(8 * (size_t)w.max_tasks);
    w.zero_bytes = o;                                       // everything above starts from zero (one memset)
    w.off_tasks = take(16 * (size_t)w.max_tasks);
    w.off_rank = take(4 * (size_t)(n > 0 ? n : 1));
    w.off_records = take((w.wide ? 32 : 16) * (size_t)(n > 0 ? n : 1));
    w.total = o;
    *out = w;
    return true;
}

int dynvox_run(const char* what, int num_points, int num_features, DynSrc src, const float* pc_range,
               const float* voxel_size, const int* grid_size, int max_voxels, int* voxel_coords, float* voxel_features,
               int* voxel_counts, int* num_voxels, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    SEEVCN_PROF("dynamic_voxelize", st);
    VoxGeom g;
    VoxGeomD gd;
    for (int i = 0; i < 3; ++i) {
        SEEVCN_REQUIRE(grid_size[i] > 0 && voxel_size[i] > 0.f, "%s: bad grid", what);
        g.lo[i] = pc_range[i]; g.vs[i] = voxel_size[i]; g.g[i] = grid_size[i];
        gd.lo[i] = (double)pc_range[i]; gd.vs[i] = (double)voxel_size[i]; gd.inv_vs[i] = 1.0 / (double)voxel_size[i];
    }
    DynWs w;
    SEEVCN_REQUIRE(dyn_layout(num_points, num_features, src.nbatch, grid_size, &w), "%s: batch x grid too large", what);
    if (workspace_bytes < w.total) {
        seevcn_set_error("%s: workspace %zu < %zu", what, workspace_bytes, w.total);
        return SEEVCN_E_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    int* meta = reinterpret_cast<int*>(ws + w.off_meta);
    unsigned* hist = reinterpret_cast<unsigned*>(ws + w.off_hist);
    auto* scan_status = reinterpret_cast<unsigned long long*>(ws + w.off_scan_status);
    auto* red_status = reinterpret_cast<unsigned long long*>(ws + w.off_red_status);
    int4* tasks = reinterpret_cast<int4*>(ws + w.off_tasks);
    unsigned* rank = reinterpret_cast<unsigned*>(ws + w.off_rank);
    uint4* records = reinterpret_cast<uint4*>(ws + w.off_records);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(ws, 0, w.zero_bytes, st));
    const int sms = seevcn_num_sms();
    const int grid_pts = (int)std::min<long long>(div_up((long long)num_points, 256ll), (long long)sms * 16);
    dynvox_count_kernel<<<grid_pts, 256, 0, st>>>(src, g, num_points, w.logw, hist, rank);
    SEEVCN_LAUNCH_CHECK();
    dynvox_scan_kernel<<<w.scan_tiles, kScanThreads, 0, st>>>(w.nb, hist, scan_status, meta, tasks);
    SEEVCN_LAUNCH_CHECK();
    const int words = 1 << (w.logw - 5);
    const int accw = w.wide ? 12 : 8;
    const size_t per_warp = (size_t)words * 6 + (size_t)kChunk * accw * 4;
    const int warps = per_warp * 8 <= 56 * 1024 ? 8 : 4;
    const size_t smem = per_warp * warps;
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / (warps * 32), (220 * 1024) / (smem + 1024)));
    const int grid_red = sms * ctas_per_sm;
    if (w.wide) {
        dynvox_scatter_kernel<true><<<grid_pts, 256, 0, st>>>(src, g, gd, num_points, w.logw, hist, rank, records);
        SEEVCN_LAUNCH_CHECK();
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(dynvox_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dynvox_reduce_kernel<true><<<grid_red, warps * 32, smem, st>>>(w.logw, num_features, g, gd, tasks, meta, records, red_status,
                                                                      max_voxels, voxel_coords, voxel_features, voxel_counts, num_voxels);
    } else {
        dynvox_scatter_kernel<false><<<grid_pts, 256, 0, st>>>(src, g, gd, num_points, w.logw, hist, rank, records);
        SEEVCN_LAUNCH_CHECK();
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(dynvox_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dynvox_reduce_kernel<false><<<grid_red, warps * 32, smem, st>>>(w.logw, num_features, g, gd, tasks, meta, records, red_status,
                                                                       max_voxels, voxel_coords, voxel_features, voxel_counts, num_voxels);
    }
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
}  // namespace

extern "C" size_t seevcn_dynamic_voxelize_workspace_bytes(int num_points, int num_features, int batch_size, const int* grid_size) {
    DynWs w;
    if (!grid_size || !dyn_layout(num_points, num_features, batch_size, grid_size, &w)) return 0;
    return w.total;
}

extern "C" int seevcn_dynamic_voxelize(int num_points, int num_features, const float* points, const float* pc_range,
                                       const float* voxel_size, const int* grid_size, int max_voxels, int sorted,
                                       int batch_size, int* voxel_coords, float* voxel_features, int* voxel_counts, int* num_voxels,
                                       void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    (void)sorted;   // rows always come out in the reference's torch.unique order
    SEEVCN_REQUIRE(num_points >= 0 && max_voxels >= 0, "dynamic_voxelize: negative size");
    SEEVCN_REQUIRE(batch_size >= 1, "dynamic_voxelize: batch_size=%d (batch_dict['batch_size'] of the reference) must be >= 1", batch_size);
    SEEVCN_REQUIRE(num_features >= 3 && num_features < kMaxFeat, "dynamic_voxelize: num_features=%d outside [3,%d]",
                   num_features, kMaxFeat - 1);
    SEEVCN_REQUIRE(pc_range && voxel_size && grid_size && num_voxels, "dynamic_voxelize: null pointer");
    cudaStream_t st = as_stream(stream);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(num_voxels, 0, sizeof(int), st));
    if (num_points == 0 || max_voxels == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && voxel_coords && voxel_features && voxel_counts && workspace,
                   "dynamic_voxelize: null pointer");
    DynSrc src{};
    src.points = points; src.c = num_features; src.nbatch = batch_size;
    return dynvox_run("dynamic_voxelize", num_points, num_features, src, pc_range, voxel_size, grid_size, max_voxels,
                      voxel_coords, voxel_features, voxel_counts, num_voxels, workspace, workspace_bytes, st);
}

extern "C" int seevcn_dynamic_voxelize_spliced(int num_frames, int pts_per_frame, const float* frame_pts,
                                               const unsigned char* frame_keep, int num_obj,
                                               int pts_per_obj, const float* obj_pts, const int* obj_frame, const int* obj_count,
                                              const float* pc_range, const float* voxel_size, const int* grid_size,
                                              int max_voxels, int sorted, int* voxel_coords, float* voxel_features,
                                              int* voxel_counts, int* num_voxels, void* workspace, size_t workspace_bytes,
                                              seevcn_stream_t stream) {
    (void)sorted;
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
    src.c = 3; src.nbatch = num_frames > 0 ? num_frames : 1;
    src.n_frame_pts = (int)n_frame; src.pts_per_frame = pts_per_frame > 0 ? pts_per_frame : 1; src.frame_pts = frame_pts; src.frame_keep = frame_keep;
    src.n_obj_pts = (int)n_obj; src.pts_per_obj = pts_per_obj > 0 ? pts_per_obj : 1; src.obj_pts = obj_pts; src.obj_frame = obj_frame; src.obj_count = obj_count;
    return dynvox_run("dynamic_voxelize_frames", (int)(n_frame + n_obj), 3, src, pc_range, voxel_size, grid_size, max_voxels,
                      voxel_coords, voxel_features, voxel_counts, num_voxels, workspace, workspace_bytes, st);
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
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(ws + w.off_scan, 0, w.scan_bytes, st));
    seevcn_scan::exclusive_scan_u32_kernel<<<w.scan_tiles, seevcn_scan::kScanThreads, 0, st>>>(
        num_points, reinterpret_cast<const unsigned*>(flag), reinterpret_cast<unsigned*>(rank),
        reinterpret_cast<unsigned long long*>(ws + w.off_scan), reinterpret_cast<int*>(ws + w.off_scan + 8 * (size_t)w.scan_tiles),
        nullptr);
    SEEVCN_LAUNCH_CHECK();
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
