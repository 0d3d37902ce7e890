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
//   4. reduce   CTA tiles of consecutive non-empty buckets: per-bucket occupancy bitmap of the W cells in shared memory ->
//               popcount prefix = the voxel's row inside the bucket; integer accumulation in shared memory; rows written
//               at the tile's global row offset, which comes from a second look-back chain over the tiles
// The mean is accumulated as 2^-23-voxel fixed-point offsets from the voxel's origin in integers, so it does not depend
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
// (31 bits each).  meta[0] = tile ticket.  Also cuts the task list into the reduce kernel's tiles: a task weighs its
// records + task_weight, tile k = the tasks whose weighted start lies in [k * tile_weight, (k+1) * tile_weight), so a tile
// holds < tile_weight + (its last task's) records and <= tile_weight / task_weight tasks.  tile_info[k] = {its first task,
// its first record}.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
dynvox_scan_kernel(int nb, unsigned* __restrict__ hist, unsigned long long* __restrict__ status, int* __restrict__ meta,
                   int4* __restrict__ tasks, int tile_weight, int task_weight, int max_red_tiles, int2* __restrict__ tile_info) {
    __shared__ int s_tile;
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_excl;
    __shared__ unsigned s_tot[2];
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
    const bool last_tile = (long long)(tile + 1) * kScanTile >= nb;
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
            const unsigned long long tot = excl + agg;
            s_tot[0] = (unsigned)(tot >> 31); s_tot[1] = (unsigned)(tot & 0x7fffffffull);
            if (last_tile) { meta[1] = (int)s_tot[0]; meta[2] = (int)s_tot[1]; }
            if (tile == 0) tile_info[0] = make_int2(0, 0);
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
            if (v[e]) {
                tasks[ne] = make_int4(i0 + e, (int)run, (int)v[e], 0);
                // tile boundaries k * tile_weight in (weighted start, weighted end]: the next task is the first at or after them
                const unsigned long long ws = (unsigned long long)run + (unsigned long long)task_weight * ne;
                const unsigned long long we = ws + v[e] + task_weight;
                for (unsigned long long k = ws / (unsigned)tile_weight + 1; k * (unsigned)tile_weight <= we; ++k)
                    if (k <= (unsigned long long)max_red_tiles) tile_info[k] = make_int2((int)ne + 1, (int)(run + v[e]));
                ++ne;
            }
            run += v[e];
        }
    }
}

// Fixed point of a coordinate inside its voxel: u = (p - origin) / voxel_size in [0,1) up to fp32 rounding of the floor;
// stored as (u + 0.25) * 2^23, a positive 24-bit integer (resolution 1.2e-8 of a voxel edge, far below an fp32 ulp of
// the coordinates).  256 of them sum without carry in 32 bits (the small-bucket path), 2^40 in 64 bits.
constexpr double kFixOne = 8388608.0;      // 2^23
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

// 4. reduce.  Persistent CTAs take tiles (consecutive buckets, cut by the scan kernel) from a ticket, so tiles start in
// order.  Everything is a thread per record or a warp per bucket on shared memory; nothing waits on global memory per
// bucket:
//   a. every bucket of the tile gets an occupancy bitmap of its W cells in shared memory; threads set the bits of their
//      records (a bucket with more than kSmall records is streamed by the whole CTA instead, straight from global memory)
//   b. a warp per bucket: popcount prefix over the bitmap words -> row of a cell inside its bucket (ascending cell =
//      torch.unique order) and the bucket's voxel count
//   c. the CTA scans the counts, publishes the tile's total and resolves its global row offset with a look-back over the
//      earlier tiles, every thread reading one predecessor per step
//   d. threads add their records into shared-memory integer accumulators, one row per voxel of the tile (big buckets:
//      one after the other, streamed by the whole CTA), then a thread per row writes coords / mean / count.
// Shared memory: task[T] int4 | rows_all[T] | rows_small[T] | bitmap[T][W/32] u32 | prefix[T][W/32] u16 | acc[kAccRows][ACCW] u32
//   with T = tile_weight / task_weight tasks: 48 KB at W = 1024, four CTAs per SM.
// Measured (B200, 8 frames x 180k points -> 467k voxels): ~15 us per tile of ~680 records / 42 buckets / 240 voxels, of which
// ~3.5 us is the look-back; 592 tiles in flight.  The kernel is bound by the latency of its barrier-separated phases.
// acc row: {x lo, x hi, y lo, y hi, z lo, z hi, count, cell | task << 16} (+ 4 fp32 sums when WIDE).
constexpr int kSmall = 256;       // buckets up to this many records take the thread-per-record path
constexpr int kRedThreads = 256;
constexpr int kRedWarps = kRedThreads / 32;
constexpr int kTileWeight = 1024; // records + kTaskWeight per bucket
constexpr int kAccRows = 640;     // voxel rows accumulated at a time (20 KB of shared memory)
constexpr int kRecSlots = (kTileWeight + kSmall) / kRedThreads;   // records per thread
static_assert(kRecSlots * kRedThreads == kTileWeight + kSmall, "tile capacity must be a multiple of the CTA size");

__device__ __forceinline__ void add_u64_split(unsigned* lo_hi, unsigned long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(lo_hi, lo);
    const unsigned carry = (old + lo) < old ? 1u : 0u;
    if (hi + carry) atomicAdd(lo_hi + 1, hi + carry);
}

struct VoxOut {
    int c, max_voxels;
    int* coords; float* features; int* counts;
};

// One output row: decode the merge key, mean = voxel origin + fixed-point mean offset (float64, rounded once).
template <bool WIDE>
__device__ __forceinline__ void write_voxel_row(const VoxOut& o, const VoxGeom& g, const VoxGeomD& gd, long long row, int logw,
                                                int bucket, unsigned cell, const unsigned* a) {
    if (row >= o.max_voxels) return;
    const unsigned pc = a[6];
    const unsigned long long yz = (unsigned long long)g.g[1] * g.g[2];
    const unsigned long long key = ((unsigned long long)(unsigned)bucket << logw) | cell;
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
    reinterpret_cast<int4*>(o.coords)[row] = make_int4(bb, cc[2], cc[1], cc[0]);   // [b,z,y,x]
    const double inv = 1.0 / ((double)pc * kFixOne);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const unsigned long long sum = ((unsigned long long)a[2 * j + 1] << 32) | a[2 * j];
        const double u = (double)sum * inv - kFixBias;
        o.features[(size_t)row * o.c + j] = (float)(gd.lo[j] + ((double)cc[j] + u) * gd.vs[j]);
    }
    if (WIDE)
        for (int j = 3; j < o.c; ++j) o.features[(size_t)row * o.c + j] = __fdiv_rn(__uint_as_float(a[8 + j - 3]), (float)pc);
    o.counts[row] = (int)pc;
}

template <bool WIDE>
__device__ __forceinline__ void acc_add(unsigned* a, const uint4& q, unsigned tag) {
    add_u64_split(a + 0, q.y); add_u64_split(a + 2, q.z); add_u64_split(a + 4, q.w);
    atomicAdd(a + 6, 1u);
    a[7] = tag;
}
// Same row layout for a bucket of at most kSmall records: the low words cannot overflow, the high words stay zero and
// nothing is read back (four fire-and-forget shared-memory reductions per record).
static_assert(kSmall <= 256, "small-bucket sums must fit 32 bits: kSmall * 1.25 * 2^24 <= 2^32");
__device__ __forceinline__ void acc_add_small(unsigned* a, const uint4& q, unsigned tag) {
    atomicAdd(a + 0, q.y); atomicAdd(a + 2, q.z); atomicAdd(a + 4, q.w);
    atomicAdd(a + 6, 1u);
    a[7] = tag;
}

template <bool WIDE>
__global__ void __launch_bounds__(kRedThreads)
dynvox_reduce_kernel(int logw, VoxGeom g, VoxGeomD gd, const int4* __restrict__ tasks, int* __restrict__ meta,
                     const int2* __restrict__ tile_info, int task_weight, int max_tile_tasks,
                     const uint4* __restrict__ records, unsigned long long* __restrict__ status, VoxOut out,
                     int* __restrict__ num_voxels) {
    constexpr int ACCW = WIDE ? 12 : 8;
    constexpr int RS = WIDE ? 2 : 1;
    constexpr int cap = kTileWeight + kSmall;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_tile;
    __shared__ unsigned s_bigmask[8];                  // bit i: task i of the tile is a big bucket (max_tile_tasks <= 256)
    __shared__ unsigned long long s_wsum[kRedWarps];
    __shared__ unsigned long long s_lb[kRedWarps];     // per warp: bit 63 = found a PREFIX, low bits = sum up to it
    __shared__ unsigned long long s_base;
    const int words = 1 << (logw - 5);                     // 32 (W = 1024) or 64
    const int T = max_tile_tasks;
    int4* s_task = reinterpret_cast<int4*>(s_raw);
    int* s_rows_all = reinterpret_cast<int*>(s_raw + (size_t)T * 16);            // voxels per task -> row offset in the tile
    int* s_rows_small = s_rows_all + T;                                          // same, counting small buckets only
    unsigned* s_bm = reinterpret_cast<unsigned*>(s_rows_small + T);
    unsigned short* s_pre = reinterpret_cast<unsigned short*>(s_bm + (size_t)T * words);
    unsigned* s_acc = reinterpret_cast<unsigned*>(s_raw + (((size_t)T * (24 + 6 * (size_t)words)) + 15 & ~(size_t)15));
    const int lane = lane_id(), warp = warp_id();
    const int ntasks = meta[1], total_rec = meta[2];
    const long long total_w = (long long)total_rec + (long long)task_weight * ntasks;
    const int num_tiles = (int)((total_w + kTileWeight - 1) / kTileWeight);

    while (true) {
        __syncthreads();                                   // the previous tile is done with shared memory and s_tile
        // the ticket is drawn when the tile really starts: later tiles wait for this one's voxel count in their look-back
        if (threadIdx.x == 0) s_tile = atomicAdd(&meta[3], 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= num_tiles) break;
        const int2 ti0 = tile_info[tile];
        int2 ti1 = make_int2(ntasks, total_rec);
        if (tile + 1 < num_tiles) ti1 = tile_info[tile + 1];
        const int t0 = ti0.x, nt = ti1.x - ti0.x;          // 1 <= nt <= T
        const int rec0 = ti0.y;
        const int nstage = min(cap, ti1.y - rec0);         // covers every small bucket of the tile completely
        if (threadIdx.x < 8) s_bigmask[threadIdx.x] = 0u;
        for (int w = threadIdx.x; w < nt * words; w += kRedThreads) s_bm[w] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < nt; i += kRedThreads) {
            const int4 d = tasks[t0 + i];
            s_task[i] = make_int4(d.x, d.y - rec0, d.z, 0);   // {bucket, first record (tile relative), records, -}
            if (d.z > kSmall) atomicOr(&s_bigmask[i >> 5], 1u << (i & 31));
        }
        // ---- a. occupancy bitmaps.  Small buckets: a thread per record (kept in registers for step d).
        uint4 my_rec[kRecSlots];
        int my_task[kRecSlots];
#pragma unroll
        for (int e = 0; e < kRecSlots; ++e) {
            const int r = (int)threadIdx.x + e * kRedThreads;
            if (r < nstage) my_rec[e] = records[(size_t)(rec0 + r) * RS];
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kRecSlots; ++e) {
            const int r = (int)threadIdx.x + e * kRedThreads;
            my_task[e] = -1;
            if (r >= nstage) continue;
            int lo = 0, hi = nt - 1;                       // last task starting at or before r
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_task[mid].y <= r) lo = mid; else hi = mid - 1; }
            if (s_task[lo].z > kSmall) continue;           // big bucket: streamed by the whole CTA below
            my_task[e] = lo;
            const unsigned cell = my_rec[e].x;
            atomicOr(&s_bm[lo * words + (cell >> 5)], 1u << (cell & 31));
        }
        bool any_big = false;
        for (int mw = 0; mw < (nt + 31) / 32; ++mw) {      // big buckets (rare): the whole CTA streams the records from global memory
            for (unsigned m = s_bigmask[mw]; m; m &= m - 1) {
                any_big = true;
                const int i = mw * 32 + __ffs(m) - 1;
                const int4 d = s_task[i];
                const uint4* rec = records + (size_t)(rec0 + d.y) * RS;
                for (int r0 = 0; r0 < d.z; r0 += 4 * kRedThreads) {
                    unsigned cell[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + u * kRedThreads + (int)threadIdx.x;
                        cell[u] = r < d.z ? rec[(size_t)r * RS].x : 0xffffffffu;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (cell[u] != 0xffffffffu) atomicOr(&s_bm[i * words + (cell[u] >> 5)], 1u << (cell[u] & 31));
                }
            }
        }
        __syncthreads();
        // ---- b. a warp per bucket: popcount prefix per bitmap word, voxels of the bucket
        for (int i0 = warp; i0 < nt; i0 += 2 * kRedWarps) {   // two buckets per iteration: independent shuffle chains
            const int i1 = i0 + kRedWarps;
            const bool has1 = i1 < nt;
            int tot0 = 0, tot1 = 0;
            for (int w0 = 0; w0 < words; w0 += 32) {
                const int c0 = __popc(s_bm[i0 * words + w0 + lane]);
                const int c1 = has1 ? __popc(s_bm[i1 * words + w0 + lane]) : 0;
                int inc0 = c0, inc1 = c1;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const int v0 = __shfl_up_sync(0xffffffffu, inc0, sft), v1 = __shfl_up_sync(0xffffffffu, inc1, sft);
                    if (lane >= sft) { inc0 += v0; inc1 += v1; }
                }
                s_pre[i0 * words + w0 + lane] = (unsigned short)(tot0 + inc0 - c0);
                if (has1) s_pre[i1 * words + w0 + lane] = (unsigned short)(tot1 + inc1 - c1);
                tot0 += __shfl_sync(0xffffffffu, inc0, 31);
                tot1 += __shfl_sync(0xffffffffu, inc1, 31);
            }
            if (lane == 0) {
                s_rows_all[i0] = tot0; s_rows_small[i0] = ((s_bigmask[i0 >> 5] >> (i0 & 31)) & 1u) ? 0 : tot0;
                if (has1) { s_rows_all[i1] = tot1; s_rows_small[i1] = ((s_bigmask[i1 >> 5] >> (i1 & 31)) & 1u) ? 0 : tot1; }
            }
        }
        __syncthreads();
        // ---- c. exclusive scans of both counts (thread-serial runs + warp scans), tile total
        const int per = (nt + kRedThreads - 1) / kRedThreads;
        const int lo_i = min((int)threadIdx.x * per, nt), hi_i = min(lo_i + per, nt);
        unsigned long long tsum = 0ull;                    // low 32: all, high 32: small only
        for (int i = lo_i; i < hi_i; ++i) tsum += (unsigned long long)(unsigned)s_rows_all[i] | ((unsigned long long)(unsigned)s_rows_small[i] << 32);
        unsigned long long tinc = tsum;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, tinc, sft); if (lane >= sft) tinc += v; }
        if (lane == 31) s_wsum[warp] = tinc;
        __syncthreads();
        unsigned long long wbase = 0ull, aggp = 0ull;
#pragma unroll
        for (int w = 0; w < kRedWarps; ++w) { if (w < warp) wbase += s_wsum[w]; aggp += s_wsum[w]; }
        unsigned long long run = wbase + tinc - tsum;
        for (int i = lo_i; i < hi_i; ++i) {
            const unsigned long long v = (unsigned long long)(unsigned)s_rows_all[i] | ((unsigned long long)(unsigned)s_rows_small[i] << 32);
            s_rows_all[i] = (int)(unsigned)run; s_rows_small[i] = (int)(unsigned)(run >> 32);
            run += v;
        }
        const int agg = (int)(unsigned)aggp, small_rows = (int)(unsigned)(aggp >> 32);
        if (threadIdx.x == 0) {
            seevcn_scan::st_status(status + tile, (tile == 0 ? seevcn_scan::kFlagPrefix : seevcn_scan::kFlagAgg) | (unsigned long long)agg);
            s_base = 0ull;
        }
        // ---- d. small buckets: integer accumulation, kAccRows voxels of the tile at a time (usually all of them); the
        // tile's global row offset is resolved after the first round of additions, when the earlier tiles had time to publish
        long long tile_base = -1;
        for (int c0 = 0; c0 < small_rows || tile_base < 0; c0 += kAccRows) {
            const int nrows = max(0, min(kAccRows, small_rows - c0));
            __syncthreads();                               // s_rows_* complete; the rows of the previous round are written
            for (int w = threadIdx.x; w < nrows * ACCW; w += kRedThreads) s_acc[w] = 0u;
            __syncthreads();
#pragma unroll
            for (int e = 0; e < kRecSlots; ++e) {
                const int i = my_task[e];
                if (i < 0) continue;
                const unsigned cell = my_rec[e].x;
                const int w = i * words + (int)(cell >> 5);
                const int row = s_rows_small[i] + (int)s_pre[w] + __popc(s_bm[w] & ((1u << (cell & 31)) - 1u)) - c0;
                if (row < 0 || row >= kAccRows) continue;
                unsigned* a = s_acc + row * ACCW;
                acc_add_small(a, my_rec[e], cell | ((unsigned)i << 16));
                if (WIDE) {
                    const uint4 x = records[(size_t)(rec0 + (int)threadIdx.x + e * kRedThreads) * RS + 1];
                    atomicAdd(reinterpret_cast<float*>(a + 8), __uint_as_float(x.x));
                    atomicAdd(reinterpret_cast<float*>(a + 9), __uint_as_float(x.y));
                    atomicAdd(reinterpret_cast<float*>(a + 10), __uint_as_float(x.z));
                    atomicAdd(reinterpret_cast<float*>(a + 11), __uint_as_float(x.w));
                }
            }
            if (tile_base < 0) {
                // look-back over the earlier tiles, kRedThreads of them per step, nearest first (thread x reads tile j - x).
                // Needed: every status between this tile and the nearest PREFIX; entries behind that PREFIX may still be
                // empty and are not waited for.
                for (int j = tile - 1; j >= 0;) {
                    const int idx = j - (int)threadIdx.x;
                    const unsigned long long st = idx >= 0 ? seevcn_scan::ld_status(status + idx) : seevcn_scan::kFlagPrefix;
                    const unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
                    const unsigned em = __ballot_sync(0xffffffffu, (st >> 62) == 0ull);
                    const int first = pm ? __ffs(pm) - 1 : 32;
                    const bool ok = (em & (first >= 32 ? 0xffffffffu : ((1u << first) - 1u))) == 0u;
                    const unsigned long long part = seevcn_scan::warp_sum_u64(lane <= first && ok ? (st & seevcn_scan::kValueMask) : 0ull);
                    __syncthreads();                       // s_lb / s_base of the previous step consumed
                    if (lane == 0) s_lb[warp] = part | (pm ? (1ull << 63) : 0ull) | (ok ? 0ull : (1ull << 62));
                    __syncthreads();
                    bool done = false, retry = false;
                    unsigned long long add = 0ull;
#pragma unroll
                    for (int w = 0; w < kRedWarps; ++w) {
                        if (!done && !retry) {
                            if ((s_lb[w] >> 62) & 1ull) retry = true;
                            else { add += s_lb[w] & ((1ull << 62) - 1ull); done = (s_lb[w] >> 63) != 0ull; }
                        }
                    }
                    if (retry) continue;                   // uniform: a needed predecessor has not published yet, read again
                    if (threadIdx.x == 0) s_base += add;
                    if (done) break;                       // uniform: every thread read the same s_lb
                    j -= kRedThreads;
                }
            }
            __syncthreads();
            if (tile_base < 0) {
                tile_base = (long long)s_base;
                if (threadIdx.x == 0) {
                    if (tile > 0) seevcn_scan::st_status(status + tile, seevcn_scan::kFlagPrefix | (unsigned long long)(tile_base + agg));
                    if (tile == num_tiles - 1) *num_voxels = (int)(tile_base + agg);
                }
            }
            for (int row = threadIdx.x; row < nrows; row += kRedThreads) {
                const unsigned* a = s_acc + row * ACCW;
                const int i = (int)(a[7] >> 16);
                write_voxel_row<WIDE>(out, g, gd, tile_base + s_rows_all[i] + (c0 + row - s_rows_small[i]), logw, s_task[i].x, a[7] & 0xffffu, a);
            }
        }
        // ---- big buckets, one after the other: the whole CTA streams the records into the bucket's rows
        if (any_big) {
            for (int mw = 0; mw < (nt + 31) / 32; ++mw) {
                for (unsigned m = s_bigmask[mw]; m; m &= m - 1) {
                    const int i = mw * 32 + __ffs(m) - 1;
                    const int4 d = s_task[i];
                    const int n = d.z;
                    const uint4* rec = records + (size_t)(rec0 + d.y) * RS;
                    const int nv = (i + 1 < nt ? s_rows_all[i + 1] : agg) - s_rows_all[i];   // <= W
                    for (int c0 = 0; c0 < nv; c0 += kAccRows) {
                        const int nrows = min(kAccRows, nv - c0);
                        __syncthreads();                   // the rows of the previous round are written
                        for (int w = threadIdx.x; w < nrows * ACCW; w += kRedThreads) s_acc[w] = 0u;
                        __syncthreads();
                        for (int r0 = 0; r0 < n; r0 += 4 * kRedThreads) {
                            uint4 q[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int r = r0 + u * kRedThreads + (int)threadIdx.x;
                                q[u] = r < n ? rec[(size_t)r * RS] : make_uint4(0xffffffffu, 0, 0, 0);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const unsigned cell = q[u].x != 0xffffffffu ? q[u].x : 0u;
                                const int w = i * words + (int)(cell >> 5);
                                const int vr = (int)s_pre[w] + __popc(s_bm[w] & ((1u << (cell & 31)) - 1u)) - c0;
                                const bool in = q[u].x != 0xffffffffu && vr >= 0 && vr < kAccRows;
                                unsigned* a = s_acc + (in ? vr : 0) * ACCW;
                                // a warp whose 32 records all fall into ONE voxel (dense blobs: hundreds of points per voxel)
                                // sums them with warp reductions: one atomic per field instead of 32 colliding ones
                                const unsigned inm = __ballot_sync(0xffffffffu, in);
                                const int lead = __ffs(inm) - 1;
                                const int lead_vr = __shfl_sync(0xffffffffu, vr, lead < 0 ? 0 : lead);
                                if (__popc(inm) > 4 && __all_sync(0xffffffffu, !in || vr == lead_vr)) {
                                    unsigned long long sum[3];
                                    const unsigned qq[3] = {in ? q[u].y : 0u, in ? q[u].z : 0u, in ? q[u].w : 0u};
#pragma unroll
                                    for (int j = 0; j < 3; ++j) {
                                        const unsigned lo16 = __reduce_add_sync(0xffffffffu, qq[j] & 0xffffu);
                                        const unsigned hi16 = __reduce_add_sync(0xffffffffu, qq[j] >> 16);
                                        sum[j] = ((unsigned long long)hi16 << 16) + lo16;
                                    }
                                    if (lane == lead) {
                                        add_u64_split(a + 0, sum[0]); add_u64_split(a + 2, sum[1]); add_u64_split(a + 4, sum[2]);
                                        atomicAdd(a + 6, (unsigned)__popc(inm));
                                        a[7] = cell;
                                    }
                                } else if (in) {
                                    acc_add<WIDE>(a, q[u], cell);
                                }
                                if (WIDE && in) {
                                    const uint4 x = rec[(size_t)(r0 + u * kRedThreads + (int)threadIdx.x) * RS + 1];
                                    atomicAdd(reinterpret_cast<float*>(a + 8), __uint_as_float(x.x));
                                    atomicAdd(reinterpret_cast<float*>(a + 9), __uint_as_float(x.y));
                                    atomicAdd(reinterpret_cast<float*>(a + 10), __uint_as_float(x.z));
                                    atomicAdd(reinterpret_cast<float*>(a + 11), __uint_as_float(x.w));
                                }
                            }
                        }
                        __syncthreads();
                        for (int v = threadIdx.x; v < nrows; v += kRedThreads)
                            write_voxel_row<WIDE>(out, g, gd, tile_base + s_rows_all[i] + c0 + v, logw, d.x, s_acc[v * ACCW + 7] & 0xffffu,
                                                  s_acc + v * ACCW);
                    }
                }
            }
        }
    }
}

// ref: mean_vfe.py:23-29.  One thread per (voxel, feature).  NP = float (the reference's voxel_num_points after
// load_data_to_gpu) or int (straight from the voxel generator).
template <typename NP>
__global__ void __launch_bounds__(256)
mean_vfe_kernel(int m, int t, int c, const float* __restrict__ voxels, const NP* __restrict__ num_points,
                float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)m * c) return;
    const int v = (int)(e / c), j = (int)(e - (long long)v * c);
    const float* src = voxels + (size_t)v * t * c + j;
    float s = 0.f;
    for (int i = 0; i < t; ++i) s = __fadd_rn(s, src[(size_t)i * c]);   // torch.sum over dim 1, sequential for t<=~32
    const float norm = fmaxf((float)num_points[v], 1.0f);              // clamp_min(1.0)
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
// All kernels take F frames of `stride` rows each (the single-frame entry is F = 1); row q of frame f takes part iff
// q < counts[f] (counts == NULL: all rows).  Voxel ids, caps and outputs are per frame: slot f * max_voxels + id.
__global__ void __launch_bounds__(256)
hardvox_insert_kernel(int frames, int stride, int c, const float* __restrict__ points, const int* __restrict__ counts, VoxGeom g,
                      unsigned long long hmask, unsigned long long* __restrict__ keys, int* __restrict__ first,
                      int* __restrict__ slot_of_point) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)frames * stride) return;
    const int f = (int)(p / stride), q = (int)(p - (long long)f * stride);
    const float* row = points + (size_t)p * c;
    int cx, cy, cz;
    if ((counts && q >= counts[f]) || !voxel_coord(g, row[0], row[1], row[2], cx, cy, cz)) { slot_of_point[p] = -1; return; }
    const unsigned long long key =
        (unsigned long long)((((long long)f * g.g[0] + cx) * g.g[1] + cy) * (long long)g.g[2] + cz);
    unsigned long long slot = mix64(key) & hmask;
    while (true) {
        const unsigned long long prev = atomicCAS(&keys[slot], kEmpty, key);
        if (prev == kEmpty || prev == key) break;
        slot = (slot + 1) & hmask;
    }
    atomicMin(&first[slot], (int)p);
    slot_of_point[p] = (int)slot;
}

__global__ void __launch_bounds__(256)
hardvox_flag_kernel(int n, const int* __restrict__ first, const int* __restrict__ slot_of_point, int* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = slot_of_point[p];
    flag[p] = (s >= 0 && first[s] == p) ? 1 : 0;
}

// coords4 != 0: coordinates (F, max_voxels, 4) [frame, z, y, x] (the collated form); else (max_voxels, 3) zyx.
__global__ void __launch_bounds__(256)
hardvox_assign_kernel(int frames, int stride, int max_voxels, VoxGeom g, const unsigned long long* __restrict__ keys,
                      const int* __restrict__ slot_of_point, const int* __restrict__ flag, const int* __restrict__ rank,
                      int* __restrict__ vid_of_slot, int coords4, int* __restrict__ coordinates, int* __restrict__ num_voxels) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)frames * stride) return;
    const int f = (int)(p / stride), q = (int)(p - (long long)f * stride);
    const int frame_rank0 = rank[(size_t)f * stride];          // voxels opened by earlier frames
    if (q == stride - 1) num_voxels[f] = min(rank[p] + flag[p] - frame_rank0, max_voxels);
    if (!flag[p]) return;
    const int s = slot_of_point[p];
    const int vid = rank[p] - frame_rank0;                     // first-seen order inside the frame
    if (vid >= max_voxels) return;   // vid_of_slot stays -1
    const size_t o = (size_t)f * max_voxels + vid;
    vid_of_slot[s] = (int)o;
    const unsigned long long key = keys[s];
    const int z = (int)(key % (unsigned long long)g.g[2]);
    const int y = (int)((key / (unsigned long long)g.g[2]) % (unsigned long long)g.g[1]);
    const int x = (int)((key / ((unsigned long long)g.g[2] * g.g[1])) % (unsigned long long)g.g[0]);
    if (coords4) reinterpret_cast<int4*>(coordinates)[o] = make_int4(f, z, y, x);
    else { coordinates[o * 3 + 0] = z; coordinates[o * 3 + 1] = y; coordinates[o * 3 + 2] = x; }
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
hardvox_gather_kernel(int frames, int max_voxels, int t, int c, const float* __restrict__ points, const int* __restrict__ num_voxels,
                      const int* __restrict__ sel, const int* __restrict__ cnt, float* __restrict__ voxels,
                      int* __restrict__ num_points_per_voxel) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)frames * max_voxels * t) return;
    const long long slot = e / t;
    const int j = (int)(e - slot * t);
    const int f = (int)(slot / max_voxels), vid = (int)(slot - (long long)f * max_voxels);
    if (vid >= num_voxels[f]) return;
    const int k = min(cnt[slot], t);
    if (j == 0) num_points_per_voxel[slot] = k;
    float* dst = voxels + (size_t)e * c;
    if (j < k) {
        const float* src = points + (size_t)sel[e] * c;
        for (int ff = 0; ff < c; ++ff) dst[ff] = src[ff];
    } else {
        for (int ff = 0; ff < c; ++ff) dst[ff] = 0.f;
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
    SEEVCN_PROF("mean_vfe", as_stream(stream));
    mean_vfe_kernel<float><<<(unsigned)div_up(total, 256ll), 256, 0, as_stream(stream)>>>(
        num_voxels, max_points, num_features, voxels, voxel_num_points, voxel_features);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_mean_vfe_int(int num_voxels, int max_points, int num_features, const float* voxels,
                                   const int* voxel_num_points, float* voxel_features, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_voxels >= 0 && max_points >= 0 && num_features >= 0, "mean_vfe: negative size");
    if (num_voxels == 0 || num_features == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(voxels && voxel_num_points && voxel_features, "mean_vfe: null pointer");
    const long long total = (long long)num_voxels * num_features;
    SEEVCN_PROF("mean_vfe", as_stream(stream));
    mean_vfe_kernel<int><<<(unsigned)div_up(total, 256ll), 256, 0, as_stream(stream)>>>(
        num_voxels, max_points, num_features, voxels, voxel_num_points, voxel_features);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}


namespace {

struct DynWs {
    int logw, nb, scan_tiles, max_tasks, wide, task_weight, max_tile_tasks, max_red_tiles;
    size_t off_meta, off_hist, off_scan_status, off_red_status, zero_bytes, off_tile_first, off_tasks, off_rank, off_records, total;
};

// Bucket width: W = 2^logw consecutive keys per bucket.  1024 keys = 25 y-columns of the Waymo grid at one x: a LiDAR frame
// puts ~20 points into a non-empty bucket, and a bucket's occupancy bitmap is one word per lane of a warp.
int dyn_logw(double) { return 10; }

bool dyn_layout(long long n, int c, int batch, const int* grid, DynWs* out) {
    DynWs w{};
    const double span = (double)(batch > 0 ? batch : 1) * grid[0] * grid[1] * grid[2];
    w.logw = dyn_logw(span);
    const double nbd = span / (double)(1ull << w.logw) + 1.0;
    if (nbd >= 1073741824.0 || n >= (1ll << 30)) return false;
    const long long np = n > 0 ? n : 1;
    w.nb = (int)nbd;
    w.scan_tiles = div_up(w.nb, kScanTile);
    w.max_tasks = (int)std::min<long long>(w.nb, np);
    w.wide = c > 3 ? 1 : 0;
    // tile = kTileWeight of (records + task_weight per bucket): at most max_tile_tasks buckets, whose bitmaps and prefixes
    // (6 B per bitmap word) take 24 KB of shared memory
    const int words = 1 << (w.logw - 5);
    w.max_tile_tasks = (24 * 1024) / (6 * words);                    // 128 buckets at W = 1024
    w.task_weight = kTileWeight / w.max_tile_tasks;                  // 8
    w.max_red_tiles = (int)div_up(np + (long long)w.task_weight * w.max_tasks, (long long)kTileWeight) + 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    w.off_meta = take(64);
    w.off_hist = take(4 * (size_t)w.nb);
    w.off_scan_status = take(8 * (size_t)w.scan_tiles);
    w.off_red_status = take(8 * (size_t)w.max_red_tiles);
    w.zero_bytes = o;                                       // everything above starts from zero (one memset)
    w.off_tile_first = take(8 * ((size_t)w.max_red_tiles + 1));
    w.off_tasks = take(16 * (size_t)w.max_tasks);
    w.off_rank = take(4 * (size_t)np);
    w.off_records = take((w.wide ? 32 : 16) * (size_t)np);
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
    int2* tile_first = reinterpret_cast<int2*>(ws + w.off_tile_first);
    int4* tasks = reinterpret_cast<int4*>(ws + w.off_tasks);
    unsigned* rank = reinterpret_cast<unsigned*>(ws + w.off_rank);
    uint4* records = reinterpret_cast<uint4*>(ws + w.off_records);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(ws, 0, w.zero_bytes, st));
    const int sms = seevcn_num_sms();
    const int grid_pts = (int)std::min<long long>(div_up((long long)num_points, 256ll), (long long)sms * 16);
    {
        SEEVCN_PROF("dynvox_count_kernel", st);
        dynvox_count_kernel<<<grid_pts, 256, 0, st>>>(src, g, num_points, w.logw, hist, rank);
        SEEVCN_LAUNCH_CHECK();
    }
    {
        SEEVCN_PROF("dynvox_scan_kernel", st);
        dynvox_scan_kernel<<<w.scan_tiles, kScanThreads, 0, st>>>(w.nb, hist, scan_status, meta, tasks, kTileWeight, w.task_weight,
                                                                 w.max_red_tiles, tile_first);
        SEEVCN_LAUNCH_CHECK();
    }
    const int words = 1 << (w.logw - 5);
    const int accw = w.wide ? 12 : 8;
    const size_t smem = align_up((size_t)w.max_tile_tasks * (24 + 6 * (size_t)words), 16) + (size_t)kAccRows * accw * 4;
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / kRedThreads, (220 * 1024) / (smem + 1024)));
    const int grid_red = std::min(w.max_red_tiles, sms * ctas_per_sm);
    VoxOut vo{num_features, max_voxels, voxel_coords, voxel_features, voxel_counts};
    if (w.wide) {
        dynvox_scatter_kernel<true><<<grid_pts, 256, 0, st>>>(src, g, gd, num_points, w.logw, hist, rank, records);
        SEEVCN_LAUNCH_CHECK();
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(dynvox_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SEEVCN_PROF("dynvox_reduce_kernel", st);
        dynvox_reduce_kernel<true><<<grid_red, kRedThreads, smem, st>>>(w.logw, g, gd, tasks, meta, tile_first, w.task_weight,
                                                                       w.max_tile_tasks, records, red_status, vo, num_voxels);
    } else {
        {
            SEEVCN_PROF("dynvox_scatter_kernel", st);
            dynvox_scatter_kernel<false><<<grid_pts, 256, 0, st>>>(src, g, gd, num_points, w.logw, hist, rank, records);
            SEEVCN_LAUNCH_CHECK();
        }
        SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(dynvox_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SEEVCN_PROF("dynvox_reduce_kernel", st);
        dynvox_reduce_kernel<false><<<grid_red, kRedThreads, smem, st>>>(w.logw, g, gd, tasks, meta, tile_first, w.task_weight,
                                                                        w.max_tile_tasks, records, red_status, vo, num_voxels);
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

namespace {
int hardvox_run(const char* what, int frames, int stride, int num_features, const float* points, const int* counts,
                const float* pc_range, const float* voxel_size, const int* grid_size, int max_points, int max_voxels,
                int coords4, float* voxels, int* coordinates, int* num_points_per_voxel, int* num_voxels, void* workspace,
                size_t workspace_bytes, cudaStream_t st) {
    const long long n = (long long)frames * stride;
    SEEVCN_REQUIRE(n < (1ll << 31) && (long long)frames * max_voxels * max_points < (1ll << 31), "%s: more than 2^31 rows", what);
    const HardWs w = hard_layout((int)n, max_points, frames * max_voxels);
    if (workspace_bytes < w.total) {
        seevcn_set_error("%s: workspace %zu < %zu", what, workspace_bytes, w.total);
        return SEEVCN_E_WORKSPACE;
    }
    VoxGeom g;
    for (int i = 0; i < 3; ++i) {
        g.lo[i] = pc_range[i]; g.vs[i] = voxel_size[i]; g.g[i] = grid_size[i];
        SEEVCN_REQUIRE(grid_size[i] > 0 && voxel_size[i] > 0.f, "%s: bad grid", what);
    }
    SEEVCN_PROF("hard_voxelize", st);
    char* ws = static_cast<char*>(workspace);
    auto* keys = reinterpret_cast<unsigned long long*>(ws + w.off_keys);
    int* first = reinterpret_cast<int*>(ws + w.off_first);
    int* vid = reinterpret_cast<int*>(ws + w.off_vid);
    int* slot = reinterpret_cast<int*>(ws + w.off_slot);
    int* flag = reinterpret_cast<int*>(ws + w.off_flag);
    int* rank = reinterpret_cast<int*>(ws + w.off_rank);
    int* sel = reinterpret_cast<int*>(ws + w.off_sel);
    int* cnt = reinterpret_cast<int*>(ws + w.off_cnt);
    const size_t mv = (size_t)frames * max_voxels;
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(keys, 0xff, w.nslots * 8, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(first, 0x7f, w.nslots * 4, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(vid, 0xff, w.nslots * 4, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(sel, 0x7f, mv * max_points * 4, st));
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(cnt, 0, mv * 4, st));
    if (frames > 1) SEEVCN_CUDA_CHECK(cudaMemsetAsync(num_points_per_voxel, 0, mv * 4, st));   // padded slots read as empty voxels
    const int gp = (int)div_up(n, 256ll);
    hardvox_insert_kernel<<<gp, 256, 0, st>>>(frames, stride, num_features, points, counts, g, w.nslots - 1, keys, first, slot);
    SEEVCN_LAUNCH_CHECK();
    hardvox_flag_kernel<<<gp, 256, 0, st>>>((int)n, first, slot, flag);
    SEEVCN_LAUNCH_CHECK();
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(ws + w.off_scan, 0, w.scan_bytes, st));
    seevcn_scan::exclusive_scan_u32_kernel<<<w.scan_tiles, seevcn_scan::kScanThreads, 0, st>>>(
        (int)n, reinterpret_cast<const unsigned*>(flag), reinterpret_cast<unsigned*>(rank),
        reinterpret_cast<unsigned long long*>(ws + w.off_scan), reinterpret_cast<int*>(ws + w.off_scan + 8 * (size_t)w.scan_tiles),
        nullptr);
    SEEVCN_LAUNCH_CHECK();
    hardvox_assign_kernel<<<gp, 256, 0, st>>>(frames, stride, max_voxels, g, keys, slot, flag, rank, vid, coords4, coordinates, num_voxels);
    SEEVCN_LAUNCH_CHECK();
    hardvox_pick_kernel<<<gp, 256, 0, st>>>((int)n, max_points, slot, vid, sel, cnt);
    SEEVCN_LAUNCH_CHECK();
    const long long tot = (long long)mv * max_points;
    hardvox_gather_kernel<<<(unsigned)div_up(tot, 256ll), 256, 0, st>>>(frames, max_voxels, max_points, num_features, points,
                                                                        num_voxels, sel, cnt, voxels, num_points_per_voxel);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
}  // namespace

// ---- the two point-level steps in front of the voxel generator (data_processor.py:78-103) ----
namespace {
// ref: common_utils.mask_points_by_range (pcdet/utils/common_utils.py:60-63): x, y inclusive
__global__ void __launch_bounds__(256)
range_flag_kernel(int n, int c, const float* __restrict__ points, float x0, float y0, float x1, float y1, unsigned* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float x = points[(size_t)p * c], y = points[(size_t)p * c + 1];
    flag[p] = (x >= x0 && x <= x1 && y >= y0 && y <= y1) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
range_scatter_kernel(int n, int c, const float* __restrict__ points, const unsigned* __restrict__ flag_in,
                     const unsigned* __restrict__ pos, float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !flag_in[p]) return;
    for (int f = 0; f < c; ++f) out[(size_t)pos[p] * c + f] = points[(size_t)p * c + f];
}
// ref: DataProcessor.shuffle_points (data_processor.py:93-103): points[np.random.permutation(n)], here a keyed bijection
__global__ void __launch_bounds__(256)
shuffle_kernel(int n, int c, unsigned key, const float* __restrict__ points, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned src = feistel_perm((unsigned)j, (unsigned)n, key);
    for (int f = 0; f < c; ++f) out[(size_t)j * c + f] = points[(size_t)src * c + f];
}
}  // namespace

extern "C" size_t seevcn_mask_points_by_range_workspace_bytes(int num_points) {
    const size_t np = (size_t)(num_points > 0 ? num_points : 1);
    return align_up(np * 4, 256) * 2 + align_up(8 * div_up(np, (size_t)seevcn_scan::kScanTile) + 64, 256);
}

extern "C" int seevcn_mask_points_by_range(int num_points, int num_features, const float* points, const float* limit_range,
                                           float* out, int* out_count, void* workspace, size_t workspace_bytes,
                                           seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_points >= 0 && num_features >= 2, "mask_points_by_range: bad sizes");
    SEEVCN_REQUIRE(limit_range && out_count, "mask_points_by_range: null pointer");
    cudaStream_t st = as_stream(stream);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(int), st));
    if (num_points == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && out && workspace, "mask_points_by_range: null pointer");
    if (workspace_bytes < seevcn_mask_points_by_range_workspace_bytes(num_points)) {
        seevcn_set_error("mask_points_by_range: workspace too small");
        return SEEVCN_E_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    const size_t np = (size_t)num_points;
    unsigned* flag = reinterpret_cast<unsigned*>(ws);
    unsigned* pos = reinterpret_cast<unsigned*>(ws + align_up(np * 4, 256));
    char* scan = ws + 2 * align_up(np * 4, 256);
    const int tiles = (int)div_up(np, (size_t)seevcn_scan::kScanTile);
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(scan, 0, 8 * (size_t)tiles + 64, st));
    const int gp = div_up(num_points, 256);
    range_flag_kernel<<<gp, 256, 0, st>>>(num_points, num_features, points, limit_range[0], limit_range[1], limit_range[3],
                                         limit_range[4], flag);
    SEEVCN_LAUNCH_CHECK();
    seevcn_scan::exclusive_scan_u32_kernel<<<tiles, seevcn_scan::kScanThreads, 0, st>>>(
        num_points, flag, pos, reinterpret_cast<unsigned long long*>(scan), reinterpret_cast<int*>(scan + 8 * (size_t)tiles),
        reinterpret_cast<unsigned*>(out_count));
    SEEVCN_LAUNCH_CHECK();
    range_scatter_kernel<<<gp, 256, 0, st>>>(num_points, num_features, points, flag, pos, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_shuffle_points(int num_points, int num_features, unsigned seed, const float* points, float* out,
                                     seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_points >= 0 && num_features >= 1, "shuffle_points: bad sizes");
    if (num_points == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && out && points != out, "shuffle_points: null or aliased pointer");
    shuffle_kernel<<<div_up(num_points, 256), 256, 0, as_stream(stream)>>>(num_points, num_features, mix32(seed ^ 0x5bd1e995u), points, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" unsigned seevcn_shuffle_perm(unsigned j, unsigned n, unsigned seed) {
    return feistel_perm(j, n, mix32(seed ^ 0x5bd1e995u));
}

extern "C" size_t seevcn_hard_voxelize_frames_workspace_bytes(int num_frames, int stride, int max_points, int max_voxels) {
    const long long n = (long long)num_frames * stride;
    if (n >= (1ll << 31) || (long long)num_frames * max_voxels >= (1ll << 31)) return 0;
    return hard_layout((int)n, max_points, num_frames * max_voxels).total;
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
    return hardvox_run("hard_voxelize", 1, num_points, num_features, points, nullptr, pc_range, voxel_size, grid_size, max_points,
                       max_voxels, 0, voxels, coordinates, num_points_per_voxel, num_voxels, workspace, workspace_bytes, st);
}

extern "C" int seevcn_hard_voxelize_frames(int num_frames, int stride, int num_features, const float* points, const int* counts,
                                           const float* pc_range, const float* voxel_size, const int* grid_size,
                                           int max_points, int max_voxels, float* voxels, int* coordinates,
                                           int* num_points_per_voxel, int* num_voxels, void* workspace, size_t workspace_bytes,
                                           seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_frames >= 0 && stride >= 0 && max_points >= 1 && max_voxels >= 0, "hard_voxelize_frames: bad sizes");
    SEEVCN_REQUIRE(num_features >= 3, "hard_voxelize_frames: num_features must be >= 3");
    SEEVCN_REQUIRE(pc_range && voxel_size && grid_size && (num_voxels || num_frames == 0), "hard_voxelize_frames: null pointer");
    cudaStream_t st = as_stream(stream);
    if (num_frames == 0) return SEEVCN_OK;
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(num_voxels, 0, sizeof(int) * (size_t)num_frames, st));
    if (stride == 0 || max_voxels == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && voxels && coordinates && num_points_per_voxel && workspace, "hard_voxelize_frames: null pointer");
    return hardvox_run("hard_voxelize_frames", num_frames, stride, num_features, points, counts, pc_range, voxel_size, grid_size,
                       max_points, max_voxels, 1, voxels, coordinates, num_points_per_voxel, num_voxels, workspace, workspace_bytes, st);
}
