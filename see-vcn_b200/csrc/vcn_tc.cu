// Stage 5 — the VCN shared-MLP layers on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// One generic kernel computes a 1x1-conv / Linear layer in the "channels x points" orientation
//
//     D[c, p] = sum_k W[c, k] * X[p, k]        (M = 128 output channels, N = 256 points, K step 64)
//
// so that in TMEM a LANE is an output channel and the COLUMNS are points.  That makes the fused
// epilogues cheap: bias / per-object bias / activation are per-lane constants, and the max-pool over
// points (torch.max(feature, dim=2), VCN_VC.py:98,102; AdaptiveMaxPool1d :121) is a running max inside
// one thread — no shuffles, no second pass, and the 1024-wide activations are never written.
//
// Structure (persistent, warp specialised, one CTA per SM):
//   warp 0   TMA producer: W tile (128 x 64 bf16) + X tile (256 x 64 bf16) per stage, SWIZZLE_128B,
//            4-stage mbarrier ring (4 x 48 KB)
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 ->
//            fp32), 4 per stage, accumulating into one of TWO 256-column TMEM accumulators
//   warp 2   TMEM allocator (512 columns)
//   warps 4-11 epilogue (two per TMEM lane quadrant, one per column half): tcgen05.ld 32 lanes x 32
//            columns at a time, bias/act/max, bf16 or fp32 stores;
//            overlaps the next tile's MMAs through the double-buffered accumulator
// Tiles that share an X tile are adjacent in the schedule so X is read from HBM once and from L2
// for the other cout/128 - 1 channel tiles.
#include <cuda.h>
#include "vcn_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;      // output channels per tile  (UMMA M)
constexpr int BN = 256;      // points per tile           (UMMA N)
constexpr int BK = 64;       // K elements per stage = 128 bytes of bf16 = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int TMEM_COLS = 512;

struct TcArgs {
    int rows, cout, kblocks, rows_per_obj, act;
    int num_m_tiles, num_tiles;
    const float* bias;       // (cout) or null
    const float* obj_bias;   // (rows / rows_per_obj, cout) or null
    __nv_bfloat16* Y; int ldy;
    float* Yf32; int ldyf;
    float* colmax;           // (rows / rows_per_obj, cout) or null; pre-filled with -inf
    int ksplit, kb_per_split; // split-K (FC layers, few rows): tile t covers k-blocks [ks*kb_per_split, ...) of K
    float* part;             // split-K partial sums (ksplit, rows, cout) fp32, raw accumulators (no bias / act)
};

using namespace tcptx;

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10),
// both K-major, N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__global__ void __launch_bounds__(NUM_THREADS, 1)
vcn_linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t* full = bars;                  // [STAGES]
    uint64_t* empty = bars + STAGES;        // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;    // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();      // everything above overlapped the tail of the preceding kernel

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int ks = tile % a.ksplit, mn = tile / a.ksplit;
                const int m_tile = mn % a.num_m_tiles, n_tile = mn / a.num_m_tiles;
                const int kb0 = ks * a.kb_per_split, kb1 = min(a.kblocks, kb0 + a.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
                    tma_load_2d(smem_a + stage * A_BYTES, &tmap_w, kb * BK, m_tile * BM, &full[stage]);
                    tma_load_2d(smem_b + stage * B_BYTES, &tmap_x, kb * BK, n_tile * BN, &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int ks = tile % a.ksplit;
                const int kb0 = ks * a.kb_per_split, kb1 = min(a.kblocks, kb0 + a.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * A_BYTES);
                    const uint32_t b_addr = smem_u32(smem_b + stage * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = make_desc(a_addr + k * UMMA_K * 2);
                        const uint64_t db = make_desc(b_addr + k * UMMA_K * 2);
                        tc_mma(d_tmem, da, db, kIdesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[stage]);   // frees the smem slot when these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tfull[acc]);         // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // 8 epilogue warps: TMEM lane quadrant q = warp % 4 (hardware restriction), column half h.
        const int q = warp & 3, h = (warp - 4) >> 2;
        // act as one branch-free op: v = max(v, slope * v) — slope 1 = none, 0 = ReLU, 0.01 = LeakyReLU
        const float slope = a.act == ACT_RELU ? 0.f : a.act == ACT_LEAKY ? 0.01f : 1.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int ks = tile % a.ksplit, mn = tile / a.ksplit;
            const int m_tile = mn % a.num_m_tiles, n_tile = mn / a.num_m_tiles;
            const int c = m_tile * BM + q * 32 + lane;
            const int row0 = n_tile * BN;
            const int nvalid = min(BN, a.rows - row0);
            const bool c_ok = c < a.cout;
            const int obj_first = row0 / a.rows_per_obj, obj_last = (row0 + nvalid - 1) / a.rows_per_obj;
            const bool one_obj = obj_first == obj_last && !a.part;   // split-K always takes the generic path
            float bt = (a.bias && c_ok) ? a.bias[c] : 0.f;            // bias (+ per-object bias when the tile is one object)
            if (one_obj && a.obj_bias && c_ok) bt += a.obj_bias[(size_t)obj_first * a.cout + c];
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + h * (BN / 2);
            float run_max = -__int_as_float(0x7f800000);
            int run_obj = obj_first;
            // max-only tiles inside one object track the RAW accumulator: bias and the monotone activation
            // commute with max and are applied once at the end
            const bool raw_max = !a.Y && !a.Yf32 && one_obj;
#pragma unroll 1
            for (int ch = 0; ch < BN / 64; ++ch) {
                const int p0 = h * (BN / 2) + ch * 32;     // first point (column) of this chunk
                if (p0 >= nvalid) break;                   // warp-uniform
                uint32_t r[32];
                tc_ld32(t_base + ch * 32, r);
                if (a.part) {
                    // split-K slice: raw accumulators, one coalesced 128-byte store per row and warp
                    if (c_ok) {
                        float* pp = a.part + ((size_t)ks * a.rows + row0 + p0) * a.cout + c;
                        if (p0 + 32 <= nvalid) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) pp[(size_t)j * a.cout] = __uint_as_float(r[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (p0 + j < nvalid) pp[(size_t)j * a.cout] = __uint_as_float(r[j]);
                        }
                    }
                    continue;
                }
                if (one_obj && c_ok && p0 + 32 <= nvalid) {
                    // fast path: full chunk, one object, valid channel
                    if (a.Y) {
                        __nv_bfloat16* yp = a.Y + (size_t)(row0 + p0) * a.ldy + c;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float v = __uint_as_float(r[j]) + bt;
                            v = fmaxf(v, slope * v);
                            yp[(size_t)j * a.ldy] = __float2bfloat16(v);
                            run_max = fmaxf(run_max, v);
                        }
                    } else if (a.Yf32) {
                        float* yp = a.Yf32 + (size_t)(row0 + p0) * a.ldyf + c;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float v = __uint_as_float(r[j]) + bt;
                            v = fmaxf(v, slope * v);
                            yp[(size_t)j * a.ldyf] = v;
                            run_max = fmaxf(run_max, v);
                        }
                    } else {
                        // max-pool only: bias and the monotone activation commute with max -> applied once below
#pragma unroll
                        for (int j = 0; j < 32; ++j) run_max = fmaxf(run_max, __uint_as_float(r[j]));
                    }
                } else if (c_ok) {
                    // generic path: ragged tail and / or several objects inside the tile (FC layers: one object per row)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {   // unrolled so r[] keeps static register indices
                        const int p = p0 + j;
                        if (p >= nvalid) continue;
                        const int row = row0 + p;
                        const int o = row / a.rows_per_obj;
                        float v = __uint_as_float(r[j]) + bt;
                        if (!one_obj && a.obj_bias) v += a.obj_bias[(size_t)o * a.cout + c];
                        v = fmaxf(v, slope * v);
                        if (a.Y) a.Y[(size_t)row * a.ldy + c] = __float2bfloat16(v);
                        if (a.Yf32) a.Yf32[(size_t)row * a.ldyf + c] = v;
                        if (a.colmax) {
                            if (o != run_obj) {
                                atomic_max_float(&a.colmax[(size_t)run_obj * a.cout + c], run_max);
                                run_obj = o; run_max = -__int_as_float(0x7f800000);
                            }
                            run_max = fmaxf(run_max, raw_max ? __uint_as_float(r[j]) : v);
                        }
                    }
                }
            }
            // release the accumulator before the global atomics
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (a.colmax && c_ok && run_max > -__int_as_float(0x7f800000)) {
                float v = run_max;
                if (raw_max) { v = run_max + bt; v = fmaxf(v, slope * v); }
                atomic_max_float(&a.colmax[(size_t)run_obj * a.cout + c], v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// fp32 3-vector input layer (K = 3 is not a tensor-core shape): Y[r, c] = act(W[c,:] . x[r,:] + b[c]) -> bf16.
// One thread per (row, 8 channels): 16-byte stores.  ref: pose_encoder.0 / mlp_conv1.0, VCN_VC.py:117,86-87
__global__ void __launch_bounds__(256)
pointwise3_kernel(size_t rows, int cout, const float* __restrict__ X, const float* __restrict__ W,
                  const float* __restrict__ bias, int act, __nv_bfloat16* __restrict__ Y, int ldy) {
    // cout * 4 floats (w0 w1 w2 b), plus 4 floats of padding per group of 8 channels: lanes of a warp read
    // 16 different groups at once, and an unpadded 128-byte group stride would put them all in one bank
    extern __shared__ float s_w[];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) {
        float* d = s_w + i * 4 + (i >> 3) * 4;
        d[0] = W[i * 3 + 0]; d[1] = W[i * 3 + 1]; d[2] = W[i * 3 + 2];
        d[3] = bias ? bias[i] : 0.f;
    }
    __syncthreads();
    const int groups = cout / 8;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * groups) return;
    const size_t r = e / groups;
    const int g = (int)(e - r * groups);
    const float x = X[r * 3 + 0], y = X[r * 3 + 1], z = X[r * 3 + 2];
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(&s_w[(g * 8 + j) * 4 + g * 4]);
        o[j] = __float2bfloat16(apply_act(fmaf(w.z, z, fmaf(w.y, y, fmaf(w.x, x, w.w))), act));
    }
    *reinterpret_cast<uint4*>(Y + r * ldy + g * 8) = *reinterpret_cast<const uint4*>(o);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 row-major (rows, kpad) tensor, box = 64 elements (128 B) x box_rows, SWIZZLE_128B, zero OOB fill
int make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t kpad, uint64_t ld_elems, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { seevcn_set_error("cuTensorMapEncodeTiled not available from the driver"); return SEEVCN_E_CUDA; }
    cuuint64_t dims[2] = {kpad, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { seevcn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SEEVCN_E_CUDA; }
    return SEEVCN_OK;
}

}  // namespace

// X (rows, ldx) bf16 with ldx >= L.kpad and columns [cin, kpad) zero.
int vcn_linear_tc(const LinearW& L, int rows, const __nv_bfloat16* X, int ldx, const float* obj_bias, int rows_per_obj,
                  int act, __nv_bfloat16* Y, int ldy, float* Yf32, float* colmax, cudaStream_t st) {
    if (rows == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(L.kpad % BK == 0 && ldx >= L.kpad, "vcn_linear_tc: K=%d must be padded to %d (ldx %d)", L.cin, BK, ldx);
    SEEVCN_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx * 2) % 16 == 0, "vcn_linear_tc: X not 16-byte aligned");
    // per device and cheap: set on every launch (a process-wide "done" flag would skip the second GPU of a process)
    SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(vcn_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CUtensorMap tw, tx;
    int rc = make_tmap(&tw, L.w16, (uint64_t)L.cout, (uint64_t)L.kpad, (uint64_t)L.kpad, BM);
    if (rc != SEEVCN_OK) return rc;
    rc = make_tmap(&tx, X, (uint64_t)rows, (uint64_t)L.kpad, (uint64_t)ldx, BN);
    if (rc != SEEVCN_OK) return rc;
    TcArgs a{};
    a.rows = rows; a.cout = L.cout; a.kblocks = L.kpad / BK; a.rows_per_obj = rows_per_obj; a.act = act;
    a.num_m_tiles = div_up(L.cout, BM);
    a.num_tiles = a.num_m_tiles * div_up(rows, BN);
    a.bias = L.b; a.obj_bias = obj_bias; a.Y = Y; a.ldy = ldy; a.Yf32 = Yf32; a.ldyf = L.cout; a.colmax = colmax;
    a.ksplit = 1; a.kb_per_split = a.kblocks; a.part = nullptr;
    const int grid = a.num_tiles < seevcn_num_sms() ? a.num_tiles : seevcn_num_sms();
    vcn_linear_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tw, tx, a);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

int vcn_pointwise3(const LinearW& L, size_t rows, const float* X, int act, __nv_bfloat16* Y, int ldy, cudaStream_t st) {
    if (rows == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(L.cin == 3 && L.cout % 8 == 0 && ldy % 8 == 0, "vcn_pointwise3: needs cin == 3, cout %% 8 == 0");
    const size_t total = rows * (size_t)(L.cout / 8);
    pointwise3_kernel<<<(unsigned)div_up(total, (size_t)256), 256, L.cout * 16 + (L.cout / 8) * 16, st>>>(rows, L.cout, X, L.w, L.b, act, Y, ldy);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}


// ---- per-object FC layers (few rows, K up to 1024): split-K over the SMs + a deterministic reduce ----
namespace {
// Y = act(sum_ks part[ks] + bias); writes fp32 (optional) and a bf16 copy padded to ldb (the next layer's X)
__global__ void __launch_bounds__(256)
fc_reduce_kernel(int rows, int cout, int ksplit, const float* __restrict__ part, const float* __restrict__ bias, int act,
                 float* __restrict__ Yf32, __nv_bfloat16* __restrict__ Yb16, int ldb) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int groups = cout / 4;
    if (e >= rows * groups) return;
    const int r = e / groups, c = (e - r * groups) * 4;
    float4 s = *reinterpret_cast<const float4*>(part + (size_t)r * cout + c);
    for (int k = 1; k < ksplit; ++k) {      // fixed order: bitwise reproducible
        const float4 t = *reinterpret_cast<const float4*>(part + ((size_t)k * rows + r) * cout + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + c); s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w; }
    s.x = apply_act(s.x, act); s.y = apply_act(s.y, act); s.z = apply_act(s.z, act); s.w = apply_act(s.w, act);
    if (Yf32) *reinterpret_cast<float4*>(Yf32 + (size_t)r * cout + c) = s;
    if (Yb16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(s.x, s.y), hi = __floats2bfloat162_rn(s.z, s.w);
        uint2 v; v.x = *reinterpret_cast<uint32_t*>(&lo); v.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(Yb16 + (size_t)r * ldb + c) = v;
    }
}
}  // namespace

static int fc_ksplit(int rows, int cout, int kblocks, int* kb_per_split) {
    const int mn = div_up(cout, BM) * div_up(rows, BN);
    int ksplit = seevcn_num_sms() / mn;                       // fill the machine once
    ksplit = ksplit < 1 ? 1 : ksplit > 16 ? 16 : ksplit;
    ksplit = ksplit > kblocks ? kblocks : ksplit;
    *kb_per_split = div_up(kblocks, ksplit);
    return div_up(kblocks, *kb_per_split);
}
size_t vcn_fc_part_bytes(int rows, int cout, int cin) {
    int kbps;
    const int ks = fc_ksplit(rows, cout, (int)align_up((size_t)cin, BK) / BK, &kbps);
    return (size_t)ks * rows * cout * 4;
}

// X bf16 (rows, ldx >= kpad, zero padded).  part: >= vcn_fc_part_bytes(rows, cout, cin).  cout % 128 == 0.
// Leaves the *ksplit split-K slices (slice k at part + k * rows * cout, no bias, no activation) for the consumer to sum.
int vcn_fc_tc_partials(const LinearW& L, int rows, const __nv_bfloat16* X, int ldx, float* part, int* ksplit, cudaStream_t st) {
    *ksplit = 0;
    if (rows == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(L.kpad % BK == 0 && ldx >= L.kpad && L.cout % 4 == 0, "vcn_fc_tc: bad layer shape");
    // per device and cheap: set on every launch (a process-wide "done" flag would skip the second GPU of a process)
    SEEVCN_CUDA_CHECK(cudaFuncSetAttribute(vcn_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CUtensorMap tw, tx;
    int rc = make_tmap(&tw, L.w16, (uint64_t)L.cout, (uint64_t)L.kpad, (uint64_t)L.kpad, BM);
    if (rc != SEEVCN_OK) return rc;
    rc = make_tmap(&tx, X, (uint64_t)rows, (uint64_t)L.kpad, (uint64_t)ldx, BN);
    if (rc != SEEVCN_OK) return rc;
    TcArgs a{};
    a.rows = rows; a.cout = L.cout; a.kblocks = L.kpad / BK; a.rows_per_obj = 1; a.act = ACT_NONE;
    a.num_m_tiles = div_up(L.cout, BM);
    const int mn = a.num_m_tiles * div_up(rows, BN);
    a.ksplit = fc_ksplit(rows, L.cout, a.kblocks, &a.kb_per_split);
    a.num_tiles = mn * a.ksplit;
    a.part = part;
    const int grid = a.num_tiles < seevcn_num_sms() ? a.num_tiles : seevcn_num_sms();
    SEEVCN_CUDA_CHECK(launch_pdl(vcn_linear_tc_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, tw, tx, a));
    SEEVCN_LAUNCH_CHECK();
    *ksplit = a.ksplit;
    return SEEVCN_OK;
}

int vcn_fc_tc(const LinearW& L, int rows, const __nv_bfloat16* X, int ldx, int act, float* Yf32, __nv_bfloat16* Yb16, int ldb,
              float* part, cudaStream_t st) {
    if (rows == 0) return SEEVCN_OK;
    int ksplit = 0;
    const int rc = vcn_fc_tc_partials(L, rows, X, ldx, part, &ksplit, st);
    if (rc != SEEVCN_OK) return rc;
    const int total = rows * (L.cout / 4);
    SEEVCN_CUDA_CHECK(launch_pdl(fc_reduce_kernel, dim3(div_up(total, 256)), dim3(256), 0, st, rows, L.cout, ksplit,
                                 static_cast<const float*>(part), L.b, act, Yf32, Yb16, ldb));
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

// ---- standalone layer op (C-ABI): lets the parity tests drive the tcgen05 kernel on its own ----
namespace {
__global__ void pack_bf16_kernel(size_t rows, int cols, const float* __restrict__ src, int lds,
                                 __nv_bfloat16* __restrict__ dst, int ldd) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * (size_t)ldd) return;
    const size_t r = e / ldd; const int c = (int)(e - r * ldd);
    dst[e] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
}
}  // namespace

extern "C" size_t seevcn_linear_bf16_workspace_bytes(int rows, int cin, int cout) {
    const size_t kpad = align_up((size_t)(cin > 0 ? cin : 1), BK);
    return align_up((size_t)rows * kpad * 2, 256) + align_up((size_t)cout * kpad * 2, 256) + 256;
}

extern "C" int seevcn_linear_bf16(int rows, int cin, int cout, const float* X, const float* W, const float* bias,
                                  const float* obj_bias, int rows_per_obj, int act, float* Y, float* colmax,
                                  void* workspace, size_t workspace_bytes, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(rows >= 0 && cin >= 1 && cout >= 1 && rows_per_obj >= 1, "linear_bf16: bad sizes");
    SEEVCN_REQUIRE(act >= 0 && act <= 2, "linear_bf16: act must be 0 (none), 1 (relu) or 2 (leaky relu)");
    if (rows == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(X && W && workspace && (Y || colmax), "linear_bf16: null pointer");
    if (workspace_bytes < seevcn_linear_bf16_workspace_bytes(rows, cin, cout)) {
        seevcn_set_error("linear_bf16: workspace too small");
        return SEEVCN_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int kpad = (int)align_up((size_t)cin, BK);
    auto* xb = static_cast<__nv_bfloat16*>(workspace);
    auto* wb = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(workspace) + align_up((size_t)rows * kpad * 2, 256));
    pack_bf16_kernel<<<(unsigned)div_up((size_t)rows * kpad, (size_t)256), 256, 0, st>>>(rows, cin, X, cin, xb, kpad);
    SEEVCN_LAUNCH_CHECK();
    pack_bf16_kernel<<<(unsigned)div_up((size_t)cout * kpad, (size_t)256), 256, 0, st>>>(cout, cin, W, cin, wb, kpad);
    SEEVCN_LAUNCH_CHECK();
    LinearW L;
    L.w = W; L.b = bias; L.w16 = wb; L.cin = cin; L.cout = cout; L.ldw = cin; L.kpad = kpad;
    return vcn_linear_tc(L, rows, xb, kpad, obj_bias, rows_per_obj, act, nullptr, 0, Y, colmax, st);
}
