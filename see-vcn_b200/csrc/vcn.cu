// Stages 2+5 — VCN forward: viewer-centred canonicalisation + shared-MLP encoder + FC decoder.
//
// Replaces VCN_VC.forward (see/surface_completion/models/vcn/models/VCN_VC.py:178-213),
// VCN_CN.forward (VCN_CN.py:142-157), FeatureEncoder.forward (VCN_VC.py:95-106) and the
// transforms of utils/transform.py:33-58,91-161.
//
// This file holds the model container, the per-object frame/pose/transform kernels and
// the fp32 SIMT validation path (precision=1).  The bf16 tcgen05 path (precision=0) lives
// in vcn_tc.cu and plugs in through vcn_linear_tc().
//
// Restructuring relative to the reference graph (all exact in real arithmetic):
//  * BatchNorm (eval) is folded into the preceding conv by the caller;
//  * torch.cat([global.expand, local]) -> conv 512->512 is split: the global half of the
//    weight times the per-object global feature becomes a per-object bias (VCN_VC.py:100-101);
//  * max-pool over points is fused into the producing layer's epilogue — the (B,1024,N)
//    activations of pose_encoder.4 / mlp_conv2.3 are never written;
//  * rotate/centre/canonicalise are fused into the kernels that read or write the points.
#include <vector>
#include <stdlib.h>
#include "vcn_common.cuh"

int vcn_linear_tc(const LinearW& L, int rows, const __nv_bfloat16* X, int ldx, const float* obj_bias,
                  int rows_per_obj, int act, __nv_bfloat16* Y, int ldy, float* Yf32, float* colmax,
                  cudaStream_t st);   // vcn_tc.cu
int vcn_pointwise3(const LinearW& L, size_t rows, const float* X, int act, __nv_bfloat16* Y, int ldy, cudaStream_t st);
size_t vcn_fc_part_bytes(int rows, int cout, int cin);
int vcn_fc_tc(const LinearW& L, int rows, const __nv_bfloat16* X, int ldx, int act, float* Yf32, __nv_bfloat16* Yb16, int ldb,
              float* part, cudaStream_t st);
// split-K partial sums only (no reduce launch): the consumer sums the *ksplit slices of `part` in index order and adds L.b
int vcn_fc_tc_partials(const LinearW& L, int rows, const __nv_bfloat16* X, int ldx, float* part, int* ksplit, cudaStream_t st);
// vcn_chain.cu: fused per-point chains (tcgen05, activations kept in TMEM)
int vcn_chain_pose(const seevcn_vcn_model* M, int num_obj, int n, const float* input, const VcnFrame* frames,
                   float* pose_feat, cudaStream_t st);
int vcn_chain_enc1(const seevcn_vcn_model* M, int num_obj, int n, const float* input, const VcnFrame* frames,
                   const VcnPose* poses, __nv_bfloat16* F, float* g256, cudaStream_t st);
int vcn_chain_enc2(const seevcn_vcn_model* M, int num_obj, int n, const __nv_bfloat16* F, const float* obj_bias_part,
                   int obj_bias_splits, float* feat, cudaStream_t st);

namespace {

// ---------------------------------------------------------------- per-object kernels --

struct VcnMaxInit { float* dst[3] = {nullptr, nullptr, nullptr}; int width[3] = {0, 0, 0}; };

// One CTA per object.  VC: theta = atan2(mean y, mean x); a = -theta; fview = p . R(a);
// mean = mean(fview); writes centred = fview - mean  (VCN_VC.py:185-190).
// CN: centre = gt[:3]; a = -heading; pc = rotate(p - centre, a) / length  (VCN_CN.py:146-147).
// out rows (n,3) fp32: VC -> centred cloud (pose-encoder input); CN -> canonical cloud (encoder input).
__global__ void __launch_bounds__(256)
vcn_frame_kernel(int n, int viewer_centred, const float* __restrict__ input, const float* __restrict__ gt_boxes,
                 VcnFrame* __restrict__ frames, float* __restrict__ out, VcnMaxInit init = VcnMaxInit{}) {
    __shared__ float red[3][8];
    __shared__ VcnFrame fr;
    const int o = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    // -inf into this object's rows of the max-pool targets of the chains that follow (saves their fill launches)
#pragma unroll
    for (int t = 0; t < 3; ++t)
        if (init.dst[t])
            for (int i = threadIdx.x; i < init.width[t]; i += 256) init.dst[t][(size_t)o * init.width[t] + i] = -__builtin_huge_valf();
    const float* p = input + (size_t)o * n * 3;
    float* q = out + (size_t)o * n * 3;
    auto block_sum3 = [&](float a, float b, float c, float& ra, float& rb, float& rc) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
            c += __shfl_xor_sync(0xffffffffu, c, off);
        }
        __syncthreads();
        if (lane_id() == 0) { red[0][warp_id()] = a; red[1][warp_id()] = b; red[2][warp_id()] = c; }
        __syncthreads();
        ra = rb = rc = 0.f;
        for (int w = 0; w < 8; ++w) { ra += red[0][w]; rb += red[1][w]; rc += red[2][w]; }
    };
    if (viewer_centred) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) { sx += p[i * 3 + 0]; sy += p[i * 3 + 1]; }
        float mx, my, mz;
        block_sum3(sx, sy, sz, mx, my, mz);
        mx /= (float)n; my /= (float)n;
        const float theta = atan2f(my, mx);
        const float ca = cosf(-theta), sa = sinf(-theta);
        sx = sy = sz = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) {
            const float x = p[i * 3 + 0], y = p[i * 3 + 1], z = p[i * 3 + 2];
            sx += x * ca - y * sa; sy += x * sa + y * ca; sz += z;
        }
        float m0, m1, m2;
        block_sum3(sx, sy, sz, m0, m1, m2);
        m0 /= (float)n; m1 /= (float)n; m2 /= (float)n;
        if (threadIdx.x == 0) {
            fr.ca = ca; fr.sa = sa; fr.theta = theta; fr.scale = 1.f;
            fr.mean[0] = m0; fr.mean[1] = m1; fr.mean[2] = m2; fr.pad = 0.f;
            frames[o] = fr;
        }
        if (out) for (int i = threadIdx.x; i < n; i += 256) {
            const float x = p[i * 3 + 0], y = p[i * 3 + 1], z = p[i * 3 + 2];
            q[i * 3 + 0] = (x * ca - y * sa) - m0;
            q[i * 3 + 1] = (x * sa + y * ca) - m1;
            q[i * 3 + 2] = z - m2;
        }
    } else {
        const float* g = gt_boxes + (size_t)o * 7;
        const float c0 = g[0], c1 = g[1], c2 = g[2], len = g[3], h = g[6];
        const float ca = cosf(-h), sa = sinf(-h);
        if (threadIdx.x == 0) {
            fr.ca = ca; fr.sa = sa; fr.theta = h; fr.scale = len;
            fr.mean[0] = c0; fr.mean[1] = c1; fr.mean[2] = c2; fr.pad = 0.f;
            frames[o] = fr;
        }
        if (out) for (int i = threadIdx.x; i < n; i += 256) {
            const float x = p[i * 3 + 0] - c0, y = p[i * 3 + 1] - c1, z = p[i * 3 + 2] - c2;
            q[i * 3 + 0] = (x * ca - y * sa) / len;
            q[i * 3 + 1] = (x * sa + y * ca) / len;
            q[i * 3 + 2] = z / len;
        }
    }
}

// rel_pose (9) -> centre, rot (Gram-Schmidt, VCN_VC.py:12-49), and the reg_rot / reg_centre outputs (VCN_VC.py:211-212).
__device__ void pose_from_rel(int o, const float* r, const VcnFrame& f, VcnPose* __restrict__ poses,
                              float* __restrict__ reg_rot, float* __restrict__ reg_centre) {
    VcnPose P;
    P.centre[0] = f.mean[0] + r[0]; P.centre[1] = f.mean[1] + r[1]; P.centre[2] = f.mean[2] + r[2];
    float x[3] = {r[3], r[4], r[5]}, yr[3] = {r[6], r[7], r[8]}, z[3], y[3];
    float m = fmaxf(sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]), 1e-8f);
    x[0] /= m; x[1] /= m; x[2] /= m;
    z[0] = x[1] * yr[2] - x[2] * yr[1]; z[1] = x[2] * yr[0] - x[0] * yr[2]; z[2] = x[0] * yr[1] - x[1] * yr[0];
    m = fmaxf(sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]), 1e-8f);
    z[0] /= m; z[1] /= m; z[2] /= m;
    y[0] = z[1] * x[2] - z[2] * x[1]; y[1] = z[2] * x[0] - z[0] * x[2]; y[2] = z[0] * x[1] - z[1] * x[0];
    for (int i = 0; i < 3; ++i) { P.rot[i * 3 + 0] = x[i]; P.rot[i * 3 + 1] = y[i]; P.rot[i * 3 + 2] = z[i]; }
    P.pad0 = 0.f; P.pad1[0] = P.pad1[1] = P.pad1[2] = 0.f;
    poses[o] = P;
    const float ct = cosf(f.theta), st = sinf(f.theta);
    if (reg_rot) {   // rot . [[c,s,0],[-s,c,0],[0,0,1]]
        float* R = reg_rot + (size_t)o * 9;
        for (int i = 0; i < 3; ++i) {
            R[i * 3 + 0] = P.rot[i * 3 + 0] * ct - P.rot[i * 3 + 1] * st;
            R[i * 3 + 1] = P.rot[i * 3 + 0] * st + P.rot[i * 3 + 1] * ct;
            R[i * 3 + 2] = P.rot[i * 3 + 2];
        }
    }
    if (reg_centre) {
        float* c = reg_centre + (size_t)o * 3;
        c[0] = P.centre[0] * ct - P.centre[1] * st;
        c[1] = P.centre[0] * st + P.centre[1] * ct;
        c[2] = P.centre[2];
    }
}

// Thread per object (layer-wise paths).
__global__ void vcn_pose_kernel(int num_obj, const float* __restrict__ rel_pose, int ld, const VcnFrame* __restrict__ frames,
                                VcnPose* __restrict__ poses, float* __restrict__ reg_rot, float* __restrict__ reg_centre) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= num_obj) return;
    pose_from_rel(o, rel_pose + (size_t)o * ld, frames[o], poses, reg_rot, reg_centre);
}

// The tail of the pose branch in one launch, one CTA (288 threads = 9 warps) per object:
// h = leaky(sum of the split-K slices of pose_fc.0 + b0)  (the slices summed in index order: what fc_reduce_kernel does),
// rel = W2 h + b2 (pose_fc.2, 512 -> 9: warp j owns output j, lanes split K as linear_skinny_kernel does), then the pose.
constexpr int kPoseHidden = 512;
__global__ void __launch_bounds__(288)
vcn_pose_tail_kernel(int num_obj, int ksplit, const float* __restrict__ part, const float* __restrict__ b0,
                     const float* __restrict__ w2, const float* __restrict__ b2, const VcnFrame* __restrict__ frames,
                     VcnPose* __restrict__ poses, float* __restrict__ reg_rot, float* __restrict__ reg_centre) {
    __shared__ float h[kPoseHidden];
    __shared__ float rel[9];
    const int o = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    for (int c = threadIdx.x; c < kPoseHidden; c += 288) {
        float s = part[(size_t)o * kPoseHidden + c];
        for (int k = 1; k < ksplit; ++k) s += part[((size_t)k * num_obj + o) * kPoseHidden + c];
        if (b0) s += b0[c];
        h[c] = apply_act(s, ACT_LEAKY);
    }
    __syncthreads();
    const int j = warp_id();
    float acc = 0.f;
    for (int k = lane_id(); k < kPoseHidden; k += 32) acc = fmaf(h[k], w2[(size_t)j * kPoseHidden + k], acc);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane_id() == 0) rel[j] = acc + (b2 ? b2[j] : 0.f);
    __syncthreads();
    if (threadIdx.x == 0) pose_from_rel(o, rel, frames[o], poses, reg_rot, reg_centre);
}

// pc_cn = (fview - centre) . rot^T  (VCN_VC.py:200); fview recomputed from the raw input.
__global__ void __launch_bounds__(256)
vcn_canon_kernel(int n, const float* __restrict__ input, const VcnFrame* __restrict__ frames,
                 const VcnPose* __restrict__ poses, float* __restrict__ out) {
    const int o = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const VcnFrame f = frames[o];
    const VcnPose P = poses[o];
    const float* p = input + ((size_t)o * n + i) * 3;
    const float x = p[0], y = p[1], z = p[2];
    const float d0 = (x * f.ca - y * f.sa) - P.centre[0];
    const float d1 = (x * f.sa + y * f.ca) - P.centre[1];
    const float d2 = z - P.centre[2];
    float* q = out + ((size_t)o * n + i) * 3;
    q[0] = d0 * P.rot[0] + d1 * P.rot[1] + d2 * P.rot[2];
    q[1] = d0 * P.rot[3] + d1 * P.rot[4] + d2 * P.rot[5];
    q[2] = d0 * P.rot[6] + d1 * P.rot[7] + d2 * P.rot[8];
}

// coarse (canonical) -> sensor view.  VC: (c . rot + centre) rotated by +theta (VCN_VC.py:205-208).
// CN: (c * length) rotated by +heading, + centre (VCN_CN.py:153-155).
__global__ void __launch_bounds__(256)
vcn_output_kernel(int m, int viewer_centred, const float* __restrict__ coarse_cn, const VcnFrame* __restrict__ frames,
                  const VcnPose* __restrict__ poses, float* __restrict__ coarse, int ksplit = 0,
                  const float* __restrict__ bias = nullptr) {
    const int o = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= m) return;
    const VcnFrame f = frames[o];
    float cs[3];
    const float* c = coarse_cn + ((size_t)o * m + i) * 3;
    if (ksplit > 0) {   // coarse_cn holds the split-K slices of shape_fc.4: sum them in index order, add the bias
        const size_t slice = (size_t)gridDim.y * m * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float s = c[d];
            for (int k = 1; k < ksplit; ++k) s += c[k * slice + d];
            cs[d] = s + (bias ? bias[i * 3 + d] : 0.f);
        }
        c = cs;
    }
    const float ct = cosf(f.theta), st = sinf(f.theta);
    float v0, v1, v2;
    if (viewer_centred) {
        const VcnPose P = poses[o];
        const float w0 = c[0] * P.rot[0] + c[1] * P.rot[3] + c[2] * P.rot[6] + P.centre[0];
        const float w1 = c[0] * P.rot[1] + c[1] * P.rot[4] + c[2] * P.rot[7] + P.centre[1];
        const float w2 = c[0] * P.rot[2] + c[1] * P.rot[5] + c[2] * P.rot[8] + P.centre[2];
        v0 = w0 * ct - w1 * st; v1 = w0 * st + w1 * ct; v2 = w2;
    } else {
        const float w0 = c[0] * f.scale, w1 = c[1] * f.scale, w2 = c[2] * f.scale;
        v0 = (w0 * ct - w1 * st) + f.mean[0]; v1 = (w0 * st + w1 * ct) + f.mean[1]; v2 = w2 + f.mean[2];
    }
    float* q = coarse + ((size_t)o * m + i) * 3;
    q[0] = v0; q[1] = v1; q[2] = v2;
}

__global__ void fill_kernel(size_t n, float v, float* p) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void bf16_to_f32_kernel(size_t n, const __nv_bfloat16* __restrict__ src, float* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __bfloat162float(src[i]);
}

__global__ void f32_to_bf16_kernel(size_t rows, int cols, const float* __restrict__ src, int lds,
                                   __nv_bfloat16* __restrict__ dst, int ldd) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (e >= rows * (size_t)ldd) return;
    const size_t r = e / ldd; const int c = (int)(e - r * ldd);
    dst[e] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
}

// -------------------------------------------------------------- fp32 SIMT linear layer --
// Y[r, c] = act(sum_k X[r,k] W[c,k] + bias[c] + obj_bias[r / rows_per_obj, c]); optional column max per object.
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
linear_f32_kernel(int rows, int cin, int cout, const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                  const float* __restrict__ bias, const float* __restrict__ obj_bias, int rows_per_obj, int act,
                  float* __restrict__ Y, int ldy, float* __restrict__ colmax) {
    __shared__ __align__(16) float Xs[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < cin; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + i * 256;
            const int r = e / BK, k = e % BK;
            const int gr = row0 + r, gc = col0 + r, gk = k0 + k;
            Xs[k][r] = (gr < rows && gk < cin) ? X[(size_t)gr * ldx + gk] : 0.f;
            Ws[k][r] = (gc < cout && gk < cin) ? W[(size_t)gc * ldw + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = col0 + tx * 4 + j;
        if (c >= cout) continue;
        const float bc = bias ? bias[c] : 0.f;
        float best = 0.f; int best_obj = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + ty * 4 + i;
            if (r >= rows) continue;
            const int obj = r / rows_per_obj;
            float v = acc[i][j] + bc;
            if (obj_bias) v += obj_bias[(size_t)obj * cout + c];
            v = apply_act(v, act);
            if (Y) Y[(size_t)r * ldy + c] = v;
            if (colmax) {
                if (obj != best_obj) {
                    if (best_obj >= 0) atomic_max_float(&colmax[(size_t)best_obj * cout + c], best);
                    best_obj = obj; best = v;
                } else best = fmaxf(best, v);
            }
        }
        if (colmax && best_obj >= 0) atomic_max_float(&colmax[(size_t)best_obj * cout + c], best);
    }
}

// Skinny layer (cout <= 16, e.g. pose_fc.2: 512 -> 9): one warp per row, lanes split K, shuffle reduce.
__global__ void __launch_bounds__(256)
linear_skinny_kernel(int rows, int cin, int cout, const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                     const float* __restrict__ bias, int act, float* __restrict__ Y, int ldy) {
    const int row = blockIdx.x * 8 + warp_id();
    if (row >= rows) return;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    const float* x = X + (size_t)row * ldx;
#pragma unroll 4
    for (int k = lane_id(); k < cin; k += 32) {
        const float xv = x[k];
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < cout) acc[c] = fmaf(xv, W[(size_t)c * ldw + k], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
        if (lane_id() == 0 && c < cout) Y[(size_t)row * ldy + c] = apply_act(acc[c] + (bias ? bias[c] : 0.f), act);
    }
}

int linear_f32(const LinearW& L, int rows, const float* X, int ldx, const float* obj_bias, int rows_per_obj, int act,
               float* Y, int ldy, float* colmax, cudaStream_t st) {
    if (rows == 0) return SEEVCN_OK;
    if (L.cout <= 16 && !obj_bias && !colmax && Y) {
        linear_skinny_kernel<<<div_up(rows, 8), 256, 0, st>>>(rows, L.cin, L.cout, X, ldx, L.w, L.ldw, L.b, act, Y, ldy);
        SEEVCN_LAUNCH_CHECK();
        return SEEVCN_OK;
    }
    dim3 grid(div_up(L.cout, BN), div_up(rows, BM));
    linear_f32_kernel<<<grid, 256, 0, st>>>(rows, L.cin, L.cout, X, ldx, L.w, L.ldw, L.b, obj_bias, rows_per_obj, act, Y,
                                            ldy, colmax);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

int fill(float* p, size_t n, float v, cudaStream_t st) {
    if (n == 0) return SEEVCN_OK;
    fill_kernel<<<(unsigned)div_up(n, (size_t)256), 256, 0, st>>>(n, v, p);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

// ----------------------------------------------------------------- workspace layout --
constexpr int kChunkObjF32 = 16;   // objects per pass of the fp32 path (activations 2 x 32 MB at N=1024)
constexpr int kChunkObjTC = 512;   // bf16 path: objects per pass (enc1 -> global FC -> enc2).  One pass for a typical batch: the
                                   // persistent chain kernels get long tile queues and the per-pass FC launches are not repeated;
                                   // enc2 streams enc1's bf16 output (0.5 MB/object) from HBM under its MMAs either way (ncu:
                                   // 128-object passes did not keep it in L2)

struct VcnWs {
    size_t frames, poses, pts3, pose_feat, h512, rel, g256, objbias, feat, fc_a, fc_b, coarse_cn, act_a, act_b,
        fcx_a, fcx_b, fc_part, total;
    int chunk;
};

VcnWs vcn_ws(int num_obj, int n, int num_coarse, int precision) {
    VcnWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    const size_t B = (size_t)(num_obj > 0 ? num_obj : 1);
    w.chunk = precision == 1 ? kChunkObjF32 : kChunkObjTC;
    if ((size_t)w.chunk > B) w.chunk = (int)B;
    const size_t rows = (size_t)w.chunk * n;
    w.frames = take(B * sizeof(VcnFrame));
    w.poses = take(B * sizeof(VcnPose));
    w.pts3 = take(rows * 3 * 4);
    w.pose_feat = take(B * 1024 * 4);
    w.h512 = take(B * 512 * 4);
    w.rel = take(B * 16 * 4);
    w.g256 = take(B * 256 * 4);
    w.objbias = take(B * 512 * 4);
    w.feat = take(B * 1024 * 4);
    w.fc_a = take(B * 1024 * 4);
    w.fc_b = take(B * 1024 * 4);
    w.coarse_cn = take(B * (size_t)num_coarse * 3 * 4);
    const size_t esz = precision == 1 ? 4 : 2;
    w.act_a = take(rows * 512 * esz);
    w.act_b = take(rows * 512 * esz);
    w.fcx_a = take(B * 1024 * 2);    // bf16 path: bf16 copies of the per-object feature vectors
    w.fcx_b = take(B * 1024 * 2);
    {   // split-K partial sums of the per-object FC layers: the largest of the layers that use them
        size_t pb = vcn_fc_part_bytes((int)B, 1024, 1024);
        const size_t p4 = vcn_fc_part_bytes((int)B, 3 * num_coarse, 1024), p5 = vcn_fc_part_bytes((int)B, 512, 1024);
        pb = p4 > pb ? p4 : pb; pb = p5 > pb ? p5 : pb;
        w.fc_part = take(pb);
    }
    w.total = o;
    return w;
}

}  // namespace

// ------------------------------------------------------------------------- C ABI -----

extern "C" int seevcn_vcn_create(const seevcn_vcn_params* p, seevcn_vcn_model** out_model, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(p && out_model, "vcn_create: null pointer");
    SEEVCN_REQUIRE(p->num_coarse > 0, "vcn_create: num_coarse must be > 0");
    struct Src { LinearW* dst; const float* w; const float* b; int cout, cin, ldw, col0; };
    auto* m = new seevcn_vcn_model();
    m->num_coarse = p->num_coarse;
    m->viewer_centred = p->viewer_centred;
    std::vector<Src> srcs;
    if (p->viewer_centred) {
        SEEVCN_REQUIRE(p->pose_enc0_w && p->pose_enc2_w && p->pose_enc4_w && p->pose_fc0_w && p->pose_fc2_w,
                       "vcn_create: VCN_VC needs pose_encoder / pose_fc weights");
        srcs.push_back({&m->pose_enc0, p->pose_enc0_w, p->pose_enc0_b, 64, 3, 3, 0});
        srcs.push_back({&m->pose_enc2, p->pose_enc2_w, p->pose_enc2_b, 128, 64, 64, 0});
        srcs.push_back({&m->pose_enc4, p->pose_enc4_w, p->pose_enc4_b, 1024, 128, 128, 0});
        srcs.push_back({&m->pose_fc0, p->pose_fc0_w, p->pose_fc0_b, 512, 1024, 1024, 0});
        srcs.push_back({&m->pose_fc2, p->pose_fc2_w, p->pose_fc2_b, 9, 512, 512, 0});
    }
    SEEVCN_REQUIRE(p->enc1_0_w && p->enc1_3_w && p->enc2_0_w && p->enc2_3_w && p->fc0_w && p->fc2_w && p->fc4_w,
                   "vcn_create: missing encoder / shape_fc weights");
    srcs.push_back({&m->enc1_0, p->enc1_0_w, p->enc1_0_b, 128, 3, 3, 0});
    srcs.push_back({&m->enc1_3, p->enc1_3_w, p->enc1_3_b, 256, 128, 128, 0});
    srcs.push_back({&m->enc2_0_global, p->enc2_0_w, p->enc2_0_b, 512, 256, 512, 0});   // columns [0,256): global half + bias
    srcs.push_back({&m->enc2_0_local, p->enc2_0_w, nullptr, 512, 256, 512, 256});      // columns [256,512): per-point half
    srcs.push_back({&m->enc2_3, p->enc2_3_w, p->enc2_3_b, 1024, 512, 512, 0});
    srcs.push_back({&m->fc0, p->fc0_w, p->fc0_b, 1024, 1024, 1024, 0});
    srcs.push_back({&m->fc2, p->fc2_w, p->fc2_b, 1024, 1024, 1024, 0});
    srcs.push_back({&m->fc4, p->fc4_w, p->fc4_b, 3 * p->num_coarse, 1024, 1024, 0});
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
    std::vector<size_t> off_w, off_b, off_w16;
    for (auto& s : srcs) {
        const int kpad = (int)align_up((size_t)s.cin, 64);
        off_w.push_back(take((size_t)s.cout * s.cin * 4));
        off_b.push_back(take((size_t)s.cout * 4));
        off_w16.push_back(take((size_t)s.cout * kpad * 2));
    }
    cudaStream_t st = as_stream(stream);
    char* blob = nullptr;
    SEEVCN_CUDA_CHECK(cudaMalloc(&blob, o));
    m->blob = blob; m->blob_bytes = o;
    SEEVCN_CUDA_CHECK(cudaMemsetAsync(blob, 0, o, st));
    for (size_t i = 0; i < srcs.size(); ++i) {
        auto& s = srcs[i];
        float* w = reinterpret_cast<float*>(blob + off_w[i]);
        float* b = reinterpret_cast<float*>(blob + off_b[i]);
        auto* w16 = reinterpret_cast<__nv_bfloat16*>(blob + off_w16[i]);
        const int kpad = (int)align_up((size_t)s.cin, 64);
        SEEVCN_CUDA_CHECK(cudaMemcpy2DAsync(w, (size_t)s.cin * 4, s.w + s.col0, (size_t)s.ldw * 4, (size_t)s.cin * 4,
                                            s.cout, cudaMemcpyDeviceToDevice, st));
        if (s.b) SEEVCN_CUDA_CHECK(cudaMemcpyAsync(b, s.b, (size_t)s.cout * 4, cudaMemcpyDeviceToDevice, st));
        const size_t tot = (size_t)s.cout * kpad;
        f32_to_bf16_kernel<<<(unsigned)div_up(tot, (size_t)256), 256, 0, st>>>(s.cout, s.cin, w, s.cin, w16, kpad);
        SEEVCN_LAUNCH_CHECK();
        s.dst->w = w; s.dst->b = b; s.dst->w16 = w16;
        s.dst->cin = s.cin; s.dst->cout = s.cout; s.dst->ldw = s.cin; s.dst->kpad = kpad;
    }
    SEEVCN_CUDA_CHECK(cudaStreamSynchronize(st));
    *out_model = m;
    return SEEVCN_OK;
}

extern "C" void seevcn_vcn_destroy(seevcn_vcn_model* model) {
    if (!model) return;
    if (model->blob) cudaFree(model->blob);
    delete model;
}

extern "C" size_t seevcn_vcn_workspace_bytes(const seevcn_vcn_model* model, int num_obj, int n_pts) {
    if (!model) return 0;
    const size_t a = vcn_ws(num_obj, n_pts, model->num_coarse, 0).total;
    const size_t b = vcn_ws(num_obj, n_pts, model->num_coarse, 1).total;
    return a > b ? a : b;
}

#define TRY(expr) do { int _rc = (expr); if (_rc != SEEVCN_OK) return _rc; } while (0)

extern "C" int seevcn_vcn_forward(const seevcn_vcn_model* M, int num_obj, int n, const float* input,
                                  const float* gt_boxes, float* coarse, float* reg_rot, float* reg_centre,
                                  void* workspace, size_t workspace_bytes, int precision, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(M, "vcn_forward: null model");
    SEEVCN_REQUIRE(num_obj >= 0 && n >= 1, "vcn_forward: bad sizes");
    SEEVCN_REQUIRE(precision >= 0 && precision <= 2, "vcn_forward: precision must be 0 (bf16 fused tcgen05 chains), 1 (fp32) or 2 (bf16, one tcgen05 GEMM per layer)");
    if (num_obj == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(input && coarse && workspace, "vcn_forward: null pointer");
    SEEVCN_REQUIRE(M->viewer_centred || gt_boxes, "vcn_forward: VCN_CN needs gt_boxes");
    SEEVCN_REQUIRE(num_obj <= 65535, "vcn_forward: num_obj > 65535 per call");
    const VcnWs w = vcn_ws(num_obj, n, M->num_coarse, precision == 1 ? 1 : 0);
    if (workspace_bytes < w.total) {
        seevcn_set_error("vcn_forward: workspace %zu < %zu", workspace_bytes, w.total);
        return SEEVCN_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    SEEVCN_PROF("vcn_forward", st);
    char* ws = static_cast<char*>(workspace);
    auto* frames = reinterpret_cast<VcnFrame*>(ws + w.frames);
    auto* poses = reinterpret_cast<VcnPose*>(ws + w.poses);
    float* pts3 = reinterpret_cast<float*>(ws + w.pts3);
    float* pose_feat = reinterpret_cast<float*>(ws + w.pose_feat);
    float* h512 = reinterpret_cast<float*>(ws + w.h512);
    float* rel = reinterpret_cast<float*>(ws + w.rel);
    float* g256 = reinterpret_cast<float*>(ws + w.g256);
    float* objbias = reinterpret_cast<float*>(ws + w.objbias);
    float* feat = reinterpret_cast<float*>(ws + w.feat);
    float* fc_a = reinterpret_cast<float*>(ws + w.fc_a);
    float* fc_b = reinterpret_cast<float*>(ws + w.fc_b);
    float* coarse_cn = reinterpret_cast<float*>(ws + w.coarse_cn);
    const bool tc = precision != 1;
    float* actA = reinterpret_cast<float*>(ws + w.act_a);
    float* actB = reinterpret_cast<float*>(ws + w.act_b);
    auto* actA16 = reinterpret_cast<__nv_bfloat16*>(ws + w.act_a);
    auto* actB16 = reinterpret_cast<__nv_bfloat16*>(ws + w.act_b);
    const float NEG_INF = -__builtin_huge_valf();

    auto* fcxA = reinterpret_cast<__nv_bfloat16*>(ws + w.fcx_a);
    auto* fcxB = reinterpret_cast<__nv_bfloat16*>(ws + w.fcx_b);
    // per-point layer: X (rows, cin) -> act(X W^T + b [+ obj_bias]); fp32 SIMT or bf16 tcgen05 by path.
    // K = 3 input layers read the fp32 points directly on both paths.
    auto pp_layer = [&](const LinearW& L, int rows, const void* X, int ldx, const float* ob, int act, void* Y, int ldy,
                        float* colmax) -> int {
        if (!tc) return linear_f32(L, rows, static_cast<const float*>(X), ldx, ob, n, act, static_cast<float*>(Y), ldy, colmax, st);
        if (L.cin == 3) return vcn_pointwise3(L, (size_t)rows, static_cast<const float*>(X), act, static_cast<__nv_bfloat16*>(Y), ldy, st);
        return vcn_linear_tc(L, rows, static_cast<const __nv_bfloat16*>(X), ldx, ob, n, act,
                             static_cast<__nv_bfloat16*>(Y), ldy, nullptr, colmax, st);
    };
    // per-object FC layer: X fp32 (B, cin) -> fp32 Y (B, cout); the tcgen05 path converts X to bf16 first
    auto fc_rows = [&](const LinearW& L, int rows, const float* X, int act, float* Y, __nv_bfloat16* xb) -> int {
        if (!tc || L.cout % 128 != 0) return linear_f32(L, rows, X, L.cin, nullptr, 1, act, Y, L.cout, nullptr, st);
        const size_t tot = (size_t)rows * L.kpad;
        f32_to_bf16_kernel<<<(unsigned)div_up(tot, (size_t)256), 256, 0, st>>>(rows, L.cin, X, L.cin, xb, L.kpad);
        SEEVCN_LAUNCH_CHECK();
        return vcn_linear_tc(L, rows, xb, L.kpad, nullptr, 1, act, nullptr, 0, Y, nullptr, st);
    };
    auto fc_layer = [&](const LinearW& L, const float* X, int act, float* Y, __nv_bfloat16* xb) -> int {
        return fc_rows(L, num_obj, X, act, Y, xb);
    };
    float* fc_part = reinterpret_cast<float*>(ws + w.fc_part);
    auto to_bf16 = [&](const float* X, int rows, int cols, __nv_bfloat16* xb) -> int {
        const size_t tot = (size_t)rows * cols;
        SEEVCN_CUDA_CHECK(launch_pdl(f32_to_bf16_kernel, dim3((unsigned)div_up(tot, (size_t)256)), dim3(256), 0, st,
                                     (size_t)rows, cols, X, cols, xb, cols));
        SEEVCN_LAUNCH_CHECK();
        return SEEVCN_OK;
    };
    // num_coarse with 3*num_coarse not a multiple of 128: last FC on CUDA cores from the fp32 copy of its input
    auto linear_f32_from_bf16 = [&](const LinearW& L, int rows, const __nv_bfloat16* xb, float* Y) -> int {
        bf16_to_f32_kernel<<<(unsigned)div_up((size_t)rows * L.cin, (size_t)256), 256, 0, st>>>((size_t)rows * L.cin, xb, fc_b);
        SEEVCN_LAUNCH_CHECK();
        return linear_f32(L, rows, fc_b, L.cin, nullptr, 1, ACT_NONE, Y, L.cout, nullptr, st);
    };
    auto pts_in = [&](int) -> int { return SEEVCN_OK; };
    const void* ptsX = pts3;
    const int ptsLd = 3;
    void* A = tc ? static_cast<void*>(actA16) : static_cast<void*>(actA);
    void* Bf = tc ? static_cast<void*>(actB16) : static_cast<void*>(actB);

    if (precision == 0) {
        // ---- fused tcgen05 chains (vcn_chain.cu): one launch per chain, activations stay in TMEM ----
        VcnMaxInit init;
        init.dst[0] = g256; init.width[0] = 256;
        init.dst[1] = feat; init.width[1] = 1024;
        int ks = 0;
        if (M->viewer_centred) {
            init.dst[2] = pose_feat; init.width[2] = 1024;
            SEEVCN_CUDA_CHECK(launch_pdl(vcn_frame_kernel, dim3(num_obj), dim3(256), 0, st, n, 1, input, nullptr, frames, nullptr, init));
            SEEVCN_LAUNCH_CHECK();
            TRY(vcn_chain_pose(M, num_obj, n, input, frames, pose_feat, st));
            TRY(to_bf16(pose_feat, num_obj, 1024, fcxA));
            SEEVCN_REQUIRE(M->pose_fc0.cout == kPoseHidden && M->pose_fc2.cout == 9, "vcn_forward: unexpected pose_fc shape");
            TRY(vcn_fc_tc_partials(M->pose_fc0, num_obj, fcxA, 1024, fc_part, &ks, st));
            SEEVCN_CUDA_CHECK(launch_pdl(vcn_pose_tail_kernel, dim3(num_obj), dim3(288), 0, st, num_obj, ks, fc_part, M->pose_fc0.b,
                                         M->pose_fc2.w, M->pose_fc2.b, frames, poses, reg_rot, reg_centre));
            SEEVCN_LAUNCH_CHECK();
        } else {
            SEEVCN_CUDA_CHECK(launch_pdl(vcn_frame_kernel, dim3(num_obj), dim3(256), 0, st, n, 0, input, gt_boxes, frames, nullptr, init));
            SEEVCN_LAUNCH_CHECK();
        }
        for (int o0 = 0; o0 < num_obj; o0 += w.chunk) {
            const int nb = std::min(w.chunk, num_obj - o0);
            TRY(vcn_chain_enc1(M, nb, n, input + (size_t)o0 * n * 3, frames + o0, poses + o0, actA16, g256 + (size_t)o0 * 256, st));
            TRY(to_bf16(g256 + (size_t)o0 * 256, nb, 256, fcxA));
            // per-object bias of mlp_conv2.0 = W[:, :256] . global + b: the chain sums the split-K slices itself
            TRY(vcn_fc_tc_partials(M->enc2_0_global, nb, fcxA, 256, fc_part, &ks, st));
            TRY(vcn_chain_enc2(M, nb, n, actA16, fc_part, ks, feat + (size_t)o0 * 1024, st));
        }
        TRY(to_bf16(feat, num_obj, 1024, fcxA));
        TRY(vcn_fc_tc(M->fc0, num_obj, fcxA, 1024, ACT_RELU, nullptr, fcxB, 1024, fc_part, st));
        TRY(vcn_fc_tc(M->fc2, num_obj, fcxB, 1024, ACT_RELU, nullptr, fcxA, 1024, fc_part, st));
        if ((3 * M->num_coarse) % 128 == 0) {
            TRY(vcn_fc_tc_partials(M->fc4, num_obj, fcxA, 1024, fc_part, &ks, st));
            SEEVCN_CUDA_CHECK(launch_pdl(vcn_output_kernel, dim3(div_up(M->num_coarse, 256), num_obj), dim3(256), 0, st,
                                         M->num_coarse, M->viewer_centred, fc_part, frames, poses, coarse, ks, M->fc4.b));
            SEEVCN_LAUNCH_CHECK();
            return SEEVCN_OK;
        }
        TRY(linear_f32_from_bf16(M->fc4, num_obj, fcxA, coarse_cn));
        vcn_output_kernel<<<dim3(div_up(M->num_coarse, 256), num_obj), 256, 0, st>>>(M->num_coarse, M->viewer_centred,
                                                                                     coarse_cn, frames, poses, coarse);
        SEEVCN_LAUNCH_CHECK();
        return SEEVCN_OK;
    }
    if (M->viewer_centred) {
        TRY(fill(pose_feat, (size_t)num_obj * 1024, NEG_INF, st));
        for (int o0 = 0; o0 < num_obj; o0 += w.chunk) {
            const int nb = std::min(w.chunk, num_obj - o0), rows = nb * n;
            vcn_frame_kernel<<<nb, 256, 0, st>>>(n, 1, input + (size_t)o0 * n * 3, nullptr, frames + o0, pts3);
            SEEVCN_LAUNCH_CHECK();
            TRY(pts_in(rows));
            TRY(pp_layer(M->pose_enc0, rows, ptsX, ptsLd, nullptr, ACT_LEAKY, A, 64, nullptr));
            TRY(pp_layer(M->pose_enc2, rows, A, 64, nullptr, ACT_LEAKY, Bf, 128, nullptr));
            TRY(pp_layer(M->pose_enc4, rows, Bf, 128, nullptr, ACT_NONE, nullptr, 0, pose_feat + (size_t)o0 * 1024));
        }
        TRY(fc_layer(M->pose_fc0, pose_feat, ACT_LEAKY, h512, fcxA));
        TRY(linear_f32(M->pose_fc2, num_obj, h512, 512, nullptr, 1, ACT_NONE, rel, 16, nullptr, st));
        vcn_pose_kernel<<<div_up(num_obj, 128), 128, 0, st>>>(num_obj, rel, 16, frames, poses, reg_rot, reg_centre);
        SEEVCN_LAUNCH_CHECK();
    }
    TRY(fill(g256, (size_t)num_obj * 256, NEG_INF, st));
    TRY(fill(feat, (size_t)num_obj * 1024, NEG_INF, st));
    for (int o0 = 0; o0 < num_obj; o0 += w.chunk) {
        const int nb = std::min(w.chunk, num_obj - o0), rows = nb * n;
        if (M->viewer_centred) {
            vcn_canon_kernel<<<dim3(div_up(n, 256), nb), 256, 0, st>>>(n, input + (size_t)o0 * n * 3, frames + o0,
                                                                       poses + o0, pts3);
        } else {
            vcn_frame_kernel<<<nb, 256, 0, st>>>(n, 0, input + (size_t)o0 * n * 3, gt_boxes + (size_t)o0 * 7, frames + o0,
                                                 pts3);
        }
        SEEVCN_LAUNCH_CHECK();
        TRY(pts_in(rows));
        TRY(pp_layer(M->enc1_0, rows, ptsX, ptsLd, nullptr, ACT_RELU, A, 128, nullptr));
        TRY(pp_layer(M->enc1_3, rows, A, 128, nullptr, ACT_NONE, Bf, 256, g256 + (size_t)o0 * 256));
        // per-object bias = W[:, :256] . global + b   (the torch.cat + expand of VCN_VC.py:100-101, folded)
        TRY(fc_rows(M->enc2_0_global, nb, g256 + (size_t)o0 * 256, ACT_NONE, objbias + (size_t)o0 * 512, fcxA));
        TRY(pp_layer(M->enc2_0_local, rows, Bf, 256, objbias + (size_t)o0 * 512, ACT_RELU, A, 512, nullptr));
        TRY(pp_layer(M->enc2_3, rows, A, 512, nullptr, ACT_NONE, nullptr, 0, feat + (size_t)o0 * 1024));
    }
    TRY(fc_layer(M->fc0, feat, ACT_RELU, fc_a, fcxA));
    TRY(fc_layer(M->fc2, fc_a, ACT_RELU, fc_b, fcxB));
    TRY(fc_layer(M->fc4, fc_b, ACT_NONE, coarse_cn, fcxA));
    vcn_output_kernel<<<dim3(div_up(M->num_coarse, 256), num_obj), 256, 0, st>>>(M->num_coarse, M->viewer_centred,
                                                                                 coarse_cn, frames, poses, coarse);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
