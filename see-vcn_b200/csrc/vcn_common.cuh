// VCN model container + the small per-object kernels shared by the fp32 and tcgen05 paths.
#pragma once
#include "common.cuh"

// Per-object viewer frame (VCN_VC.py:185-190) or GT frame (VCN_CN.py:146-147, transform.py:91-161)
struct __align__(16) VcnFrame {
    float ca, sa;        // cos/sin of the canonicalising z-rotation angle a (= -theta or -heading)
    float theta;         // angle to rotate back by (+theta / +heading)
    float scale;         // CN: box length gt[:,3]; VC: 1
    float mean[3];       // VC: mean of the rotated cloud; CN: box centre (subtracted BEFORE rotating)
    float pad;
};

// Regressed pose (VCN_VC.py:194-198): centre (3) + rot (3x3 row-major, columns x,y,z)
struct __align__(16) VcnPose {
    float centre[3];
    float pad0;
    float rot[9];
    float pad1[3];
};

struct LinearW {
    const float* w = nullptr;           // (cout, ldw) fp32
    const float* b = nullptr;           // (cout)
    const __nv_bfloat16* w16 = nullptr; // packed bf16 copy, (cout, kpad) row-major, K zero-padded to 64
    int cin = 0, cout = 0, ldw = 0, kpad = 0;
};

struct seevcn_vcn_model {
    int num_coarse = 1024;
    int viewer_centred = 1;
    LinearW pose_enc0, pose_enc2, pose_enc4, pose_fc0, pose_fc2;
    LinearW enc1_0, enc1_3, enc2_0_global, enc2_0_local, enc2_3, fc0, fc2, fc4;
    void* blob = nullptr;   // one device allocation holding every copy above
    size_t blob_bytes = 0;
};

// Programmatic dependent launch (the VCN forward is ~20 short dependent kernels on one stream).  Every kernel of the
// forward starts with pdl_launch_dependents() — the next kernel's CTAs may be scheduled and run their prologue (barrier
// init, TMEM allocation, tensor-map prefetch, staging of constant weights) while this one is still running — and calls
// pdl_wait() before it touches anything a predecessor wrote (or still reads): that returns when the preceding grid has
// completed and its writes are visible.  Since every kernel waits at its top, completion of the predecessor implies
// completion of everything before it.  Both are no-ops in a kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Activation codes used by every linear kernel
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };   // LeakyReLU slope 0.01 (torch default)

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
    return v;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    v += 0.f;   // -0.0 -> +0.0: as an int, -0.0 would order below every negative float
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
