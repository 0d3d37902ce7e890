// Microbenchmark: cycles per tcgen05.mma for the shapes / operand sources the fused chains could use.
// One CTA per SM issues REPS back-to-back MMAs (operands: whatever is in smem / TMEM), commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../see-vcn_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

constexpr int REPS = 512;

__device__ __forceinline__ uint32_t idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// variant: bit0 = A from TMEM, bit1 = alternate between two accumulators, n = UMMA N, kadv = advance operands like a real k loop
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int variant, int n, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t scratch[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&scratch[0], 1); mbar_init(&scratch[1], 1); mbar_init(&bar, (variant & 8) ? 2 : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const bool ts = variant & 1, alt = variant & 2;
    long long t0 = 0, t1 = 0;
    const bool dual = variant & 8;
    if ((warp == 1 || (dual && warp == 2)) && lane == 0) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
        const uint32_t id = idesc(128, n);
        t0 = clock64();
        if (variant & 4) {
            // lean issue loop: descriptors precomputed, 8 MMAs per iteration, no per-MMA address arithmetic
            uint64_t dbs[8], das[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) dbs[i] = make_desc(b_addr + (i >> 2) * 16384 + (i & 3) * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) das[i] = make_desc(a_addr + i * 32);
            const uint32_t d0 = tmem + (dual && warp == 2 ? 128 : 0), d1 = d0 + (alt ? 256 : 0);
            for (int r = 0; r < (dual ? REPS / 2 : REPS); r += 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t d = i < 4 ? d0 : d1;
                    if (ts) tc_mma_ts(d, tmem + 384 + (i & 3) * 8, dbs[i], id, 1u);
                    else tc_mma(d, das[i & 3], dbs[i], id, 1u);
                    if ((variant & 16) && (i & 3) == 3) tc_commit(&scratch[warp & 1]);
                    if ((variant & 32) && i == 7) tc_commit(&scratch[warp & 1]);
                }
            }
        } else
        for (int r = 0; r < REPS; ++r) {
            const int k = r & 3;                       // 4 k-steps inside a 128-byte swizzle row, like the real loop
            const uint32_t d = tmem + ((alt && (r & 4)) ? 256 : 0);   // switch accumulator every 4 MMAs
            const uint64_t db = make_desc(b_addr + ((r >> 2) & 1) * 32768 / 2 + k * 32);
            if (ts) tc_mma_ts(d, tmem + 384 + k * 8, db, id, r > 7);
            else tc_mma(d, make_desc(a_addr + k * 32), db, id, r > 7);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (warp == 1) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory"); }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const int smem = 97 * 1024 + 1024;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[64] = {}; names[13] = "TS lean 2w"; names[29] = "TS lean 2w commit/4"; names[45] = "TS lean 2w commit/8"; names[28] = "SS lean 2w commit/4"; const char* unused[16] = {"SS same-acc", "TS same-acc", "SS alt-acc ", "TS alt-acc ", "SS same lean", "TS same lean", "SS alt lean", "TS alt lean","","","","","SS lean 2warps","TS lean 2warps","",""};
    for (int grid : {148})
        for (int n : {64, 128, 256})
            for (int v : {13, 13 + 16, 13 + 32, 12 + 16}) {
                if (n == 256 && (v & 2)) { /* accumulators at col 0 and 256 still fit (A at 384 overlaps: timing only) */ }
                mma_rate_kernel<<<grid, 128, smem>>>(v, n, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                const double per = (double)mx / REPS;
                printf("grid %3d  N=%3d  %s : %7.1f clk/MMA  -> %6.0f MAC/clk/SM\n", grid, n, names[v], per, 128.0 * n * 16 / per);
            }
    return 0;
}
