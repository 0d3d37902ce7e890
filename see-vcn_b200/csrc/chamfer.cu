// Parity metric — Chamfer distance (forward only).
// Replaces chamfer_dist_kernel, see/surface_completion/models/vcn/extensions/chamfer_dist/chamfer.cu:15-145:
// for every point of xyz1 the squared distance to its nearest neighbour in xyz2 (and vice versa).
// One thread per query, the other cloud tiled through shared memory.
#include "common.cuh"

namespace {
constexpr int kT = 256, kTile = 1024;

__global__ void __launch_bounds__(kT)
nn_dist_kernel(int n, int m, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dist) {
    __shared__ float sx[kTile], sy[kTile], sz[kTile];
    const int bi = blockIdx.y;
    const int i = blockIdx.x * kT + threadIdx.x;
    const float* bp = b + (size_t)bi * m * 3;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < n) { const float* p = a + ((size_t)bi * n + i) * 3; x = p[0]; y = p[1]; z = p[2]; }
    float best = __int_as_float(0x7f800000);
    for (int j0 = 0; j0 < m; j0 += kTile) {
        const int cnt = min(kTile, m - j0);
        __syncthreads();
        for (int f = threadIdx.x; f < cnt * 3; f += kT) {
            const float v = bp[(size_t)j0 * 3 + f];
            const int p = f / 3, c = f - 3 * p;
            (c == 0 ? sx : c == 1 ? sy : sz)[p] = v;
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float dx = x - sx[j], dy = y - sy[j], dz = z - sz[j];
            best = fminf(best, dx * dx + dy * dy + dz * dz);
        }
    }
    if (i < n) dist[(size_t)bi * n + i] = best;
}
}  // namespace

extern "C" int seevcn_chamfer(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, float* dist2,
                              seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && n >= 0 && m >= 0, "chamfer: negative size");
    if (b == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(b <= 65535, "chamfer: b > 65535");
    SEEVCN_REQUIRE(xyz1 && xyz2, "chamfer: null pointer");
    cudaStream_t st = as_stream(stream);
    if (dist1 && n > 0) { nn_dist_kernel<<<dim3(div_up(n, kT), b), kT, 0, st>>>(n, m, xyz1, xyz2, dist1); SEEVCN_LAUNCH_CHECK(); }
    if (dist2 && m > 0) { nn_dist_kernel<<<dim3(div_up(m, kT), b), kT, 0, st>>>(m, n, xyz2, xyz1, dist2); SEEVCN_LAUNCH_CHECK(); }
    return SEEVCN_OK;
}
