// Stage 4a — gather_operation / grouping_operation (forward).
//
// Replaces gather_points_kernel_fast (pointnet2_batch/src/sampling_gpu.cu:15-51) and
// group_points_kernel_fast (pointnet2_batch/src/group_points_gpu.cu:53-92).
//
// The reference launches one thread per OUTPUT ELEMENT with the channel on blockIdx.y, so
// every index is re-read C times.  Here one thread owns one (batch, output position), reads
// its index once and loops over channels: writes stay coalesced along the position axis,
// index traffic drops C-fold.  HBM-bound: 4 B idx + 4*C read + 4*C write per position.
#include "common.cuh"

namespace {

// out[b,c,j] = points[b,c,idx[b,j]] for j in [0, m); grid (ceil(m/256), B)
__global__ void __launch_bounds__(256)
gather_kernel(int c, int n, int m, const float* __restrict__ points, const int* __restrict__ idx,
              float* __restrict__ out) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int src = idx[(size_t)b * m + j];
    const float* p = points + (size_t)b * c * n + src;
    float* o = out + (size_t)b * c * m + j;
#pragma unroll 4
    for (int ch = 0; ch < c; ++ch) o[(size_t)ch * m] = __ldg(p + (size_t)ch * n);
}

}  // namespace

extern "C" int seevcn_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                                    float* out, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, "gather_points: negative size");
    if (b == 0 || c == 0 || npoints == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && idx && out, "gather_points: null pointer");
    SEEVCN_REQUIRE(b <= 65535, "gather_points: b > 65535");
    dim3 grid(div_up(npoints, 256), b);
    SEEVCN_PROF("gather_points", as_stream(stream));
    gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, npoints, points, idx, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

// out (B,C,P,S) is gather with m = P*S over the flattened (P,S) index tensor.
extern "C" int seevcn_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                   const int* idx, float* out, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "group_points: negative size");
    const long long m = (long long)npoints * nsample;
    SEEVCN_REQUIRE(m < (1ll << 31), "group_points: npoints*nsample overflows int32");
    if (b == 0 || c == 0 || m == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(points && idx && out, "group_points: null pointer");
    SEEVCN_REQUIRE(b <= 65535, "group_points: b > 65535");
    dim3 grid(div_up((int)m, 256), b);
    SEEVCN_PROF("group_points", as_stream(stream));
    gather_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, (int)m, points, idx, out);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
