// Row (e) of the path — the end-of-path collection "for the detector" (reference: commu_utils.py:50-111 all_gather of
// pickled results, sc_multiproc.py:81-85).  One rank's batch travels as ONE contiguous record
//     header (header_words int32: objects, voxels, frame offset, feature width, 0...) | completed clouds (fp32) |
//     voxel coords (int32 x 4) | voxel features (fp32 x 3) | points per voxel (int32)
// packed by one kernel, then delivered into slot `rank` of every peer's receive buffer with one device-to-device copy
// per peer (symmetric-memory mappings: plain cudaMemcpyAsync over NVLink, no SM of the compute kernels involved).
// HBM-bound: 4 B read + 4 B written per record word.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
gather_pack_kernel(int header_words, int4 hdr, long long n_clu, long long m, const float* __restrict__ clustered,
                   const int* __restrict__ coords, const float* __restrict__ feats, const int* __restrict__ nums,
                   unsigned* __restrict__ send) {
    const long long total = header_words + n_clu + m * 8;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        unsigned v;
        if (i < header_words) {
            v = i == 0 ? (unsigned)hdr.x : i == 1 ? (unsigned)hdr.y : i == 2 ? (unsigned)hdr.z : i == 3 ? (unsigned)hdr.w : 0u;
        } else {
            long long j = i - header_words;
            if (j < n_clu) v = __float_as_uint(clustered[j]);
            else if ((j -= n_clu) < m * 4) v = (unsigned)coords[j];
            else if ((j -= m * 4) < m * 3) v = __float_as_uint(feats[j]);
            else v = (unsigned)nums[j - m * 3];
        }
        send[i] = v;
    }
}

}  // namespace

extern "C" int seevcn_gather_pack(int header_words, int num_objects, int num_voxels, int frame_offset, int feat_width,
                                  long long clustered_words, const float* clustered, const int* coords, const float* feats,
                                  const int* nums, void* send, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(header_words >= 4 && num_objects >= 0 && num_voxels >= 0 && clustered_words >= 0, "gather_pack: bad sizes");
    SEEVCN_REQUIRE(feat_width == 3, "gather_pack: voxel features must be 3 wide");
    SEEVCN_REQUIRE(send && (clustered_words == 0 || clustered) && (num_voxels == 0 || (coords && feats && nums)),
                   "gather_pack: null pointer");
    const long long total = header_words + clustered_words + (long long)num_voxels * 8;
    const long long want = (total + 255) / 256;
    const int grid = (int)(want < (long long)seevcn_num_sms() * 8 ? want : (long long)seevcn_num_sms() * 8);
    gather_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(header_words, make_int4(num_objects, num_voxels, frame_offset, feat_width),
                                                            clustered_words, (long long)num_voxels, clustered, coords, feats, nums,
                                                            static_cast<unsigned*>(send));
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

extern "C" int seevcn_gather_broadcast(const void* send, size_t bytes, void* const* dst, int num_dst, seevcn_stream_t stream) {
    SEEVCN_REQUIRE(num_dst >= 0 && (num_dst == 0 || dst) && (bytes == 0 || send), "gather_broadcast: null pointer");
    for (int r = 0; r < num_dst; ++r) {
        SEEVCN_REQUIRE(dst[r], "gather_broadcast: null destination %d", r);
        SEEVCN_CUDA_CHECK(cudaMemcpyAsync(dst[r], send, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    }
    return SEEVCN_OK;
}
