// placeholder until the tcgen05 kernel lands (next commit)
#include "vcn_common.cuh"
int vcn_linear_tc(const LinearW&, int, const __nv_bfloat16*, int, const float*, int, int, __nv_bfloat16*, int, float*,
                  float*, cudaStream_t) {
    seevcn_set_error("vcn_linear_tc: tcgen05 path not built yet");
    return SEEVCN_E_UNSUPPORTED;
}
