// C-ABI plumbing: version, per-thread error message, device check.
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void seevcn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void seevcn_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long seevcn_launch_count(void) { return g_launches.load(); }

extern "C" int seevcn_abi_version(void) { return 1; }
extern "C" const char* seevcn_last_error(void) { return g_err; }

extern "C" int seevcn_check_device(int dev) {
    cudaDeviceProp prop;
    SEEVCN_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        seevcn_set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        return SEEVCN_E_UNSUPPORTED;
    }
    return SEEVCN_OK;
}
