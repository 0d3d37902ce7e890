// C-ABI plumbing: version, per-thread error message, device check.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
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

extern "C" int seevcn_abi_version(void) { return 2; }
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

// SM count of the current device, cached per device (no process-wide "the device" assumption).
int seevcn_num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ---- optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg) ----
// Off by default: a disabled scope costs one relaxed load.  Enabled, every scope records two events on the
// stream its kernels are launched on; seevcn_prof_report() synchronises them and sums per name.
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

SeevcnProfScope::SeevcnProfScope(const char* name, cudaStream_t st) : st_(st), slot_(-1) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{name, prof_event(), prof_event()};
    cudaEventRecord(r.e0, st_);
    slot_ = (long)g_prof.size();
    g_prof.push_back(r);
}
SeevcnProfScope::~SeevcnProfScope() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if ((size_t)slot_ < g_prof.size()) cudaEventRecord(g_prof[slot_].e1, st_);
}

extern "C" int seevcn_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    const int prev = g_prof_on.exchange(on ? 1 : 0);
    if (on) {
        for (auto& r : g_prof) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
        g_prof.clear();
    }
    return prev;
}

extern "C" int seevcn_prof_report(char* buf, size_t cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long, double>> acc;
    std::vector<std::string> order;
    for (auto& r : g_prof) {
        if (cudaEventSynchronize(r.e1) != cudaSuccess) { seevcn_set_error("prof_report: event sync failed"); return SEEVCN_E_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (!acc.count(r.name)) order.push_back(r.name);
        auto& a = acc[r.name];
        a.first += 1; a.second += ms;
        g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1);
    }
    g_prof.clear();
    size_t off = 0;
    if (cap) buf[0] = 0;
    for (auto& n : order) {
        const int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s %ld %.6f\n", n.c_str(), acc[n].first, acc[n].second);
        if (w < 0 || off + (size_t)w >= cap) { seevcn_set_error("prof_report: buffer too small"); return SEEVCN_E_INVALID; }
        off += (size_t)w;
    }
    return SEEVCN_OK;
}

// ---- small results to the host without the copy engines ----
// The pipeline needs two tiny device results on the host while the GPU keeps running (box counts: how many objects to
// launch for; number of voxels).  A cudaMemcpyAsync for them queues behind the multi-megabyte result downloads on the
// same DMA engine; SM stores into pinned (UVA-mapped) host memory on the compute stream do not.
namespace {
__global__ void copy_words_to_host_kernel(  /* either side may be pinned host memory */
   const unsigned* __restrict__ src, unsigned* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
    __threadfence_system();
}
}  // namespace

extern "C" int seevcn_copy_to_pinned(const void* src_device, void* dst_pinned_host, size_t bytes, seevcn_stream_t stream) {
    if (bytes == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(src_device && dst_pinned_host, "copy_to_pinned: null pointer");
    SEEVCN_REQUIRE(bytes % 4 == 0 && ((uintptr_t)src_device & 3) == 0 && ((uintptr_t)dst_pinned_host & 3) == 0,
                   "copy_to_pinned: size and pointers must be multiples of 4 bytes");
    const size_t n = bytes / 4;
    copy_words_to_host_kernel<<<(unsigned)div_up(n, (size_t)256), 256, 0, as_stream(stream)>>>(
        static_cast<const unsigned*>(src_device), static_cast<unsigned*>(dst_pinned_host), n);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}

// The other direction: a few words (object ids) from pinned host memory into device memory with SM loads, so the
// upload does not queue behind the next batch's bulk H2D copy on the copy engine.
extern "C" int seevcn_copy_from_pinned(const void* src_pinned_host, void* dst_device, size_t bytes, seevcn_stream_t stream) {
    if (bytes == 0) return SEEVCN_OK;
    SEEVCN_REQUIRE(src_pinned_host && dst_device, "copy_from_pinned: null pointer");
    SEEVCN_REQUIRE(bytes % 4 == 0 && ((uintptr_t)src_pinned_host & 3) == 0 && ((uintptr_t)dst_device & 3) == 0,
                   "copy_from_pinned: size and pointers must be multiples of 4 bytes");
    const size_t n = bytes / 4;
    copy_words_to_host_kernel<<<(unsigned)div_up(n, (size_t)256), 256, 0, as_stream(stream)>>>(
        static_cast<const unsigned*>(src_pinned_host), static_cast<unsigned*>(dst_device), n);
    SEEVCN_LAUNCH_CHECK();
    return SEEVCN_OK;
}
