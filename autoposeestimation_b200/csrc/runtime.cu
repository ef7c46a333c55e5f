// Library-level plumbing: version, error message, launch counter, device properties.
#include "ape_common.cuh"
#include <atomic>
#include <cstring>

namespace ape {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return APE_ERR_CUDA;
    }
    return APE_OK;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        n = p.multiProcessorCount;
    }
    return n;
}

}  // namespace ape

extern "C" __attribute__((visibility("default"))) int ape_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) const char* ape_last_error(void) { return ape::g_err; }
extern "C" __attribute__((visibility("default"))) uint64_t ape_launch_count(void) { return ape::g_launches.load(); }
