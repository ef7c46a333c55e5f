// Library-level plumbing: version, error message, launch counter, device properties.
#include "ape_common.cuh"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

struct ProfEntry { const char* label; cudaEvent_t e0, e1; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;
bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void prof_push(const char* label, cudaEvent_t e0, cudaEvent_t e1) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({label, e0, e1});
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("APE_PDL"); on = e ? (atoi(e) != 0) : 1; }
    return on != 0;
}

int sm_count() {
    static int n[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    dev &= 63;
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        n[dev] = v;
    }
    return n[dev];
}

}  // namespace ape

extern "C" __attribute__((visibility("default"))) int ape_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) const char* ape_last_error(void) { return ape::g_err; }
extern "C" __attribute__((visibility("default"))) uint64_t ape_launch_count(void) { return ape::g_launches.load(); }

// Enable/disable per-launch timing; enabling clears previously collected entries.
extern "C" __attribute__((visibility("default"))) int ape_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(ape::g_prof_mu);
    if (on) {
        for (auto& e : ape::g_prof) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
        ape::g_prof.clear();
    }
    ape::g_prof_on.store(on != 0);
    return APE_OK;
}
// Writes "label launches total_ms\n" per kernel label (device time between the events that bracket each
// launch); synchronises on the recorded events.  Returns the number of bytes needed (including the NUL).
extern "C" __attribute__((visibility("default"))) int ape_profile_report(char* buf, int buflen) {
    std::lock_guard<std::mutex> lk(ape::g_prof_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& e : ape::g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(e.e1) == cudaSuccess && cudaEventElapsedTime(&ms, e.e0, e.e1) == cudaSuccess) {
            auto& a = agg[e.label]; a.first += 1; a.second += ms;
        }
    }
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && buflen > 0) { strncpy(buf, out.c_str(), (size_t)buflen - 1); buf[buflen - 1] = 0; }
    return (int)out.size() + 1;
}
