// Shared helpers for the sm_100a kernels behind include/ape_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/ape_b200.h"

namespace ape {

// ---- error reporting (thread-local message behind ape_last_error) -------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  check_launch(const char* what);     // cudaGetLastError -> APE_OK / APE_ERR_CUDA
int  sm_count();                         // of the CURRENT device
// cudaFuncSetAttribute is per device: `static ape::PerDevice done; if (done.first()) { ...set attributes... }` runs the block
// once per device of the process (the current device at the time of the call).
struct PerDevice {
    bool seen[64] = {};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (seen[dev]) return false;
        seen[dev] = true;
        return true;
    }
};

// Optional per-launch device timing (bench.py roofline leg): CUDA events on the launching stream around
// every kernel, aggregated by label in ape_profile_report().  Disabled by default (one branch per launch).
bool prof_enabled();
void prof_push(const char* label, cudaEvent_t e0, cudaEvent_t e1);
// Every launch site is also an NVTX range named after the kernel label (header-only NVTX 3: a no-op unless a tool such as
// Nsight Systems / Compute is attached), so a timeline shows "gemm.pn.heads1", "icp_p2p", ... instead of mangled names.
struct ProfScope {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t s;
    const char* label;
    ProfScope(const char* l, cudaStream_t st) : s(st), label(l) {
        nvtxRangePushA(l);
        if (prof_enabled()) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); }
    }
    ~ProfScope() {
        if (e0) { cudaEventRecord(e1, s); prof_push(label, e0, e1); }
        nvtxRangePop();
    }
};

#define APE_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::ape::set_error(__VA_ARGS__);                       \
            return APE_ERR_INVALID;                              \
        }                                                        \
    } while (0)

#define APE_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ::ape::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
            return APE_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------
// The kernels of the per-frame pipeline are launched with the stream-serialization attribute: a kernel's CTAs may be
// scheduled while the previous kernel on the stream is still draining its last CTAs, and every such kernel calls
// pdl_sync() before it touches global memory, which waits for the COMPLETION (and memory flush) of that previous kernel:
// the semantics of a normal launch, minus the launch gap.  pdl_sync() also lets the next kernel be scheduled early.
// APE_PDL=0 turns the attribute off (normal launches; pdl_sync() is then a no-op).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 16-byte load that does not allocate in L1 (inputs that are read exactly once)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Row-major 3x3 "base" of a normalised quaternion (w,x,y,z), fp32, written as in
// DenseFusion/tools/utils.py:46-67 / lib/loss_refiner.py:19-29.
__device__ __forceinline__ void quat_to_base(float w, float x, float y, float z, float* R) {
    R[0] = 1.0f - 2.0f * (y * y + z * z);
    R[1] = 2.0f * x * y - 2.0f * w * z;
    R[2] = 2.0f * w * y + 2.0f * x * z;
    R[3] = 2.0f * x * y + 2.0f * z * w;
    R[4] = 1.0f - 2.0f * (x * x + z * z);
    R[5] = -2.0f * w * x + 2.0f * y * z;
    R[6] = -2.0f * w * y + 2.0f * x * z;
    R[7] = 2.0f * w * x + 2.0f * y * z;
    R[8] = 1.0f - 2.0f * (x * x + y * y);
}

}  // namespace ape
