// Colour-embedding gather at the sampled pixels (DenseFusion/lib/network.py:100-102: `emb = torch.gather(emb, 2, choose)`)
// as a stand-alone kernel, so that an encoder map that lives in HOST memory never has to cross the bus as a whole.
//
// `out_img` may be a device pointer or a pointer into mapped pinned host memory (cudaHostAlloc / torch pin_memory:
// under unified addressing the host pointer is directly usable by a kernel).  For a [B,32,hw] fp32 map of a 120 x 160
// crop and 500 sampled points per object only 500 of 19 200 columns are needed: 4.1 MB of 157 MB per batch of 64.
// Measured on this pool's B200 (PCIe 5 x16, tools/zc_bench.cu): full-map cudaMemcpyAsync 2.83 ms; zero-copy gather
// with lane = channel 2.58 ms (every lane its own 32-byte PCIe read), with lane = POINT 1.26 ms (`choose` is
// ascending, so neighbouring lanes fall into the same 128-byte line now and then and the requests merge); channels-last
// ([B,hw,32]: one point = one contiguous 128-byte read) 0.089 ms.  Run beside the step's kernels the NCHW zero-copy gather
// is not free: about half of its stand-alone duration shows up in the step time (its CTAs wait microseconds for every PCIe
// read and share the SMs with the persistent GEMM CTAs; the shared-memory carve-out preference makes no difference), which
// is why the Runner prefers the host pool while the host keeps up (densefusion/estimate_poses.py).
#include "ape_common.cuh"
#include <cstdlib>

namespace ape {

// NCHW map: thread = one sampled point, CH channels per thread (all loads issued before the first store).
// grid = (ceil(N / 128), B, 32 / CH), block = 128.
template <int CH>
__global__ void __launch_bounds__(128)
gather_emb_nchw_kernel(const float* __restrict__ img, int hw, const int64_t* __restrict__ choose, int N, float* __restrict__ emb, int B)
{
  // grid.y = min(B, 16): 256 CTAs walk the objects.  Reading a pinned host map, one CTA per (object, point block, channel
  // block) = 1024 CTAs of waiting loads got in the way of the step's kernels on the other stream (a whole batch through the
  // zero-copy path: 1.41 ms per step; 16 object rows: 1.27 ms = the gather's stand-alone time, i.e. fully overlapped;
  // 4 rows: 1.79 ms, too few requests in flight) -- profiles/r02_e2e_diag.txt.
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const int n = blockIdx.x * 128 + threadIdx.x;
    const int c0 = blockIdx.z * CH;
    if (n >= N) return;
    int64_t col = choose[(size_t)b * N + n];
    col = col < 0 ? 0 : (col >= hw ? hw - 1 : col);      // torch.gather would raise; never read outside the map
    const float* src = img + ((size_t)b * 32 + c0) * hw + col;
    float e[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) e[j] = __ldg(src + (size_t)j * hw);
    float* dst = emb + ((size_t)b * 32 + c0) * N + n;
#pragma unroll
    for (int j = 0; j < CH; ++j) dst[(size_t)j * N] = e[j];
  }
}

// Channels-last map [B,hw,32]: one warp reads 8 points, lane = channel (one 128-byte line per point).
// grid = (ceil(N / 64), B), block = 256.
__global__ void __launch_bounds__(256)
gather_emb_nhwc_kernel(const float* __restrict__ img, int hw, const int64_t* __restrict__ choose, int N, float* __restrict__ emb)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int n0 = (blockIdx.x * 8 + warp) * 8;
    if (n0 >= N) return;
    int64_t ci = 0;
    if (lane < 8) ci = choose[(size_t)b * N + min(n0 + lane, N - 1)];
    ci = ci < 0 ? 0 : (ci >= hw ? hw - 1 : ci);
    float e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t c = __shfl_sync(0xffffffffu, ci, j);
        e[j] = __ldg(img + ((size_t)b * hw + c) * 32 + lane);
    }
    float* dst = emb + ((size_t)b * 32 + lane) * N + n0;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (n0 + j < N) dst[j] = e[j];
}

}  // namespace ape

extern "C" __attribute__((visibility("default")))
int ape_gather_emb(const float* out_img, int hw, int layout, const int64_t* choose, int B, int N, float* emb, void* stream)
{
    APE_REQUIRE(out_img && choose && emb, "ape_gather_emb: null pointer");
    APE_REQUIRE(B > 0 && N > 0 && hw > 0, "ape_gather_emb: bad sizes");
    APE_REQUIRE(layout == APE_EMB_NCHW || layout == APE_EMB_NHWC, "ape_gather_emb: layout must be APE_EMB_NCHW or APE_EMB_NHWC");
    cudaStream_t s = (cudaStream_t)stream;
    ape::ProfScope prof_("gather_emb", s);
    static int max_y = -1;
    if (max_y < 0) { const char* e = getenv("APE_GATHER_MAX_OBJ_CTAS"); max_y = e ? atoi(e) : 16; }
    if (layout == APE_EMB_NCHW)
        ape::gather_emb_nchw_kernel<8><<<dim3((N + 127) / 128, (max_y > 0 && max_y < B) ? max_y : B, 4), 128, 0, s>>>(out_img, hw, choose, N, emb, B);
    else
        ape::gather_emb_nhwc_kernel<<<dim3((N + 63) / 64, B), 256, 0, s>>>(out_img, hw, choose, N, emb);
    ape::count_launch();
    return ape::check_launch("gather_emb");
}
