// Development micro-benchmark (not product): how should the 500 chosen columns of a HOST-resident encoder map
// [B,32,hw] fp32 reach the GPU?  (a) zero-copy gather from mapped pinned memory, (b) multi-threaded host gather into a
// pinned staging buffer + small H2D, (c) the full-map cudaMemcpyAsync of round 1.  Also times the device-resident gather
// and a channels-last (NHWC) zero-copy gather.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/zc_bench tools/zc_bench.cu -lpthread
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int B = 64, C = 32, HW = 19200, N = 500;

// lane = channel, PTS points per warp; all loads in flight before the first store
template <int PTS>
__global__ void gather_nchw(const float* __restrict__ img, const int64_t* __restrict__ choose, float* __restrict__ emb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * PTS;
    if (n0 >= N) return;
    const float* src = img + ((size_t)b * C + lane) * HW;
    float e[PTS];
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
        const int n = min(n0 + j, N - 1);
        const int64_t c = choose[(size_t)b * N + n];
        e[j] = __ldg(src + c);
    }
    float* dst = emb + ((size_t)b * C + lane) * N + n0;
#pragma unroll
    for (int j = 0; j < PTS; ++j) if (n0 + j < N) dst[j] = e[j];
}

// lane = point (consecutive chosen points), loop over channels: neighbouring lanes hit neighbouring (sorted) columns
template <int CH>
__global__ void gather_nchw_lanept(const float* __restrict__ img, const int64_t* __restrict__ choose, float* __restrict__ emb) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.z * CH;
    if (n >= N) return;
    const int64_t col = choose[(size_t)b * N + n];
    float e[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) e[j] = __ldg(img + ((size_t)b * C + c0 + j) * HW + col);
#pragma unroll
    for (int j = 0; j < CH; ++j) emb[((size_t)b * C + c0 + j) * N + n] = e[j];
}

// channels-last map [B,hw,32]: one warp reads one point's 128 contiguous bytes
__global__ void gather_nhwc(const float* __restrict__ img, const int64_t* __restrict__ choose, float* __restrict__ emb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 8;
    if (n0 >= N) return;
    float e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = min(n0 + j, N - 1);
        e[j] = __ldg(img + ((size_t)b * HW + choose[(size_t)b * N + n]) * C + lane);
    }
    float* dst = emb + ((size_t)b * C + lane) * N + n0;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (n0 + j < N) dst[j] = e[j];
}

static void host_gather(const float* img, const int64_t* choose, float* out, int threads) {
    std::vector<std::thread> th;
    const int rows = B * C;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=] {
            for (int r = t; r < rows; r += threads) {
                const int b = r / C;
                const float* src = img + (size_t)r * HW;
                const int64_t* ch = choose + (size_t)b * N;
                float* dst = out + (size_t)r * N;
                for (int n = 0; n < N; ++n) dst[n] = src[ch[n]];
            }
        });
    for (auto& x : th) x.join();
}

template <typename F>
static float time_ms(F f, int reps, cudaStream_t s) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1, s); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int nsets = 3;
    const size_t map_elems = (size_t)B * C * HW;
    printf("hardware_concurrency=%u\n", std::thread::hardware_concurrency());
    float* h_img[nsets]; float* h_nhwc;
    for (int i = 0; i < nsets; ++i) CK(cudaHostAlloc(&h_img[i], map_elems * 4, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_nhwc, map_elems * 4, cudaHostAllocDefault));
    int64_t* h_choose; CK(cudaHostAlloc(&h_choose, (size_t)B * N * 8, cudaHostAllocDefault));
    float* h_stage; CK(cudaHostAlloc(&h_stage, (size_t)B * C * N * 4, cudaHostAllocDefault));
    std::mt19937 rng(1);
    for (int i = 0; i < nsets; ++i) for (size_t j = 0; j < map_elems; ++j) h_img[i][j] = (float)((j * 2654435761u + i) & 0xffff) / 65536.f;
    for (int b = 0; b < B; ++b) for (int p = 0; p < HW; ++p) for (int c = 0; c < C; ++c)
        h_nhwc[((size_t)b * HW + p) * C + c] = h_img[0][((size_t)b * C + c) * HW + p];
    for (int b = 0; b < B; ++b) {
        std::vector<int> all(HW); for (int i = 0; i < HW; ++i) all[i] = i;
        std::shuffle(all.begin(), all.end(), rng);
        std::sort(all.begin(), all.begin() + N);
        for (int n = 0; n < N; ++n) h_choose[(size_t)b * N + n] = all[n];
    }
    float *d_img, *d_emb, *d_emb2; int64_t* d_choose;
    CK(cudaMalloc(&d_img, map_elems * 4)); CK(cudaMalloc(&d_emb, (size_t)B * C * N * 4)); CK(cudaMalloc(&d_emb2, (size_t)B * C * N * 4));
    CK(cudaMalloc(&d_choose, (size_t)B * N * 8));
    CK(cudaMemcpy(d_choose, h_choose, (size_t)B * N * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_img, h_img[0], map_elems * 4, cudaMemcpyHostToDevice));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    int k = 0;
    // reference result (device-resident gather)
    gather_nchw<8><<<dim3((N + 63) / 64, B), 256, 0, s>>>(d_img, d_choose, d_emb2);
    CK(cudaStreamSynchronize(s));
    std::vector<float> ref((size_t)B * C * N), got((size_t)B * C * N);
    CK(cudaMemcpy(ref.data(), d_emb2, ref.size() * 4, cudaMemcpyDeviceToHost));
    auto check = [&](const char* what) {
        cudaMemcpy(got.data(), d_emb, got.size() * 4, cudaMemcpyDeviceToHost);
        size_t bad = 0; for (size_t i = 0; i < got.size(); ++i) bad += got[i] != ref[i];
        printf("   check %-28s mismatches=%zu\n", what, bad);
    };
    const double useful_mb = (double)B * C * N * 4 / 1e6;
    float ms;
    ms = time_ms([&] { cudaMemcpyAsync(d_img, h_img[k++ % nsets], map_elems * 4, cudaMemcpyHostToDevice, s); }, 10, s);
    printf("full map H2D memcpy           : %8.3f ms  (%.1f GB/s)\n", ms, map_elems * 4 / ms / 1e6);
    CK(cudaMemcpy(d_img, h_img[0], map_elems * 4, cudaMemcpyHostToDevice));
    ms = time_ms([&] { gather_nchw<8><<<dim3((N + 63) / 64, B), 256, 0, s>>>(d_img, d_choose, d_emb); }, 20, s);
    printf("device-resident gather nchw<8>: %8.3f ms\n", ms);
    ms = time_ms([&] { gather_nchw<8><<<dim3((N + 63) / 64, B), 256, 0, s>>>(h_img[k++ % nsets], d_choose, d_emb); }, 10, s);
    printf("zero-copy gather nchw<8>      : %8.3f ms  (%.1f GB/s useful, %.1f GB/s of 32B sectors)\n", ms, useful_mb / ms, useful_mb * 8 / ms);
    gather_nchw<8><<<dim3((N + 63) / 64, B), 256, 0, s>>>(h_img[0], d_choose, d_emb); cudaStreamSynchronize(s); check("zero-copy nchw<8>");
    ms = time_ms([&] { gather_nchw<16><<<dim3((N + 127) / 128, B), 256, 0, s>>>(h_img[k++ % nsets], d_choose, d_emb); }, 10, s);
    printf("zero-copy gather nchw<16>     : %8.3f ms\n", ms);
    ms = time_ms([&] { gather_nchw<4><<<dim3((N + 31) / 32, B), 256, 0, s>>>(h_img[k++ % nsets], d_choose, d_emb); }, 10, s);
    printf("zero-copy gather nchw<4>      : %8.3f ms\n", ms);
    ms = time_ms([&] { gather_nchw_lanept<8><<<dim3((N + 127) / 128, B, 4), 128, 0, s>>>(h_img[k++ % nsets], d_choose, d_emb); }, 10, s);
    printf("zero-copy gather lane=pt <8>  : %8.3f ms\n", ms);
    gather_nchw_lanept<8><<<dim3((N + 127) / 128, B, 4), 128, 0, s>>>(h_img[0], d_choose, d_emb); cudaStreamSynchronize(s); check("zero-copy lane=pt");
    ms = time_ms([&] { gather_nchw_lanept<16><<<dim3((N + 63) / 64, B, 2), 64, 0, s>>>(h_img[k++ % nsets], d_choose, d_emb); }, 10, s);
    printf("zero-copy gather lane=pt <16> : %8.3f ms\n", ms);
    ms = time_ms([&] { gather_nhwc<<<dim3((N + 63) / 64, B), 256, 0, s>>>(h_nhwc, d_choose, d_emb); }, 10, s);
    printf("zero-copy gather nhwc (128 B) : %8.3f ms  (%.1f GB/s useful)\n", ms, useful_mb / ms);
    gather_nhwc<<<dim3((N + 63) / 64, B), 256, 0, s>>>(h_nhwc, d_choose, d_emb); cudaStreamSynchronize(s); check("zero-copy nhwc");
    // host gather
    for (int t : {1, 2, 4, 8, 16, 32, 64}) {
        if (t > 2 * (int)std::thread::hardware_concurrency()) break;
        host_gather(h_img[0], h_choose, h_stage, t);
        auto t0 = std::chrono::steady_clock::now();
        const int reps = 6;
        for (int i = 0; i < reps; ++i) host_gather(h_img[(i + 1) % nsets], h_choose, h_stage, t);
        const double hms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        printf("host gather %2d threads        : %8.3f ms\n", t, hms);
    }
    ms = time_ms([&] { cudaMemcpyAsync(d_emb, h_stage, (size_t)B * C * N * 4, cudaMemcpyHostToDevice, s); }, 20, s);
    printf("staging H2D (4.1 MB)          : %8.3f ms  (%.1f GB/s)\n", ms, useful_mb / ms);
    return 0;
}
