// Voxel-grid down-sampling, open3d 0.9 semantics (sm_100a).  Replaces pcd.voxel_down_sample(v) at
// pc_reconstruction/open3d_utils.py:21, :198 and create_pointcloud.py:312:
//   origin = min_bound - v/2;  voxel = floor((p - origin)/v);  output = mean of the points of a voxel.
// open3d emits voxels in unordered_map order (not reproducible); this kernel emits them sorted by
// (ix,iy,iz) and sums the points of a voxel in ascending input index, which is exactly what
// oracle/icp.py:voxel_down_sample does, so the two agree bit for bit.
// One CTA per cloud: (voxel key | point index) packed in 64 bits, bitonic sort in shared memory,
// segmented mean over the sorted runs.
#include "ape_common.cuh"
#include <cfloat>

namespace ape {

constexpr int kVoxThreads = 1024;
constexpr int kVoxMax = APE_VOXEL_MAX_POINTS;        // 16384 keys * 8 B = 128 KB shared memory

__global__ void __launch_bounds__(kVoxThreads)
voxel_down_sample_kernel(const double* __restrict__ points, const int32_t* __restrict__ offset, double voxel,
                         double* __restrict__ out_points, int32_t* __restrict__ out_counts)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ double s_min[3][kVoxThreads / 32];
    __shared__ double s_origin[3];
    __shared__ int s_wtot[kVoxThreads / 32];
    __shared__ int s_carry;
    __shared__ int s_bad;

    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = offset[c], n = offset[c + 1] - p0;
    const double* P = points + 3 * (size_t)p0;
    double* O = out_points + 3 * (size_t)p0;
    if (n <= 0) { if (tid == 0) out_counts[c] = 0; return; }
    if (n > kVoxMax) { if (tid == 0) out_counts[c] = -1; return; }

    // ---- min bound
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
    for (int i = tid; i < n; i += kVoxThreads)
#pragma unroll
        for (int a = 0; a < 3; ++a) lo[a] = fmin(lo[a], P[3 * i + a]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        if (lane == 0) s_min[a][warp] = lo[a];
    }
    if (tid == 0) { s_bad = 0; s_carry = 0; }
    __syncthreads();
    if (tid < 3) {
        double m = DBL_MAX;
        for (int w = 0; w < kVoxThreads / 32; ++w) m = fmin(m, s_min[tid][w]);
        s_origin[tid] = __dsub_rn(m, __dmul_rn(voxel, 0.5));
    }
    __syncthreads();
    const double ox = s_origin[0], oy = s_origin[1], oz = s_origin[2];

    // ---- keys (16 bits per axis | 14-bit point index), padded to a power of two with ~0
    int npad = 1;
    while (npad < n) npad <<= 1;
    for (int i = tid; i < npad; i += kVoxThreads) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const double fx = floor(__ddiv_rn(__dsub_rn(P[3 * i], ox), voxel));
            const double fy = floor(__ddiv_rn(__dsub_rn(P[3 * i + 1], oy), voxel));
            const double fz = floor(__ddiv_rn(__dsub_rn(P[3 * i + 2], oz), voxel));
            if (!(fx >= 0.0 && fx < 65536.0 && fy >= 0.0 && fy < 65536.0 && fz >= 0.0 && fz < 65536.0)) s_bad = 1;
            k = ((unsigned long long)(unsigned)fx << 46) | ((unsigned long long)(unsigned)fy << 30) |
                ((unsigned long long)(unsigned)fz << 14) | (unsigned long long)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) out_counts[c] = -2; return; }     // > 65535 voxels along an axis

    // ---- bitonic sort (ascending)
    for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (npad >> 1); t += kVoxThreads) {
                const int i = 2 * t - (t & (stride - 1));      // lower element of the pair
                const int j = i + stride;
                const bool up = ((i & size) == 0);
                const unsigned long long a = keys[i], b = keys[j];
                if ((a > b) == up) { keys[i] = b; keys[j] = a; }
            }
            __syncthreads();
        }
    }

    // ---- segmented mean over runs of equal voxel key; run rank by block scan of the head flags
    for (int base = 0; base < n; base += kVoxThreads) {
        const int i = base + tid;
        bool head = false;
        unsigned long long vk = 0;
        if (i < n) {
            vk = keys[i] >> 14;
            head = (i == 0) || ((keys[i - 1] >> 14) != vk);
        }
        int incl = head ? 1 : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_wtot[w];
        if (head) {
            double sx = 0.0, sy = 0.0, sz = 0.0;
            int cnt = 0;
            for (int j = i; j < n && (keys[j] >> 14) == vk; ++j) {
                const int pi = (int)(keys[j] & 0x3fffull);
                sx = __dadd_rn(sx, P[3 * pi]); sy = __dadd_rn(sy, P[3 * pi + 1]); sz = __dadd_rn(sz, P[3 * pi + 2]);
                ++cnt;
            }
            const int slot = before + incl - 1;
            const double dc = (double)cnt;
            O[3 * slot] = __ddiv_rn(sx, dc); O[3 * slot + 1] = __ddiv_rn(sy, dc); O[3 * slot + 2] = __ddiv_rn(sz, dc);
        }
        __syncthreads();
        if (tid == kVoxThreads - 1) s_carry = before + incl;
        __syncthreads();
    }
    if (tid == 0) out_counts[c] = s_carry;
}

// ------------------------------------------------------------------------------------------------------------------
// Clouds of more than APE_VOXEL_MAX_POINTS points (a 480 x 640 frame can label 300 k pixels; a merged object cloud plus a
// new view exceeds 16 k routinely): same result, bit for bit, through global memory.
//   keys     (13 bits of voxel index per axis | 25-bit point index) -- unique, so any comparison sort gives the order
//            "voxel (ix,iy,iz), then ascending input index" that the single-CTA kernel and the oracle use;
//   sort     16 384-key chunks by the shared-memory bitonic network, then log2(#chunks) merge-path passes (one 4096-key
//            output tile per CTA, two binary searches + a sequential 16-key merge per thread);
//   means    head flags per 1024-key tile -> tile counts -> one-CTA exclusive scan -> every run head sums its points in
//            key order (= ascending input index) and writes slot = tile base + rank inside the tile.
constexpr int kVoxIdxBits = 25, kVoxAxisBits = 13;
constexpr int kVoxChunk = APE_VOXEL_MAX_POINTS;              // keys per bitonic chunk
constexpr int kVoxTile = 4096;                               // keys per merge CTA (256 threads x 16)
constexpr int kVoxPart = 128;                                // partial-minimum CTAs

__global__ void __launch_bounds__(256)
vox_minbound_kernel(const double* __restrict__ P, int n, double* __restrict__ partial /* [kVoxPart][3] */)
{
    __shared__ double s_min[3][8];
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
#pragma unroll
        for (int a = 0; a < 3; ++a) lo[a] = fmin(lo[a], P[3 * (size_t)i + a]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        if ((threadIdx.x & 31) == 0) s_min[a][threadIdx.x >> 5] = lo[a];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double m = DBL_MAX;
        for (int w = 0; w < 8; ++w) m = fmin(m, s_min[threadIdx.x][w]);
        partial[3 * blockIdx.x + threadIdx.x] = m;
    }
}

__global__ void __launch_bounds__(256)
vox_keys_kernel(const double* __restrict__ P, int n, int n_pad, double voxel, const double* __restrict__ partial, int n_partial,
                unsigned long long* __restrict__ keys, int* __restrict__ bad)
{
    __shared__ double s_origin[3];
    if (threadIdx.x < 3) {
        double m = DBL_MAX;
        for (int g = 0; g < n_partial; ++g) m = fmin(m, partial[3 * g + threadIdx.x]);       // fmin is exact: any order
        s_origin[threadIdx.x] = __dsub_rn(m, __dmul_rn(voxel, 0.5));
    }
    __syncthreads();
    const double ox = s_origin[0], oy = s_origin[1], oz = s_origin[2];
    const double lim = (double)(1 << kVoxAxisBits);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_pad; i += gridDim.x * 256) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const double fx = floor(__ddiv_rn(__dsub_rn(P[3 * (size_t)i], ox), voxel));
            const double fy = floor(__ddiv_rn(__dsub_rn(P[3 * (size_t)i + 1], oy), voxel));
            const double fz = floor(__ddiv_rn(__dsub_rn(P[3 * (size_t)i + 2], oz), voxel));
            if (!(fx >= 0.0 && fx < lim && fy >= 0.0 && fy < lim && fz >= 0.0 && fz < lim)) { *bad = 1; }
            else k = ((unsigned long long)(unsigned)fx << (kVoxIdxBits + 2 * kVoxAxisBits)) |
                     ((unsigned long long)(unsigned)fy << (kVoxIdxBits + kVoxAxisBits)) |
                     ((unsigned long long)(unsigned)fz << kVoxIdxBits) | (unsigned long long)i;
        }
        keys[i] = k;
    }
}

__global__ void __launch_bounds__(kVoxThreads)
vox_sort_chunks_kernel(unsigned long long* __restrict__ keys)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* g = keys + (size_t)blockIdx.x * kVoxChunk;
    for (int i = threadIdx.x; i < kVoxChunk; i += kVoxThreads) s[i] = g[i];
    __syncthreads();
    for (int size = 2; size <= kVoxChunk; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (kVoxChunk >> 1); t += kVoxThreads) {
                const int i = 2 * t - (t & (stride - 1));
                const int j = i + stride;
                const bool up = ((i & size) == 0);
                const unsigned long long a = s[i], b = s[j];
                if ((a > b) == up) { s[i] = b; s[j] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < kVoxChunk; i += kVoxThreads) g[i] = s[i];
}

// number of elements taken from A among the first d outputs of merge(A[0:la], B[0:lb]) (keys are unique)
__device__ __forceinline__ int merge_path(const unsigned long long* A, int la, const unsigned long long* B, int lb, int d) {
    int lo = max(0, d - lb), hi = min(d, la);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] < B[d - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
vox_merge_kernel(const unsigned long long* __restrict__ src, unsigned long long* __restrict__ dst, int n_pad, int run)
{
    __shared__ unsigned long long s_in[kVoxTile];
    __shared__ int s_split[2];
    const int tiles_per_pair = 2 * run / kVoxTile;
    const int pair = blockIdx.x / tiles_per_pair, tile = blockIdx.x - pair * tiles_per_pair;
    const size_t base = (size_t)pair * 2 * run;
    const int la = (int)min((size_t)run, (size_t)n_pad - base);
    const int lb = (int)min((size_t)run, (size_t)n_pad - base - la);
    const unsigned long long* A = src + base;
    const unsigned long long* B = A + la;
    const int d0 = tile * kVoxTile, d1 = min(d0 + kVoxTile, la + lb);
    if (d0 >= la + lb) return;
    if (threadIdx.x < 2) s_split[threadIdx.x] = merge_path(A, la, B, lb, threadIdx.x == 0 ? d0 : d1);
    __syncthreads();
    const int i0 = s_split[0], i1 = s_split[1], j0 = d0 - i0, j1 = d1 - i1;
    const int na = i1 - i0, nb = j1 - j0;
    for (int k = threadIdx.x; k < na; k += 256) s_in[k] = A[i0 + k];
    for (int k = threadIdx.x; k < nb; k += 256) s_in[na + k] = B[j0 + k];
    __syncthreads();
    const unsigned long long* sa = s_in;
    const unsigned long long* sb = s_in + na;
    const int per = kVoxTile / 256;
    const int t0 = min(threadIdx.x * per, na + nb), t1 = min(t0 + per, na + nb);
    int ia = merge_path(sa, na, sb, nb, t0), ib = t0 - ia;
    unsigned long long* out = dst + base + d0;
    for (int k = t0; k < t1; ++k) {
        const bool take_a = ib >= nb || (ia < na && sa[ia] < sb[ib]);
        out[k] = take_a ? sa[ia++] : sb[ib++];
    }
}

__global__ void __launch_bounds__(1024)
vox_count_heads_kernel(const unsigned long long* __restrict__ keys, int n, int32_t* __restrict__ tile_count)
{
    __shared__ int s_w[32];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const int head = (i < n) && (i == 0 || (keys[i - 1] >> kVoxIdxBits) != (keys[i] >> kVoxIdxBits));
    const int c = warp_sum(head);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += s_w[w];
        tile_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
vox_scan_tiles_kernel(const int32_t* __restrict__ tile_count, int n_tiles, int32_t* __restrict__ tile_base, int32_t* __restrict__ out_count,
                      const int* __restrict__ bad)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_tiles; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int x = i < n_tiles ? tile_count[i] : 0;
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (i < n_tiles) tile_base[i] = before + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = *bad ? -2 : s_carry;
}

__global__ void __launch_bounds__(1024)
vox_emit_kernel(const double* __restrict__ P, const unsigned long long* __restrict__ keys, int n, const int32_t* __restrict__ tile_base,
                double* __restrict__ O)
{
    __shared__ int s_w[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 1024 + threadIdx.x;
    unsigned long long vk = 0;
    bool head = false;
    if (i < n) {
        vk = keys[i] >> kVoxIdxBits;
        head = (i == 0) || ((keys[i - 1] >> kVoxIdxBits) != vk);
    }
    int incl = head ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (!head) return;
    int slot = tile_base[blockIdx.x] + incl - 1;
    for (int w = 0; w < warp; ++w) slot += s_w[w];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int cnt = 0;
    for (int j = i; j < n && (keys[j] >> kVoxIdxBits) == vk; ++j) {
        const size_t pi = (size_t)(keys[j] & ((1ull << kVoxIdxBits) - 1ull));
        sx = __dadd_rn(sx, P[3 * pi]); sy = __dadd_rn(sy, P[3 * pi + 1]); sz = __dadd_rn(sz, P[3 * pi + 2]);
        ++cnt;
    }
    const double dc = (double)cnt;
    O[3 * (size_t)slot] = __ddiv_rn(sx, dc); O[3 * (size_t)slot + 1] = __ddiv_rn(sy, dc); O[3 * (size_t)slot + 2] = __ddiv_rn(sz, dc);
}

}  // namespace ape

extern "C" __attribute__((visibility("default")))
int ape_voxel_down_sample(const double* points, const int32_t* offset, int n_clouds, double voxel_size,
                          double* out_points, int32_t* out_counts, void* stream)
{
    APE_REQUIRE(points && offset && out_points && out_counts, "ape_voxel_down_sample: null pointer");
    APE_REQUIRE(n_clouds >= 0 && voxel_size > 0.0, "ape_voxel_down_sample: bad sizes (open3d raises for voxel_size <= 0)");
    if (n_clouds == 0) return APE_OK;
    const int smem = ape::kVoxMax * 8;
    static ape::PerDevice attr_done;
    if (attr_done.first()) {
        APE_CUDA(cudaFuncSetAttribute(ape::voxel_down_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    ape::ProfScope prof_("voxel_down_sample", (cudaStream_t)stream);
    ape::voxel_down_sample_kernel<<<n_clouds, ape::kVoxThreads, smem, (cudaStream_t)stream>>>(points, offset, voxel_size,
                                                                                            out_points, out_counts);
    ape::count_launch();
    return ape::check_launch("ape_voxel_down_sample");
}

// One cloud of any size (n_points given by value: the host knows it), same output as the batched kernel.
extern "C" __attribute__((visibility("default")))
int ape_voxel_down_sample_large(const double* points, int n_points, double voxel_size, double* out_points, int32_t* out_count,
                                void* stream)
{
    APE_REQUIRE(points && out_points && out_count, "ape_voxel_down_sample_large: null pointer");
    APE_REQUIRE(n_points >= 0 && voxel_size > 0.0, "ape_voxel_down_sample_large: bad sizes");
    APE_REQUIRE(n_points < (1 << ape::kVoxIdxBits), "ape_voxel_down_sample_large: at most 2^25 - 1 points");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_points == 0) { APE_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s)); return APE_OK; }
    const int n_pad = (n_points + ape::kVoxChunk - 1) / ape::kVoxChunk * ape::kVoxChunk;
    const int n_tiles = (n_points + 1023) / 1024;
    // stream-ordered scratch: two key buffers, partial minima, tile counts / bases, range flag
    const size_t key_bytes = (size_t)n_pad * 8;
    const size_t bytes = 2 * key_bytes + ape::kVoxPart * 3 * sizeof(double) + 2 * (size_t)n_tiles * 4 + 16;
    unsigned char* scratch = nullptr;
    APE_CUDA(cudaMallocAsync((void**)&scratch, bytes, s));
    unsigned long long* ka = reinterpret_cast<unsigned long long*>(scratch);
    unsigned long long* kb = ka + n_pad;
    double* partial = reinterpret_cast<double*>(kb + n_pad);
    int32_t* tile_count = reinterpret_cast<int32_t*>(partial + ape::kVoxPart * 3);
    int32_t* tile_base = tile_count + n_tiles;
    int* bad = reinterpret_cast<int*>(tile_base + n_tiles);
    static ape::PerDevice attr_done;
    if (attr_done.first()) {
        APE_CUDA(cudaFuncSetAttribute(ape::vox_sort_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ape::kVoxChunk * 8));
    }
    ape::ProfScope prof_("voxel_down_sample_large", s);
    APE_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    ape::vox_minbound_kernel<<<ape::kVoxPart, 256, 0, s>>>(points, n_points, partial);
    ape::vox_keys_kernel<<<(n_pad + 255) / 256 < 1184 ? (n_pad + 255) / 256 : 1184, 256, 0, s>>>(points, n_points, n_pad, voxel_size, partial,
                                                                                                 ape::kVoxPart, ka, bad);
    ape::vox_sort_chunks_kernel<<<n_pad / ape::kVoxChunk, ape::kVoxThreads, ape::kVoxChunk * 8, s>>>(ka);
    unsigned long long *src = ka, *dst = kb;
    int launches = 3;
    for (long long run = ape::kVoxChunk; run < n_pad; run *= 2) {
        const int pairs = (int)((n_pad + 2 * run - 1) / (2 * run));
        ape::vox_merge_kernel<<<pairs * (int)(2 * run / ape::kVoxTile), 256, 0, s>>>(src, dst, n_pad, (int)run);
        unsigned long long* t = src; src = dst; dst = t;
        ++launches;
    }
    ape::vox_count_heads_kernel<<<n_tiles, 1024, 0, s>>>(src, n_points, tile_count);
    ape::vox_scan_tiles_kernel<<<1, 1024, 0, s>>>(tile_count, n_tiles, tile_base, out_count, bad);
    ape::vox_emit_kernel<<<n_tiles, 1024, 0, s>>>(points, src, n_points, tile_base, out_points);
    ape::count_launch(launches + 3);
    int rc = ape::check_launch("ape_voxel_down_sample_large");
    cudaFreeAsync(scratch, s);
    return rc;
}
