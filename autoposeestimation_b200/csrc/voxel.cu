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

}  // namespace ape

extern "C" __attribute__((visibility("default")))
int ape_voxel_down_sample(const double* points, const int32_t* offset, int n_clouds, double voxel_size,
                          double* out_points, int32_t* out_counts, void* stream)
{
    APE_REQUIRE(points && offset && out_points && out_counts, "ape_voxel_down_sample: null pointer");
    APE_REQUIRE(n_clouds >= 0 && voxel_size > 0.0, "ape_voxel_down_sample: bad sizes (open3d raises for voxel_size <= 0)");
    if (n_clouds == 0) return APE_OK;
    const int smem = ape::kVoxMax * 8;
    static bool attr_set = false;
    if (!attr_set) {
        APE_CUDA(cudaFuncSetAttribute(ape::voxel_down_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    ape::ProfScope prof_("voxel_down_sample", (cudaStream_t)stream);
    ape::voxel_down_sample_kernel<<<n_clouds, ape::kVoxThreads, smem, (cudaStream_t)stream>>>(points, offset, voxel_size,
                                                                                            out_points, out_counts);
    ape::count_launch();
    return ape::check_launch("ape_voxel_down_sample");
}
