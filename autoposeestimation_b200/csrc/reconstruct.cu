// The sequential register-and-merge loop of the object-cloud reconstruction (pc_reconstruction/create_pointcloud.py:286-312)
// as ONE host call without a host synchronisation: per further view
//     target_down = voxel(cloud), source_down = voxel(view)          one launch over the two clouds   (open3d_utils.py:76-77)
//     T = ICP(source_down -> target_down)                             one launch (512-thread CTA)       (:96-104)
//     cloud = voxel([T * source_down ; target_down])                  transform + concatenate, one voxel launch (:299-312)
// All sizes stay on the device: the kernels below only glue the ragged layouts together (offsets / counts that the voxel
// grid and the ICP kernel read from device memory).  The loop is sequential by construction (every view registers against
// the cloud accumulated so far): what this removes is the six host round trips per view of the Python loop.
#include "ape_common.cuh"

namespace ape {

// ints: [0] = 0, [1..2] = voxel counts of (cloud, view), [3..5] = offsets of the packed pair, [6..7] = offsets of the merged
// cloud, [8] = points in the accumulated cloud, [9] = status (0, or the first negative voxel status seen)
enum { kZero = 0, kCnt2 = 1, kOff2 = 3, kMoff = 6, kCloud = 8, kStatus = 9, kInts = 16 };

__global__ void recon_init_kernel(int32_t* ints, int n0)
{
    if (threadIdx.x < kInts) ints[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) ints[kCloud] = n0;
}

// pack[cloud .. cloud + n_view) = view; offsets of the pair (cloud, view)
__global__ void recon_pack_kernel(const double* __restrict__ view, int n_view, double* __restrict__ pack, int32_t* __restrict__ ints)
{
    int cnt = ints[kCloud];
    if (cnt < 0) cnt = 0;                                   // a failed voxel grid: the status is already recorded
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n_view; i += gridDim.x * blockDim.x) pack[3 * (size_t)cnt + i] = view[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { ints[kOff2] = 0; ints[kOff2 + 1] = cnt; ints[kOff2 + 2] = cnt + n_view; }
}

// merged = [T * source_down ; target_down]  (source first, create_pointcloud.py:299-301), its offsets, status
__global__ void recon_merge_kernel(const double* __restrict__ down, const double* __restrict__ T, double* __restrict__ merged,
                                   int32_t* __restrict__ ints)
{
    const int nt_raw = ints[kCnt2], ns_raw = ints[kCnt2 + 1], s0 = ints[kOff2 + 1];
    const int nt = nt_raw < 0 ? 0 : nt_raw, ns = ns_raw < 0 ? 0 : ns_raw;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const double r00 = T[0], r01 = T[1], r02 = T[2], t0 = T[3], r10 = T[4], r11 = T[5], r12 = T[6], t1 = T[7],
                 r20 = T[8], r21 = T[9], r22 = T[10], t2 = T[11];
    for (int i = tid; i < ns; i += nthr) {
        const double x = down[3 * (size_t)(s0 + i)], y = down[3 * (size_t)(s0 + i) + 1], z = down[3 * (size_t)(s0 + i) + 2];
        merged[3 * (size_t)i] = x * r00 + y * r01 + z * r02 + t0;
        merged[3 * (size_t)i + 1] = x * r10 + y * r11 + z * r12 + t1;
        merged[3 * (size_t)i + 2] = x * r20 + y * r21 + z * r22 + t2;
    }
    for (int i = tid; i < 3 * nt; i += nthr) merged[3 * (size_t)ns + i] = down[i];
    if (tid == 0) {
        ints[kMoff] = 0; ints[kMoff + 1] = ns + nt;
        if (ints[kStatus] == 0 && (nt_raw < 0 || ns_raw < 0)) ints[kStatus] = nt_raw < 0 ? nt_raw : ns_raw;
    }
}

__global__ void recon_finish_kernel(int32_t* __restrict__ ints, int32_t* __restrict__ out_count, int32_t* __restrict__ out_status)
{
    const int c = ints[kCloud];
    int st = ints[kStatus];
    if (st == 0 && c < 0) st = c;
    *out_count = c < 0 ? 0 : c;
    *out_status = st;
}

}  // namespace ape

extern "C" __attribute__((visibility("default")))
size_t ape_reconstruct_work_bytes(int total_points)
{
    const size_t p = (size_t)(total_points > 0 ? total_points : 0);
    return 3 * (24 * p) + 256 + 16 * 8 + 4 * 8 + ape_icp_work_bytes((int)p, (int)p) + 64;
}

extern "C" __attribute__((visibility("default")))
int ape_reconstruct_run(const double* points, const int32_t* offset_host, int n_views, double voxel_size, double threshold,
                        double rel_fitness, double rel_rmse, int max_iter, double* out_points, int32_t* out_count,
                        int32_t* out_status, void* work, void* stream)
{
    APE_REQUIRE(offset_host && out_points && out_count && out_status && work, "ape_reconstruct_run: null pointer");
    APE_REQUIRE(n_views >= 0 && voxel_size > 0.0 && threshold > 0.0, "ape_reconstruct_run: bad sizes");
    APE_REQUIRE((((uintptr_t)work) & 7) == 0, "ape_reconstruct_run: work must be 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const int P = n_views > 0 ? offset_host[n_views] : 0;
    APE_REQUIRE(P >= 0 && (P == 0 || points), "ape_reconstruct_run: bad offsets");
    for (int v = 0; v < n_views; ++v) APE_REQUIRE(offset_host[v + 1] >= offset_host[v], "ape_reconstruct_run: offsets must not decrease");
    unsigned char* w = reinterpret_cast<unsigned char*>(work);
    double* pack = reinterpret_cast<double*>(w);   w += 24 * (size_t)P;
    double* down = reinterpret_cast<double*>(w);   w += 24 * (size_t)P;
    double* merged = reinterpret_cast<double*>(w); w += 24 * (size_t)P;
    int32_t* ints = reinterpret_cast<int32_t*>(w); w += 256;
    double* T = reinterpret_cast<double*>(w);      w += 16 * 8;
    double* info = reinterpret_cast<double*>(w);   w += 4 * 8;
    void* icp_work = w;
    int rc;
    bool started = false;
    for (int v = 0; v < n_views; ++v) {
        const int n = offset_host[v + 1] - offset_host[v];
        if (n == 0) continue;                               // create_pointcloud.py:288 (`if len(source.points) == 0: continue`)
        const double* view = points + 3 * (size_t)offset_host[v];
        if (!started) {                                     // :290-292: the first non-empty surface starts the cloud
            ape::recon_init_kernel<<<1, 32, 0, s>>>(ints, n);
            APE_CUDA(cudaMemcpyAsync(pack, view, 24 * (size_t)n, cudaMemcpyDeviceToDevice, s));
            ape::count_launch();
            started = true;
            continue;
        }
        ape::recon_pack_kernel<<<(3 * n + 255) / 256, 256, 0, s>>>(view, n, pack, ints);
        ape::count_launch();
        if ((rc = ape::check_launch("recon_pack"))) return rc;
        if ((rc = ape_voxel_down_sample(pack, ints + ape::kOff2, 2, voxel_size, down, ints + ape::kCnt2, s))) return rc;
        if ((rc = ape_icp_p2p_ex(down, ints + ape::kOff2 + 1, ints + ape::kCnt2 + 1, down, ints + ape::kZero, 1, P, P, threshold,
                                 rel_fitness, rel_rmse, max_iter, nullptr, T, info, icp_work, s))) return rc;
        ape::recon_merge_kernel<<<32, 256, 0, s>>>(down, T, merged, ints);
        ape::count_launch();
        if ((rc = ape::check_launch("recon_merge"))) return rc;
        if ((rc = ape_voxel_down_sample(merged, ints + ape::kMoff, 1, voxel_size, pack, ints + ape::kCloud, s))) return rc;
    }
    if (!started) {
        APE_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
        APE_CUDA(cudaMemsetAsync(out_status, 0, sizeof(int32_t), s));
        return APE_OK;
    }
    APE_CUDA(cudaMemcpyAsync(out_points, pack, 24 * (size_t)P, cudaMemcpyDeviceToDevice, s));
    ape::recon_finish_kernel<<<1, 1, 0, s>>>(ints, out_count, out_status);
    ape::count_launch();
    return ape::check_launch("recon_finish");
}
