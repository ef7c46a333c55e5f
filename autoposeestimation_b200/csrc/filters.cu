// Point-cloud outlier filters between back-projection and ICP (sm_100a), open3d 0.9.0 semantics.  Replaces, at
// pc_reconstruction/open3d_utils.py:158-166 and :198-211,
//   pcd.remove_radius_outlier(nb_points, radius)           keep i  <=>  #{j : |p_j - p_i|^2 < radius^2} > nb_points   (self counts)
//   pcd.compute_mahalanobis_distance()                     sqrt((p - mean)^T cov^-1 (p - mean)), population covariance from cumulants
//   pcd.remove_statistical_outlier(nb_neighbors, ratio)    avg_i = mean of the distances to the nb_neighbors nearest points (self
//                                                          included, summed in ascending order); keep i <=> 0 < avg_i < mean + ratio*std
//                                                          (std with Bessel's correction over the valid points)
// Ragged batches: points [P,3] fp64 with int32 offsets [C+1] (the layout of ape_voxel_down_sample / ape_icp_p2p).  Neighbour
// searches are tiled brute force: the cloud streams through shared memory in 1024-point tiles, each thread owns one query,
// distances are fp64 ((dx^2 + dy^2) + dz^2 with every operation rounded, as FLANN's L2 functor computes them), so the
// counts / neighbour sets are those of an exact search.  Clouds here are 2-5 k points after the 2 mm voxel grid.
#include "ape_common.cuh"
#include <cfloat>

namespace ape {

constexpr int kFiltThreads = 256;
constexpr int kFiltTile = 1024;
constexpr int kFiltMaxK = 64;

__device__ __forceinline__ double dist2(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// grid = (ceil(max_cloud / 256), C)
__global__ void __launch_bounds__(kFiltThreads)
radius_outlier_kernel(const double* __restrict__ points, const int32_t* __restrict__ offset, int nb_points, double radius2,
                      uint8_t* __restrict__ keep, int32_t* __restrict__ n_neighbors)
{
    __shared__ double s_p[kFiltTile][3];
    const int c = blockIdx.y;
    const int p0 = offset[c], n = offset[c + 1] - p0;
    if ((int)(blockIdx.x * kFiltThreads) >= n) return;
    const double* P = points + 3 * (size_t)p0;
    const int i = blockIdx.x * kFiltThreads + threadIdx.x;
    const bool live = i < n;
    const int ii = live ? i : n - 1;
    const double qx = P[3 * ii], qy = P[3 * ii + 1], qz = P[3 * ii + 2];
    int cnt = 0;
    for (int t0 = 0; t0 < n; t0 += kFiltTile) {
        const int m = min(kFiltTile, n - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < 3 * m; k += kFiltThreads) (&s_p[0][0])[k] = P[3 * (size_t)t0 + k];
        __syncthreads();
        for (int j = 0; j < m; ++j) cnt += dist2(s_p[j][0], s_p[j][1], s_p[j][2], qx, qy, qz) < radius2 ? 1 : 0;
    }
    if (live) {
        keep[p0 + i] = cnt > nb_points ? 1 : 0;
        if (n_neighbors) n_neighbors[p0 + i] = cnt;
    }
}

// avg[i] = (sum of sqrt of the k smallest d^2, ascending) / (number found); -1 for an empty result (open3d's marker)
__global__ void __launch_bounds__(kFiltThreads)
knn_mean_distance_kernel(const double* __restrict__ points, const int32_t* __restrict__ offset, int k, double* __restrict__ avg)
{
    __shared__ double s_p[kFiltTile][3];
    const int c = blockIdx.y;
    const int p0 = offset[c], n = offset[c + 1] - p0;
    if ((int)(blockIdx.x * kFiltThreads) >= n) return;
    const double* P = points + 3 * (size_t)p0;
    const int i = blockIdx.x * kFiltThreads + threadIdx.x;
    const bool live = i < n;
    const int ii = live ? i : n - 1;
    const double qx = P[3 * ii], qy = P[3 * ii + 1], qz = P[3 * ii + 2];
    double best[kFiltMaxK];                       // ascending; local memory (dynamic indexing), L1 resident
    int have = 0;
    for (int t0 = 0; t0 < n; t0 += kFiltTile) {
        const int m = min(kFiltTile, n - t0);
        __syncthreads();
        for (int q = threadIdx.x; q < 3 * m; q += kFiltThreads) (&s_p[0][0])[q] = P[3 * (size_t)t0 + q];
        __syncthreads();
        for (int j = 0; j < m; ++j) {
            const double d = dist2(s_p[j][0], s_p[j][1], s_p[j][2], qx, qy, qz);
            if (have == k && !(d < best[k - 1])) continue;
            int pos = have < k ? have : k - 1;
            while (pos > 0 && d < best[pos - 1]) { best[pos] = best[pos - 1]; --pos; }
            best[pos] = d;
            if (have < k) ++have;
        }
    }
    if (live) {
        double s = 0.0;
        for (int j = 0; j < have; ++j) s = __dadd_rn(s, sqrt(best[j]));
        avg[p0 + i] = have > 0 ? __ddiv_rn(s, (double)have) : -1.0;
    }
}

// Block-wide fp64 sum, fixed tree (reproducible).  Every thread gets the total.
__device__ double block_sum(double v, double* s_part /* [threads/32] */) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    return t;
}

// One CTA per cloud: threshold = mean + ratio * std over the valid (> 0) averages; keep = 0 < avg < threshold.
// ratio: per-cloud value from `ratio_dev` (e.g. the Mahalanobis std the reference feeds in) or the scalar.
__global__ void __launch_bounds__(1024)
statistical_threshold_kernel(const double* __restrict__ avg, const int32_t* __restrict__ offset, const double* __restrict__ ratio_dev,
                             double ratio_scalar, uint8_t* __restrict__ keep, double* __restrict__ threshold_out)
{
    __shared__ double s_part[32];
    const int c = blockIdx.x;
    const int p0 = offset[c], n = offset[c + 1] - p0;
    const double* A = avg + p0;
    double s = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double a = A[i]; if (a > 0.0) { s += a; cnt += 1.0; } }
    s = block_sum(s, s_part); cnt = block_sum(cnt, s_part);
    const double mean = s / cnt;
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double a = A[i]; if (a > 0.0) q += (a - mean) * (a - mean); }
    q = block_sum(q, s_part);
    const double sd = sqrt(q / (cnt - 1.0));
    const double thr = mean + (ratio_dev ? ratio_dev[c] : ratio_scalar) * sd;
    for (int i = threadIdx.x; i < n; i += blockDim.x) keep[p0 + i] = (A[i] > 0.0 && A[i] < thr) ? 1 : 0;
    if (threadIdx.x == 0 && threshold_out) threshold_out[c] = thr;
}

// One CTA per cloud: cumulants -> mean, population covariance -> analytic 3x3 inverse -> per-point distance (optional) and
// the population std of |distance| (numpy's np.std, which is what open3d_utils.py:200-201 / :207-208 computes from it).
__global__ void __launch_bounds__(1024)
mahalanobis_kernel(const double* __restrict__ points, const int32_t* __restrict__ offset, double* __restrict__ dist,
                   double* __restrict__ std_out)
{
    __shared__ double s_part[32];
    const int c = blockIdx.x;
    const int p0 = offset[c], n = offset[c + 1] - p0;
    if (n <= 0) { if (threadIdx.x == 0 && std_out) std_out[c] = 0.0; return; }
    const double* P = points + 3 * (size_t)p0;
    double cu[9] = {};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
        cu[0] += x; cu[1] += y; cu[2] += z; cu[3] += x * x; cu[4] += x * y; cu[5] += x * z; cu[6] += y * y; cu[7] += y * z; cu[8] += z * z;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) cu[k] = block_sum(cu[k], s_part) / (double)n;
    const double a = cu[3] - cu[0] * cu[0], b = cu[4] - cu[0] * cu[1], cc = cu[5] - cu[0] * cu[2];
    const double d = cu[6] - cu[1] * cu[1], e = cu[7] - cu[1] * cu[2], f = cu[8] - cu[2] * cu[2];
    // inverse of the symmetric [[a,b,cc],[b,d,e],[cc,e,f]] by cofactors
    const double c00 = d * f - e * e, c01 = cc * e - b * f, c02 = b * e - cc * d;
    const double c11 = a * f - cc * cc, c12 = b * cc - a * e, c22 = a * d - b * b;
    const double det = a * c00 + b * c01 + cc * c02;
    const double i00 = c00 / det, i01 = c01 / det, i02 = c02 / det, i11 = c11 / det, i12 = c12 / det, i22 = c22 / det;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = P[3 * i] - cu[0], y = P[3 * i + 1] - cu[1], z = P[3 * i + 2] - cu[2];
        const double m = sqrt(x * (i00 * x + i01 * y + i02 * z) + y * (i01 * x + i11 * y + i12 * z) + z * (i02 * x + i12 * y + i22 * z));
        if (dist) dist[p0 + i] = m;
        s += fabs(m);
    }
    s = block_sum(s, s_part);
    const double mu = s / (double)n;
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = P[3 * i] - cu[0], y = P[3 * i + 1] - cu[1], z = P[3 * i + 2] - cu[2];
        const double m = fabs(sqrt(x * (i00 * x + i01 * y + i02 * z) + y * (i01 * x + i11 * y + i12 * z) + z * (i02 * x + i12 * y + i22 * z)));
        q += (m - mu) * (m - mu);
    }
    q = block_sum(q, s_part);
    if (threadIdx.x == 0 && std_out) std_out[c] = sqrt(q / (double)n);
}

// Ordered per-cloud compaction of the kept points: out cloud c occupies out[offset[c] : offset[c] + counts[c]].
__global__ void __launch_bounds__(1024)
compact_points_kernel(const double* __restrict__ points, const int32_t* __restrict__ offset, const uint8_t* __restrict__ keep,
                      double* __restrict__ out, int32_t* __restrict__ counts, int32_t* __restrict__ index)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p0 = offset[c], n = offset[c + 1] - p0;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int k = (i < n && keep[p0 + i]) ? 1 : 0;
        int incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (k) {
            const int slot = p0 + before + incl - 1;
            out[3 * (size_t)slot] = points[3 * (size_t)(p0 + i)];
            out[3 * (size_t)slot + 1] = points[3 * (size_t)(p0 + i) + 1];
            out[3 * (size_t)slot + 2] = points[3 * (size_t)(p0 + i) + 2];
            if (index) index[slot] = i;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[c] = s_carry;
}

}  // namespace ape

#define APE_API extern "C" __attribute__((visibility("default")))

APE_API int ape_radius_outlier(const double* points, const int32_t* offset, int n_clouds, int max_cloud_points, int nb_points,
                               double radius, uint8_t* keep, int32_t* n_neighbors, void* stream)
{
    APE_REQUIRE(points && offset && keep, "ape_radius_outlier: null pointer");
    APE_REQUIRE(n_clouds >= 0 && max_cloud_points >= 0 && radius > 0.0, "ape_radius_outlier: bad sizes");
    if (n_clouds == 0 || max_cloud_points == 0) return APE_OK;
    APE_REQUIRE(n_clouds <= 65535, "ape_radius_outlier: more than 65535 clouds per call (split the batch)");
    ape::ProfScope prof_("radius_outlier", (cudaStream_t)stream);
    dim3 grid((max_cloud_points + ape::kFiltThreads - 1) / ape::kFiltThreads, n_clouds);
    ape::radius_outlier_kernel<<<grid, ape::kFiltThreads, 0, (cudaStream_t)stream>>>(points, offset, nb_points, radius * radius, keep,
                                                                                   n_neighbors);
    ape::count_launch();
    return ape::check_launch("ape_radius_outlier");
}

APE_API int ape_mahalanobis(const double* points, const int32_t* offset, int n_clouds, double* dist, double* std_out, void* stream)
{
    APE_REQUIRE(points && offset && (dist || std_out), "ape_mahalanobis: null pointer");
    APE_REQUIRE(n_clouds >= 0, "ape_mahalanobis: bad sizes");
    if (n_clouds == 0) return APE_OK;
    ape::ProfScope prof_("mahalanobis", (cudaStream_t)stream);
    ape::mahalanobis_kernel<<<n_clouds, 1024, 0, (cudaStream_t)stream>>>(points, offset, dist, std_out);
    ape::count_launch();
    return ape::check_launch("ape_mahalanobis");
}

APE_API int ape_statistical_outlier(const double* points, const int32_t* offset, int n_clouds, int max_cloud_points, int nb_neighbors,
                                    const double* std_ratio_dev, double std_ratio, uint8_t* keep, double* avg_dist /* [P] scratch+out */,
                                    double* threshold, void* stream)
{
    APE_REQUIRE(points && offset && keep && avg_dist, "ape_statistical_outlier: null pointer");
    APE_REQUIRE(n_clouds >= 0 && max_cloud_points >= 0, "ape_statistical_outlier: bad sizes");
    APE_REQUIRE(nb_neighbors >= 1 && nb_neighbors <= ape::kFiltMaxK, "ape_statistical_outlier: nb_neighbors must be in [1, %d]", ape::kFiltMaxK);
    if (n_clouds == 0 || max_cloud_points == 0) return APE_OK;
    APE_REQUIRE(n_clouds <= 65535, "ape_statistical_outlier: more than 65535 clouds per call (split the batch)");
    cudaStream_t s = (cudaStream_t)stream;
    ape::ProfScope prof_("statistical_outlier", s);
    dim3 grid((max_cloud_points + ape::kFiltThreads - 1) / ape::kFiltThreads, n_clouds);
    ape::knn_mean_distance_kernel<<<grid, ape::kFiltThreads, 0, s>>>(points, offset, nb_neighbors, avg_dist);
    ape::statistical_threshold_kernel<<<n_clouds, 1024, 0, s>>>(avg_dist, offset, std_ratio_dev, std_ratio, keep, threshold);
    ape::count_launch(2);
    return ape::check_launch("ape_statistical_outlier");
}

APE_API int ape_compact_points(const double* points, const int32_t* offset, const uint8_t* keep, int n_clouds, double* out_points,
                               int32_t* out_counts, int32_t* out_index, void* stream)
{
    APE_REQUIRE(points && offset && keep && out_points && out_counts, "ape_compact_points: null pointer");
    APE_REQUIRE(n_clouds >= 0, "ape_compact_points: bad sizes");
    APE_REQUIRE(out_points != points, "ape_compact_points: output must not alias the input");
    if (n_clouds == 0) return APE_OK;
    ape::ProfScope prof_("compact_points", (cudaStream_t)stream);
    ape::compact_points_kernel<<<n_clouds, 1024, 0, (cudaStream_t)stream>>>(points, offset, keep, out_points, out_counts, out_index);
    ape::count_launch();
    return ape::check_launch("ape_compact_points");
}
