// Post-network pose math on the device (sm_100a): confidence arg-max, pose of the winning point,
// cloud re-expressed in the predicted frame, and the fp64 pose composition of a refinement step.
// Replaces DenseFusion/tools/utils.py:7-18 (my_estimator_prediction), :43-86 (get_new_points),
// :20-40 (my_refined_prediction -> transformations.py:1254-1278, :1320-1341, :1361-1363) and the
// canonical loop body of tools/eval_linemod.py:92-110, removing the per-object D2H/H2D round trips.
#include "ape_common.cuh"
#include <cfloat>

namespace ape {

// quaternion_matrix (transformations.py:1266-1278), 3x3 part, fp64.  q need not be unit length.
__device__ __forceinline__ void quaternion_matrix3(const double* q, double* M /*9, row-major*/) {
    const double n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n < 2.220446049250313e-16 * 4.0) {          // _EPS (transformations.py:1669)
        M[0] = 1; M[1] = 0; M[2] = 0; M[3] = 0; M[4] = 1; M[5] = 0; M[6] = 0; M[7] = 0; M[8] = 1;
        return;
    }
    const double s = sqrt(2.0 / n);
    const double w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;   // q *= sqrt(2/n); outer(q,q)
    M[0] = 1.0 - y * y - z * z; M[1] = x * y - z * w;       M[2] = x * z + y * w;
    M[3] = x * y + z * w;       M[4] = 1.0 - x * x - z * z; M[5] = y * z - x * w;
    M[6] = x * z - y * w;       M[7] = y * z + x * w;       M[8] = 1.0 - x * x - y * y;
}

// quaternion_from_matrix(M, isprecise=True) for a 4x4 whose last row/col is (0,0,0,1)
// (transformations.py:1321-1341, sign rule :1361-1362).  R row-major 3x3.
__device__ __forceinline__ void quaternion_from_matrix_precise(const double* R, double* q) {
    const double m33 = 1.0;
    double t = R[0] + R[4] + R[8] + m33;
    if (t > m33) {
        q[0] = t; q[3] = R[3] - R[1]; q[2] = R[2] - R[6]; q[1] = R[7] - R[5];
    } else {
        int i = 0, j = 1, k = 2;
        if (R[4] > R[0]) { i = 1; j = 2; k = 0; }
        if (R[8] > R[3 * i + i]) { i = 2; j = 0; k = 1; }
        t = R[3 * i + i] - (R[3 * j + j] + R[3 * k + k]) + m33;
        double v[4];
        v[i] = t;
        v[j] = R[3 * i + j] + R[3 * j + i];
        v[k] = R[3 * k + i] + R[3 * i + k];
        v[3] = R[3 * k + j] - R[3 * j + k];
        q[0] = v[3]; q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
    }
    const double s = 0.5 / sqrt(t * m33);
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
    if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
}

// ------------------------------------------------------------------------------ a8/a9
// One CTA per object: arg-max of the confidences (lowest index on ties), the winning point's
// normalised quaternion and translation, then new_points = (cloud - t) @ base.
constexpr int kSelThreads = 256;
__global__ void __launch_bounds__(kSelThreads)
pose_select_kernel(const float* __restrict__ pred_r, const float* __restrict__ pred_t, const float* __restrict__ pred_c,
                   const float* __restrict__ cloud, int N, int32_t* __restrict__ which_max, float* __restrict__ my_r,
                   float* __restrict__ my_t, float* __restrict__ new_points, double* __restrict__ pose)
{
    pdl_sync();                                           // programmatic dependent launch: see ape_common.cuh
    __shared__ float s_v[kSelThreads / 32];
    __shared__ int s_i[kSelThreads / 32];
    __shared__ float s_pose[16];
    const int b = blockIdx.x;
    const float* c = pred_c + (size_t)b * N;
    float bv = -FLT_MAX; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += kSelThreads) {
        const float v = c[i];
        if (v > bv) { bv = v; bi = i; }                 // ascending i per thread: first max kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kSelThreads / 32; ++w)
            if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
        if (bi == 0x7fffffff) bi = 0;
        const float* r = pred_r + ((size_t)b * N + bi) * 4;
        // pred_r / torch.norm(pred_r, dim=2)  (tools/utils.py:8, :45)
        const float nrm = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        const float w_ = r[0] / nrm, x = r[1] / nrm, y = r[2] / nrm, z = r[3] / nrm;
        const float* t = pred_t + ((size_t)b * N + bi) * 3;
        const float* p = cloud + ((size_t)b * N + bi) * 3;
        const float tx = p[0] + t[0], ty = p[1] + t[1], tz = p[2] + t[2];   // points + pred_t (:16, :78)
        quat_to_base(w_, x, y, z, s_pose);
        s_pose[9] = tx; s_pose[10] = ty; s_pose[11] = tz;
        which_max[b] = bi;
        my_r[4 * b] = w_; my_r[4 * b + 1] = x; my_r[4 * b + 2] = y; my_r[4 * b + 3] = z;
        my_t[3 * b] = tx; my_t[3 * b + 1] = ty; my_t[3 * b + 2] = tz;
        if (pose) {
            double* o = pose + 7 * (size_t)b;
            o[0] = w_; o[1] = x; o[2] = y; o[3] = z; o[4] = tx; o[5] = ty; o[6] = tz;
        }
    }
    __syncthreads();
    if (new_points) {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = s_pose[i];
        const float tx = s_pose[9], ty = s_pose[10], tz = s_pose[11];
        for (int i = threadIdx.x; i < N; i += kSelThreads) {
            const float* p = cloud + ((size_t)b * N + i) * 3;
            const float dx = p[0] - tx, dy = p[1] - ty, dz = p[2] - tz;
            float* o = new_points + ((size_t)b * N + i) * 3;
            // (points - t) @ base : out_j = sum_k d_k * R[k][j]  (tools/utils.py:83)
            o[0] = dx * R[0] + dy * R[3] + dz * R[6];
            o[1] = dx * R[1] + dy * R[4] + dz * R[7];
            o[2] = dx * R[2] + dy * R[5] + dz * R[8];
        }
    }
}

// ------------------------------------------------------------------------------ a11
// One CTA per object.  Thread 0 composes the pose in fp64, the block then (optionally) maps the
// cloud into the composed frame for the next refinement iteration.
__global__ void __launch_bounds__(kSelThreads)
pose_compose_kernel(const double* __restrict__ pose_in, const float* __restrict__ r2, const float* __restrict__ t2,
                    double* __restrict__ pose_out, const float* __restrict__ cloud, int N,
                    float* __restrict__ next_points)
{
    pdl_sync();                                           // programmatic dependent launch: see ape_common.cuh
    __shared__ float s_R[9];
    __shared__ float s_T[3];
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        const double* pi = pose_in + 7 * (size_t)b;
        double M1[9], M2[9], Mf[9], q[4], t1[3], tt[3];
        quaternion_matrix3(pi, M1);
        t1[0] = pi[4]; t1[1] = pi[5]; t1[2] = pi[6];
        // pred_r / torch.norm(pred_r) in fp32 (tools/utils.py:24, eval_linemod.py:100), then widened
        const float a = r2[4 * b], c = r2[4 * b + 1], d = r2[4 * b + 2], e = r2[4 * b + 3];
        const float nrm = sqrtf(a * a + c * c + d * d + e * e);
        const double q2[4] = {(double)(a / nrm), (double)(c / nrm), (double)(d / nrm), (double)(e / nrm)};
        quaternion_matrix3(q2, M2);
        const double t2d[3] = {(double)t2[3 * b], (double)t2[3 * b + 1], (double)t2[3 * b + 2]};
        // my_mat_final = my_mat @ my_mat_2  (tools/utils.py:31)
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j)
                Mf[3 * i + j] = M1[3 * i] * M2[j] + M1[3 * i + 1] * M2[3 + j] + M1[3 * i + 2] * M2[6 + j];
            tt[i] = M1[3 * i] * t2d[0] + M1[3 * i + 1] * t2d[1] + M1[3 * i + 2] * t2d[2] + t1[i];
        }
        quaternion_from_matrix_precise(Mf, q);
        double* o = pose_out + 7 * (size_t)b;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; o[4] = tt[0]; o[5] = tt[1]; o[6] = tt[2];
        if (next_points) {
            // eval_linemod.py:92-94 of the NEXT iteration: R = quaternion_matrix(my_r)[:3,:3].astype(fp32),
            // T = my_t.astype(fp32)
            double Mn[9];
            quaternion_matrix3(q, Mn);
            for (int i = 0; i < 9; ++i) s_R[i] = (float)Mn[i];
            for (int i = 0; i < 3; ++i) s_T[i] = (float)tt[i];
        }
    }
    if (!next_points) return;
    __syncthreads();
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = s_R[i];
    const float tx = s_T[0], ty = s_T[1], tz = s_T[2];
    for (int i = threadIdx.x; i < N; i += kSelThreads) {
        const float* p = cloud + ((size_t)b * N + i) * 3;
        const float dx = p[0] - tx, dy = p[1] - ty, dz = p[2] - tz;
        float* o = next_points + ((size_t)b * N + i) * 3;
        o[0] = dx * R[0] + dy * R[3] + dz * R[6];      // bmm(points - T, R)  (eval_linemod.py:97)
        o[1] = dx * R[1] + dy * R[4] + dz * R[7];
        o[2] = dx * R[2] + dy * R[5] + dz * R[8];
    }
}

}  // namespace ape

extern "C" __attribute__((visibility("default"))) int ape_pose_select(const float* pred_r, const float* pred_t, const float* pred_c, const float* cloud,
                               int B, int N, int32_t* which_max, float* my_r, float* my_t, float* new_points,
                               double* pose, void* stream)
{
    APE_REQUIRE(pred_r && pred_t && pred_c && cloud && which_max && my_r && my_t, "ape_pose_select: null pointer");
    APE_REQUIRE(B >= 0 && N > 0, "ape_pose_select: bad sizes");
    if (B == 0) return APE_OK;
    ape::ProfScope prof_("pose_select", (cudaStream_t)stream);
    APE_CUDA(ape::launch_pdl(ape::pose_select_kernel, dim3(B), dim3(ape::kSelThreads), 0, (cudaStream_t)stream, pred_r, pred_t, pred_c, cloud, N,
                             which_max, my_r, my_t, new_points, pose));
    ape::count_launch();
    return ape::check_launch("ape_pose_select");
}

extern "C" __attribute__((visibility("default"))) int ape_pose_compose(const double* pose_in, const float* r2, const float* t2, int B, double* pose_out,
                                const float* cloud, int N, float* next_points, void* stream)
{
    APE_REQUIRE(pose_in && r2 && t2 && pose_out, "ape_pose_compose: null pointer");
    APE_REQUIRE(B >= 0, "ape_pose_compose: bad sizes");
    APE_REQUIRE(!next_points || (cloud && N > 0), "ape_pose_compose: next_points needs cloud and N");
    if (B == 0) return APE_OK;
    ape::ProfScope prof_("pose_compose", (cudaStream_t)stream);
    APE_CUDA(ape::launch_pdl(ape::pose_compose_kernel, dim3(B), dim3(ape::kSelThreads), 0, (cudaStream_t)stream, pose_in, r2, t2, pose_out, cloud, N,
                             next_points));
    ape::count_launch();
    return ape::check_launch("ape_pose_compose");
}
