// Brute-force nearest neighbours and the fused ADD / ADD-S metric (sm_100a).
//   ape_knn          replaces knn() DenseFusion/lib/knn/src/knn.h:12 (kernels knn.cu:36, :113)
//   ape_add_metric   replaces loss_refiner.py:39-49 / eval_linemod.py:118-130
// The reference writes and re-reads an N x M fp32 distance matrix (10.4 MB per 500x2600
// instance); here the reference points of one instance live in shared memory, each thread
// keeps its queries in registers and the matrix is never materialised (compulsory traffic
// only: ~41 KB per instance).  Arithmetic is selectable so indices are bit-exact with either
// reference implementation (APE_KNN_ARITH_CPU / APE_KNN_ARITH_FMA).
#include "ape_common.cuh"
#include <cfloat>

namespace ape {

template <bool FMA>
__device__ __forceinline__ float dist3(float rx, float ry, float rz, float qx, float qy, float qz) {
    const float dx = __fsub_rn(rx, qx), dy = __fsub_rn(ry, qy), dz = __fsub_rn(rz, qz);
    if (FMA) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

constexpr int kKnnThreads = 256;
constexpr int kKnnQ = 2;            // queries per thread (register tile)
constexpr int kKnnRefTile = 2048;   // reference points staged per pass (float4 each: 32 KB)

// Scan `n` staged reference points for kKnnQ queries.  Points are visited in ascending index and
// replaced only on strict '<', so the lowest index wins exact ties (knn.cu:156-173, knn_cpu.cpp:30).
// Groups of 4 references are reduced with min first; the index is resolved only when the group
// beats the running best, which keeps the common path at one FMNMX per pair.
template <bool FMA>
__device__ __forceinline__ void scan_tile(const float4* __restrict__ s_ref, int n, int index0,
                                          const float (&qx)[kKnnQ], const float (&qy)[kKnnQ], const float (&qz)[kKnnQ],
                                          float (&best)[kKnnQ], int (&bidx)[kKnnQ])
{
    int j = 0;
    for (; j + 4 <= n; j += 4) {
        const float4 r0 = s_ref[j], r1 = s_ref[j + 1], r2 = s_ref[j + 2], r3 = s_ref[j + 3];
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            const float d0 = dist3<FMA>(r0.x, r0.y, r0.z, qx[q], qy[q], qz[q]);
            const float d1 = dist3<FMA>(r1.x, r1.y, r1.z, qx[q], qy[q], qz[q]);
            const float d2 = dist3<FMA>(r2.x, r2.y, r2.z, qx[q], qy[q], qz[q]);
            const float d3 = dist3<FMA>(r3.x, r3.y, r3.z, qx[q], qy[q], qz[q]);
            const float m = fminf(fminf(d0, d1), fminf(d2, d3));
            if (m < best[q]) {
                best[q] = m;
                bidx[q] = index0 + j + (d0 == m ? 0 : (d1 == m ? 1 : (d2 == m ? 2 : 3)));
            }
        }
    }
    for (; j < n; ++j) {
        const float4 r = s_ref[j];
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            const float d = dist3<FMA>(r.x, r.y, r.z, qx[q], qy[q], qz[q]);
            if (d < best[q]) { best[q] = d; bidx[q] = index0 + j; }
        }
    }
}

// ref [B,3,N], query [B,3,M] (SoA rows as the reference lays them out), idx [B,1,M] 1-based.
// grid = (ceil(M / (threads*Q)), B)
template <bool FMA>
__global__ void __launch_bounds__(kKnnThreads)
knn3_top1_kernel(const float* __restrict__ ref, const float* __restrict__ query, int64_t* __restrict__ idx,
                 int N, int M)
{
    __shared__ float4 s_ref[kKnnRefTile];
    const int b = blockIdx.y;
    const float* R = ref + (size_t)b * 3 * N;
    const float* Q = query + (size_t)b * 3 * M;
    const int q0 = (blockIdx.x * kKnnThreads + threadIdx.x) * kKnnQ;
    float qx[kKnnQ], qy[kKnnQ], qz[kKnnQ], best[kKnnQ];
    int bidx[kKnnQ];
#pragma unroll
    for (int q = 0; q < kKnnQ; ++q) {
        const int m = min(q0 + q, M - 1);
        qx[q] = Q[m]; qy[q] = Q[M + m]; qz[q] = Q[2 * M + m];
        best[q] = FLT_MAX; bidx[q] = 0;
    }
    for (int t0 = 0; t0 < N; t0 += kKnnRefTile) {
        const int n = min(kKnnRefTile, N - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kKnnThreads)
            s_ref[i] = make_float4(R[t0 + i], R[N + t0 + i], R[2 * N + t0 + i], 0.f);
        __syncthreads();
        scan_tile<FMA>(s_ref, n, t0, qx, qy, qz, best, bidx);
    }
#pragma unroll
    for (int q = 0; q < kKnnQ; ++q)
        if (q0 + q < M) idx[(size_t)b * M + q0 + q] = (int64_t)bidx[q] + 1;
}

// Generic shape: any D, 1 <= k <= kKnnMaxK.  One thread per query, sorted k-list per thread.
constexpr int kKnnMaxK = 64;
template <bool FMA>
__global__ void __launch_bounds__(128)
knn_generic_kernel(const float* __restrict__ ref, const float* __restrict__ query, int64_t* __restrict__ idx,
                   int D, int N, int M, int k)
{
    const int b = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* R = ref + (size_t)b * D * N;
    const float* Q = query + (size_t)b * D * M;
    float bd[kKnnMaxK];
    int bi[kKnnMaxK];
    int have = 0;
    for (int r = 0; r < N; ++r) {
        float acc = 0.f;
        for (int d = 0; d < D; ++d) {
            const float diff = __fsub_rn(R[(size_t)d * N + r], Q[(size_t)d * M + m]);
            acc = FMA ? __fmaf_rn(diff, diff, acc) : __fadd_rn(acc, __fmul_rn(diff, diff));
        }
        if (have == k && !(acc < bd[k - 1])) continue;
        int pos = have < k ? have : k - 1;
        while (pos > 0 && acc < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = acc; bi[pos] = r + 1;
        if (have < k) ++have;
    }
    for (int j = 0; j < k; ++j) idx[((size_t)b * k + j) * M + m] = bi[j];
}

// ------------------------------------------------------------------------------ ADD / ADD-S
// One CTA per instance.  pred_i = model_i * R^T + t (row-major base of the normalised quaternion),
// ADD-S: nearest target point per pred point (CPU arithmetic: the query is pred, the reference is
// target, as knn(target, pred) in loss_refiner.py:44), distance = sqrt of the winning d2.
// ADD: ||pred_i - target_i||.  dis = mean over model points (fp32, fixed-order block reduction).
constexpr int kAddThreads = 256;
__global__ void __launch_bounds__(kAddThreads)
add_metric_kernel(const float* __restrict__ quat, const float* __restrict__ trans,
                  const float* __restrict__ model, int64_t model_stride, int n_model,
                  const float* __restrict__ target, int64_t target_stride, int n_target,
                  const uint8_t* __restrict__ symmetric, float* __restrict__ dis, int32_t* __restrict__ nn_index,
                  float* __restrict__ std_out /* unbiased std of the per-point distances (loss.py:50), or NULL */)
{
    __shared__ float4 s_ref[kKnnRefTile];
    __shared__ float s_part[kAddThreads / 32];
    __shared__ float s_mean;
    __shared__ float s_d[kKnnRefTile];                   // per-point distances (std_out only)
    const int b = blockIdx.x;
    const float* Mp = model + (size_t)b * model_stride;
    const float* Tg = target + (size_t)b * target_stride;
    float w = quat[4 * b], x = quat[4 * b + 1], y = quat[4 * b + 2], z = quat[4 * b + 3];
    const float nrm = sqrtf(w * w + x * x + y * y + z * z);
    w /= nrm; x /= nrm; y /= nrm; z /= nrm;
    float R[9];
    quat_to_base(w, x, y, z, R);
    const float tx = trans[3 * b], ty = trans[3 * b + 1], tz = trans[3 * b + 2];
    const bool sym = symmetric && symmetric[b];

    float acc = 0.f;
    for (int p0 = 0; p0 < n_model; p0 += kAddThreads * kKnnQ) {       // usually one pass (500 <= 512)
        const int q0 = p0 + threadIdx.x * kKnnQ;
        float qx[kKnnQ], qy[kKnnQ], qz[kKnnQ], best[kKnnQ];
        int bidx[kKnnQ];
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            const int i = min(q0 + q, n_model - 1);
            const float mx = Mp[3 * i], my = Mp[3 * i + 1], mz = Mp[3 * i + 2];
            // model @ base^T + t  (loss_refiner.py:31, :39): pred_a = sum_k m_k * R[a][k] + t_a
            qx[q] = (mx * R[0] + my * R[1] + mz * R[2]) + tx;
            qy[q] = (mx * R[3] + my * R[4] + mz * R[5]) + ty;
            qz[q] = (mx * R[6] + my * R[7] + mz * R[8]) + tz;
            best[q] = FLT_MAX; bidx[q] = 0;
        }
        if (sym) {
            for (int t0 = 0; t0 < n_target; t0 += kKnnRefTile) {
                const int n = min(kKnnRefTile, n_target - t0);
                __syncthreads();
                for (int i = threadIdx.x; i < n; i += kAddThreads)
                    s_ref[i] = make_float4(Tg[3 * (t0 + i)], Tg[3 * (t0 + i) + 1], Tg[3 * (t0 + i) + 2], 0.f);
                __syncthreads();
                scan_tile<false>(s_ref, n, t0, qx, qy, qz, best, bidx);
            }
        }
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            if (q0 + q < n_model) {
                float d;
                if (sym) {
                    d = sqrtf(best[q]);
                    if (nn_index) nn_index[(size_t)b * n_model + q0 + q] = bidx[q];
                } else {
                    // ADD pairs model point i with target row i, so an instance with flag 0 needs n_target >= n_model (the first
                    // n_model target rows are the partners); if the caller got that wrong nothing is read out of bounds and
                    // the instance's distance comes out as NaN
                    const int i = q0 + q, it = min(i, n_target - 1);
                    const float dx = qx[q] - Tg[3 * it], dy = qy[q] - Tg[3 * it + 1], dz = qz[q] - Tg[3 * it + 2];
                    d = i < n_target ? sqrtf(dx * dx + dy * dy + dz * dz) : nanf("");
                    if (nn_index) nn_index[(size_t)b * n_model + i] = it;
                }
                acc += d;
                if (std_out) s_d[q0 + q] = d;                  // n_model <= kKnnRefTile (checked by the caller)
            }
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kAddThreads / 32; ++i) s += s_part[i];
        s_mean = s / (float)n_model;
        dis[b] = s_mean;
    }
    if (std_out) {
        // torch.std: two passes, Bessel's correction
        __syncthreads();
        const float mean = s_mean;
        float q = 0.f;
        for (int i = threadIdx.x; i < n_model; i += kAddThreads) {
            const float d = s_d[i] - mean;
            q += d * d;
        }
        q = warp_sum(q);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = q;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kAddThreads / 32; ++i) s += s_part[i];
            std_out[b] = sqrtf(s / (float)(n_model - 1));
        }
    }
}


// Gradient w.r.t. the RAW quaternion r (q = r / |r|, Rm = base(q)) from D = d loss / d Rm (row-major like Rm):
// through quat_to_base (tools/utils.py:46-67) and through the normalisation.
__device__ __forceinline__ void base_backward(float w, float x, float y, float z, float nrm, const float* D, float* d_r) {
    const float gw = -2.f * z * D[1] + 2.f * y * D[2] + 2.f * z * D[3] - 2.f * x * D[5] - 2.f * y * D[6] + 2.f * x * D[7];
    const float gx = 2.f * y * D[1] + 2.f * z * D[2] + 2.f * y * D[3] - 4.f * x * D[4] - 2.f * w * D[5] + 2.f * z * D[6] +
                     2.f * w * D[7] - 4.f * x * D[8];
    const float gy = -4.f * y * D[0] + 2.f * x * D[1] + 2.f * w * D[2] + 2.f * x * D[3] + 2.f * z * D[5] - 2.f * w * D[6] +
                     2.f * z * D[7] - 4.f * y * D[8];
    const float gz = -4.f * z * D[0] - 2.f * w * D[1] + 2.f * x * D[2] + 2.f * w * D[3] - 4.f * z * D[4] + 2.f * y * D[5] +
                     2.f * x * D[6] + 2.f * y * D[7];
    const float dot = gw * w + gx * x + gy * y + gz * z;
    d_r[0] = (gw - w * dot) / nrm; d_r[1] = (gx - x * dot) / nrm;
    d_r[2] = (gy - y * dot) / nrm; d_r[3] = (gz - z * dot) / nrm;
}

// ------------------------------------------------------------------------------ refiner loss, forward + backward
// Loss_refine (DenseFusion/lib/loss_refiner.py:12-64) for B objects with its gradient, one CTA per object:
//   q = r / |r|, Rm = base(q), pred_j = Rm m_j + t (:39); symmetric: tgt_j = nearest target to pred_j (:41-47, the
//   index is a constant of the backward pass exactly as the reference detaches it); dis = mean_j |pred_j - tgt_j| (:49)
//   d dis / d pred_j = (pred_j - tgt_j) / (M |.|);  d_t = sum_j,  dRm = sum_j g_j m_j^T,  d_r through base() and
//   the normalisation.  Also new_points = (points - t) Rm and new_target = (target - t) Rm (:51-60).
constexpr int kLossThreads = 256;
__global__ void __launch_bounds__(kLossThreads)
refine_loss_kernel(const float* __restrict__ quat, const float* __restrict__ trans, const float* __restrict__ model,
                   const float* __restrict__ target, int n_mesh, const float* __restrict__ points, int n_points,
                   const uint8_t* __restrict__ symmetric, float* __restrict__ dis, float* __restrict__ d_r,
                   float* __restrict__ d_t, float* __restrict__ new_points, float* __restrict__ new_target)
{
    __shared__ float4 s_ref[kKnnRefTile];
    __shared__ float s_part[kLossThreads / 32][13];
    const int b = blockIdx.x;
    const float* Mp = model + (size_t)b * n_mesh * 3;
    const float* Tg = target + (size_t)b * n_mesh * 3;
    const float r0 = quat[4 * b], r1 = quat[4 * b + 1], r2 = quat[4 * b + 2], r3 = quat[4 * b + 3];
    const float nrm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
    const float w = r0 / nrm, x = r1 / nrm, y = r2 / nrm, z = r3 / nrm;
    float R[9];
    quat_to_base(w, x, y, z, R);
    const float tx = trans[3 * b], ty = trans[3 * b + 1], tz = trans[3 * b + 2];
    const bool sym = symmetric && symmetric[b];
    float acc[13] = {};                                    // dis, d_t[3], dRm[9]
    for (int p0 = 0; p0 < n_mesh; p0 += kLossThreads * kKnnQ) {
        const int q0 = p0 + threadIdx.x * kKnnQ;
        float qx[kKnnQ], qy[kKnnQ], qz[kKnnQ], best[kKnnQ], mx[kKnnQ], my[kKnnQ], mz[kKnnQ];
        int bidx[kKnnQ];
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            const int i = min(q0 + q, n_mesh - 1);
            mx[q] = Mp[3 * i]; my[q] = Mp[3 * i + 1]; mz[q] = Mp[3 * i + 2];
            qx[q] = (mx[q] * R[0] + my[q] * R[1] + mz[q] * R[2]) + tx;
            qy[q] = (mx[q] * R[3] + my[q] * R[4] + mz[q] * R[5]) + ty;
            qz[q] = (mx[q] * R[6] + my[q] * R[7] + mz[q] * R[8]) + tz;
            best[q] = FLT_MAX; bidx[q] = q0 + q < n_mesh ? q0 + q : 0;
        }
        if (sym) {
            for (int t0 = 0; t0 < n_mesh; t0 += kKnnRefTile) {
                const int n = min(kKnnRefTile, n_mesh - t0);
                __syncthreads();
                for (int i = threadIdx.x; i < n; i += kLossThreads)
                    s_ref[i] = make_float4(Tg[3 * (t0 + i)], Tg[3 * (t0 + i) + 1], Tg[3 * (t0 + i) + 2], 0.f);
                __syncthreads();
                scan_tile<false>(s_ref, n, t0, qx, qy, qz, best, bidx);
            }
        }
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            if (q0 + q < n_mesh) {
                const int i = bidx[q];
                const float dx = qx[q] - Tg[3 * i], dy = qy[q] - Tg[3 * i + 1], dz = qz[q] - Tg[3 * i + 2];
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                acc[0] += d;
                const float inv = d > 0.f ? 1.0f / (d * (float)n_mesh) : 0.f;
                const float gx = dx * inv, gy = dy * inv, gz = dz * inv;
                acc[1] += gx; acc[2] += gy; acc[3] += gz;
                acc[4] += gx * mx[q]; acc[5] += gx * my[q]; acc[6] += gx * mz[q];
                acc[7] += gy * mx[q]; acc[8] += gy * my[q]; acc[9] += gy * mz[q];
                acc[10] += gz * mx[q]; acc[11] += gz * my[q]; acc[12] += gz * mz[q];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < 13; ++i) s_part[threadIdx.x >> 5][i] = acc[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        float s[13];
        for (int i = 0; i < 13; ++i) {
            s[i] = 0.f;
            for (int k = 0; k < kLossThreads / 32; ++k) s[i] += s_part[k][i];
        }
        dis[b] = s[0] / (float)n_mesh;
        if (d_t) { d_t[3 * b] = s[1]; d_t[3 * b + 1] = s[2]; d_t[3 * b + 2] = s[3]; }
        if (d_r) {
            base_backward(w, x, y, z, nrm, s + 4, d_r + 4 * b);
        }
    }
    // next-iteration cloud and target in the predicted frame: (p - t) . Rm  (row vector times ori_base)
    if (new_points) {
        const float* P = points + (size_t)b * n_points * 3;
        float* O = new_points + (size_t)b * n_points * 3;
        for (int i = threadIdx.x; i < n_points; i += kLossThreads) {
            const float px = P[3 * i] - tx, py = P[3 * i + 1] - ty, pz = P[3 * i + 2] - tz;
            O[3 * i] = px * R[0] + py * R[3] + pz * R[6];
            O[3 * i + 1] = px * R[1] + py * R[4] + pz * R[7];
            O[3 * i + 2] = px * R[2] + py * R[5] + pz * R[8];
        }
    }
    if (new_target) {
        float* O = new_target + (size_t)b * n_mesh * 3;
        for (int i = threadIdx.x; i < n_mesh; i += kLossThreads) {
            const float px = Tg[3 * i] - tx, py = Tg[3 * i + 1] - ty, pz = Tg[3 * i + 2] - tz;
            O[3 * i] = px * R[0] + py * R[3] + pz * R[6];
            O[3 * i + 1] = px * R[1] + py * R[4] + pz * R[7];
            O[3 * i + 2] = px * R[2] + py * R[5] + pz * R[8];
        }
    }
}

// ------------------------------------------------------------------------------ estimator loss, forward + backward
// `Loss` (DenseFusion/lib/loss.py:12-73) for the N per-point candidate poses of ONE object, with its gradient:
//   q_i = r_i / |r_i|, pred_ij = base(q_i) m_j + (p_i + t_i)  (:14-38);  symmetric and not refine: tgt_ij = nearest target
//   point to pred_ij (:40-47; the index is a constant of the backward pass, the reference detaches it);
//   d_ij = |pred_ij - tgt_ij|, dis_i = mean_j d_ij, std_i = unbiased std_j d_ij (:49-50),
//   loss = mean_i [(dis_i + 2 std_i) c_i - w log c_i]  (:53, the fork's objective).
// One CTA per candidate; the [N,M,3] `pred` tensor is written only when the caller wants the reference's 5th return value,
// the N*M-query kNN and its N*M x M distance matrix (4 GB at 1000 x 1000, knn.h:33) never exist.
//   d loss / d d_ij = (c_i / N) [1/M + 2 (d_ij - dis_i) / ((M - 1) std_i)],  d d_ij / d pred_ij = (pred_ij - tgt_ij) / d_ij,
//   d_t_i = sum_j,  d Rm_i = sum_j g_ij m_j^T -> d_r_i (base_backward),  d_c_i = ((dis_i + 2 std_i) - w / c_i) / N.
// Pass 1 keeps (pred - tgt, d) of every model point in shared memory (dynamic: 16 B x M), pass 2 applies the weights.
__global__ void __launch_bounds__(kLossThreads)
estimator_loss_kernel(const float* __restrict__ pred_r, const float* __restrict__ pred_t, const float* __restrict__ pred_c,
                      const float* __restrict__ points, const float* __restrict__ model, const float* __restrict__ target,
                      int n_mesh, int sym, float w_reg, int n_cand, float* __restrict__ dis, float* __restrict__ std_out,
                      float* __restrict__ term /* (dis + 2 std) c - w log c per candidate */, float* __restrict__ d_r,
                      float* __restrict__ d_t, float* __restrict__ d_c, float* __restrict__ pred_out /* [N,M,3] or NULL */)
{
    __shared__ float4 s_ref[kKnnRefTile];
    __shared__ float s_part[kLossThreads / 32][12];
    __shared__ float s_stat[2];
    extern __shared__ float4 s_g[];                        // [n_mesh]: pred - tgt (xyz), d
    const int b = blockIdx.x;
    const float r0 = pred_r[4 * b], r1 = pred_r[4 * b + 1], r2 = pred_r[4 * b + 2], r3 = pred_r[4 * b + 3];
    const float nrm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
    const float w = r0 / nrm, x = r1 / nrm, y = r2 / nrm, z = r3 / nrm;
    float R[9];
    quat_to_base(w, x, y, z, R);
    const float tx = points[3 * b] + pred_t[3 * b], ty = points[3 * b + 1] + pred_t[3 * b + 1], tz = points[3 * b + 2] + pred_t[3 * b + 2];
    float acc = 0.f;
    for (int p0 = 0; p0 < n_mesh; p0 += kLossThreads * kKnnQ) {
        const int q0 = p0 + threadIdx.x * kKnnQ;
        float qx[kKnnQ], qy[kKnnQ], qz[kKnnQ], best[kKnnQ];
        int bidx[kKnnQ];
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            const int i = min(q0 + q, n_mesh - 1);
            const float mx = model[3 * i], my = model[3 * i + 1], mz = model[3 * i + 2];
            qx[q] = (mx * R[0] + my * R[1] + mz * R[2]) + tx;
            qy[q] = (mx * R[3] + my * R[4] + mz * R[5]) + ty;
            qz[q] = (mx * R[6] + my * R[7] + mz * R[8]) + tz;
            best[q] = FLT_MAX; bidx[q] = i;
            if (pred_out && q0 + q < n_mesh) {
                float* o = pred_out + ((size_t)b * n_mesh + i) * 3;
                o[0] = qx[q]; o[1] = qy[q]; o[2] = qz[q];
            }
        }
        if (sym) {
            for (int t0 = 0; t0 < n_mesh; t0 += kKnnRefTile) {
                const int n = min(kKnnRefTile, n_mesh - t0);
                __syncthreads();
                for (int i = threadIdx.x; i < n; i += kLossThreads)
                    s_ref[i] = make_float4(target[3 * (t0 + i)], target[3 * (t0 + i) + 1], target[3 * (t0 + i) + 2], 0.f);
                __syncthreads();
                scan_tile<false>(s_ref, n, t0, qx, qy, qz, best, bidx);
            }
        }
#pragma unroll
        for (int q = 0; q < kKnnQ; ++q) {
            if (q0 + q < n_mesh) {
                const int i = bidx[q];
                const float dx = qx[q] - target[3 * i], dy = qy[q] - target[3 * i + 1], dz = qz[q] - target[3 * i + 2];
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                s_g[q0 + q] = make_float4(dx, dy, dz, d);
                acc += d;
            }
        }
    }
    // mean, then Bessel-corrected std in a second pass (torch.std)
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < kLossThreads / 32; ++k) s += s_part[k][0];
        s_stat[0] = s / (float)n_mesh;
    }
    __syncthreads();
    const float mean = s_stat[0];
    float qq = 0.f;
    for (int i = threadIdx.x; i < n_mesh; i += kLossThreads) { const float e = s_g[i].w - mean; qq += e * e; }
    qq = warp_sum(qq);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][0] = qq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < kLossThreads / 32; ++k) s += s_part[k][0];
        s_stat[1] = sqrtf(s / (float)(n_mesh - 1));
    }
    __syncthreads();
    const float sd = s_stat[1];
    const float c = pred_c[b];
    const float inv_n = 1.0f / (float)n_cand;
    if (threadIdx.x == 0) {
        dis[b] = mean; std_out[b] = sd;
        term[b] = (mean + 2.0f * sd) * c - w_reg * logf(c);
        if (d_c) d_c[b] = ((mean + 2.0f * sd) - w_reg / c) * inv_n;
    }
    if (!d_r) return;
    // pass 2: weighted gradient sums
    const float g = c * inv_n;
    const float k_mean = g / (float)n_mesh;
    const float k_std = sd > 0.f ? 2.0f * g / ((float)(n_mesh - 1) * sd) : 0.f;
    float a[12] = {};                                      // d_t[3], dRm[9]
    for (int i = threadIdx.x; i < n_mesh; i += kLossThreads) {
        const float4 v = s_g[i];
        const float coef = v.w > 0.f ? (k_mean + k_std * (v.w - mean)) / v.w : 0.f;
        const float gx = v.x * coef, gy = v.y * coef, gz = v.z * coef;
        const float mx = model[3 * i], my = model[3 * i + 1], mz = model[3 * i + 2];
        a[0] += gx; a[1] += gy; a[2] += gz;
        a[3] += gx * mx; a[4] += gx * my; a[5] += gx * mz;
        a[6] += gy * mx; a[7] += gy * my; a[8] += gy * mz;
        a[9] += gz * mx; a[10] += gz * my; a[11] += gz * mz;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = warp_sum(a[i]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < 12; ++i) s_part[threadIdx.x >> 5][i] = a[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        float s[12];
        for (int i = 0; i < 12; ++i) {
            s[i] = 0.f;
            for (int k = 0; k < kLossThreads / 32; ++k) s[i] += s_part[k][i];
        }
        d_t[3 * b] = s[0]; d_t[3 * b + 1] = s[1]; d_t[3 * b + 2] = s[2];
        base_backward(w, x, y, z, nrm, s + 3, d_r + 4 * b);
    }
}

// loss = mean_i term_i (fixed-order reduction), which = arg-max c (lowest index on ties), dis of that candidate, and the
// cloud / target in its frame (loss.py:55-69): new_points = (points - t) Rm, new_target = (target - t) Rm.  One CTA.
__global__ void __launch_bounds__(kLossThreads)
estimator_loss_finish_kernel(const float* __restrict__ pred_r, const float* __restrict__ pred_t, const float* __restrict__ pred_c,
                             const float* __restrict__ points, const float* __restrict__ target, int n_mesh, int n_cand,
                             const float* __restrict__ term, const float* __restrict__ dis, float* __restrict__ loss_out /* [2]: loss, dis[which] */,
                             int32_t* __restrict__ which_out, float* __restrict__ new_points, float* __restrict__ new_target)
{
    __shared__ float s_sum[kLossThreads];
    __shared__ float s_c[kLossThreads];
    __shared__ int s_i[kLossThreads];
    float sum = 0.f, bc = -FLT_MAX;
    int bi = 0;
    for (int i = threadIdx.x; i < n_cand; i += kLossThreads) {
        sum += term[i];
        const float c = pred_c[i];
        if (c > bc) { bc = c; bi = i; }                    // ascending i per thread: first maximum kept
    }
    s_sum[threadIdx.x] = sum; s_c[threadIdx.x] = bc; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kLossThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            const float c2 = s_c[threadIdx.x + o]; const int i2 = s_i[threadIdx.x + o];
            if (c2 > s_c[threadIdx.x] || (c2 == s_c[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_c[threadIdx.x] = c2; s_i[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    const int which = s_i[0];
    if (threadIdx.x == 0) { loss_out[0] = s_sum[0] / (float)n_cand; loss_out[1] = dis[which]; *which_out = which; }
    const float r0 = pred_r[4 * which], r1 = pred_r[4 * which + 1], r2 = pred_r[4 * which + 2], r3 = pred_r[4 * which + 3];
    const float nrm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
    float R[9];
    quat_to_base(r0 / nrm, r1 / nrm, r2 / nrm, r3 / nrm, R);
    const float tx = pred_t[3 * which] + points[3 * which], ty = pred_t[3 * which + 1] + points[3 * which + 1],
                tz = pred_t[3 * which + 2] + points[3 * which + 2];
    for (int pass = 0; pass < 2; ++pass) {
        const float* P = pass == 0 ? points : target;
        float* O = pass == 0 ? new_points : new_target;
        const int n = pass == 0 ? n_cand : n_mesh;
        if (!O) continue;
        for (int i = threadIdx.x; i < n; i += kLossThreads) {
            const float px = P[3 * i] - tx, py = P[3 * i + 1] - ty, pz = P[3 * i + 2] - tz;
            O[3 * i] = px * R[0] + py * R[3] + pz * R[6];
            O[3 * i + 1] = px * R[1] + py * R[4] + pz * R[7];
            O[3 * i + 2] = px * R[2] + py * R[5] + pz * R[8];
        }
    }
}

}  // namespace ape

extern "C" __attribute__((visibility("default"))) int ape_knn(const float* ref, const float* query, int64_t* idx, int B, int D, int N, int M, int k,
                       int arith, void* stream)
{
    APE_REQUIRE(B >= 0 && D > 0 && N > 0 && M >= 0 && k > 0, "ape_knn: bad sizes");
    if (B == 0 || M == 0) return APE_OK;                       // empty query set: nothing to write
    APE_REQUIRE(ref && query && idx, "ape_knn: null pointer");
    APE_REQUIRE(k <= N, "ape_knn: k (%d) > number of reference points (%d)", k, N);
    APE_REQUIRE(arith == APE_KNN_ARITH_CPU || arith == APE_KNN_ARITH_FMA, "ape_knn: unknown arithmetic mode");
    if (B == 0 || M == 0) return APE_OK;
    APE_REQUIRE(B <= 65535, "ape_knn: batch > 65535 (split the batch)");
    cudaStream_t s = (cudaStream_t)stream;
    ape::ProfScope prof_("knn", s);
    if (D == 3 && k == 1) {
        dim3 grid((M + ape::kKnnThreads * ape::kKnnQ - 1) / (ape::kKnnThreads * ape::kKnnQ), B);
        if (arith == APE_KNN_ARITH_FMA) ape::knn3_top1_kernel<true><<<grid, ape::kKnnThreads, 0, s>>>(ref, query, idx, N, M);
        else ape::knn3_top1_kernel<false><<<grid, ape::kKnnThreads, 0, s>>>(ref, query, idx, N, M);
    } else {
        if (k > ape::kKnnMaxK) {
            ape::set_error("ape_knn: k=%d > %d is not implemented", k, ape::kKnnMaxK);
            return APE_ERR_UNSUPPORTED;
        }
        dim3 grid((M + 127) / 128, B);
        if (arith == APE_KNN_ARITH_FMA) ape::knn_generic_kernel<true><<<grid, 128, 0, s>>>(ref, query, idx, D, N, M, k);
        else ape::knn_generic_kernel<false><<<grid, 128, 0, s>>>(ref, query, idx, D, N, M, k);
    }
    ape::count_launch();
    return ape::check_launch("ape_knn");
}

extern "C" __attribute__((visibility("default"))) int ape_add_metric(const float* quat, const float* trans, const float* model_points, int64_t model_stride,
                              int n_model, const float* target, int64_t target_stride, int n_target,
                              const uint8_t* symmetric, int B, float* dis, int32_t* nn_index, void* stream)
{
    APE_REQUIRE(quat && trans && model_points && target && dis, "ape_add_metric: null pointer");
    APE_REQUIRE(B >= 0 && n_model > 0 && n_target > 0, "ape_add_metric: bad sizes");
    APE_REQUIRE(symmetric || n_model <= n_target, "ape_add_metric: ADD pairs model point i with target row i (n_target >= n_model)");
    if (B == 0) return APE_OK;
    ape::ProfScope prof_("add_metric", (cudaStream_t)stream);
    ape::add_metric_kernel<<<B, ape::kAddThreads, 0, (cudaStream_t)stream>>>(
        quat, trans, model_points, model_stride, n_model, target, target_stride, n_target, symmetric, dis, nn_index, nullptr);
    ape::count_launch();
    return ape::check_launch("ape_add_metric");
}

// Candidate-pose distances of the estimator loss (lib/loss.py:30-50): B = the N per-point candidate poses of one object,
// model / target shared (stride 0).  dis[i] = mean_j |pred_ij - tgt_j| and its unbiased std, with the nearest-neighbour
// target for symmetric objects -- without materialising pred [N,M,3] or the N*M-query kNN.
extern "C" __attribute__((visibility("default"))) int ape_add_metric_std(const float* quat, const float* trans, const float* model_points, int64_t model_stride,
                                  int n_model, const float* target, int64_t target_stride, int n_target,
                                  const uint8_t* symmetric, int B, float* dis, float* std_out, void* stream)
{
    APE_REQUIRE(quat && trans && model_points && target && dis && std_out, "ape_add_metric_std: null pointer");
    APE_REQUIRE(B >= 0 && n_model > 1 && n_target > 0, "ape_add_metric_std: bad sizes");
    APE_REQUIRE(symmetric || n_model <= n_target, "ape_add_metric_std: ADD pairs model point i with target row i (n_target >= n_model)");
    APE_REQUIRE(n_model <= ape::kKnnRefTile, "ape_add_metric_std: at most %d model points", ape::kKnnRefTile);
    if (B == 0) return APE_OK;
    ape::ProfScope prof_("add_metric_std", (cudaStream_t)stream);
    ape::add_metric_kernel<<<B, ape::kAddThreads, 0, (cudaStream_t)stream>>>(
        quat, trans, model_points, model_stride, n_model, target, target_stride, n_target, symmetric, dis, nullptr, std_out);
    ape::count_launch();
    return ape::check_launch("ape_add_metric_std");
}

extern "C" __attribute__((visibility("default"))) int ape_refine_loss(const float* quat, const float* trans, const float* model_points, const float* target,
                               int n_mesh, const float* points, int n_points, const uint8_t* symmetric, int B, float* dis,
                               float* d_r, float* d_t, float* new_points, float* new_target, void* stream)
{
    APE_REQUIRE(quat && trans && model_points && target && dis, "ape_refine_loss: null pointer");
    APE_REQUIRE(B >= 0 && n_mesh > 0, "ape_refine_loss: bad sizes");
    APE_REQUIRE(!new_points || (points && n_points > 0), "ape_refine_loss: new_points needs the input cloud");
    APE_REQUIRE((!new_points || new_points != points) && (!new_target || new_target != target), "ape_refine_loss: outputs must not alias inputs");
    if (B == 0) return APE_OK;
    ape::ProfScope prof_("train.refine_loss", (cudaStream_t)stream);
    ape::refine_loss_kernel<<<B, ape::kLossThreads, 0, (cudaStream_t)stream>>>(quat, trans, model_points, target, n_mesh, points,
                                                                              n_points, symmetric, dis, d_r, d_t, new_points, new_target);
    ape::count_launch();
    return ape::check_launch("ape_refine_loss");
}

// Estimator loss `Loss` (lib/loss.py:12-73), forward + backward, for the n_cand per-point candidate poses of one object.
extern "C" __attribute__((visibility("default")))
int ape_estimator_loss(const float* pred_r, const float* pred_t, const float* pred_c, const float* points, const float* model_points,
                       const float* target, int n_cand, int n_mesh, int symmetric, float w, float* loss_dis /* [2] */,
                       int32_t* which_max, float* dis, float* std_out, float* term, float* d_r, float* d_t, float* d_c,
                       float* new_points, float* new_target, float* pred_out, void* stream)
{
    APE_REQUIRE(pred_r && pred_t && pred_c && points && model_points && target && loss_dis && which_max && dis && std_out && term,
                "ape_estimator_loss: null pointer");
    APE_REQUIRE(n_cand > 0 && n_mesh > 1, "ape_estimator_loss: bad sizes");
    APE_REQUIRE(n_mesh <= 8192, "ape_estimator_loss: at most 8192 mesh points");
    APE_REQUIRE((d_r != nullptr) == (d_t != nullptr), "ape_estimator_loss: d_r and d_t go together");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t dyn = (size_t)n_mesh * sizeof(float4);
    static ape::PerDevice attr_done;
    if (attr_done.first())
        APE_CUDA(cudaFuncSetAttribute(ape::estimator_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * (int)sizeof(float4)));
    {
        ape::ProfScope prof_("estimator_loss", s);
        ape::estimator_loss_kernel<<<n_cand, ape::kLossThreads, dyn, s>>>(pred_r, pred_t, pred_c, points, model_points, target, n_mesh,
                                                                           symmetric ? 1 : 0, w, n_cand, dis, std_out, term, d_r, d_t, d_c, pred_out);
        ape::count_launch();
        int rc = ape::check_launch("ape_estimator_loss");
        if (rc) return rc;
    }
    ape::ProfScope prof_("estimator_loss_finish", s);
    ape::estimator_loss_finish_kernel<<<1, ape::kLossThreads, 0, s>>>(pred_r, pred_t, pred_c, points, target, n_mesh, n_cand, term, dis,
                                                                      loss_dis, which_max, new_points, new_target);
    ape::count_launch();
    return ape::check_launch("ape_estimator_loss (finish)");
}
