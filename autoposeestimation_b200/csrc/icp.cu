// Point-to-point ICP: one whole registration per CTA, all iterations inside one persistent kernel
// (sm_100a).  Replaces o3d.registration.registration_icp(..., PointToPoint, criteria) as called at
// pc_reconstruction/open3d_utils.py:76-104 (open3d 0.9.0 semantics restated in oracle/icp.py).
//
// Per registration:
//   build   : uniform grid over the (fixed) target cloud, cells of half the threshold (two-cell reach) when that fits
//             the 4096-cell budget, else cells >= threshold (one-cell reach); counting sort of the
//             target into cell order (scratch in global memory: read-only afterwards, L1/L2-resident)
//   iterate : [apply update to the working source cloud] -> nearest target point: own cell first, then the
//             neighbouring cells that can still beat the running best
//             (fp64, d2=(dx*dx+dy*dy)+dz*dz, strict d2<r2, lowest original index on
//             exact ties) -> block reduction of n, sum d2, sum p, sum q -> fitness / rmse / stop rule
//             -> centred 3x3 covariance (second pass over the stored correspondences) -> one-thread
//             fp64 Jacobi SVD (Kabsch / Eigen::umeyama without scaling) -> T = update * T
// Reductions use a fixed tree, so results are run-to-run reproducible.
#include "ape_common.cuh"
#include <cfloat>
#include <cstdlib>

namespace ape {

// Search steps around the query's cell, nearest first.  Step 0 = the query's own cell, 1 / 2 = the cells to its left /
// right in the same x-row, 3.. = the other x-rows (dy, dz), ordered by distance; steps 0..10 are the 3 x 3 neighbourhood
// (one-cell reach), 11..26 complete the 5 x 5 neighbourhood (two-cell reach).
constexpr int kIcpSteps = 27;
constexpr int kIcpSteps1 = 11;
// bit k set when step k has dy (dz) == the given offset; index = offset + 2
__host__ __device__ constexpr uint32_t steps_dy(int i) { return i == 0 ? 0x2818800u : i == 1 ? 0x0280288u : i == 2 ? 0x0006067u : i == 3 ? 0x0500510u : 0x5061000u; }
__host__ __device__ constexpr uint32_t steps_dz(int i) { return i == 0 ? 0x1982000u : i == 1 ? 0x00281A0u : i == 2 ? 0x000181Fu : i == 3 ? 0x0050640u : 0x6604000u; }
__constant__ signed char c_run_dy[kIcpSteps] = {0, 0, 0,  -1, 1, 0, 0,  -1, 1, -1, 1,  -2, 2, 0, 0,  -2, -2, 2, 2, -1, 1, -1, 1,  -2, 2, -2, 2};
__constant__ signed char c_run_dz[kIcpSteps] = {0, 0, 0,  0, 0, -1, 1,  -1, -1, 1, 1,  0, 0, -2, 2,  -1, 1, -1, 1, -2, -2, 2, 2,  -2, -2, 2, 2};

constexpr int kIcpThreads = 128;      // small CTAs, 4 or 6 registrations per SM: one CTA's serial Kabsch/SVD step overlaps the others' searches
constexpr int kIcpWarps = kIcpThreads / 32;
constexpr int kIcpMaxCells = 4096;

// (Tried and dropped, profiles/r02_icp_ab.txt: the cell-sorted target copy in SHARED memory -- 28 B per point, 3 CTAs per SM --
// 375 k registrations/s with 128 threads and 510 k with 256 against 555 k for this version: the scattered candidate loads
// are not what bounds the search, occupancy and the serial steps are; seeding the search with the previous iteration's
// partner: -2 %; 8 CTAs per SM at 64 registers: +1 %.)
struct IcpSmem {
    int cell_start[kIcpMaxCells + 2];        // [c] = first cell-sorted position of cell c, [c + 1] = one past its last
    double red[32][12];                      // up to 32 warps per CTA
    double out[12];
    double U[12];          // current update (3x4 row-major)
    double T[12];          // accumulated transform (3x4 row-major)
    double bbmin[3], bbmax[3];
    int scan_carry;
    signed char run_dy[32], run_dz[32];   // copy of c_run_* (lanes index it with different steps: no constant-cache replays)
};

template <int K, int NT>
__device__ __forceinline__ void block_sum(double (&v)[K], IcpSmem& s) {
    // result in s.out[0..K); safe to call back-to-back (leading barrier protects s.out readers)
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) s.red[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) a += s.red[w][threadIdx.x];
        s.out[threadIdx.x] = a;
    }
    __syncthreads();
}

__device__ __forceinline__ double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// Kabsch rotation from sigma = (1/n) sum (q-qm)(p-pm)^T via one-sided Jacobi SVD (fp64).
// R = U diag(1,1,det(U)det(V)) V^T with singular values sorted descending (Eigen::umeyama Eq. 39-40).
__device__ void kabsch_rotation(const double* sigma, double* R) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = sigma[i];
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                alpha += A[3 * k + p] * A[3 * k + p];
                beta += A[3 * k + q] * A[3 * k + q];
                gamma += A[3 * k + p] * A[3 * k + q];
            }
            if (fabs(gamma) > 1e-17 * sqrt(alpha * beta) && gamma != 0.0) {
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double ap = A[3 * k + p], aq = A[3 * k + q];
                    A[3 * k + p] = c * ap - s * aq; A[3 * k + q] = s * ap + c * aq;
                    const double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = c * vp - s * vq; V[3 * k + q] = s * vp + c * vq;
                }
            }
        }
        if (!rotated) break;
    }
    double sv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) sv[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    // sort columns by singular value, descending (3-element network)
#define APE_SWAPCOL(a, b)                                                                     \
    if (sv[a] < sv[b]) {                                                                      \
        double tmp = sv[a]; sv[a] = sv[b]; sv[b] = tmp;                                       \
        for (int k = 0; k < 3; ++k) {                                                         \
            tmp = A[3 * k + a]; A[3 * k + a] = A[3 * k + b]; A[3 * k + b] = tmp;              \
            tmp = V[3 * k + a]; V[3 * k + a] = V[3 * k + b]; V[3 * k + b] = tmp;              \
        }                                                                                     \
    }
    APE_SWAPCOL(0, 1) APE_SWAPCOL(1, 2) APE_SWAPCOL(0, 1)
#undef APE_SWAPCOL
    double Um[9];
    const double tiny = 1e-13 * sv[0];
    if (!(sv[0] > 0.0)) {                         // sigma == 0: no information, identity
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    for (int k = 0; k < 3; ++k) Um[3 * k] = A[3 * k] / sv[0];
    if (sv[1] > tiny) {
        for (int k = 0; k < 3; ++k) Um[3 * k + 1] = A[3 * k + 1] / sv[1];
    } else {                                      // rank 1: any unit vector orthogonal to u0
        const double ax = fabs(Um[0]), ay = fabs(Um[3]), az = fabs(Um[6]);
        double e[3] = {0, 0, 0};
        e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
        double w[3] = {Um[3] * e[2] - Um[6] * e[1], Um[6] * e[0] - Um[0] * e[2], Um[0] * e[1] - Um[3] * e[0]};
        const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        for (int k = 0; k < 3; ++k) Um[3 * k + 1] = w[k] / n;
    }
    if (sv[2] > tiny) {
        for (int k = 0; k < 3; ++k) Um[3 * k + 2] = A[3 * k + 2] / sv[2];
    } else {                                      // rank <= 2: complete to a right-handed frame
        Um[2] = Um[3] * Um[7] - Um[6] * Um[4];
        Um[5] = Um[6] * Um[1] - Um[0] * Um[7];
        Um[8] = Um[0] * Um[4] - Um[3] * Um[1];
    }
    const double sgn = (det3(Um) * det3(V) < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = Um[3 * i] * V[3 * j] + Um[3 * i + 1] * V[3 * j + 1] + sgn * Um[3 * i + 2] * V[3 * j + 2];
}

// Candidate j at squared distance d against the running best: strict '<' replaces, an exact tie goes to the lowest
// ORIGINAL target index (worig maps cell-sorted position -> original index; bpos < 0 while nothing is accepted).
__device__ __forceinline__ void icp_consider(double d, int j, double& best, int& bpos, int& borig, const int32_t* __restrict__ worig) {
    if (d < best) { best = d; bpos = j; borig = 0x7fffffff; }
    else if (d == best && bpos >= 0) {
        if (borig == 0x7fffffff) borig = worig[bpos];
        const int o = worig[j];
        if (o < borig) { bpos = j; borig = o; }
    }
}

__device__ __forceinline__ int cell_coord(double v, double lo, double inv_h) {
    return (int)floor((v - lo) * inv_h);
}

template <int MINB, int NT /* threads per CTA */>
__global__ void __launch_bounds__(NT, MINB)
icp_p2p_kernel(const double* __restrict__ source, const int32_t* __restrict__ src_offset, const int32_t* __restrict__ src_count,
               const double* __restrict__ target, const int32_t* __restrict__ tgt_offset, int n_reg,
               double threshold, double rel_fitness, double rel_rmse, int max_iter,
               const double* __restrict__ init, double* __restrict__ transform, double* __restrict__ info,
               double* __restrict__ wrk_src, double* __restrict__ wrk_tgt, int32_t* __restrict__ wrk_tgt_orig,
               int32_t* __restrict__ wrk_corr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IcpSmem& s = *reinterpret_cast<IcpSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const double r2 = threshold * threshold;
    if (tid < kIcpSteps) { s.run_dy[tid] = c_run_dy[tid]; s.run_dz[tid] = c_run_dz[tid]; }   // visible after the first barrier below

    for (int reg = blockIdx.x; reg < n_reg; reg += gridDim.x) {
        const int s0 = src_offset[reg], Ns = src_count ? max(0, min(src_count[reg], src_offset[reg + 1] - s0)) : src_offset[reg + 1] - s0;
        const int t0 = tgt_offset[reg], Nt = tgt_offset[reg + 1] - t0;
        const double* src = source + 3 * (size_t)s0;
        const double* tgt = target + 3 * (size_t)t0;
        double* wsrc = wrk_src + 3 * (size_t)s0;
        double* wtgt = wrk_tgt + 3 * (size_t)t0;             // cell-sorted target copy (read-only after the build, L1/L2 resident)
        int32_t* worig = wrk_tgt_orig + t0;
        int32_t* corr = wrk_corr + s0;

        // ---------------- target bounding box
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int j = tid; j < Nt; j += NT) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double v = tgt[3 * j + a];
                lo[a] = fmin(lo[a], v); hi[a] = fmax(hi[a], v);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
        }
        __syncthreads();
        if ((tid & 31) == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { s.red[tid >> 5][a] = lo[a]; s.red[tid >> 5][3 + a] = hi[a]; }
        }
        __syncthreads();
        if (tid < 3) {
            double l = DBL_MAX, h = -DBL_MAX;
            for (int w = 0; w < (NT / 32); ++w) { l = fmin(l, s.red[w][tid]); h = fmax(h, s.red[w][3 + tid]); }
            s.bbmin[tid] = l; s.bbmax[tid] = h;
        }
        __syncthreads();
        // grid geometry (uniform across the CTA).  Cells of half the threshold with a two-cell reach when that fits the
        // cell budget (the candidates of a query shrink from a 30 x 10 x 10 mm run to little more than its own
        // 5 mm cell), else cells of the threshold with a one-cell reach, widened until the grid fits.
        // reach * h > threshold by a hair, so a target within the threshold is never more than `reach` cells away.
        int reach = 2;
        double h = threshold * (0.5 + 1e-9);
        int dim[3];
        for (;;) {
            long cells = 1;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double ext = (Nt > 0) ? (s.bbmax[a] - s.bbmin[a]) : 0.0;
                double d = floor(ext / h) + 1.0;
                if (!(d < 1e6)) d = 1e6;
                dim[a] = (int)d;
                cells *= dim[a];
            }
            if (cells <= kIcpMaxCells) break;
            if (reach == 2) { reach = 1; h = threshold * (1.0 + 1e-9); }
            else h *= 1.25;
        }
        const double inv_h = 1.0 / h;
        const int ncell = dim[0] * dim[1] * dim[2];
        const double g0 = s.bbmin[0], g1 = s.bbmin[1], g2 = s.bbmin[2];

        // ---------------- counting sort of the target into cell order
        // cell_start[c + 1] first counts cell c, becomes its first position (exclusive scan, shifted by one) and then serves
        // as the cell's write cursor: after the scatter it is one past the cell's last point = the first position of cell
        // c + 1, i.e. cell c occupies [cell_start[c], cell_start[c + 1]) without a second per-cell array.
        for (int c = tid; c <= ncell + 1; c += NT) s.cell_start[c] = 0;
        __syncthreads();
        for (int j = tid; j < Nt; j += NT) {
            const int cx = min(max(cell_coord(tgt[3 * j], g0, inv_h), 0), dim[0] - 1);
            const int cy = min(max(cell_coord(tgt[3 * j + 1], g1, inv_h), 0), dim[1] - 1);
            const int cz = min(max(cell_coord(tgt[3 * j + 2], g2, inv_h), 0), dim[2] - 1);
            atomicAdd(&s.cell_start[(cz * dim[1] + cy) * dim[0] + cx + 2], 1);
        }
        __syncthreads();
        // inclusive scan of cell_start[2..ncell+1] in chunks of NT: [c + 1] = points in cells 0..c-1 = begin of c
        if (tid == 0) s.scan_carry = 0;
        __syncthreads();
        for (int c0 = 2; c0 <= ncell + 1; c0 += NT) {
            const int c = c0 + tid;
            int v = (c <= ncell + 1) ? s.cell_start[c] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((tid & 31) >= o) incl += t;
            }
            int* wtot = reinterpret_cast<int*>(&s.red[0][0]);
            if ((tid & 31) == 31) wtot[tid >> 5] = incl;
            __syncthreads();
            int before = s.scan_carry;
            for (int w = 0; w < (tid >> 5); ++w) before += wtot[w];
            if (c <= ncell + 1) s.cell_start[c] = before + incl;
            __syncthreads();
            if (tid == NT - 1) s.scan_carry = before + incl;
            __syncthreads();
        }
        // now cell_start[c + 2] = points in cells 0..c = begin of cell c + 1; shift the view by one: cursor of cell c
        // is cell_start[c + 1] and must start at begin(c) = (old) cell_start[c + 1]; the scan above wrote begin(c + 1)
        // into [c + 2], so [c + 1] already holds begin(c) for c >= 1 and [1] = 0 holds begin(0).
        for (int j = tid; j < Nt; j += NT) {
            const double x = tgt[3 * j], y = tgt[3 * j + 1], z = tgt[3 * j + 2];
            const int cx = min(max(cell_coord(x, g0, inv_h), 0), dim[0] - 1);
            const int cy = min(max(cell_coord(y, g1, inv_h), 0), dim[1] - 1);
            const int cz = min(max(cell_coord(z, g2, inv_h), 0), dim[2] - 1);
            const int c = (cz * dim[1] + cy) * dim[0] + cx;
            const int pos = atomicAdd(&s.cell_start[c + 1], 1);
            wtgt[3 * pos] = x; wtgt[3 * pos + 1] = y; wtgt[3 * pos + 2] = z;
            worig[pos] = j;
        }
        // ---------------- initial transform
        if (tid < 12) {
            const double idv = (tid % 5 == 0) ? 1.0 : 0.0;        // 3x4 identity: entries 0,5,10
            const double v = init ? init[16 * (size_t)reg + tid] : idv;
            s.T[tid] = v; s.U[tid] = v;
        }
        __threadfence_block();
        __syncthreads();

        double fitness = 0.0, rmse = 0.0, ncorr = 0.0;
        int it = 0;
        bool first = true;
        for (;;) {
            // ---- (re)evaluate correspondences; the working cloud is updated in place by s.U
            const double* U = s.U;                                // (broadcast reads: keeps 24 registers free)
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};            // n, sum d2, sum p(3), sum q(3)
            // The point loop is warp-uniform (every lane runs every trip, lanes past Ns are masked) so that the lanes can
            // be brought back together with a warp vote before each scan: without it each lane scanned its runs on its
            // own and the inner loop ran with 7.5 of 32 lanes active (profiles/r01d).
            for (int i0 = (tid & ~31); i0 < Ns; i0 += NT) {
                const int i = i0 + (tid & 31);
                const bool live = i < Ns;
                double px = 0.0, py = 0.0, pz = 0.0;
                if (live) {
                    const double* pin = first ? (src + 3 * i) : (wsrc + 3 * i);
                    const double x0 = pin[0], y0 = pin[1], z0 = pin[2];
                    px = fma(U[0], x0, fma(U[1], y0, fma(U[2], z0, U[3])));
                    py = fma(U[4], x0, fma(U[5], y0, fma(U[6], z0, U[7])));
                    pz = fma(U[8], x0, fma(U[9], y0, fma(U[10], z0, U[11])));
                    wsrc[3 * i] = px; wsrc[3 * i + 1] = py; wsrc[3 * i + 2] = pz;
                }
                const int cx = cell_coord(px, g0, inv_h), cy = cell_coord(py, g1, inv_h), cz = cell_coord(pz, g2, inv_h);
                // best starts at r^2: only candidates with d2 < r^2 can be accepted (open3d's strict radius test).  The
                // query's own cell is scanned first; every other step is an x-row of cells that is skipped when its
                // distance in (y, z) already exceeds the current best, and otherwise trimmed in x to the cells that can
                // still hold a point within the best.  In the common case (nearest neighbour 1-2 mm away, 5 mm cells)
                // that leaves the own cell and one or two neighbours.  Skipping needs boxdist^2 > best strictly (an exact
                // tie in a skipped cell could otherwise win on original index); delta widens the cells against binning
                // round-off; the (y, z) gaps are kept as float LOWER bounds (round-down), the x gaps in fp64.
                double best = r2; int bpos = -1, borig = 0x7fffffff;
                const double delta = 1e-6 * h;
                float sy_l1, sy_l2, sy_r1, sy_r2, sz_l1, sz_l2, sz_r1, sz_r2;
                {
                    const double yl = fmax(py - (g1 + cy * h) - delta, 0.0), yr = fmax((g1 + (cy + 1) * h) - py - delta, 0.0);
                    const double zl = fmax(pz - (g2 + cz * h) - delta, 0.0), zr = fmax((g2 + (cz + 1) * h) - pz - delta, 0.0);
                    float t;
                    t = __double2float_rd(yl); sy_l1 = __fmul_rd(t, t); t = __double2float_rd(yl + h); sy_l2 = __fmul_rd(t, t);
                    t = __double2float_rd(yr); sy_r1 = __fmul_rd(t, t); t = __double2float_rd(yr + h); sy_r2 = __fmul_rd(t, t);
                    t = __double2float_rd(zl); sz_l1 = __fmul_rd(t, t); t = __double2float_rd(zl + h); sz_l2 = __fmul_rd(t, t);
                    t = __double2float_rd(zr); sz_r1 = __fmul_rd(t, t); t = __double2float_rd(zr + h); sz_r2 = __fmul_rd(t, t);
                }
                const double xl = fmax(px - (g0 + cx * h) - delta, 0.0), xr = fmax((g0 + (cx + 1) * h) - px - delta, 0.0);
                const double xl1 = xl * xl, xl2 = (xl + h) * (xl + h), xr1 = xr * xr, xr2 = (xr + h) * (xr + h);
                uint32_t need = 0;                                // the steps this lane may still have to take
                if (live && Nt > 0 && cx >= -reach && cx < dim[0] + reach && cy >= -reach && cy < dim[1] + reach &&
                    cz >= -reach && cz < dim[2] + reach) {
                    need = reach == 2 ? (1u << kIcpSteps) - 1u : (1u << kIcpSteps1) - 1u;
#pragma unroll
                    for (int d = -2; d <= 2; ++d) {               // rows outside the grid
                        if (cy + d < 0 || cy + d >= dim[1]) need &= ~steps_dy(d + 2);
                        if (cz + d < 0 || cz + d >= dim[2]) need &= ~steps_dz(d + 2);
                    }
                }
                double pruned_at = DBL_MAX;                       // the best the bulk pruning below was last done for
                for (;;) {
                    // every lane advances to ITS next step that survives the box test and holds points (short loop) ...
                    int jb = 0, je = 0;
                    if (best < pruned_at) {                       // drop every row whose y gap or z gap alone exceeds the best
                        pruned_at = best;
                        if ((double)sy_l1 > best) need &= ~steps_dy(1);
                        if ((double)sy_r1 > best) need &= ~steps_dy(3);
                        if ((double)sy_l2 > best) need &= ~steps_dy(0);
                        if ((double)sy_r2 > best) need &= ~steps_dy(4);
                        if ((double)sz_l1 > best) need &= ~steps_dz(1);
                        if ((double)sz_r1 > best) need &= ~steps_dz(3);
                        if ((double)sz_l2 > best) need &= ~steps_dz(0);
                        if ((double)sz_r2 > best) need &= ~steps_dz(4);
                    }
                    while (need) {
                        const int k = __ffs(need) - 1;
                        need &= need - 1;
                        const int dy = s.run_dy[k], dz = s.run_dz[k];
                        const float fy = dy == 0 ? 0.f : (dy == -1 ? sy_l1 : (dy == 1 ? sy_r1 : (dy == -2 ? sy_l2 : sy_r2)));
                        const float fz = dz == 0 ? 0.f : (dz == -1 ? sz_l1 : (dz == 1 ? sz_r1 : (dz == -2 ? sz_l2 : sz_r2)));
                        const double lower = (double)fy + (double)fz;
                        if (lower > best) continue;
                        const double rem = best - lower;          // what is left for the x gap
                        const int nl = (xl1 <= rem) + (reach == 2 && xl2 <= rem);
                        const int nr = (xr1 <= rem) + (reach == 2 && xr2 <= rem);
                        int xa = k == 0 ? cx : (k == 2 ? cx + 1 : cx - nl);
                        int xb = k == 0 ? cx : (k == 1 ? cx - 1 : cx + nr);
                        xa = max(xa, 0); xb = min(xb, dim[0] - 1);
                        if (xa > xb) continue;
                        const int row = ((cz + dz) * dim[1] + (cy + dy)) * dim[0];
                        jb = s.cell_start[row + xa]; je = s.cell_start[row + xb + 1];   // cells of an x-row are contiguous
                        if (je > jb) break;
                    }
                    // ... and the warp meets here, so the scans of the 32 lanes run side by side
                    // (measured: working in rounds of four candidates across steps instead of whole runs is 8 % slower)
                    if (!__any_sync(0xffffffffu, je > jb || need != 0u)) break;
                    // groups of four candidates are reduced with min first (independent chains, no branch);
                    // only a group that reaches the running best takes the slow path with the tie rule
                    int j = jb;
                    for (; j + 4 <= je; j += 4) {
                        double d[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double dx = px - wtgt[3 * (j + u)], dy = py - wtgt[3 * (j + u) + 1], dz = pz - wtgt[3 * (j + u) + 2];
                            d[u] = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        }
                        if (fmin(fmin(d[0], d[1]), fmin(d[2], d[3])) <= best) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) icp_consider(d[u], j + u, best, bpos, borig, worig);
                        }
                    }
                    for (; j < je; ++j) {
                        const double dx = px - wtgt[3 * j], dy = py - wtgt[3 * j + 1], dz = pz - wtgt[3 * j + 2];
                        icp_consider(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)), j, best, bpos, borig, worig);
                    }
                }
                if (live) {
                    if (bpos >= 0 && best < r2) {
                        corr[i] = bpos;
                        acc[0] += 1.0; acc[1] += best;
                        acc[2] += px; acc[3] += py; acc[4] += pz;
                        acc[5] += wtgt[3 * bpos]; acc[6] += wtgt[3 * bpos + 1]; acc[7] += wtgt[3 * bpos + 2];
                    } else {
                        corr[i] = -1;
                    }
                }
            }
            block_sum<8, NT>(acc, s);
            const double n = s.out[0];
            const double pfit = fitness, prmse = rmse;
            ncorr = n;
            if (n > 0.0) { fitness = n / (double)Ns; rmse = sqrt(s.out[1] / n); }
            else { fitness = 0.0; rmse = 0.0; }
            if (!first) {
                ++it;
                if ((fabs(pfit - fitness) < rel_fitness && fabs(prmse - rmse) < rel_rmse) || it >= max_iter) break;
            } else {
                first = false;
                if (max_iter <= 0) break;
            }
            // ---- Kabsch update from the current correspondences
            const double inv_n = n > 0.0 ? 1.0 / n : 0.0;
            const double pm[3] = {s.out[2] * inv_n, s.out[3] * inv_n, s.out[4] * inv_n};
            const double qm[3] = {s.out[5] * inv_n, s.out[6] * inv_n, s.out[7] * inv_n};
            double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = tid; i < Ns; i += NT) {
                const int j = corr[i];
                if (j >= 0) {
                    const double p0 = wsrc[3 * i] - pm[0], p1 = wsrc[3 * i + 1] - pm[1], p2 = wsrc[3 * i + 2] - pm[2];
                    const double q0 = wtgt[3 * j] - qm[0], q1 = wtgt[3 * j + 1] - qm[1], q2 = wtgt[3 * j + 2] - qm[2];
                    cov[0] += q0 * p0; cov[1] += q0 * p1; cov[2] += q0 * p2;
                    cov[3] += q1 * p0; cov[4] += q1 * p1; cov[5] += q1 * p2;
                    cov[6] += q2 * p0; cov[7] += q2 * p1; cov[8] += q2 * p2;
                }
            }
            block_sum<9, NT>(cov, s);
            if (tid == 0) {
                double Um[12];
                if (n > 0.0) {
                    double sigma[9], R[9];
                    for (int k = 0; k < 9; ++k) sigma[k] = s.out[k] * inv_n;
                    kabsch_rotation(sigma, R);
                    for (int a = 0; a < 3; ++a) {
                        Um[4 * a] = R[3 * a]; Um[4 * a + 1] = R[3 * a + 1]; Um[4 * a + 2] = R[3 * a + 2];
                        Um[4 * a + 3] = qm[a] - (R[3 * a] * pm[0] + R[3 * a + 1] * pm[1] + R[3 * a + 2] * pm[2]);
                    }
                } else {                                         // no correspondences: identity update
                    for (int k = 0; k < 12; ++k) Um[k] = (k % 5 == 0) ? 1.0 : 0.0;
                }
                double Tn[12];
                for (int a = 0; a < 3; ++a) {                    // T = update * T
                    for (int b = 0; b < 4; ++b)
                        Tn[4 * a + b] = Um[4 * a] * s.T[b] + Um[4 * a + 1] * s.T[4 + b] + Um[4 * a + 2] * s.T[8 + b]
                                        + (b == 3 ? Um[4 * a + 3] : 0.0);
                }
                for (int k = 0; k < 12; ++k) { s.T[k] = Tn[k]; s.U[k] = Um[k]; }
            }
            __syncthreads();
        }
        if (tid < 16) transform[16 * (size_t)reg + tid] = tid < 12 ? s.T[tid] : (tid == 15 ? 1.0 : 0.0);
        if (info && tid == 0) {
            info[4 * (size_t)reg] = fitness; info[4 * (size_t)reg + 1] = rmse;
            info[4 * (size_t)reg + 2] = (double)it; info[4 * (size_t)reg + 3] = ncorr;
        }
        __syncthreads();
    }
}

}  // namespace ape

extern "C" __attribute__((visibility("default"))) size_t ape_icp_work_bytes(int total_source_points, int total_target_points)
{
    const size_t s = (size_t)(total_source_points > 0 ? total_source_points : 0);
    const size_t t = (size_t)(total_target_points > 0 ? total_target_points : 0);
    return 24 * s + 24 * t + 4 * t + 4 * s + 64;
}

// src_count (optional): registration r uses src_count[r] points starting at src_offset[r] (a gapped ragged layout, e.g. the
// output of ape_voxel_down_sample, without repacking on the host).
extern "C" __attribute__((visibility("default"))) int ape_icp_p2p_ex(const double* source, const int32_t* src_offset, const int32_t* src_count,
                           const double* target, const int32_t* tgt_offset, int n_reg, int total_source_points, int total_target_points,
                           double threshold, double rel_fitness, double rel_rmse, int max_iter, const double* init,
                           double* transform, double* info, void* work, void* stream)
{
    APE_REQUIRE(source && src_offset && target && tgt_offset && transform && work, "ape_icp_p2p: null pointer");
    APE_REQUIRE(n_reg >= 0 && total_source_points >= 0 && total_target_points >= 0, "ape_icp_p2p: bad sizes");
    APE_REQUIRE(threshold > 0.0, "ape_icp_p2p: threshold must be > 0 (open3d returns the identity otherwise)");
    APE_REQUIRE((((uintptr_t)work) & 7) == 0, "ape_icp_p2p: work must be 8-byte aligned");
    if (n_reg == 0) return APE_OK;
    double* wsrc = reinterpret_cast<double*>(work);
    double* wtgt = wsrc + 3 * (size_t)total_source_points;
    int32_t* worig = reinterpret_cast<int32_t*>(wtgt + 3 * (size_t)total_target_points);
    int32_t* wcorr = worig + total_target_points;
    const int smem = (int)sizeof(ape::IcpSmem);
    static ape::PerDevice attr_done;
    if (attr_done.first()) {
        APE_CUDA(cudaFuncSetAttribute(ape::icp_p2p_kernel<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        APE_CUDA(cudaFuncSetAttribute(ape::icp_p2p_kernel<6, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    // (Measured and dropped: CTAs drawing registrations from an atomic counter instead of a fixed share -- 535 k against
    // 557 k registrations/s on 3552 problems; co-resident CTAs already absorb each other's idle time.)
    // Resident CTAs per SM: 4 (128 registers, no spills) or 6 (80 registers, a few spills).  The search is latency-bound
    // (fixed-latency fp64 chains, 4 warps per scheduler), so 6 per SM deliver 15 % more registrations per second when the
    // grid stays full (measured: 3552 registrations, 245 k/s against 213 k/s), but a CTA then takes 1.3x as long per
    // registration: pick the one with the shorter makespan for this batch (1184 registrations: 2 rounds of 4 per SM beat
    // 1.33 -> 2 rounds of 6 per SM).
    const int sms = ape::sm_count();
    // Few registrations (the sequential reconstruction loop registers ONE view at a time): a CTA per registration leaves
    // the GPU empty, so the CTA is made 512 threads wide -- the searches of one registration run 4x as parallel; the
    // fixed-tree reductions then sum in a different (still fixed) order than the 128-thread kernel.  APE_ICP_WIDE=0 disables.
    // Measured on the 30-view reconstruction loop (29 registrations of ~800 against ~1800 points, 10 iterations on average,
    // tools/recon_profile.py): 1099 us per registration with 128 threads, 771 with 256, 632 with 512, 619 with 1024 -- what
    // is left is the latency of one query's grid walk plus the one-thread SVD (tools/icp_latency.py: 55-67 us set-up and
    // first search, 40 us per iteration far from convergence, 16 us near it).
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("APE_ICP_WIDE"); wide = e ? atoi(e) : 1; }
    if (wide && n_reg <= sms) {
        static ape::PerDevice attr_w;
        if (attr_w.first()) APE_CUDA(cudaFuncSetAttribute(ape::icp_p2p_kernel<1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ape::ProfScope prof_("icp_p2p", (cudaStream_t)stream);
        ape::icp_p2p_kernel<1, 512><<<n_reg, 512, smem, (cudaStream_t)stream>>>(
            source, src_offset, src_count, target, tgt_offset, n_reg, threshold, rel_fitness, rel_rmse, max_iter, init, transform, info,
            wsrc, wtgt, worig, wcorr);
        ape::count_launch();
        return ape::check_launch("ape_icp_p2p (wide)");
    }
    const double t4 = (double)((n_reg + 4 * sms - 1) / (4 * sms)), t6 = 1.30 * (double)((n_reg + 6 * sms - 1) / (6 * sms));
    const int minb = (n_reg > 4 * sms && t6 < t4) ? 6 : 4;
    const int grid = n_reg < sms * minb ? n_reg : sms * minb;
    ape::ProfScope prof_("icp_p2p", (cudaStream_t)stream);
    if (minb == 6)
        ape::icp_p2p_kernel<6, 128><<<grid, 128, smem, (cudaStream_t)stream>>>(
            source, src_offset, src_count, target, tgt_offset, n_reg, threshold, rel_fitness, rel_rmse, max_iter, init, transform, info,
            wsrc, wtgt, worig, wcorr);
    else
        ape::icp_p2p_kernel<4, 128><<<grid, 128, smem, (cudaStream_t)stream>>>(
            source, src_offset, src_count, target, tgt_offset, n_reg, threshold, rel_fitness, rel_rmse, max_iter, init, transform, info,
            wsrc, wtgt, worig, wcorr);
    ape::count_launch();
    return ape::check_launch("ape_icp_p2p");
}

extern "C" __attribute__((visibility("default"))) int ape_icp_p2p(const double* source, const int32_t* src_offset, const double* target,
                           const int32_t* tgt_offset, int n_reg, int total_source_points, int total_target_points,
                           double threshold, double rel_fitness, double rel_rmse, int max_iter, const double* init,
                           double* transform, double* info, void* work, void* stream)
{
    return ape_icp_p2p_ex(source, src_offset, nullptr, target, tgt_offset, n_reg, total_source_points, total_target_points, threshold,
                          rel_fitness, rel_rmse, max_iter, init, transform, info, work, stream);
}
