/* Oracle (TEST INFRASTRUCTURE, not product): brute-force k-NN restating the
 * arithmetic of the reference's two implementations.
 *   mode 0  DenseFusion/lib/knn/src/cpu/knn_cpu.cpp:8-16  -- fp32, d += diff*diff per
 *           dimension, each product and each sum rounded (gcc, no FMA contraction)
 *   mode 1  DenseFusion/lib/knn/src/cuda/knn.cu:86-92     -- ssd = fma(tmp,tmp,ssd)
 *           (nvcc contracts `ssd += tmp*tmp`; FFMA in the sm_100a SASS)
 * Selection: the k smallest in ascending distance, ascending index on exact
 * ties, 1-based (knn_cpu.cpp:21-43 stable bubble sort; knn.cu:113-176 strict '<').
 * Layout as the reference: ref [B,D,N], query [B,D,M], idx [B,k,M] int64.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static float dist2(const float* ref, const float* qry, int D, int N, int M, int r, int q, int mode)
{
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) {
        volatile float diff = ref[(size_t)d * N + r] - qry[(size_t)d * M + q];
        if (mode == 1) {
            acc = fmaf(diff, diff, acc);
        } else {
            volatile float sq = diff * diff;
            volatile float s = acc + sq;
            acc = s;
        }
    }
    return acc;
}

int oracle_knn(const float* ref, const float* query, int64_t* idx,
               int B, int D, int N, int M, int k, int mode)
{
    if (k < 1 || k > N) return -1;
    float* bd = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* bi = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    for (int b = 0; b < B; ++b) {
        const float* R = ref + (size_t)b * D * N;
        const float* Q = query + (size_t)b * D * M;
        int64_t* O = idx + (size_t)b * k * M;
        for (int q = 0; q < M; ++q) {
            int have = 0;
            for (int r = 0; r < N; ++r) {
                float d = dist2(R, Q, D, N, M, r, q, mode);
                if (have == k && !(d < bd[k - 1])) continue;
                int pos = have < k ? have : k - 1;
                while (pos > 0 && d < bd[pos - 1]) { /* strict: equal stays behind */
                    if (pos < k) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; }
                    --pos;
                }
                bd[pos] = d; bi[pos] = r + 1;
                if (have < k) ++have;
            }
            for (int j = 0; j < k; ++j) O[(size_t)j * M + q] = bi[j];
        }
    }
    free(bd); free(bi);
    return 0;
}

/* ADD / ADD-S metric of loss_refiner.py:39-49 / eval_linemod.py:118-130 in fp32:
 * pred = model * R^T + t (row-major base from the normalised quaternion as in
 * loss_refiner.py:19-29, fp32), sym: nearest target per pred point (mode-0
 * arithmetic) then mean ||pred - target[nn]||, else mean ||pred_i - target_i||.
 * Returns the mean distance; accumulates in double for a stable reference value. */
double oracle_add_metric(const float* quat_wxyz, const float* t, const float* model, int Mq,
                         const float* target, int Nt, int symmetric)
{
    float q0 = quat_wxyz[0], q1 = quat_wxyz[1], q2 = quat_wxyz[2], q3 = quat_wxyz[3];
    float n = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 /= n; q1 /= n; q2 /= n; q3 /= n;
    float R[9] = {
        1.0f - 2.0f * (q2 * q2 + q3 * q3), 2.0f * q1 * q2 - 2.0f * q0 * q3, 2.0f * q0 * q2 + 2.0f * q1 * q3,
        2.0f * q1 * q2 + 2.0f * q3 * q0, 1.0f - 2.0f * (q1 * q1 + q3 * q3), -2.0f * q0 * q1 + 2.0f * q2 * q3,
        -2.0f * q0 * q2 + 2.0f * q1 * q3, 2.0f * q0 * q1 + 2.0f * q2 * q3, 1.0f - 2.0f * (q1 * q1 + q2 * q2)};
    double acc = 0.0;
    for (int i = 0; i < Mq; ++i) {
        const float* m = model + 3 * i;
        float p[3];
        for (int a = 0; a < 3; ++a)
            p[a] = (m[0] * R[3 * a + 0] + m[1] * R[3 * a + 1] + m[2] * R[3 * a + 2]) + t[a];
        const float* g;
        if (symmetric) {
            int best = 0; float bd = INFINITY;
            for (int r = 0; r < Nt; ++r) {
                volatile float dx = target[3 * r] - p[0], dy = target[3 * r + 1] - p[1], dz = target[3 * r + 2] - p[2];
                volatile float sx = dx * dx, sy = dy * dy, sz = dz * dz;
                volatile float s1 = sx + sy; volatile float s2 = s1 + sz;
                if (s2 < bd) { bd = s2; best = r; }
            }
            g = target + 3 * best;
        } else {
            g = target + 3 * i;
        }
        float dx = p[0] - g[0], dy = p[1] - g[1], dz = p[2] - g[2];
        acc += (double)sqrtf(dx * dx + dy * dy + dz * dz);
    }
    return acc / (double)Mq;
}
