// TEST INFRASTRUCTURE: extern "C" shim around the reference's own knn_cpu
// (DenseFusion/lib/knn/src/cpu/knn_cpu.cpp:4, compiled UNMODIFIED from /root/reference by
// oracle/Makefile).  Plays the role of knn.h:54-63 (scratch alloc + batch loop).
#include <cstdint>
#include <cstdlib>

void knn_cpu(float* ref_dev, int ref_width, float* query_dev, int query_width, int height, int k,
             float* dist_dev, long* ind_dev, long* ind_buf);

extern "C" int ref_knn_cpu(const float* ref, const float* query, int64_t* idx, int B, int D, int N, int M, int k)
{
    float* dist = (float*)malloc(sizeof(float) * (size_t)N * M);
    long* buf = (long*)malloc(sizeof(long) * (size_t)N);
    for (int b = 0; b < B; ++b)
        knn_cpu(const_cast<float*>(ref) + (size_t)b * D * N, N, const_cast<float*>(query) + (size_t)b * D * M, M,
                D, k, dist, reinterpret_cast<long*>(idx) + (size_t)b * k * M, buf);
    free(dist); free(buf);
    return 0;
}
