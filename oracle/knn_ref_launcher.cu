// Bench/test-only launcher around the reference's own CUDA kNN (TEST INFRASTRUCTURE).
// The reference kernel file DenseFusion/lib/knn/src/cuda/knn.cu is torch-free; it is
// compiled UNMODIFIED from /root/reference by oracle/Makefile and linked with this
// extern "C" shim, which plays the role of knn.h:33-40 (scratch alloc + batch loop).
#include <cuda_runtime.h>
#include <cstdint>

void knn_device(float* ref_dev, int ref_nb, float* query_dev, int query_nb, int dim, int k,
                float* dist_dev, long* ind_dev, cudaStream_t stream);   // knn.cu:217

extern "C" int ref_knn_cuda(const float* ref, const float* query, int64_t* idx, float* dist_scratch,
                            int B, int D, int N, int M, int k, void* stream)
{
    for (int b = 0; b < B; ++b)
        knn_device(const_cast<float*>(ref) + (size_t)b * D * N, N,
                   const_cast<float*>(query) + (size_t)b * D * M, M, D, k, dist_scratch,
                   reinterpret_cast<long*>(idx) + (size_t)b * k * M, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}
