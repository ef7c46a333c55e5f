// tcgen05 / TMEM / TMA GEMM tile for the DenseFusion 1x1-conv stacks (sm_100a only).
//
//   D[128 x 128] (fp32, TMEM) = sum over 3 passes of  A_p[128 x K] * W_p[128 x K]^T
//
// Operands are *split bf16*: every fp32 activation/weight x is stored as hi = bf16(x),
// lo = bf16(x - hi).  The three passes (A_lo,W_hi), (A_hi,W_lo), (A_hi,W_hi) recover ~16 mantissa
// bits per operand with fp32 accumulation, which is what the 1e-4 m / 1e-3 rad parity gate against
// the fp32 reference needs (plain bf16 flips the confidence arg-max).  Mechanically
// the kernel is a plain K-major x K-major GEMM with K' = 3K.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer
// (one lane issues tcgen05.mma, tcgen05.commit releases smem stages), warps 2-5 = epilogue
// (tcgen05.ld of their TMEM lane quadrant -> bias + ReLU -> either split-bf16 store for the next
// layer or the masked column sum that implements AvgPool1d without ever writing the [1024 x N] map).
// 3-stage smem ring x 32 KB; two CTAs fit per SM so one tile's epilogue overlaps another's mainloop.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "ape_common.cuh"

namespace ape {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16;
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr int kSmemBytes = kStages * (kStageBytesA + kStageBytesB) + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t kTmemCols = 128;

enum { EPI_RELU_SPLIT = 0, EPI_RELU_COLSUM = 1, EPI_HEAD_OUT = 2 };

struct Params {
    int M, N, K, groups;          // rows (multiple of 128), outputs per group (multiple of 128), K (multiple of 64)
    int a_k0, a_kg;               // first K column of A, and extra K offset per group
    const float* bias;            // [groups*N], or [n_obj][groups*N] when bias_obj_rows > 0
    int bias_obj_rows;            // rows per object for per-object bias (0 = shared bias)
    int mode;
    __nv_bfloat16 *o_hi, *o_lo;   // EPI_RELU_SPLIT: output [M, o_ld], column = o_c0 + g*N + n
    int o_ld, o_c0;
    float* colsum;                // EPI_RELU_COLSUM: [M/128, groups*N] per-tile column sums of valid rows
    int rows_per_obj, valid_rows; // rows (r % rows_per_obj) >= valid_rows are padding
    // EPI_HEAD_OUT (gemm_tc2.cuh only; PoseNet conv3_{r,t,c} with conv4_{r,t,c} of the object's class folded into the
    // epilogue, network.py:115-130): groups = (r, t, c), N == 128;  pred = relu(D + bias) . w4[class]^T + b4[class]
    const float* w4[3];           // [num_obj*4,128], [num_obj*3,128], [num_obj,128]
    const float* b4[3];
    const int64_t* obj;           // [batch] class ids
    int num_obj, batch;
    float* pred[3];               // pred_r [batch,valid_rows,4], pred_t [batch,valid_rows,3], pred_c [batch,valid_rows] (sigmoid)
    // training forward (gemm_tc2.cuh only; csrc/train.cu): plain bf16 = the A_hi*W_hi pass alone, no lo stores, and the
    // ReLU sign bits of the pooled layer kept for the backward pass
    int passes;                   // 1: A_hi*W_hi only;  0 or 3: the three split-bf16 passes
    int pass_mask;                // gemm_tc2.cuh: when != 0 it overrides `passes`: bit 0 = A_lo*W_hi, bit 1 = A_hi*W_lo, bit 2 = A_hi*W_hi
    int hi_only;                  // EPI_RELU_SPLIT: skip the lo store
    uint32_t* relu_bits;          // EPI_RELU_COLSUM: [M, groups*N/32] bit j of word c/32 = (valid row && relu(x)[c] > 0)
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (-> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM quadrant -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile (rows x 64 bf16, 8-row groups 1024 B apart):
// start address >> 4, LBO = 1 (unused for swizzled K-major), SBO = 1024 >> 4, version = 1 (sm_100),
// layout type 2 = SWIZZLE_128B  (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28 (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ---------------------------------------------------------------------------------- kernel
// grid = (N/128, M/128, groups).  Tensor maps: box {64 (K), 128 (rows)}, SWIZZLE_128B.
__global__ void __launch_bounds__(kThreads, 2)
gemm_split_bf16_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                       const Params p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_a = smem;
    unsigned char* smem_b = smem + kStages * kStageBytesA;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * (kStageBytesA + kStageBytesB));
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* s_colsum = reinterpret_cast<float*>(smem_a);     // reused after the mainloop (EPI_RELU_COLSUM)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y, g = blockIdx.z;
    const int kb_per_pass = p.K / BK;
    const int total_iters = 3 * kb_per_pass;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const int a_k = p.a_k0 + g * p.a_kg;
            const int a_row = m_tile * BM;
            const int w_row = g * p.N + n_tile * BN;
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int pass = it / kb_per_pass, kb = it - pass * kb_per_pass;
                // pass 0: A_lo*W_hi, pass 1: A_hi*W_lo, pass 2: A_hi*W_hi (small terms first)
                const CUtensorMap* ma = (pass == 0) ? &map_a_lo : &map_a_hi;
                const CUtensorMap* mw = (pass == 1) ? &map_w_lo : &map_w_hi;
                mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
                tma_load_2d(smem_a + s * kStageBytesA, ma, &full_bar[s], a_k + kb * BK, a_row);
                tma_load_2d(smem_b + s * kStageBytesB, mw, &full_bar[s], kb * BK, w_row);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + s * kStageBytesA));
                const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + s * kStageBytesB));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4) units
                    umma_bf16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                              (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);           // frees the smem stage once these MMAs retire
            }
            umma_commit(tmem_full_bar);               // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int quad = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int row = m_tile * BM + quad * 32 + lane;
        const int col_g = g * p.N + n_tile * BN;                      // first column within [groups*N]
        const float* bias = p.bias + (p.bias_obj_rows > 0 ? (size_t)(row / p.bias_obj_rows) * (size_t)(p.groups * p.N) : 0) + col_g;
        const bool valid = (p.mode != EPI_RELU_COLSUM) || ((row % p.rows_per_obj) < p.valid_rows);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(__uint_as_float(v[j]) + __ldg(bias + c0 + j), 0.0f);
            if (p.mode == EPI_RELU_SPLIT) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0));
                    const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
                    hi[j] = pack_bf16x2(h0, h1); lo[j] = pack_bf16x2(l0, l1);
                }
                const size_t off = (size_t)row * p.o_ld + p.o_c0 + col_g + c0;
                uint4* dh = reinterpret_cast<uint4*>(p.o_hi + off);
                uint4* dl = reinterpret_cast<uint4*>(p.o_lo + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            } else {
                // masked column sum over this warp's 32 rows: butterfly transpose-reduce, lane j ends
                // with the sum of column c0 + j
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = valid ? f[j] : 0.0f;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        const float send = upper ? f[i] : f[i + off];
                        const float keep = upper ? f[i + off] : f[i];
                        f[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                s_colsum[quad * BN + c0 + lane] = f[0];
            }
        }
        if (p.mode == EPI_RELU_COLSUM) {
            asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps only
            const int t = threadIdx.x - 64;                            // 0..127: one column each
            // fixed order over the quadrants -> deterministic
            const float sum = (s_colsum[t] + s_colsum[BN + t]) + (s_colsum[2 * BN + t] + s_colsum[3 * BN + t]);
            p.colsum[(size_t)m_tile * (size_t)(p.groups * p.N) + col_g + t] = sum;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace tc
}  // namespace ape
