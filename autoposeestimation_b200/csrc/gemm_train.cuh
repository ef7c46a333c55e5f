// Backward GEMMs of the PoseRefineNet training step (DenseFusion/tools/train.py:215-233) on tcgen05 / TMEM / TMA.
// bf16 operands, fp32 accumulation (BASELINE config 5 is a bf16 training step).
//
//   DGRAD :  dX[R, N]   = dY[R, K] * W[K, N]                       ( * ReLU mask of the forward activation )
//            A = dY, row-major [R, K]            -> K-major operand,  TMA box {64 k, 128 rows}
//            B = W,  row-major [K = out, N = in] -> MN-major operand, TMA boxes {64 n, 64 k}
//   WGRAD :  dW[M, N]  += dY[R, M]^T * X[R, N]   summed over a slice of the R rows (split-K across CTAs)
//            A = dY, row-major [R, M] -> MN-major operand, boxes {64 m, 64 k};  B = X likewise.
//
// Neither form needs a transposed copy of an activation or of a weight: the MN-major shared-memory descriptor
// (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units for SWIZZLE_128B) reads the
// row-major boxes exactly as TMA lands them: 8 k-rows x 128 B per swizzle atom (SBO = 1024 B), consecutive
// 64-element chunks of M/N one box apart (LBO = 64 rows x 128 B = 8192 B), 16 k-rows per MMA = +2048 B.
//
// Same skeleton as gemm_tc2.cuh: persistent CTAs, 4-stage x 48 KB TMA ring across work items, two 256-column TMEM
// accumulators, warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
#pragma once
#include "gemm_tc2.cuh"

namespace ape {
namespace tr {

using namespace ape::tc;
using ape::tc2::kStages2;
using ape::tc2::kStage;
using ape::tc2::kStageA;
using ape::tc2::kTmemCols2;
using ape::tc2::mbar_arrive;

enum { BWD_DGRAD = 0, BWD_WGRAD = 1 };

constexpr int kBoxBytes = 64 * 64 * 2;                 // one MN-major box: 64 k-rows x 128 B
constexpr int kScratch = 4 * 256 * 4;                  // per-quadrant column sums (bias gradient)
constexpr int kSmemBytesBwd = kStages2 * kStage + kScratch + 256 /*barriers*/ + 1024 /*align slack*/;

struct BwdParams {
    int mode;
    int M, N, K, groups;          // DGRAD: M = rows (x128), N = outputs / group (x64), K = contraction (x64)
                                  // WGRAD: M = dW rows / group (x128), N = dW cols (x64), K = rows summed over (x64)
    int a_c0, a_cg;               // A column origin and per-group step (DGRAD: + k, WGRAD: + m)
    int b_c0, b_cg, b_r0, b_rg;   // B origin: DGRAD col = b_c0 + n, row = b_r0 + g*b_rg + k; WGRAD col = b_c0 + g*b_cg + n, row = k
    // DGRAD epilogue: v = acc (+ out[r, c] when add_out); columns n >= mask_from: v = mask[r, .] > 0 ? v : 0 and
    // bias_grad[g*N + n] += column sum; out[r, o_c0 + g*o_cg + n] = bf16(v)
    __nv_bfloat16* out; int o_ld, o_c0, o_cg;
    const __nv_bfloat16* mask; int m_ld, m_c0, m_cg, mask_from;
    int add_out;
    float* bias_grad;
    // WGRAD epilogue: dw[(g*dw_rg + m) * dw_ld + n] += acc   (fp32 red.add: the flat gradient accumulates over
    // objects, refinement iterations and K slices, as dis.backward() accumulates in train.py:222)
    float* dw; int dw_ld, dw_rg;
    int k_splits;
};

__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_major(int m, int n, int a_mn, int b_mn) {
    return make_idesc_bf16(m, n) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct Item { int g, m_tile, n0, bn, kb0, kb1; };

__device__ __forceinline__ Item decode_item(int t, const BwdParams& p, int m_tiles, int n_tiles) {
    Item r;
    const int n_idx = t % n_tiles; t /= n_tiles;
    r.m_tile = t % m_tiles; t /= m_tiles;
    r.g = t % p.groups; t /= p.groups;                  // t = K slice (WGRAD), 0 for DGRAD
    r.n0 = n_idx * 256;
    r.bn = min(256, p.N - r.n0);
    const int kb = p.K / BK;
    r.kb0 = (int)((long long)kb * t / p.k_splits);
    r.kb1 = (int)((long long)kb * (t + 1) / p.k_splits);
    return r;
}

// grid = min(#items, #SMs), 192 threads.  map_a: DGRAD box {64, 128 rows}, WGRAD box {64, 64 rows}; map_b: box {64, 64 rows}.
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_bwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const BwdParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    float* s_colsum = reinterpret_cast<float*>(smem + kStages2 * kStage);             // [4][256]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * kStage + kScratch);
    uint64_t* empty_bar = full_bar + kStages2;
    uint64_t* tfull_bar = empty_bar + kStages2;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool wgrad = p.mode == BWD_WGRAD;
    const int m_tiles = p.M / BM;
    const int n_tiles = (p.N + 255) / 256;
    const int total = p.k_splits * p.groups * m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b);
#pragma unroll
        for (int s = 0; s < kStages2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols2);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const Item w = decode_item(t, p, m_tiles, n_tiles);
                const uint32_t bytes = (uint32_t)(kStageA + w.bn * 128);
                const int a_col = p.a_c0 + w.g * p.a_cg + (wgrad ? w.m_tile * BM : 0);
                const int b_col = p.b_c0 + w.g * p.b_cg + w.n0;
                const int b_row = p.b_r0 + w.g * p.b_rg;
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    const int s = it % kStages2;
                    const uint32_t ph = (uint32_t)(it / kStages2) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    unsigned char* sa = smem + s * kStage;
                    unsigned char* sb = sa + kStageA;
                    mbar_expect_tx(&full_bar[s], bytes);
                    if (wgrad) {
                        tma_load_2d(sa, &map_a, &full_bar[s], a_col, kb * BK);
                        tma_load_2d(sa + kBoxBytes, &map_a, &full_bar[s], a_col + 64, kb * BK);
                    } else {
                        tma_load_2d(sa, &map_a, &full_bar[s], a_col + kb * BK, w.m_tile * BM);
                    }
                    for (int j = 0; j < w.bn / 64; ++j)
                        tma_load_2d(sb + j * kBoxBytes, &map_b, &full_bar[s], b_col + 64 * j, b_row + kb * BK);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, lt = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
                const Item w = decode_item(t, p, m_tiles, n_tiles);
                const int acc = lt & 1;
                const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
                mbar_wait(&tempty_bar[acc], aph ^ 1u);
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16_major(BM, w.bn, wgrad ? 1 : 0, 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
                    const int s = it % kStages2;
                    const uint32_t ph = (uint32_t)(it / kStages2) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kStage);
                    const uint64_t a_desc = wgrad ? make_smem_desc_mn_sw128(sa) : make_smem_desc_sw128(sa);
                    const uint64_t b_desc = make_smem_desc_mn_sw128(sa + kStageA);
                    const uint64_t a_step = wgrad ? 128u : 2u;              // 16 k: +2048 B (MN-major) / +32 B (K-major)
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, a_desc + a_step * (uint64_t)k, b_desc + (uint64_t)(128 * k), idesc,
                                  (kb > w.kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        // ===== epilogue =====
        const int quad = warp & 3;
        int lt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
            const Item w = decode_item(t, p, m_tiles, n_tiles);
            const int acc = lt & 1;
            const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
            const int row = w.m_tile * BM + quad * 32 + lane;
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256);
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (wgrad) {
                float* dst = p.dw + (size_t)(w.g * p.dw_rg + row) * p.dw_ld + w.n0;
                const bool has_k = w.kb1 > w.kb0;
#pragma unroll 1
                for (int c0 = 0; c0 < w.bn; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    if (c0 + 32 >= w.bn) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    if (has_k) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                       __uint_as_float(v[j + 3]));
                    }
                }
            } else {
                __nv_bfloat16* orow = p.out + (size_t)row * p.o_ld + p.o_c0 + w.g * p.o_cg + w.n0;
                const __nv_bfloat16* mrow = p.mask ? p.mask + (size_t)row * p.m_ld + p.m_c0 + w.g * p.m_cg + w.n0 : nullptr;
#pragma unroll 1
                for (int c0 = 0; c0 < w.bn; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    if (c0 + 32 >= w.bn) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    const bool masked = (w.n0 + c0) >= p.mask_from;       // mask_from is a multiple of 32
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.add_out) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 a = *reinterpret_cast<const uint4*>(orow + c0 + 8 * q);
                            const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                f[8 * q + 2 * e] += __uint_as_float(u[e] << 16);
                                f[8 * q + 2 * e + 1] += __uint_as_float(u[e] & 0xffff0000u);
                            }
                        }
                    }
                    if (masked && mrow) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 a = *reinterpret_cast<const uint4*>(mrow + c0 + 8 * q);
                            const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                // forward activations are post-ReLU: positive <=> non-zero magnitude bits, sign clear
                                if (!(__uint_as_float(u[e] << 16) > 0.0f)) f[8 * q + 2 * e] = 0.0f;
                                if (!(__uint_as_float(u[e] & 0xffff0000u) > 0.0f)) f[8 * q + 2 * e + 1] = 0.0f;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * q + 2 * e], f[8 * q + 2 * e + 1]);
                            o[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        *reinterpret_cast<uint4*>(orow + c0 + 8 * q) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    if (p.bias_grad && masked) {
                        // column sums over this warp's 32 rows (butterfly transpose-reduce): lane j <- column c0 + j
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
                        s_colsum[quad * 256 + c0 + lane] = f[0];
                    }
                }
                if (p.bias_grad) {
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int tt = threadIdx.x - 64;
                    for (int c = tt; c < w.bn; c += 128) {
                        if (w.n0 + c >= p.mask_from) {
                            const float sum = (s_colsum[c] + s_colsum[256 + c]) + (s_colsum[512 + c] + s_colsum[768 + c]);
                            atomicAdd(p.bias_grad + w.g * p.N + w.n0 + c, sum);
                        }
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols2);
    }
}

}  // namespace tr
}  // namespace ape
