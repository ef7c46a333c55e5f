// Back-to-back tcgen05 GEMM for the PoseNet heads (sm_100a only): conv1_{r,t,c} -> conv2_{r,t,c} in ONE kernel.
//
//   H1 = relu(PF[R,384] * W1^T + GB[object])      network.py:107-109   (1920 = 3 x 640 outputs, K = 384)
//   H2 = relu(H1_h[R,640] * W2_h^T + b2_h)        network.py:111-113   (3 heads x 256 outputs, K = 640)
//
// Unfused (gemm_tc2.cuh) the [R,1920] split-bf16 H1 is written to HBM by one launch and read back by the next:
// 0.25 GB each way at batch 64 x 500 points, which makes both launches as much HBM- as tensor-bound.  Here H1 never
// leaves the SM.  One work item = (128-row block, head); per item, for each of the five 128-column chunks j of H1_h:
//
//   G1(j)  : acc1[j & 1] (128 TMEM columns) = PF_blk * W1_h[j]^T        3 split-bf16 passes x 6 K-blocks, 128x128 MMAs
//   epi1(j): acc1 -> + per-object bias, ReLU, split hi/lo -> A2[j & 1]   shared memory, in the K-major 128-byte-swizzled
//                                                                        layout a UMMA A operand wants (2 K-blocks x hi/lo)
//   G2(j)  : acc2 (256 TMEM columns) += A2[j & 1] * W2_h[:, chunk j]^T   3 passes x 2 K-blocks, 128x256 MMAs
//
// and at the end epi2: acc2 -> + bias, ReLU, split -> TMA store of H2.  The MMA warp issues G1(j+1) before G2(j), so the
// tensor pipe works on the next chunk while the epilogue warps convert the current one; acc1 and A2 are double buffered.
// TMEM: 2 x 128 + 256 = 512 columns.  Shared memory: 3-stage x 32 KB operand ring (G1: 16 KB of PF + 16 KB of W1;
// G2: 32 KB of W2) + 2 x 64 KB A2 buffers = 224 KB.  epi2 stages its TMA stores in the (then idle) A2 buffer 1.
// The accumulation order of G2 differs from the unfused kernel (passes interleave per chunk instead of per layer), so the
// two paths agree to fp32 rounding, not bit for bit.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
#pragma once
#include "gemm_tc2.cuh"

namespace ape {
namespace tc4 {

using namespace ape::tc;
using ape::tc2::mbar_arrive;
using ape::tc2::tma_store_2d;
using ape::tc2::bulk_commit;
using ape::tc2::bulk_wait_read;
using ape::tc2::bulk_wait_all;

constexpr int kK1 = 384, kN1 = 640, kN2 = 256, kHeads = 3;
constexpr int kChunk = 128;                      // H1 columns per chunk = K of one G2 step
constexpr int kChunks = kN1 / kChunk;            // 5
constexpr int kStageF = 32 * 1024;
constexpr int kBlk = BM * BK * 2;                // one [128 x 64] bf16 K-block = 16 KB
constexpr int kA2Buf = 4 * kBlk;                 // [hi kb0 | hi kb1 | lo kb0 | lo kb1] = 64 KB
constexpr int kSmemBytesF = 224 * 1024 + 256 /*barriers*/ + 1024 /*align slack*/;   // ring + A2 buffers = 224 KB in both variants
constexpr int kG1Stages = 3 * (kK1 / BK);        // 18
constexpr int kG2Stages = 3 * (kChunk / BK);     // 6

struct FusedParams {
    int M;                     // rows (multiple of 128)
    const float* gb;           // [n_obj, 1920] per-object bias of conv1_{r,t,c} (bias + global-feature part)
    int rows_per_obj;          // Np
    const float* b2;           // [768]
};

// A2BUFS = 2: double-buffered A2 (2 x 64 KB) + 3 ring stages;  A2BUFS = 1: one A2 buffer + 5 ring stages (epi1 converts
// its first K-block while the G2 of the previous chunk is still reading the buffer, and only then waits for it).
template <int A2BUFS>
__global__ void __launch_bounds__(kThreads, 1)
heads12_fused_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                     const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                     const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                     const FusedParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    constexpr int kStagesF = A2BUFS == 2 ? 3 : 5;
    constexpr int kMaxStages = 5;
    unsigned char* a2 = smem + kStagesF * kStageF;                      // A2BUFS x 64 KB
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(a2 + A2BUFS * kA2Buf);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* t1full = empty_bar + kMaxStages;     // [2] G1 chunk complete in acc1[b]
    uint64_t* t1empty = t1full + 2;              // [2] epilogue has drained acc1[b]
    uint64_t* a2full = t1empty + 2;              // [2] epilogue has written A2[b]
    uint64_t* a2empty = a2full + 2;              // [2] G2 has finished reading A2[b]
    uint64_t* t2full = a2empty + 2;              // acc2 complete
    uint64_t* t2empty = t2full + 1;              // epilogue has drained acc2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t2empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = (p.M / BM) * kHeads;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w1_hi); tma_prefetch_desc(&map_w1_lo);
        tma_prefetch_desc(&map_w2_hi); tma_prefetch_desc(&map_w2_lo);
        tma_prefetch_desc(&map_o_hi); tma_prefetch_desc(&map_o_lo);
#pragma unroll
        for (int s = 0; s < kStagesF; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(&t1full[b], 1); mbar_init(&t1empty[b], 4);
            mbar_init(&a2full[b], 4); mbar_init(&a2empty[b], 1);
        }
        mbar_init(t2full, 1); mbar_init(t2empty, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: the stages in exactly the order the MMA warp consumes them =====
        if (lane == 0) {
            int it = 0;
            auto stage_wait = [&](int& s) {
                s = it % kStagesF;
                mbar_wait(&empty_bar[s], ((uint32_t)(it / kStagesF) & 1u) ^ 1u);
                ++it;
            };
            for (int t = blockIdx.x; t < n_items; t += gridDim.x) {
                const int h = t % kHeads, m_tile = t / kHeads;
                const int a_row = m_tile * BM;
                auto load_g1 = [&](int j) {
                    const int w_row = h * kN1 + j * kChunk;
                    for (int i = 0; i < kG1Stages; ++i) {
                        int s; stage_wait(s);
                        const int pass = i / (kK1 / BK), kb = i - pass * (kK1 / BK);
                        unsigned char* sa = smem + s * kStageF;
                        mbar_expect_tx(&full_bar[s], 2 * kBlk);
                        tma_load_2d(sa, pass == 0 ? &map_a_lo : &map_a_hi, &full_bar[s], kb * BK, a_row);
                        tma_load_2d(sa + kBlk, pass == 1 ? &map_w1_lo : &map_w1_hi, &full_bar[s], kb * BK, w_row);
                    }
                };
                auto load_g2 = [&](int j) {
                    const int w_row = h * kN2;
                    for (int i = 0; i < kG2Stages; ++i) {
                        int s; stage_wait(s);
                        const int pass = i / (kChunk / BK), kb2 = i - pass * (kChunk / BK);
                        unsigned char* sb = smem + s * kStageF;
                        const CUtensorMap* mw = pass == 1 ? &map_w2_lo : &map_w2_hi;
                        const int k = j * kChunk + kb2 * BK;
                        mbar_expect_tx(&full_bar[s], 2 * kBlk);
                        tma_load_2d(sb, mw, &full_bar[s], k, w_row);
                        tma_load_2d(sb + kBlk, mw, &full_bar[s], k, w_row + 128);
                    }
                };
                for (int j = 0; j < kChunks; ++j) {
                    load_g1(j);
                    if (j >= 1) load_g2(j - 1);
                }
                load_g2(kChunks - 1);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, lt = 0;
            const uint32_t idesc1 = make_idesc_bf16(BM, kChunk), idesc2 = make_idesc_bf16(BM, kN2);
            const uint32_t acc2 = tmem_base + 256u;
            for (int t = blockIdx.x; t < n_items; t += gridDim.x, ++lt) {
                auto mma_g1 = [&](int j) {
                    const int b = j & 1;
                    const uint32_t use = (uint32_t)(b == 0 ? lt * 3 + (j >> 1) : lt * 2 + (j >> 1));
                    mbar_wait(&t1empty[b], (use & 1u) ^ 1u);          // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(b * kChunk);
                    for (int i = 0; i < kG1Stages; ++i, ++it) {
                        const int s = it % kStagesF;
                        mbar_wait(&full_bar[s], (uint32_t)(it / kStagesF) & 1u);
                        tc_fence_after();
                        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + s * kStageF));
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + s * kStageF + kBlk));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(d, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc1, (i > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&t1full[b]);
                };
                auto mma_g2 = [&](int j) {
                    const int b = A2BUFS == 2 ? (j & 1) : 0;
                    const uint32_t use = A2BUFS == 2 ? (uint32_t)(b == 0 ? lt * 3 + (j >> 1) : lt * 2 + (j >> 1)) : (uint32_t)(lt * kChunks + j);
                    if (j == 0) mbar_wait(t2empty, ((uint32_t)lt & 1u) ^ 1u);   // epilogue has drained acc2 of the previous item
                    mbar_wait(&a2full[b], use & 1u);                  // epilogue has written this chunk of H1
                    tc_fence_after();
                    const unsigned char* abuf = a2 + b * kA2Buf;
                    for (int i = 0; i < kG2Stages; ++i, ++it) {
                        const int s = it % kStagesF;
                        const int pass = i / (kChunk / BK), kb2 = i - pass * (kChunk / BK);
                        mbar_wait(&full_bar[s], (uint32_t)(it / kStagesF) & 1u);
                        tc_fence_after();
                        // pass 0: A_lo * W_hi, pass 1: A_hi * W_lo, pass 2: A_hi * W_hi
                        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(abuf + ((pass == 0 ? 2 : 0) + kb2) * kBlk));
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + s * kStageF));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(acc2, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc2, (j > 0 || i > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&a2empty[b]);
                    if (j == kChunks - 1) umma_commit(t2full);
                };
                for (int j = 0; j < kChunks; ++j) {
                    mma_g1(j);
                    if (j >= 1) mma_g2(j - 1);
                }
                mma_g2(kChunks - 1);
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int quad = warp & 3;
        const uint32_t sw = (uint32_t)lane & 7u;                      // 128-byte swizzle: 16 B chunk ^= row % 8
        const uint32_t row_off = (uint32_t)(quad * 32 + lane) * 128u;
        // epi2 staging: the rows of this warp's own quadrant in blocks 0 (hi) and 1 (lo) of A2 buffer 1, which only this
        // warp ever writes, so the only hazard (its own pending TMA stores) is covered by bulk_wait_read below
        unsigned char* stg_hi = a2 + (A2BUFS - 1) * kA2Buf + quad * 4096;
        unsigned char* stg_lo = a2 + (A2BUFS - 1) * kA2Buf + kBlk + quad * 4096;
        int lt = 0;
        for (int t = blockIdx.x; t < n_items; t += gridDim.x, ++lt) {
            const int h = t % kHeads, m_tile = t / kHeads;
            const int row0 = m_tile * BM + quad * 32;
            const float* gb = p.gb + (size_t)(row0 / p.rows_per_obj) * (size_t)(kHeads * kN1) + h * kN1;
            for (int j = 0; j < kChunks; ++j) {
                const int b = j & 1;
                const uint32_t use = (uint32_t)(b == 0 ? lt * 3 + (j >> 1) : lt * 2 + (j >> 1));
                float bl = __ldg(gb + j * kChunk + lane);
                mbar_wait(&t1full[b], use & 1u);
                tc_fence_after();
                const int ab = A2BUFS == 2 ? b : 0;
                const uint32_t ause = A2BUFS == 2 ? use : (uint32_t)(lt * kChunks + j);
                const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * kChunk);
                unsigned char* abuf = a2 + ab * kA2Buf;
#pragma unroll 1
                for (int kb2 = 0; kb2 < kChunk / BK; ++kb2) {
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + (uint32_t)(kb2 * 64 + 32 * hh), v);
                        const bool last = (hh == 1) && (kb2 == kChunk / BK - 1);
                        if (last) {                                   // last read of this accumulator: hand it back early
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&t1empty[b]);
                        }
                        const float bcur = bl;
                        if (!last) bl = __ldg(gb + j * kChunk + kb2 * 64 + 32 * hh + 32 + lane);
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float f0 = fmaxf(__uint_as_float(v[2 * q]) + __shfl_sync(0xffffffffu, bcur, 2 * q), 0.0f);
                            const float f1 = fmaxf(__uint_as_float(v[2 * q + 1]) + __shfl_sync(0xffffffffu, bcur, 2 * q + 1), 0.0f);
                            const __nv_bfloat162 hb = __floats2bfloat162_rn(f0, f1);
                            const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hb);
                            const __nv_bfloat162 lb = __floats2bfloat162_rn(f0 - __uint_as_float(hu << 16),
                                                                            f1 - __uint_as_float(hu & 0xffff0000u));
                            hi[16 * hh + q] = hu; lo[16 * hh + q] = *reinterpret_cast<const uint32_t*>(&lb);
                        }
                    }
                    if (kb2 == 0) {
                        mbar_wait(&a2empty[ab], (ause & 1u) ^ 1u);    // the previous G2 on this buffer has finished reading it
                        if (lane == 0) bulk_wait_read<0>();           // (this warp's epi2 stores of the previous item)
                        __syncwarp();
                    }
                    unsigned char* sh = abuf + kb2 * kBlk + row_off;
                    unsigned char* sl = abuf + (2 + kb2) * kBlk + row_off;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t off = ((uint32_t)q ^ sw) * 16u;
                        *reinterpret_cast<uint4*>(sh + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                        *reinterpret_cast<uint4*>(sl + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                    }
                }
                fence_proxy_async();                                  // generic-proxy writes -> visible to the UMMA reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&a2full[ab]);
            }
            // ---- epi2: acc2 -> bias, ReLU, split -> TMA store of H2[:, h*256 ...]
            const float* b2 = p.b2 + h * kN2;
            float bl = __ldg(b2 + lane);
            mbar_wait(t2full, (uint32_t)lt & 1u);
            tc_fence_after();
            const uint32_t t_addr2 = tmem_base + ((uint32_t)(quad * 32) << 16) + 256u;
#pragma unroll 1
            for (int c0 = 0; c0 < kN2; c0 += 64) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr2 + (uint32_t)(c0 + 32 * hh), v);
                    const bool last = (hh == 1) && (c0 + 64 >= kN2);
                    if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(t2empty);
                    }
                    const float bcur = bl;
                    if (!last) bl = __ldg(b2 + c0 + 32 * hh + 32 + lane);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float f0 = fmaxf(__uint_as_float(v[2 * q]) + __shfl_sync(0xffffffffu, bcur, 2 * q), 0.0f);
                        const float f1 = fmaxf(__uint_as_float(v[2 * q + 1]) + __shfl_sync(0xffffffffu, bcur, 2 * q + 1), 0.0f);
                        const __nv_bfloat162 hb = __floats2bfloat162_rn(f0, f1);
                        const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hb);
                        const __nv_bfloat162 lb = __floats2bfloat162_rn(f0 - __uint_as_float(hu << 16),
                                                                        f1 - __uint_as_float(hu & 0xffff0000u));
                        hi[16 * hh + q] = hu; lo[16 * hh + q] = *reinterpret_cast<const uint32_t*>(&lb);
                    }
                }
                if (lane == 0) bulk_wait_read<0>();                   // the previous 64-column box has left the staging rows
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t off = (uint32_t)lane * 128u + ((uint32_t)q ^ sw) * 16u;
                    *reinterpret_cast<uint4*>(stg_hi + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                    *reinterpret_cast<uint4*>(stg_lo + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_o_hi, stg_hi, h * kN2 + c0, row0);
                    tma_store_2d(&map_o_lo, stg_lo, h * kN2 + c0, row0);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc4
}  // namespace ape
