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
// accumulators, warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue (two warps per TMEM lane quadrant, each
// draining half of the tile's columns: with K as short as 512 the epilogue, not the MMA, is the critical path).
// Same-address fp32 atomics serialise at ~130 ns each on B200 (measured: 1024-deep chains cost > 100 us), so bias
// gradients are accumulated per CTA in shared memory and flushed once, spread over kBiasCopies partial vectors.
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
using ape::tc2::tma_store_2d;
using ape::tc2::bulk_commit;
using ape::tc2::bulk_wait_read;
using ape::tc2::bulk_wait_all;

enum { BWD_DGRAD = 0, BWD_WGRAD = 1 };

constexpr int kBoxBytes = 64 * 64 * 2;                 // one MN-major box: 64 k-rows x 128 B
constexpr int kScratch = 4 * 256 * 4 + 1024 * 4;       // per-quadrant column sums + per-CTA bias-gradient accumulators
constexpr int kThreadsBwd = 320;                       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int kBiasCopies = 8;                         // bias partial sums are spread over this many copies (short atomic chains)
constexpr int kStgChunk = 32 * 64 * 2;                  // one epilogue chunk: 32 rows x 64 bf16, 128-byte swizzled (TMA box)
constexpr int kStgWarpBwd = 2 * kStgChunk;             // two chunk buffers per epilogue warp
constexpr int kStagingBwd = 8 * kStgWarpBwd;           // 64 KB: DGRAD only -- it overlays ring stage 3 (DGRAD runs 3 stages) + 16 KB
constexpr int kRingBwd = 3 * kStage + kStagingBwd;     // = 4 stages + 16 KB
constexpr int kSmemBytesBwd = kRingBwd + kScratch + 512 /*barriers*/ + 1024 /*align slack*/;

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
    float* bias_grad;             // [kBiasCopies][bg_stride] partial sums (caller reduces the copies); index g*N + n within a copy
    int bg_stride;
    // WGRAD epilogue: dw[(g*dw_rg + m) * dw_ld + n] += acc   (fp32 red.add: the flat gradient accumulates over
    // objects, refinement iterations and K slices, as dis.backward() accumulates in train.py:222)
    float* dw; int dw_ld, dw_rg;
    int k_splits;
    int b_box_chunks;             // 64-column chunks per B box (map_b); A of WGRAD always has 2
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

// One TMA op for a whole MN-major operand stage: the row-major matrix [R, C] is described to TMA as a 3-D tensor
// (64 columns, R rows, C/64 column chunks); a box {64, 64 rows, n chunks} lands chunk-major in shared memory, i.e. exactly
// the [chunk][k-row][128 B] layout the MN-major descriptor reads (LBO = 8192).  Chunks past the matrix are zero-filled.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Item { int g, m_tile, n0, bn, kb0, kb1; };

__device__ __forceinline__ Item decode_item(int t, const BwdParams& p, int m_tiles, int n_tiles) {
    Item r;
    // With a grid that is a multiple of n_tiles the static round-robin would hand a CTA the SAME n tile every time (all
    // 256-wide or all 128-wide tiles when N = 384): rotate by the CTA's local tile counter so the widths alternate.
    const int rot = (gridDim.x % n_tiles == 0) ? t / (int)gridDim.x : 0;
    const int n_idx = (t % n_tiles + rot) % n_tiles; t /= n_tiles;
    r.m_tile = t % m_tiles; t /= m_tiles;
    r.g = t % p.groups; t /= p.groups;                  // t = K slice (WGRAD), 0 for DGRAD
    r.n0 = n_idx * 256;
    r.bn = min(256, p.N - r.n0);
    const int kb = p.K / BK;
    r.kb0 = (int)((long long)kb * t / p.k_splits);
    r.kb1 = (int)((long long)kb * (t + 1) / p.k_splits);
    return r;
}

// grid = min(#items, #SMs), 320 threads.  map_a: DGRAD 2-D box {64, 128 rows}, WGRAD 3-D box {64, 64 rows, 2 chunks};
// map_b: 3-D box {64, 64 rows, b_box_chunks}; map_mask / map_out: 2-D box {64, 32 rows} (DGRAD epilogue chunks).
__global__ void __launch_bounds__(kThreadsBwd, 1)
gemm_bf16_bwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_mask, const __grid_constant__ CUtensorMap map_out, const BwdParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* staging = smem + 3 * kStage;                                        // DGRAD epilogue chunks (over ring stage 3)
    float* s_colsum = reinterpret_cast<float*>(smem + kRingBwd);                       // [4][256]
    float* s_bias = s_colsum + 4 * 256;                                                // [groups * N <= 1024]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRingBwd + kScratch);
    uint64_t* empty_bar = full_bar + kStages2;
    uint64_t* tfull_bar = empty_bar + kStages2;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* stg_bar = tempty_bar + 2;                                                // [8 warps][2 buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_bar + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool wgrad = p.mode == BWD_WGRAD;
    const int m_tiles = p.M / BM;
    const int n_tiles = (p.N + 255) / 256;
    const int total = p.k_splits * p.groups * m_tiles * n_tiles;
    const int n_stages = wgrad ? kStages2 : 3;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b);
        if (!wgrad) { tma_prefetch_desc(&map_mask); tma_prefetch_desc(&map_out); }
#pragma unroll
        for (int s = 0; s < kStages2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < 16; ++s) mbar_init(&stg_bar[s], 1);
#pragma unroll
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols2);
    for (int i = threadIdx.x; i < 1024; i += kThreadsBwd) s_bias[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const Item w = decode_item(t, p, m_tiles, n_tiles);
                const uint32_t bytes = (uint32_t)(kStageA + p.b_box_chunks * kBoxBytes);    // full boxes, zero-filled or not
                const int a_col = p.a_c0 + w.g * p.a_cg + (wgrad ? w.m_tile * BM : 0);
                const int b_col = p.b_c0 + w.g * p.b_cg + w.n0;
                const int b_row = p.b_r0 + w.g * p.b_rg;
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    unsigned char* sa = smem + s * kStage;
                    unsigned char* sb = sa + kStageA;
                    mbar_expect_tx(&full_bar[s], bytes);
                    if (wgrad) tma_load_3d(sa, &map_a, &full_bar[s], 0, kb * BK, a_col >> 6);
                    else tma_load_2d(sa, &map_a, &full_bar[s], a_col + kb * BK, w.m_tile * BM);
                    tma_load_3d(sb, &map_b, &full_bar[s], 0, b_row + kb * BK, b_col >> 6);
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int s = 0, lt = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
                const Item w = decode_item(t, p, m_tiles, n_tiles);
                const int acc = lt & 1;
                const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
                mbar_wait(&tempty_bar[acc], aph ^ 1u);
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16_major(BM, w.bn, wgrad ? 1 : 0, 1);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
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
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        // ===== epilogue: warps 2..9; TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        uint32_t ld_ph0 = 0u, ld_ph1 = 0u;                      // completed-load counters of this warp's two staging buffers
        int lt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
            const Item w = decode_item(t, p, m_tiles, n_tiles);
            const int acc = lt & 1;
            const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
            const int row = w.m_tile * BM + quad * 32 + lane;
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256);
            if (wgrad) {
                mbar_wait(&tfull_bar[acc], aph);
                tc_fence_after();
                float* dst = p.dw + (size_t)(w.g * p.dw_rg + row) * p.dw_ld + w.n0;
                const bool has_k = w.kb1 > w.kb0;
                const int c_beg = half * (w.bn >> 1), c_end = c_beg + (w.bn >> 1);
#pragma unroll 1
                for (int c0 = c_beg; c0 < c_end; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    if (c0 + 32 >= c_end) {
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
                // DGRAD: global traffic goes through 128-byte-swizzled staging chunks (32 rows x 64 columns) moved by TMA:
                // mask (and addend) chunks are fetched at the start of the tile, before the accumulator is waited for;
                // the result overwrites the mask chunk in place and leaves by TMA store.  (Per-lane row accesses from
                // registers made the LSU the bottleneck: 32 distinct lines per request.)
                // bn >= 128: each half takes bn/2 columns = 1 or 2 chunks, chunk i in buffer i.
                // bn == 64 : half 0 takes the chunk (buffer 0 = mask / result, buffer 1 = addend), half 1 only keeps in step.
                const int n_chunks = w.bn >= 128 ? (w.bn >> 7) : (half == 0 ? 1 : 0);
                const int c_beg = w.bn >= 128 ? half * (w.bn >> 1) : 0;
                const int row0 = w.m_tile * BM + quad * 32;
                const int o_col = p.o_c0 + w.g * p.o_cg + w.n0;
                const int m_col = p.m_c0 + w.g * p.m_cg + w.n0;
                unsigned char* buf0 = staging + (warp - 2) * kStgWarpBwd;
                uint64_t* bar0 = stg_bar + (warp - 2) * 2;
                const bool has_mask = p.mask != nullptr;
                (void)row;
                // ---- prefetch (previous tile's stores must have finished reading the buffers)
                if (lane == 0) {
                    bulk_wait_read<0>();
                    for (int ch = 0; ch < n_chunks; ++ch) {
                        const int c0 = c_beg + 64 * ch;
                        if (has_mask && (w.n0 + c0) >= p.mask_from) {
                            mbar_expect_tx(&bar0[ch], kStgChunk);
                            tma_load_2d(buf0 + ch * kStgChunk, &map_mask, &bar0[ch], m_col + c0, row0);
                        }
                    }
                    if (p.add_out && n_chunks > 0) {
                        mbar_expect_tx(&bar0[1], kStgChunk);
                        tma_load_2d(buf0 + kStgChunk, &map_out, &bar0[1], o_col + c_beg, row0);
                    }
                }
                __syncwarp();
                mbar_wait(&tfull_bar[acc], aph);
                tc_fence_after();
                if (n_chunks == 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                const uint32_t sw = (uint32_t)lane & 7u;
                for (int ch = 0; ch < n_chunks; ++ch) {
                    const int c0 = c_beg + 64 * ch;
                    const bool masked = has_mask && (w.n0 + c0) >= p.mask_from;
                    unsigned char* buf = buf0 + ch * kStgChunk;
                    if (masked) {
                        if (ch == 0) { mbar_wait(&bar0[0], ld_ph0 & 1u); ++ld_ph0; } else { mbar_wait(&bar0[1], ld_ph1 & 1u); ++ld_ph1; }
                    }
                    if (p.add_out) { mbar_wait(&bar0[1], ld_ph1 & 1u); ++ld_ph1; }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + (uint32_t)(c0 + 32 * h), v);
                        if (h == 1 && ch == n_chunks - 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                        }
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (p.add_out) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 a = *reinterpret_cast<const uint4*>(buf0 + kStgChunk + (uint32_t)lane * 128u + (((uint32_t)(4 * h + q)) ^ sw) * 16u);
                                const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    f[8 * q + 2 * e] += __uint_as_float(u[e] << 16);
                                    f[8 * q + 2 * e + 1] += __uint_as_float(u[e] & 0xffff0000u);
                                }
                            }
                        }
                        if (masked) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 a = *reinterpret_cast<const uint4*>(buf + (uint32_t)lane * 128u + (((uint32_t)(4 * h + q)) ^ sw) * 16u);
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
                                const __nv_bfloat162 hh = __floats2bfloat162_rn(f[8 * q + 2 * e], f[8 * q + 2 * e + 1]);
                                o[e] = *reinterpret_cast<const uint32_t*>(&hh);
                            }
                            *reinterpret_cast<uint4*>(buf + (uint32_t)lane * 128u + (((uint32_t)(4 * h + q)) ^ sw) * 16u) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                        if (p.bias_grad && masked) {
                            // column sums over this warp's 32 rows (butterfly transpose-reduce): lane j <- column c0 + 32h + j
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
                            s_colsum[quad * 256 + c0 + 32 * h + lane] = f[0];
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_out, buf, o_col + c0, row0);
                        bulk_commit();
                    }
                }
                if (p.bias_grad) {
                    // combine the four lane quadrants and accumulate per CTA; one thread owns a column, so the running
                    // sums need no atomics until the CTA flushes them once at the end
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    const int c = threadIdx.x - 64;                    // 0..255
                    if (c < w.bn && w.n0 + c >= p.mask_from)
                        s_bias[w.g * p.N + w.n0 + c] += (s_colsum[c] + s_colsum[256 + c]) + (s_colsum[512 + c] + s_colsum[768 + c]);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            }
        }
        if (!wgrad && lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols2);
    }
    if (!wgrad && p.bias_grad) {
        float* dst = p.bias_grad + (size_t)(blockIdx.x % kBiasCopies) * p.bg_stride;
        for (int i = threadIdx.x; i < p.groups * p.N; i += kThreadsBwd) {
            const float v = s_bias[i];
            if (v != 0.f) atomicAdd(dst + i, v);
        }
    }
}

}  // namespace tr
}  // namespace ape
