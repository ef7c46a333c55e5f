// Per-object dense layers on the tensor cores (sm_100a): out[b, o] = act(bias[o] + sum_k W[o, k] * x[b, k_off(o) + k]) for a
// batch of 17..256 objects -- the global-feature half of PoseNet's conv1_{r,t,c} (network.py:67-68, 107-109 folded into a
// per-object bias) and PoseRefineNet's conv1_{r,t} / conv2_{r,t} (:192-199).
//
// The products are tiny (64 x 1024 x 1920 at most) but the fp32 SIMT kernel they replace re-read the activations from L2
// once per 8 outputs and took 17-20 us per launch, five launches per step.  Here the roles of a GEMM are SWAPPED so that
// the batch can be small: the WEIGHT rows are the M dimension of the MMA (128 outputs per CTA), the objects are N
// (128 or 256 columns, zero / stale rows beyond the batch are computed and ignored), split-bf16 operands with the three
// products A_lo*W_hi + A_hi*W_lo + A_hi*W_hi as everywhere else.  K is split into slices of 128 over the grid
// (grid = outputs/128 x K/128: 120 CTAs for the 1024 -> 1920 layer) so that no CTA streams more than 6 operand stages;
// every CTA writes its fp32 partial tile to a workspace and the LAST CTA to arrive at a tile (atomic ticket) adds the
// slices in slice order -- the result does not depend on the arrival order -- applies bias / ReLU and writes the layer's
// output transposed back to [object, output], as fp32 and optionally as split bf16 (the next dense layer's operand).
#pragma once
#include "gemm_tc2.cuh"

namespace ape {
namespace tcd {

using namespace ape::tc;

constexpr int kSlice = 128;                    // K per CTA
constexpr int kStagesD = 4;
constexpr int kStageW = 128 * BK * 2;          // 16 KB: 128 weight rows x 64 K
constexpr int kSmemDense = kStagesD * (kStageW + 256 * BK * 2) + 256 + 1024;

struct DenseParams {
    int n_out, K, bp, batch;                   // outputs (multiple of 128), K (multiple of 128), padded batch (128 / 256), objects
    int x_k0, x_kg, rows_per_group;            // x column of output row o: x_k0 + (o / rows_per_group) * x_kg  (rows_per_group 0: no groups)
    const float* bias;                         // [n_out]
    int relu;
    float* out; int out_ld;                    // [batch, out_ld] fp32
    __nv_bfloat16 *xo_hi, *xo_lo; int xo_ld;   // optional split-bf16 copy of the output [*, xo_ld] (NULL: none)
    float* partial;                            // [K / 128][n_out][bp] fp32 workspace
    int* ticket;                               // [n_out / 128], zero before the first launch; the last CTA re-zeroes it
};

__global__ void __launch_bounds__(kThreads, 1)
dense_swapped_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                     const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo, const DenseParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = kStageW + p.bp * BK * 2;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStagesD * (kStageW + 256 * BK * 2));
    uint64_t* empty_bar = full_bar + kStagesD;
    uint64_t* tfull_bar = empty_bar + kStagesD;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    int* s_last = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o_tile = blockIdx.x, slice = blockIdx.y, n_slices = gridDim.y;
    const int o_row = o_tile * 128;
    const int x_k = p.x_k0 + (p.rows_per_group > 0 ? (o_row / p.rows_per_group) * p.x_kg : 0) + slice * kSlice;
    constexpr int kIters = 3 * (kSlice / BK);              // three products x two 64-wide K blocks

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo); tma_prefetch_desc(&map_x_hi); tma_prefetch_desc(&map_x_lo);
#pragma unroll
        for (int s = 0; s < kStagesD; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.bp);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");       // programmatic dependent launch: see ape_common.cuh
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < kIters; ++it) {
                const int s = it % kStagesD;
                mbar_wait(&empty_bar[s], ((uint32_t)(it / kStagesD) & 1u) ^ 1u);
                const int pass = it / (kSlice / BK), kb = it % (kSlice / BK);
                // pass 0: W_hi * x_lo, pass 1: W_lo * x_hi, pass 2: W_hi * x_hi (small terms first)
                const CUtensorMap* mw = pass == 1 ? &map_w_lo : &map_w_hi;
                const CUtensorMap* mx = pass == 0 ? &map_x_lo : &map_x_hi;
                unsigned char* sw = smem + s * stage_bytes;
                unsigned char* sx = sw + kStageW;
                mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                tma_load_2d(sw, mw, &full_bar[s], slice * kSlice + kb * BK, o_row);
                tma_load_2d(sx, mx, &full_bar[s], x_k + kb * BK, 0);
                if (p.bp > 128) tma_load_2d(sx + 128 * BK * 2, mx, &full_bar[s], x_k + kb * BK, 128);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, p.bp);
            for (int it = 0; it < kIters; ++it) {
                const int s = it % kStagesD;
                mbar_wait(&full_bar[s], (uint32_t)(it / kStagesD) & 1u);
                tc_fence_after();
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + s * stage_bytes));
                const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + s * stage_bytes + kStageW));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_bf16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tfull_bar);
        }
    } else {
        // ===== epilogue: thread = output row o (TMEM lane), columns = objects =====
        const int quad = warp & 3;
        const int o = o_row + quad * 32 + lane;
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        float* part = p.partial + ((size_t)slice * p.n_out + o) * p.bp;
        for (int c0 = 0; c0 < p.bp; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            if (c0 < p.batch) {                                  // columns beyond the batch are never read back
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(part + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) *s_last = (atomicAdd(p.ticket + o_tile, 1) == n_slices - 1) ? 1 : 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_last) {
            __threadfence();
            const float bias = __ldg(p.bias + o);
            for (int c0 = 0; c0 < p.batch; c0 += 4) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int sl = 0; sl < n_slices; ++sl) {          // slice order: independent of which CTA arrived last
                    const float4 q = __ldcg(reinterpret_cast<const float4*>(p.partial + ((size_t)sl * p.n_out + o) * p.bp + c0));
                    acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
                }
                const float r[4] = {acc.x + bias, acc.y + bias, acc.z + bias, acc.w + bias};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = c0 + j;
                    if (b < p.batch) {
                        const float y = p.relu ? fmaxf(r[j], 0.f) : r[j];
                        p.out[(size_t)b * p.out_ld + o] = y;      // lanes = consecutive outputs: coalesced
                        if (p.xo_hi) {
                            const __nv_bfloat16 h = __float2bfloat16_rn(y);
                            p.xo_hi[(size_t)b * p.xo_ld + o] = h;
                            p.xo_lo[(size_t)b * p.xo_ld + o] = __float2bfloat16_rn(y - __bfloat162float(h));
                        }
                    }
                }
            }
            if (threadIdx.x == 64) p.ticket[o_tile] = 0;         // ready for the next launch (stream order)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.bp);
    }
}

}  // namespace tcd
}  // namespace ape
