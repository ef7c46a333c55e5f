// Per-object dense layers on the tensor cores (sm_100a): out[b, o] = act(bias[o] + sum_k W[o, k] * x[b, k_off(o) + k]) for a
// batch of 17..256 objects -- the global-feature half of PoseNet's conv1_{r,t,c} (network.py:67-68, 107-109 folded into a
// per-object bias) and PoseRefineNet's conv1_{r,t} / conv2_{r,t} (:192-199).
//
// The products are tiny (64 x 1024 x 1920 at most) but the fp32 SIMT kernel they replace re-read the activations from L2
// once per 8 outputs and took 17-20 us per launch, five launches per step.  Here the roles of a GEMM are SWAPPED so that
// the batch can be small: the WEIGHT rows are the M dimension of the MMA (128 outputs per CTA), the objects are N
// (128 or 256 columns, zero / stale rows beyond the batch are computed and ignored), split-bf16 operands with the three
// products A_lo*W_hi + A_hi*W_lo + A_hi*W_hi as everywhere else.  K is split into slices of 128 over the grid
// (grid = outputs/128 x K/128: 120 CTAs for the 1024 -> 1920 layer) so that no CTA streams more than 6 operand stages.
// The K slices of one output tile form a thread-block CLUSTER (4 or 8 CTAs): every CTA leaves its fp32 partial tile in its
// own shared memory (the operand ring is free by then), and after a cluster barrier each CTA sums a share of the objects
// across all slices through distributed shared memory, in slice order (deterministic), applies bias / ReLU and writes the
// layer's output transposed back to [object, output], as fp32 and optionally as split bf16 (the next dense layer's
// operand).  No workspace, no atomics.  (First version: partial tiles through a global workspace + last-arriver ticket:
// 24 us per launch, slower than the SIMT kernel it was to replace.)
#pragma once
#include "gemm_tc2.cuh"
#include <cooperative_groups.h>

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
    float* s_part = reinterpret_cast<float*>(smem);        // [bp objects][128 outputs] fp32, aliases the operand ring after the MMAs
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();

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
        // ===== epilogue, part 1: thread = output row (TMEM lane); partial tile -> own shared memory, object-major =====
        const int quad = warp & 3;
        const int ol = quad * 32 + lane;
        mbar_wait(tfull_bar, 0);                                 // all MMAs retired: the operand ring is free
        tc_fence_after();
        for (int c0 = 0; c0 < p.batch; c0 += 32) {               // columns beyond the batch are never read
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) s_part[(c0 + j) * 128 + ol] = __uint_as_float(v[j]);   // lanes = consecutive outputs: no bank conflicts
        }
    }
    cluster.sync();                                              // every slice's partial tile is in its CTA's shared memory
    if (warp >= 2) {
        // ===== part 2: this CTA finishes the objects b = rank, rank + n_slices, ... of the tile =====
        const int ol = (warp & 3) * 32 + lane;
        const int o = o_row + ol;
        const float bias = __ldg(p.bias + o);
        const float* peer[8];
        for (int sl = 0; sl < n_slices; ++sl) peer[sl] = cluster.map_shared_rank(s_part, sl);
        for (int b = (int)cluster.block_rank(); b < p.batch; b += n_slices) {
            float acc = 0.f;
            for (int sl = 0; sl < n_slices; ++sl) acc += peer[sl][b * 128 + ol];          // slice order: deterministic
            acc += bias;
            const float y = p.relu ? fmaxf(acc, 0.f) : acc;
            p.out[(size_t)b * p.out_ld + o] = y;                  // lanes = consecutive outputs: coalesced
            if (p.xo_hi) {
                const __nv_bfloat16 h = __float2bfloat16_rn(y);
                p.xo_hi[(size_t)b * p.xo_ld + o] = h;
                p.xo_lo[(size_t)b * p.xo_ld + o] = __float2bfloat16_rn(y - __bfloat162float(h));
            }
        }
    }
    cluster.sync();                                              // peers may still be reading this CTA's partial tile
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.bp);
    }
}

}  // namespace tcd
}  // namespace ape
