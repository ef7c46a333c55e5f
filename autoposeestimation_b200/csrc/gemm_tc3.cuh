// CTA-pair tcgen05 GEMM for the DenseFusion 1x1-conv stacks (sm_100a only), third generation.
//
//   D[256 x bn] (fp32, TMEM of two SMs) = sum over 3 split-bf16 passes of  A_p[256 x K] * W_p[bn x K]^T ,  bn in {128, 256}
//
// ncu on the second generation (profiles/r01b_ncu_gemm.csv) showed the tensor pipe 60-80 % busy with the warps
// stalled on nothing they could fix: a 128 x 256 tile per SM needs 48 KB of TMA fill plus 48 KB of MMA operand reads
// per 512 tensor cycles, i.e. 192 B/clk of shared-memory traffic against 128 B/clk.  Here two CTAs of a cluster
// (one TPC) share every B tile through `tcgen05.mma.cta_group::2`: each CTA stages its own 128 rows of A and HALF of
// the B rows (32 KB per k-block instead of 48 KB), the leader CTA issues one M=256 MMA that reads both halves, and
// each CTA keeps the 128 x bn accumulator of its own rows in its own TMEM.  Per SM that is 64 B/clk of fill and
// 64 B/clk of operand reads, and 64 B/clk of L2 -> SM traffic instead of 96.
//
// Roles (320 threads per CTA): warp 0 = TMA producer (both CTAs, `cp.async.bulk.tensor...cta_group::2` signalling the
// leader's full barrier), warp 1 = TMEM allocator (both) + MMA issuer (leader only; `tcgen05.commit...multicast`
// frees the smem stage in both CTAs and publishes the accumulator to both epilogues), warps 2-9 = epilogue (two
// warps per TMEM lane quadrant, each takes half of the tile's columns; they release the accumulator on the leader's
// barrier).  6-stage x 32 KB ring across tile boundaries, two 256-column accumulators, wide tiles scheduled before
// narrow ones (longest first) over the 74 CTA pairs.
#pragma once
#include "gemm_tc2.cuh"

namespace ape {
namespace tc3 {

using namespace ape::tc;
using ape::tc2::bulk_commit;
using ape::tc2::bulk_wait_all;
using ape::tc2::bulk_wait_read;
using ape::tc2::tma_store_2d;

constexpr int kStages3 = 5;
constexpr int kStageA3 = BM * BK * 2;                 // 16 KB: this CTA's 128 rows of A
constexpr int kStageB3 = 128 * BK * 2;                // 16 KB: this CTA's half of the B rows (bn/2 <= 128)
constexpr int kStage3 = kStageA3 + kStageB3;          // 32 KB
constexpr int kEpiWarps = 8;
constexpr int kThreads3 = 64 + 32 * kEpiWarps;        // 320
constexpr int kStgBuf3 = 32 * 64 * 2;                 // one 32-row x 64-col bf16 box = 4 KB
constexpr int kStgWarp3 = 2 * kStgBuf3;               // hi + lo
constexpr int kStaging3 = kEpiWarps * kStgWarp3;      // 64 KB (also the 4 KB column-sum scratch)
constexpr int kSmemBytes3 = kStages3 * kStage3 + kStaging3 + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr uint32_t kTmemCols3 = 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into this CTA's smem whose completion bytes are counted on a barrier given as shared::cluster address
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct PairTile { int g, mp, n0, bn; };

// Wide (256-column) tiles first, then the 128-column ones: longest-processing-time-first over the CTA pairs.
struct PairSched {
    int mp_tiles, n_wide, n_narrow, wide_total, total;
    __device__ PairSched(const Params& p, int bn_full) {
        mp_tiles = p.M / (2 * BM);
        n_wide = bn_full == 256 ? p.N / 256 : 0;
        n_narrow = (p.N - n_wide * 256) / 128;
        wide_total = p.groups * mp_tiles * n_wide;
        total = wide_total + p.groups * mp_tiles * n_narrow;
    }
    __device__ __forceinline__ PairTile get(int t) const {
        PairTile r;
        int n_idx, rest;
        if (t < wide_total) { n_idx = t % n_wide; rest = t / n_wide; r.n0 = n_idx * 256; r.bn = 256; }
        else { const int u = t - wide_total; n_idx = u % n_narrow; rest = u / n_narrow; r.n0 = n_wide * 256 + n_idx * 128; r.bn = 128; }
        r.mp = rest % mp_tiles;
        r.g = rest / mp_tiles;
        return r;
    }
};

// grid = 2 * min(#pair tiles, #SMs / 2), cluster (2,1,1).  Load maps: A box {64 (K), 128 rows}, W box {64 (K), 64 rows},
// SWIZZLE_128B.  Store maps (EPI_RELU_SPLIT): box {64 cols, 32 rows}, SWIZZLE_128B.  p.M must be a multiple of 256.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads3, 1)
gemm_split_bf16_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                            const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                            const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                            const Params p, const int bn_full)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* staging = smem + kStages3 * kStage3;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStaging3);   // used in the leader CTA
    uint64_t* empty_bar = full_bar + kStages3;                               // each CTA waits on its own
    uint64_t* tfull_bar = empty_bar + kStages3;                              // [2] each CTA waits on its own
    uint64_t* tempty_bar = tfull_bar + 2;                                    // [2] used in the leader CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const PairSched sched(p, bn_full);
    const int kb_per_pass = p.K / BK;
    const int iters_per_tile = 3 * kb_per_pass;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        if (p.mode == EPI_RELU_SPLIT) { tma_prefetch_desc(&map_o_hi); tma_prefetch_desc(&map_o_lo); }
#pragma unroll
        for (int s = 0; s < kStages3; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols3);
    tc_fence_before();
    cluster_sync_all();                         // barriers of both CTAs initialised, TMEM allocated in both
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            int it = 0;
            for (int t = pair; t < sched.total; t += n_pairs) {
                const PairTile tl = sched.get(t);
                const int a_k = p.a_k0 + tl.g * p.a_kg;
                const int a_row = (2 * tl.mp + (int)rank) * BM;
                const int bnh = tl.bn >> 1;                                     // B rows staged by this CTA
                const int w_row = tl.g * p.N + tl.n0 + (int)rank * bnh;
                const uint32_t bytes_pair = 2u * (uint32_t)(kStageA3 + bnh * BK * 2);
                for (int i = 0; i < iters_per_tile; ++i, ++it) {
                    const int s = it % kStages3;
                    const uint32_t ph = (uint32_t)(it / kStages3) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const int pass = i / kb_per_pass, kb = i - pass * kb_per_pass;
                    // pass 0: A_lo*W_hi, pass 1: A_hi*W_lo, pass 2: A_hi*W_hi (small terms first)
                    const CUtensorMap* ma = (pass == 0) ? &map_a_lo : &map_a_hi;
                    const CUtensorMap* mw = (pass == 1) ? &map_w_lo : &map_w_hi;
                    unsigned char* sa = smem + s * kStage3;
                    unsigned char* sb = sa + kStageA3;
                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);   // the leader's full barrier
                    if (leader) mbar_expect_tx(&full_bar[s], bytes_pair);
                    tma_load_2d_pair(sa, ma, fb, a_k + kb * BK, a_row);
                    tma_load_2d_pair(sb, mw, fb, kb * BK, w_row);
                    if (bnh > 64) tma_load_2d_pair(sb + 64 * BK * 2, mw, fb, kb * BK, w_row + 64);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader && lane == 0) {
            int it = 0, lt = 0;
            for (int t = pair; t < sched.total; t += n_pairs, ++lt) {
                const PairTile tl = sched.get(t);
                const int acc = lt & 1;
                const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
                mbar_wait(&tempty_bar[acc], aph ^ 1u);            // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16(2 * BM, tl.bn);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int i = 0; i < iters_per_tile; ++i, ++it) {
                    const int s = it % kStages3;
                    const uint32_t ph = (uint32_t)(it / kStages3) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + s * kStage3));
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + s * kStage3 + kStageA3));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(&empty_bar[s]);              // frees this smem stage in both CTAs
                }
                umma_commit_pair(&tfull_bar[acc]);                // accumulator complete, both epilogues may read
            }
        }
    } else {
        // ===== epilogue: warps 2..9; TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int half = ew >> 2;
        unsigned char* stg = staging + ew * kStgWarp3;            // [hi|lo][32 rows x 128 B], SWIZZLE_128B
        float* s_colsum = reinterpret_cast<float*>(staging);      // EPI_RELU_COLSUM: [4][256]
        const int GN = p.groups * p.N;
        const uint32_t te_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
        const uint32_t te_leader1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
        int lt = 0;
        for (int t = pair; t < sched.total; t += n_pairs, ++lt) {
            const PairTile tl = sched.get(t);
            const int acc = lt & 1;
            const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
            const int m_tile = 2 * tl.mp + (int)rank;
            const int row0 = m_tile * BM + quad * 32;
            const int row = row0 + lane;
            const int col_g = tl.g * p.N + tl.n0;                 // first column within [groups*N]
            const int cw = tl.bn >> 1;                            // columns handled by this warp
            const int cbeg = half * cw;
            const float* bias = p.bias + (p.bias_obj_rows > 0 ? (size_t)(row0 / p.bias_obj_rows) * (size_t)GN : 0) + col_g;
            float bl = __ldg(bias + cbeg + lane);                 // prefetched one 32-column chunk ahead
            const bool valid = (p.mode != EPI_RELU_COLSUM) || ((row % p.rows_per_obj) < p.valid_rows);
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256);
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (p.mode == EPI_RELU_SPLIT) {
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + cw; c0 += 64) {
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + (uint32_t)(c0 + 32 * h), v);
                        const bool last = (h == 1) && (c0 + 64 >= cbeg + cw);
                        if (last) {                               // last read of this accumulator: hand it back early
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(acc ? te_leader1 : te_leader0);
                        }
                        const float bcur = bl;
                        if (!last) bl = __ldg(bias + c0 + 32 * h + 32 + lane);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float f0 = fmaxf(__uint_as_float(v[2 * j]) + __shfl_sync(0xffffffffu, bcur, 2 * j), 0.0f);
                            const float f1 = fmaxf(__uint_as_float(v[2 * j + 1]) + __shfl_sync(0xffffffffu, bcur, 2 * j + 1), 0.0f);
                            const __nv_bfloat162 hb = __floats2bfloat162_rn(f0, f1);
                            const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hb);
                            const __nv_bfloat162 lb = __floats2bfloat162_rn(f0 - __uint_as_float(hu << 16),
                                                                            f1 - __uint_as_float(hu & 0xffff0000u));
                            hi[16 * h + j] = hu; lo[16 * h + j] = *reinterpret_cast<const uint32_t*>(&lb);
                        }
                    }
                    // single staging buffer: the previous chunk's TMA store read it while this chunk was converted
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    unsigned char* sh = stg;
                    unsigned char* sl = stg + kStgBuf3;
                    const uint32_t sw = (uint32_t)lane & 7u;          // 128-byte swizzle: 16 B chunk ^= row % 8
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t off = (uint32_t)lane * 128u + ((uint32_t)j ^ sw) * 16u;
                        *reinterpret_cast<uint4*>(sh + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        *reinterpret_cast<uint4*>(sl + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int oc = p.o_c0 + col_g + c0;
                        tma_store_2d(&map_o_hi, sh, oc, row0);
                        tma_store_2d(&map_o_lo, sl, oc, row0);
                        bulk_commit();
                    }
                }
            } else {
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + cw; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    const bool last = c0 + 32 >= cbeg + cw;
                    if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc ? te_leader1 : te_leader0);
                    }
                    const float bcur = bl;
                    if (!last) bl = __ldg(bias + c0 + 32 + lane);
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = fmaxf(__uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bcur, j), 0.0f);
                        f[j] = valid ? x : 0.0f;
                    }
                    // masked column sum over this warp's 32 rows: butterfly transpose-reduce, lane j ends
                    // with the sum of column c0 + j
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
                asm volatile("bar.sync 1, 256;" ::: "memory");        // the eight epilogue warps only
                const int c = threadIdx.x - 64;                        // 0..255: one column each
                if (c < tl.bn) {
                    // fixed order over the quadrants -> deterministic
                    const float sum = (s_colsum[c] + s_colsum[256 + c]) + (s_colsum[512 + c] + s_colsum[768 + c]);
                    p.colsum[(size_t)m_tile * (size_t)GN + col_g + c] = sum;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");        // scratch is free for the next tile
            }
        }
        if (p.mode == EPI_RELU_SPLIT && lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();                         // no CTA leaves while its pair may still signal it or read its smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols3);
    }
}

}  // namespace tc3
}  // namespace ape
