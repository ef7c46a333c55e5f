// Persistent tcgen05 / TMEM / TMA GEMM for the DenseFusion 1x1-conv stacks (sm_100a only), second generation.
//
//   D[128 x bn] (fp32, TMEM) = sum over 3 split-bf16 passes of  A_p[128 x K] * W_p[bn x K]^T ,  bn in {128, 256}
//
// What changed against gemm_tc.cuh (kept as APE_GEMM_TCGEN05_V1 for A/B runs):
//   * 128 x 256 tiles: one tcgen05.mma (M=128, N=256, K=16) reads 4 KB of A + 8 KB of B from shared memory per
//     128 tensor-pipe cycles (96 B/clk); the 128 x 128 tile needs 8 KB per 64 cycles = 128 B/clk, which is the
//     whole shared-memory bandwidth of an SM and is why v1 topped out near 1000 TFLOP/s.
//   * persistent CTAs (one per SM, static round-robin over tiles, n fastest so the A row-block stays L2-hot),
//     4-stage x 48 KB TMA ring that runs across tile boundaries;
//   * two TMEM accumulators (2 x 256 columns): the epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue stores through shared memory + TMA (cp.async.bulk.tensor store, one 128-byte-swizzled 32-row x
//     64-column box of hi and of lo per warp and chunk; the TMA engine drains the staging buffer while the warp
//     converts the next chunk) instead of 16-byte-per-row scattered global stores.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
// (TMEM lane quadrant = warp % 4).
#pragma once
#include "gemm_tc.cuh"

namespace ape {
namespace tc2 {

using namespace ape::tc;   // PTX wrappers, Params, descriptors

constexpr int kStages2 = 4;
constexpr int kStageA = BM * BK * 2;                  // 16 KB
constexpr int kStageB = 256 * BK * 2;                 // 32 KB (bn = 128 tiles use the first half)
constexpr int kStage = kStageA + kStageB;             // 48 KB
constexpr int kStgBuf = 32 * 64 * 2;                  // one 32-row x 64-col bf16 box = 4 KB
constexpr int kStgWarp = 2 /*hi, lo*/ * kStgBuf;      // 8 KB per epilogue warp
constexpr int kStaging = 4 * kStgWarp;                // 32 KB (also the 4 KB column-sum scratch)
constexpr int kSmemBytes2 = kStages2 * kStage + kStaging + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr uint32_t kTmemCols2 = 512;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct Tile { int g, m_tile, n0, bn; };

// Tail splitting: with T wide tiles on G persistent CTAs the last round is only partly filled (conv5: 512 tiles on 148
// SMs = 3.46 rounds, heads2: 768 = 5.19).  When the leftover `rem = T mod G` wide tiles fit one round as 2*rem HALF-width
// tiles, items >= `full` are those halves (two adjacent items share the A tile through L2): 3.46 rounds become ~3.55
// instead of 4.  Returns the item count; `full` = first item index of the split region (== item count when unused).
__device__ __forceinline__ int tail_split(int wide_tiles, int grid, int bn_full, int N, int* full) {
    const int rem = wide_tiles % grid;
    if (bn_full == 256 && (N % 256) == 0 && rem > 0 && 2 * rem <= grid && wide_tiles > grid) {
        *full = wide_tiles - rem;
        return wide_tiles + rem;
    }
    *full = wide_tiles;
    return wide_tiles;
}

// Static tile order: n fastest, then m, then group.  `wide` tiles are 256 columns (last one 128 when N % 256 != 0).
__device__ __forceinline__ Tile decode_tile(int t, const Params& p, int m_tiles, int n_tiles, int bn_full, int full) {
    Tile r;
    int half = -1;
    if (t >= full) { const int u = t - full; half = u & 1; t = full + (u >> 1); }
    const int n_idx = t % n_tiles;
    const int rest = t / n_tiles;
    r.m_tile = rest % m_tiles;
    r.g = rest / m_tiles;
    r.n0 = n_idx * bn_full;
    r.bn = min(bn_full, p.N - r.n0);
    if (half >= 0) { r.n0 += half * 128; r.bn = 128; }
    return r;
}

// grid = min(#tiles, #SMs).  Load maps: box {64 (K), 128 (rows)}, SWIZZLE_128B.  Store maps (EPI_RELU_SPLIT):
// box {64 (cols), 32 (rows)}, SWIZZLE_128B.
// EPI8: warps 2..9 drain the accumulator, two per TMEM lane quadrant, each taking half of the tile's columns.  For the
// layers whose mainloop is short (conv2 / e_conv2 with K = 64, and every layer of the plain-bf16 training forward) the
// epilogue, not the MMA, is the critical path; the deep layers keep EPI8 = false (192 threads, warps 2..5, 4 ring stages:
// measured, K = 256 x 3 passes already loses more from the shorter ring than it gains).  Results are bit-identical: the
// arithmetic per element is the same.
constexpr int kThreadsEpi8 = 320;
template <bool EPI8>
__global__ void __launch_bounds__(EPI8 ? kThreadsEpi8 : kThreads, 1)
gemm_split_bf16_persistent_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                                  const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                                  const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                                  const Params p, const int bn_full)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    // EPI8 trades one ring stage for a private hi+lo staging pair per epilogue warp (3 x 48 KB + 8 x 8 KB = 208 KB)
    constexpr int NS = EPI8 ? 3 : kStages2;
    constexpr int kStagingK = EPI8 ? 8 * kStgWarp : kStaging;
    unsigned char* staging = smem + NS * kStage;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingK);
    uint64_t* empty_bar = full_bar + kStages2;
    uint64_t* tfull_bar = empty_bar + kStages2;       // [2]
    uint64_t* tempty_bar = tfull_bar + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = p.M / BM;
    const int n_tiles = (p.N + bn_full - 1) / bn_full;
    int full_tiles;
    const int total_tiles = tail_split(p.groups * m_tiles * n_tiles, (int)gridDim.x, bn_full, p.N, &full_tiles);
    const int kb_per_pass = p.K / BK;
    // which of the three split-bf16 products run (small terms first): bit 0 = A_lo*W_hi, bit 1 = A_hi*W_lo, bit 2 = A_hi*W_hi
    const int pass_mask = p.pass_mask ? (p.pass_mask & 7) : (p.passes == 1 ? 4 : 7);
    const int n_pass = __popc(pass_mask);
    const int iters_per_tile = n_pass * kb_per_pass;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        if (p.mode == EPI_RELU_SPLIT) { tma_prefetch_desc(&map_o_hi); tma_prefetch_desc(&map_o_lo); }
#pragma unroll
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], EPI8 ? 8 : 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols2);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: when this grid was launched with the stream-serialization attribute its CTAs may
    // start (barrier init, TMEM allocation, descriptor prefetch above) while the previous kernel on the stream is still
    // draining its last tiles; everything below reads or overwrites that kernel's buffers, so all threads wait here for its
    // completion (a no-op for a normal launch).  launch_dependents lets the NEXT kernel do the same with this one.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const Tile tl = decode_tile(t, p, m_tiles, n_tiles, bn_full, full_tiles);
                const int a_k = p.a_k0 + tl.g * p.a_kg;
                const int a_row = tl.m_tile * BM;
                const int w_row = tl.g * p.N + tl.n0;
                const uint32_t bytes = (uint32_t)(kStageA + tl.bn * BK * 2);
                for (int i = 0; i < iters_per_tile; ++i, ++it) {
                    const int s = it % NS;
                    const uint32_t ph = (uint32_t)(it / NS) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const int pass_i = i / kb_per_pass, kb = i - pass_i * kb_per_pass;
                    // pass 0: A_lo*W_hi, pass 1: A_hi*W_lo, pass 2: A_hi*W_hi (small terms first); plain bf16 = pass 2 alone
                    // = the pass_i-th set bit of pass_mask
                    int pass = 0;
                    for (int seen = 0; pass < 3; ++pass)
                        if ((pass_mask >> pass) & 1) { if (seen == pass_i) break; ++seen; }
                    const CUtensorMap* ma = (pass == 0) ? &map_a_lo : &map_a_hi;
                    const CUtensorMap* mw = (pass == 1) ? &map_w_lo : &map_w_hi;
                    unsigned char* sa = smem + s * kStage;
                    unsigned char* sb = sa + kStageA;
                    mbar_expect_tx(&full_bar[s], bytes);
                    tma_load_2d(sa, ma, &full_bar[s], a_k + kb * BK, a_row);
                    tma_load_2d(sb, mw, &full_bar[s], kb * BK, w_row);
                    if (tl.bn > 128) tma_load_2d(sb + 128 * BK * 2, mw, &full_bar[s], kb * BK, w_row + 128);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int it = 0, lt = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
                const Tile tl = decode_tile(t, p, m_tiles, n_tiles, bn_full, full_tiles);
                const int acc = lt & 1;
                const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
                mbar_wait(&tempty_bar[acc], aph ^ 1u);            // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t idesc = make_idesc_bf16(BM, tl.bn);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int i = 0; i < iters_per_tile; ++i, ++it) {
                    const int s = it % NS;
                    const uint32_t ph = (uint32_t)(it / NS) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + s * kStage));
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + s * kStage + kStageA));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);                   // frees the smem stage once these MMAs retire
                }
                umma_commit(&tfull_bar[acc]);                     // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int quad = warp & 3;
        const int half = EPI8 ? (warp - 2) >> 2 : 0;              // EPI8: which half of the tile's columns this warp drains
        unsigned char* stg = staging + (EPI8 ? warp - 2 : quad) * kStgWarp;   // [hi|lo][32 rows x 128 B], SWIZZLE_128B
        float* s_colsum = reinterpret_cast<float*>(staging);      // EPI_RELU_COLSUM: [4][256]
        const int GN = p.groups * p.N;
        int lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const Tile tl = decode_tile(t, p, m_tiles, n_tiles, bn_full, full_tiles);
            const int acc = lt & 1;
            const uint32_t aph = (uint32_t)(lt >> 1) & 1u;
            const int row0 = tl.m_tile * BM + quad * 32;
            const int row = row0 + lane;
            const int col_g = tl.g * p.N + tl.n0;                 // first column within [groups*N]
            const float* bias = p.bias + (p.bias_obj_rows > 0 ? (size_t)(row0 / p.bias_obj_rows) * (size_t)GN : 0) + col_g;
            const int c_beg = EPI8 ? half * (tl.bn >> 1) : 0;
            const int c_end = EPI8 ? c_beg + (tl.bn >> 1) : tl.bn;
            float bl = __ldg(bias + c_beg + lane);                // bias is prefetched one 32-column chunk ahead
            const bool valid = (p.mode != EPI_RELU_COLSUM) || ((row % p.rows_per_obj) < p.valid_rows);
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256);
            // EPI_HEAD_OUT: stage the last-layer weights of this tile's object class before waiting for the accumulator
            float* s_w4 = reinterpret_cast<float*>(stg);          // [4][128], rows >= n_out zero-filled
            const int hb_obj = row0 / p.rows_per_obj;
            const int n_out = tl.g == 0 ? 4 : (tl.g == 1 ? 3 : 1);
            int cls = 0;
            if (p.mode == EPI_HEAD_OUT) {
                if (hb_obj < p.batch) {
                    cls = (int)p.obj[hb_obj];
                    cls = cls < 0 ? 0 : (cls >= p.num_obj ? p.num_obj - 1 : cls);
                }
                const float* wsrc = p.w4[tl.g] + (size_t)cls * n_out * 128;
                __syncwarp();
                for (int i = lane; i < 4 * 128; i += 32) s_w4[i] = i < n_out * 128 ? __ldg(wsrc + i) : 0.0f;
                __syncwarp();
            }
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (p.mode == EPI_RELU_SPLIT) {
#pragma unroll 1
                for (int c0 = c_beg; c0 < c_end; c0 += 64) {
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + (uint32_t)(c0 + 32 * h), v);
                        const bool last = (h == 1) && (c0 + 64 >= c_end);
                        if (last) {                                   // last read of this accumulator: hand it back early
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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
                    unsigned char* sl = stg + kStgBuf;
                    const uint32_t sw = (uint32_t)lane & 7u;          // 128-byte swizzle: 16 B chunk ^= row % 8
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t off = (uint32_t)lane * 128u + ((uint32_t)j ^ sw) * 16u;
                        *reinterpret_cast<uint4*>(sh + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        if (!p.hi_only) *reinterpret_cast<uint4*>(sl + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int oc = p.o_c0 + col_g + c0;
                        tma_store_2d(&map_o_hi, sh, oc, row0);
                        if (!p.hi_only) tma_store_2d(&map_o_lo, sl, oc, row0);
                        bulk_commit();
                    }
                }
            } else if (p.mode == EPI_HEAD_OUT) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < tl.bn; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    const bool last = c0 + 32 >= tl.bn;
                    if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    const float bcur = bl;
                    if (!last) bl = __ldg(bias + c0 + 32 + lane);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float f[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) f[q] = fmaxf(__uint_as_float(v[j + q]) + __shfl_sync(0xffffffffu, bcur, j + q), 0.0f);
                        const float4 w0 = *reinterpret_cast<const float4*>(s_w4 + c0 + j);
                        const float4 w1 = *reinterpret_cast<const float4*>(s_w4 + 128 + c0 + j);
                        const float4 w2 = *reinterpret_cast<const float4*>(s_w4 + 256 + c0 + j);
                        const float4 w3 = *reinterpret_cast<const float4*>(s_w4 + 384 + c0 + j);
                        a0 = fmaf(f[3], w0.w, fmaf(f[2], w0.z, fmaf(f[1], w0.y, fmaf(f[0], w0.x, a0))));
                        a1 = fmaf(f[3], w1.w, fmaf(f[2], w1.z, fmaf(f[1], w1.y, fmaf(f[0], w1.x, a1))));
                        a2 = fmaf(f[3], w2.w, fmaf(f[2], w2.z, fmaf(f[1], w2.y, fmaf(f[0], w2.x, a2))));
                        a3 = fmaf(f[3], w3.w, fmaf(f[2], w3.z, fmaf(f[1], w3.y, fmaf(f[0], w3.x, a3))));
                    }
                }
                const int n = row % p.rows_per_obj;
                if (hb_obj < p.batch && n < p.valid_rows) {
                    const float* b4 = p.b4[tl.g] + cls * n_out;
                    const size_t pt = (size_t)hb_obj * p.valid_rows + n;
                    if (tl.g == 0) {
                        *reinterpret_cast<float4*>(p.pred[0] + pt * 4) = make_float4(a0 + __ldg(b4), a1 + __ldg(b4 + 1), a2 + __ldg(b4 + 2), a3 + __ldg(b4 + 3));
                    } else if (tl.g == 1) {
                        float* o = p.pred[1] + pt * 3;
                        o[0] = a0 + __ldg(b4); o[1] = a1 + __ldg(b4 + 1); o[2] = a2 + __ldg(b4 + 2);
                    } else {
                        p.pred[2][pt] = 1.0f / (1.0f + expf(-(a0 + __ldg(b4))));
                    }
                }
            } else {
#pragma unroll 1
                for (int c0 = c_beg; c0 < c_end; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)c0, v);
                    const bool last = c0 + 32 >= c_end;
                    if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    const float bcur = bl;
                    if (!last) bl = __ldg(bias + c0 + 32 + lane);
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x = fmaxf(__uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bcur, j), 0.0f);
                        f[j] = valid ? x : 0.0f;
                    }
                    if (p.relu_bits) {
                        uint32_t bits = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) bits |= (f[j] > 0.0f ? 1u : 0u) << j;
                        p.relu_bits[(size_t)row * (size_t)(GN >> 5) + (size_t)((col_g + c0) >> 5)] = bits;
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
            }
            if (p.mode == EPI_RELU_COLSUM) {
                if (EPI8) asm volatile("bar.sync 1, 256;" ::: "memory"); else
                asm volatile("bar.sync 1, 128;" ::: "memory");        // the epilogue warps only
                const int tt = threadIdx.x - 64;                       // 0..127 (0..255 with EPI8)
                for (int c = tt; c < tl.bn; c += (EPI8 ? 256 : 128)) {
                    // fixed order over the quadrants -> deterministic
                    const float sum = (s_colsum[c] + s_colsum[256 + c]) + (s_colsum[512 + c] + s_colsum[768 + c]);
                    p.colsum[(size_t)tl.m_tile * (size_t)GN + col_g + c] = sum;
                }
                if (EPI8) asm volatile("bar.sync 1, 256;" ::: "memory"); else
                asm volatile("bar.sync 1, 128;" ::: "memory");        // scratch is free for the next tile
            }
        }
        if (p.mode == EPI_RELU_SPLIT && lane == 0) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols2);
    }
}

}  // namespace tc2
}  // namespace ape
