// PoseRefineNet after the pooled trunk, in ONE launch (sm_100a): AvgPool1d finish -> conv1_{r,t} -> conv2_{r,t} ->
// conv3_{r,t} rows of the object's class (network.py:187-204).  These were pool_finish + two dense_swapped launches +
// refiner_out: four dependent launches of 4-10 us each per refine iteration with almost no work in them (64 objects).
//
// One thread-block CLUSTER per (branch r | t, chunk of 64 objects).  CL = 16 (or 8) CTAs split the K = 1024 of conv1:
//   1. every CTA: its K slice of the pooled feature from the per-tile column sums (-> split bf16, written 128-byte-swizzled
//      into shared memory as the MMA's N operand), its slice of the conv1 weights by TMA (issued BEFORE the programmatic
//      dependency wait: weights do not depend on the previous kernel), 4 M tiles x 3 split-bf16 products of tcgen05.mma
//      (128 outputs x 64 objects each) into 256 TMEM columns; the partial [64 x 512] fp32 tile goes to its own shared memory.
//   2. cluster barrier; CTA r sums outputs [r*512/CL, ...) of all objects over the CL partial tiles through distributed
//      shared memory in rank order (deterministic), + bias, ReLU: these are exactly ITS K slice of conv2's input.
//   3. conv2 partial products on that slice in fp32 SIMT (4 outputs x 8 objects per thread) -> partial [64 x 128] tile;
//      cluster barrier; the same split reduction (+ bias, ReLU) gives each CTA 128/CL inputs of conv3.
//   4. conv3 partials of the object's class rows (4 or 3 outputs) are PUSHED into rank 0's shared memory, cluster barrier,
//      rank 0 sums them in rank order, adds the bias and writes r2 / t2 (the other CTAs are free to exit).
// Nothing but r2 / t2 is written to global memory.  Phase times at 64 objects (clock64 stamps, cycles): pooled operand
// 1.6 k, MMAs 4.2 k (the conv2 weight slice is fetched under them), TMEM -> partial tile + barrier 2.5 k, conv1 reduction
// 8.1 k (DSMEM: 120 KB per CTA at ~17 B/cycle -- the floor of a split-K cluster), conv2 2.5 k, its reduction 2.5 k.
#pragma once
#include "gemm_tc2.cuh"
#include <cooperative_groups.h>

namespace ape {
namespace tail {

using namespace ape::tc;

constexpr int kThreadsT = 576;                 // warp 0: TMA, warp 1: MMA issue + TMEM, warps 2-17: workers; all: reductions
constexpr int kWorkers = kThreadsT - 64;
constexpr int kObj = 64;                       // objects per cluster = N of the MMAs
constexpr int kOffX = 128 * 1024;              // [0, 128 KB): conv1 weight tiles (hi | lo) x 4 M tiles, then the partial tile
constexpr int kW2Ld = 132, kH1Ld = 68;

template <int CL> struct Cfg {
    static constexpr int KS = 1024 / CL;       // conv1 K per CTA
    static constexpr int NST = KS / 64;        // sequential 64-wide stages through the same shared memory
    static constexpr int OW = 512 / CL;        // conv1 outputs (= conv2 inputs) finished per CTA
    static constexpr int OW2 = 128 / CL;       // conv2 outputs (= conv3 inputs) finished per CTA
    static constexpr int kH1Bytes = OW * kH1Ld * 4;
    static constexpr int kXBytes = kH1Bytes > 16384 ? kH1Bytes : 16384;      // operand tiles of the objects, then h1
    static constexpr int kOffP2 = kOffX + kXBytes;
    static constexpr int kP2Stride = kObj * OW2 + 8;                         // conv2 partials [owner CTA][object][its OW2 outputs] (+8: banks)
    static constexpr int kOffW2 = kOffP2 + CL * kP2Stride * 4;
    static constexpr int kOffH2 = kOffW2 + OW * kW2Ld * 4;
    static constexpr int kOffP3 = kOffH2 + kObj * OW2 * 4;                   // [CL][64 x 4]: every CTA PUSHES its conv3 partials to rank 0
    static constexpr int kOffBar = kOffP3 + CL * kObj * 4 * 4;
    static constexpr int kSmem = kOffBar + 64 + 1024;
};

struct TailParams {
    const float* cs; int tiles_per_obj; float n_points;      // per-tile column sums [B*tiles_per_obj, 1024], points per object
    const float *b1, *w2, *b2;                               // conv1 bias [1024], conv2 weights [256, 512] fp32, bias [256]
    const float *w3r, *b3r, *w3t, *b3t;                      // conv3_{r,t} [num_obj*4 | 3, 128]
    const int64_t* obj; int num_obj, batch;
    float *r2, *t2;                                          // [batch, 4], [batch, 3]
};

#ifdef APE_TAIL_TIMING
__device__ unsigned long long g_tail_dbg[64 * 16];
#define TSTAMP(i) do { if (threadIdx.x == 64) g_tail_dbg[((blockIdx.z * 2 + blockIdx.y) * CL + blockIdx.x) * 16 + (i)] = clock64(); } while (0)
#else
#define TSTAMP(i) do { } while (0)
#endif
template <int CL>
__global__ void __launch_bounds__(kThreadsT, 1)
refiner_tail_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const TailParams p)
{
    using C = Cfg<CL>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    float* s_part = reinterpret_cast<float*>(smem);                        // [64 objects][512 outputs]
    unsigned char* s_x = smem + kOffX;                                     // hi tile 8 KB | lo tile 8 KB
    float* s_h1 = reinterpret_cast<float*>(smem + kOffX);                  // [OW][kH1Ld]   (after the MMAs)
    float* s_p2 = reinterpret_cast<float*>(smem + C::kOffP2);              // [CL owners][64 objects][OW2 outputs]
    float* s_w2 = reinterpret_cast<float*>(smem + C::kOffW2);              // [OW k][kW2Ld]: conv2 weights, transposed slice
    float* s_h2 = reinterpret_cast<float*>(smem + C::kOffH2);              // [64][OW2]
    float* s_p3 = reinterpret_cast<float*>(smem + C::kOffP3);              // [CL][64][4]   (used on rank 0)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
    uint64_t* mma_bar = full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = blockIdx.x, branch = blockIdx.y, b0 = blockIdx.z * kObj;
    const int nb = min(kObj, p.batch - b0);
    const int nj = branch == 0 ? 4 : 3;

    TSTAMP(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        mbar_init(full_bar, 1); mbar_init(mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto load_weights = [&](int st) {                                      // 8 tiles of 128 outputs x 64 K
        mbar_expect_tx(full_bar, 8 * 16384);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
                tma_load_2d(smem + (h * 4 + mt) * 16384, h == 0 ? &map_w_hi : &map_w_lo, full_bar, rank * C::KS + st * 64,
                            branch * 512 + mt * 128);
    };
    // ---- independent of the previous kernel: the weights
    if (warp == 0 && lane == 0) load_weights(0);
    TSTAMP(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");                     // programmatic dependent launch: see ape_common.cuh
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    TSTAMP(2);

    // ---- conv1 partial products of this CTA's K slice
    for (int st = 0; st < C::NST; ++st) {
        if (st > 0) {                                                      // the MMAs of the previous stage have read W and X
            mbar_wait(mma_bar, (uint32_t)(st - 1) & 1u);
            tc_fence_after();
            if (warp == 0 && lane == 0) load_weights(st);
        }
        // AvgPool1d finish for 64 objects x 64 channels -> split bf16, K-major, 128-byte swizzle (16 B chunk ^= row % 8)
        const int k0 = rank * C::KS + st * 64;
        for (int i = tid; i < kObj * 8; i += kThreadsT) {
            const int r = i >> 3, c = i & 7;
            uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = make_uint4(0, 0, 0, 0);
            if (r < nb) {
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const float* src = p.cs + (size_t)(b0 + r) * p.tiles_per_obj * 1024 + k0 + c * 8;
#pragma unroll 4
                for (int t = 0; t < p.tiles_per_obj; ++t) {
                    const float4 a = *reinterpret_cast<const float4*>(src + (size_t)t * 1024);
                    const float4 b = *reinterpret_cast<const float4*>(src + (size_t)t * 1024 + 4);
                    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
                }
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x0 = v[2 * e] / p.n_points, x1 = v[2 * e + 1] / p.n_points;
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                    h[e] = pack_bf16x2(h0, h1);
                    l[e] = pack_bf16x2(__float2bfloat16_rn(x0 - __bfloat162float(h0)), __float2bfloat16_rn(x1 - __bfloat162float(h1)));
                }
                hi4 = make_uint4(h[0], h[1], h[2], h[3]); lo4 = make_uint4(l[0], l[1], l[2], l[3]);
            }
            const int off = r * 128 + ((c ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(s_x + off) = hi4;
            *reinterpret_cast<uint4*>(s_x + 8192 + off) = lo4;
        }
        fence_proxy_async();                                               // generic-proxy writes -> visible to the MMA's async proxy
        __syncthreads();
        TSTAMP(3);
        if (warp >= 2 && st == C::NST - 1) {                               // under the MMAs: conv2 weight slice, transposed s_w2[k][o2]
            constexpr int kV = C::OW / 4;
            for (int i = tid - 64; i < 128 * kV; i += kWorkers) {
                const int o2 = i / kV, c = i % kV;
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.w2 + (size_t)(branch * 128 + o2) * 512 + rank * C::OW) + c);
                s_w2[(4 * c) * kW2Ld + o2] = v.x; s_w2[(4 * c + 1) * kW2Ld + o2] = v.y;
                s_w2[(4 * c + 2) * kW2Ld + o2] = v.z; s_w2[(4 * c + 3) * kW2Ld + o2] = v.w;
            }
        }
        if (warp == 1 && lane == 0) {
            mbar_wait(full_bar, (uint32_t)st & 1u);
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(128, kObj);
            const uint64_t x_hi = make_smem_desc_sw128(smem_u32(s_x)), x_lo = make_smem_desc_sw128(smem_u32(s_x + 8192));
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const uint64_t w_hi = make_smem_desc_sw128(smem_u32(smem + mt * 16384));
                const uint64_t w_lo = make_smem_desc_sw128(smem_u32(smem + (4 + mt) * 16384));
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {                     // W_hi * x_lo, W_lo * x_hi, W_hi * x_hi (small terms first)
                    const uint64_t a = pass == 1 ? w_lo : w_hi, b = pass == 0 ? x_lo : x_hi;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(tmem_base + (uint32_t)(mt * kObj), a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), idesc,
                                  (st > 0 || pass > 0 || k > 0) ? 1u : 0u);
                }
            }
            umma_commit(mma_bar);
        }
    }
    mbar_wait(mma_bar, (uint32_t)(C::NST - 1) & 1u);                       // all MMAs retired: W and X shared memory is free
    tc_fence_after();
    TSTAMP(4);
    if (warp >= 2) {                                                       // TMEM lane = output row -> partial tile, object-major
        const int quad = warp & 3, mt = (warp - 2) >> 2;
#pragma unroll
        for (int c0 = 0; c0 < kObj; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * kObj + c0), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) s_part[(c0 + j) * 512 + mt * 128 + quad * 32 + lane] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    TSTAMP(5);
    cluster.sync();                                                        // (1) every CTA's conv1 partial tile is in place

    TSTAMP(6);
    // ---- conv1 finish for this CTA's OW outputs (all 64 objects): + bias, ReLU -> its K slice of conv2's input
    {
        const float* peer[CL];
#pragma unroll
        for (int i = 0; i < CL; ++i) peer[i] = cluster.map_shared_rank(s_part, i);
        for (int idx = tid; idx < kObj * C::OW; idx += kThreadsT) {
            const int b = idx / C::OW, ol = idx % C::OW, o = rank * C::OW + ol;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < CL; ++i) acc += peer[i][b * 512 + o];     // rank order: deterministic
            acc += __ldg(p.b1 + branch * 512 + o);
            s_h1[ol * kH1Ld + b] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    TSTAMP(7);
    // ---- conv2 partial products over this K slice: thread = 4 outputs x 8 objects on eight warps (measured: 2 x 8 and 4 x 4
    // on sixteen warps are bound by shared-memory instruction issue, 8 x 8 on four warps by FMA latency: all 1.5x slower)
    if (warp >= 2 && warp < 10) {
        const int tx = lane, ty = warp - 2;
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll 4
        for (int k = 0; k < C::OW; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(s_w2 + k * kW2Ld + tx * 4);
            const float4 ha = *reinterpret_cast<const float4*>(s_h1 + k * kH1Ld + ty * 8);
            const float4 hb = *reinterpret_cast<const float4*>(s_h1 + k * kH1Ld + ty * 8 + 4);
            const float h[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[j][0] = fmaf(w.x, h[j], acc[j][0]); acc[j][1] = fmaf(w.y, h[j], acc[j][1]);
                acc[j][2] = fmaf(w.z, h[j], acc[j][2]); acc[j][3] = fmaf(w.w, h[j], acc[j][3]);
            }
        }
        const int o2 = tx * 4;
        float* dst = s_p2 + (o2 / C::OW2) * C::kP2Stride + (o2 % C::OW2);       // the slice its owner CTA will read, contiguous
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + (ty * 8 + j) * C::OW2) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
    TSTAMP(8);
    // conv3 rows of each object's class: this CTA's OW2 columns, fetched under the barrier, used after the conv2 reduction
    int cls = 0;
    float w3reg[C::OW2];
#pragma unroll
    for (int ol = 0; ol < C::OW2; ++ol) w3reg[ol] = 0.f;
    if (tid < kObj * 4 && (tid >> 2) < nb && (tid & 3) < nj) {
        cls = (int)p.obj[b0 + (tid >> 2)];
        cls = cls < 0 ? 0 : (cls >= p.num_obj ? p.num_obj - 1 : cls);
        const float* w = (branch == 0 ? p.w3r : p.w3t) + (size_t)(cls * nj + (tid & 3)) * 128 + rank * C::OW2;
#pragma unroll
        for (int ol = 0; ol < C::OW2; ++ol) w3reg[ol] = __ldg(w + ol);
    }
    cluster.sync();                                                        // (2) conv2 partial tiles in place
    TSTAMP(9);

    // ---- conv2 finish for this CTA's OW2 outputs, then conv3 partials of each object's class rows
    {
        const float* peer[CL];
#pragma unroll
        for (int i = 0; i < CL; ++i) peer[i] = cluster.map_shared_rank(s_p2, i);
        for (int idx = tid; idx < kObj * C::OW2; idx += kThreadsT) {       // idx = object * OW2 + local output
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < CL; ++i) acc += peer[i][rank * C::kP2Stride + idx];
            acc += __ldg(p.b2 + branch * 128 + rank * C::OW2 + idx % C::OW2);
            s_h2[idx] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    TSTAMP(10);
    if (tid < kObj * 4) {                                                  // conv3 partials -> rank 0's shared memory, slot [rank]
        const int b = tid >> 2;
        float acc = 0.f;
#pragma unroll
        for (int ol = 0; ol < C::OW2; ++ol) acc = fmaf(w3reg[ol], s_h2[b * C::OW2 + ol], acc);
        cluster.map_shared_rank(s_p3, 0)[rank * (kObj * 4) + tid] = acc;
    }
    TSTAMP(11);
    cluster.sync();                                                        // (3) every CTA's conv3 partials have landed on rank 0
    TSTAMP(12);
    if (rank == 0 && tid < kObj * 4) {
        const int b = tid >> 2, j = tid & 3;
        if (b < nb && j < nj) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < CL; ++i) acc += s_p3[i * (kObj * 4) + tid];
            if (branch == 0) p.r2[(size_t)(b0 + b) * 4 + j] = acc + __ldg(p.b3r + cls * 4 + j);
            else p.t2[(size_t)(b0 + b) * 3 + j] = acc + __ldg(p.b3t + cls * 3 + j);
        }
    }
    TSTAMP(13);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tail
}  // namespace ape
