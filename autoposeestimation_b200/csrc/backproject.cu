// Depth -> point cloud back-projection kernels (sm_100a).
//   ape_backproject_choose   : gather at fixed sampling indices, fp32 bit-exact with
//                              pipeline/utils.py:542-553
//   ape_surface_backproject  : masked scan with ordered stream compaction + fp64 rigid
//                              transform, replaces pc_reconstruction/open3d_utils.py:172-192
// Both are HBM-bound integer/byte scans; the second is the 921 600 B/frame kernel whose
// roofline is reported.
#include "ape_common.cuh"

namespace ape {

// ---------------------------------------------------------------------------------- a3
__global__ void __launch_bounds__(256)
backproject_choose_kernel(const uint16_t* __restrict__ depth, int n_frames, int H, int W,
                          const int32_t* __restrict__ frame_of, const int32_t* __restrict__ bbox,
                          const int64_t* __restrict__ choose, const float* __restrict__ cam,
                          int n_points, float* __restrict__ cloud)
{
    const int b = blockIdx.y;
    const int f = frame_of ? frame_of[b] : b;
    const int rmin = bbox[4 * b + 0], cmin = bbox[4 * b + 2], cmax = bbox[4 * b + 3];
    const int cw = cmax - cmin;
    const float ppx = cam[5 * b + 0], ppy = cam[5 * b + 1], fx = cam[5 * b + 2], fy = cam[5 * b + 3],
                scale = cam[5 * b + 4];
    const uint16_t* d = depth + (size_t)f * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += gridDim.x * blockDim.x) {
        const int64_t c = choose[(size_t)b * n_points + i];
        const int r = rmin + (int)(c / cw);
        const int col = cmin + (int)(c % cw);
        float z = 0.f, x = 0.f, y = 0.f;
        if (r >= 0 && r < H && col >= 0 && col < W) {
            // numpy order (pipeline/utils.py:549-551): pt2 = d*scale; pt0 = ((col-ppx)*pt2)/fx;
            // pt1 = ((row-ppy)*pt2)/fy -- every op rounded to fp32, no contraction.
            z = __fmul_rn((float)d[r * W + col], scale);
            x = __fdiv_rn(__fmul_rn(__fsub_rn((float)col, ppx), z), fx);
            y = __fdiv_rn(__fmul_rn(__fsub_rn((float)r, ppy), z), fy);
        }
        float* o = cloud + ((size_t)b * n_points + i) * 3;
        o[0] = x; o[1] = y; o[2] = z;
    }
}

// ---------------------------------------------------------------------------------- a4
// Three kernels, no inter-CTA waiting:
//   surface_mask_kernel : pure stream over label + depth (the 921 600 B/frame of SURVEY 8d).  Each thread owns 32
//                         consecutive pixels (2 x 16 B of label, 4 x 16 B of depth, streaming loads), builds a 32-bit
//                         validity mask and writes it (4 B per 96 B read); each WARP writes the pop-count of its
//                         1024-pixel span ("sub-chunk").  5.9 TB/s = 89 % of the measured HBM peak.
//   surface_scan_kernel : one warp per view: exclusive prefix of the view's sub-chunk counts (= first output slot of
//                         every sub-chunk), the view total, and a global list of the NON-EMPTY sub-chunks (3 of 4 are
//                         empty: the object covers a fraction of the frame).
//   surface_emit_kernel : persistent warps over that list, one task = one non-empty sub-chunk, no block-level state,
//                         no compaction list: lane l converts pixel 32 w + l of word w when its bit is set and writes
//                         it to slot = first slot + (valid pixels before word w) + (set bits below l) -- the row-major
//                         order of np.where.  Depth is re-read only for non-empty 32-pixel words.
// History (profiles/): one CTA per view walking the frame = 1.35 TB/s (latency-bound); single-pass chunked kernels
// with a decoupled look-back = 1.5 TB/s, and 0.16 TB/s when made persistent (convoy on the look-back flags); mask +
// one emit CTA per 8192-pixel chunk with a shared-memory compaction list (3 of 4 CTAs empty, two block barriers, two
// fp64 divisions and one integer division per pixel: 69 us of emit for 512 frames) = 3.5 TB/s; this version (40 us of
// emit: the prologue of the empty CTAs / warps was 40 % of all stall samples) = 4.7 TB/s for the whole call, 4.8 TB/s with
// programmatic dependent launch between the three kernels.
constexpr int kSurfThreads = 256;
constexpr int kSurfPix = 32;                          // pixels per thread
constexpr int kSurfSub = 32 * kSurfPix;               // 1024 pixels per warp = one sub-chunk
constexpr int kSurfChunk = kSurfThreads * kSurfPix;   // 8192 pixels = 24 KB of input per mask CTA
constexpr int kSurfWarps = kSurfThreads / 32;

struct SurfRegs { uint4 l0, l1, d0, d1, d2, d3; };

// ragged tail (a thread span that crosses the end of the frame): scalar, zero-filled; kept out of line so that its
// byte arrays do not inflate the register budget of the streaming path
__device__ __noinline__ void surf_load_tail(SurfRegs* r, const uint8_t* lab, const uint16_t* dep, int p, int npix) {
    uint32_t lw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dw[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 32 && p + i < npix; ++i) {
        lw[i >> 2] |= (uint32_t)lab[p + i] << (8 * (i & 3));
        dw[i >> 1] |= (uint32_t)dep[p + i] << (16 * (i & 1));
    }
    r->l0 = make_uint4(lw[0], lw[1], lw[2], lw[3]);   r->l1 = make_uint4(lw[4], lw[5], lw[6], lw[7]);
    r->d0 = make_uint4(dw[0], dw[1], dw[2], dw[3]);   r->d1 = make_uint4(dw[4], dw[5], dw[6], dw[7]);
    r->d2 = make_uint4(dw[8], dw[9], dw[10], dw[11]); r->d3 = make_uint4(dw[12], dw[13], dw[14], dw[15]);
}
__device__ __forceinline__ void surf_load(SurfRegs& r, const uint8_t* lab, const uint16_t* dep, int p, int npix) {
    if (p + kSurfPix <= npix) {
        r.l0 = ld_stream16(lab + p);      r.l1 = ld_stream16(lab + p + 16);
        r.d0 = ld_stream16(dep + p);      r.d1 = ld_stream16(dep + p + 8);
        r.d2 = ld_stream16(dep + p + 16); r.d3 = ld_stream16(dep + p + 24);
    } else if (p >= npix) {               // past the end of the frame (last chunk)
        r.l0 = r.l1 = r.d0 = r.d1 = r.d2 = r.d3 = make_uint4(0u, 0u, 0u, 0u);
    } else {
        surf_load_tail(&r, lab, dep, p, npix);
    }
}

// 4 label bytes -> 4 bits (bit i set when byte i matches: == want, or != 0 when want == 0).  Exact zero-byte test on
// x = w ^ want4: ((x & 0x7f..) + 0x7f..) | x has the top bit of a byte clear only when the byte is zero (no carries
// between bytes); plain integer ops instead of the emulated SIMD-video intrinsics.
__device__ __forceinline__ uint32_t label_bits(uint32_t w, uint32_t want4) {
    const uint32_t x = w ^ want4;
    const uint32_t nz = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;      // 0x80 per NON-zero byte of x
    const uint32_t hit = want4 ? (nz ^ 0x80808080u) : nz;                          // want == 0: label != 0
    return (hit * 0x00204081u) >> 28;                                              // top bits of the 4 bytes -> bits 0..3
}
__device__ __forceinline__ uint32_t depth_bits(uint32_t w) {
    const uint32_t nz = (((w & 0x7fff7fffu) + 0x7fff7fffu) | w) & 0x80008000u;      // 0x8000 per non-zero half
    return ((nz >> 15) | (nz >> 30)) & 3u;
}

// work layout: sub_count [n_views * n_chunks * 8] int32 (valid pixels per 1024-pixel span), then
//              masks [n_views * n_chunks * 256] u32 (one bit per pixel)
__global__ void __launch_bounds__(kSurfThreads, 4)
surface_mask_kernel(const uint8_t* __restrict__ label, const uint16_t* __restrict__ depth, int npix,
                    const int32_t* __restrict__ frame_of, const uint8_t* __restrict__ label_value,
                    int32_t* __restrict__ sub_count, uint32_t* __restrict__ masks, int n_chunks, int32_t* __restrict__ n_tasks)
{
    pdl_sync();                                           // programmatic dependent launch: see ape_common.cuh
    const int tile = blockIdx.x;
    const int v = tile / n_chunks, c = tile - v * n_chunks;
    const int f = frame_of ? frame_of[v] : v;
    const uint32_t want4 = (label_value ? (uint32_t)label_value[v] : 0u) * 0x01010101u;
    SurfRegs cur;
    surf_load(cur, label + (size_t)f * npix, depth + (size_t)f * npix, c * kSurfChunk + threadIdx.x * kSurfPix, npix);
    const uint32_t lw[8] = {cur.l0.x, cur.l0.y, cur.l0.z, cur.l0.w, cur.l1.x, cur.l1.y, cur.l1.z, cur.l1.w};
    const uint32_t dw[16] = {cur.d0.x, cur.d0.y, cur.d0.z, cur.d0.w, cur.d1.x, cur.d1.y, cur.d1.z, cur.d1.w,
                             cur.d2.x, cur.d2.y, cur.d2.z, cur.d2.w, cur.d3.x, cur.d3.y, cur.d3.z, cur.d3.w};
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) mask |= label_bits(lw[i], want4) << (4 * i);
    uint32_t dmask = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) dmask |= depth_bits(dw[i]) << (2 * i);
    mask &= dmask;
    masks[(size_t)tile * kSurfThreads + threadIdx.x] = mask;
    if (tile == 0 && threadIdx.x == 0) *n_tasks = 0;        // consumed by surface_scan_kernel (next launch on the stream)
    const int cnt = warp_sum(__popc(mask));
    if ((threadIdx.x & 31) == 0) sub_count[(size_t)tile * kSurfWarps + (threadIdx.x >> 5)] = cnt;
}

// Multi-label variant (BASELINE config 4: every frame carries the labels of all its objects): ONE pass over label + depth
// of a frame produces the validity words of all L label values, so the 921 600 B are read once per FRAME instead of once
// per (frame, object).  View v = frame * L + l.  Same work layout as above.
struct LabelSet { int n; uint8_t value[8]; };
__global__ void __launch_bounds__(kSurfThreads, 4)
surface_mask_multi_kernel(const uint8_t* __restrict__ label, const uint16_t* __restrict__ depth, int npix, LabelSet labels,
                          int32_t* __restrict__ sub_count, uint32_t* __restrict__ masks, int n_chunks, int32_t* __restrict__ n_tasks)
{
    pdl_sync();
    const int tile = blockIdx.x;                            // (frame, chunk)
    const int f = tile / n_chunks, c = tile - f * n_chunks;
    SurfRegs cur;
    surf_load(cur, label + (size_t)f * npix, depth + (size_t)f * npix, c * kSurfChunk + threadIdx.x * kSurfPix, npix);
    const uint32_t lw[8] = {cur.l0.x, cur.l0.y, cur.l0.z, cur.l0.w, cur.l1.x, cur.l1.y, cur.l1.z, cur.l1.w};
    const uint32_t dw[16] = {cur.d0.x, cur.d0.y, cur.d0.z, cur.d0.w, cur.d1.x, cur.d1.y, cur.d1.z, cur.d1.w,
                             cur.d2.x, cur.d2.y, cur.d2.z, cur.d2.w, cur.d3.x, cur.d3.y, cur.d3.z, cur.d3.w};
    uint32_t dmask = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) dmask |= depth_bits(dw[i]) << (2 * i);
    if (tile == 0 && threadIdx.x == 0) *n_tasks = 0;
    for (int l = 0; l < labels.n; ++l) {
        const uint32_t want4 = (uint32_t)labels.value[l] * 0x01010101u;
        uint32_t mask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) mask |= label_bits(lw[i], want4) << (4 * i);
        mask &= dmask;
        const size_t vt = ((size_t)f * labels.n + l) * n_chunks + c;          // tile index of view (f, l)
        masks[vt * kSurfThreads + threadIdx.x] = mask;
        const int cnt = warp_sum(__popc(mask));
        if ((threadIdx.x & 31) == 0) sub_count[vt * kSurfWarps + (threadIdx.x >> 5)] = cnt;
    }
}

// Exclusive prefix of the per-view totals -> first output slot of every view in a PACKED ragged cloud (one CTA).
__global__ void __launch_bounds__(1024)
view_offsets_kernel(const int32_t* __restrict__ counts, int n_views, int32_t* __restrict__ offsets /* [n_views + 1] */)
{
    pdl_sync();
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_views; base += 1024) {
        const int i = base + threadIdx.x;
        const int x = i < n_views ? counts[i] : 0;
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (i < n_views) offsets[i] = before + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_views] = s_carry;
}

// One warp per view: exclusive prefix of the view's sub-chunk counts -> the view total, and one task record
// {view, sub-chunk, first output slot, valid pixels} per NON-EMPTY sub-chunk appended to a global list (the order of the
// list does not matter: every task knows its output slots).  work[0] (task counter) is zeroed by the mask kernel.
__global__ void __launch_bounds__(256)
surface_scan_kernel(const int32_t* __restrict__ sub_count, int32_t* __restrict__ counts, int32_t* __restrict__ n_tasks,
                    int4* __restrict__ tasks, int n_sub, int n_views)
{
    pdl_sync();                                           // programmatic dependent launch: see ape_common.cuh
    __shared__ int s_ne[8], s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + warp;
    const bool live = v < n_views;                          // (whole warps; dead warps still meet the block barriers)
    const int32_t* c = sub_count + (size_t)(live ? v : 0) * n_sub;
    // all counts of the view are fetched up front (independent loads: one memory round trip); frames of more than
    // 32 * kScanRegs sub-chunks (512 Ki pixels) read the rest inside the loops
    constexpr int kScanRegs = 16;
    int xr[kScanRegs];
#pragma unroll
    for (int k = 0; k < kScanRegs; ++k) xr[k] = (live && 32 * k + lane < n_sub) ? c[32 * k + lane] : 0;
    // number of non-empty sub-chunks per view -> ONE atomic per CTA reserves the task slots of its 8 views
    // (same-address atomics with a return value serialise: one per view cost 8 us for 512 views)
    int ne = 0;
#pragma unroll
    for (int k = 0; k < kScanRegs; ++k) ne += xr[k] != 0;
    if (live) for (int j = 32 * kScanRegs + lane; j < n_sub; j += 32) ne += c[j] != 0;
    ne = warp_sum(ne);
    if (lane == 0) s_ne[warp] = ne;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_ne[w];
        s_base = tot ? atomicAdd(n_tasks, tot) : 0;
    }
    __syncthreads();
    if (!live) return;
    int tbase = s_base;
    for (int w = 0; w < warp; ++w) tbase += s_ne[w];
    int carry = 0;
    auto round = [&](int j, int x) {
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t nz = __ballot_sync(0xffffffffu, x != 0);
        if (x) tasks[tbase + __popc(nz & ((1u << lane) - 1u))] = make_int4(v, j, carry + incl - x, x);
        tbase += __popc(nz);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    };
#pragma unroll
    for (int k = 0; k < kScanRegs; ++k)
        if (32 * k < n_sub) round(32 * k + lane, xr[k]);
    for (int j0 = 32 * kScanRegs; j0 < n_sub; j0 += 32) round(j0 + lane, j0 + lane < n_sub ? c[j0 + lane] : 0);
    if (lane == 0) counts[v] = carry;
}

// Persistent warps over the task list (one task = one non-empty 1024-pixel sub-chunk), no block-level state.  Two
// dependent memory round trips per task: (task record + mask word), then the depth of the non-empty 32-pixel words
// (64 B per lane, all in flight together, parked in shared memory); the conversion loop itself touches no global memory
// except its stores.
// Word by word (non-empty words only, warp-uniform): lane l converts pixel 32 w + l when its bit is set and writes it
// to slot = offset + pre(w) + (set bits below l) -- the row-major order of np.where without a compaction list.  Object
// interiors are dense, so most words keep all 32 lanes busy.
constexpr int kEmitWarps = 8;
__global__ void __launch_bounds__(kEmitWarps * 32)
surface_emit_kernel(const uint16_t* __restrict__ depth, int npix, int W, const int32_t* __restrict__ frame_of,
                    const double* __restrict__ cam, const double* __restrict__ robot2cam, int capacity,
                    double* __restrict__ points, int32_t* __restrict__ pixel_index,
                    const int32_t* __restrict__ n_tasks_p, const int4* __restrict__ tasks, const uint32_t* __restrict__ masks, int n_sub,
                    const int32_t* __restrict__ view_base /* packed output: first slot of every view (capacity = total), or NULL */,
                    int views_per_frame /* > 0: view v reads frame v / views_per_frame, cam / robot2cam are per FRAME */)
{
    pdl_sync();                                           // programmatic dependent launch: see ape_common.cuh
    __shared__ __align__(16) uint16_t s_dep[kEmitWarps][kSurfSub];
    __shared__ double s_parw[kEmitWarps][16];              // robot2cam rows 0..2 (12) + ppx, ppy, fx, fy
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tasks = *n_tasks_p;
    const uint32_t w_inv = (uint32_t)(0x100000000ull / (uint32_t)W);   // floor(2^32 / W): row estimate is exact or one short
    const uint32_t lt_mask = (1u << lane) - 1u;
    const double* par = s_parw[warp];
    for (int t = blockIdx.x * kEmitWarps + warp; t < n_tasks; t += gridDim.x * kEmitWarps) {
        const int4 task = tasks[t];
        const int v = task.x, j = task.y;
        const int off = task.z + (view_base ? view_base[v] : 0);
        if (off >= capacity) continue;                      // warp-uniform
        const uint32_t mk = masks[((size_t)v * n_sub + j) * 32 + lane];   // lane l: validity of pixels [32 l, 32 l + 32)
        const int f = views_per_frame > 0 ? v / views_per_frame : (frame_of ? frame_of[v] : v);
        const int pv = views_per_frame > 0 ? f : v;         // index of the view's camera parameters
        const uint16_t* dep = depth + (size_t)f * npix;
        const int p0 = j * kSurfSub;
        __syncwarp();                                       // previous task's readers of s_dep / s_parw are done
        if (mk) {
            const int p = p0 + 32 * lane;
            uint4* dst = reinterpret_cast<uint4*>(&s_dep[warp][32 * lane]);
            if (p + 32 <= npix) {
                const uint4 a = ld_stream16(dep + p), b = ld_stream16(dep + p + 8), c = ld_stream16(dep + p + 16), d = ld_stream16(dep + p + 24);
                dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
            } else {
                for (int i = 0; i < 32; ++i) s_dep[warp][32 * lane + i] = p + i < npix ? dep[p + i] : (uint16_t)0;
            }
        }
        if (lane < 16) s_parw[warp][lane] = lane < 12 ? robot2cam[16 * (size_t)pv + lane] : cam[4 * (size_t)pv + (lane - 12)];
        __syncwarp();
        const double ppx = par[12], ppy = par[13], fx = par[14], fy = par[15];
        const double rfx = 1.0 / fx, rfy = 1.0 / fy;
        double* out = view_base ? points : points + (size_t)v * capacity * 3;
        int32_t* opix = pixel_index ? (view_base ? pixel_index : pixel_index + (size_t)v * capacity) : nullptr;
        const int pc = __popc(mk);
        int incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const int pre = incl - pc;                          // valid pixels of the span before word `lane`
        uint32_t words = __ballot_sync(0xffffffffu, pc != 0);
        while (words) {
            const int w = __ffs(words) - 1;
            words &= words - 1;
            const uint32_t wd = __shfl_sync(0xffffffffu, mk, w);
            const int slot = off + __shfl_sync(0xffffffffu, pre, w) + __popc(wd & lt_mask);
            if (!((wd >> lane) & 1u) || slot >= capacity) continue;
            const int pix = p0 + 32 * w + lane;
            const double z = (double)s_dep[warp][32 * w + lane];
            uint32_t r = __umulhi((uint32_t)pix, w_inv);
            uint32_t col = (uint32_t)pix - r * (uint32_t)W;
            if (col >= (uint32_t)W) { col -= (uint32_t)W; ++r; }
            // open3d_utils.py:185-189: p0 = ((px-ppx)*z)/fx, p1 = ((py-ppy)*z)/fy (fp64, rounded per op).  The two
            // divisions by the per-view constants are reciprocal + two fma residual corrections (the second one starts
            // from a faithful quotient, so the result is the correctly rounded quotient)
            const double ax = __dmul_rn(__dsub_rn((double)col, ppx), z);
            const double ay = __dmul_rn(__dsub_rn((double)r, ppy), z);
            double x = ax * rfx, y = ay * rfy;
            x = fma(fma(-x, fx, ax), rfx, x); y = fma(fma(-y, fy, ay), rfy, y);
            x = fma(fma(-x, fx, ax), rfx, x); y = fma(fma(-y, fy, ay), rfy, y);
            // robot2cam * [x y z 1]^T (:190-192)
            double* o = out + (size_t)slot * 3;
            o[0] = fma(par[0], x, fma(par[1], y, fma(par[2], z, par[3])));
            o[1] = fma(par[4], x, fma(par[5], y, fma(par[6], z, par[7])));
            o[2] = fma(par[8], x, fma(par[9], y, fma(par[10], z, par[11])));
            if (opix) opix[slot] = pix;
        }
    }
}

}  // namespace ape

extern "C" __attribute__((visibility("default"))) int ape_backproject_choose(const uint16_t* depth, int n_frames, int height, int width,
                                      const int32_t* frame_of, const int32_t* bbox, const int64_t* choose,
                                      const float* cam, int n_obj, int n_points, float* cloud, void* stream)
{
    APE_REQUIRE(depth && bbox && choose && cam && cloud, "ape_backproject_choose: null pointer");
    APE_REQUIRE(n_frames > 0 && height > 0 && width > 0 && n_obj >= 0 && n_points >= 0,
                "ape_backproject_choose: bad sizes");
    if (n_obj == 0 || n_points == 0) return APE_OK;
    APE_REQUIRE(n_obj <= 65535, "ape_backproject_choose: n_obj > 65535 (split the batch)");
    dim3 grid((n_points + 255) / 256, n_obj);
    ape::ProfScope prof_("backproject_choose", (cudaStream_t)stream);
    ape::backproject_choose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth, n_frames, height, width, frame_of,
                                                                          bbox, choose, cam, n_points, cloud);
    ape::count_launch();
    return ape::check_launch("ape_backproject_choose");
}

extern "C" __attribute__((visibility("default"))) size_t ape_surface_work_bytes(int n_views, int height, int width)
{
    const size_t chunks = ((size_t)(height > 0 ? height : 0) * (size_t)(width > 0 ? width : 0) + ape::kSurfChunk - 1) / ape::kSurfChunk;
    const size_t tiles = (size_t)(n_views > 0 ? n_views : 0) * chunks;
    // task counter (16 B) + one int4 task record and one count per 1024-pixel span + one mask bit per pixel
    return 16 + (16 + 4) * tiles * ape::kSurfWarps + 4 * tiles * ape::kSurfThreads;
}

extern "C" __attribute__((visibility("default"))) int ape_surface_backproject(const uint8_t* label, const uint16_t* depth, int n_frames, int height,
                                       int width, const int32_t* frame_of, const uint8_t* label_value,
                                       const double* cam, const double* robot2cam, int n_views, int capacity,
                                       double* points, int32_t* pixel_index, int32_t* counts, void* work, void* stream)
{
    APE_REQUIRE(label && depth && cam && robot2cam && points && counts && work, "ape_surface_backproject: null pointer");
    APE_REQUIRE(n_frames > 0 && height > 0 && width > 0 && n_views >= 0 && capacity > 0,
                "ape_surface_backproject: bad sizes");
    APE_REQUIRE(((size_t)height * width) % 16 == 0, "ape_surface_backproject: height*width must be a multiple of 16");
    APE_REQUIRE((size_t)height * width < (1u << 30), "ape_surface_backproject: frame too large");
    APE_REQUIRE((((uintptr_t)label) & 15) == 0 && (((uintptr_t)depth) & 15) == 0,
                "ape_surface_backproject: label/depth must be 16-byte aligned");
    APE_REQUIRE((((uintptr_t)work) & 15) == 0, "ape_surface_backproject: work must be 16-byte aligned");
    if (n_views == 0) return APE_OK;
    const int n_chunks = (int)(((size_t)height * width + ape::kSurfChunk - 1) / ape::kSurfChunk);
    APE_REQUIRE((size_t)n_views * n_chunks * ape::kSurfWarps < (1u << 31), "ape_surface_backproject: too many views (split the batch)");
    cudaStream_t s = (cudaStream_t)stream;
    const int n_tiles = n_views * n_chunks;
    // work: [task counter + 3 pad words | task records int4 | sub-chunk counts | mask words]
    const int n_sub = n_chunks * ape::kSurfWarps;            // 1024-pixel spans per view (the last ones may be empty padding)
    const size_t n_span = (size_t)n_views * n_sub;
    int32_t* n_tasks = reinterpret_cast<int32_t*>(work);
    int4* tasks = reinterpret_cast<int4*>(work) + 1;
    int32_t* sub_count = reinterpret_cast<int32_t*>(tasks + n_span);
    uint32_t* masks = reinterpret_cast<uint32_t*>(sub_count + n_span);
    {
        ape::ProfScope prof_("surface_mask", s);
        APE_CUDA(ape::launch_pdl(ape::surface_mask_kernel, dim3(n_tiles), dim3(ape::kSurfThreads), 0, s, label, depth, height * width, frame_of,
                                 label_value, sub_count, masks, n_chunks, n_tasks));
        ape::count_launch();
    }
    int rc = ape::check_launch("ape_surface_backproject (mask)");
    if (rc) return rc;
    {
        ape::ProfScope prof_("surface_scan", s);
        APE_CUDA(ape::launch_pdl(ape::surface_scan_kernel, dim3((n_views + 7) / 8), dim3(256), 0, s, sub_count, counts, n_tasks, tasks, n_sub, n_views));
        ape::count_launch();
    }
    ape::ProfScope prof_("surface_emit", s);
    const size_t want_ctas = (n_span + ape::kEmitWarps - 1) / ape::kEmitWarps;      // at most one task per warp is ever needed
    const size_t max_ctas = (size_t)ape::sm_count() * 8;
    APE_CUDA(ape::launch_pdl(ape::surface_emit_kernel, dim3((unsigned)(want_ctas < max_ctas ? want_ctas : max_ctas)), dim3(ape::kEmitWarps * 32), 0, s,
                             depth, height * width, width, frame_of, cam, robot2cam, capacity, points, pixel_index, n_tasks, tasks, masks, n_sub,
                             (const int32_t*)nullptr, 0));
    ape::count_launch();
    return ape::check_launch("ape_surface_backproject (emit)");
}

// a4 for frames that carry several object labels, packed ragged output.
extern "C" __attribute__((visibility("default")))
int ape_surface_backproject_multi(const uint8_t* label, const uint16_t* depth, int n_frames, int height, int width,
                                  const uint8_t* label_values_host, int n_labels, const double* cam, const double* robot2cam,
                                  int total_capacity, double* points, int32_t* pixel_index, int32_t* counts, int32_t* offsets,
                                  void* work, void* stream)
{
    APE_REQUIRE(label && depth && label_values_host && cam && robot2cam && points && counts && offsets && work,
                "ape_surface_backproject_multi: null pointer");
    APE_REQUIRE(n_frames >= 0 && height > 0 && width > 0 && total_capacity > 0, "ape_surface_backproject_multi: bad sizes");
    APE_REQUIRE(n_labels >= 1 && n_labels <= 8, "ape_surface_backproject_multi: 1..8 label values per frame");
    APE_REQUIRE(((size_t)height * width) % 16 == 0, "ape_surface_backproject_multi: height*width must be a multiple of 16");
    APE_REQUIRE((size_t)height * width < (1u << 30), "ape_surface_backproject_multi: frame too large");
    APE_REQUIRE((((uintptr_t)label) & 15) == 0 && (((uintptr_t)depth) & 15) == 0 && (((uintptr_t)work) & 15) == 0,
                "ape_surface_backproject_multi: label / depth / work must be 16-byte aligned");
    if (n_frames == 0) return APE_OK;
    const int n_views = n_frames * n_labels;
    const int n_chunks = (int)(((size_t)height * width + ape::kSurfChunk - 1) / ape::kSurfChunk);
    APE_REQUIRE((size_t)n_views * n_chunks * ape::kSurfWarps < (1u << 31), "ape_surface_backproject_multi: too many views (split the batch)");
    cudaStream_t s = (cudaStream_t)stream;
    ape::LabelSet ls;
    ls.n = n_labels;
    for (int i = 0; i < 8; ++i) ls.value[i] = i < n_labels ? label_values_host[i] : 0;
    for (int i = 0; i < n_labels; ++i) APE_REQUIRE(ls.value[i] != 0, "ape_surface_backproject_multi: label value 0 is the background");
    const int n_sub = n_chunks * ape::kSurfWarps;
    const size_t n_span = (size_t)n_views * n_sub;
    int32_t* n_tasks = reinterpret_cast<int32_t*>(work);
    int4* tasks = reinterpret_cast<int4*>(work) + 1;
    int32_t* sub_count = reinterpret_cast<int32_t*>(tasks + n_span);
    uint32_t* masks = reinterpret_cast<uint32_t*>(sub_count + n_span);
    {
        ape::ProfScope prof_("surface_mask_multi", s);
        APE_CUDA(ape::launch_pdl(ape::surface_mask_multi_kernel, dim3(n_frames * n_chunks), dim3(ape::kSurfThreads), 0, s, label, depth,
                                 height * width, ls, sub_count, masks, n_chunks, n_tasks));
        ape::count_launch();
    }
    int rc = ape::check_launch("ape_surface_backproject_multi (mask)");
    if (rc) return rc;
    {
        ape::ProfScope prof_("surface_scan", s);
        APE_CUDA(ape::launch_pdl(ape::surface_scan_kernel, dim3((n_views + 7) / 8), dim3(256), 0, s, sub_count, counts, n_tasks, tasks, n_sub, n_views));
        ape::count_launch();
    }
    {
        ape::ProfScope prof_("view_offsets", s);
        APE_CUDA(ape::launch_pdl(ape::view_offsets_kernel, dim3(1), dim3(1024), 0, s, (const int32_t*)counts, n_views, offsets));
        ape::count_launch();
    }
    ape::ProfScope prof_("surface_emit", s);
    const size_t want_ctas = (n_span + ape::kEmitWarps - 1) / ape::kEmitWarps;
    const size_t max_ctas = (size_t)ape::sm_count() * 8;
    APE_CUDA(ape::launch_pdl(ape::surface_emit_kernel, dim3((unsigned)(want_ctas < max_ctas ? want_ctas : max_ctas)), dim3(ape::kEmitWarps * 32), 0, s,
                             depth, height * width, width, (const int32_t*)nullptr, cam, robot2cam, total_capacity, points, pixel_index,
                             (const int32_t*)n_tasks, (const int4*)tasks, (const uint32_t*)masks, n_sub, (const int32_t*)offsets, n_labels));
    ape::count_launch();
    return ape::check_launch("ape_surface_backproject_multi (emit)");
}
