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
// Two embarrassingly parallel kernels, no inter-CTA waiting:
//   surface_mask_kernel : pure stream over label + depth (the 921 600 B/frame of SURVEY 8d).  Each thread owns 32
//                         consecutive pixels (2 x 16 B of label, 4 x 16 B of depth, streaming loads), builds a 32-bit
//                         validity mask, writes it (4 B per 96 B read) and the CTA writes its chunk's pop-count.
//   surface_emit_kernel : per 8192-pixel chunk with at least one valid pixel: base offset = sum of the counts of the
//                         chunks before it in the same view (<= 37 words), block scan of the mask pop-counts, compaction
//                         of the in-chunk offsets into shared memory (row-major = np.where order), then ALL threads
//                         convert the compacted pixels (balanced fp64 work, contiguous 24-byte point stores).  Depth is
//                         re-read only for valid pixels.
// History (profiles/): one CTA per view walking the frame = 1.35 TB/s (latency-bound); single-pass chunked kernels
// with a decoupled look-back = 1.5 TB/s, and 0.16 TB/s when made persistent (convoy on the look-back flags).
constexpr int kSurfThreads = 256;
constexpr int kSurfPix = 32;                       // pixels per thread
constexpr int kSurfChunk = kSurfThreads * kSurfPix;   // 8192 pixels = 24 KB of input per CTA

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

// 4 label bytes -> 4 bits (bit i set when byte i matches: != 0 if want == 0 else == want), SIMD-in-register
__device__ __forceinline__ uint32_t label_bits(uint32_t w, uint32_t want4) {
    // __vcmpeq4 gives 0xff per equal byte; want4 = want replicated (0 -> "byte == 0", inverted below)
    const uint32_t eq = __vcmpeq4(w, want4);
    const uint32_t hit = want4 ? eq : ~eq;                 // want == 0: label != 0
    // gather the top bit of each byte into bits 0..3
    return ((hit & 0x80808080u) * 0x00204081u) >> 28;
}
__device__ __forceinline__ uint32_t depth_bits(uint32_t w) {
    const uint32_t nz = ~__vcmpeq2(w, 0u);                 // 0xffff per non-zero half
    return ((nz >> 15) & 1u) | ((nz >> 30) & 2u);
}

// work layout: counts [n_views * n_chunks] int32, then masks [n_views * n_chunks * 256] u32 (one bit per pixel)
__global__ void __launch_bounds__(kSurfThreads, 4)
surface_mask_kernel(const uint8_t* __restrict__ label, const uint16_t* __restrict__ depth, int npix,
                    const int32_t* __restrict__ frame_of, const uint8_t* __restrict__ label_value,
                    int32_t* __restrict__ chunk_count, uint32_t* __restrict__ masks, int n_chunks)
{
    __shared__ int s_warp_tot[kSurfThreads / 32];
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
    const int cnt = warp_sum(__popc(mask));
    if ((threadIdx.x & 31) == 0) s_warp_tot[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kSurfThreads / 32; ++w) t += s_warp_tot[w];
        chunk_count[tile] = t;
    }
}

__global__ void __launch_bounds__(kSurfThreads)
surface_emit_kernel(const uint16_t* __restrict__ depth, int H, int W, const int32_t* __restrict__ frame_of,
                    const double* __restrict__ cam, const double* __restrict__ robot2cam, int capacity,
                    double* __restrict__ points, int32_t* __restrict__ pixel_index, int32_t* __restrict__ counts,
                    const int32_t* __restrict__ chunk_count, const uint32_t* __restrict__ masks, int n_chunks)
{
    __shared__ uint16_t s_list[kSurfChunk];                // compacted in-chunk offsets of the valid pixels
    __shared__ int s_warp_tot[kSurfThreads / 32];
    __shared__ int s_base;
    __shared__ double s_par[16];                           // robot2cam rows 0..2 (12) + ppx, ppy, fx, fy
    const int tile = blockIdx.x;
    const int v = tile / n_chunks, c = tile - v * n_chunks;
    const int total = chunk_count[tile];
    if (total == 0 && c != n_chunks - 1) return;           // nothing to emit (uniform across the CTA)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {                                       // base = valid pixels of the chunks before this one
        int sum = 0;
        for (int j = lane; j < c; j += 32) sum += chunk_count[v * n_chunks + j];
        sum = warp_sum(sum);
        if (lane == 0) {
            s_base = sum;
            if (c == n_chunks - 1) counts[v] = sum + total;
        }
    } else if (warp == 1 && lane < 16) {
        s_par[lane] = lane < 12 ? robot2cam[16 * v + lane] : cam[4 * v + (lane - 12)];
    }
    if (total == 0) return;
    const uint32_t mask = masks[(size_t)tile * kSurfThreads + threadIdx.x];
    const int cnt = __popc(mask);
    int incl = cnt;                                        // warp inclusive scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int w = 0; w < kSurfThreads / 32; ++w) before += (w < warp) ? s_warp_tot[w] : 0;
    {
        int slot = before + incl - cnt;
        uint32_t m = mask;
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            s_list[slot++] = (uint16_t)(threadIdx.x * kSurfPix + i);
        }
    }
    __syncthreads();
    const int f = frame_of ? frame_of[v] : v;
    const uint16_t* dep = depth + (size_t)f * H * W;
    const int p0 = c * kSurfChunk;
    const int base = s_base;
    const double* T = s_par;
    const double ppx = s_par[12], ppy = s_par[13], fx = s_par[14], fy = s_par[15];
    double* out = points + (size_t)v * capacity * 3;
    int32_t* opix = pixel_index ? pixel_index + (size_t)v * capacity : nullptr;
    for (int i = threadIdx.x; i < total; i += kSurfThreads) {
        const int slot = base + i;
        if (slot >= capacity) break;
        const int pix = p0 + (int)s_list[i];
        const int r = pix / W, col = pix - r * W;
        const double z = (double)dep[pix];
        // open3d_utils.py:185-189: p0 = ((px-ppx)*z)/fx, p1 = ((py-ppy)*z)/fy (fp64, rounded per op)
        const double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)col, ppx), z), fx);
        const double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)r, ppy), z), fy);
        // robot2cam * [x y z 1]^T (:190-192)
        double* o = out + (size_t)slot * 3;
        o[0] = fma(T[0], x, fma(T[1], y, fma(T[2], z, T[3])));
        o[1] = fma(T[4], x, fma(T[5], y, fma(T[6], z, T[7])));
        o[2] = fma(T[8], x, fma(T[9], y, fma(T[10], z, T[11])));
        if (opix) opix[slot] = pix;
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
    return 4 * tiles + 4 * tiles * ape::kSurfThreads;        // chunk counts + one mask bit per pixel
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
    APE_REQUIRE((((uintptr_t)work) & 3) == 0, "ape_surface_backproject: work must be 4-byte aligned");
    if (n_views == 0) return APE_OK;
    const int n_chunks = (int)(((size_t)height * width + ape::kSurfChunk - 1) / ape::kSurfChunk);
    APE_REQUIRE((size_t)n_views * n_chunks < (1u << 31), "ape_surface_backproject: too many views (split the batch)");
    cudaStream_t s = (cudaStream_t)stream;
    const int n_tiles = n_views * n_chunks;
    int32_t* chunk_count = reinterpret_cast<int32_t*>(work);
    uint32_t* masks = reinterpret_cast<uint32_t*>(work) + n_tiles;
    {
        ape::ProfScope prof_("surface_mask", s);
        ape::surface_mask_kernel<<<n_tiles, ape::kSurfThreads, 0, s>>>(label, depth, height * width, frame_of, label_value,
                                                                       chunk_count, masks, n_chunks);
        ape::count_launch();
    }
    int rc = ape::check_launch("ape_surface_backproject (mask)");
    if (rc) return rc;
    ape::ProfScope prof_("surface_emit", s);
    ape::surface_emit_kernel<<<n_tiles, ape::kSurfThreads, 0, s>>>(depth, height, width, frame_of, cam, robot2cam, capacity, points,
                                                                   pixel_index, counts, chunk_count, masks, n_chunks);
    ape::count_launch();
    return ape::check_launch("ape_surface_backproject (emit)");
}
