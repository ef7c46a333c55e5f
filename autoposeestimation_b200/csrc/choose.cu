// Device-side mask -> bbox -> choose (+ back-projection) for the option-6 geometry block (sm_100a).
// Replaces, per detected object, the host work of pipeline/utils.py:524-553:
//   mask_label = (label == value); bbox = get_bbox(mask_label)           (datasets/myDatasetAugmented/dataset.py:342-380)
//   choose = flat crop indices of (mask_label & depth != 0), row-major   (:529)
//   > N candidates: a uniformly random N-subset, ascending               (:532-537, np.random.shuffle of a 0/1 vector)
//   <= N candidates: cyclic 'wrap' padding                               (:539)
//   cloud = fp32 back-projection at choose                               (:542-553, same arithmetic as ape_backproject_choose)
// so that the segmentation output never has to leave the device and the 300 k-element xmap/ymap lists (:518-519) are gone.
//
// Random subset: the reference draws it from numpy's global MT19937 stream, which cannot be replayed on the device.
// Here candidate c gets the key mix32(seed ^ c * 0x9E3779B9) and the N smallest keys are kept (ties: lower index first) --
// the same distribution (every N-subset equally likely for a good mixer), bit-exact against oracle/geometry.py with the
// same seed.  For exact replay of a host RNG keep using ape_backproject_choose with host-made `choose`.
// One CTA per object: bbox by a full-frame scan of the label, then a 4-pass radix select over the crop, then an ordered
// emission (block scan) -- every pass re-derives keys from the pixel index, nothing is stored per candidate.
#include "ape_common.cuh"

namespace ape {

constexpr int kChThreads = 1024;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {          // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t cand_key(uint32_t seed, uint32_t c) { return mix32(seed ^ (c * 0x9E3779B9u)); }

// dataset.py:350-358: an extent strictly between two multiples of 40 is raised to the upper one
__device__ __forceinline__ int round_up_border(int e) { return (e % 40) ? (e / 40 + 1) * 40 : e; }

__global__ void __launch_bounds__(kChThreads)
mask_bbox_choose_kernel(const uint8_t* __restrict__ label, const uint16_t* __restrict__ depth, int H, int W,
                        const int32_t* __restrict__ frame_of, const uint8_t* __restrict__ label_value,
                        const uint32_t* __restrict__ seeds, const float* __restrict__ cam, int n_points,
                        int32_t* __restrict__ bbox_out, int32_t* __restrict__ n_cand_out, int64_t* __restrict__ choose,
                        float* __restrict__ cloud)
{
    __shared__ int s_box[4];                      // rmin, rmax (inclusive), cmin, cmax (inclusive)
    __shared__ int s_hist[256];
    __shared__ int s_w[kChThreads / 32];
    __shared__ int s_total, s_carry, s_tie_carry;
    __shared__ uint32_t s_prefix;
    __shared__ int s_need;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = frame_of ? frame_of[b] : b;
    const uint8_t want = label_value ? label_value[b] : 255;
    const uint8_t* L = label + (size_t)f * H * W;
    const uint16_t* D = depth + (size_t)f * H * W;
    if (tid == 0) { s_box[0] = H; s_box[1] = -1; s_box[2] = W; s_box[3] = -1; }
    __syncthreads();
    // ---- bbox of the label mask (full-frame scan, 4 pixels per load)
    {
        int rmin = H, rmax = -1, cmin = W, cmax = -1;
        const int npix = H * W;
        for (int p = tid * 4; p < npix; p += kChThreads * 4) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(L + p);          // W % 4 == 0 (checked by the host)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (((v >> (8 * k)) & 0xffu) == want) {
                    const int r = (p + k) / W, c = (p + k) - r * W;
                    rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
                }
            }
        }
        if (rmax >= 0) { atomicMin(&s_box[0], rmin); atomicMax(&s_box[1], rmax); atomicMin(&s_box[2], cmin); atomicMax(&s_box[3], cmax); }
    }
    __syncthreads();
    if (s_box[1] < 0) {                                         // no labelled pixel: object skipped (:530-531)
        if (tid == 0) { n_cand_out[b] = 0; bbox_out[4 * b] = bbox_out[4 * b + 1] = bbox_out[4 * b + 2] = bbox_out[4 * b + 3] = 0; }
        return;
    }
    // ---- get_bbox (dataset.py:342-380); all threads compute the same four integers
    int rmin = s_box[0], rmax = s_box[1] + 1, cmin = s_box[2], cmax = s_box[3] + 1;
    {
        const int r_b = round_up_border(rmax - rmin), c_b = round_up_border(cmax - cmin);
        const int cr = (rmin + rmax) / 2, cc = (cmin + cmax) / 2;
        rmin = cr - r_b / 2; rmax = cr + r_b / 2; cmin = cc - c_b / 2; cmax = cc + c_b / 2;
        if (rmin < 0) { rmax -= rmin; rmin = 0; }
        if (cmin < 0) { cmax -= cmin; cmin = 0; }
        if (rmax > H) { rmin -= rmax - H; rmax = H; }
        if (cmax > W) { cmin -= cmax - W; cmax = W; }
    }
    const int cw = cmax - cmin, ncrop = (rmax - rmin) * cw;
    const uint32_t seed = seeds ? seeds[b] : 0u;
    auto is_cand = [&](int c) -> bool {
        const int r = rmin + c / cw, col = cmin + c % cw;
        if (r < 0 || r >= H || col < 0 || col >= W) return false;
        const int p = r * W + col;
        return L[p] == want && D[p] != 0;
    };
    // ---- number of candidates
    {
        int cnt = 0;
        for (int c = tid; c < ncrop; c += kChThreads) cnt += is_cand(c) ? 1 : 0;
        cnt = warp_sum(cnt);
        if (lane == 0) s_w[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < kChThreads / 32; ++w) t += s_w[w];
            s_total = t; s_prefix = 0u; s_need = n_points; s_carry = 0; s_tie_carry = 0;
            n_cand_out[b] = t;
            bbox_out[4 * b] = rmin; bbox_out[4 * b + 1] = rmax; bbox_out[4 * b + 2] = cmin; bbox_out[4 * b + 3] = cmax;
        }
        __syncthreads();
    }
    const int total = s_total;
    if (total == 0) return;
    const bool subsample = total > n_points;
    uint32_t thr = 0xffffffffu;        // keep key < thr, plus the first `tie_quota` candidates with key == thr
    int tie_quota = 0;
    if (subsample) {
        // ---- radix select: the n_points-th smallest key, 8 bits per pass from the top
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int i = tid; i < 256; i += kChThreads) s_hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
            for (int c = tid; c < ncrop; c += kChThreads) {
                if (is_cand(c)) {
                    const uint32_t k = cand_key(seed, (uint32_t)c);
                    if ((k & pmask) == prefix) atomicAdd(&s_hist[(k >> shift) & 0xffu], 1);
                }
            }
            __syncthreads();
            if (tid == 0) {
                int need = s_need, d = 0;
                while (d < 255 && s_hist[d] < need) { need -= s_hist[d]; ++d; }
                s_need = need;                                  // still to take inside bucket d
                s_prefix = prefix | ((uint32_t)d << shift);
            }
            __syncthreads();
        }
        thr = s_prefix; tie_quota = s_need;                    // keys equal to thr: take the first s_need in index order
    }
    // ---- ordered emission (row-major = ascending choose), block scan per chunk of 1024 crop pixels
    const float ppx = cam[5 * b], ppy = cam[5 * b + 1], fx = cam[5 * b + 2], fy = cam[5 * b + 3], scale = cam[5 * b + 4];
    int64_t* CH = choose + (size_t)b * n_points;
    float* CL = cloud ? cloud + (size_t)b * n_points * 3 : nullptr;
    auto emit = [&](int slot, int c) {
        CH[slot] = c;
        if (CL) {
            const int r = rmin + c / cw, col = cmin + c % cw;
            const float z = __fmul_rn((float)D[r * W + col], scale);
            CL[3 * slot] = __fdiv_rn(__fmul_rn(__fsub_rn((float)col, ppx), z), fx);
            CL[3 * slot + 1] = __fdiv_rn(__fmul_rn(__fsub_rn((float)r, ppy), z), fy);
            CL[3 * slot + 2] = z;
        }
    };
    for (int base = 0; base < ncrop; base += kChThreads) {
        const int c = base + tid;
        bool cand = c < ncrop && is_cand(c);
        bool tie = false;
        if (cand && subsample) {
            const uint32_t k = cand_key(seed, (uint32_t)c);
            tie = k == thr;
            cand = k < thr;
        }
        // two scans: strictly-below-threshold candidates, and ties (admitted in index order up to tie_quota)
        int v0 = cand ? 1 : 0, v1 = tie ? 1 : 0;
        int i0 = v0, i1 = v1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += t0; i1 += t1; }
        }
        if (lane == 31) s_w[warp] = i0 | (i1 << 16);           // chunk <= 1024 pixels: both counts fit 16 bits... per warp <= 32
        __syncthreads();
        int b0 = 0, b1 = 0;
        for (int w = 0; w < warp; ++w) { b0 += s_w[w] & 0xffff; b1 += s_w[w] >> 16; }
        const int ties_before = s_tie_carry + b1 + i1 - v1;    // ties with a lower index
        const bool take_tie = tie && ties_before < tie_quota;
        // slot = (#kept strictly-below before me) + (#admitted ties before me)
        const int below_before = s_carry + b0 + i0 - v0;
        const int adm_before = min(ties_before, tie_quota);
        if (cand || take_tie) {
            const int slot = below_before + adm_before;
            if (subsample) emit(slot, c);
            else {
                // wrap padding (:539): candidate j also fills slots j + total, j + 2 total, ...
                for (int s2 = slot; s2 < n_points; s2 += total) emit(s2, c);
            }
        }
        __syncthreads();
        if (tid == kChThreads - 1) { s_carry = below_before + v0; s_tie_carry = ties_before + v1; }
        __syncthreads();
    }
}

}  // namespace ape

extern "C" __attribute__((visibility("default")))
int ape_mask_bbox_choose(const uint8_t* label, const uint16_t* depth, int n_frames, int height, int width, const int32_t* frame_of,
                         const uint8_t* label_value, const uint32_t* seeds, const float* cam, int n_obj, int n_points,
                         int32_t* bbox, int32_t* n_candidates, int64_t* choose, float* cloud, void* stream)
{
    APE_REQUIRE(label && depth && cam && bbox && n_candidates && choose, "ape_mask_bbox_choose: null pointer");
    APE_REQUIRE(n_frames > 0 && height > 0 && width > 0 && n_obj >= 0 && n_points > 0, "ape_mask_bbox_choose: bad sizes");
    APE_REQUIRE(width % 4 == 0 && height * (long long)width < (1ll << 31), "ape_mask_bbox_choose: width must be a multiple of 4");
    APE_REQUIRE(height % 40 == 0 && width % 40 == 0, "ape_mask_bbox_choose: get_bbox's border list assumes multiples of 40 (480 x 640)");
    if (n_obj == 0) return APE_OK;
    ape::ProfScope prof_("mask_bbox_choose", (cudaStream_t)stream);
    ape::mask_bbox_choose_kernel<<<n_obj, ape::kChThreads, 0, (cudaStream_t)stream>>>(label, depth, height, width, frame_of, label_value,
                                                                                     seeds, cam, n_points, bbox, n_candidates, choose, cloud);
    ape::count_launch();
    return ape::check_launch("ape_mask_bbox_choose");
}
