// PoseRefineNet training step (DenseFusion/tools/train.py:215-233) -- included at the end of net.cu.
//
//   forward   : the inference trunk in plain bf16 (one tcgen05 pass instead of the three split-bf16 passes; BASELINE
//               config 5 is a bf16 training step), activations kept: PF [R,384], H5 [R,512], the ReLU sign bits of
//               conv6 (the [R,1024] map itself is never written, as in inference), AP, G1, G2
//   backward  : heads on SIMT fp32 (batch x 1024 matrices), trunk on gemm_train.cuh (dgrad / wgrad on tcgen05 straight
//               from the row-major buffers), conv1 / e_conv1 weight gradients on SIMT
//   optimizer : Adam (train.py:149, :410) over ONE flat fp32 parameter vector; its flat gradient is the buffer the
//               caller all-reduces over NCCL (the only collective of the scope)
//
// Flat parameter layout (floats; O = num_obj).  Grouped layers are contiguous so the grouped GEMMs see one matrix:
//   head block: conv1.w [64,3] | e_conv1.w [64,32] | biases of conv1, e_conv1, conv2|e_conv2, conv5 |
//               conv2.w, e_conv2.w [2x128,64] | conv5.w [512,384]
//   tail block: conv6.w [1024,512] | conv1_r.w, conv1_t.w [2x512,1024] | conv2_r.w, conv2_t.w [2x128,512] |
//               conv3_r.w [4O,128] | conv3_t.w [3O,128] | biases of conv6, conv1_{r,t}, conv2_{r,t}, conv3_r, conv3_t
//   (ape_refiner_trainer_layout gives every offset by name; the blocks exist for the overlapped all-reduce, see train_layout)
#include "gemm_train.cuh"

namespace ape {

struct TrainLayout {
    size_t w1, we1, w2e2, w5, w6, wh1, wh2, w3r, w3t, b1, be1, b2e2, b5, b6, bh1, bh2, b3r, b3t, bulk, total;
};
static TrainLayout train_layout(int O) {
    // Two contiguous blocks.  TAIL block [bulk, total): conv6 + the heads, weights AND biases -- 89 % of the vector, and
    // complete as soon as the LAST iteration's conv6 weight gradient is done, i.e. before the backward pass walks conv5,
    // conv2 and conv1: its all-reduce overlaps that tail (ape_refiner_trainer_wait_bulk).  HEAD block [0, bulk): the rest.
    TrainLayout L; size_t o = 0;
    // (the head block's biases sit in front of its big weights so that conv2|e_conv2 .. conv2_{r,t} stay one contiguous
    // range across the block boundary: one fp32 -> bf16 refresh launch, ape_refiner_trainer_sync_weights)
    L.w1 = o; o += 64 * 3;      L.we1 = o; o += 64 * 32;
    L.b1 = o; o += 64; L.be1 = o; o += 64; L.b2e2 = o; o += 256; L.b5 = o; o += 512;
    L.w2e2 = o; o += 256 * 64;  L.w5 = o; o += 512 * 384;
    L.bulk = o;
    L.w6 = o; o += 1024 * 512;  L.wh1 = o; o += 1024 * 1024; L.wh2 = o; o += 256 * 512;
    L.w3r = o; o += (size_t)4 * O * 128; L.w3t = o; o += (size_t)3 * O * 128;
    L.b6 = o; o += 1024; L.bh1 = o; o += 1024; L.bh2 = o; o += 256; L.b3r = o; o += 4 * O; L.b3t = o; o += 3 * O;
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n)
{
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= n) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    } else {
        for (size_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
}

// packed front-end weights [e_conv1^T 32x64 | conv1^T 3x64 | b1 | be1] (frontend_kernel) from the flat vector
__global__ void pack_frontend_kernel(const float* __restrict__ w1, const float* __restrict__ we1, const float* __restrict__ b1,
                                     const float* __restrict__ be1, float* __restrict__ fw)
{
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) { const int k = i >> 6, c = i & 63; fw[i] = we1[c * 32 + k]; }
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) { const int k = i >> 6, c = i & 63; fw[32 * 64 + i] = w1[c * 3 + k]; }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { fw[32 * 64 + 3 * 64 + i] = b1[i]; fw[32 * 64 + 3 * 64 + 64 + i] = be1[i]; }
}

// Backward of conv3_{r,t} restricted to the object's class (network.py:199-204) and of the ReLU in front of it:
//   dz2[b, c] = G2[b, c] > 0 ? sum_j d[b, j] * w3[(o*nj + j), c % 128] : 0;   dW3, db3, db2 accumulate (atomics)
__global__ void __launch_bounds__(256)
head3_bwd_kernel(const float* __restrict__ d_r, const float* __restrict__ d_t, const int64_t* __restrict__ obj, int num_obj,
                 const bf16* __restrict__ g2, const float* __restrict__ w3r, const float* __restrict__ w3t,
                 bf16* __restrict__ dz2, float* __restrict__ gw3r, float* __restrict__ gw3t, float* __restrict__ gb3r,
                 float* __restrict__ gb3t, float* __restrict__ gb2)
{
    const int b = blockIdx.x, c = threadIdx.x, h = c >> 7, k = c & 127;
    int o = (int)obj[b];
    o = o < 0 ? 0 : (o >= num_obj ? num_obj - 1 : o);
    const int nj = h == 0 ? 4 : 3;
    const float* d = h == 0 ? d_r + 4 * b : d_t + 3 * b;
    const float* w = (h == 0 ? w3r : w3t) + (size_t)o * nj * 128;
    float* gw = (h == 0 ? gw3r : gw3t) + (size_t)o * nj * 128;
    const float x = __bfloat162float(g2[(size_t)b * 256 + c]);
    float s = 0.f;
    for (int j = 0; j < nj; ++j) {
        const float dj = d[j];
        s = fmaf(dj, w[j * 128 + k], s);
        atomicAdd(gw + j * 128 + k, dj * x);
    }
    const float dz = x > 0.f ? s : 0.f;
    dz2[(size_t)b * 256 + c] = __float2bfloat16_rn(dz);
    atomicAdd(gb2 + c, dz);
    if (k < nj) atomicAdd((h == 0 ? gb3r : gb3t) + o * nj + k, d[k]);
}

// Backward of AvgPool1d + the ReLU of conv6 (network.py:163-167): dY6[r, c] = bit(r, c) ? g6[b, c] : 0 (bf16), where
// g6 = dAP / N already, and db6[c] += sum_r dY6[r, c].  One CTA per 128-row tile; thread = 8 consecutive columns (one
// byte of sign bits in, one 16-byte store out per row), two row phases per CTA, 4 rows in flight per thread.
__global__ void __launch_bounds__(256)
dy6_kernel(const uint32_t* __restrict__ bits, const bf16* __restrict__ g6 /* dAP [Bp,1024] */, float inv_n, int Np, int rows_live,
           bf16* __restrict__ dy6, float* __restrict__ part6 /* [tiles][1024]: this tile's column sums (no atomics: see gemm_train.cuh) */)
{
    __shared__ uint32_t s_cnt[128][8];
    const int row0 = blockIdx.x * 128;
    const int cg = threadIdx.x & 127, rh = threadIdx.x >> 7;
    const int c = cg * 8;
    const int b = row0 / Np;                                  // Np is a multiple of 128: one object per tile
    const bool live = row0 < rows_live;
    uint32_t u[4] = {0u, 0u, 0u, 0u};
    if (live) {
        const uint4 gv = *reinterpret_cast<const uint4*>(g6 + (size_t)b * 1024 + c);          // 8 bf16 of dAP; AvgPool1d backward = / N
        const uint32_t gu[4] = {gv.x, gv.y, gv.z, gv.w};
        float ge[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { ge[2 * q] = __uint_as_float(gu[q] << 16) * inv_n; ge[2 * q + 1] = __uint_as_float(gu[q] & 0xffff0000u) * inv_n; }
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(ge[0], ge[1]), h1 = __floats2bfloat162_rn(ge[2], ge[3]);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(ge[4], ge[5]), h3 = __floats2bfloat162_rn(ge[6], ge[7]);
        u[0] = *reinterpret_cast<const uint32_t*>(&h0); u[1] = *reinterpret_cast<const uint32_t*>(&h1);
        u[2] = *reinterpret_cast<const uint32_t*>(&h2); u[3] = *reinterpret_cast<const uint32_t*>(&h3);
    }
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(bits);       // little-endian: byte j of a word = columns 8j..8j+7
    uint32_t cnt[8] = {};
    for (int r0 = rh; r0 < 128; r0 += 8) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = live ? (uint32_t)__ldg(bytes + ((size_t)row0 + r0 + 2 * i) * 128 + cg) : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t b0 = (w[i] >> (2 * q)) & 1u, b1 = (w[i] >> (2 * q + 1)) & 1u;
                o[q] = (b0 ? (u[q] & 0xffffu) : 0u) | (b1 ? (u[q] & 0xffff0000u) : 0u);
                cnt[2 * q] += b0; cnt[2 * q + 1] += b1;
            }
            *reinterpret_cast<uint4*>(dy6 + ((size_t)row0 + r0 + 2 * i) * 1024 + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    // the bias gradient sums the bf16-rounded values the GEMMs see; the two row phases are combined through smem
    if (rh == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s_cnt[cg][j] = cnt[j];
    }
    __syncthreads();
    if (rh == 0) {
        float o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[2 * q] = (float)(cnt[2 * q] + s_cnt[cg][2 * q]) * __uint_as_float(u[q] << 16);
            o[2 * q + 1] = (float)(cnt[2 * q + 1] + s_cnt[cg][2 * q + 1]) * __uint_as_float(u[q] & 0xffff0000u);
        }
        float4* dst = reinterpret_cast<float4*>(part6 + (size_t)blockIdx.x * 1024 + c);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// Folds the partial sums of one backward pass into the flat gradient (one short launch instead of deep atomic chains):
//   biases of conv5 | conv2,e_conv2 | conv1,e_conv1 from kBiasCopies copies, conv1 / e_conv1 weights from kW1Copies
//   copies, conv6 bias from the per-tile sums of dy6_kernel.
constexpr int kW1Copies = 16;
constexpr int kBiasPart = 2048;                 // floats per bias copy: [b5 512 | b2e2 256 | b1,be1 128 | pad 128 | bh1 1024]
constexpr int kW1Part = 64 * 3 + 64 * 32;       // conv1.w | e_conv1.w, contiguous in the flat layout
__global__ void __launch_bounds__(1024)
fold_partials_kernel(const float* __restrict__ part_b, const float* __restrict__ part_w1, const float* __restrict__ part6, int tiles,
                     float* __restrict__ g_b5, float* __restrict__ g_b2e2, float* __restrict__ g_b1, float* __restrict__ g_w1,
                     float* __restrict__ g_b6, float* __restrict__ g_bh1, int phase /* 0: conv6 / conv1_{r,t} biases (their partials are
                     complete after the conv6 weight gradient), 1: the rest (end of the backward pass) */)
{
    __shared__ float s_red[32][33];
    if (blockIdx.x < 32) {
        if (phase != 0) return;
        // conv6 bias: block = 32 columns x 32 tile segments, 4 loads in flight per thread, fixed-order tree over the segments
        const int cx = threadIdx.x & 31, sg = threadIdx.x >> 5, c = blockIdx.x * 32 + cx;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int t = sg;
        for (; t + 96 < tiles; t += 128) {
            s0 += part6[(size_t)t * 1024 + c]; s1 += part6[(size_t)(t + 32) * 1024 + c];
            s2 += part6[(size_t)(t + 64) * 1024 + c]; s3 += part6[(size_t)(t + 96) * 1024 + c];
        }
        for (; t < tiles; t += 32) s0 += part6[(size_t)t * 1024 + c];
        s_red[sg][cx] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (sg == 0) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) s += s_red[k][cx];
            g_b6[c] += s;
        }
        return;
    }
    const int i = (blockIdx.x - 32) * 1024 + threadIdx.x;
    if ((phase == 0) != (i >= 4096)) return;                 // [4096, 5120) = conv1_{r,t} bias: phase 0
    if (i < 896) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < tr::kBiasCopies; ++k) s += part_b[k * kBiasPart + i];
        float* dst = i < 512 ? g_b5 + i : (i < 768 ? g_b2e2 + (i - 512) : g_b1 + (i - 768));
        *dst += s;
    } else if (i >= 1024 && i < 1024 + kW1Part) {
        const int j = i - 1024;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kW1Copies; ++k) s += part_w1[k * kW1Part + j];
        g_w1[j] += s;
    } else if (i >= 4096 && i < 4096 + 1024) {
        const int j = i - 4096;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < tr::kBiasCopies; ++k) s += part_b[k * kBiasPart + 1024 + j];
        g_bh1[j] += s;
    }
}

// Weight gradients of the K=3 / K=32 first layers: dW1[c, k] += sum_r dz1[r, c] * cloud[r, k],
// dWe1[c, k] += sum_r dze1[r, c] * emb[r, k];  dz = dPF[:, 0:128] (bf16).  Persistent CTAs over 64-row chunks.
__global__ void __launch_bounds__(256)
conv1_wgrad_kernel(const bf16* __restrict__ dpf, int ld, const float* __restrict__ cloud, const float* __restrict__ emb,
                   int B, int N, int Np, float* __restrict__ gw1, float* __restrict__ gwe1)
{
    __shared__ float s_dz[64][129];
    __shared__ __align__(16) float s_x[64][36];
    const int tid = threadIdx.x, c = tid & 63, kq = tid >> 6;
    float acc_e[8] = {}, acc_x[3] = {};
    const int chunks = B * (Np / 64);
    for (int ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const int b = ch / (Np / 64), n0 = (ch % (Np / 64)) * 64;
        if (n0 >= N) continue;                                  // whole chunk is padding (uniform per CTA)
        __syncthreads();
        for (int i = tid; i < 64 * 16; i += 256) {              // 64 rows x 16 chunks of 8 bf16 (16-byte loads)
            const int r = i >> 4, ch8 = i & 15;
            const uint4 v = *reinterpret_cast<const uint4*>(dpf + ((size_t)b * Np + n0 + r) * ld + 8 * ch8);
            const uint32_t u4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s_dz[r][8 * ch8 + 2 * e] = __uint_as_float(u4[e] << 16);
                s_dz[r][8 * ch8 + 2 * e + 1] = __uint_as_float(u4[e] & 0xffff0000u);
            }
        }
        for (int i = tid; i < 32 * 64; i += 256) {              // emb [B,32,N]: n fastest
            const int k = i >> 6, r = i & 63;
            s_x[r][k] = (n0 + r < N) ? emb[((size_t)b * 32 + k) * N + n0 + r] : 0.f;
        }
        for (int i = tid; i < 64 * 3; i += 256) {
            const int r = i / 3, k = i - 3 * r;
            s_x[r][32 + k] = (n0 + r < N) ? cloud[((size_t)b * N + n0 + r) * 3 + k] : 0.f;
        }
        __syncthreads();
        for (int r = 0; r < 64; ++r) {
            const float de = s_dz[r][64 + c];
            const float4 xa = *reinterpret_cast<const float4*>(&s_x[r][kq * 8]), xb = *reinterpret_cast<const float4*>(&s_x[r][kq * 8 + 4]);
            acc_e[0] = fmaf(de, xa.x, acc_e[0]); acc_e[1] = fmaf(de, xa.y, acc_e[1]); acc_e[2] = fmaf(de, xa.z, acc_e[2]);
            acc_e[3] = fmaf(de, xa.w, acc_e[3]); acc_e[4] = fmaf(de, xb.x, acc_e[4]); acc_e[5] = fmaf(de, xb.y, acc_e[5]);
            acc_e[6] = fmaf(de, xb.z, acc_e[6]); acc_e[7] = fmaf(de, xb.w, acc_e[7]);
            if (kq == 0) {
                const float dx = s_dz[r][c];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc_x[j] = fmaf(dx, s_x[r][32 + j], acc_x[j]);
            }
        }
    }
    // gw1 / gwe1 point at kW1Copies partial copies [conv1.w 192 | e_conv1.w 2048]: short atomic chains
    const size_t copy = (size_t)(blockIdx.x % 16) * (64 * 3 + 64 * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(gwe1 + copy + c * 32 + kq * 8 + j, acc_e[j]);
    if (kq == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) atomicAdd(gw1 + copy + c * 3 + j, acc_x[j]);
    }
}

// Adam as torch.optim.Adam (train.py:149): p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

}  // namespace ape

// ================================================================================================ host side
struct BfMat {                        // bf16 row-major matrix with its TMA maps
    bf16* p = nullptr; int rows = 0, cols = 0;
    CUtensorMap kmaj;                 // box {64 cols, 128 rows}: K-major operand (rows = M or N, cols = K)
    CUtensorMap mn2, mn4;             // 3-D view (64 cols, rows, cols/64 chunks), box {64, 64 rows, 2 | min(4, chunks)}: a whole
                                      // MN-major operand stage (rows = K, chunks = M or N) in one TMA op
    int mn4_chunks = 0;
    CUtensorMap st;                   // box {64 cols, 32 rows}:  one epilogue chunk (DGRAD mask / addend loads, result stores)
};

struct ape_trainer {
    ape_net net;                      // forward buffers and weights in the inference trunk's own structure (net.train = 1)
    ape::TrainLayout L;
    float *params = nullptr, *grads = nullptr;      // caller-owned flat vectors
    int num_obj = 0, max_batch = 0, max_points = 0;
    BfMat Wb2, Wb5, Wb6, Wbh1, Wbh2;  // bf16 copies of conv2|e_conv2, conv5, conv6, conv1_{r,t}, conv2_{r,t} (one contiguous allocation)
    BfMat PFm, H5m, dY6, dZ5, dPF;    // PFm / H5m alias net.PF.hi / net.H5.hi
    BfMat APm, G1m, G2m, dZ2h, dZ1h, g6; // heads: [Bp,1024], [Bp,1024], [Bp,256] activations and their gradients (bf16)
    int bp_max = 0;
    float *part_b = nullptr, *part_w1 = nullptr, *part6 = nullptr;   // partial sums folded by fold_partials_kernel
    // scratch of ape_refiner_trainer_step (allocated on first use, sized for max_batch x max(max_points, mesh points))
    float *s_r = nullptr, *s_t = nullptr, *s_dr = nullptr, *s_dt = nullptr, *s_pts[2] = {nullptr, nullptr}, *s_tgt[2] = {nullptr, nullptr};
    int s_mesh = 0;
    cudaEvent_t bulk_event = nullptr;  // recorded when the tail block of the flat gradient is final (last iteration only)
    int record_bulk = 0;
};

static int make_bf_maps(BfMat& m) {
    int rc = make_map_box(&m.kmaj, m.p, m.rows, m.cols, m.cols, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if ((rc = make_map_box(&m.st, m.p, m.rows, m.cols, m.cols, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { ape::set_error("cuTensorMapEncodeTiled entry point not available"); return APE_ERR_CUDA; }
    const int chunks = m.cols / 64;
    m.mn4_chunks = chunks < 4 ? chunks : 4;
    for (int which = 0; which < 2; ++which) {
        const int bc = which == 0 ? (chunks < 2 ? chunks : 2) : m.mn4_chunks;
        cuuint64_t dims[3] = {64, (cuuint64_t)m.rows, (cuuint64_t)chunks};
        cuuint64_t strides[2] = {(cuuint64_t)m.cols * 2, 128};                 // bytes: next row, next 64-column chunk
        cuuint32_t box[3] = {64, 64, (cuuint32_t)bc};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = fn(which == 0 ? &m.mn2 : &m.mn4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, m.p, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ape::set_error("cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r); return APE_ERR_CUDA; }
    }
    return APE_OK;
}

extern "C" __attribute__((visibility("default")))
int64_t ape_refiner_trainer_layout(int num_obj, int64_t* offsets24)
{
    if (num_obj <= 0) return -1;
    const ape::TrainLayout L = ape::train_layout(num_obj);
    if (offsets24) {
        // reference state_dict order (weight, bias per layer): feat.conv1, feat.e_conv1, feat.conv2, feat.e_conv2, feat.conv5,
        // feat.conv6, conv1_r, conv1_t, conv2_r, conv2_t, conv3_r, conv3_t
        const size_t o[24] = {L.w1, L.b1, L.we1, L.be1, L.w2e2, L.b2e2, L.w2e2 + 128 * 64, L.b2e2 + 128, L.w5, L.b5, L.w6, L.b6,
                              L.wh1, L.bh1, L.wh1 + 512 * 1024, L.bh1 + 512, L.wh2, L.bh2, L.wh2 + 128 * 512, L.bh2 + 128,
                              L.w3r, L.b3r, L.w3t, L.b3t};
        for (int i = 0; i < 24; ++i) offsets24[i] = (int64_t)o[i];
    }
    return (int64_t)L.total;
}

extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_destroy(ape_trainer* tr)
{
    if (!tr) return APE_OK;
    for (void* p : tr->net.allocs) cudaFree(p);
    if (tr->bulk_event) cudaEventDestroy(tr->bulk_event);
    delete tr;
    return APE_OK;
}

extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_sync_weights(ape_trainer* tr, void* stream)
{
    APE_REQUIRE(tr, "ape_refiner_trainer_sync_weights: null handle");
    cudaStream_t s = (cudaStream_t)stream;
    const ape::TrainLayout& L = tr->L;
    const size_t n = 256 * 64 + 512 * 384 + 1024 * 512 + 1024 * 1024 + 256 * 512;   // conv2|e_conv2 ... conv2_{r,t}: contiguous in the flat vector
    ape::f32_to_bf16_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(tr->params + L.w2e2, tr->Wb2.p, n);
    ape::pack_frontend_kernel<<<1, 256, 0, s>>>(tr->params + L.w1, tr->params + L.we1, tr->params + L.b1, tr->params + L.be1, tr->net.fw.p);
    ape::count_launch(2);
    return ape::check_launch("trainer sync_weights");
}

extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_create(float* params, float* grads, int num_obj, int max_batch, int max_points, ape_trainer** out)
{
    APE_REQUIRE(params && grads && out, "ape_refiner_trainer_create: null pointer");
    APE_REQUIRE(num_obj > 0 && max_batch > 0 && max_points > 0, "ape_refiner_trainer_create: bad sizes");
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
        ape::set_error("ape_refiner_trainer_create: no CUDA device (there is no CPU fallback)");
        return APE_ERR_CUDA;
    }
    ape_trainer* tr = new ape_trainer();
    tr->L = ape::train_layout(num_obj);
    tr->params = params; tr->grads = grads; tr->num_obj = num_obj; tr->max_batch = max_batch; tr->max_points = max_points;
    const ape::TrainLayout& L = tr->L;
    ape_net* net = &tr->net;
    net->kind = APE_NET_REFINER; net->num_obj = num_obj; net->max_batch = max_batch; net->max_points = max_points;
    net->np_max = (max_points + 127) / 128 * 128;
    net->gemm_impl = APE_GEMM_TCGEN05; net->train = 1;
    const size_t R = ((size_t)max_batch * net->np_max + 255) / 256 * 256;
    int rc = APE_OK;
#define TRY(x) do { if ((rc = (x)) != APE_OK) { ape_refiner_trainer_destroy(tr); return rc; } } while (0)
    // weights: fp32 views into the flat vector (SIMT layers), bf16 copies for the tensor-core layers
    net->b_c2e2.p = params + L.b2e2; net->b_c5.p = params + L.b5; net->b_c6.p = params + L.b6;
    net->Wr1.p = params + L.wh1; net->br1.p = params + L.bh1; net->Wr2.p = params + L.wh2; net->br2.p = params + L.bh2;
    net->w3r.p = params + L.w3r; net->b3r.p = params + L.b3r; net->w3t.p = params + L.w3t; net->b3t.p = params + L.b3t;
    TRY(alloc_f32(net, net->fw, 32 * 64 + 3 * 64 + 128));
    {
        bf16* wb = nullptr;
        TRY(dev_alloc(net, (void**)&wb, (size_t)(256 * 64 + 512 * 384 + 1024 * 512 + 1024 * 1024 + 256 * 512) * 2));
        tr->Wb2.p = wb; tr->Wb2.rows = 256; tr->Wb2.cols = 64;
        tr->Wb5.p = wb + 256 * 64; tr->Wb5.rows = 512; tr->Wb5.cols = 384;
        tr->Wb6.p = tr->Wb5.p + 512 * 384; tr->Wb6.rows = 1024; tr->Wb6.cols = 512;
        tr->Wbh1.p = tr->Wb6.p + 1024 * 512; tr->Wbh1.rows = 1024; tr->Wbh1.cols = 1024;
        tr->Wbh2.p = tr->Wbh1.p + 1024 * 1024; tr->Wbh2.rows = 256; tr->Wbh2.cols = 512;
        SplitMat* W[5] = {&net->W_c2e2, &net->W_c5, &net->W_c6, &net->Wh1b, &net->Wh2b};
        BfMat* Bm[5] = {&tr->Wb2, &tr->Wb5, &tr->Wb6, &tr->Wbh1, &tr->Wbh2};
        for (int i = 0; i < 5; ++i) {            // the forward kernel only dereferences the hi maps when passes == 1
            TRY(make_bf_maps(*Bm[i]));
            W[i]->hi = W[i]->lo = Bm[i]->p; W[i]->rows = Bm[i]->rows; W[i]->cols = Bm[i]->cols;
            W[i]->map_hi = W[i]->map_lo = Bm[i]->kmaj;
        }
    }
    {   // head activations and their gradients, bf16, rows padded to whole 128-row tiles (padding rows stay zero in the
        // gradient buffers, so they contribute nothing)
        tr->bp_max = (max_batch + 127) / 128 * 128;
        BfMat* Hm[6] = {&tr->APm, &tr->G1m, &tr->G2m, &tr->dZ2h, &tr->dZ1h, &tr->g6};
        const int hc[6] = {1024, 1024, 256, 256, 1024, 1024};
        for (int i = 0; i < 6; ++i) {
            TRY(dev_alloc(net, (void**)&Hm[i]->p, (size_t)tr->bp_max * hc[i] * 2));
            APE_CUDA(cudaMemset(Hm[i]->p, 0, (size_t)tr->bp_max * hc[i] * 2));
            Hm[i]->rows = tr->bp_max; Hm[i]->cols = hc[i];
            TRY(make_bf_maps(*Hm[i]));
        }
        SplitMat* S[3] = {&net->APb, &net->G1b, &net->G2b};
        for (int i = 0; i < 3; ++i) {
            S[i]->hi = S[i]->lo = Hm[i]->p; S[i]->rows = Hm[i]->rows; S[i]->cols = Hm[i]->cols;
            S[i]->map_hi = S[i]->map_lo = Hm[i]->kmaj; S[i]->st_hi = S[i]->st_lo = Hm[i]->st;
        }
    }
    TRY(alloc_split(net, net->PF, R, 384));      // the front end writes hi and lo; GEMM epilogues store hi only
    {
        SplitMat& H = net->H5;
        TRY(dev_alloc(net, (void**)&H.hi, R * 512 * 2));
        APE_CUDA(cudaMemset(H.hi, 0, R * 512 * 2));
        H.lo = H.hi; H.rows = (int)R; H.cols = 512;
        TRY(make_map(&H.map_hi, H.hi, R, 512, 512)); H.map_lo = H.map_hi;
        TRY(make_store_map(&H.st_hi, H.hi, R, 512, 512)); H.st_lo = H.st_hi;
    }
    TRY(alloc_f32(net, net->CS, (R / 128) * 1024));
    TRY(alloc_f32(net, net->AP, (size_t)max_batch * 1024));
    TRY(alloc_f32(net, net->G1, (size_t)max_batch * 1024));
    TRY(alloc_f32(net, net->G2, (size_t)max_batch * 256));
    TRY(dev_alloc(net, (void**)&net->relu_bits, R * 32 * sizeof(uint32_t)));
    APE_CUDA(cudaMemset(net->relu_bits, 0, R * 32 * sizeof(uint32_t)));
    tr->PFm.p = net->PF.hi; tr->PFm.rows = (int)R; tr->PFm.cols = 384; TRY(make_bf_maps(tr->PFm));
    tr->H5m.p = net->H5.hi; tr->H5m.rows = (int)R; tr->H5m.cols = 512; TRY(make_bf_maps(tr->H5m));
    BfMat* G[3] = {&tr->dY6, &tr->dZ5, &tr->dPF};
    const int gc[3] = {1024, 512, 384};
    for (int i = 0; i < 3; ++i) {
        TRY(dev_alloc(net, (void**)&G[i]->p, R * (size_t)gc[i] * 2));
        APE_CUDA(cudaMemset(G[i]->p, 0, R * (size_t)gc[i] * 2));
        G[i]->rows = (int)R; G[i]->cols = gc[i];
        TRY(make_bf_maps(*G[i]));
    }
    TRY(dev_alloc(net, (void**)&tr->part_b, (size_t)ape::tr::kBiasCopies * ape::kBiasPart * 4));
    TRY(dev_alloc(net, (void**)&tr->part_w1, (size_t)ape::kW1Copies * ape::kW1Part * 4));
    TRY(dev_alloc(net, (void**)&tr->part6, (R / 128) * 1024 * 4));
#undef TRY
    static ape::PerDevice attr_done;
    if (attr_done.first()) {
        cudaError_t e = cudaFuncSetAttribute(ape::tc2::gemm_split_bf16_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             ape::tc2::kSmemBytes2);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(ape::tc2::gemm_split_bf16_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     ape::tc2::kSmemBytes2);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(ape::tr::gemm_bf16_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ape::tr::kSmemBytesBwd);
        if (e != cudaSuccess) { ape::set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); ape_refiner_trainer_destroy(tr); return APE_ERR_CUDA; }
    }
    rc = ape_refiner_trainer_sync_weights(tr, nullptr);
    if (rc) { ape_refiner_trainer_destroy(tr); return rc; }
    APE_CUDA(cudaStreamSynchronize(nullptr));
    APE_CUDA(cudaEventCreateWithFlags(&tr->bulk_event, cudaEventDisableTiming));
    *out = tr;
    return APE_OK;
}

// Training forward: identical call surface to ape_refiner_forward; keeps the activations for the backward pass.
extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_forward(ape_trainer* tr, const float* new_points, const float* emb, const int64_t* obj, int B, int N,
                                float* r2, float* t2, void* stream)
{
    APE_REQUIRE(tr, "ape_refiner_trainer_forward: null handle");
    return ape_refiner_forward(&tr->net, new_points, emb, obj, B, N, r2, t2, stream);
}

// `mask` / `out`: the matrices behind p.mask / p.out (DGRAD only; their chunk maps feed the TMA-staged epilogue)
static int run_bwd_gemm(const BfMat& A, const BfMat& Bm, const BfMat* mask, const BfMat* out, ape::tr::BwdParams p, cudaStream_t s,
                        const char* label)
{
    ape::ProfScope prof_(label, s);
    const bool wgrad = p.mode == ape::tr::BWD_WGRAD;
    const int tiles = p.groups * (p.M / 128) * ((p.N + 255) / 256);
    if (wgrad) {
        const int kb = p.K / 64;
        int ks = ape::sm_count() / tiles;
        p.k_splits = ks < 1 ? 1 : (ks > kb ? kb : ks);
    } else {
        p.k_splits = 1;
    }
    const int total = tiles * p.k_splits;
    const int grid = total < ape::sm_count() ? total : ape::sm_count();
    if (!wgrad && (!out || (p.add_out && p.N != 64) || (p.mask_from % 64) != 0)) {
        ape::set_error("run_bwd_gemm: unsupported DGRAD epilogue configuration");
        return APE_ERR_UNSUPPORTED;
    }
    p.b_box_chunks = Bm.mn4_chunks;
    ape::tr::gemm_bf16_bwd_kernel<<<grid, ape::tr::kThreadsBwd, ape::tr::kSmemBytesBwd, s>>>(
        wgrad ? A.mn2 : A.kmaj, Bm.mn4, mask ? mask->st : A.kmaj, out ? out->st : A.kmaj, p);
    ape::count_launch();
    return ape::check_launch(label);
}

// Backward of the forward that ape_refiner_trainer_forward just ran (same new_points / emb / obj / B / N):
// grads += d(sum_b <d_r[b], r2[b]> + <d_t[b], t2[b]>) / d params.
extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_backward(ape_trainer* tr, const float* new_points, const float* emb, const int64_t* obj, int B, int N,
                                 const float* d_r, const float* d_t, void* stream)
{
    APE_REQUIRE(tr && new_points && emb && obj && d_r && d_t, "ape_refiner_trainer_backward: null pointer");
    APE_REQUIRE(B > 0 && N > 0 && B <= tr->max_batch && N <= tr->max_points, "ape_refiner_trainer_backward: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    ape_net* net = &tr->net;
    const ape::TrainLayout& L = tr->L;
    float* G = tr->grads;
    const int Np = (N + 127) / 128 * 128, M = (B * Np + 255) / 256 * 256;
    int rc;
    APE_CUDA(cudaMemsetAsync(tr->part_b, 0, (size_t)ape::tr::kBiasCopies * ape::kBiasPart * 4, s));
    APE_CUDA(cudaMemsetAsync(tr->part_w1, 0, (size_t)ape::kW1Copies * ape::kW1Part * 4, s));
    const int Bp = (B + 127) / 128 * 128;
    ape::tr::BwdParams p;
    {   // heads: conv3 (class rows, SIMT) then conv2_{r,t} / conv1_{r,t} on the tensor cores (bf16, K or M = padded batch)
        ape::ProfScope prof_("train.head3_bwd", s);
        ape::head3_bwd_kernel<<<B, 256, 0, s>>>(d_r, d_t, obj, tr->num_obj, tr->G2m.p, net->w3r.p, net->w3t.p, tr->dZ2h.p,
                                               G + L.w3r, G + L.w3t, G + L.b3r, G + L.b3t, G + L.bh2);
        ape::count_launch();
        if ((rc = ape::check_launch("head3_bwd"))) return rc;
    }
    // dW(conv2_{r,t}) [2 x 128, 512] += dZ2h[:, g]^T * G1[:, g]
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_WGRAD; p.M = 128; p.N = 512; p.K = Bp; p.groups = 2; p.a_cg = 128; p.b_cg = 512;
    p.dw = G + L.wh2; p.dw_ld = 512; p.dw_rg = 128;
    if ((rc = run_bwd_gemm(tr->dZ2h, tr->G1m, nullptr, nullptr, p, s, "gemm.train.head2_wgrad"))) return rc;
    // dZ1h [Bp, 2 x 512] = (dZ2h[:, g] * W2[g]) masked by G1 > 0, bias gradient of conv1_{r,t}
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_DGRAD; p.M = Bp; p.N = 512; p.K = 128; p.groups = 2; p.a_cg = 128; p.b_rg = 128;
    p.out = tr->dZ1h.p; p.o_ld = 1024; p.o_cg = 512; p.mask = tr->G1m.p; p.m_ld = 1024; p.m_cg = 512;
    p.bias_grad = tr->part_b + 1024; p.bg_stride = ape::kBiasPart;
    if ((rc = run_bwd_gemm(tr->dZ2h, tr->Wbh2, &tr->G1m, &tr->dZ1h, p, s, "gemm.train.head2_dgrad"))) return rc;
    // dW(conv1_{r,t}) [1024, 1024] += dZ1h^T * AP
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_WGRAD; p.M = 1024; p.N = 1024; p.K = Bp; p.groups = 1; p.dw = G + L.wh1; p.dw_ld = 1024;
    if ((rc = run_bwd_gemm(tr->dZ1h, tr->APm, nullptr, nullptr, p, s, "gemm.train.head1_wgrad"))) return rc;
    // dAP [Bp, 1024] = dZ1h * W1   (AvgPool1d's 1/N is applied where dY6 is formed)
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_DGRAD; p.M = Bp; p.N = 1024; p.K = 1024; p.groups = 1; p.out = tr->g6.p; p.o_ld = 1024;
    if ((rc = run_bwd_gemm(tr->dZ1h, tr->Wbh1, nullptr, &tr->g6, p, s, "gemm.train.head1_dgrad"))) return rc;
    {
        ape::ProfScope prof_("train.dy6", s);
        ape::dy6_kernel<<<M / 128, 256, 0, s>>>(net->relu_bits, tr->g6.p, 1.0f / (float)N, Np, B * Np, tr->dY6.p, tr->part6);
        ape::count_launch();
        if ((rc = ape::check_launch("dy6"))) return rc;
    }
    // conv6: dW6 [1024, 512] += dY6^T * H5 ;  dZ5 = (dY6 * W6) masked by H5 > 0, db5
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_WGRAD; p.M = 1024; p.N = 512; p.K = M; p.groups = 1; p.dw = G + L.w6; p.dw_ld = 512;
    if ((rc = run_bwd_gemm(tr->dY6, tr->H5m, nullptr, nullptr, p, s, "gemm.train.wgrad6"))) return rc;
    {   // the conv6 / conv1_{r,t} bias partials are final too: fold them now, then the whole tail block of the gradient is done
        ape::ProfScope prof_("train.fold_partials", s);
        ape::fold_partials_kernel<<<32 + 5, 1024, 0, s>>>(tr->part_b, tr->part_w1, tr->part6, M / 128, G + L.b5, G + L.b2e2,
                                                       G + L.b1, G + L.w1, G + L.b6, G + L.bh1, 0);
        ape::count_launch();
        if ((rc = ape::check_launch("fold_partials (tail block)"))) return rc;
        if (tr->record_bulk && tr->bulk_event) APE_CUDA(cudaEventRecord(tr->bulk_event, s));
    }
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_DGRAD; p.M = M; p.N = 512; p.K = 1024; p.groups = 1; p.out = tr->dZ5.p; p.o_ld = 512;
    p.mask = tr->H5m.p; p.m_ld = 512; p.bias_grad = tr->part_b; p.bg_stride = ape::kBiasPart;
    if ((rc = run_bwd_gemm(tr->dY6, tr->Wb6, &tr->H5m, &tr->dZ5, p, s, "gemm.train.dgrad6"))) return rc;
    // conv5 (input = pointfeat_3 = PF[:, 0:384], network.py:160-162)
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_WGRAD; p.M = 512; p.N = 384; p.K = M; p.groups = 1; p.dw = G + L.w5; p.dw_ld = 384;
    if ((rc = run_bwd_gemm(tr->dZ5, tr->PFm, nullptr, nullptr, p, s, "gemm.train.wgrad5"))) return rc;
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_DGRAD; p.M = M; p.N = 384; p.K = 512; p.groups = 1; p.out = tr->dPF.p; p.o_ld = 384;
    p.mask = tr->PFm.p; p.m_ld = 384; p.mask_from = 128;
    p.bias_grad = tr->part_b + 512 - 128; p.bg_stride = ape::kBiasPart;          // columns 128..383 = conv2 | e_conv2
    if ((rc = run_bwd_gemm(tr->dZ5, tr->Wb5, &tr->PFm, &tr->dPF, p, s, "gemm.train.dgrad5"))) return rc;
    // conv2 | e_conv2 (two groups): dW [2 x 128, 64] += dZ[:, 128 + g*128 ...]^T * PF[:, g*64 ...]; then
    // dPF[:, g*64 ...] = (dZ * W_g + the conv5 share already there) masked by PF > 0, db1 | dbe1
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_WGRAD; p.M = 128; p.N = 64; p.K = M; p.groups = 2; p.a_c0 = 128; p.a_cg = 128; p.b_cg = 64;
    p.dw = G + L.w2e2; p.dw_ld = 64; p.dw_rg = 128;
    if ((rc = run_bwd_gemm(tr->dPF, tr->PFm, nullptr, nullptr, p, s, "gemm.train.wgrad2"))) return rc;
    memset(&p, 0, sizeof(p));
    p.mode = ape::tr::BWD_DGRAD; p.M = M; p.N = 64; p.K = 128; p.groups = 2; p.a_c0 = 128; p.a_cg = 128; p.b_rg = 128;
    p.out = tr->dPF.p; p.o_ld = 384; p.o_cg = 64; p.add_out = 1; p.mask = tr->PFm.p; p.m_ld = 384; p.m_cg = 64;
    p.bias_grad = tr->part_b + 768; p.bg_stride = ape::kBiasPart;
    if ((rc = run_bwd_gemm(tr->dPF, tr->Wb2, &tr->PFm, &tr->dPF, p, s, "gemm.train.dgrad2"))) return rc;
    {
        ape::ProfScope prof_("train.conv1_wgrad", s);
        ape::conv1_wgrad_kernel<<<5 * ape::sm_count(), 256, 0, s>>>(tr->dPF.p, 384, new_points, emb, B, N, Np, tr->part_w1,
                                                                   tr->part_w1 + 64 * 3);
        ape::count_launch();
        if ((rc = ape::check_launch("conv1_wgrad"))) return rc;
    }
    {
        ape::ProfScope prof_("train.fold_partials", s);
        ape::fold_partials_kernel<<<32 + 5, 1024, 0, s>>>(tr->part_b, tr->part_w1, tr->part6, M / 128, G + L.b5, G + L.b2e2,
                                                       G + L.b1, G + L.w1, G + L.b6, G + L.bh1, 1);
        ape::count_launch();
        if ((rc = ape::check_launch("fold_partials"))) return rc;
    }
    return APE_OK;
}

extern "C" __attribute__((visibility("default")))
int ape_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, void* stream)
{
    APE_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step > 0, "ape_adam_step: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const float bc1 = 1.f - powf(beta1, (float)step), bc2s = sqrtf(1.f - powf(beta2, (float)step));
    ape::ProfScope prof_("train.adam", s);
    ape::adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, (size_t)n, lr, beta1, beta2, eps,
                                                                 bc1, bc2s, grad_scale);
    ape::count_launch();
    return ape::check_launch("adam");
}


// Whole accumulation phase of one optimizer step for B objects (train.py:215-223 batched): zero the flat gradient, then
// `iterations` x (training forward -> Loss_refine forward + backward -> backward), the cloud / target re-expressed in the
// predicted frame between iterations.  One host call, no host synchronisation, graph-capturable after the first call
// (the first call allocates scratch).  dis [iterations, B] out.
extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_step(ape_trainer* tr, const float* points, const float* emb, const int64_t* obj, const float* target,
                             const float* model_points, const uint8_t* symmetric, int B, int N, int n_mesh, int iterations,
                             int zero_grad, float* dis, void* stream)
{
    APE_REQUIRE(tr && points && emb && obj && target && model_points && dis, "ape_refiner_trainer_step: null pointer");
    APE_REQUIRE(B > 0 && N > 0 && n_mesh > 0 && iterations > 0, "ape_refiner_trainer_step: bad sizes");
    APE_REQUIRE(B <= tr->max_batch && N <= tr->max_points, "ape_refiner_trainer_step: B=%d N=%d exceed the trainer's workspace (%d, %d)",
                B, N, tr->max_batch, tr->max_points);
    cudaStream_t s = (cudaStream_t)stream;
    ape_net* net = &tr->net;
    if (!tr->s_r || tr->s_mesh < n_mesh) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &cap);
        APE_REQUIRE(cap == cudaStreamCaptureStatusNone, "ape_refiner_trainer_step: first call (scratch allocation) must not be captured");
        int rc;
        const size_t mb = (size_t)tr->max_batch;
        if (!tr->s_r) {
            if ((rc = dev_alloc(net, (void**)&tr->s_r, mb * 4 * 4)) || (rc = dev_alloc(net, (void**)&tr->s_t, mb * 3 * 4)) ||
                (rc = dev_alloc(net, (void**)&tr->s_dr, mb * 4 * 4)) || (rc = dev_alloc(net, (void**)&tr->s_dt, mb * 3 * 4)) ||
                (rc = dev_alloc(net, (void**)&tr->s_pts[0], mb * tr->max_points * 3 * 4)) ||
                (rc = dev_alloc(net, (void**)&tr->s_pts[1], mb * tr->max_points * 3 * 4))) return rc;
        }
        if ((rc = dev_alloc(net, (void**)&tr->s_tgt[0], mb * n_mesh * 3 * 4)) || (rc = dev_alloc(net, (void**)&tr->s_tgt[1], mb * n_mesh * 3 * 4)))
            return rc;                                        // (a smaller earlier pair stays owned by the handle until destroy)
        tr->s_mesh = n_mesh;
    }
    if (zero_grad) APE_CUDA(cudaMemsetAsync(tr->grads, 0, tr->L.total * sizeof(float), s));
    const float* p_cur = points;
    const float* t_cur = target;
    for (int it = 0; it < iterations; ++it) {
        int rc = ape_refiner_trainer_forward(tr, p_cur, emb, obj, B, N, tr->s_r, tr->s_t, stream);
        if (rc) return rc;
        const bool last = it == iterations - 1;
        float* p_nxt = last ? nullptr : tr->s_pts[it & 1];
        float* t_nxt = last ? nullptr : tr->s_tgt[it & 1];
        rc = ape_refine_loss(tr->s_r, tr->s_t, model_points, t_cur, n_mesh, p_cur, N, symmetric, B, dis + (size_t)it * B, tr->s_dr, tr->s_dt,
                             p_nxt, t_nxt, stream);
        if (rc) return rc;
        tr->record_bulk = last ? 1 : 0;                     // the tail block is final after the LAST iteration's conv6 weight gradient
        rc = ape_refiner_trainer_backward(tr, p_cur, emb, obj, B, N, tr->s_dr, tr->s_dt, stream);
        tr->record_bulk = 0;
        if (rc) return rc;
        if (!last) { p_cur = p_nxt; t_cur = t_nxt; }
    }
    return APE_OK;
}

// Data-parallel overlap: [bulk_begin, total) of the flat gradient (conv6 + heads: 89 %) is final when the last iteration's
// conv6 weight gradient has been written, before the backward pass walks conv5 / conv2 / conv1.  ape_refiner_trainer_step
// records an event there; ape_refiner_trainer_wait_bulk makes `side_stream` wait for it, so that the caller's all-reduce of
// the tail block on that stream overlaps the rest of the backward pass (the head block follows on the main stream).
extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_wait_bulk(ape_trainer* tr, void* side_stream, int64_t* bulk_begin)
{
    APE_REQUIRE(tr, "ape_refiner_trainer_wait_bulk: null handle");
    if (bulk_begin) *bulk_begin = (int64_t)tr->L.bulk;
    APE_REQUIRE(tr->bulk_event, "ape_refiner_trainer_wait_bulk: trainer without event");
    APE_CUDA(cudaStreamWaitEvent((cudaStream_t)side_stream, tr->bulk_event, 0));     // never recorded yet: a no-op
    return APE_OK;
}

// Adam on the trainer's own flat vectors followed by the bf16 weight refresh: the tail of an optimizer step in one call.
extern "C" __attribute__((visibility("default")))
int ape_refiner_trainer_adam(ape_trainer* tr, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2, float eps, int step,
                             float grad_scale, void* stream)
{
    APE_REQUIRE(tr, "ape_refiner_trainer_adam: null handle");
    int rc = ape_adam_step(tr->params, tr->grads, exp_avg, exp_avg_sq, (int64_t)tr->L.total, lr, beta1, beta2, eps, step, grad_scale, stream);
    if (rc) return rc;
    return ape_refiner_trainer_sync_weights(tr, stream);
}
