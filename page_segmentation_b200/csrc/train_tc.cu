// Training step of the FCN graphs on the tensor cores (BASELINE configs[4]; ocr4all_pixel_classifier/lib/network.py:127-242
// `train_dataset` with batch 1, lib/metrics.py:8-9 `loss`, lib/model.py:45-92 / :206-234 the graphs).
//
// Mixed precision: fp32 master weights, gradients and optimizer state (train.cu's Adam); activations and activation
// gradients are bf16 in the inference path's plane-major layout [C/8][H][W][8] (one page per step, as the reference
// trains), every product is accumulated in fp32 in TMEM.
//
//   forward            conv_umma_kernel (conv_umma.cu) over operand images rebuilt from the master weights each step
//   input gradients    the same implicit GEMM with swapped roles: dX = corr(dY, W^T flipped); the 2x2 stride-2 transposed
//                      convolutions become 1x1 GEMMs over the space-to-depth form of dY ([tap][C][h/2][w/2], written
//                      by whoever produces that gradient, never by a separate transposition)
//   weight gradients   wgrad_tc_kernel below: a pixel-reduction GEMM, K = pixels.  The plane-major layout IS the
//                      MN-major SWIZZLE_NONE canonical operand layout of tcgen05 (8 pixels x 8 channels = one 128-byte
//                      core matrix), so X rows and dY rows go from TMA straight into the MMA: A = X (M = vertical taps
//                      x input-channel planes, the tap shift is a +16-byte start address), B = dY (N = output channels),
//                      one accumulator per horizontal tap, 5 x N <= 512 TMEM columns.  Partial sums of the CTAs are
//                      combined with fp32 atomics into the flat gradient buffer.
//   loss               logits 1x1 + softmax cross entropy + d logits + the logits layer's input gradient in one pass
//
// Concatenations are plane-aligned buffers (every source padded to whole 8-channel planes), a skip tensor is a plane
// range of the buffer of the concatenation it feeds, and its gradient is the matching plane range of that buffer's
// gradient.
#include "common.cuh"
#include "umma_ptx.cuh"

#include <algorithm>

namespace pcs {
namespace {

using namespace ptx;
typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack2<bf16>(f[0], f[1]), pack2<bf16>(f[2], f[3]), pack2<bf16>(f[4], f[5]), pack2<bf16>(f[6], f[7]));
}

// x/255 of the uint8 page into channel 0 of a zero-padded one-plane tensor (architecture.py:67-68, model.py:20-26)
__global__ void __launch_bounds__(256) tc_input_kernel(const uint8_t* __restrict__ img, int h, int w, uint4* __restrict__ out, int H, int W) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)H * W; i += (size_t)gridDim.x * 256) {
        const int r = (int)(i / W), c = (int)(i % W);
        const float v = (r < h && c < w) ? (float)img[(size_t)r * w + c] / 255.0f : 0.f;
        out[i] = make_uint4(pack2<bf16>(v, 0.f), 0u, 0u, 0u);
    }
}

// Operand images of conv_umma_kernel, [ntile][chunk][tap][plane][npad][8], from the fp32 master weights:
// element (tap t, padded input channel c, column n) = w[in_off[c] + out_off[n] + t' * s_tap], t' = flip ? taps-1-t : t
struct WimgDesc {
    const float* w;
    uint16_t* dst;
    const int* in_off;
    const int* out_off;
    int taps, nchunks, npad, ntiles, s_tap, flip;
    unsigned long long total;
};
__global__ void __launch_bounds__(256) tc_wimg_kernel(const WimgDesc* __restrict__ descs) {
    const WimgDesc d = descs[blockIdx.y];
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < d.total; i += (unsigned long long)gridDim.x * 256) {
        const int e = (int)(i & 7);
        unsigned long long r = i >> 3;
        const int nn = (int)(r % d.npad); r /= d.npad;
        const int pl = (int)(r & 1); r >>= 1;
        const int t = (int)(r % d.taps); r /= d.taps;
        const int chunk = (int)(r % d.nchunks);
        const int nt = (int)(r / d.nchunks);
        const int io = d.in_off[chunk * 16 + pl * 8 + e], oo = d.out_off[nt * d.npad + nn];
        const float v = (io >= 0 && oo >= 0) ? __ldg(d.w + io + oo + (d.flip ? d.taps - 1 - t : t) * d.s_tap) : 0.f;
        const bf16 b = __float2bfloat16_rn(v);
        d.dst[i] = *reinterpret_cast<const uint16_t*>(&b);
    }
}

// dst = [y > 0] * (a + b) on plane-major tensors of `planes` x h x w 16-byte units (b, y optional).  s2d: dst is the
// space-to-depth form [tap = (y&1)*2 + (x&1)][plane][h/2][w/2] the stride-2 layers' backward kernels read.
__global__ void __launch_bounds__(256)
tc_combine_kernel(uint4* __restrict__ dst, const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ y,
                  int planes, int h, int w, int s2d) {
    const size_t units = (size_t)planes * h * w;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < units; i += (size_t)gridDim.x * 256) {
        float fa[8];
        unpack8(a[i], fa);
        if (b) {
            float fb[8];
            unpack8(b[i], fb);
#pragma unroll
            for (int k = 0; k < 8; ++k) fa[k] += fb[k];
        }
        if (y) {
            float fy[8];
            unpack8(y[i], fy);
#pragma unroll
            for (int k = 0; k < 8; ++k) fa[k] = fy[k] > 0.f ? fa[k] : 0.f;
        }
        size_t o = i;
        if (s2d) {
            const int q = (int)(i / ((size_t)h * w));
            const size_t rem = i - (size_t)q * h * w;
            const int yy = (int)(rem / w), xx = (int)(rem % w);
            const int tap = (yy & 1) * 2 + (xx & 1);
            o = (((size_t)(tap * planes + q) * (h >> 1)) + (yy >> 1)) * (w >> 1) + (xx >> 1);
        }
        dst[o] = pack8(fa);
    }
}

// MaxPooling2D(2,2) backward on plane-major tensors: the gradient of a pooled pixel goes to the first maximum of its window
// in row-major order (TensorFlow / torch convention; train.cu does the same in fp32); `skip` (optional, may alias dst) is
// added: the gradient that reaches the same tensor through a skip connection.
__global__ void __launch_bounds__(256)
tc_pool_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ g, const uint4* skip, uint4* dst, int planes, int h, int w) {
    const int h2 = h >> 1, w2 = w >> 1;
    const size_t units = (size_t)planes * h2 * w2;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < units; i += (size_t)gridDim.x * 256) {
        const int q = (int)(i / ((size_t)h2 * w2));
        const size_t rem = i - (size_t)q * h2 * w2;
        const int r = (int)(rem / w2), c = (int)(rem % w2);
        const size_t o00 = ((size_t)q * h + 2 * r) * w + 2 * c;
        const size_t off[4] = {o00, o00 + 1, o00 + w, o00 + w + 1};
        float v[4][8], gg[8], out[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack8(y[off[k]], v[k]);
        unpack8(g[i], gg);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (skip) unpack8(skip[off[k]], out[k]);
            else {
#pragma unroll
                for (int e = 0; e < 8; ++e) out[k][e] = 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int best = 0;
            float bv = v[0][e];
#pragma unroll
            for (int k = 1; k < 4; ++k) if (v[k][e] > bv) { bv = v[k][e]; best = k; }
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k][e] += (k == best) ? gg[e] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[off[k]] = pack8(out[k]);
    }
}

// db[c] += sum over the pixels of channel c; g has `planes` planes of hw units, plane p holds channels (p % qp) * 8 .. + 7
// (qp < planes: the space-to-depth form, whose taps fold into the same channel).  grid = (planes, blocks per plane)
__global__ void __launch_bounds__(256) tc_bias_grad_kernel(const uint4* __restrict__ g, size_t hw, int qp, int co, float* __restrict__ db) {
    const uint4* p = g + (size_t)blockIdx.x * hw;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (size_t i = (size_t)blockIdx.y * 256 + threadIdx.x; i < hw; i += (size_t)gridDim.y * 256) {
        float f[8];
        unpack8(p[i], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    __shared__ float s[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = acc[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = 0.f;
        for (int k = 0; k < 8; ++k) v += s[k][threadIdx.x];
        const int c = ((int)blockIdx.x % qp) * 8 + (int)threadIdx.x;
        if (c < co && v != 0.f) atomicAdd(db + c, v);
    }
}

// Logits layer + loss (model.py:88, metrics.py:8-9) in one pass over the pixels: logits = b + W a over the concatenated
// input planes, softmax cross entropy inside the crop, d logits = (softmax - onehot) / (hc * wc), and the input gradient
// W^T d logits written where its consumers want it: the deconv5 planes in space-to-depth form, the conv2 planes (skip)
// as a full-resolution tensor.  d logits are kept (float4 per pixel) for the weight-gradient pass below.
constexpr int TC_NC = 4, TC_LOGIT_CP = 56;
struct LogitsParams {
    const uint4* a;              // [planes][H][W]
    int planes, planes_d5;       // all input planes / those of deconv5 (the rest is conv2)
    int H, W, hc, wc, ncls, ci;
    const float* w;              // [ncls][ci]
    const float* bias;
    const int* ch_map;           // padded channel -> real channel or -1
    const uint8_t* labels;       // [hc][wc]
    float4* dl;                  // [H][W]
    uint4* g_d5s;                // [4][planes_d5][H/2][W/2]
    uint4* g_skip;               // [planes - planes_d5][H][W] or null
    double* loss_sum;
};
__global__ void __launch_bounds__(256) tc_logits_kernel(const LogitsParams p) {
    __shared__ float s_w[TC_LOGIT_CP][TC_NC];
    __shared__ float s_b[TC_NC];
    for (int i = threadIdx.x; i < TC_LOGIT_CP * TC_NC; i += 256) {
        const int c = i / TC_NC, k = i % TC_NC;
        const int rc = c < p.planes * 8 ? p.ch_map[c] : -1;
        s_w[c][k] = (rc >= 0 && k < p.ncls) ? p.w[(size_t)k * p.ci + rc] : 0.f;
    }
    if (threadIdx.x < TC_NC) s_b[threadIdx.x] = threadIdx.x < p.ncls ? p.bias[threadIdx.x] : 0.f;
    __syncthreads();
    const size_t hw = (size_t)p.H * p.W;
    const float inv = 1.f / ((float)p.hc * (float)p.wc);
    double local = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < hw; i += (size_t)gridDim.x * 256) {
        const int r = (int)(i / p.W), c = (int)(i % p.W);
        float lg[TC_NC];
#pragma unroll
        for (int k = 0; k < TC_NC; ++k) lg[k] = s_b[k];
        for (int q = 0; q < p.planes; ++q) {
            float f[8];
            unpack8(__ldg(p.a + (size_t)q * hw + i), f);
#pragma unroll
            for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int k = 0; k < TC_NC; ++k) lg[k] = fmaf(f[e], s_w[q * 8 + e][k], lg[k]);
        }
        float dl[TC_NC] = {0.f, 0.f, 0.f, 0.f};
        if (r < p.hc && c < p.wc) {
            float mx = lg[0];
            for (int k = 1; k < p.ncls; ++k) mx = fmaxf(mx, lg[k]);
            float sum = 0.f;
            for (int k = 0; k < p.ncls; ++k) sum += expf(lg[k] - mx);
            const int lab = p.labels[(size_t)r * p.wc + c];
            const float lse = mx + logf(sum);
            float lg_lab = 0.f;
#pragma unroll
            for (int k = 0; k < TC_NC; ++k) if (k == lab) lg_lab = lg[k];
            local += (double)(lse - lg_lab);
#pragma unroll
            for (int k = 0; k < TC_NC; ++k) dl[k] = k < p.ncls ? (expf(lg[k] - lse) - (k == lab ? 1.f : 0.f)) * inv : 0.f;
        }
        p.dl[i] = make_float4(dl[0], dl[1], dl[2], dl[3]);
        const int tap = (r & 1) * 2 + (c & 1);
        const size_t hw4 = hw >> 2;
        const size_t s2d_pix = (size_t)(r >> 1) * (p.W >> 1) + (c >> 1);
        for (int q = 0; q < p.planes; ++q) {
            float gch[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < TC_NC; ++k) v = fmaf(dl[k], s_w[q * 8 + e][k], v);
                gch[e] = v;
            }
            if (q < p.planes_d5) p.g_d5s[(size_t)(tap * p.planes_d5 + q) * hw4 + s2d_pix] = pack8(gch);
            else p.g_skip[(size_t)(q - p.planes_d5) * hw + i] = pack8(gch);
        }
    }
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(p.loss_sum, local);
}

// dW[k][c] += sum_px a[px][c] * dl[px][k], db[k] += sum_px dl[px][k]; grid = (input planes, blocks per plane)
__global__ void __launch_bounds__(256)
tc_logits_wgrad_kernel(const uint4* __restrict__ a, const float4* __restrict__ dl, size_t hw, int ncls, int ci, const int* __restrict__ ch_map,
                       float* __restrict__ dw, float* __restrict__ db) {
    const uint4* ap = a + (size_t)blockIdx.x * hw;
    float acc[8][TC_NC], bsum[TC_NC];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int k = 0; k < TC_NC; ++k) acc[e][k] = 0.f;
#pragma unroll
    for (int k = 0; k < TC_NC; ++k) bsum[k] = 0.f;
    for (size_t i = (size_t)blockIdx.y * 256 + threadIdx.x; i < hw; i += (size_t)gridDim.y * 256) {
        const float4 d = __ldg(dl + i);
        const float dk[TC_NC] = {d.x, d.y, d.z, d.w};
        float f[8];
        unpack8(__ldg(ap + i), f);
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
            for (int k = 0; k < TC_NC; ++k) acc[e][k] = fmaf(f[e], dk[k], acc[e][k]);
#pragma unroll
        for (int k = 0; k < TC_NC; ++k) bsum[k] += dk[k];
    }
    __shared__ float s[8][8 * TC_NC + TC_NC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int k = 0; k < TC_NC; ++k) {
            float v = acc[e][k];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s[warp][e * TC_NC + k] = v;
        }
#pragma unroll
    for (int k = 0; k < TC_NC; ++k) {
        float v = bsum[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s[warp][8 * TC_NC + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8 * TC_NC + TC_NC) {
        float v = 0.f;
        for (int k = 0; k < 8; ++k) v += s[k][threadIdx.x];
        if (threadIdx.x < 8 * TC_NC) {
            const int e = threadIdx.x / TC_NC, k = threadIdx.x % TC_NC;
            const int rc = ch_map[blockIdx.x * 8 + e];
            if (rc >= 0 && k < ncls) atomicAdd(dw + (size_t)k * ci + rc, v);
        } else if (blockIdx.x == 0) {
            const int k = threadIdx.x - 8 * TC_NC;
            if (k < ncls) atomicAdd(db + k, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradients on tcgen05: dW[dy][dx][ci][co] = sum_{r,c} X[r+dy-pad][c+dx-pad][ci] * dY[r][c][co]
// ---------------------------------------------------------------------------------------------------------------------
// One CTA = one group of vertical taps (blockIdx.y: rows d0 .. d0+nr-1 of the kernel) over a share of the page, walked as
// (strip of DT pixels) x (band of B rows).  Per band the X rows it needs ([row][plane][XT pixels], one TMA box per row)
// sit in shared memory (XT = DT + the halo, rounded up so that every row starts 128-byte aligned); dY rows ([plane][DT pixels]) stream through a small ring.  For a dY row and a 16-pixel K step
//   A = 128 rows of the X block starting at that row: 16 groups of 8 channels, group g = (vertical tap g / P, plane g % P),
//       MN-major, SBO = XT * 16 bytes (one plane row), LBO = 128 bytes (8 pixels); horizontal tap dx = +16 bytes
//   B = the dY row: N = output channels, MN-major, SBO = DT * 16 bytes
//   D[dx] += A B   (KX accumulators of N columns)
// Groups beyond nr * P read whatever follows in shared memory; those accumulator rows are never looked at.
constexpr int kWgThreads = 192;
constexpr int kWgMaxSlots = 4;

struct WgradParams {
    int H, W;
    int P, npad, nq;            // X planes, N, N / 8
    int pad;                    // (k - 1) / 2
    int B, xr;                  // dY rows per band, X rows per block (B + max nr - 1)
    int strips, bands;
    int nslots;
    uint32_t x_row_bytes, x_alloc, dy_slot_bytes;
    int ntiles_m, d0[5], nr[5];
    float* acc;                 // [ntiles_m][KX][npad][128] fp32 partial sums of all CTAs (zeroed by the caller), scattered into
                                // the weight layout by tc_wgrad_scatter_kernel
    int swap_strides;           // diagnosis: exchange LBO and SBO
};

// acc[tile][dx][n][m] -> dw[m_off[channel of row m] + n_off[n] + ((d0 + j) * kx + dx) * s_tap] += value
struct ScatterDesc {
    const float* acc;
    float* dw;
    const int* m_off;           // [P * 8]: X padded channel -> element offset of its weights or -1
    const int* n_off;           // [npad]
    int P, npad, kx, ntiles_m, d0[5], nr[5], s_tap;
    unsigned total;
};
__global__ void __launch_bounds__(256) tc_wgrad_scatter_kernel(const ScatterDesc* __restrict__ descs) {
    const ScatterDesc d = descs[blockIdx.y];
    for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < d.total; i += gridDim.x * 256) {
        const int m = (int)(i & 127);
        unsigned r = i >> 7;
        const int n = (int)(r % d.npad); r /= d.npad;
        const int dx = (int)(r % d.kx);
        const int tile = (int)(r / d.kx);
        const int g = m >> 3, j = g / d.P, pl = g - j * d.P;
        if (j >= d.nr[tile]) continue;
        const int moff = d.m_off[pl * 8 + (m & 7)], noff = d.n_off[n];
        if (moff < 0 || noff < 0) continue;
        const float v = d.acc[i];
        if (v != 0.f) d.dw[moff + noff + ((d.d0[tile] + j) * d.kx + dx) * d.s_tap] += v;
    }
}

template <int KX>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const WgradParams p, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmdy) {
    constexpr int DT = KX == 5 ? 112 : 128, XT = KX == 5 ? 120 : 128, KSTEPS = DT / 16;      // XT * 16 B: a multiple of the 128-byte TMA alignment
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_xfull, s_xempty, s_dyfull[kWgMaxSlots], s_dyempty[kWgMaxSlots], s_done;
    __shared__ uint32_t s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* xblk = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* dyring = xblk + p.x_alloc;

    if (warp == 0 && lane == 0) {
        mbar_init(&s_xfull, 1); mbar_init(&s_xempty, 1); mbar_init(&s_done, 1);
        for (int s = 0; s < kWgMaxSlots; ++s) { mbar_init(&s_dyfull[s], 1); mbar_init(&s_dyempty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const int d0 = p.d0[blockIdx.y], nr = p.nr[blockIdx.y];
    const int items = p.strips * p.bands;

    if (warp == 0) {
        if (lane == 0) {
            int slot = 0;
            uint32_t xph = 0, dph = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int strip = item % p.strips, band = item / p.strips;
                const int x0 = strip * DT, r0 = band * p.B;
                mbar_wait(&s_xempty, xph ^ 1u);                  // the MMAs of the previous band have left the block
                xph ^= 1u;
                mbar_expect_tx(&s_xfull, (uint32_t)p.xr * p.x_row_bytes);
                for (int j = 0; j < p.xr; ++j)
                    tma_load_3d(xblk + (size_t)j * p.x_row_bytes, &tmx, &s_xfull, (x0 - p.pad) * 2, r0 + d0 - p.pad + j, 0);
                for (int rb = 0; rb < p.B; ++rb) {
                    mbar_wait(&s_dyempty[slot], dph ^ 1u);
                    mbar_expect_tx(&s_dyfull[slot], p.dy_slot_bytes);
                    tma_load_3d(dyring + (size_t)slot * p.dy_slot_bytes, &tmdy, &s_dyfull[slot], x0 * 2, r0 + rb, 0);
                    if (++slot == p.nslots) { slot = 0; dph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.npad >> 3) << 17) | (8u << 24);
        const uint32_t a_sbo = (uint32_t)XT * 16u, b_sbo = (uint32_t)DT * 16u, lbo = 128u;
        const uint32_t a_hi = (uint32_t)(make_desc(0, 0, p.swap_strides ? lbo : a_sbo) >> 32);
        const uint32_t b_hi = (uint32_t)(make_desc(0, 0, p.swap_strides ? lbo : b_sbo) >> 32);
        const uint32_t a_lo_c = (((p.swap_strides ? a_sbo : lbo) >> 4) & 0x3fffu) << 16;
        const uint32_t b_lo_c = (((p.swap_strides ? b_sbo : lbo) >> 4) & 0x3fffu) << 16;
        int slot = 0;
        uint32_t xph = 0, dph = 0;
        bool first = true;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            mbar_wait(&s_xfull, xph);
            xph ^= 1u;
            for (int rb = 0; rb < p.B; ++rb) {
                mbar_wait(&s_dyfull[slot], dph);
                tc_fence_after();
                if (leader) {
                    const uint32_t a_row = smem_u32(xblk) + (uint32_t)rb * p.x_row_bytes;
                    const uint32_t b_row = smem_u32(dyring) + (uint32_t)slot * p.dy_slot_bytes;
#pragma unroll
                    for (int s = 0; s < KSTEPS; ++s) {
                        const uint32_t b_lo = (((b_row + (uint32_t)s * 256u) >> 4) & 0x3fffu) | b_lo_c;
#pragma unroll
                        for (int dx = 0; dx < KX; ++dx) {
                            const uint32_t a_lo = (((a_row + (uint32_t)(s * 16 + dx) * 16u) >> 4) & 0x3fffu) | a_lo_c;
                            tc_mma(tmem_base + (uint32_t)(dx * p.npad), a_lo, a_hi, b_lo, b_hi, idesc, (first && s == 0) ? 0u : 1u);
                        }
                    }
                    tc_commit(&s_dyempty[slot]);
                    if (rb == p.B - 1) tc_commit(&s_xempty);
                }
                first = false;
                __syncwarp();
                if (++slot == p.nslots) { slot = 0; dph ^= 1u; }
            }
        }
        if (leader) tc_commit(&s_done);
        __syncwarp();
    } else {
        // one lane = one accumulator row; consecutive lanes add into consecutive floats of acc (a warp-wide reduction
        // touches four 32-byte sectors; adding straight into the weight layout touched 32 and made this epilogue as long
        // as the main loop of the low-resolution layers)
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const bool valid = ((m >> 3) / p.P) < nr;
        mbar_wait(&s_done, 0);
        tc_fence_after();
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* acc = p.acc + (size_t)blockIdx.y * KX * p.npad * 128 + m;
        for (int dx = 0; dx < KX; ++dx) {
            for (int c16 = 0; c16 < p.npad / 16; ++c16) {
                uint32_t v[16];
                tmem_ld16(t_lane + (uint32_t)(dx * p.npad + c16 * 16), v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float f = __uint_as_float(v[i]);
                        if (f != 0.f) atomicAdd(acc + (size_t)(dx * p.npad + c16 * 16 + i) * 128, f);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct Ten {
    bf16* p = nullptr;
    int cp = 0, h = 0, w = 0;
    int planes() const { return cp / 8; }
    size_t units() const { return (size_t)(cp / 8) * h * w; }
    size_t bytes() const { return units() * 16; }
    uint4* u() const { return reinterpret_cast<uint4*>(p); }
};
Ten view(const Ten& t, int plane0, int nplanes) {
    Ten v = t;
    v.p = t.p + (size_t)plane0 * t.h * t.w * 8;
    v.cp = nplanes * 8;
    return v;
}

// plane-major tensor as a u64 tensor (2w, h, planes): box = (2 * px) x 1 row x planes -> shared memory [plane][px][16 B]
int make_row_map(pcs_ctx* ctx, CUtensorMap* tm, const Ten& t, int box_px, int box_planes) {
    EncodeTiledFn enc = tc_get_encode();
    if (!enc) return set_err(ctx, PCS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[3] = {(cuuint64_t)t.w * 2, (cuuint64_t)t.h, (cuuint64_t)t.planes()};
    const cuuint64_t strides[2] = {(cuuint64_t)t.w * 16, (cuuint64_t)t.h * t.w * 16};
    const cuuint32_t box[3] = {(cuuint32_t)box_px * 2, 1, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, t.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(ctx, PCS_ERR_CUDA, "cuTensorMapEncodeTiled (wgrad) failed with %d (cp=%d w=%d h=%d)", (int)r, t.cp, t.w, t.h);
    return PCS_OK;
}

struct LDef { const char* name; int kind, k, ci, co, relu; };       // kind: 0 conv, 1 transposed s1 (flipped form), 2 transposed 2x2 s2, 3 logits
const LDef kSkip[13] = {{"conv1", 0, 5, 1, 20, 1},   {"conv2", 0, 5, 20, 30, 0},  {"conv3", 0, 5, 30, 40, 1},   {"conv4", 0, 5, 40, 40, 0},
                        {"conv5", 0, 5, 40, 60, 1},  {"conv6", 0, 5, 60, 60, 0},  {"conv7", 0, 5, 60, 80, 1},   {"deconv1", 1, 5, 80, 80, 1},
                        {"deconv2", 2, 2, 80, 60, 1}, {"deconv3", 1, 5, 120, 40, 1}, {"deconv4", 2, 2, 100, 30, 1}, {"deconv5", 2, 2, 70, 20, 0},
                        {"logits", 3, 1, 50, 0, 0}};
const LDef kPlain[13] = {{"conv1", 0, 5, 1, 20, 1},   {"conv2", 0, 5, 20, 30, 0},  {"conv3", 0, 5, 30, 40, 1},  {"conv4", 0, 5, 40, 40, 0},
                         {"conv5", 0, 5, 40, 60, 1},  {"conv6", 0, 5, 60, 60, 0},  {"conv7", 0, 5, 60, 80, 1},  {"deconv1", 1, 5, 80, 80, 1},
                         {"deconv2", 2, 2, 80, 60, 1}, {"deconv3", 1, 5, 60, 40, 1}, {"deconv4", 2, 2, 40, 30, 1}, {"deconv5", 2, 2, 30, 20, 0},
                         {"logits", 3, 1, 20, 0, 0}};
enum { L_CONV1, L_CONV2, L_CONV3, L_CONV4, L_CONV5, L_CONV6, L_CONV7, L_DECONV1, L_DECONV2, L_DECONV3, L_DECONV4, L_DECONV5, L_LOGITS, L_COUNT };

struct Seg { int real, padded; };                                   // one source of a concatenation

struct Image { uint16_t* d = nullptr; int npad = 0, ntiles = 0, nchunks = 0, taps = 0; };

struct TcLayer {
    LDef def;
    long long w_off = 0, b_off = 0;
    std::vector<Seg> in;         // input sources (padded channel layout of the input tensor)
    int in_cp = 0;               // padded input channels
    int co_p8 = 0;               // output channels padded to whole planes
    Image fwd, dgrad;
    int co_t = 0;                // stride-2 layers: padded channels per tap of the forward GEMM
    const int* wg_m_off = nullptr; const int* wg_n_off = nullptr;
    int wg_npad = 0;
    float* wg_acc = nullptr;     // partial-sum buffer of the weight-gradient kernel (inside TrainTc::d_wg_acc)
};

// groups of vertical taps of a weight-gradient launch: as many kernel rows as fit 16 eight-channel groups
int wgrad_tiles(int P, int kx, int* d0, int* nr) {
    const int per = kx == 1 ? 1 : std::min(5, 16 / P);
    int n = 0;
    for (int d = 0; d < kx; d += per) { d0[n] = d; nr[n] = std::min(per, kx - d); ++n; }
    return n;
}
size_t wgrad_acc_floats(int P, int kx, int npad) {
    int d0[5], nr[5];
    return (size_t)wgrad_tiles(P, kx, d0, nr) * kx * npad * 128;
}

int npad_for(int cols) { return cols <= 32 ? 32 : cols <= 48 ? 48 : cols <= 64 ? 64 : cols <= 80 ? 80 : 64; }

}  // namespace

struct TrainTc {
    int arch = 0, skip = 0, ncls = 0, h = 0, w = 0, H = 0, W = 0;
    TcLayer L[L_COUNT];
    std::vector<void*> allocs;
    std::vector<WimgDesc> h_descs;
    WimgDesc* d_descs = nullptr;
    unsigned long long max_img_total = 0;
    float* d_zero_bias = nullptr;
    float* d_wg_acc = nullptr;               // partial sums of all weight-gradient launches of a step
    size_t wg_acc_floats = 0;
    std::vector<ScatterDesc> h_scatter;      // decoder layers first (phase 1), then the encoder layers (phase 2)
    int n_scatter_decoder = 0;
    ScatterDesc* d_scatter = nullptr;
    const float* scatter_grads = nullptr;    // gradient buffer the device scatter table was built for
    unsigned max_scatter_total = 0;
    const int* d_logit_map = nullptr;
    float4* d_dl = nullptr;
    Ten x, conv1, d5cat, deconv5, conv2, pool2, d4cat, deconv4, conv3, conv4, pool4, d3cat, deconv3, conv5, d2cat, deconv2, conv6, pool6, conv7, deconv1;
    Ten g_d5s, g_conv2, g_conv1, g_d4cat, g_d4s, g_conv4, g_pool2, t_conv3, g_d3cat, g_d2cat, g_d2s, t_conv5, g_pool4, g_pool6, g_conv7, g_deconv1;
    const float* params = nullptr;
    const float* desc_params = nullptr;      // parameter buffer the device descriptor table was built for
    float* grads = nullptr;
};

namespace {

int tc_alloc(pcs_ctx* ctx, TrainTc* t, void** p, size_t bytes) {
    PCS_CUDA(ctx, cudaMalloc(p, bytes));
    t->allocs.push_back(*p);
    PCS_CUDA(ctx, cudaMemsetAsync(*p, 0, bytes, ctx->stream));
    return PCS_OK;
}
int tc_ten(pcs_ctx* ctx, TrainTc* t, Ten* out, int cp, int h, int w) {
    out->cp = cp; out->h = h; out->w = w;
    void* p = nullptr;
    PCS_TRY(tc_alloc(ctx, t, &p, out->bytes()));
    out->p = reinterpret_cast<bf16*>(p);
    return PCS_OK;
}
int tc_upload_ints(pcs_ctx* ctx, TrainTc* t, const std::vector<int>& v, const int** out) {
    void* p = nullptr;
    PCS_TRY(tc_alloc(ctx, t, &p, std::max<size_t>(v.size(), 1) * sizeof(int)));
    PCS_CUDA(ctx, cudaMemcpyAsync(p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));            // v is a temporary of the caller
    *out = reinterpret_cast<const int*>(p);
    return PCS_OK;
}

// padded channel index -> real channel index over concatenated sources
std::vector<int> seg_map(const std::vector<Seg>& segs) {
    std::vector<int> m;
    int base = 0;
    for (const Seg& s : segs) {
        for (int c = 0; c < s.padded; ++c) m.push_back(c < s.real ? base + c : -1);
        base += s.real;
    }
    return m;
}

int make_image(pcs_ctx* ctx, TrainTc* t, Image* img, long long w_off, const std::vector<int>& in_off, const std::vector<int>& out_off, int taps,
               int npad, int s_tap, int flip) {
    img->npad = npad; img->taps = taps;
    img->nchunks = ((int)in_off.size() + 15) / 16;
    img->ntiles = ((int)out_off.size() + npad - 1) / npad;
    std::vector<int> io = in_off, oo = out_off;
    io.resize((size_t)img->nchunks * 16, -1);
    oo.resize((size_t)img->ntiles * npad, -1);
    WimgDesc d{};
    PCS_TRY(tc_upload_ints(ctx, t, io, &d.in_off));
    PCS_TRY(tc_upload_ints(ctx, t, oo, &d.out_off));
    d.total = (unsigned long long)img->ntiles * img->nchunks * taps * 2 * npad * 8;
    void* p = nullptr;
    PCS_TRY(tc_alloc(ctx, t, &p, d.total * 2));
    img->d = reinterpret_cast<uint16_t*>(p);
    d.dst = img->d;
    d.w = reinterpret_cast<const float*>(w_off);                  // offset; rebased on the parameter buffer at each step
    d.taps = taps; d.nchunks = img->nchunks; d.npad = npad; d.ntiles = img->ntiles; d.s_tap = s_tap; d.flip = flip;
    t->h_descs.push_back(d);
    t->max_img_total = std::max(t->max_img_total, d.total);
    return PCS_OK;
}

int setup_layer(pcs_ctx* ctx, TrainTc* t, int li, const std::vector<Seg>& in) {
    TcLayer& L = t->L[li];
    L.in = in;
    const std::vector<int> imap = seg_map(in);
    L.in_cp = (int)imap.size();
    const int Ci = L.def.ci, Co = L.def.co, KK = L.def.k * L.def.k;
    L.co_p8 = pad8(Co);
    if (L.def.kind == 0 || L.def.kind == 1) {
        // forward: w[Co][Ci][25]
        std::vector<int> io(imap.size()), oo(Co);
        for (size_t c = 0; c < imap.size(); ++c) io[c] = imap[c] >= 0 ? imap[c] * KK : -1;
        for (int o = 0; o < Co; ++o) oo[o] = o * Ci * KK;
        PCS_TRY(make_image(ctx, t, &L.fwd, L.w_off, io, oo, KK, npad_for(Co), 1, 0));
        if (li != L_CONV1) {
            // input gradient: dX = corr(dY, w'), w'[ci][co][t] = w[co][ci][24 - t]
            std::vector<int> gi(L.co_p8), go(imap.size());
            for (int o = 0; o < L.co_p8; ++o) gi[o] = o < Co ? o * Ci * KK : -1;
            for (size_t c = 0; c < imap.size(); ++c) go[c] = imap[c] >= 0 ? imap[c] * KK : -1;
            PCS_TRY(make_image(ctx, t, &L.dgrad, L.w_off, gi, go, KK, npad_for((int)imap.size()), 1, 1));
        }
        std::vector<int> mo(imap.size()), no(pad16(Co), -1);
        for (size_t c = 0; c < imap.size(); ++c) mo[c] = imap[c] >= 0 ? imap[c] * KK : -1;
        for (int o = 0; o < Co; ++o) no[o] = o * Ci * KK;
        L.wg_npad = (int)no.size();
        if (L.wg_npad < 32) { L.wg_npad = 32; no.resize(32, -1); }
        PCS_TRY(tc_upload_ints(ctx, t, mo, &L.wg_m_off));
        PCS_TRY(tc_upload_ints(ctx, t, no, &L.wg_n_off));
    } else if (L.def.kind == 2) {
        // k2[tap][Co][Ci]; forward GEMM column J = tap * co_t + o
        L.co_t = Co <= 32 ? 32 : 64;
        std::vector<int> io(imap.size()), oo(4 * L.co_t, -1);
        for (size_t c = 0; c < imap.size(); ++c) io[c] = imap[c];
        for (int tp = 0; tp < 4; ++tp)
            for (int o = 0; o < Co; ++o) oo[tp * L.co_t + o] = (tp * Co + o) * Ci;
        PCS_TRY(make_image(ctx, t, &L.fwd, L.w_off, io, oo, 1, 128, 0, 0));
        // input gradient: 1x1 GEMM over the space-to-depth dY, channel c' = tap * co_p8 + o
        std::vector<int> gi(4 * L.co_p8, -1), go(imap.size());
        for (int tp = 0; tp < 4; ++tp)
            for (int o = 0; o < Co; ++o) gi[tp * L.co_p8 + o] = (tp * Co + o) * Ci;
        for (size_t c = 0; c < imap.size(); ++c) go[c] = imap[c];
        PCS_TRY(make_image(ctx, t, &L.dgrad, L.w_off, gi, go, 1, (int)imap.size() <= 80 ? 80 : 128, 0, 0));
        std::vector<int> mo(imap.size());
        for (size_t c = 0; c < imap.size(); ++c) mo[c] = imap[c];
        L.wg_npad = 4 * L.co_p8;
        PCS_TRY(tc_upload_ints(ctx, t, mo, &L.wg_m_off));
        PCS_TRY(tc_upload_ints(ctx, t, gi, &L.wg_n_off));
    }
    return PCS_OK;
}

unsigned blocks_for(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16); }

// ---- launches ---------------------------------------------------------------------------------------------------------
int conv_launch(pcs_ctx* ctx, const Image& img, const Ten& src, const float* bias, int cout, int relu, const Ten& out, const Ten* pool, int k) {
    UmmaConvArgs a;
    a.src[0].p = src.p; a.src[0].c = src.cp; a.src[0].cp = src.cp;
    a.nsrc = 1; a.n = 1; a.h = src.h; a.w = src.w; a.k = k; a.pad = (k - 1) / 2;
    a.wmma = img.d; a.b32 = bias; a.cout = cout; a.npad = img.npad; a.nchunks = img.nchunks; a.relu = relu; a.mode = 0;
    a.out = out.p; a.out_cp = out.cp;
    if (pool) { a.pool_out = pool->p; a.pool_cp = pool->cp; }
    if ((src.cp + 15) / 16 != img.nchunks) return set_err(ctx, PCS_ERR_STATE, "train_tc: operand image has %d chunks, source %d channels", img.nchunks, src.cp);
    return launch_conv_umma(ctx, a);
}

int deconv_fwd_launch(pcs_ctx* ctx, const TcLayer& L, const Ten& src, const float* bias, const Ten& out) {
    UmmaConvArgs a;
    a.src[0].p = src.p; a.src[0].c = src.cp; a.src[0].cp = src.cp;
    a.nsrc = 1; a.n = 1; a.h = src.h; a.w = src.w; a.k = 1; a.pad = 0;
    a.wmma = L.fwd.d; a.b32 = bias; a.cout = L.def.co; a.npad = 128; a.nchunks = L.fwd.nchunks; a.relu = L.def.relu; a.mode = 1; a.co_t = L.co_t;
    a.out = out.p; a.out_cp = out.cp;
    return launch_conv_umma(ctx, a);
}

int wgrad_launch_raw(pcs_ctx* ctx, const Ten& X, const Ten& dY, int kx, int npad, float* acc) {
    const int DT = kx == 5 ? 112 : 128, XT = kx == 5 ? 120 : 128;
    WgradParams p{};
    p.H = dY.h; p.W = dY.w; p.P = X.planes(); p.npad = npad; p.nq = p.npad / 8; p.pad = (kx - 1) / 2;
    if (X.h != dY.h || X.w != dY.w) return set_err(ctx, PCS_ERR_STATE, "train_tc: wgrad operands differ in size");
    if (p.P < 1 || p.P > 16 || kx * p.npad > 512 || p.npad > 256 || (p.npad & 15)) return set_err(ctx, PCS_ERR_STATE, "train_tc: wgrad shape P=%d N=%d", p.P, p.npad);
    p.ntiles_m = wgrad_tiles(p.P, kx, p.d0, p.nr);
    int nr_max = 0;
    for (int i = 0; i < p.ntiles_m; ++i) nr_max = std::max(nr_max, p.nr[i]);
    p.x_row_bytes = (uint32_t)p.P * XT * 16;
    p.dy_slot_bytes = (uint32_t)p.nq * DT * 16;
    const uint32_t budget = 222 * 1024, slack = 16u * XT * 16;
    int best_b = 0, best_s = 0;
    for (int s = kWgMaxSlots; s >= 2; --s) {
        const long long left = (long long)budget - 1024 - slack - (long long)s * p.dy_slot_bytes;
        const int b = (int)std::min<long long>(32, left / (long long)p.x_row_bytes - (nr_max - 1));
        if (b > best_b && (best_b < 4 || s >= 3)) { best_b = b; best_s = s; }
        if (best_b >= 8) break;
    }
    if (best_b < 1) return set_err(ctx, PCS_ERR_STATE, "train_tc: wgrad tile does not fit shared memory (P=%d N=%d)", p.P, p.npad);
    p.B = std::min(best_b, std::max(1, p.H)); p.nslots = best_s;
    p.strips = (p.W + DT - 1) / DT;
    {   // among the band heights that fit, the one that spreads strips x bands most evenly over the CTAs (ties: the taller band)
        const int ctas = std::max(1, (ctx->sm_count + p.ntiles_m - 1) / p.ntiles_m);
        double best_eff = -1.0;
        int best = p.B;
        for (int b = p.B; b >= std::max(1, p.B / 2); --b) {
            const int items = p.strips * ((p.H + b - 1) / b);
            const int grid = std::min(items, ctas);
            const double rows_done = (double)((items + grid - 1) / grid) * b * grid;          // row-strips the slowest CTA pattern pays for
            const double eff = (double)p.strips * p.H / rows_done;
            if (eff > best_eff + 1e-9) { best_eff = eff; best = b; }
        }
        p.B = best;
    }
    p.xr = p.B + nr_max - 1;
    p.x_alloc = ((uint32_t)p.xr * p.x_row_bytes + 1023) / 1024 * 1024;
    p.bands = (p.H + p.B - 1) / p.B;
    p.acc = acc;
    { const char* e = getenv("PCSEG_WGRAD_SWAP"); p.swap_strides = e && e[0] == '1'; }
    CUtensorMap tmx, tmdy;
    PCS_TRY(make_row_map(ctx, &tmx, X, XT, p.P));
    PCS_TRY(make_row_map(ctx, &tmdy, dY, DT, p.nq));
    const size_t smem = std::max<size_t>((size_t)p.x_alloc + (size_t)p.nslots * p.dy_slot_bytes + slack + 1024, kSoloSmem);
    const int items = p.strips * p.bands;
    const dim3 grid(std::min(items, std::max(1, (ctx->sm_count + p.ntiles_m - 1) / p.ntiles_m)), p.ntiles_m);
    static bool attr_set[64][2] = {};
    if (ctx->device >= 64 || !attr_set[ctx->device][kx == 5]) {
        if (kx == 5) PCS_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        else PCS_CUDA(ctx, cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        if (ctx->device < 64) attr_set[ctx->device][kx == 5] = true;
    }
    if (kx == 5) wgrad_tc_kernel<5><<<grid, kWgThreads, smem, ctx->stream>>>(p, tmx, tmdy);
    else wgrad_tc_kernel<1><<<grid, kWgThreads, smem, ctx->stream>>>(p, tmx, tmdy);
    PCS_LAUNCH_CHECK(ctx, "wgrad_tc_kernel");
    return PCS_OK;
}

int wgrad_launch(pcs_ctx* ctx, TrainTc* t, const TcLayer& L, const Ten& X, const Ten& dY) {
    (void)t;
    return wgrad_launch_raw(ctx, X, dY, L.def.kind == 2 ? 1 : 5, L.wg_npad, L.wg_acc);
}

ScatterDesc scatter_desc(const float* acc, float* dw, const int* m_off, const int* n_off, int P, int npad, int kx, int s_tap) {
    ScatterDesc d{};
    d.acc = acc; d.dw = dw; d.m_off = m_off; d.n_off = n_off; d.P = P; d.npad = npad; d.kx = kx; d.s_tap = s_tap;
    d.ntiles_m = wgrad_tiles(P, kx, d.d0, d.nr);
    d.total = (unsigned)((size_t)d.ntiles_m * kx * npad * 128);
    return d;
}

// launches tc_wgrad_scatter_kernel over descs[first, first + count) of the device table
int scatter_launch(pcs_ctx* ctx, const ScatterDesc* d_descs, int first, int count, unsigned max_total) {
    tc_wgrad_scatter_kernel<<<dim3(std::min<unsigned>((max_total + 255) / 256, 96), (unsigned)count), 256, 0, ctx->stream>>>(d_descs + first);
    PCS_LAUNCH_CHECK(ctx, "tc_wgrad_scatter_kernel");
    return PCS_OK;
}

int bias_grad_launch(pcs_ctx* ctx, TrainTc* t, const TcLayer& L, const Ten& g, int qp) {
    const size_t hw = (size_t)g.h * g.w;
    const dim3 grid(g.planes(), (unsigned)std::max<size_t>(1, std::min<size_t>((hw + 2047) / 2048, 64)));
    tc_bias_grad_kernel<<<grid, 256, 0, ctx->stream>>>(g.u(), hw, qp, L.def.co, t->grads + L.b_off);
    PCS_LAUNCH_CHECK(ctx, "tc_bias_grad_kernel");
    return PCS_OK;
}

int combine_launch(pcs_ctx* ctx, const Ten& dst, const Ten& a, const Ten* b, const Ten* y, int s2d) {
    tc_combine_kernel<<<blocks_for(a.units()), 256, 0, ctx->stream>>>(dst.u(), a.u(), b ? b->u() : nullptr, y ? y->u() : nullptr, a.planes(), a.h, a.w, s2d);
    PCS_LAUNCH_CHECK(ctx, "tc_combine_kernel");
    return PCS_OK;
}

int pool_bwd_launch(pcs_ctx* ctx, const Ten& y, const Ten& g_pool, const Ten* skip, const Ten& dst) {
    tc_pool_bwd_kernel<<<blocks_for(g_pool.units()), 256, 0, ctx->stream>>>(y.u(), g_pool.u(), skip ? skip->u() : nullptr, dst.u(), y.planes(), y.h, y.w);
    PCS_LAUNCH_CHECK(ctx, "tc_pool_bwd_kernel");
    return PCS_OK;
}

}  // namespace

// The weight-gradient kernel on its own (a primitive like train.cu's): dw[co][ci][k][k] += sum_px x[px + tap][ci] * dy[px][co]
// for plane-major bf16 tensors x ([x_planes][H][W][8]) and dy ([dy_planes][H][W][8]), k = 5 ('same' correlation) or 1.
int train_tc_wgrad(pcs_ctx* ctx, const void* d_x, int x_planes, const void* d_dy, int dy_planes, int H, int W, int k, int ci, int co, float* d_dw) {
    if ((k != 5 && k != 1) || x_planes < 1 || x_planes > 16 || ci < 1 || ci > x_planes * 8 || co < 1 || co > dy_planes * 8 || H < 1 || W < 1)
        return set_err(ctx, PCS_ERR_ARG, "train_tc_wgrad: bad shape (k=%d, %d x planes for %d channels, %d dy planes for %d channels)", k, x_planes, ci, dy_planes, co);
    int npad = std::max(32, pad16(co));
    if (k * npad > 512 || npad > 256) return set_err(ctx, PCS_ERR_ARG, "train_tc_wgrad: %d output channels exceed the accumulator columns", co);
    const int KK = k * k;
    std::vector<int> mo(x_planes * 8, -1), no(npad, -1);
    for (int c = 0; c < ci; ++c) mo[c] = c * KK;
    for (int o = 0; o < co; ++o) no[o] = o * ci * KK;
    const size_t acc_floats = wgrad_acc_floats(x_planes, k, npad);
    const size_t maps_bytes = ((mo.size() + no.size()) * sizeof(int) + 255) / 256 * 256;
    PCS_TRY(scratch_reserve(ctx, maps_bytes + 256 + acc_floats * sizeof(float)));
    char* base = reinterpret_cast<char*>(ctx->scratch);
    int* d_maps = reinterpret_cast<int*>(base);
    ScatterDesc* d_desc = reinterpret_cast<ScatterDesc*>(base + maps_bytes);
    float* d_acc = reinterpret_cast<float*>(base + maps_bytes + 256);
    static_assert(sizeof(ScatterDesc) <= 256, "scatter descriptor slot");
    const ScatterDesc desc = scatter_desc(d_acc, d_dw, d_maps, d_maps + mo.size(), x_planes, npad, k, 1);
    PCS_CUDA(ctx, cudaMemcpyAsync(d_maps, mo.data(), mo.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PCS_CUDA(ctx, cudaMemcpyAsync(d_maps + mo.size(), no.data(), no.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PCS_CUDA(ctx, cudaMemcpyAsync(d_desc, &desc, sizeof(desc), cudaMemcpyHostToDevice, ctx->stream));
    PCS_CUDA(ctx, cudaMemsetAsync(d_acc, 0, acc_floats * sizeof(float), ctx->stream));
    PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));            // the maps and the descriptor are temporaries
    Ten X, dY;
    X.p = reinterpret_cast<bf16*>(const_cast<void*>(d_x)); X.cp = x_planes * 8; X.h = H; X.w = W;
    dY.p = reinterpret_cast<bf16*>(const_cast<void*>(d_dy)); dY.cp = dy_planes * 8; dY.h = H; dY.w = W;
    PCS_TRY(wgrad_launch_raw(ctx, X, dY, k, npad, d_acc));
    return scatter_launch(ctx, d_desc, 0, 1, desc.total);
}

int train_tc_destroy(pcs_ctx* ctx, TrainTc* t) {
    if (!t) return PCS_OK;
    cudaStreamSynchronize(ctx->stream);
    for (void* p : t->allocs) cudaFree(p);
    delete t;
    return PCS_OK;
}

int train_tc_create(pcs_ctx* ctx, int arch, int n_classes, int h, int w, const long long* offsets, int n_offsets, TrainTc** out) {
    if (arch != PCS_ARCH_FCN_SKIP && arch != PCS_ARCH_FCN) return set_err(ctx, PCS_ERR_ARG, "train_tc: the tensor-core training step covers fcn_skip and fcn");
    if (n_classes < 1 || n_classes > TC_NC) return set_err(ctx, PCS_ERR_ARG, "train_tc: 1..%d classes (got %d); use the fp32 engine", TC_NC, n_classes);
    if (h < 1 || w < 1 || n_offsets != 2 * L_COUNT + 1) return set_err(ctx, PCS_ERR_ARG, "train_tc: bad page size or offset table (%d entries)", n_offsets);
    TrainTc* t = new TrainTc();
    t->arch = arch; t->skip = arch == PCS_ARCH_FCN_SKIP; t->ncls = n_classes; t->h = h; t->w = w;
    t->H = (h + 31) / 32 * 32; t->W = (w + 31) / 32 * 32;
    const LDef* defs = t->skip ? kSkip : kPlain;
    for (int i = 0; i < L_COUNT; ++i) {
        t->L[i].def = defs[i];
        if (i == L_LOGITS) t->L[i].def.co = n_classes;
        t->L[i].w_off = offsets[2 * i]; t->L[i].b_off = offsets[2 * i + 1];
        const long long wsize = (long long)t->L[i].def.co * t->L[i].def.ci * t->L[i].def.k * t->L[i].def.k;
        if (offsets[2 * i + 1] - offsets[2 * i] != wsize || offsets[2 * i + 2] - offsets[2 * i + 1] != t->L[i].def.co) {
            delete t;
            return set_err(ctx, PCS_ERR_ARG, "train_tc: offset table does not match the %s graph at layer %d", t->skip ? "fcn_skip" : "fcn", i);
        }
    }
    const int S = t->skip, H = t->H, W = t->W, H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
    int rc = PCS_OK;
    auto T = [&](Ten* ten, int cp, int hh, int ww) { if (rc == PCS_OK) rc = tc_ten(ctx, t, ten, cp, hh, ww); };
    T(&t->x, 8, H, W); T(&t->conv1, 24, H, W);
    T(&t->d5cat, S ? 56 : 24, H, W);
    t->deconv5 = view(t->d5cat, 0, 3);
    if (S) t->conv2 = view(t->d5cat, 3, 4); else T(&t->conv2, 32, H, W);
    T(&t->pool2, 32, H2, W2);
    T(&t->d4cat, S ? 72 : 32, H2, W2);
    t->deconv4 = view(t->d4cat, 0, 4);
    if (S) t->conv3 = view(t->d4cat, 4, 5); else T(&t->conv3, 40, H2, W2);
    T(&t->conv4, 40, H2, W2); T(&t->pool4, 40, H4, W4);
    T(&t->d3cat, S ? 104 : 40, H4, W4);
    t->deconv3 = view(t->d3cat, 0, 5);
    if (S) t->conv5 = view(t->d3cat, 5, 8); else T(&t->conv5, 64, H4, W4);
    T(&t->d2cat, S ? 128 : 64, H4, W4);
    t->deconv2 = view(t->d2cat, 0, 8);
    if (S) t->conv6 = view(t->d2cat, 8, 8); else T(&t->conv6, 64, H4, W4);
    T(&t->pool6, 64, H8, W8); T(&t->conv7, 80, H8, W8); T(&t->deconv1, 80, H8, W8);
    // gradients
    T(&t->g_d5s, 96, H2, W2); T(&t->g_conv2, 32, H, W); T(&t->g_conv1, 24, H, W);
    T(&t->g_d4cat, S ? 72 : 32, H2, W2); T(&t->g_d4s, 128, H4, W4); T(&t->g_conv4, 40, H2, W2); T(&t->g_pool2, 32, H2, W2); T(&t->t_conv3, 40, H2, W2);
    T(&t->g_d3cat, S ? 104 : 40, H4, W4); T(&t->g_d2cat, S ? 128 : 64, H4, W4); T(&t->g_d2s, 256, H8, W8); T(&t->t_conv5, 64, H4, W4);
    T(&t->g_pool4, 40, H4, W4); T(&t->g_pool6, 64, H8, W8); T(&t->g_conv7, 80, H8, W8); T(&t->g_deconv1, 80, H8, W8);
    void* p = nullptr;
    if (rc == PCS_OK) rc = tc_alloc(ctx, t, &p, (size_t)H * W * sizeof(float4));
    t->d_dl = reinterpret_cast<float4*>(p);
    if (rc == PCS_OK) rc = tc_alloc(ctx, t, &p, 256 * sizeof(float));
    t->d_zero_bias = reinterpret_cast<float*>(p);
    // layers: input sources in the padded channel layout of their input tensor
    const std::vector<Seg> in[L_COUNT] = {
        {{1, 8}}, {{20, 24}}, {{30, 32}}, {{40, 40}}, {{40, 40}}, {{60, 64}}, {{60, 64}}, {{80, 80}}, {{80, 80}},
        S ? std::vector<Seg>{{60, 64}, {60, 64}} : std::vector<Seg>{{60, 64}},
        S ? std::vector<Seg>{{40, 40}, {60, 64}} : std::vector<Seg>{{40, 40}},
        S ? std::vector<Seg>{{30, 32}, {40, 40}} : std::vector<Seg>{{30, 32}},
        S ? std::vector<Seg>{{20, 24}, {30, 32}} : std::vector<Seg>{{20, 24}}};
    for (int i = 0; i < L_COUNT && rc == PCS_OK; ++i) {
        if (i == L_LOGITS) { t->L[i].in = in[i]; rc = tc_upload_ints(ctx, t, seg_map(in[i]), &t->d_logit_map); }
        else rc = setup_layer(ctx, t, i, in[i]);
    }
    if (rc == PCS_OK) rc = tc_alloc(ctx, t, &p, t->h_descs.size() * sizeof(WimgDesc));
    t->d_descs = reinterpret_cast<WimgDesc*>(p);
    if (rc == PCS_OK) {
        // partial-sum buffers and the scatter table of the weight-gradient launches, in backward order
        const int order[12] = {L_DECONV5, L_DECONV4, L_DECONV3, L_DECONV2, L_DECONV1, L_CONV7, L_CONV6, L_CONV5, L_CONV4, L_CONV3, L_CONV2, L_CONV1};
        size_t total = 0;
        for (int li : order) {
            TcLayer& L = t->L[li];
            total += wgrad_acc_floats(L.in_cp / 8, L.def.kind == 2 ? 1 : 5, L.wg_npad);
        }
        t->wg_acc_floats = total;
        rc = tc_alloc(ctx, t, &p, total * sizeof(float));
        t->d_wg_acc = reinterpret_cast<float*>(p);
        size_t at = 0;
        for (int k = 0; k < 12 && rc == PCS_OK; ++k) {
            TcLayer& L = t->L[order[k]];
            const int kx = L.def.kind == 2 ? 1 : 5;
            L.wg_acc = t->d_wg_acc + at;
            at += wgrad_acc_floats(L.in_cp / 8, kx, L.wg_npad);
            // dw holds the OFFSET of the layer's kernel until the step rebases it on the gradient buffer
            ScatterDesc d = scatter_desc(L.wg_acc, reinterpret_cast<float*>(L.w_off), L.wg_m_off, L.wg_n_off, L.in_cp / 8, L.wg_npad, kx, L.def.kind == 2 ? 0 : 1);
            t->h_scatter.push_back(d);
            t->max_scatter_total = std::max(t->max_scatter_total, d.total);
            if (order[k] == L_DECONV1) t->n_scatter_decoder = k + 1;
        }
        if (rc == PCS_OK) rc = tc_alloc(ctx, t, &p, t->h_scatter.size() * sizeof(ScatterDesc));
        t->d_scatter = reinterpret_cast<ScatterDesc*>(p);
    }
    if (rc == PCS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_err(ctx, PCS_ERR_CUDA, "train_tc: setup failed");
    if (rc != PCS_OK) { train_tc_destroy(ctx, t); return rc; }
    *out = t;
    return PCS_OK;
}

// phase bits: 1 = forward + loss + decoder backward (logits .. deconv1), 2 = encoder backward (conv7 .. conv1)
int train_tc_step(pcs_ctx* ctx, TrainTc* t, int phases, const uint8_t* d_img, const uint8_t* d_labels, const float* d_params, float* d_grads,
                  double* d_loss) {
    const int saved_precision = ctx->precision;
    ctx->precision = PCS_PREC_BF16;
    struct Restore { pcs_ctx* c; int v; ~Restore() { c->precision = v; } } restore{ctx, saved_precision};
    t->params = d_params; t->grads = d_grads;
    const int S = t->skip;
    TcLayer* L = t->L;
    auto bias = [&](int li) { return d_params + L[li].b_off; };
    if (phases & 1) {
        // operand images from the current master weights (one launch)
        if (t->desc_params != d_params) {                         // the descriptor table points into the parameter buffer
            std::vector<WimgDesc> descs = t->h_descs;
            for (WimgDesc& d : descs) d.w = d_params + reinterpret_cast<intptr_t>(d.w);
            PCS_CUDA(ctx, cudaMemcpyAsync(t->d_descs, descs.data(), descs.size() * sizeof(WimgDesc), cudaMemcpyHostToDevice, ctx->stream));
            PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));    // descs is a temporary
            t->desc_params = d_params;
        }
        tc_wimg_kernel<<<dim3((unsigned)std::min<unsigned long long>((t->max_img_total + 255) / 256, 64), (unsigned)t->h_descs.size()), 256, 0, ctx->stream>>>(t->d_descs);
        PCS_LAUNCH_CHECK(ctx, "tc_wimg_kernel");
        if (t->scatter_grads != d_grads) {
            std::vector<ScatterDesc> sd = t->h_scatter;
            for (ScatterDesc& d : sd) d.dw = d_grads + reinterpret_cast<intptr_t>(d.dw);
            PCS_CUDA(ctx, cudaMemcpyAsync(t->d_scatter, sd.data(), sd.size() * sizeof(ScatterDesc), cudaMemcpyHostToDevice, ctx->stream));
            PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));    // sd is a temporary
            t->scatter_grads = d_grads;
        }
        PCS_CUDA(ctx, cudaMemsetAsync(t->d_wg_acc, 0, t->wg_acc_floats * sizeof(float), ctx->stream));
        size_t nparams = (size_t)L[L_LOGITS].b_off + t->ncls;
        PCS_CUDA(ctx, cudaMemsetAsync(d_grads, 0, nparams * sizeof(float), ctx->stream));
        PCS_CUDA(ctx, cudaMemsetAsync(d_loss, 0, sizeof(double), ctx->stream));

        // ---- forward (model.py:45-92 / :206-234)
        tc_input_kernel<<<blocks_for((size_t)t->H * t->W), 256, 0, ctx->stream>>>(d_img, t->h, t->w, t->x.u(), t->H, t->W);
        PCS_LAUNCH_CHECK(ctx, "tc_input_kernel");
        PCS_TRY(conv_launch(ctx, L[L_CONV1].fwd, t->x, bias(L_CONV1), 20, 1, t->conv1, nullptr, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV2].fwd, t->conv1, bias(L_CONV2), 30, 0, t->conv2, &t->pool2, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV3].fwd, t->pool2, bias(L_CONV3), 40, 1, t->conv3, nullptr, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV4].fwd, t->conv3, bias(L_CONV4), 40, 0, t->conv4, &t->pool4, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV5].fwd, t->pool4, bias(L_CONV5), 60, 1, t->conv5, nullptr, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV6].fwd, t->conv5, bias(L_CONV6), 60, 0, t->conv6, &t->pool6, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV7].fwd, t->pool6, bias(L_CONV7), 80, 1, t->conv7, nullptr, 5));
        PCS_TRY(conv_launch(ctx, L[L_DECONV1].fwd, t->conv7, bias(L_DECONV1), 80, 1, t->deconv1, nullptr, 5));
        PCS_TRY(deconv_fwd_launch(ctx, L[L_DECONV2], t->deconv1, bias(L_DECONV2), t->deconv2));
        PCS_TRY(conv_launch(ctx, L[L_DECONV3].fwd, S ? t->d2cat : t->deconv2, bias(L_DECONV3), 40, 1, t->deconv3, nullptr, 5));
        PCS_TRY(deconv_fwd_launch(ctx, L[L_DECONV4], S ? t->d3cat : t->deconv3, bias(L_DECONV4), t->deconv4));
        PCS_TRY(deconv_fwd_launch(ctx, L[L_DECONV5], S ? t->d4cat : t->deconv4, bias(L_DECONV5), t->deconv5));

        // ---- logits, loss, d logits, gradient of the logits layer and of its input
        LogitsParams lp{};
        lp.a = t->d5cat.u(); lp.planes = t->d5cat.planes(); lp.planes_d5 = 3; lp.H = t->H; lp.W = t->W; lp.hc = t->h; lp.wc = t->w;
        lp.ncls = t->ncls; lp.ci = L[L_LOGITS].def.ci; lp.w = d_params + L[L_LOGITS].w_off; lp.bias = bias(L_LOGITS); lp.ch_map = t->d_logit_map;
        lp.labels = d_labels; lp.dl = t->d_dl; lp.g_d5s = t->g_d5s.u(); lp.g_skip = S ? t->g_conv2.u() : nullptr; lp.loss_sum = d_loss;
        tc_logits_kernel<<<blocks_for((size_t)t->H * t->W), 256, 0, ctx->stream>>>(lp);
        PCS_LAUNCH_CHECK(ctx, "tc_logits_kernel");
        {
            const size_t hw = (size_t)t->H * t->W;
            const dim3 grid(lp.planes, (unsigned)std::max<size_t>(1, std::min<size_t>((hw + 4095) / 4096, 96)));
            tc_logits_wgrad_kernel<<<grid, 256, 0, ctx->stream>>>(lp.a, t->d_dl, hw, t->ncls, lp.ci, t->d_logit_map, d_grads + L[L_LOGITS].w_off,
                                                                 d_grads + L[L_LOGITS].b_off);
            PCS_LAUNCH_CHECK(ctx, "tc_logits_wgrad_kernel");
        }

        // ---- decoder backward
        const float* zb = t->d_zero_bias;
        // deconv5 (linear): dY = g_d5s (space-to-depth)
        PCS_TRY(wgrad_launch(ctx, t, L[L_DECONV5], S ? t->d4cat : t->deconv4, t->g_d5s));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_DECONV5], t->g_d5s, 3));
        PCS_TRY(conv_launch(ctx, L[L_DECONV5].dgrad, t->g_d5s, zb, t->g_d4cat.cp, 0, t->g_d4cat, nullptr, 1));
        // deconv4 (relu)
        { Ten g = view(t->g_d4cat, 0, 4); PCS_TRY(combine_launch(ctx, t->g_d4s, g, nullptr, &t->deconv4, 1)); }
        PCS_TRY(wgrad_launch(ctx, t, L[L_DECONV4], S ? t->d3cat : t->deconv3, t->g_d4s));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_DECONV4], t->g_d4s, 4));
        PCS_TRY(conv_launch(ctx, L[L_DECONV4].dgrad, t->g_d4s, zb, t->g_d3cat.cp, 0, t->g_d3cat, nullptr, 1));
        // deconv3 (relu, 5x5)
        { Ten g = view(t->g_d3cat, 0, 5);
          PCS_TRY(combine_launch(ctx, g, g, nullptr, &t->deconv3, 0));
          PCS_TRY(wgrad_launch(ctx, t, L[L_DECONV3], S ? t->d2cat : t->deconv2, g));
          PCS_TRY(bias_grad_launch(ctx, t, L[L_DECONV3], g, 5));
          PCS_TRY(conv_launch(ctx, L[L_DECONV3].dgrad, g, zb, t->g_d2cat.cp, 0, t->g_d2cat, nullptr, 5)); }
        // deconv2 (relu)
        { Ten g = view(t->g_d2cat, 0, 8); PCS_TRY(combine_launch(ctx, t->g_d2s, g, nullptr, &t->deconv2, 1)); }
        PCS_TRY(wgrad_launch(ctx, t, L[L_DECONV2], t->deconv1, t->g_d2s));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_DECONV2], t->g_d2s, 8));
        PCS_TRY(conv_launch(ctx, L[L_DECONV2].dgrad, t->g_d2s, zb, 80, 0, t->g_deconv1, nullptr, 1));
        // deconv1 (relu, 5x5)
        PCS_TRY(combine_launch(ctx, t->g_deconv1, t->g_deconv1, nullptr, &t->deconv1, 0));
        PCS_TRY(wgrad_launch(ctx, t, L[L_DECONV1], t->conv7, t->g_deconv1));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_DECONV1], t->g_deconv1, 10));
        PCS_TRY(conv_launch(ctx, L[L_DECONV1].dgrad, t->g_deconv1, zb, 80, 0, t->g_conv7, nullptr, 5));
        PCS_TRY(scatter_launch(ctx, t->d_scatter, 0, t->n_scatter_decoder, t->max_scatter_total));
    }
    if (phases & 2) {
        const float* zb = t->d_zero_bias;
        // conv7 (relu)
        PCS_TRY(combine_launch(ctx, t->g_conv7, t->g_conv7, nullptr, &t->conv7, 0));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV7], t->pool6, t->g_conv7));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV7], t->g_conv7, 10));
        PCS_TRY(conv_launch(ctx, L[L_CONV7].dgrad, t->g_conv7, zb, 64, 0, t->g_pool6, nullptr, 5));
        // pool6 -> conv6 (linear); the skip gradient already sits in g_d2cat[8:16]
        Ten g_conv6 = S ? view(t->g_d2cat, 8, 8) : t->g_d2cat;      // fcn: g_d2cat (64) has been consumed by deconv2's backward
        PCS_TRY(pool_bwd_launch(ctx, t->conv6, t->g_pool6, S ? &g_conv6 : nullptr, g_conv6));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV6], t->conv5, g_conv6));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV6], g_conv6, 8));
        PCS_TRY(conv_launch(ctx, L[L_CONV6].dgrad, g_conv6, zb, 64, 0, t->t_conv5, nullptr, 5));
        // conv5 (relu): skip gradient g_d3cat[5:13] + conv6's input gradient
        Ten g_conv5 = S ? view(t->g_d3cat, 5, 8) : t->t_conv5;
        PCS_TRY(combine_launch(ctx, g_conv5, g_conv5, S ? &t->t_conv5 : nullptr, &t->conv5, 0));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV5], t->pool4, g_conv5));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV5], g_conv5, 8));
        PCS_TRY(conv_launch(ctx, L[L_CONV5].dgrad, g_conv5, zb, 40, 0, t->g_pool4, nullptr, 5));
        // pool4 -> conv4 (linear, no skip)
        PCS_TRY(pool_bwd_launch(ctx, t->conv4, t->g_pool4, nullptr, t->g_conv4));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV4], t->conv3, t->g_conv4));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV4], t->g_conv4, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV4].dgrad, t->g_conv4, zb, 40, 0, t->t_conv3, nullptr, 5));
        // conv3 (relu): skip gradient g_d4cat[4:9] + conv4's input gradient
        Ten g_conv3 = S ? view(t->g_d4cat, 4, 5) : t->t_conv3;
        PCS_TRY(combine_launch(ctx, g_conv3, g_conv3, S ? &t->t_conv3 : nullptr, &t->conv3, 0));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV3], t->pool2, g_conv3));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV3], g_conv3, 5));
        PCS_TRY(conv_launch(ctx, L[L_CONV3].dgrad, g_conv3, zb, 32, 0, t->g_pool2, nullptr, 5));
        // pool2 -> conv2 (linear); the skip gradient (from the logits) already sits in g_conv2
        PCS_TRY(pool_bwd_launch(ctx, t->conv2, t->g_pool2, S ? &t->g_conv2 : nullptr, t->g_conv2));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV2], t->conv1, t->g_conv2));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV2], t->g_conv2, 4));
        PCS_TRY(conv_launch(ctx, L[L_CONV2].dgrad, t->g_conv2, zb, 24, 0, t->g_conv1, nullptr, 5));
        // conv1 (relu)
        PCS_TRY(combine_launch(ctx, t->g_conv1, t->g_conv1, nullptr, &t->conv1, 0));
        PCS_TRY(wgrad_launch(ctx, t, L[L_CONV1], t->x, t->g_conv1));
        PCS_TRY(bias_grad_launch(ctx, t, L[L_CONV1], t->g_conv1, 3));
        if (t->scatter_grads != d_grads) return set_err(ctx, PCS_ERR_STATE, "train_tc: phase 2 with a gradient buffer phase 1 has not seen");
        PCS_TRY(scatter_launch(ctx, t->d_scatter, t->n_scatter_decoder, (int)t->h_scatter.size() - t->n_scatter_decoder, t->max_scatter_total));
    }
    return PCS_OK;
}

}  // namespace pcs
