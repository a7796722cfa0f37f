// Network head + colour epilogue.
//
// Fuses, per output pixel of the cropped grid (model.py:29-42 crop):
//   [FCN] deconv5 = Conv2DTranspose(20, 2x2, s2, linear) over concat[deconv4, conv3]   model.py:83
//         concat [deconv5, conv2]                                                      model.py:85
//   logits = Conv2D(n_classes, 1x1) + bias                                            model.py:88 / :199 / :231
//   prob = softmax(logit), pred = argmax(logit) (first max wins)                       network.py:258-259
//   color = LUT[pred]; overlay[(1-binary)==0] = 0; inverted[binary==0] = 0             output.py:44-60
// The deconv5 result stays in fp32 registers; logits weights are fp32.
#include "common.cuh"

namespace pcs {

template <typename T> __device__ __forceinline__ float hf(T v);
template <> __device__ __forceinline__ float hf<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float hf<__half>(__half v) { return __half2float(v); }

constexpr int DCO = 20;      // deconv5 output channels of both FCN variants

struct HeadParams {
    int has_deconv;
    const void* d0; const void* d1; int dc0, dcp0, dc1, dcp1;
    const float* dw32; const float* db32; int dcin;
    const void* skip; int skip_c, skip_cp; int has_skip;
    const float* lw32; const float* lb32; int n_classes;
    int hp, wp, h, w;
    const uint8_t* binary; uint8_t* labels; float* logits; float* prob;
    const uint8_t* lut; uint8_t* color; uint8_t* overlay; uint8_t* inverted;
};

template <typename T, int NC>
__global__ void __launch_bounds__(256) head_kernel(HeadParams p) {
    extern __shared__ __align__(16) float s_head[];
    // layout: dw [4][dcin][DCO] | db [DCO] | lw [cin_total][NC] | lb [NC]
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int dcin = p.has_deconv ? p.dcin : 0;
    float* s_dw = s_head;
    float* s_db = s_dw + 4 * dcin * DCO;
    float* s_lw = s_db + DCO;
    const int cin_total = (p.has_deconv ? DCO : 0) + (p.has_skip ? p.skip_c : 0);
    float* s_lb = s_lw + cin_total * NC;
    for (int i = tid; i < 4 * dcin * DCO; i += 256) s_dw[i] = __ldg(p.dw32 + i);
    if (p.has_deconv && tid < DCO) s_db[tid] = __ldg(p.db32 + tid);
    for (int i = tid; i < cin_total * NC; i += 256) {
        const int k = i % NC, c = i / NC;
        s_lw[i] = k < p.n_classes ? __ldg(p.lw32 + (size_t)c * p.n_classes + k) : 0.f;
    }
    if (tid < NC) s_lb[tid] = tid < p.n_classes ? __ldg(p.lb32 + tid) : 0.f;
    __syncthreads();

    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, page = blockIdx.z;
    if (x >= p.w || y >= p.h) return;

    float lg[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) lg[k] = s_lb[k];
    int row = 0;
    if (p.has_deconv) {
        float d5[DCO];
#pragma unroll
        for (int o = 0; o < DCO; ++o) d5[o] = s_db[o];
        const int t = (y & 1) * 2 + (x & 1);
        const int hh = p.hp / 2, wh = p.wp / 2;
        const T* a0 = reinterpret_cast<const T*>(p.d0);
        const T* a1 = reinterpret_cast<const T*>(p.d1);
        for (int c = 0; c < dcin; ++c) {
            const float a = hf(c < p.dc0 ? a0[act_idx(page, p.dcp0, hh, wh, c, y >> 1, x >> 1)]
                                         : a1[act_idx(page, p.dcp1, hh, wh, c - p.dc0, y >> 1, x >> 1)]);
            const float4* wv = reinterpret_cast<const float4*>(s_dw + ((size_t)t * dcin + c) * DCO);
#pragma unroll
            for (int q = 0; q < DCO / 4; ++q) {
                const float4 w4 = wv[q];
                d5[4 * q + 0] = fmaf(a, w4.x, d5[4 * q + 0]);
                d5[4 * q + 1] = fmaf(a, w4.y, d5[4 * q + 1]);
                d5[4 * q + 2] = fmaf(a, w4.z, d5[4 * q + 2]);
                d5[4 * q + 3] = fmaf(a, w4.w, d5[4 * q + 3]);
            }
        }
#pragma unroll
        for (int o = 0; o < DCO; ++o)
#pragma unroll
            for (int k = 0; k < NC; ++k) lg[k] = fmaf(d5[o], s_lw[o * NC + k], lg[k]);
        row = DCO;
    }
    if (p.has_skip) {
        // one 128-bit load per 8-channel plane (the plane-major layout keeps a pixel's 8 channels in one 16-byte unit): a
        // 2-byte load per channel cost the U-Net head (64 channels per pixel) half a millisecond per 8 pages
        const T* s = reinterpret_cast<const T*>(p.skip);
        for (int c0 = 0; c0 < p.skip_c; c0 += 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(s + act_idx(page, p.skip_cp, p.hp, p.wp, c0, y, x)));
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (c0 + i >= p.skip_c) break;
                const float a = hf(e[i]);
#pragma unroll
                for (int k = 0; k < NC; ++k) lg[k] = fmaf(a, s_lw[(row + c0 + i) * NC + k], lg[k]);
            }
        }
    }

    // argmax, first maximum wins (np.argmax)
    int best = 0;
    float bv = lg[0];
#pragma unroll
    for (int k = 1; k < NC; ++k)
        if (k < p.n_classes && lg[k] > bv) { bv = lg[k]; best = k; }
    const size_t opix = ((size_t)page * p.h + y) * p.w + x;
    if (p.labels) p.labels[opix] = (uint8_t)best;
    if (p.logits)
        for (int k = 0; k < p.n_classes; ++k) p.logits[opix * p.n_classes + k] = lg[k];
    if (p.prob) {
        float e[NC], sum = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            e[k] = k < p.n_classes ? expf(lg[k] - bv) : 0.f;
            sum += e[k];
        }
        for (int k = 0; k < p.n_classes; ++k) p.prob[opix * p.n_classes + k] = e[k] / sum;
    }
    if (p.color || p.overlay || p.inverted) {
        const uint8_t r = p.lut[best * 3 + 0], g = p.lut[best * 3 + 1], b = p.lut[best * 3 + 2];
        const uint8_t bin = p.binary ? p.binary[opix] : 1;
        if (p.color) { p.color[opix * 3 + 0] = r; p.color[opix * 3 + 1] = g; p.color[opix * 3 + 2] = b; }
        if (p.overlay) {
            // overlay[(1 - binary) == 0] = 0  -> kept where binary != 1
            const bool keep = (uint8_t)(1 - bin) != 0;
            p.overlay[opix * 3 + 0] = keep ? r : 0; p.overlay[opix * 3 + 1] = keep ? g : 0; p.overlay[opix * 3 + 2] = keep ? b : 0;
        }
        if (p.inverted) {
            const bool keep = bin != 0;                     // inverted[binary == 0] = 0
            p.inverted[opix * 3 + 0] = keep ? r : 0; p.inverted[opix * 3 + 1] = keep ? g : 0; p.inverted[opix * 3 + 2] = keep ? b : 0;
        }
    }
}

template <typename T, int NC>
static int launch_head_nc(pcs_ctx* ctx, const HeadParams& p, int n) {
    const int dcin = p.has_deconv ? p.dcin : 0;
    const int cin_total = (p.has_deconv ? DCO : 0) + (p.has_skip ? p.skip_c : 0);
    const size_t smem = ((size_t)4 * dcin * DCO + DCO + (size_t)cin_total * NC + NC) * sizeof(float);
    if (smem > 200 * 1024) return set_err(ctx, PCS_ERR_ARG, "head: weights do not fit in shared memory");
    if (smem > 48 * 1024)
        PCS_CUDA(ctx, cudaFuncSetAttribute(head_kernel<T, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.w + 31) / 32, (p.h + 7) / 8, n), block(32, 8);
    head_kernel<T, NC><<<grid, block, smem, ctx->stream>>>(p);
    PCS_LAUNCH_CHECK(ctx, "head_kernel");
    return PCS_OK;
}

template <typename T>
static int launch_head_t(pcs_ctx* ctx, const HeadParams& p, int n) {
    const int nc = p.n_classes;
    if (nc <= 2) return launch_head_nc<T, 2>(ctx, p, n);
    if (nc <= 3) return launch_head_nc<T, 3>(ctx, p, n);
    if (nc <= 4) return launch_head_nc<T, 4>(ctx, p, n);
    if (nc <= 6) return launch_head_nc<T, 6>(ctx, p, n);
    if (nc <= 8) return launch_head_nc<T, 8>(ctx, p, n);
    if (nc <= 16) return launch_head_nc<T, 16>(ctx, p, n);
    return set_err(ctx, PCS_ERR_ARG, "head: n_classes %d > %d", nc, kMaxClasses);
}

int launch_head(pcs_ctx* ctx, const HeadArgs& a) {
    if (a.has_deconv && a.dcout != DCO) return set_err(ctx, PCS_ERR_ARG, "head: fused deconv expects %d channels", DCO);
    if (a.has_deconv && (a.dcin % 1)) return PCS_ERR_ARG;
    HeadParams p{};
    p.has_deconv = a.has_deconv;
    p.d0 = a.dsrc[0].p; p.dc0 = a.dsrc[0].c; p.dcp0 = a.dsrc[0].cp;
    p.d1 = a.dnsrc > 1 ? a.dsrc[1].p : nullptr; p.dc1 = a.dnsrc > 1 ? a.dsrc[1].c : 0; p.dcp1 = a.dnsrc > 1 ? a.dsrc[1].cp : 0;
    p.dw32 = a.dw32; p.db32 = a.db32; p.dcin = a.dcin;
    p.skip = a.skip.p; p.skip_c = a.skip.c; p.skip_cp = a.skip.cp; p.has_skip = a.has_skip;
    p.lw32 = a.lw32; p.lb32 = a.lb32; p.n_classes = a.n_classes;
    p.hp = a.hp; p.wp = a.wp; p.h = a.h; p.w = a.w;
    p.binary = a.binary; p.labels = a.labels; p.logits = a.logits; p.prob = a.prob;
    p.lut = a.lut; p.color = a.color; p.overlay = a.overlay; p.inverted = a.inverted;
    if (ctx->precision == PCS_PREC_BF16) return launch_head_t<__nv_bfloat16>(ctx, p, a.n);
    return launch_head_t<__half>(ctx, p, a.n);
}

// ---------------------------------------------------------------------------
// generate_output_masks on an existing class map (output.py:44-60).
// Labels not present in the LUT map to (0,0,0).
// ---------------------------------------------------------------------------
// One thread = 16 consecutive pixels: one aligned 128-bit load of the class bytes and one of the binary bytes,
// three aligned 128-bit stores (48 bytes) per image; every register index is a compile-time constant.
__global__ void __launch_bounds__(256)
masks_kernel(const uint8_t* __restrict__ labels, const uint8_t* __restrict__ binary, size_t npix,
             const uint8_t* __restrict__ lut, int n_lut, uint8_t* __restrict__ color, uint8_t* __restrict__ overlay,
             uint8_t* __restrict__ inverted, int vec_ok) {
    __shared__ uint32_t s_lut[256];                               // r | g << 8 | b << 16
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_lut[i] = i < n_lut ? ((uint32_t)lut[i * 3] | ((uint32_t)lut[i * 3 + 1] << 8) | ((uint32_t)lut[i * 3 + 2] << 16)) : 0u;
    __syncthreads();
    const size_t nchunks = (npix + 15) / 16;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < nchunks; k += (size_t)gridDim.x * blockDim.x) {
        const size_t p0 = k * 16;
        const bool full = vec_ok && p0 + 16 <= npix;
        uint32_t lw[4], bw[4];
        if (full) {
            const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(labels + p0));
            lw[0] = l4.x; lw[1] = l4.y; lw[2] = l4.z; lw[3] = l4.w;
            if (binary) {
                const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(binary + p0));
                bw[0] = b4.x; bw[1] = b4.y; bw[2] = b4.z; bw[3] = b4.w;
            } else {
                bw[0] = bw[1] = bw[2] = bw[3] = 0x01010101u;
            }
        } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                lw[w] = 0; bw[w] = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const size_t pi = p0 + w * 4 + q;
                    const uint32_t l = pi < npix ? labels[pi] : 0, b = (pi < npix && binary) ? binary[pi] : 1;
                    lw[w] |= l << (8 * q); bw[w] |= b << (8 * q);
                }
            }
        }
        // 16 pixels x 3 bytes = 12 words per image; pixel q contributes bytes 3q .. 3q+2
        uint32_t wc[12], wo[12], wi[12];
#pragma unroll
        for (int w = 0; w < 12; ++w) { wc[w] = 0; wo[w] = 0; wi[w] = 0; }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint32_t lab = (lw[q >> 2] >> (8 * (q & 3))) & 0xffu, bin = (bw[q >> 2] >> (8 * (q & 3))) & 0xffu;
            const uint32_t v = s_lut[lab];
            const uint32_t vo = ((uint8_t)(1u - bin) != 0) ? v : 0u;       // overlay[(1 - binary) == 0] = 0
            const uint32_t vi = bin != 0 ? v : 0u;                          // inverted[binary == 0] = 0
            const int byte0 = 3 * q, w0 = byte0 >> 2, sh = 8 * (byte0 & 3);
            wc[w0] |= v << sh; wo[w0] |= vo << sh; wi[w0] |= vi << sh;
            if (sh > 8) {                                                    // the 3 bytes straddle a word boundary
                wc[w0 + 1] |= v >> (32 - sh); wo[w0 + 1] |= vo >> (32 - sh); wi[w0 + 1] |= vi >> (32 - sh);
            }
        }
        const size_t b0 = p0 * 3;
        if (full) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (color) *reinterpret_cast<uint4*>(color + b0 + 16 * t) = make_uint4(wc[4 * t], wc[4 * t + 1], wc[4 * t + 2], wc[4 * t + 3]);
                if (overlay) *reinterpret_cast<uint4*>(overlay + b0 + 16 * t) = make_uint4(wo[4 * t], wo[4 * t + 1], wo[4 * t + 2], wo[4 * t + 3]);
                if (inverted) *reinterpret_cast<uint4*>(inverted + b0 + 16 * t) = make_uint4(wi[4 * t], wi[4 * t + 1], wi[4 * t + 2], wi[4 * t + 3]);
            }
        } else {
            const size_t nbytes = npix * 3;
#pragma unroll
            for (int b = 0; b < 48; ++b) {
                if (b0 + b < nbytes) {
                    if (color) color[b0 + b] = (uint8_t)(wc[b >> 2] >> (8 * (b & 3)));
                    if (overlay) overlay[b0 + b] = (uint8_t)(wo[b >> 2] >> (8 * (b & 3)));
                    if (inverted) inverted[b0 + b] = (uint8_t)(wi[b >> 2] >> (8 * (b & 3)));
                }
            }
        }
    }
}

int launch_masks(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W,
                 const uint8_t* d_lut, int n_lut, uint8_t* d_color, uint8_t* d_overlay, uint8_t* d_inverted) {
    if (n <= 0 || H <= 0 || W <= 0 || n_lut < 0 || n_lut > 256) return set_err(ctx, PCS_ERR_ARG, "masks: bad argument");
    const size_t npix = (size_t)n * H * W;
    const size_t nchunks = (npix + 15) / 16;
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 16, (nchunks + 255) / 256);
    const int vec_ok = (((uintptr_t)d_color | (uintptr_t)d_overlay | (uintptr_t)d_inverted | (uintptr_t)d_labels | (uintptr_t)d_binary) & 15) == 0;
    masks_kernel<<<blocks, 256, 0, ctx->stream>>>(d_labels, d_binary, npix, d_lut, n_lut, d_color, d_overlay, d_inverted, vec_ok);
    PCS_LAUNCH_CHECK(ctx, "masks_kernel");
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// Evaluation counts behind fgpa / fgoverlap_per_class (lib/image_ops.py:8-55): over the foreground pixels (bin != 0)
// out[0] = their number, out[1] = those with pred != mask, out[2 + p * (n + 2) + m] = the confusion matrix of
// (pred, mask) with every class value above n_classes folded into the last bucket.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eval_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ bin, size_t n,
                   int n_classes, unsigned long long* __restrict__ out) {
    extern __shared__ unsigned int s_cnt[];                       // [2 + (n_classes + 2)^2]
    const int nb = n_classes + 2, total = 2 + nb * nb;
    for (int i = threadIdx.x; i < total; i += 256) s_cnt[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        if (!bin[i]) continue;
        const int p = pred[i], m = mask[i];
        atomicAdd(&s_cnt[0], 1u);
        if (p != m) atomicAdd(&s_cnt[1], 1u);
        atomicAdd(&s_cnt[2 + min(p, nb - 1) * nb + min(m, nb - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += 256)
        if (s_cnt[i]) atomicAdd(&out[i], (unsigned long long)s_cnt[i]);
}

// counts the fp16 values of an activation tensor that sit at the saturation bound +-65504 (0x7BFF): pack2<__half> stores
// anything beyond the fp16 range there (umma_ptx.cuh), and an exact 65504 in a healthy network is as good as impossible
__global__ void __launch_bounds__(256) saturation_scan_kernel(const uint4* __restrict__ a, size_t units, unsigned long long* __restrict__ count) {
    unsigned local = 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < units; i += (size_t)gridDim.x * 256) {
        const uint4 v = __ldg(a + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) local += __popc(__vcmpeq2(w[k] & 0x7fff7fffu, 0x7bff7bffu)) >> 4;
    }
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
}

int launch_saturation_scan(pcs_ctx* ctx, const void* d_act, size_t n_halves, unsigned long long* d_count) {
    const size_t units = n_halves / 8;
    if (!units) return PCS_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((units + 255) / 256, 148 * 16);
    saturation_scan_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(d_act), units, d_count);
    PCS_LAUNCH_CHECK(ctx, "saturation_scan_kernel");
    return PCS_OK;
}

int launch_eval_counts(pcs_ctx* ctx, const uint8_t* d_pred, const uint8_t* d_mask, const uint8_t* d_bin, size_t n, int n_classes,
                       unsigned long long* d_out) {
    if (n_classes < 1 || n_classes > 62) return set_err(ctx, PCS_ERR_ARG, "eval_counts: 1..62 classes");
    const int nb = n_classes + 2;
    const size_t words = 2 + (size_t)nb * nb;
    PCS_CUDA(ctx, cudaMemsetAsync(d_out, 0, words * sizeof(unsigned long long), ctx->stream));
    // a block's shared counters are 32 bit: keep a block's share of the pixels below 2^32
    const unsigned blocks = (unsigned)std::max<size_t>(std::min<size_t>((size_t)ctx->sm_count * 8, (n + 255) / 256), (n >> 31) + 1);
    eval_counts_kernel<<<blocks, 256, words * sizeof(unsigned int), ctx->stream>>>(d_pred, d_mask, d_bin, n, n_classes, d_out);
    PCS_LAUNCH_CHECK(ctx, "eval_counts_kernel");
    return PCS_OK;
}

}  // namespace pcs
