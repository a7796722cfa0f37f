// Training step primitives (BASELINE configs[4]; ocr4all_pixel_classifier/lib/network.py:167-242 `train_dataset`,
// lib/metrics.py:8-9 `loss`, Keras Adam with clipnorm as compiled at network.py:91-103): forward with kept
// activations, backward, loss and the optimizer for the FCN graphs of lib/model.py:45-92 / :206-234.
//
// First version: fp32 on the CUDA cores, one page per step per GPU (the reference trains with batch 1,
// network.py:151-161), planar [C][H][W] tensors.  Every convolution of the graph - Conv2D, Conv2DTranspose with
// stride 1 (a correlation with the flipped kernel), their input gradients (correlations with the transposed /
// flipped kernel) and the 1x1 logits - is ONE 'same' correlation kernel over host-arranged weights
// w[C_out][C_in][k][k]; weight gradients are pixel reductions per (C_out, C_in) pair.  The tensor-core forward of the
// inference path is not used here: training needs the fp32 activations the reference keeps.
#include "common.cuh"

#include <cstdint>

namespace pcs {
namespace {

constexpr int TX = 32, TY = 8, CO_T = 8;

// y[co][r][c] = act(b[co] + sum_ci sum_t x[ci][r + ky - P][c + kx - P] * w[co][ci][t]), zero padding, P = (K - 1) / 2.
// ACC: add to y instead of overwriting (second source of a concatenation, gradient accumulation).
// A block computes a 32 x (TY * RPT) pixel tile for CO_T output channels; a thread owns RPT pixels of one column
// (rows ty, ty + TY, ...), so that the two float4 weight loads of a tap feed RPT * CO_T multiply-adds.
constexpr int RPT = 2;

template <int K>
__global__ void __launch_bounds__(TX * TY)
corr2d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y, int Ci,
              int Co, int H, int W, int relu, int acc_out) {
    constexpr int P = (K - 1) / 2, IW = TX + K - 1, IH = TY * RPT + K - 1, KK = K * K;
    constexpr int CI_T = K == 1 ? 8 : 4;                             // input channels staged per barrier pair
    __shared__ float s_in[CI_T][IH][IW + 1];
    __shared__ __align__(16) float s_w[CI_T][KK][CO_T];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY * RPT, co0 = blockIdx.z * CO_T;
    float acc[RPT][CO_T];
#pragma unroll
    for (int p = 0; p < RPT; ++p)
#pragma unroll
        for (int o = 0; o < CO_T; ++o) acc[p][o] = 0.f;
    for (int ci0 = 0; ci0 < Ci; ci0 += CI_T) {
        const int nci = min(CI_T, Ci - ci0);
        __syncthreads();
        for (int i = tid; i < nci * IH * IW; i += TX * TY) {
            const int q = i / (IH * IW), rem = i - q * IH * IW, r = rem / IW, c = rem - r * IW, gy = y0 + r - P, gx = x0 + c - P;
            s_in[q][r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(x + (size_t)(ci0 + q) * H * W + (size_t)gy * W + gx) : 0.f;
        }
        for (int i = tid; i < nci * KK * CO_T; i += TX * TY) {
            const int q = i / (KK * CO_T), rem = i - q * KK * CO_T, t = rem / CO_T, o = rem - t * CO_T;
            s_w[q][t][o] = (co0 + o < Co) ? __ldg(w + ((size_t)(co0 + o) * Ci + ci0 + q) * KK + t) : 0.f;
        }
        __syncthreads();
        for (int q = 0; q < nci; ++q) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&s_w[q][ky * K + kx][0]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&s_w[q][ky * K + kx][4]);
#pragma unroll
                    for (int p = 0; p < RPT; ++p) {
                        const float v = s_in[q][ty + p * TY + ky][tx + kx];
                        acc[p][0] = fmaf(v, w0.x, acc[p][0]); acc[p][1] = fmaf(v, w0.y, acc[p][1]);
                        acc[p][2] = fmaf(v, w0.z, acc[p][2]); acc[p][3] = fmaf(v, w0.w, acc[p][3]);
                        acc[p][4] = fmaf(v, w1.x, acc[p][4]); acc[p][5] = fmaf(v, w1.y, acc[p][5]);
                        acc[p][6] = fmaf(v, w1.z, acc[p][6]); acc[p][7] = fmaf(v, w1.w, acc[p][7]);
                    }
                }
        }
    }
    const int gx = x0 + tx;
    if (gx >= W) return;
#pragma unroll
    for (int p = 0; p < RPT; ++p) {
        const int gy = y0 + ty + p * TY;
        if (gy >= H) continue;
#pragma unroll
        for (int o = 0; o < CO_T; ++o) {
            if (co0 + o >= Co) break;
            float* dst = y + ((size_t)(co0 + o) * H + gy) * W + gx;
            float v = acc[p][o] + (b ? __ldg(b + co0 + o) : 0.f);
            if (acc_out) v += *dst;
            if (relu) v = fmaxf(v, 0.f);
            *dst = v;
        }
    }
}

// dw[co][ci][t] = sum_{r,c} dy[co][r][c] * x[ci][r + ky - P][c + kx - P].  A block owns WG_CO output channels of one input
// channel over a band of rows: the K*K shifted x values of a pixel are loaded once and used for all WG_CO gradients
// (one block per (co, ci) pair re-read them per output channel, and conv1's 20 pairs left 128 SMs idle); bands are
// combined with atomicAdd into the zero-filled dw.
constexpr int WG_CO = 4;

template <int K>
__global__ void __launch_bounds__(TX * TY)
wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int Ci, int Co, int H, int W, int band) {
    // the band is walked in 32 x 8 pixel tiles staged in shared memory (x with its halo, dy of the WG_CO channels):
    // no bounds tests and no repeated global loads in the 25 x WG_CO inner products
    constexpr int WG_R = 4;                                          // tile rows per thread: 32 x 32 pixel tiles, a quarter of the barriers
    constexpr int P = (K - 1) / 2, KK = K * K, IW = TX + K - 1, IH = TY * WG_R + K - 1;
    __shared__ float s_in[IH][IW + 1];
    __shared__ float s_red[TY][WG_CO * KK];
    const int co0 = blockIdx.x * WG_CO, ci = blockIdx.y;
    const int r0 = blockIdx.z * band, r1 = min(H, r0 + band);
    const float* xp = x + (size_t)ci * H * W;
    const size_t plane = (size_t)H * W;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    float acc[WG_CO][KK];
#pragma unroll
    for (int o = 0; o < WG_CO; ++o)
#pragma unroll
        for (int t = 0; t < KK; ++t) acc[o][t] = 0.f;
    for (int y0 = r0; y0 < r1; y0 += TY * WG_R)
        for (int x0 = 0; x0 < W; x0 += TX) {
            __syncthreads();
            for (int i = tid; i < IH * IW; i += TX * TY) {
                const int r = i / IW, c = i - r * IW, gy = y0 + r - P, gx = x0 + c - P;
                s_in[r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(xp + (size_t)gy * W + gx) : 0.f;
            }
            __syncthreads();
            const int gx = x0 + tx;
            if (gx >= W) continue;
#pragma unroll 1
            for (int q = 0; q < WG_R; ++q) {
                const int ly = ty + q * TY, gy = y0 + ly;
                if (gy >= r1) break;
                float g[WG_CO];
                bool any = false;
#pragma unroll
                for (int o = 0; o < WG_CO; ++o) {
                    g[o] = (co0 + o < Co) ? __ldg(dy + (size_t)(co0 + o) * plane + (size_t)gy * W + gx) : 0.f;
                    any |= g[o] != 0.f;
                }
                if (!any) continue;
#pragma unroll
                for (int ky = 0; ky < K; ++ky)
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        const float xv = s_in[ly + ky][tx + kx];
#pragma unroll
                        for (int o = 0; o < WG_CO; ++o) acc[o][ky * K + kx] = fmaf(g[o], xv, acc[o][ky * K + kx]);
                    }
            }
        }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < WG_CO; ++o)
#pragma unroll
        for (int t = 0; t < KK; ++t) {
            float v = acc[o][t];
            for (int k = 16; k; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
            if (tx == 0) s_red[ty][o * KK + t] = v;
        }
    __syncthreads();
    if (tid < WG_CO * KK) {
        const int o = tid / KK, t = tid - o * KK;
        if (co0 + o < Co) {
            float v = 0.f;
            for (int k = 0; k < TY; ++k) v += s_red[k][tid];
            atomicAdd(dw + ((size_t)(co0 + o) * Ci + ci) * KK + t, v);
        }
    }
}

// db[c] = sum over the plane; one block per channel
__global__ void __launch_bounds__(256) plane_sum_kernel(const float* __restrict__ dy, float* __restrict__ db, size_t plane) {
    const float* p = dy + (size_t)blockIdx.x * plane;
    float v = 0.f;
    for (size_t i = threadIdx.x; i < plane; i += 256) v += p[i];
    __shared__ float s[8];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { for (int k = 1; k < 8; ++k) v += s[k]; db[blockIdx.x] = v; }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        if (!(y[i] > 0.f)) dy[i] = 0.f;
}

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W) {
    const int ho = H / 2, wo = W / 2;
    const size_t n = (size_t)C * ho * wo;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int c = (int)(i / ((size_t)ho * wo)), r = (int)((i / wo) % ho), q = (int)(i % wo);
        const float* p = x + ((size_t)c * H + 2 * r) * W + 2 * q;
        y[i] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
    }
}
// the gradient goes to the first maximum of the window in row-major order (TensorFlow / torch argmax convention)
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int C, int H, int W, int acc_out) {
    const int ho = H / 2, wo = W / 2;
    const size_t n = (size_t)C * ho * wo;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int c = (int)(i / ((size_t)ho * wo)), r = (int)((i / wo) % ho), q = (int)(i % wo);
        const size_t o = ((size_t)c * H + 2 * r) * W + 2 * q;
        const float v[4] = {x[o], x[o + 1], x[o + W], x[o + W + 1]};
        int best = 0;
        for (int k = 1; k < 4; ++k) if (v[k] > v[best]) best = k;
        const float g = dy[i];
        const size_t off[4] = {o, o + 1, o + W, o + W + 1};
        for (int k = 0; k < 4; ++k) {
            const float d = k == best ? g : 0.f;
            dx[off[k]] = acc_out ? dx[off[k]] + d : d;
        }
    }
}

// Conv2DTranspose(2x2, stride 2): y[o][2r+i][2c+j] = act(b[o] + sum_ci x[ci][r][c] * k2[i*2+j][o][ci]).
// The three kernels tile one channel axis by D2_T in registers, so that a value loaded from the big tensor feeds D2_T
// multiply-adds, with the matching weight slice in shared memory (the first versions did one load per multiply-add).
constexpr int D2_T = 8;

// grid = (pixel blocks of the INPUT grid, ceil(Co / D2_T)); a thread owns one input pixel and D2_T output channels (x 4 taps)
__global__ void __launch_bounds__(256)
deconv2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ k2, const float* __restrict__ b, float* __restrict__ y, int Ci,
                   int Co, int h, int w, int relu) {
    extern __shared__ float s_k[];                                  // [4][D2_T][Ci]
    const int o0 = blockIdx.y * D2_T;
    for (int i = threadIdx.x; i < 4 * D2_T * Ci; i += 256) {
        const int tap = i / (D2_T * Ci), rem = i - tap * D2_T * Ci, o = rem / Ci, ci = rem - o * Ci;
        s_k[i] = (o0 + o < Co) ? __ldg(k2 + ((size_t)tap * Co + o0 + o) * Ci + ci) : 0.f;
    }
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x, plane = (size_t)h * w;
    if (p >= plane) return;
    const int r = (int)(p / w), c = (int)(p % w);
    float acc[4][D2_T];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int o = 0; o < D2_T; ++o) acc[t][o] = 0.f;
    for (int ci = 0; ci < Ci; ++ci) {
        const float xv = __ldg(x + (size_t)ci * plane + p);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int o = 0; o < D2_T; ++o) acc[t][o] = fmaf(xv, s_k[(t * D2_T + o) * Ci + ci], acc[t][o]);
    }
#pragma unroll
    for (int o = 0; o < D2_T; ++o) {
        if (o0 + o >= Co) break;
        const float bias = __ldg(b + o0 + o);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = acc[t][o] + bias;
            if (relu) v = fmaxf(v, 0.f);
            y[((size_t)(o0 + o) * 2 * h + 2 * r + (t >> 1)) * 2 * w + 2 * c + (t & 1)] = v;
        }
    }
}

// dx[ci][r][c] = sum_{tap,o} dy[o][2r+i][2c+j] * k2[tap][o][ci]; grid = (pixel blocks, ceil(Ci / D2_T))
__global__ void __launch_bounds__(256)
deconv2_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ k2, float* __restrict__ dx, int Ci, int Co, int h, int w) {
    extern __shared__ float s_k[];                                  // [4][Co][D2_T]
    const int c0 = blockIdx.y * D2_T;
    for (int i = threadIdx.x; i < 4 * Co * D2_T; i += 256) {
        const int to = i / D2_T, q = i - to * D2_T;
        s_k[i] = (c0 + q < Ci) ? __ldg(k2 + (size_t)to * Ci + c0 + q) : 0.f;
    }
    __syncthreads();
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x, plane = (size_t)h * w;
    if (p >= plane) return;
    const int r = (int)(p / w), c = (int)(p % w);
    float acc[D2_T];
#pragma unroll
    for (int q = 0; q < D2_T; ++q) acc[q] = 0.f;
    for (int o = 0; o < Co; ++o) {
        const float* g = dy + ((size_t)o * 2 * h + 2 * r) * 2 * w + 2 * c;
        const float2 top = *reinterpret_cast<const float2*>(g), bot = *reinterpret_cast<const float2*>(g + 2 * w);
        const float gv[4] = {top.x, top.y, bot.x, bot.y};
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int q = 0; q < D2_T; ++q) acc[q] = fmaf(gv[t], s_k[(t * Co + o) * D2_T + q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < D2_T; ++q)
        if (c0 + q < Ci) dx[(size_t)(c0 + q) * plane + p] = acc[q];
}

// dk2[tap][o][ci] = sum_{r,c} dy[o][2r+i][2c+j] * x[ci][r][c]; grid = (Co, ceil(Ci / D2_T)): a block owns one output channel,
// D2_T input channels and all four taps
__global__ void __launch_bounds__(256)
deconv2_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dk2, int Ci, int Co, int h, int w) {
    const int o = blockIdx.x, c0 = blockIdx.y * D2_T;
    const size_t plane = (size_t)h * w;
    float acc[4][D2_T];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int q = 0; q < D2_T; ++q) acc[t][q] = 0.f;
    for (size_t p = threadIdx.x; p < plane; p += 256) {
        const int r = (int)(p / w), c = (int)(p % w);
        const float* g = dy + ((size_t)o * 2 * h + 2 * r) * 2 * w + 2 * c;
        const float2 top = *reinterpret_cast<const float2*>(g), bot = *reinterpret_cast<const float2*>(g + 2 * w);
        const float gv[4] = {top.x, top.y, bot.x, bot.y};
        if (gv[0] == 0.f && gv[1] == 0.f && gv[2] == 0.f && gv[3] == 0.f) continue;
#pragma unroll
        for (int q = 0; q < D2_T; ++q) {
            const float xv = (c0 + q < Ci) ? __ldg(x + (size_t)(c0 + q) * plane + p) : 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t][q] = fmaf(gv[t], xv, acc[t][q]);
        }
    }
    __shared__ float s_red[8][4 * D2_T];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int q = 0; q < D2_T; ++q) {
            float v = acc[t][q];
            for (int k = 16; k; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][t * D2_T + q] = v;
        }
    __syncthreads();
    if (threadIdx.x < 4 * D2_T) {
        const int t = threadIdx.x / D2_T, q = threadIdx.x - t * D2_T;
        if (c0 + q < Ci) {
            float v = 0.f;
            for (int k = 0; k < 8; ++k) v += s_red[k][threadIdx.x];
            dk2[((size_t)t * Co + o) * Ci + c0 + q] = v;
        }
    }
}

// metrics.loss: mean over the Hc x Wc crop of the sparse softmax cross entropy from logits [C][H][W];
// dlogits = (softmax - onehot) / (Hc * Wc) inside the crop, 0 outside; loss_sum accumulates the un-normalised sum
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, int C, int H, int W, int Hc, int Wc,
                  float* __restrict__ dlogits, double* __restrict__ loss_sum) {
    const size_t plane = (size_t)H * W;
    double local = 0.0;
    const float inv = 1.f / ((float)Hc * (float)Wc);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < plane; i += (size_t)gridDim.x * 256) {
        const int r = (int)(i / W), c = (int)(i % W);
        if (r >= Hc || c >= Wc) { for (int k = 0; k < C; ++k) dlogits[k * plane + i] = 0.f; continue; }
        float mx = logits[i];
        for (int k = 1; k < C; ++k) mx = fmaxf(mx, logits[k * plane + i]);
        float sum = 0.f;
        for (int k = 0; k < C; ++k) sum += expf(logits[k * plane + i] - mx);
        const int lab = labels[(size_t)r * Wc + c];
        const float lse = mx + logf(sum);
        local += (double)(lse - logits[(size_t)lab * plane + i]);
        for (int k = 0; k < C; ++k) dlogits[k * plane + i] = (expf(logits[k * plane + i] - lse) - (k == lab ? 1.f : 0.f)) * inv;
    }
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss_sum, local);
}

// x/255 of the uint8 page into the zero-padded fp32 plane (architecture.py:67-68, model.py:20-26)
__global__ void __launch_bounds__(256) input_plane_kernel(const uint8_t* __restrict__ img, int h, int w, float* __restrict__ out, int H, int W) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < (size_t)H * W; i += (size_t)gridDim.x * 256) {
        const int r = (int)(i / W), c = (int)(i % W);
        out[i] = (r < h && c < w) ? (float)img[(size_t)r * w + c] / 255.0f : 0.f;
    }
}

// Keras Adam with clipnorm (TF <= 2.5 clips every variable separately).  grid = (variables, kAdamBlocks): first the
// squared norms (atomicAdd per block into sumsq[var]), then  m, v, p  with lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t),
// p -= lr_t * m / (sqrt(v) + eps)
constexpr int kAdamBlocks = 16;

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, const long long* __restrict__ offs, float gscale, float* __restrict__ sumsq) {
    const long long a = offs[blockIdx.x], b = offs[blockIdx.x + 1];
    float ss = 0.f;
    for (long long i = a + (long long)blockIdx.y * 256 + threadIdx.x; i < b; i += 256LL * gridDim.y) { const float gi = g[i] * gscale; ss = fmaf(gi, gi, ss); }
    __shared__ float s[8];
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) ss += s[k];
        if (ss != 0.f) atomicAdd(&sumsq[blockIdx.x], ss);
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, const long long* __restrict__ offs,
            const float* __restrict__ sumsq, float lr_t, float b1, float b2, float eps, float clipnorm, float gscale) {
    const long long a = offs[blockIdx.x], b = offs[blockIdx.x + 1];
    const float norm = sqrtf(sumsq[blockIdx.x]);
    const float sc = ((clipnorm > 0.f && norm > clipnorm) ? clipnorm / norm : 1.f) * gscale;       // tf.clip_by_norm
    for (long long i = a + (long long)blockIdx.y * 256 + threadIdx.x; i < b; i += 256LL * gridDim.y) {
        const float gi = g[i] * sc;
        const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

inline unsigned blocks_for(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16); }

}  // namespace

int train_corr2d(pcs_ctx* ctx, const float* x, const float* w, const float* b, float* y, int Ci, int Co, int H, int W, int k, int relu, int acc) {
    if (Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || (k != 1 && k != 5)) return set_err(ctx, PCS_ERR_ARG, "train_corr2d: bad shape or kernel size %d", k);
    dim3 grid((W + TX - 1) / TX, (H + TY * RPT - 1) / (TY * RPT), (Co + CO_T - 1) / CO_T), block(TX, TY);
    if (grid.y > 65535 || grid.z > 65535) return set_err(ctx, PCS_ERR_ARG, "train_corr2d: grid too large");
    if (k == 5) corr2d_kernel<5><<<grid, block, 0, ctx->stream>>>(x, w, b, y, Ci, Co, H, W, relu, acc);
    else corr2d_kernel<1><<<grid, block, 0, ctx->stream>>>(x, w, b, y, Ci, Co, H, W, relu, acc);
    PCS_LAUNCH_CHECK(ctx, "corr2d_kernel");
    return PCS_OK;
}

int train_wgrad(pcs_ctx* ctx, const float* x, const float* dy, float* dw, int Ci, int Co, int H, int W, int k) {
    if (Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || (k != 1 && k != 5) || Ci > 65535) return set_err(ctx, PCS_ERR_ARG, "train_wgrad: bad shape");
    PCS_CUDA(ctx, cudaMemsetAsync(dw, 0, (size_t)Co * Ci * k * k * sizeof(float), ctx->stream));
    const int pairs = ((Co + WG_CO - 1) / WG_CO) * Ci;
    int bands = std::max(1, std::min((H + 7) / 8, (148 * 6 + pairs - 1) / pairs));       // enough blocks to fill the GPU a few times
    const int band = ((H + bands - 1) / bands + 31) / 32 * 32;
    bands = (H + band - 1) / band;
    const dim3 grid((Co + WG_CO - 1) / WG_CO, Ci, bands);
    if (k == 5) wgrad_kernel<5><<<grid, dim3(TX, TY), 0, ctx->stream>>>(x, dy, dw, Ci, Co, H, W, band);
    else wgrad_kernel<1><<<grid, dim3(TX, TY), 0, ctx->stream>>>(x, dy, dw, Ci, Co, H, W, band);
    PCS_LAUNCH_CHECK(ctx, "wgrad_kernel");
    return PCS_OK;
}

int train_plane_sum(pcs_ctx* ctx, const float* dy, float* db, int C, size_t plane) {
    plane_sum_kernel<<<C, 256, 0, ctx->stream>>>(dy, db, plane);
    PCS_LAUNCH_CHECK(ctx, "plane_sum_kernel");
    return PCS_OK;
}

int train_relu_bwd(pcs_ctx* ctx, float* dy, const float* y, size_t n) {
    relu_bwd_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(dy, y, n);
    PCS_LAUNCH_CHECK(ctx, "relu_bwd_kernel");
    return PCS_OK;
}

int train_maxpool(pcs_ctx* ctx, const float* x, float* y, const float* dy, float* dx, int C, int H, int W, int acc) {
    if ((H | W) & 1) return set_err(ctx, PCS_ERR_ARG, "train_maxpool: odd plane");
    const size_t n = (size_t)C * (H / 2) * (W / 2);
    if (dx) maxpool_bwd_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(x, dy, dx, C, H, W, acc);
    else maxpool_fwd_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(x, y, C, H, W);
    PCS_LAUNCH_CHECK(ctx, "maxpool kernel");
    return PCS_OK;
}

int train_deconv2(pcs_ctx* ctx, int mode, const float* x, const float* k2, const float* b, float* y, const float* dy, float* dx, float* dk2,
                  int Ci, int Co, int h, int w, int relu) {
    const unsigned pblocks = (unsigned)(((size_t)h * w + 255) / 256);
    const size_t smem = (size_t)4 * D2_T * (mode == 0 ? Ci : Co) * sizeof(float);
    if (smem > 48 * 1024) return set_err(ctx, PCS_ERR_ARG, "train_deconv2: %d channels exceed the shared-memory weight slice", mode == 0 ? Ci : Co);
    if (mode != 0 && (reinterpret_cast<uintptr_t>(dy) & 7)) return set_err(ctx, PCS_ERR_ARG, "train_deconv2: dy must be 8-byte aligned");
    if (mode == 0) deconv2_fwd_kernel<<<dim3(pblocks, (Co + D2_T - 1) / D2_T), 256, smem, ctx->stream>>>(x, k2, b, y, Ci, Co, h, w, relu);
    else if (mode == 1) deconv2_bwd_data_kernel<<<dim3(pblocks, (Ci + D2_T - 1) / D2_T), 256, smem, ctx->stream>>>(dy, k2, dx, Ci, Co, h, w);
    else deconv2_wgrad_kernel<<<dim3(Co, (Ci + D2_T - 1) / D2_T), 256, 0, ctx->stream>>>(x, dy, dk2, Ci, Co, h, w);
    PCS_LAUNCH_CHECK(ctx, "deconv2 kernel");
    return PCS_OK;
}

int train_softmax_ce(pcs_ctx* ctx, const float* logits, const uint8_t* labels, int C, int H, int W, int Hc, int Wc, float* dlogits, double* loss_sum) {
    PCS_CUDA(ctx, cudaMemsetAsync(loss_sum, 0, sizeof(double), ctx->stream));
    softmax_ce_kernel<<<blocks_for((size_t)H * W), 256, 0, ctx->stream>>>(logits, labels, C, H, W, Hc, Wc, dlogits, loss_sum);
    PCS_LAUNCH_CHECK(ctx, "softmax_ce_kernel");
    return PCS_OK;
}

int train_input_plane(pcs_ctx* ctx, const uint8_t* img, int h, int w, float* out, int H, int W) {
    input_plane_kernel<<<blocks_for((size_t)H * W), 256, 0, ctx->stream>>>(img, h, w, out, H, W);
    PCS_LAUNCH_CHECK(ctx, "input_plane_kernel");
    return PCS_OK;
}

int train_adam(pcs_ctx* ctx, float* p, const float* g, float* m, float* v, const long long* d_offsets, int nvars, float lr_t, float b1,
               float b2, float eps, float clipnorm, float gscale) {
    PCS_TRY(scratch_reserve(ctx, (size_t)nvars * sizeof(float) + 256));
    float* sumsq = reinterpret_cast<float*>(ctx->scratch);
    PCS_CUDA(ctx, cudaMemsetAsync(sumsq, 0, (size_t)nvars * sizeof(float), ctx->stream));
    grad_sumsq_kernel<<<dim3(nvars, kAdamBlocks), 256, 0, ctx->stream>>>(g, d_offsets, gscale, sumsq);
    PCS_LAUNCH_CHECK(ctx, "grad_sumsq_kernel");
    adam_kernel<<<dim3(nvars, kAdamBlocks), 256, 0, ctx->stream>>>(p, g, m, v, d_offsets, sumsq, lr_t, b1, b2, eps, clipnorm, gscale);
    PCS_LAUNCH_CHECK(ctx, "adam_kernel");
    return PCS_OK;
}

}  // namespace pcs
