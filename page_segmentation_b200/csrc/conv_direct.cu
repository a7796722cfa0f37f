// CUDA-core (fp32 FMA) direct convolution kernels.
//
// Roles: (1) the first layer (C_in = 1, uint8 input, K = 25) which is not
// tensor-core shaped; (2) the numerics twin of the tcgen05 path: same operand
// rounding (activations/weights in bf16|fp16, fp32 accumulate), selectable with
// pcs_set_engine(PCS_ENGINE_DIRECT), used by the parity tests to localise
// differences layer by layer.
//
// TF/Keras semantics restated from ocr4all_pixel_classifier/lib/model.py:45-92:
// Conv2D 'same' (cross-correlation), Conv2DTranspose k5 s1 (pre-flipped into
// correlation form on the host), Conv2DTranspose k2 s2, MaxPooling2D(2,2).
#include "common.cuh"

namespace pcs {

__constant__ float c_u8_lut[256];    // (float)(v / 255.0): architecture.py:67-68 at the float32 boundary
static bool g_lut_ready[64] = {false};

static int ensure_lut(pcs_ctx* ctx) {
    if (ctx->device < 64 && g_lut_ready[ctx->device]) return PCS_OK;
    float h[256];
    for (int v = 0; v < 256; ++v) h[v] = (float)((double)v / 255.0);
    PCS_CUDA(ctx, cudaMemcpyToSymbol(c_u8_lut, h, sizeof(h)));
    if (ctx->device < 64) g_lut_ready[ctx->device] = true;
    return PCS_OK;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

constexpr int COT = 16;     // output channels per block
constexpr int CC = 8;       // input channels per smem chunk
constexpr int TW = 32;      // tile width  (pixels)
constexpr int TH = 16;      // tile height (pixels): thread (tx,ty) owns rows 2ty, 2ty+1
constexpr int KMAX = 5;

struct DirectParams {
    const void* s0; const void* s1;
    int c0, cp0, c1, cp1;
    int img_h, img_w;       // uint8 input bounds
    int upsample;
    int h, w, k, pad;
    const float* w32; const float* b32;
    int cin, cout, relu;
    void* out; int out_cp;
    void* pool; int pool_cp;
    int cout_tiles;
};

template <typename T, bool SRC_U8>
__global__ void __launch_bounds__(256) conv_direct_kernel(DirectParams p) {
    __shared__ float s_in[CC][TH + KMAX - 1][TW + KMAX - 1 + 1];
    __shared__ __align__(16) float s_w[KMAX * KMAX][CC][COT];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + tx;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int page = blockIdx.z / p.cout_tiles;
    const int ot = blockIdx.z % p.cout_tiles;
    const int o0 = ot * COT;
    const int k = p.k, taps = k * k;
    const int ph = TH + k - 1, pw = TW + k - 1;
    const int in_h = p.upsample ? p.h / 2 : p.h, in_w = p.upsample ? p.w / 2 : p.w;

    float acc[2][COT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int o = 0; o < COT; ++o) acc[r][o] = 0.f;

    for (int cbase = 0; cbase < p.cin; cbase += CC) {
        __syncthreads();
        // ---- stage the input patch chunk ----
        for (int i = tid; i < ph * pw * CC; i += 256) {
            const int c = i % CC;
            const int px = (i / CC) % pw;
            const int py = i / (CC * pw);
            const int ci = cbase + c;
            int gy = y0 + py - p.pad, gx = x0 + px - p.pad;
            float v = 0.f;
            if (ci < p.cin && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) {
                if (p.upsample) { gy >>= 1; gx >>= 1; }
                if (SRC_U8) {
                    if (gy < p.img_h && gx < p.img_w)
                        v = c_u8_lut[reinterpret_cast<const uint8_t*>(p.s0)[((size_t)page * p.img_h + gy) * p.img_w + gx]];
                } else if (ci < p.c0) {
                    v = to_f(reinterpret_cast<const T*>(p.s0)[act_idx(page, p.cp0, in_h, in_w, ci, gy, gx)]);
                } else {
                    v = to_f(reinterpret_cast<const T*>(p.s1)[act_idx(page, p.cp1, in_h, in_w, ci - p.c0, gy, gx)]);
                }
            }
            s_in[c][py][px] = v;
        }
        // ---- stage the weight chunk [tap][c][o] ----
        for (int i = tid; i < taps * CC * COT; i += 256) {
            const int o = i % COT;
            const int c = (i / COT) % CC;
            const int t = i / (COT * CC);
            const int ci = cbase + c, oo = o0 + o;
            s_w[t][c][o] = (ci < p.cin && oo < p.cout) ? __ldg(p.w32 + ((size_t)t * p.cin + ci) * p.cout + oo) : 0.f;
        }
        __syncthreads();
        const int cmax = min(CC, p.cin - cbase);
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx) {
                const int t = dy * k + dx;
                for (int c = 0; c < cmax; ++c) {
                    const float a0 = s_in[c][2 * ty + dy][tx + dx];
                    const float a1 = s_in[c][2 * ty + 1 + dy][tx + dx];
                    const float4* wv = reinterpret_cast<const float4*>(&s_w[t][c][0]);
#pragma unroll
                    for (int q = 0; q < COT / 4; ++q) {
                        const float4 w4 = wv[q];
                        acc[0][4 * q + 0] = fmaf(a0, w4.x, acc[0][4 * q + 0]);
                        acc[0][4 * q + 1] = fmaf(a0, w4.y, acc[0][4 * q + 1]);
                        acc[0][4 * q + 2] = fmaf(a0, w4.z, acc[0][4 * q + 2]);
                        acc[0][4 * q + 3] = fmaf(a0, w4.w, acc[0][4 * q + 3]);
                        acc[1][4 * q + 0] = fmaf(a1, w4.x, acc[1][4 * q + 0]);
                        acc[1][4 * q + 1] = fmaf(a1, w4.y, acc[1][4 * q + 1]);
                        acc[1][4 * q + 2] = fmaf(a1, w4.z, acc[1][4 * q + 2]);
                        acc[1][4 * q + 3] = fmaf(a1, w4.w, acc[1][4 * q + 3]);
                    }
                }
            }
    }

    // ---- epilogue: bias, activation, rounding, store, optional 2x2 max-pool ----
    const int x = x0 + tx;
    const int yA = y0 + 2 * ty;
    T vals[2][COT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int o = 0; o < COT; ++o) {
            float v = 0.f;
            if (o0 + o < p.cout) {
                v = acc[r][o] + __ldg(p.b32 + o0 + o);
                if (p.relu) v = fmaxf(v, 0.f);
            }
            vals[r][o] = from_f<T>(v);
        }
    if (p.out && x < p.w) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int y = yA + r;
            if (y < p.h) {
                T* base = reinterpret_cast<T*>(p.out);
                const uint4* s4 = reinterpret_cast<const uint4*>(&vals[r][0]);
                *reinterpret_cast<uint4*>(base + act_idx(page, p.out_cp, p.h, p.w, o0, y, x)) = s4[0];
                *reinterpret_cast<uint4*>(base + act_idx(page, p.out_cp, p.h, p.w, o0 + 8, y, x)) = s4[1];
            }
        }
    }
    if (p.pool) {
        // rows yA, yA+1 are in this thread; x pair via shuffle (h, w are even on the /32-padded grid)
        T pooled[COT];
#pragma unroll
        for (int o = 0; o < COT; ++o) {
            float m = fmaxf(to_f(vals[0][o]), to_f(vals[1][o]));
            float other = __shfl_xor_sync(0xffffffffu, m, 1);
            pooled[o] = from_f<T>(fmaxf(m, other));
        }
        if ((tx & 1) == 0 && x < p.w && yA < p.h) {
            T* base = reinterpret_cast<T*>(p.pool);
            const uint4* s4 = reinterpret_cast<const uint4*>(&pooled[0]);
            *reinterpret_cast<uint4*>(base + act_idx(page, p.pool_cp, p.h / 2, p.w / 2, o0, yA / 2, x / 2)) = s4[0];
            *reinterpret_cast<uint4*>(base + act_idx(page, p.pool_cp, p.h / 2, p.w / 2, o0 + 8, yA / 2, x / 2)) = s4[1];
        }
    }
}

// ---------------------------------------------------------------------------
// First layer of both FCN variants: Conv2D(20, 5x5, 'same', relu) on the uint8 page
// (model.py:50 / :211).  C_in = 1, so there is no contraction to hand to the tensor
// cores; instead every weight is an immediate constant-bank operand of an FFMA
// (fully unrolled 25 taps x 20 outputs), each thread owns 4 consecutive pixels and
// the stream is pure FFMA.  fp32 math on the exact (float)(v/255.0) inputs.
// ---------------------------------------------------------------------------
constexpr int C1_K = 5, C1_CO = 20, C1_PX = 4;
__constant__ float c_conv1_w[C1_K * C1_K * C1_CO];
__constant__ float c_conv1_b[C1_CO];
static int64_t g_conv1_owner[64] = {0};      // generation stamp of the model whose weights sit in the constant bank

template <typename T>
__global__ void __launch_bounds__(256) conv1_c20_kernel(const uint8_t* __restrict__ img, int img_h, int img_w, int h, int w,
                                                       T* __restrict__ out, int out_cp) {
    constexpr int TWX = 32 * C1_PX, THY = 8;
    __shared__ float s_in[THY + C1_K - 1][TWX + C1_K - 1 + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x0 = blockIdx.x * TWX, y0 = blockIdx.y * THY, page = blockIdx.z;
    const uint8_t* src = img + (size_t)page * img_h * img_w;
    for (int i = tid; i < (THY + 4) * (TWX + 4); i += 256) {
        const int px = i % (TWX + 4), py = i / (TWX + 4);
        const int gy = y0 + py - 2, gx = x0 + px - 2;
        float v = 0.f;
        if (gy >= 0 && gy < img_h && gx >= 0 && gx < img_w) v = c_u8_lut[src[(size_t)gy * img_w + gx]];
        s_in[py][px] = v;
    }
    __syncthreads();
    float acc[C1_PX][C1_CO];
#pragma unroll
    for (int q = 0; q < C1_PX; ++q)
#pragma unroll
        for (int o = 0; o < C1_CO; ++o) acc[q][o] = c_conv1_b[o];
#pragma unroll
    for (int dy = 0; dy < C1_K; ++dy) {
        float row[C1_PX + C1_K - 1];
#pragma unroll
        for (int i = 0; i < C1_PX + C1_K - 1; ++i) row[i] = s_in[ty + dy][tx * C1_PX + i];
#pragma unroll
        for (int dx = 0; dx < C1_K; ++dx)
#pragma unroll
            for (int o = 0; o < C1_CO; ++o) {
                const float wv = c_conv1_w[(dy * C1_K + dx) * C1_CO + o];
#pragma unroll
                for (int q = 0; q < C1_PX; ++q) acc[q][o] = fmaf(row[q + dx], wv, acc[q][o]);
            }
    }
    const int y = y0 + ty;
    if (y >= h) return;
#pragma unroll
    for (int q = 0; q < C1_PX; ++q) {
        const int x = x0 + tx * C1_PX + q;
        if (x >= w) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {            // 32 padded channels = 4 planes of 8
            T v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int o = g * 8 + e;
                v[e] = from_f<T>(o < C1_CO ? fmaxf(acc[q][o < C1_CO ? o : 0], 0.f) : 0.f);
            }
            *reinterpret_cast<uint4*>(out + act_idx(page, out_cp, h, w, g * 8, y, x)) = *reinterpret_cast<const uint4*>(v);
        }
    }
}

static int launch_conv1_c20(pcs_ctx* ctx, const DirectConvArgs& a) {
    if (ctx->device >= 64 || g_conv1_owner[ctx->device] != ctx->model_stamp) {
        // w32 is [25][1][20] == the constant-bank layout; device-to-device copy on the stream
        PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_conv1_w, a.w32, sizeof(float) * C1_K * C1_K * C1_CO, 0, cudaMemcpyDeviceToDevice, ctx->stream));
        PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_conv1_b, a.b32, sizeof(float) * C1_CO, 0, cudaMemcpyDeviceToDevice, ctx->stream));
        if (ctx->device < 64) g_conv1_owner[ctx->device] = ctx->model_stamp;
    }
    dim3 grid((a.w + 127) / 128, (a.h + 7) / 8, a.n), block(32, 8);
    if (ctx->precision == PCS_PREC_BF16)
        conv1_c20_kernel<__nv_bfloat16><<<grid, block, 0, ctx->stream>>>(reinterpret_cast<const uint8_t*>(a.src[0].p), a.img_h, a.img_w, a.h, a.w,
                                                                         reinterpret_cast<__nv_bfloat16*>(a.out), a.out_cp);
    else
        conv1_c20_kernel<__half><<<grid, block, 0, ctx->stream>>>(reinterpret_cast<const uint8_t*>(a.src[0].p), a.img_h, a.img_w, a.h, a.w,
                                                                  reinterpret_cast<__half*>(a.out), a.out_cp);
    PCS_LAUNCH_CHECK(ctx, "conv1_c20_kernel");
    return PCS_OK;
}

int launch_conv_direct(pcs_ctx* ctx, const DirectConvArgs& a) {
    if (a.k < 1 || a.k > KMAX) return set_err(ctx, PCS_ERR_ARG, "conv_direct: kernel size %d unsupported", a.k);
    if (a.out_cp % COT || (a.pool_out && a.pool_cp % COT))
        return set_err(ctx, PCS_ERR_ARG, "conv_direct: channel stride must be a multiple of %d", COT);
    if ((a.h & 1) || (a.w & 1)) return set_err(ctx, PCS_ERR_ARG, "conv_direct: odd grid %dx%d", a.h, a.w);
    PCS_TRY(ensure_lut(ctx));
    if (a.src_u8 && a.k == C1_K && a.cout == C1_CO && a.relu && a.out_cp == 32 && !a.pool_out && !a.upsample && a.fast_first)
        return launch_conv1_c20(ctx, a);
    DirectParams p{};
    p.s0 = a.src[0].p; p.c0 = a.src[0].c; p.cp0 = a.src[0].cp;
    p.s1 = a.nsrc > 1 ? a.src[1].p : nullptr; p.c1 = a.nsrc > 1 ? a.src[1].c : 0; p.cp1 = a.nsrc > 1 ? a.src[1].cp : 0;
    p.img_h = a.img_h; p.img_w = a.img_w; p.upsample = a.upsample;
    p.h = a.h; p.w = a.w; p.k = a.k; p.pad = a.pad;
    p.w32 = a.w32; p.b32 = a.b32; p.cin = a.cin; p.cout = a.cout; p.relu = a.relu;
    p.out = a.out; p.out_cp = a.out_cp; p.pool = a.pool_out; p.pool_cp = a.pool_cp;
    const int cp = a.out ? a.out_cp : a.pool_cp;
    p.cout_tiles = cp / COT;
    dim3 grid((a.w + TW - 1) / TW, (a.h + TH - 1) / TH, a.n * p.cout_tiles), block(32, 8);
    const bool bf = ctx->precision == PCS_PREC_BF16;
    if (a.src_u8) {
        if (bf) conv_direct_kernel<__nv_bfloat16, true><<<grid, block, 0, ctx->stream>>>(p);
        else conv_direct_kernel<__half, true><<<grid, block, 0, ctx->stream>>>(p);
    } else {
        if (bf) conv_direct_kernel<__nv_bfloat16, false><<<grid, block, 0, ctx->stream>>>(p);
        else conv_direct_kernel<__half, false><<<grid, block, 0, ctx->stream>>>(p);
    }
    PCS_LAUNCH_CHECK(ctx, "conv_direct_kernel");
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// Conv2DTranspose(2x2, stride 2, 'same'):  y[2h+i, 2w+j, o] = b[o] + sum_c x[h,w,c] K[i,j,o,c]
// weights arrive as [tap = i*2+j][cin][cout].  One thread per output pixel.
// ---------------------------------------------------------------------------
struct DeconvParams {
    const void* s0; const void* s1;
    int c0, cp0, c1, cp1;
    int h, w;                 // input grid
    const float* w32; const float* b32;
    int cin, cout, relu;
    void* out; int out_cp; int cout_tiles;
};

template <typename T>
__global__ void __launch_bounds__(256) deconv_s2_direct_kernel(DeconvParams p) {
    extern __shared__ __align__(16) float s_dw[];    // [4][cin][COT]
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int page = blockIdx.z / p.cout_tiles;
    const int o0 = (blockIdx.z % p.cout_tiles) * COT;
    for (int i = tid; i < 4 * p.cin * COT; i += 256) {
        const int o = i % COT, c = (i / COT) % p.cin, t = i / (COT * p.cin);
        s_dw[i] = (o0 + o < p.cout) ? __ldg(p.w32 + ((size_t)t * p.cin + c) * p.cout + o0 + o) : 0.f;
    }
    __syncthreads();
    const int X = blockIdx.x * 32 + tx, Y = blockIdx.y * 8 + ty;
    if (X >= 2 * p.w || Y >= 2 * p.h) return;
    const int t = (Y & 1) * 2 + (X & 1);
    float acc[COT];
#pragma unroll
    for (int o = 0; o < COT; ++o) acc[o] = 0.f;
    const T* a0 = reinterpret_cast<const T*>(p.s0);
    const T* a1 = reinterpret_cast<const T*>(p.s1);
    for (int c = 0; c < p.cin; ++c) {
        const float a = to_f(c < p.c0 ? a0[act_idx(page, p.cp0, p.h, p.w, c, Y >> 1, X >> 1)]
                                      : a1[act_idx(page, p.cp1, p.h, p.w, c - p.c0, Y >> 1, X >> 1)]);
        const float4* wv = reinterpret_cast<const float4*>(s_dw + ((size_t)t * p.cin + c) * COT);
#pragma unroll
        for (int q = 0; q < COT / 4; ++q) {
            const float4 w4 = wv[q];
            acc[4 * q + 0] = fmaf(a, w4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(a, w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a, w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(a, w4.w, acc[4 * q + 3]);
        }
    }
    T vals[COT];
#pragma unroll
    for (int o = 0; o < COT; ++o) {
        float v = 0.f;
        if (o0 + o < p.cout) {
            v = acc[o] + __ldg(p.b32 + o0 + o);
            if (p.relu) v = fmaxf(v, 0.f);
        }
        vals[o] = from_f<T>(v);
    }
    T* base = reinterpret_cast<T*>(p.out);
    const uint4* s4 = reinterpret_cast<const uint4*>(&vals[0]);
    *reinterpret_cast<uint4*>(base + act_idx(page, p.out_cp, 2 * p.h, 2 * p.w, o0, Y, X)) = s4[0];
    *reinterpret_cast<uint4*>(base + act_idx(page, p.out_cp, 2 * p.h, 2 * p.w, o0 + 8, Y, X)) = s4[1];
}

int launch_deconv_s2_direct(pcs_ctx* ctx, const DeconvS2Args& a) {
    if (a.out_cp % COT) return set_err(ctx, PCS_ERR_ARG, "deconv_s2: channel stride must be a multiple of %d", COT);
    DeconvParams p{};
    p.s0 = a.src[0].p; p.c0 = a.src[0].c; p.cp0 = a.src[0].cp;
    p.s1 = a.nsrc > 1 ? a.src[1].p : nullptr; p.c1 = a.nsrc > 1 ? a.src[1].c : 0; p.cp1 = a.nsrc > 1 ? a.src[1].cp : 0;
    p.h = a.h; p.w = a.w; p.w32 = a.w32; p.b32 = a.b32; p.cin = a.cin; p.cout = a.cout; p.relu = a.relu;
    p.out = a.out; p.out_cp = a.out_cp; p.cout_tiles = a.out_cp / COT;
    const size_t smem = (size_t)4 * a.cin * COT * sizeof(float);
    if (smem > 200 * 1024) return set_err(ctx, PCS_ERR_ARG, "deconv_s2: C_in %d too large", a.cin);
    dim3 grid((2 * a.w + 31) / 32, (2 * a.h + 7) / 8, a.n * p.cout_tiles), block(32, 8);
    if (ctx->precision == PCS_PREC_BF16) {
        if (smem > 48 * 1024)
            PCS_CUDA(ctx, cudaFuncSetAttribute(deconv_s2_direct_kernel<__nv_bfloat16>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        deconv_s2_direct_kernel<__nv_bfloat16><<<grid, block, smem, ctx->stream>>>(p);
    } else {
        if (smem > 48 * 1024)
            PCS_CUDA(ctx, cudaFuncSetAttribute(deconv_s2_direct_kernel<__half>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        deconv_s2_direct_kernel<__half><<<grid, block, smem, ctx->stream>>>(p);
    }
    PCS_LAUNCH_CHECK(ctx, "deconv_s2_direct_kernel");
    return PCS_OK;
}

}  // namespace pcs
