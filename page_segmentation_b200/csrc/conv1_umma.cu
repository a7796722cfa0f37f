// First layer of the network on the tensor cores:
//   FCN variants: Conv2D(20, 5x5, 'same', relu) over the uint8 page  (model.py:50 / :211, x/255 of architecture.py:67-68);
//   U-Net:        Conv2D(64, 3x3, 'same', relu)                      (model.py:156, conv1a) -- same kernel, template
//                 parameters <kernel size, padded C_out, rows per tile> = <5, 32, 8> / <3, 64, 4>.
//
// C_in = 1, so the contraction runs over the 25 taps.  The builder warps expand every input row rho
// of the tile once into shared memory as   E_rho[x'] = (in[rho][x0-2+x' + 0..7])  -- one 16-byte unit
// (8 operand elements) per pixel, i.e. exactly one K-major SWIZZLE_NONE "plane" of the canonical UMMA
// layout.  Output row r then needs K = (dy, dx-slot): its A operand for the K=16 step ks is the pair of
// expanded rows (r+2ks, r+2ks+1) -- the same smem rows serve five different output rows -- and the B
// operand holds W[dy][dx] in slot dx < 5 and zeros elsewhere.  Pixels 0..255 are exact in bf16/fp16; the
// fp32 weights are split into hi + lo operand halves (two MMAs, ~16 mantissa bits; bf16 operands) or used as one fp16
// operand (11 bits, like the weights of every other layer) and 1/255 is applied to the fp32 accumulator.
//
// Warp roles (320 threads): warp 0 loads the weight image; warp 1 = MMA issuer + TMEM owner;
// warps 2-5 and 10-13 = epilogue (bias, ReLU, pack, plane-major stores; two warps per TMEM lane quarter
// splitting the rows); warps 6-9 = builders.
#include "common.cuh"
#include "umma_ptx.cuh"

namespace pcs {
namespace {
using namespace ptx;

// KSZ = kernel size (5 | 3), NP = padded C_out (32 | 64), R = output rows per tile (2 accumulator stages of R * NP TMEM columns)
template <int KSZ, int NP, int R> struct C1Cfg {
    static constexpr int KS = (KSZ + 1) / 2;                  // K = 16 steps: vertical tap pairs (2 ks, 2 ks + 1)
    static constexpr int ROWS = R + 2 * KS - 1;               // expanded rows per stage (r + 2 ks + 1 <= R - 1 + 2 KS - 1)
    static constexpr int STAGE_BYTES = ROWS * 128 * 16;
    static constexpr int B_BYTES = 2 * KS * 2 * NP * 16;      // [hi|lo][ks][plane][n][8]
    static constexpr int CVT_COPY = ROWS * 144 + 14;          // elements per converted copy: an odd number of words (943 for 13 rows, 511
                                                              // for 7), so the windows of odd pixel slots (copy 1, word m+1) use other banks
    static_assert(2 * R * NP <= 512, "two accumulator stages must fit TMEM");
};
constexpr int C1_SW = 124;         // valid output pixels per 128-pixel strip
constexpr int C1_STAGES = 3;
constexpr int C1_ROW_BYTES = 128 * 16;
constexpr int C1_THREADS = 448;     // warp 0 weights, warp 1 MMA, warps 2-5 + 10-13 epilogue (2 per lane quarter), warps 6-9 builders
constexpr int C1_CVT_W = 144;      // operand elements per converted patch row (4-byte aligned rows)
constexpr int C1_MAXN = 64;

struct Conv1Params {
    const uint8_t* img; int img_h, img_w;   // real page
    int n, h, w;                            // padded grid
    const uint8_t* wimg;                    // C1_B_BYTES operand image
    float bias[C1_MAXN];                    // by value: read as constant-bank FFMA operands, no shared-memory traffic
    void* out; int out_cp;
    void* pair_out;                         // 5x5 / 20 channels only: [n][h][w + 1][8] pixel-pair units of channels 16..19 (see the epilogue), or null
    int strips, rowblocks, num_tiles;
    int halves;                             // operand halves of the weights: 2 = hi + lo (bf16 operands: ~16 mantissa bits), 1 = hi only
                                            // (fp16 operands: 11 bits, what every other layer's weights have)
    int dbg;                                // diagnosis only (PCSEG_C1_DEBUG): 1 = no output stores, 2 = hi MMAs only, 4 = no expansion
};

template <typename T, int KSZ, int NP, int C1_R>
__global__ void __launch_bounds__(C1_THREADS, 1) conv1_umma_kernel(const Conv1Params p) {
    using Cfg = C1Cfg<KSZ, NP, C1_R>;
    constexpr int C1_N = NP, C1_ROWS = Cfg::ROWS, C1_STAGE_BYTES = Cfg::STAGE_BYTES, C1_B_BYTES = Cfg::B_BYTES, C1_CVT_COPY = Cfg::CVT_COPY;
    constexpr int KS = Cfg::KS, PAD = KSZ / 2;
    constexpr uint32_t IDESC = (1u << 4) | ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 7) |
                               ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 10) |
                               ((uint32_t)(C1_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[C1_STAGES], s_empty[C1_STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    // every patch pixel converted ONCE to the operand type, stored twice: copy c holds element e at index e + c, so
    // that the 8-element window of an odd pixel slot is 4-byte aligned in copy 1 (and of an even slot in copy 0)
    __shared__ __align__(16) T s_cvt[C1_STAGES][2 * C1_CVT_COPY];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_b = base;                                   // weight image
    uint8_t* stages = base + ((C1_B_BYTES + 1023) / 1024) * 1024;

    if (warp == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(p.wimg);
        uint4* dst = reinterpret_cast<uint4*>(s_b);
        for (int i = lane; i < C1_B_BYTES / 16; i += 32) dst[i] = __ldg(src + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) {
            for (int s = 0; s < C1_STAGES; ++s) { mbar_init(&s_full[s], 4); mbar_init(&s_empty[s], 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(&s_tfull[a], 1); mbar_init(&s_tempty[a], 8); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    griddep_launch();
    griddep_wait();
    const uint32_t tmem_base = s_tmem_base;
    const int tiles_per_page = p.strips * p.rowblocks;

    if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const bool leader = elect_one();
        const uint32_t hi = (uint32_t)(make_desc(0, 0, 128) >> 32);
        constexpr uint32_t a_lbo = ((uint32_t)(C1_ROW_BYTES >> 4) & 0x3fffu) << 16;       // K halves = consecutive rows
        constexpr uint32_t b_lbo = (((uint32_t)C1_N * 16u >> 4) & 0x3fffu) << 16;
        const uint32_t b_lo0 = ((smem_u32(s_b) >> 4) & 0x3fffu) | b_lbo;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            mbar_wait(&s_tempty[acc], acc_phase ^ 1u);
            mbar_wait(&s_full[stage], phase);
            tc_fence_after();
            if (leader) {
                const uint32_t a_lo0 = ((smem_u32(stages + (size_t)stage * C1_STAGE_BYTES) >> 4) & 0x3fffu) | a_lbo;
                const uint32_t d0 = tmem_base + (uint32_t)(acc * C1_R * C1_N);
#pragma unroll
                for (int r = 0; r < C1_R; ++r) {
#pragma unroll
                    for (int half = 0; half < ((p.dbg & 2) ? 1 : p.halves); ++half) {        // weights hi, then lo
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint32_t a_off = (uint32_t)((r + 2 * ks) * (C1_ROW_BYTES >> 4));
                            const uint32_t b_off = (uint32_t)((half * KS + ks) * (2 * C1_N));   // 16-byte units
                            tc_mma(d0 + (uint32_t)(r * C1_N), a_lo0 + a_off, hi, b_lo0 + b_off, hi, IDESC, (half | ks) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(&s_empty[stage]);
                tc_commit(&s_tfull[acc]);
            }
            __syncwarp();
            if (++stage == C1_STAGES) { stage = 0; phase ^= 1u; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if ((warp >= 2 && warp < 6) || warp >= 10) {
        // ===================== epilogue =====================
        const int quarter = warp & 3, rgroup = warp >= 10 ? 1 : 0;
        const int nplanes = p.out_cp >> 3;                 // 3 (tensor engine: 20 -> 24 channels) or 4
        int acc = 0;
        uint32_t acc_phase = 0;
        T* out = reinterpret_cast<T*>(p.out);
        const float inv255 = 1.0f / 255.0f;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int page = tile / tiles_per_page;
            const int rem = tile - page * tiles_per_page;
            const int rb = rem % p.rowblocks, strip = rem / p.rowblocks;
            const int m = quarter * 32 + lane;
            const int x = strip * C1_SW + m;
            const bool xok = m < C1_SW && x < p.w;
            const int y0 = rb * C1_R;
            mbar_wait(&s_tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * C1_R * C1_N);
#pragma unroll 1
            for (int r = rgroup; r < C1_R; r += 2) {
                const int y = y0 + r;
                const bool store = xok && y < p.h && !(p.dbg & 1);
                uint32_t v[C1_N];
#pragma unroll
                for (int c16 = 0; c16 < C1_N / 16; ++c16)
                    tmem_ld16(t_lane + (uint32_t)(r * C1_N + c16 * 16), *reinterpret_cast<uint32_t(*)[16]>(&v[c16 * 16]));
                tmem_ld_wait();
                if (!store) continue;
#pragma unroll
                for (int g = 0; g < C1_N / 8; ++g) {
                    if (g >= nplanes) break;
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = fmaxf(fmaf(__uint_as_float(v[g * 8 + i]), inv255, p.bias[g * 8 + i]), 0.f);
                    *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, g * 8, y, x)) =
                        make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7]));
                }
                if constexpr (KSZ == 5) {
                    if (p.pair_out) {
                        // Channels 16..19 travel as pixel-pair units: unit i of a row holds [ch 16..19 of pixel i - 1 | of pixel i],
                        // i = 0 .. w (w + 1 units: the first and the last are half padding), so that one 16-byte K half of conv2's
                        // MMA covers TWO horizontal taps of these four channels (conv_fold.cu, PX).  Columns 20..23 of this layer's
                        // GEMM are channels 16..19 of the pixel to the RIGHT (weights shifted by one tap), so a lane builds the unit
                        // x + 1 from its own accumulators; lane x = 0 also writes unit 0.
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = fmaxf(fmaf(__uint_as_float(v[16 + i]), inv255, p.bias[16 + i]), 0.f);
                        if (x + 1 >= p.w) { f[4] = 0.f; f[5] = 0.f; f[6] = 0.f; f[7] = 0.f; }      // 'same' padding of conv2's input
                        T* unit = reinterpret_cast<T*>(p.pair_out) + (((size_t)page * p.h + y) * (p.w + 1) + x + 1) * 8;
                        *reinterpret_cast<uint4*>(unit) =
                            make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7]));
                        if (x == 0) *reinterpret_cast<uint4*>(unit - 8) = make_uint4(0u, 0u, pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (warp >= 6 && warp < 10) {
        // ===================== builders: expand input rows into K-major planes =====================
        // phase 1: the 13 x 135-byte uint8 patch is fetched with independent, coalesced byte loads -- for the
        // NEXT tile, so that the global-load latency hides behind the current tile's work; phase 2: each byte is
        // converted once (pixels 0..255 are exact in bf16/fp16) into the two shifted copies; phase 3: every
        // thread assembles the 16-byte unit (8 consecutive pixels) of "its" pixel slot for all 13 rows with four
        // aligned 32-bit loads and one 128-bit store per row.
        const int xq = threadIdx.x - 192;                 // 0..127 = pixel slot of the strip patch
        int stage = 0;
        uint32_t phase = 0;
        // byte (row j, column xq) for the 13 rows, plus one of the 13 x 8 look-ahead bytes (columns 128..135)
        constexpr int NV = C1_ROWS + 1;
        uint8_t next[NV];
        const int xrow = xq >> 3, xcol = 128 + (xq & 7);  // the look-ahead byte of this thread (xq < 104)
        auto fetch = [&](int tile, uint8_t (&dst)[NV]) {
            const int page = tile / tiles_per_page;
            const int rem = tile - page * tiles_per_page;
            const int rb = rem % p.rowblocks, strip = rem / p.rowblocks;
            const int gxb = strip * C1_SW - PAD, gy0 = rb * C1_R - PAD;
            const uint8_t* src = p.img + (size_t)page * p.img_h * p.img_w;
            const int gx = gxb + xq;
            const bool xin = gx >= 0 && gx < p.img_w;
            const uint8_t* col = src + (ptrdiff_t)gy0 * p.img_w + gx;
#pragma unroll
            for (int j = 0; j < C1_ROWS; ++j) {
                const int gy = gy0 + j;
                dst[j] = (xin && gy >= 0 && gy < p.img_h) ? __ldg(col + (ptrdiff_t)j * p.img_w) : (uint8_t)0;
            }
            const int gy = gy0 + xrow, gx2 = gxb + xcol;
            dst[C1_ROWS] = (xrow < C1_ROWS && gy >= 0 && gy < p.img_h && gx2 < p.img_w)
                               ? __ldg(src + (size_t)gy * p.img_w + gx2) : (uint8_t)0;
        };
        if ((int)blockIdx.x < p.num_tiles) fetch(blockIdx.x, next);
        const int par = xq & 1;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            mbar_wait(&s_empty[stage], phase ^ 1u);           // stage (incl. its conversion buffer) is free
            T* cvt = s_cvt[stage];
#pragma unroll
            for (int j = 0; j < C1_ROWS; ++j) {
                const T v = T((float)next[j]);
                cvt[j * C1_CVT_W + xq] = v;
                cvt[C1_CVT_COPY + j * C1_CVT_W + xq + 1] = v;
            }
            if (xrow < C1_ROWS) {
                const T v = T((float)next[C1_ROWS]);
                cvt[xrow * C1_CVT_W + xcol] = v;
                cvt[C1_CVT_COPY + xrow * C1_CVT_W + xcol + 1] = v;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tile + (int)gridDim.x < p.num_tiles) fetch(tile + gridDim.x, next);
            uint8_t* dst = stages + (size_t)stage * C1_STAGE_BYTES + (size_t)xq * 16;
            const uint32_t* win = reinterpret_cast<const uint32_t*>(cvt + par * C1_CVT_COPY) + ((xq + par) >> 1);
#pragma unroll
            for (int row = 0; row < ((p.dbg & 4) ? 1 : C1_ROWS); ++row) {
                const uint32_t* w4 = win + row * (C1_CVT_W / 2);
                *reinterpret_cast<uint4*>(dst + (size_t)row * C1_ROW_BYTES) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[stage]);                      // one arrival per builder warp
            // (the conversion buffer is reused three tiles later, after s_empty: every builder has arrived, hence read it)
            if (++stage == C1_STAGES) { stage = 0; phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

// Operand image [hi|lo][ks][plane 0..1][n][e 0..7]: value W[dy = 2ks+plane][dx = e][n] for dy, e < ksz (ksz = 5: 3 K steps,
// n < 32; ksz = 3: 2 K steps, n < 64).
// pairx (5x5, 20 channels): columns 20..23 = channels 16..19 of the pixel one to the right, W[dy][dx = e - 1][16 + n - 20].
size_t conv1_umma_weight_image(const float* w32 /*[ksz*ksz][1][cout]*/, int ksz, int cout, int precision, std::vector<uint16_t>& out, bool pairx) {
    const int KS = (ksz + 1) / 2, NP = ksz == 5 ? 32 : 64;
    out.assign((size_t)2 * KS * 2 * NP * 8, 0);
    auto to16 = [&](float v) -> uint16_t {
        if (precision == PCS_PREC_BF16) { __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b); }
        __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h);
    };
    auto from16 = [&](uint16_t u) -> float {
        if (precision == PCS_PREC_BF16) { uint32_t x = (uint32_t)u << 16; float f; memcpy(&f, &x, 4); return f; }
        __half_raw hr; hr.x = u; return __half2float(__half(hr));
    };
    for (int ks = 0; ks < KS; ++ks)
        for (int pl = 0; pl < 2; ++pl)
            for (int n = 0; n < NP; ++n)
                for (int e = 0; e < 8; ++e) {
                    const int dy = 2 * ks + pl;
                    if (dy >= ksz) continue;
                    float w;
                    if (n < cout) {
                        if (e >= ksz) continue;
                        w = w32[(size_t)(dy * ksz + e) * cout + n];
                    } else if (pairx && n < cout + 4 && e >= 1 && e <= ksz) {
                        w = w32[(size_t)(dy * ksz + e - 1) * cout + (n - 4)];
                    } else continue;
                    const uint16_t hi = to16(w);
                    const uint16_t lo = to16(w - from16(hi));
                    const size_t idx = (((size_t)ks * 2 + pl) * NP + n) * 8 + e;
                    out[idx] = hi;
                    out[(size_t)KS * 2 * NP * 8 + idx] = lo;
                }
    return out.size() * sizeof(uint16_t);
}

namespace {
template <typename T, int KSZ, int NP, int R>
int launch_conv1_t(pcs_ctx* ctx, Conv1Params& p) {
    using Cfg = C1Cfg<KSZ, NP, R>;
    p.strips = (p.w + C1_SW - 1) / C1_SW;
    p.rowblocks = (p.h + R - 1) / R;
    p.num_tiles = p.n * p.strips * p.rowblocks;
    const size_t smem = std::max<size_t>(((Cfg::B_BYTES + 1023) / 1024) * 1024 + (size_t)C1_STAGES * Cfg::STAGE_BYTES + 1024, kSoloSmem);
    const int grid = std::min(p.num_tiles, ctx->sm_count);
    static bool set[64] = {};                // the attribute is per device (and per instantiation: this is a template)
    if (ctx->device >= 64 || !set[ctx->device]) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(conv1_umma_kernel<T, KSZ, NP, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (ctx->device < 64) set[ctx->device] = true;
    }
    PCS_CUDA(ctx, launch_kernel_pdl(conv1_umma_kernel<T, KSZ, NP, R>, dim3(grid), dim3(C1_THREADS), smem, ctx->stream, ctx->pdl, p));
    PCS_LAUNCH_CHECK(ctx, "conv1_umma_kernel");
    return PCS_OK;
}
}  // namespace

bool conv1_umma_supported(int ksz, int cout) { return (ksz == 5 && cout <= 32) || (ksz == 3 && cout <= 64); }

int launch_conv1_umma(pcs_ctx* ctx, const uint8_t* d_image, int n, int img_h, int img_w, int h, int w, const void* wimg,
                      const float* h_bias /*host, cout values*/, int ksz, int cout, void* out, int out_cp, void* pair_out) {
    if (!conv1_umma_supported(ksz, cout)) return set_err(ctx, PCS_ERR_ARG, "conv1_umma: no instantiation for a %dx%d kernel with %d outputs", ksz, ksz, cout);
    if (pair_out && (ksz != 5 || cout != 20 || out_cp != 16))
        return set_err(ctx, PCS_ERR_ARG, "conv1_umma: the pixel-pair hand-off is for 5x5 / 20 channels with two whole planes");
    if ((out_cp & 7) || (!pair_out && out_cp < cout) || out_cp > (ksz == 5 ? 32 : 64))
        return set_err(ctx, PCS_ERR_ARG, "conv1_umma: output stride of %d channels", out_cp);
    Conv1Params p{};
    p.img = d_image; p.img_h = img_h; p.img_w = img_w; p.n = n; p.h = h; p.w = w;
    p.wimg = reinterpret_cast<const uint8_t*>(wimg); p.out = out; p.out_cp = out_cp;
    for (int i = 0; i < C1_MAXN; ++i) p.bias[i] = i < cout ? h_bias[i] : 0.f;
    p.pair_out = pair_out;
    if (pair_out) for (int i = 0; i < 4; ++i) p.bias[20 + i] = h_bias[16 + i];
    { const char* e = getenv("PCSEG_C1_DEBUG"); p.dbg = e ? atoi(e) : 0; }
    const bool bf = ctx->precision == PCS_PREC_BF16;
    // fp16 operands carry 11 mantissa bits: the weights of this layer are used like those of every other layer (one operand).
    // With bf16 operands (8 bits) the hi + lo split keeps the first layer, whose inputs are exact, out of the error budget.
    // PCSEG_C1_SPLIT=1 / 0 forces either (measured on A4 pages, fp16: class-map agreement with the fp64 oracle 99.982 % with
    // the split, 99.980 % without; conv1 1.02 -> 0.79 ms per 64 pages).
    static const int split_env = [] { const char* e = getenv("PCSEG_C1_SPLIT"); return e ? atoi(e) : -1; }();
    p.halves = split_env >= 0 ? (split_env ? 2 : 1) : (bf ? 2 : 1);
    if (ksz == 5) return bf ? launch_conv1_t<__nv_bfloat16, 5, 32, 8>(ctx, p) : launch_conv1_t<__half, 5, 32, 8>(ctx, p);
    return bf ? launch_conv1_t<__nv_bfloat16, 3, 64, 4>(ctx, p) : launch_conv1_t<__half, 3, 64, 4>(ctx, p);
}

}  // namespace pcs
