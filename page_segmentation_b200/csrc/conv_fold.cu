// "Marching, dx-folded" implicit-GEMM 5x5 'same' convolution for the small-channel encoder layers
// (conv2, conv3, conv4 of ocr4all_pixel_classifier/lib/model.py:52-57 / :212-215).
//
// With C_out of 30-40 the plain kernel (conv_umma.cu) issues one N=32..48 MMA per tap and re-reads the
// 4 KB A operand from shared memory for each of them: it is shared-memory-read bound at ~40 % tensor
// activity.  Here the five horizontal taps are folded into the N dimension:
//     D_y[j][dx*NPAD + o] = sum_{dy, c} in[y+dy-2][x0-2+j][c] * W[dy][dx][c][o]        (one N' = 5*NPAD MMA per (chunk, dy))
//     out[y][x0+t][o]     = b[o] + sum_dx D_y[t+dx][dx*NPAD + o]                       (shifted sum in the epilogue)
// so A is read once per 5 taps and the MMA (N' = 160 / 240) is tensor-bound.  Further:
//   * all weights of the layer stay resident in shared memory (51-115 KB), loaded once per CTA;
//   * a CTA marches down a 124-pixel strip: input rows stream through a ring (one TMA box per row holding
//     all channel chunks), every output row needs ONE new input row (halo re-read 132/128 instead of 12/8);
//   * two TMEM accumulator stages (2 x N' columns) let the epilogue of row y overlap the MMAs of row y+1;
//   * 8 epilogue warps in two groups; a group owns a row pair (keeps the even row for the fused 2x2
//     max-pool); the dx shift is a warp shuffle plus a 10-value-per-warp exchange through shared memory.
#include "common.cuh"
#include "umma_ptx.cuh"

namespace pcs {
namespace {
using namespace ptx;

template <typename T> __device__ __forceinline__ float2 unpack2f(uint32_t v);
template <> __device__ __forceinline__ float2 unpack2f<__nv_bfloat16>(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2f<__half>(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

constexpr int F_THREADS = 64 + 8 * 32;      // warp 0 producer, warp 1 MMA, warps 2-9 epilogue (2 groups x 4 quarters)
constexpr int F_RING = 8;                   // input-row ring entries
constexpr int F_SW = 124;                   // valid output pixels per strip

struct FoldParams {
    int n, h, w;
    int seg_rows, segs, strips, num_items;  // work item = (page, strip, segment of seg_rows output rows)
    const uint8_t* wimg;                    // [chunk][dy][plane][N' rows][8]: resident operand image
    const float* bias;
    int cout, relu;
    void* out; int out_cp;
    void* pool; int pool_cp;
    uint32_t w_bytes;
};

template <typename T, int NPAD, int NCH>
__global__ void __launch_bounds__(F_THREADS, 1)
conv_fold_kernel(const FoldParams p, const __grid_constant__ CUtensorMap tm) {
    constexpr int NF = 5 * NPAD;                                     // folded N
    constexpr uint32_t ROW_BYTES = NCH * 2 * 2048;                   // one ring entry: all chunks of one input row
    constexpr uint32_t WDY_BYTES = 2 * NF * 16;                      // weights of one (chunk, dy)
    constexpr uint32_t IDESC = (1u << 4) | ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 7) |
                               ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 10) |
                               ((uint32_t)(NF >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    static_assert(NF <= 256 && NF % 16 == 0, "folded N must be a legal UMMA N");
    static_assert(2 * NF <= 512, "two accumulator stages must fit TMEM");

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[F_RING], s_empty[F_RING], s_wfull, s_tfull[2][2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_bias[NPAD];
    __shared__ __align__(16) float s_xchg[2][2][4][10][16];          // [group][parity][quarter][(lane,dx) slot][16 ch]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = base;
    uint8_t* ring = base + ((p.w_bytes + 1023) / 1024) * 1024;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < F_RING; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        mbar_init(&s_wfull, 1);
        for (int g = 0; g < 2; ++g)
            for (int a = 0; a < 2; ++a) mbar_init(&s_tfull[g][a], 1);
        for (int a = 0; a < 2; ++a) mbar_init(&s_tempty[a], 4);       // TMEM stage drained (whichever group read it)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < NPAD; i += blockDim.x) s_bias[i] = i < p.cout ? __ldg(p.bias + i) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const int items_per_page = p.strips * p.segs;

    if (warp == 0) {
        // ===================== producer: resident weights, then the input-row stream =====================
        if (lane == 0) {
            mbar_expect_tx(&s_wfull, p.w_bytes);
            for (uint32_t off = 0; off < p.w_bytes; off += WDY_BYTES) bulk_load(s_w + off, p.wimg + off, WDY_BYTES, &s_wfull);
            uint32_t k = 0;                                           // running ring position
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                const int page = item / items_per_page;
                const int rem = item - page * items_per_page;
                const int seg = rem % p.segs, strip = rem / p.segs;
                const int ys = seg * p.seg_rows;
                const int rows = min(p.seg_rows, p.h - ys);
                const int x0 = strip * F_SW - 2;
                for (int i = 0; i < rows + 4; ++i, ++k) {
                    const uint32_t slot = k % F_RING, pass = k / F_RING;
                    mbar_wait(&s_empty[slot], (pass & 1u) ^ 1u);
                    mbar_expect_tx(&s_full[slot], ROW_BYTES);
                    // box = 256 u64 (128 px x 16 B) x 1 row x (2 * NCH) planes
                    tma_load_4d(ring + (size_t)slot * ROW_BYTES, &tm, &s_full[slot], x0 * 2, ys - 2 + i, 0, page);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint32_t hi = (uint32_t)(make_desc(0, 0, 128) >> 32);
        constexpr uint32_t a_lbo = ((2048u >> 4) & 0x3fffu) << 16;                    // the two K planes of a chunk
        constexpr uint32_t b_lbo = (((uint32_t)NF * 16u >> 4) & 0x3fffu) << 16;
        mbar_wait(&s_wfull, 0);
        const uint32_t ring_lo = (smem_u32(ring) >> 4) & 0x3fffu;
        const uint32_t b_lo0 = ((smem_u32(s_w) >> 4) & 0x3fffu) | b_lbo;
        uint32_t k = 0;                                                               // ring position of input row 0 of the item
        uint32_t use[2] = {0, 0};                                                     // uses of TMEM stage 0 / 1
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            const int page = item / items_per_page;
            const int rem = item - page * items_per_page;
            const int seg = rem % p.segs;
            const int ys = seg * p.seg_rows;
            const int rows = min(p.seg_rows, p.h - ys);
            (void)page;
            for (int i = 0; i < 4; ++i) {                                            // halo rows of the segment
                const uint32_t kk = k + i;
                mbar_wait(&s_full[kk % F_RING], (kk / F_RING) & 1u);
            }
            for (int r = 0; r < rows; ++r) {
                const int g = (r >> 1) & 1, st = r & 1;
                const uint32_t kn = k + r + 4;                                       // newest input row needed
                mbar_wait(&s_full[kn % F_RING], (kn / F_RING) & 1u);
                mbar_wait(&s_tempty[st], (use[st] & 1u) ^ 1u);
                tc_fence_after();
                if (leader) {
                    const uint32_t d = tmem_base + (uint32_t)(st * NF);
#pragma unroll
                    for (int dy = 0; dy < 5; ++dy) {
                        const uint32_t slot = (k + r + dy) % F_RING;
                        const uint32_t a_row = ring_lo + slot * (ROW_BYTES >> 4);
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            const uint32_t a_lo = (a_row + (uint32_t)c * (4096u >> 4)) | a_lbo;
                            const uint32_t b_lo = b_lo0 + (uint32_t)((c * 5 + dy) * (WDY_BYTES >> 4));
                            tc_mma(d, a_lo, hi, b_lo, hi, IDESC, (dy | c) ? 1u : 0u);
                        }
                    }
                    tc_commit(&s_empty[(k + r) % F_RING]);                            // input row r of the segment is dead
                    tc_commit(&s_tfull[g][st]);
                    if (r == rows - 1)
                        for (int i = 1; i <= 4; ++i) tc_commit(&s_empty[(k + r + i) % F_RING]);
                }
                __syncwarp();
                ++use[st];
            }
            k += rows + 4;
        }
    } else {
        // ===================== epilogue: shifted sum over dx, bias, activation, store, pool =====================
        const int quarter = warp & 3, group = (warp - 2) >> 2;
        const int j = quarter * 32 + lane;                           // patch column; output pixel t = j
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t bar_id = 2 + group;
        T* out = reinterpret_cast<T*>(p.out);
        T* pool = reinterpret_cast<T*>(p.pool);
        uint32_t cnt[2] = {0, 0};                                     // uses of my group's accumulator stages
        uint32_t xpar = 0;                                            // exchange-buffer parity
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            const int page = item / items_per_page;
            const int rem = item - page * items_per_page;
            const int seg = rem % p.segs, strip = rem / p.segs;
            const int ys = seg * p.seg_rows;
            const int rows = min(p.seg_rows, p.h - ys);
            const int x = strip * F_SW + j;
            const bool xok = j < F_SW && x < p.w;
            for (int r0 = 2 * group; r0 < rows; r0 += 4) {           // my group's row pairs
                uint32_t kept[NPAD / 2];                               // even row, packed, for the 2x2 max-pool
#pragma unroll
                for (int st = 0; st < 2; ++st) {
                    const int y = ys + r0 + st;
                    mbar_wait(&s_tfull[group][st], cnt[st] & 1u);
                    tc_fence_after();
                    const uint32_t tacc = t_lane + (uint32_t)(st * NF);
#pragma unroll
                    for (int hb = 0; hb < NPAD / 16; ++hb) {
                        float accv[16];
                        uint32_t v[16];
                        tmem_ld16(tacc + (uint32_t)(hb * 16), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) accv[i] = __uint_as_float(v[i]);
                        uint32_t vd[4][16];
#pragma unroll
                        for (int dx = 1; dx <= 4; ++dx) tmem_ld16(tacc + (uint32_t)(dx * NPAD + hb * 16), vd[dx - 1]);
                        tmem_ld_wait();
                        // publish what the previous quarter's top lanes need: lane l < dx publishes (l, dx)
#pragma unroll
                        for (int dx = 1; dx <= 4; ++dx) {
                            if (lane < dx) {
                                const int slot = (dx * (dx - 1)) / 2 + lane;          // 0 | 1,2 | 3,4,5 | 6,7,8,9
                                float4* dst = reinterpret_cast<float4*>(&s_xchg[group][xpar][quarter][slot][0]);
                                dst[0] = make_float4(__uint_as_float(vd[dx - 1][0]), __uint_as_float(vd[dx - 1][1]), __uint_as_float(vd[dx - 1][2]), __uint_as_float(vd[dx - 1][3]));
                                dst[1] = make_float4(__uint_as_float(vd[dx - 1][4]), __uint_as_float(vd[dx - 1][5]), __uint_as_float(vd[dx - 1][6]), __uint_as_float(vd[dx - 1][7]));
                                dst[2] = make_float4(__uint_as_float(vd[dx - 1][8]), __uint_as_float(vd[dx - 1][9]), __uint_as_float(vd[dx - 1][10]), __uint_as_float(vd[dx - 1][11]));
                                dst[3] = make_float4(__uint_as_float(vd[dx - 1][12]), __uint_as_float(vd[dx - 1][13]), __uint_as_float(vd[dx - 1][14]), __uint_as_float(vd[dx - 1][15]));
                            }
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
                        for (int dx = 1; dx <= 4; ++dx) {
                            const int src_lane = lane + dx;
                            float wv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) wv[i] = __uint_as_float(__shfl_down_sync(0xffffffffu, vd[dx - 1][i], dx));   // all lanes
                            if (src_lane >= 32) {
                                // pixels t+dx of the next TMEM lane quarter: published through shared memory
                                const int l2 = src_lane - 32;
                                const float4* srcp = reinterpret_cast<const float4*>(&s_xchg[group][xpar][(quarter + 1) & 3][(dx * (dx - 1)) / 2 + l2][0]);
#pragma unroll
                                for (int q4 = 0; q4 < 4; ++q4) {
                                    const float4 f = srcp[q4];
                                    wv[4 * q4 + 0] = f.x; wv[4 * q4 + 1] = f.y; wv[4 * q4 + 2] = f.z; wv[4 * q4 + 3] = f.w;
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) accv[i] += wv[i];      // quarter 3, lanes >= 28: garbage, never stored (j >= 124)
                        }
                        xpar ^= 1u;
                        // bias, activation, rounding
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float a = accv[2 * i] + s_bias[hb * 16 + 2 * i], b = accv[2 * i + 1] + s_bias[hb * 16 + 2 * i + 1];
                            if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                            pk[i] = pack2<T>(a, b);
                        }
                        if (out && xok && y < p.h) {
                            *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, hb * 16, y, x)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, hb * 16 + 8, y, x)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                        if (pool) {
                            if (st == 0) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) kept[hb * 8 + i] = pk[i];
                            } else {
                                uint32_t pm[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float2 a = unpack2f<T>(pk[i]), b = unpack2f<T>(kept[hb * 8 + i]);
                                    float mx = fmaxf(a.x, b.x), my = fmaxf(a.y, b.y);
                                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                                    my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, 1));
                                    pm[i] = pack2<T>(mx, my);
                                }
                                if (!(lane & 1) && xok && y < p.h) {
                                    const int ph = p.h >> 1, pw = p.w >> 1;
                                    *reinterpret_cast<uint4*>(pool + act_idx(page, p.pool_cp, ph, pw, hb * 16, y >> 1, x >> 1)) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
                                    *reinterpret_cast<uint4*>(pool + act_idx(page, p.pool_cp, ph, pw, hb * 16 + 8, y >> 1, x >> 1)) = make_uint4(pm[4], pm[5], pm[6], pm[7]);
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_tempty[st]);
                    ++cnt[st];
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn fold_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
    }
    return fn;
}

template <typename T, int NPAD, int NCH>
int launch_fold_t(pcs_ctx* ctx, const FoldConvArgs& a) {
    constexpr int NF = 5 * NPAD;
    FoldParams p{};
    p.n = a.n; p.h = a.h; p.w = a.w;
    p.wimg = reinterpret_cast<const uint8_t*>(a.wimg); p.bias = a.b32; p.cout = a.cout; p.relu = a.relu;
    p.out = a.out; p.out_cp = a.out_cp; p.pool = a.pool_out; p.pool_cp = a.pool_cp;
    p.w_bytes = (uint32_t)NCH * 5 * 2 * NF * 16;
    p.strips = (a.w + F_SW - 1) / F_SW;
    // segments: multiples of 4 rows (row pairs x 2 groups); aim at >= 4 items per CTA
    int segs = 1;
    while (segs < 64 && (size_t)a.n * p.strips * segs < (size_t)4 * ctx->sm_count && a.h / (segs * 2) >= 16) segs *= 2;
    p.seg_rows = ((a.h + segs - 1) / segs + 3) / 4 * 4;
    p.segs = (a.h + p.seg_rows - 1) / p.seg_rows;
    p.num_items = a.n * p.strips * p.segs;
    if ((a.h & 3) || (a.w & 1)) return set_err(ctx, PCS_ERR_ARG, "conv_fold: grid %dx%d must be a multiple of 4 x 2", a.h, a.w);
    EncodeTiledFn enc = fold_get_encode();
    if (!enc) return set_err(ctx, PCS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t planes = (cuuint64_t)a.src.cp / 8;
    if ((int)planes != 2 * NCH) return set_err(ctx, PCS_ERR_ARG, "conv_fold: source has %d planes, kernel expects %d", (int)planes, 2 * NCH);
    const cuuint64_t dims[4] = {(cuuint64_t)a.w * 2, (cuuint64_t)a.h, planes, (cuuint64_t)a.n};
    const cuuint64_t strides[3] = {(cuuint64_t)a.w * 16, (cuuint64_t)a.h * a.w * 16, planes * a.h * a.w * 16};
    const cuuint32_t box[4] = {256, 1, (cuuint32_t)(2 * NCH), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a.src.p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(ctx, PCS_ERR_CUDA, "conv_fold: cuTensorMapEncodeTiled failed with %d", (int)r);
    const size_t smem = ((p.w_bytes + 1023) / 1024) * 1024 + (size_t)F_RING * NCH * 4096 + 1024;
    if (smem + 11 * 1024 > 227 * 1024) return set_err(ctx, PCS_ERR_ARG, "conv_fold: %zu bytes of shared memory needed", smem);
    static size_t attr_set = 0;                    // static shared memory (exchange buffers) counts against the 227 KB too
    if (attr_set < smem) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(conv_fold_kernel<T, NPAD, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int grid = std::min(p.num_items, ctx->sm_count);
    conv_fold_kernel<T, NPAD, NCH><<<grid, F_THREADS, smem, ctx->stream>>>(p, tm);
    PCS_LAUNCH_CHECK(ctx, "conv_fold_kernel");
    return PCS_OK;
}

template <typename T>
int launch_fold_dispatch(pcs_ctx* ctx, const FoldConvArgs& a) {
    const int key = a.npad * 10 + a.nchunks;
    switch (key) {
        case 322: return launch_fold_t<T, 32, 2>(ctx, a);     // conv2: 20(32) -> 30(32)
        case 482: return launch_fold_t<T, 48, 2>(ctx, a);     // conv3: 30(32) -> 40(48)
        case 483: return launch_fold_t<T, 48, 3>(ctx, a);     // conv4: 40(48) -> 40(48)
        default: return set_err(ctx, PCS_ERR_ARG, "conv_fold: no instantiation for N=%d chunks=%d", a.npad, a.nchunks);
    }
}

}  // namespace

bool fold_supported(int k, int npad, int nchunks, int nsrc) {
    if (k != 5 || nsrc != 1) return false;
    const int key = npad * 10 + nchunks;
    return key == 322 || key == 482 || key == 483;
}

// Resident operand image [chunk][dy][plane][row = dx*NPAD + o][8]
size_t fold_weight_image(const float* w32 /*[25][cin][cout]*/, int cin, int cout, int npad, int precision, std::vector<uint16_t>& out) {
    const int nch = pad16(cin) / 16, nf = 5 * npad;
    out.assign((size_t)nch * 5 * 2 * nf * 8, 0);
    auto conv = [&](float v) -> uint16_t {
        if (precision == PCS_PREC_BF16) { __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b); }
        __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h);
    };
    for (int c = 0; c < nch; ++c)
        for (int dy = 0; dy < 5; ++dy)
            for (int pl = 0; pl < 2; ++pl)
                for (int dx = 0; dx < 5; ++dx)
                    for (int o = 0; o < npad; ++o)
                        for (int e = 0; e < 8; ++e) {
                            const int ci = c * 16 + pl * 8 + e;
                            if (ci >= cin || o >= cout) continue;
                            const float v = w32[((size_t)(dy * 5 + dx) * cin + ci) * cout + o];
                            out[((((size_t)c * 5 + dy) * 2 + pl) * nf + dx * npad + o) * 8 + e] = conv(v);
                        }
    return out.size() * sizeof(uint16_t);
}

int launch_conv_fold(pcs_ctx* ctx, const FoldConvArgs& a) {
    if (!fold_supported(a.k, a.npad, a.nchunks, 1)) return set_err(ctx, PCS_ERR_ARG, "conv_fold: unsupported layer");
    if (ctx->precision == PCS_PREC_BF16) return launch_fold_dispatch<__nv_bfloat16>(ctx, a);
    return launch_fold_dispatch<__half>(ctx, a);
}

}  // namespace pcs
