// Implicit-GEMM 'same' convolution on the 5th-generation tensor cores:
// tcgen05.mma (kind::f16, bf16|fp16 operands, fp32 accumulators in TMEM) fed by
// TMA, warp-specialised, persistent.
//
// Reference layers served: Conv2D k5/k3 'same' and Conv2DTranspose k5 s1 (as a
// pre-flipped correlation) of ocr4all_pixel_classifier/lib/model.py:45-92,
// :151-203, :206-234, with the channel concatenation of the skip connections
// read as two K segments (the concat tensor is never materialised) and
// MaxPooling2D(2,2) fused into the epilogue.
//
// GEMM view per CTA tile:  D[128 px, NPAD] += A[128 px, 16 ch] * B[16 ch, NPAD]
//   M = 128 consecutive pixels of one image row (one TMEM lane per pixel),
//   N = C_out padded to 16 (tile of <= 128),
//   K = taps x C_in, walked as (16-channel chunk) x (tap).
// A CTA owns a strip of 128 px x R rows and keeps R accumulators (R x NPAD TMEM
// columns).  Per 16-channel chunk the producer brings in, with TMA,
//   * the input patch  [2 planes][R+k-1 rows][128 px][8 ch]  (zero filled outside
//     the image = the 'same' zero border): ONE TMA box whose rows are 2 KB
//     contiguous runs of the plane-major activation layout (common.cuh), and
//   * the weights      [taps][2 planes][NPAD][8 ch]  (host pre-arranged image).
// Both are K-major, SWIZZLE_NONE canonical layouts: a "plane" holds 8 channels
// (16 B) per pixel at a 16-byte pixel pitch, so the A operand of tap (dy,dx) for
// accumulator row r is the SAME shared-memory patch addressed at byte offset
// ((r+dy)*128 + dx)*16 -- the im2col expansion never exists anywhere.  The patch
// is 128 px wide, so a strip yields 128-(k-1) valid output pixels; the last k-1
// MMA rows read wrapped neighbours and are discarded.
//
// mode EPI_DECONV serves Conv2DTranspose(2x2, stride 2): a 1x1 GEMM whose N axis
// is (tap, C_out) -- y[2h+i, 2w+j, o] = b[o] + sum_c x[h,w,c] K[i,j,o,c] -- with
// a scattering epilogue (model.py:71,79).
#include "common.cuh"
#include "umma_ptx.cuh"

namespace pcs {
namespace {

constexpr int TILE_M = 128;
constexpr int kMaxStages = 8;       // barrier arrays; the stage count in use is capped per mode (Cfg::STAGE_CAP)
constexpr int kHeadThreads = 64 + 8 * 32;    // fused head: 8 epilogue warps (2 per TMEM lane quarter)

template <int NPAD, int KS = 5, int MODE = 0> struct Cfg {
    // The stride-2 transposed conv (1x1 GEMM, N = 128) is bound by its output stream, not by the tensor pipe: two
    // rows per tile so that TWO accumulator stages fit TMEM (the MMAs of the next tile run under this tile's
    // stores), eight epilogue warps, a deeper ring of small stages.
    static constexpr bool STREAMING_DECONV = (MODE == 1 && KS == 1 && NPAD == 128);
    static constexpr int R = STREAMING_DECONV ? 2 : (NPAD <= 64 ? 8 : (NPAD <= 80 ? 6 : 4));      // accumulator rows per tile
    static constexpr int ACC = (2 * R * NPAD <= 512) ? 2 : 1;             // TMEM accumulator stages
    static constexpr int THREADS = (MODE >= 2 || STREAMING_DECONV) ? kHeadThreads : 192;
    static constexpr int STAGE_CAP = STREAMING_DECONV ? 8 : 4;
};

// one 32-byte store (STG.256): two neighbouring pixels of one channel plane
__device__ __forceinline__ void st_global_v8(void* ptr, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

enum { EPI_STORE = 0, EPI_DECONV = 1, EPI_HEAD = 2 };

// Fused FCN head (mode EPI_HEAD).  deconv5 = Conv2DTranspose(20, 2x2, s2, linear) (model.py:83 / :229)
// feeds the 1x1 logits layer (model.py:88 / :232) with no activation in between, so the two compose into
// ONE linear map per 2x2 tap:  M[tap][c][k] = sum_o K5[tap][c][o] * lw[o][k]  (host, double precision).
// The GEMM has N = 32 columns: J = tap*4 + k holds the high operand half of M, J = 16 + tap*4 + k the low
// half (M = hi + lo, both in the model precision: the composed weights keep ~16 mantissa bits).  The
// epilogue adds the conv2 share of the logits (written by conv2's own epilogue from its fp32
// accumulators, conv_fold.cu) and the folded bias, then takes softmax/argmax (network.py:258-259) per
// cropped output pixel (model.py:29-42).  deconv5, the concat and the logits tensor never exist in HBM.
constexpr int HEAD_NC = 4;
__constant__ float c_head_lb[HEAD_NC];     // logits bias + deconv5 bias + conv2 bias folded through the logits weights
int64_t g_head_owner[64] = {0};

struct HeadEpi {
    const float4* plog;                    // [n][2h][2w] conv2 share of the logits or null (model_fcn)
    int n_classes, hs, ws;                 // crop (model.py:29-42)
    uint8_t* labels; float* logits; float* prob;       // the colour masks are a separate, fully vectorised pass
};

struct UmmaParams {
    int n, h, w;                // input grid (conv: == output grid; deconv: output is 2h x 2w)
    int k, pad;
    int sw;                     // valid output pixels per strip = 128 - (k - 1)
    int nch0, nchunks;          // 16-channel chunks of source 0 / of all sources
    const uint8_t* wimg;        // [ntile][chunk][tap][plane][NPAD][8] operand image
    const float* bias;
    int cout, relu, mode, co_t; // co_t: padded channels per tap (deconv mode)
    void* out; int out_cp;
    void* pool; int pool_cp;
    int strips, rowblocks, ntiles_n, num_tiles;
    int a_rows;                 // R + k - 1
    uint32_t a_plane_stride, a_bytes, b_bytes, stage_bytes;   // a_bytes = 2 * a_plane_stride = TMA bytes
    int nstages;
    int debug_poison;
    HeadEpi head;
};

using namespace ptx;

template <typename T, int NPAD, int KS, int MODE>
__global__ void __launch_bounds__(Cfg<NPAD, KS, MODE>::THREADS, 1)
conv_umma_kernel(const UmmaParams p, const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1) {
    constexpr int R = Cfg<NPAD, KS, MODE>::R;
    constexpr int ACC = Cfg<NPAD, KS, MODE>::ACC;
    constexpr uint32_t IDESC = (1u << 4)                                            // D = f32
                               | ((sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 7)    // A format
                               | ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 10)                    // B format
                               | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);            // N, M; K-major A and B

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_bias[NPAD];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    if (p.debug_poison) {      // diagnosis: NaN-fill the stage buffers so that an early operand read is visible
        uint32_t* w = reinterpret_cast<uint32_t*>(stages);
        for (uint32_t i = threadIdx.x; i < (uint32_t)p.nstages * p.stage_bytes / 4; i += blockDim.x) w[i] = 0x7fc07fc0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nstages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int a = 0; a < ACC; ++a) { mbar_init(&s_tfull[a], 1); mbar_init(&s_tempty[a], (blockDim.x >> 5) - 2); }
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
    griddep_launch();
    griddep_wait();
    const uint32_t tmem_base = s_tmem_base;
    if (p.debug_poison && warp >= 2) {     // diagnosis: NaN-fill all accumulator columns of this warp's lane quarter
        const uint32_t nanv = 0x7fc00000u;
        for (int c = 0; c < 512; c += 16) {
            const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                ::"r"(ta), "r"(nanv) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
    }
    if (p.debug_poison) { __syncthreads(); tc_fence_after(); }

    const int tiles_per_page = p.strips * p.rowblocks * p.ntiles_n;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int page = tile / tiles_per_page;
                int rem = tile - page * tiles_per_page;
                const int nt = rem % p.ntiles_n; rem /= p.ntiles_n;
                const int rb = rem % p.rowblocks;
                const int strip = rem / p.rowblocks;
                const int x0 = strip * p.sw - p.pad, y0 = rb * R - p.pad;      // p.sw = 128 - (KS - 1)
                const uint8_t* wsrc = p.wimg + (size_t)nt * p.nchunks * p.b_bytes;
                for (int kc = 0; kc < p.nchunks; ++kc) {
                    mbar_wait(&s_empty[stage], phase ^ 1u);
                    uint8_t* a_dst = stages + (size_t)stage * p.stage_bytes;
                    uint8_t* b_dst = a_dst + p.a_bytes;
                    mbar_expect_tx(&s_full[stage], p.a_bytes + p.b_bytes);
                    const CUtensorMap* tm = kc < p.nch0 ? &tm0 : &tm1;
                    const int plane0 = (kc < p.nch0 ? kc : kc - p.nch0) * 2;
                    // box = 256 x u64 (128 px x 16 B) x a_rows x 2 planes; coordinate 0 counts 8-byte units
                    tma_load_4d(a_dst, tm, &s_full[stage], x0 * 2, y0, plane0, page);
                    bulk_load(b_dst, wsrc + (size_t)kc * p.b_bytes, p.b_bytes, &s_full[stage]);
                    if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const bool leader = elect_one();
        const uint32_t a_hi = (uint32_t)(make_desc(0, 0, 128) >> 32), b_hi = a_hi;
        const uint32_t a_lo_c = ((p.a_plane_stride >> 4) & 0x3fffu) << 16;           // LBO field of A
        constexpr uint32_t b_lo_c = (((uint32_t)NPAD * 16u >> 4) & 0x3fffu) << 16;   // LBO field of B
        constexpr uint32_t B_TAP16 = 2u * NPAD;                                      // 16-byte units per tap
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            // UpSampling2D + 2x2 convolution on the low-resolution grid (KS = 2, N column = parity * co_t + o): parity (i, j)
            // of the output pixel only sees the low-resolution taps dy <= i, dx <= j -- 9 of the 16 (parity, tap) blocks
            // of the operand image are non-zero -- so the taps beyond the largest i / j of this N tile's parities are skipped
            int tap_dy = KS - 1, tap_dx = KS - 1;
            if (MODE == EPI_DECONV && KS == 2) {
                const int nt = (tile % tiles_per_page) % p.ntiles_n;
                const int par_lo = (nt * NPAD) / p.co_t, par_hi = min(3, (nt * NPAD + NPAD - 1) / p.co_t);
                tap_dy = par_hi >> 1;
                tap_dx = par_lo != par_hi ? 1 : (par_lo & 1);
            }
            mbar_wait(&s_tempty[acc], acc_phase ^ 1u);
            tc_fence_after();
            for (int kc = 0; kc < p.nchunks; ++kc) {
                mbar_wait(&s_full[stage], phase);
                tc_fence_after();
                if (leader) {
                    const uint32_t a_base = smem_u32(stages + (size_t)stage * p.stage_bytes);
                    const uint32_t a_lo0 = ((a_base >> 4) & 0x3fffu) | a_lo_c;
                    const uint32_t b_lo0 = (((a_base + p.a_bytes) >> 4) & 0x3fffu) | b_lo_c;
                    const uint32_t d0 = tmem_base + (uint32_t)(acc * R * NPAD);
                    const uint32_t first = kc ? 1u : 0u;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
                            for (int dx = 0; dx < KS; ++dx) {
                                if (dy > tap_dy || dx > tap_dx) continue;
                                const uint32_t a_off = (uint32_t)((r + dy) * TILE_M + dx);      // 16-byte units
                                const uint32_t b_off = (uint32_t)(dy * KS + dx) * B_TAP16;
                                tc_mma(d0 + (uint32_t)(r * NPAD), a_lo0 + a_off, a_hi, b_lo0 + b_off, b_hi, IDESC,
                                       (dy | dx) ? 1u : first);
                            }
                        }
                    }
                    tc_commit(&s_empty[stage]);                    // smem slot free once these MMAs retire
                    if (kc == p.nchunks - 1) tc_commit(&s_tfull[acc]);
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
            }
            if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ===================== epilogue warps =====================
        // warps 2.. : TMEM lane quarter = warp & 3; with 8 epilogue warps two warps share a quarter
        // and split the accumulator rows between them (group 0 / 1)
        const int quarter = warp & 3;
        const int group = (warp - 2) >> 2, ngroups = ((blockDim.x >> 5) - 2) >> 2;
        const int epi_threads = blockDim.x - 64;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int page = tile / tiles_per_page;
            int rem = tile - page * tiles_per_page;
            const int nt = rem % p.ntiles_n; rem /= p.ntiles_n;
            const int rb = rem % p.rowblocks;
            const int strip = rem / p.rowblocks;
            const int m = quarter * 32 + lane;                     // MMA row = pixel within the strip
            const int x = strip * p.sw + m;
            const bool xok = m < p.sw && x < p.w;
            const int y0 = rb * R;
            const int cbase = nt * NPAD;
            // fused head: conv2's share of the logits of the 2x2 output block of input row yy is requested one row of this
            // warp AHEAD of its use -- the first one here, before the waits (the first use of a freshly requested value was
            // 37 % of this kernel's stall samples)
            const HeadEpi& hd = p.head;
            const int oh = 2 * p.h, ow = 2 * p.w;
            auto request = [&](int yy, float4 (&q)[2][2]) {
#pragma unroll
                for (int i2 = 0; i2 < 2; ++i2) {
                    const int oy = 2 * yy + i2;
                    if (MODE >= EPI_HEAD && hd.plog && xok && yy < p.h && oy < hd.hs) {
                        // the two pixels of an output row: 32 contiguous bytes per thread, 1 KB per warp
                        const float4* src = hd.plog + ((size_t)page * oh + oy) * ow + 2 * x;
                        q[i2][0] = __ldg(src);
                        q[i2][1] = __ldg(src + 1);
                    } else {
                        q[i2][0] = q[i2][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            };
            float4 nx[2][2];
            if constexpr (MODE >= EPI_HEAD) request(y0 + group, nx);
            // per-column bias of this N tile (named barrier 1 over the epilogue warps)
            asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");
            for (int i = threadIdx.x - 64; i < NPAD; i += epi_threads) {
                int o = cbase + i;
                if (MODE == EPI_DECONV) o = (cbase + i) % p.co_t;
                if (MODE >= EPI_HEAD) o = 0x7fffffff;              // all biases are folded into c_head_lb
                s_bias[i] = o < p.cout ? __ldg(p.bias + o) : 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");

            mbar_wait(&s_tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * R * NPAD);

            if constexpr (MODE >= EPI_HEAD) {
                // one thread = one input pixel of deconv5 = a 2x2 block of output pixels x 4 classes
#pragma unroll 1
                for (int r = group; r < R; r += ngroups) {
                    const int y = y0 + r;
                    uint32_t vh[16], vl[16];
                    tmem_ld16(t_lane + (uint32_t)(r * NPAD), vh);
                    tmem_ld16(t_lane + (uint32_t)(r * NPAD + 16), vl);
                    const bool rok = xok && y < p.h;
                    float4 pl[2][2];
#pragma unroll
                    for (int i2 = 0; i2 < 2; ++i2) { pl[i2][0] = nx[i2][0]; pl[i2][1] = nx[i2][1]; }
                    if (r + ngroups < R) request(y + ngroups, nx);
                    tmem_ld_wait();
#pragma unroll
                    for (int t = 0; t < 4; ++t) {                      // tap = 2*i2 + j -> output pixel (2y+i2, 2x+j)
                        const int oy = 2 * y + (t >> 1), ox = 2 * x + (t & 1);
                        if (!(rok && oy < hd.hs && ox < hd.ws)) continue;
                        const float4 pp = pl[t >> 1][t & 1];
                        float lg[HEAD_NC];
                        lg[0] = (__uint_as_float(vh[t * 4 + 0]) + __uint_as_float(vl[t * 4 + 0])) + pp.x + c_head_lb[0];
                        lg[1] = (__uint_as_float(vh[t * 4 + 1]) + __uint_as_float(vl[t * 4 + 1])) + pp.y + c_head_lb[1];
                        lg[2] = (__uint_as_float(vh[t * 4 + 2]) + __uint_as_float(vl[t * 4 + 2])) + pp.z + c_head_lb[2];
                        lg[3] = (__uint_as_float(vh[t * 4 + 3]) + __uint_as_float(vl[t * 4 + 3])) + pp.w + c_head_lb[3];
                        int best = 0;
                        float bv = lg[0];
#pragma unroll
                        for (int k = 1; k < HEAD_NC; ++k)
                            if (k < hd.n_classes && lg[k] > bv) { bv = lg[k]; best = k; }      // first maximum wins
                        const size_t opix = ((size_t)page * hd.hs + oy) * hd.ws + ox;
                        if (hd.labels) hd.labels[opix] = (uint8_t)best;
                        if (hd.logits) {
#pragma unroll
                            for (int k = 0; k < HEAD_NC; ++k)
                                if (k < hd.n_classes) hd.logits[opix * hd.n_classes + k] = lg[k];
                        }
                        if (hd.prob) {
                            float e[HEAD_NC], sum = 0.f;
#pragma unroll
                            for (int k = 0; k < HEAD_NC; ++k) { e[k] = k < hd.n_classes ? expf(lg[k] - bv) : 0.f; sum += e[k]; }
#pragma unroll
                            for (int k = 0; k < HEAD_NC; ++k)
                                if (k < hd.n_classes) hd.prob[opix * hd.n_classes + k] = e[k] / sum;
                        }
                    }
                }
            } else if (MODE == EPI_DECONV && 2 * p.co_t <= NPAD && (p.co_t & 15) == 0) {
                // both horizontal taps (2i, 2i+1) of an output row lie in this N tile: one thread owns the output pixel
                // pair (2y+i, 2x), (2y+i, 2x+1) of a plane = 32 contiguous bytes = one STG.256; a warp writes 1 KB runs
                T* out = reinterpret_cast<T*>(p.out);
                const int oh = 2 * p.h, ow = 2 * p.w;
                const int tap_base = cbase / p.co_t;                       // first tap of this N tile (even)
#pragma unroll 1
                for (int r = group; r < R; r += ngroups) {
                    const int y = y0 + r;
                    const bool ok = xok && y < p.h;
#pragma unroll 1
                    for (int c = 0; c < NPAD / 2; c += 16) {               // 16 channels of the even tap of a pair
                        const int tp = c / p.co_t, o = c - tp * p.co_t;    // tap pair within the tile, first channel
                        const int c0 = 2 * tp * p.co_t + o, c1 = c0 + p.co_t;
                        uint32_t v0[16], v1[16];
                        tmem_ld16(t_lane + (uint32_t)(r * NPAD + c0), v0);
                        tmem_ld16(t_lane + (uint32_t)(r * NPAD + c1), v1);
                        tmem_ld_wait();
                        float f0[16], f1[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float b = s_bias[c0 + i];                // same channel, same bias for both taps
                            f0[i] = __uint_as_float(v0[i]) + b;
                            f1[i] = __uint_as_float(v1[i]) + b;
                            if (p.relu) { f0[i] = fmaxf(f0[i], 0.f); f1[i] = fmaxf(f1[i], 0.f); }
                        }
                        if (ok && o < p.out_cp) {
                            const int oy = 2 * y + ((tap_base + 2 * tp) >> 1);
                            st_global_v8(out + act_idx(page, p.out_cp, oh, ow, o, oy, 2 * x),
                                         make_uint4(pack2<T>(f0[0], f0[1]), pack2<T>(f0[2], f0[3]), pack2<T>(f0[4], f0[5]), pack2<T>(f0[6], f0[7])),
                                         make_uint4(pack2<T>(f1[0], f1[1]), pack2<T>(f1[2], f1[3]), pack2<T>(f1[4], f1[5]), pack2<T>(f1[6], f1[7])));
                            if (o + 8 < p.out_cp)
                                st_global_v8(out + act_idx(page, p.out_cp, oh, ow, o + 8, oy, 2 * x),
                                             make_uint4(pack2<T>(f0[8], f0[9]), pack2<T>(f0[10], f0[11]), pack2<T>(f0[12], f0[13]), pack2<T>(f0[14], f0[15])),
                                             make_uint4(pack2<T>(f1[8], f1[9]), pack2<T>(f1[10], f1[11]), pack2<T>(f1[12], f1[13]), pack2<T>(f1[14], f1[15])));
                        }
                    }
                }
            } else {
                T* out = reinterpret_cast<T*>(p.out);
                T* pool = reinterpret_cast<T*>(p.pool);
#pragma unroll 1
                for (int rp = group; rp < R / 2; rp += ngroups) {
                    const int ya = y0 + 2 * rp;
#pragma unroll 1
                    for (int c16 = 0; c16 < NPAD / 16; ++c16) {
                        uint32_t v0[16], v1[16];
                        tmem_ld16(t_lane + (uint32_t)((2 * rp) * NPAD + c16 * 16), v0);
                        tmem_ld16(t_lane + (uint32_t)((2 * rp + 1) * NPAD + c16 * 16), v1);
                        tmem_ld_wait();
                        float f0[16], f1[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float b = s_bias[c16 * 16 + i];
                            f0[i] = __uint_as_float(v0[i]) + b;
                            f1[i] = __uint_as_float(v1[i]) + b;
                            if (p.relu) { f0[i] = fmaxf(f0[i], 0.f); f1[i] = fmaxf(f1[i], 0.f); }
                        }
                        const uint4 lo0 = make_uint4(pack2<T>(f0[0], f0[1]), pack2<T>(f0[2], f0[3]), pack2<T>(f0[4], f0[5]), pack2<T>(f0[6], f0[7]));
                        const uint4 hi0 = make_uint4(pack2<T>(f0[8], f0[9]), pack2<T>(f0[10], f0[11]), pack2<T>(f0[12], f0[13]), pack2<T>(f0[14], f0[15]));
                        const uint4 lo1 = make_uint4(pack2<T>(f1[0], f1[1]), pack2<T>(f1[2], f1[3]), pack2<T>(f1[4], f1[5]), pack2<T>(f1[6], f1[7]));
                        const uint4 hi1 = make_uint4(pack2<T>(f1[8], f1[9]), pack2<T>(f1[10], f1[11]), pack2<T>(f1[12], f1[13]), pack2<T>(f1[14], f1[15]));
                        if constexpr (MODE == EPI_DECONV) {
                            // column J = (tap, o): scatter to output pixel (2y + tap/2, 2x + tap%2)
                            const int J = cbase + c16 * 16;
                            const int tap = J / p.co_t, o = J - tap * p.co_t;
                            if (xok && o < p.out_cp) {
                                const int ox = 2 * x + (tap & 1), oh = 2 * p.h, ow = 2 * p.w;
                                const bool two = o + 8 < p.out_cp;
                                if (ya < p.h) {
                                    const int oy = 2 * ya + (tap >> 1);
                                    *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, oh, ow, o, oy, ox)) = lo0;
                                    if (two) *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, oh, ow, o + 8, oy, ox)) = hi0;
                                }
                                if (ya + 1 < p.h) {
                                    const int oy = 2 * (ya + 1) + (tap >> 1);
                                    *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, oh, ow, o, oy, ox)) = lo1;
                                    if (two) *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, oh, ow, o + 8, oy, ox)) = hi1;
                                }
                            }
                        } else {
                            const int o = cbase + c16 * 16;
                            if (out && xok && o < p.out_cp) {
                                const bool two = o + 8 < p.out_cp;
                                if (ya < p.h) {
                                    *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, o, ya, x)) = lo0;
                                    if (two) *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, o + 8, ya, x)) = hi0;
                                }
                                if (ya + 1 < p.h) {
                                    *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, o, ya + 1, x)) = lo1;
                                    if (two) *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, o + 8, ya + 1, x)) = hi1;
                                }
                            }
                            if (pool) {
                                float mx[16];
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const float a = fmaxf(f0[i], f1[i]);
                                    mx[i] = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
                                }
                                if (!(lane & 1) && xok && ya < p.h && o < p.pool_cp) {
                                    const int ph = p.h >> 1, pw = p.w >> 1;
                                    *reinterpret_cast<uint4*>(pool + act_idx(page, p.pool_cp, ph, pw, o, ya >> 1, x >> 1)) =
                                        make_uint4(pack2<T>(mx[0], mx[1]), pack2<T>(mx[2], mx[3]), pack2<T>(mx[4], mx[5]), pack2<T>(mx[6], mx[7]));
                                    if (o + 8 < p.pool_cp)
                                        *reinterpret_cast<uint4*>(pool + act_idx(page, p.pool_cp, ph, pw, o + 8, ya >> 1, x >> 1)) =
                                        make_uint4(pack2<T>(mx[8], mx[9]), pack2<T>(mx[10], mx[11]), pack2<T>(mx[12], mx[13]), pack2<T>(mx[14], mx[15]));
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[acc]);
            if (++acc == ACC) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// plane-major activation [n][cp/8][h][w][8 x 2 B] seen as a u64 tensor (2w, h, planes, pages):
// box = 256 u64 (128 px x 16 B) x rows x 2 planes x 1 -> shared memory [plane][row][128 px][16 B]
int make_act_map(pcs_ctx* ctx, CUtensorMap* tm, const ConvSrc& s, int n, int h, int w, int rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_err(ctx, PCS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t planes = (cuuint64_t)s.cp / 8;
    const cuuint64_t dims[4] = {(cuuint64_t)w * 2, (cuuint64_t)h, planes, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16, planes * h * w * 16};
    const cuuint32_t box[4] = {256, (cuuint32_t)rows, 2, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(s.p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_err(ctx, PCS_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (cp=%d w=%d h=%d n=%d rows=%d)", (int)r, s.cp, w, h, n, rows);
    return PCS_OK;
}

template <typename T, int NPAD, int KS, int MODE>
int launch_t(pcs_ctx* ctx, const UmmaConvArgs& a) {
    constexpr int R = Cfg<NPAD, KS, MODE>::R;
    UmmaParams p{};
    p.n = a.n; p.h = a.h; p.w = a.w; p.k = a.k; p.pad = a.pad;
    // 16-channel K chunks; a source with an odd number of 8-channel planes ends in a half-empty chunk whose
    // second plane is out of bounds for the tensor map (TMA zero fill) and has zero weights
    p.nch0 = (a.src[0].cp + 15) / 16;
    p.nchunks = p.nch0 + (a.nsrc > 1 ? (a.src[1].cp + 15) / 16 : 0);
    if (p.nchunks != a.nchunks) return set_err(ctx, PCS_ERR_ARG, "conv_umma: source channel chunks %d != weight image chunks %d", p.nchunks, a.nchunks);
    p.wimg = reinterpret_cast<const uint8_t*>(a.wmma); p.bias = a.b32; p.cout = a.cout; p.relu = a.relu;
    p.out = a.out; p.out_cp = a.out_cp; p.pool = a.pool_out; p.pool_cp = a.pool_cp;
    if (a.mode != MODE) return set_err(ctx, PCS_ERR_ARG, "conv_umma: mode dispatch mismatch");
    p.mode = a.mode; p.co_t = a.co_t;
    { const char* e = getenv("PCSEG_DEBUG_POISON"); p.debug_poison = e && e[0] == '1'; }
    p.sw = TILE_M - (a.k - 1);
    p.strips = (a.w + p.sw - 1) / p.sw;
    p.rowblocks = (a.h + R - 1) / R;
    const int ncols = a.mode == EPI_STORE ? a.cout : (a.mode == EPI_DECONV ? 4 * a.co_t : NPAD);
    p.ntiles_n = (ncols + NPAD - 1) / NPAD;
    p.num_tiles = a.n * p.strips * p.rowblocks * p.ntiles_n;
    p.a_rows = R + a.k - 1;
    p.a_plane_stride = (uint32_t)p.a_rows * TILE_M * 16;
    p.a_bytes = 2 * p.a_plane_stride;
    p.b_bytes = (uint32_t)a.k * a.k * 2 * NPAD * 16;
    p.stage_bytes = (p.a_bytes + p.b_bytes + 1023) / 1024 * 1024;
    const uint32_t budget = 225 * 1024 - 1024;
    p.nstages = (int)std::min<uint32_t>(Cfg<NPAD, KS, MODE>::STAGE_CAP, budget / p.stage_bytes);
    if (p.nstages < 2) return set_err(ctx, PCS_ERR_ARG, "conv_umma: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
    if ((a.h & 1) || (a.w & 1)) return set_err(ctx, PCS_ERR_ARG, "conv_umma: odd grid");
    if (a.pool_out && (p.sw & 1)) return set_err(ctx, PCS_ERR_ARG, "conv_umma: fused pooling needs an even strip width");
    const size_t smem = std::max<size_t>((size_t)p.nstages * p.stage_bytes + 1024, kSoloSmem);
    CUtensorMap tm0, tm1;
    PCS_TRY(make_act_map(ctx, &tm0, a.src[0], a.n, a.h, a.w, p.a_rows));
    if (a.nsrc > 1) PCS_TRY(make_act_map(ctx, &tm1, a.src[1], a.n, a.h, a.w, p.a_rows));
    else tm1 = tm0;
    static bool attr_set[64] = {};               // the attribute is per device
    if (ctx->device >= 64 || !attr_set[ctx->device]) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(conv_umma_kernel<T, NPAD, KS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        if (ctx->device < 64) attr_set[ctx->device] = true;
    }
    const int grid = std::min(p.num_tiles, ctx->sm_count);
    if constexpr (MODE >= EPI_HEAD) {
        const UmmaHeadArgs& h = *a.head;
        p.head.plog = reinterpret_cast<const float4*>(h.plog); p.head.n_classes = h.n_classes; p.head.hs = h.hs; p.head.ws = h.ws;
        p.head.labels = h.labels; p.head.logits = h.logits; p.head.prob = h.prob;
        if (ctx->device >= 64 || g_head_owner[ctx->device] != ctx->model_stamp) {
            PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_head_lb, h.lb_folded, sizeof(float) * HEAD_NC, 0, cudaMemcpyDeviceToDevice, ctx->stream));
            if (ctx->device < 64) g_head_owner[ctx->device] = ctx->model_stamp;
        }
    }
    PCS_CUDA(ctx, launch_kernel_pdl(conv_umma_kernel<T, NPAD, KS, MODE>, dim3(grid), dim3(Cfg<NPAD, KS, MODE>::THREADS), smem,
                                    ctx->stream, ctx->pdl, p, tm0, tm1));
    PCS_LAUNCH_CHECK(ctx, "conv_umma_kernel");
    return PCS_OK;
}

template <typename T>
int launch_npad(pcs_ctx* ctx, const UmmaConvArgs& a) {
    const int key = a.mode * 100000 + a.k * 1000 + a.npad;
    switch (key) {
        case 5032: return launch_t<T, 32, 5, EPI_STORE>(ctx, a);
        case 5048: return launch_t<T, 48, 5, EPI_STORE>(ctx, a);
        case 5064: return launch_t<T, 64, 5, EPI_STORE>(ctx, a);
        case 5080: return launch_t<T, 80, 5, EPI_STORE>(ctx, a);
        case 3064: return launch_t<T, 64, 3, EPI_STORE>(ctx, a);
        case 3128: return launch_t<T, 128, 3, EPI_STORE>(ctx, a);
        case 1080: return launch_t<T, 80, 1, EPI_STORE>(ctx, a);         // input gradients of the stride-2 transposed convs (train_tc.cu)
        case 1128: return launch_t<T, 128, 1, EPI_STORE>(ctx, a);
        case 101128: return launch_t<T, 128, 1, EPI_DECONV>(ctx, a);
        case 102128: return launch_t<T, 128, 2, EPI_DECONV>(ctx, a);      // UpSampling2D(2) + Conv2D(2x2) of the U-Net
        case 201032: return launch_t<T, 32, 1, EPI_HEAD>(ctx, a);
        default: return set_err(ctx, PCS_ERR_ARG, "conv_umma: no instantiation for mode=%d k=%d N tile %d", a.mode, a.k, a.npad);
    }
}

}  // namespace

bool umma_supported(int k, int npad) {
    const int key = k * 1000 + npad;
    return key == 5032 || key == 5048 || key == 5064 || key == 5080 || key == 3064 || key == 3128 || key == 1128 || key == 1032 || key == 2128;
}

// Operand image [ntile][chunk][tap][plane][NPAD][8]; chunk runs over the 16-channel groups of
// source 0 then source 1 (each padded to a multiple of 16 channels with zeros).
size_t umma_weight_image(const float* w32, int taps, const int* src_c, int nsrc, int cout, int npad, int precision,
                         std::vector<uint16_t>& out) {
    int cin = 0, nchunks = 0;
    for (int s = 0; s < nsrc; ++s) { cin += src_c[s]; nchunks += pad16(src_c[s]) / 16; }
    const int ntiles = (cout + npad - 1) / npad;
    out.assign((size_t)ntiles * nchunks * taps * 2 * npad * 8, 0);
    auto conv = [&](float v) -> uint16_t {
        if (precision == PCS_PREC_BF16) {
            __nv_bfloat16 b = __float2bfloat16_rn(v);
            return *reinterpret_cast<uint16_t*>(&b);
        }
        __half h = __float2half_rn(v);
        return *reinterpret_cast<uint16_t*>(&h);
    };
    for (int nt = 0; nt < ntiles; ++nt) {
        int chunk = 0, cbase = 0;
        for (int s = 0; s < nsrc; ++s) {
            for (int lc = 0; lc < pad16(src_c[s]) / 16; ++lc, ++chunk)
                for (int t = 0; t < taps; ++t)
                    for (int pl = 0; pl < 2; ++pl)
                        for (int nn = 0; nn < npad; ++nn)
                            for (int e = 0; e < 8; ++e) {
                                const int c = lc * 16 + pl * 8 + e, o = nt * npad + nn;
                                if (c >= src_c[s] || o >= cout) continue;
                                const float v = w32[((size_t)t * cin + cbase + c) * cout + o];
                                out[(((((size_t)nt * nchunks + chunk) * taps + t) * 2 + pl) * npad + nn) * 8 + e] = conv(v);
                            }
            cbase += src_c[s];
        }
    }
    return out.size() * sizeof(uint16_t);
}

// Operand image of a 2x2 stride-2 transposed convolution: one "tap" (1x1 GEMM), N column J = t * co_t + o.
size_t umma_weight_image_deconv(const float* w32 /*[4][cin][cout]*/, const int* src_c, int nsrc, int cout, int co_t,
                                int npad, int precision, std::vector<uint16_t>& out) {
    int cin = 0, nchunks = 0;
    for (int s = 0; s < nsrc; ++s) { cin += src_c[s]; nchunks += pad16(src_c[s]) / 16; }
    const int ncols = 4 * co_t, ntiles = (ncols + npad - 1) / npad;
    out.assign((size_t)ntiles * nchunks * 2 * npad * 8, 0);
    auto conv = [&](float v) -> uint16_t {
        if (precision == PCS_PREC_BF16) {
            __nv_bfloat16 b = __float2bfloat16_rn(v);
            return *reinterpret_cast<uint16_t*>(&b);
        }
        __half h = __float2half_rn(v);
        return *reinterpret_cast<uint16_t*>(&h);
    };
    for (int nt = 0; nt < ntiles; ++nt) {
        int chunk = 0, cbase = 0;
        for (int s = 0; s < nsrc; ++s) {
            for (int lc = 0; lc < pad16(src_c[s]) / 16; ++lc, ++chunk)
                for (int pl = 0; pl < 2; ++pl)
                    for (int nn = 0; nn < npad; ++nn)
                        for (int e = 0; e < 8; ++e) {
                            const int c = lc * 16 + pl * 8 + e, J = nt * npad + nn;
                            const int t = J / co_t, o = J % co_t;
                            if (c >= src_c[s] || J >= ncols || o >= cout) continue;
                            const float v = w32[((size_t)t * cin + cbase + c) * cout + o];
                            out[((((size_t)nt * nchunks + chunk) * 2 + pl) * npad + nn) * 8 + e] = conv(v);
                        }
            cbase += src_c[s];
        }
    }
    return out.size() * sizeof(uint16_t);
}

// Operand image of UpSampling2D(2, nearest) -> Conv2D(2x2, 'same') (model.py:176-196) as ONE 2x2 convolution on
// the low-resolution grid with N column J = parity * co_t + o, parity = (i, j) of the output pixel (2y+i, 2x+j):
//   out[2y+i][2x+j] = sum_{u,v} W[u][v] . in[(2y+i+u) >> 1][(2x+j+v) >> 1]
// so the low-resolution tap (dy, dx) of parity (i, j) carries  sum of W[u][v] over (i+u)>>1 == dy, (j+v)>>1 == dx
// (9 of the 16 (parity, tap) blocks are non-zero).  Layout [ntile][chunk][tap = dy*2+dx][plane][npad][8].
size_t umma_weight_image_up2(const float* w32 /*[4][cin][cout], tap = u*2+v*/, int cin, int cout, int co_t, int npad,
                             int precision, std::vector<uint16_t>& out) {
    const int nchunks = pad16(cin) / 16, ncols = 4 * co_t, ntiles = (ncols + npad - 1) / npad;
    out.assign((size_t)ntiles * nchunks * 4 * 2 * npad * 8, 0);
    auto conv = [&](float v) -> uint16_t {
        if (precision == PCS_PREC_BF16) { __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b); }
        __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h);
    };
    for (int nt = 0; nt < ntiles; ++nt)
        for (int ch = 0; ch < nchunks; ++ch)
            for (int t = 0; t < 4; ++t)
                for (int pl = 0; pl < 2; ++pl)
                    for (int nn = 0; nn < npad; ++nn)
                        for (int e = 0; e < 8; ++e) {
                            const int c = ch * 16 + pl * 8 + e, J = nt * npad + nn;
                            const int par = J / co_t, o = J % co_t;
                            if (c >= cin || J >= ncols || o >= cout) continue;
                            const int i = par >> 1, j = par & 1, dy = t >> 1, dx = t & 1;
                            float sum = 0.f;
                            bool any = false;
                            for (int u = 0; u < 2; ++u)
                                for (int v = 0; v < 2; ++v)
                                    if (((i + u) >> 1) == dy && ((j + v) >> 1) == dx) {
                                        sum += w32[((size_t)(u * 2 + v) * cin + c) * cout + o];
                                        any = true;
                                    }
                            if (any) out[(((((size_t)nt * nchunks + ch) * 4 + t) * 2 + pl) * npad + nn) * 8 + e] = conv(sum);
                        }
    return out.size() * sizeof(uint16_t);
}

// Operand image of the fused head: composed deconv5 x logits map m[tap][cin][4] split into two operand
// halves; N column J = tap*4 + k (high half), 16 + tap*4 + k (low half); layout [chunk][plane][32][8].
size_t umma_weight_image_head(const double* m /*[4][cin][4]*/, const int* src_c, int nsrc, int precision,
                              std::vector<uint16_t>& out) {
    int cin = 0, nchunks = 0;
    for (int s = 0; s < nsrc; ++s) { cin += src_c[s]; nchunks += pad16(src_c[s]) / 16; }
    constexpr int npad = 32;
    out.assign((size_t)nchunks * 2 * npad * 8, 0);
    auto split = [&](double v, uint16_t& hi, uint16_t& lo) {
        if (precision == PCS_PREC_BF16) {
            __nv_bfloat16 h = __float2bfloat16_rn((float)v);
            __nv_bfloat16 l = __float2bfloat16_rn((float)(v - (double)__bfloat162float(h)));
            hi = *reinterpret_cast<uint16_t*>(&h); lo = *reinterpret_cast<uint16_t*>(&l);
        } else {
            __half h = __float2half_rn((float)v);
            __half l = __float2half_rn((float)(v - (double)__half2float(h)));
            hi = *reinterpret_cast<uint16_t*>(&h); lo = *reinterpret_cast<uint16_t*>(&l);
        }
    };
    int chunk = 0, cbase = 0;
    for (int s = 0; s < nsrc; ++s) {
        for (int lc = 0; lc < pad16(src_c[s]) / 16; ++lc, ++chunk)
            for (int pl = 0; pl < 2; ++pl)
                for (int e = 0; e < 8; ++e) {
                    const int c = lc * 16 + pl * 8 + e;
                    if (c >= src_c[s]) continue;
                    for (int t = 0; t < 4; ++t)
                        for (int k = 0; k < 4; ++k) {
                            uint16_t hi, lo;
                            split(m[((size_t)t * cin + cbase + c) * 4 + k], hi, lo);
                            out[(((size_t)chunk * 2 + pl) * npad + t * 4 + k) * 8 + e] = hi;
                            out[(((size_t)chunk * 2 + pl) * npad + 16 + t * 4 + k) * 8 + e] = lo;
                        }
                }
        cbase += src_c[s];
    }
    return out.size() * sizeof(uint16_t);
}

int launch_conv_umma(pcs_ctx* ctx, const UmmaConvArgs& a) {
    if (a.mode >= EPI_HEAD && (!a.head || a.npad != 32 || a.head->n_classes > HEAD_NC))
        return set_err(ctx, PCS_ERR_ARG, "conv_umma: fused head needs the N=32 operand image and <= %d classes", HEAD_NC);
    const bool plain_1x1 = a.mode == 0 && a.k == 1 && (a.npad == 80 || a.npad == 128);
    if (!umma_supported(a.k, a.npad) && !plain_1x1) return set_err(ctx, PCS_ERR_ARG, "conv_umma: unsupported k=%d N=%d", a.k, a.npad);
    if (a.src[0].cp % 8 || (a.nsrc > 1 && a.src[1].cp % 8)) return set_err(ctx, PCS_ERR_ARG, "conv_umma: channel stride not a multiple of 8");
    if ((a.out && a.out_cp % 8) || (a.pool_out && a.pool_cp % 8)) return set_err(ctx, PCS_ERR_ARG, "conv_umma: output stride");
    if (ctx->precision == PCS_PREC_BF16) return launch_npad<__nv_bfloat16>(ctx, a);
    return launch_npad<__half>(ctx, a);
}

}  // namespace pcs
