// conv1 + conv2 of the FCN variants in ONE marching kernel (model.py:50-54 / :211-213):
//   conv1 = ReLU(Conv2D(20, 5x5, 'same')(x / 255)),   conv2 = Conv2D(30, 5x5, 'same')(conv1),   pool2 = MaxPool(conv2)
// The separate kernels (conv1_umma.cu, conv_fold.cu) hand conv1 over through HBM: 47 MB per A4 page written and read
// straight back, a fifth of the step for 1 % of its arithmetic.  Here a CTA marches down a 124-pixel strip like
// conv_fold_kernel does, but the input rows of conv2 never exist in global memory: per output row
//   * four builder warps expand ONE new row of the uint8 page into the K-major operand form of conv1_umma.cu
//     (one 16-byte unit = 8 consecutive pixels per pixel slot; a ring of 8 rows + a mirror of row 0 so that
//     the pair (row r, row r + 1) is always two consecutive ring entries);
//   * a second MMA thread issues conv1 for the rows ahead (K = vertical tap pairs x 8 horizontal slots, N = 32:
//     3 MMAs, 6 with bf16 operands whose weights are split hi + lo) into one of four 32-column accumulators that sit
//     in TMEM behind the conv2 ring (12 x 32 columns);
//   * four conv1-epilogue warps drain it (x 1/255, bias, ReLU, zero outside the grid = conv2's 'same' border), pack
//     and store the row into conv2's shared-memory input ring in exactly the layout the TMA box used to deliver;
//   * conv2 proceeds as in conv_fold.cu (five vertical taps folded into N' = 160, ring of output-row accumulators,
//     logits share + 2x2 max-pool in the epilogue).
// Channels 16..19 of conv1 travel as a PIXEL-PAIR plane: conv1's GEMM has 12 idle output columns, four of them now
// compute channels 16..19 of the pixel one to the right (the same taps one slot further), so the third plane holds
// (ch16-19 of x | ch16-19 of x + 1) and one K = 16 step of conv2 covers FOUR horizontal taps of those channels: 7 MMAs
// per row instead of 8.
#include "common.cuh"
#include "umma_ptx.cuh"

namespace pcs {
namespace {
using namespace ptx;

constexpr int G_SW = 124;                    // valid output pixels per strip
constexpr int G_NPAD = 32, G_NF = 5 * G_NPAD, G_NMMA = 7;
constexpr int G_SLOTS = 12, G_RING = 12;     // conv2: accumulator slots (32 TMEM columns each) / input-row ring; period 12
constexpr int G_ER = 8;                      // expanded page rows in flight (+ 1 mirror entry)
constexpr int G_D1 = 4;                      // conv1 accumulators (TMEM columns 384 ...)
constexpr uint32_t G_PLANE = 2048;           // 128 pixels x 16 bytes
constexpr uint32_t G_ROW_BYTES = 3 * G_PLANE;
constexpr uint32_t G_WDX_BYTES = 2 * G_NF * 16;
constexpr uint32_t G_W2_BYTES = G_NMMA * G_WDX_BYTES;                 // 35 840
constexpr uint32_t G_W1_HALF = 3 * 2 * 32 * 16;                       // [ks][plane][n][8] = 3 072 bytes per operand half
constexpr int G_EG = 3;                      // conv2 epilogue groups of four warps
constexpr int G_THREADS = 320 + G_EG * 128;  // warp 0 weights + conv1 MMAs, 1 conv2 MMAs, 2-5 conv1 epilogue, 6-9 builders, 10.. conv2 epilogue
constexpr int G_CVT_COPY = 146;              // elements per converted copy (73 words: odd, the two copies use different banks)
constexpr int G_LOGC = 32, G_NC = 4;
__constant__ float c_skip_lw12[G_LOGC * G_NC];
int64_t g_skip12_owner[64] = {0};

template <typename T> __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t max2<__nv_bfloat16>(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t max2<__half>(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// PCSEG_C12_STATS=1: cycles spent in the protocol's waits, summed over the CTAs (who waits for whom)
__device__ unsigned long long g_c12_stats[16];
__device__ __forceinline__ void mbar_wait_cnt(uint64_t* bar, uint32_t parity, uint32_t& cycles, int stats) {
    if (!stats) { mbar_wait(bar, parity); return; }
    const uint32_t t0 = (uint32_t)clock();              // try_wait itself suspends the thread for a while: time all of it
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++n > (1u << 24)) __trap();
    }
    cycles += (uint32_t)clock() - t0;
}

struct FusedParams {
    const uint8_t* img; int img_h, img_w;    // real page
    int n, h, w;                             // padded grid
    int strips;
    long long total_rows, rows_per_cta;
    const uint8_t* w1img;                    // conv1 operand image, G_W1_HALF bytes per half
    const uint8_t* w2img;                    // conv2 resident operand image, G_W2_BYTES
    float bias1[24];                         // conv1 bias; [20..23] = [16..19] (the pixel-pair columns)
    float bias2[32];
    int halves;                              // conv1 weight halves: 1 (fp16 operands) or 2 (bf16: hi + lo)
    void* out; int out_cp;                   // full-resolution conv2 (optional)
    void* pool; int pool_cp;
    float4* plog; int plog_nc;
    int stats;
    int dbg;                                 // timing experiments only (PCSEG_C12_DEBUG): 1 = no pair-plane MMAs, 2 = no conv1 MMAs,
                                             // 4 = conv1 epilogue stores nothing, 8 = one conv2 MMA per row
};

// (page, strip, first row, rows) of the pieces of this CTA's range, in order
struct SegIter {
    long long pos, pos1;
    int h, strips;
    int page, strip, ys, rows;
    __device__ __forceinline__ SegIter(long long p0, long long p1, int h_, int strips_) : pos(p0), pos1(p1), h(h_), strips(strips_) {}
    __device__ __forceinline__ bool next() {
        if (pos >= pos1) return false;
        const int sg = (int)(pos / h);
        ys = (int)(pos - (long long)sg * h);
        rows = (int)((long long)(h - ys) < pos1 - pos ? (long long)(h - ys) : pos1 - pos);
        page = sg / strips; strip = sg - page * strips;
        pos += rows;
        return true;
    }
};

template <typename T>
__global__ void __launch_bounds__(G_THREADS, 1) conv12_fused_kernel(const FusedParams p) {
    constexpr uint32_t FMT = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
    constexpr uint32_t IDESC0 = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t IDESC1 = IDESC0 | ((uint32_t)(32 >> 3) << 17);        // conv1: N = 32

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[G_RING], s_empty[G_RING], s_tfull[G_SLOTS], s_tempty[G_SLOTS];
    __shared__ __align__(8) uint64_t s_efull[G_ER], s_eempty[G_ER], s_t1full[G_D1], s_t1empty[G_D1], s_wfull;
    __shared__ uint32_t s_tmem_base;
    __shared__ __align__(16) T s_cvt[2][2 * G_CVT_COPY];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w2 = base;                                             // 35 840 -> 36 864
    uint8_t* s_w1 = base + 36864;                                     // 6 144
    uint8_t* s_e = s_w1 + 6144;                                       // (G_ER + 1) x 2 048
    uint8_t* ring = s_e + (G_ER + 1) * G_PLANE;                       // G_RING x 6 144

    // defined (finite) values everywhere an MMA may read with a zero weight: rows that are not built yet, the units
    // behind a row's last pixel
    // (+ 8 units behind the last ring row: the zero-weight half of the last K step reads up to five units past a row)
    for (uint32_t i = threadIdx.x; i < ((G_ER + 1) * G_PLANE + G_RING * G_ROW_BYTES) / 16 + 8; i += G_THREADS)
        reinterpret_cast<uint4*>(s_e)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < G_RING; ++s) { mbar_init(&s_full[s], 4); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < G_SLOTS; ++s) { mbar_init(&s_tfull[s], 1); mbar_init(&s_tempty[s], 4); }
        for (int s = 0; s < G_ER; ++s) { mbar_init(&s_efull[s], 4); mbar_init(&s_eempty[s], 1); }
        for (int s = 0; s < G_D1; ++s) { mbar_init(&s_t1full[s], 1); mbar_init(&s_t1empty[s], 4); }
        mbar_init(&s_wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
    const long long pos0 = (long long)blockIdx.x * p.rows_per_cta;
    const long long pos1 = pos0 + p.rows_per_cta < p.total_rows ? pos0 + p.rows_per_cta : p.total_rows;

    if (warp == 0) {
        // ===================== resident weights, then the conv1 MMAs =====================
        // conv1 has its own issuing thread: the conv2 issuer's loop is all immediates (~5 instructions per MMA) and must
        // stay that way -- with conv1's runtime ring indices in the same thread the issue loop, not the tensor pipe, set
        // the pace (1 890 cycles per row against 860 of MMAs).  The two threads never touch the same accumulators; the
        // commits of a thread track that thread's MMAs.
        if (elect_one()) {
            mbar_expect_tx(&s_wfull, G_W2_BYTES + 2 * G_W1_HALF);
            for (uint32_t off = 0; off < G_W2_BYTES; off += G_WDX_BYTES) bulk_load(s_w2 + off, p.w2img + off, G_WDX_BYTES, &s_wfull);
            bulk_load(s_w1, p.w1img, 2 * G_W1_HALF, &s_wfull);
            const uint32_t hi = (uint32_t)(make_desc(0, 0, 128) >> 32);
            constexpr uint32_t lbo_plane = ((G_PLANE >> 4) & 0x3fffu) << 16;            // K halves = expanded rows r, r + 1
            constexpr uint32_t b1_lbo = ((32u * 16u >> 4) & 0x3fffu) << 16;
            const uint32_t e_lo0 = ((smem_u32(s_e) >> 4) & 0x3fffu) | lbo_plane;
            const uint32_t b1_lo0 = ((smem_u32(s_w1) >> 4) & 0x3fffu) | b1_lbo;
            mbar_wait(&s_wfull, 0);
            // c1 = conv1 row, eb = builder index of its first vertical tap.  Inside a piece eb advances by one per row; a
            // piece of `rows` output rows has rows + 4 conv1 rows and rows + 8 page rows, so eb skips four at the end of a
            // piece (and releases those four ring entries).
            uint32_t c1 = 0, eb = 0, w_efull = 0, w_t1empty = 0;
            SegIter seg1(pos0, pos1, p.h, p.strips);
            while (seg1.next()) {
                for (int i = 0; i < seg1.rows + 4; ++i, ++c1, ++eb) {
                    mbar_wait_cnt(&s_efull[(eb + 4) % G_ER], ((eb + 4) / G_ER) & 1u, w_efull, p.stats);
                    mbar_wait_cnt(&s_t1empty[c1 % G_D1], ((c1 / G_D1) & 1u) ^ 1u, w_t1empty, p.stats);
                    tc_fence_after();
                    const uint32_t d1 = tmem_base + (uint32_t)(G_SLOTS * G_NPAD) + (c1 % G_D1) * 32u;
                    for (int half = 0; half < ((p.dbg & 2) ? 0 : p.halves); ++half) {
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) {
                            const uint32_t slot = (eb + 2 * ks) % G_ER;                  // partner = slot + 1 (entry G_ER mirrors entry 0)
                            tc_mma(d1, e_lo0 + slot * (G_PLANE >> 4), hi, b1_lo0 + (uint32_t)((half * 3 + ks) * 64), hi, IDESC1,
                                   (half | ks) ? 1u : 0u);
                        }
                    }
                    tc_commit(&s_t1full[c1 % G_D1]);
                    tc_commit(&s_eempty[eb % G_ER]);
                }
                for (int j = 0; j < 4; ++j) tc_commit(&s_eempty[(eb + j) % G_ER]);
                eb += 4;
            }
            if (p.stats) { atomicAdd(&g_c12_stats[0], w_efull); atomicAdd(&g_c12_stats[1], w_t1empty); }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== conv2 MMA issuer =====================
        if (elect_one()) {
            uint32_t total = 0;
            {
                SegIter it(pos0, pos1, p.h, p.strips);
                while (it.next()) total += (uint32_t)(it.rows + 4);
            }
            const uint32_t hi = (uint32_t)(make_desc(0, 0, 128) >> 32);
            constexpr uint32_t lbo_plane = ((G_PLANE >> 4) & 0x3fffu) << 16;            // K halves = planes 0, 1 / expanded rows r, r + 1
            constexpr uint32_t lbo_pair2 = ((32u >> 4) & 0x3fffu) << 16;                 // K halves = pair units two pixels apart
            constexpr uint32_t lbo_pair1 = ((16u >> 4) & 0x3fffu) << 16;                 // ... one pixel apart (second half: zero weights)
            constexpr uint32_t b2_lbo = (((uint32_t)G_NF * 16u >> 4) & 0x3fffu) << 16;
            mbar_wait(&s_wfull, 0);
            const uint32_t a_lo0 = (smem_u32(ring) >> 4) & 0x3fffu;
            const uint32_t b2_lo0 = ((smem_u32(s_w2) >> 4) & 0x3fffu) | b2_lbo;
            for (int s = 0; s < G_SLOTS; ++s) mbar_wait(&s_tempty[s], 0);                // every conv2 slot zeroed once
            tc_fence_after();
            uint32_t w_full = 0, w_tempty = 0;
            for (uint32_t kk = 0; kk < total; kk += G_RING) {
#pragma unroll
                for (int u = 0; u < G_RING; ++u) {
                    if (kk + u >= total) break;
                    mbar_wait_cnt(&s_full[u], ((kk + u) / G_RING) & 1u, w_full, p.stats);
                    mbar_wait_cnt(&s_tempty[(u + 4) % G_SLOTS], ((kk + u + 4) / G_SLOTS) & 1u, w_tempty, p.stats);
                    tc_fence_after();
                    uint32_t a_step, b_step, d_step;
                    asm volatile("mov.u32 %0, %3;\n\tmov.u32 %1, %4;\n\tmov.u32 %2, %5;"
                                 : "=r"(a_step), "=r"(b_step), "=r"(d_step) : "r"(a_lo0), "r"(b2_lo0), "r"(tmem_base));
                    const int s0 = u % G_SLOTS;
                    const int n1 = (G_SLOTS - s0) < 5 ? (G_SLOTS - s0) : 5;           // blocks before the ring wraps
                    const uint32_t idesc1 = IDESC0 | ((uint32_t)((n1 * G_NPAD) >> 3) << 17);
                    const uint32_t idesc2 = IDESC0 | ((uint32_t)(((5 - n1) * G_NPAD) >> 3) << 17);
#pragma unroll
                    for (int q = 0; q < G_NMMA; ++q) {
                        if (((p.dbg & 1) && q >= 5) || ((p.dbg & 8) && q >= 1)) break;
                        // q < 5: planes (0, 1) at tap q; q = 5: pair plane, taps 0..3; q = 6: pair plane, tap 4
                        const uint32_t a_off = (uint32_t)(u * (G_ROW_BYTES >> 4)) +
                                               (q < 5 ? (uint32_t)q : (uint32_t)(2 * (G_PLANE >> 4) + (q == 5 ? 0 : 4)));
                        const uint32_t a_lo = (a_step + a_off) | (q < 5 ? lbo_plane : (q == 5 ? lbo_pair2 : lbo_pair1));
                        const uint32_t b_lo = b_step + (uint32_t)(q * (G_WDX_BYTES >> 4));
                        tc_mma(d_step + (uint32_t)(s0 * G_NPAD), a_lo, hi, b_lo, hi, idesc1, 1u);   // slots are pre-zeroed
                        if (n1 < 5) tc_mma(d_step, a_lo, hi, b_lo + (uint32_t)(n1 * G_NPAD), hi, idesc2, 1u);
                    }
                    tc_commit(&s_empty[u]);
                    tc_commit(&s_tfull[s0]);
                }
            }
            if (p.stats) {
                atomicAdd(&g_c12_stats[2], w_full); atomicAdd(&g_c12_stats[3], w_tempty);
                atomicAdd(&g_c12_stats[8], total);
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ===================== conv1 epilogue: accumulator -> conv2's input ring =====================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;                            // pixel slot of the strip patch
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(G_SLOTS * G_NPAD);
        const float inv255 = 1.0f / 255.0f;
        uint32_t c = 0, w_t1full = 0, w_empty = 0;
        SegIter it(pos0, pos1, p.h, p.strips);
        while (it.next()) {
            const int x = it.strip * G_SW - 2 + m;                     // conv1 output column of this slot
            const bool in0 = x >= 0 && x < p.w, in1 = x + 1 >= 0 && x + 1 < p.w;
            for (int i = 0; i < it.rows + 4; ++i, ++c) {
                const int y = it.ys - 2 + i;
                const bool yin = y >= 0 && y < p.h;
                mbar_wait_cnt(&s_t1full[c % G_D1], (c / G_D1) & 1u, w_t1full, p.stats);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld16(t_lane + (c % G_D1) * 32u, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                tmem_ld16(t_lane + (c % G_D1) * 32u + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_t1empty[c % G_D1]);
                uint32_t pk[12];
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    const float a = fmaxf(fmaf(__uint_as_float(v[2 * j]), inv255, p.bias1[2 * j]), 0.f);
                    const float b = fmaxf(fmaf(__uint_as_float(v[2 * j + 1]), inv255, p.bias1[2 * j + 1]), 0.f);
                    const bool keep = yin && (j < 10 ? in0 : in1);     // outside the grid conv2 sees its 'same' zeros
                    pk[j] = keep ? pack2<T>(a, b) : 0u;
                }
                mbar_wait_cnt(&s_empty[c % G_RING], ((c / G_RING) & 1u) ^ 1u, w_empty, p.stats);
                uint8_t* row = ring + (size_t)(c % G_RING) * G_ROW_BYTES + (size_t)m * 16;
                if (!(p.dbg & 4)) {
                    *reinterpret_cast<uint4*>(row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(row + G_PLANE) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    *reinterpret_cast<uint4*>(row + 2 * G_PLANE) = make_uint4(pk[8], pk[9], pk[10], pk[11]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_full[c % G_RING]);
            }
        }
        if (p.stats && warp == 2 && lane == 0) { atomicAdd(&g_c12_stats[4], w_t1full); atomicAdd(&g_c12_stats[5], w_empty); }
    } else if (warp < 10) {
        // ===================== builders: one page row per step into the operand form of conv1 =====================
        const int xq = threadIdx.x - 192;                              // pixel slot 0..127: unit = page pixels gx .. gx + 7
        const int par = xq & 1;
        uint32_t b = 0, w_eempty = 0;
        SegIter it(pos0, pos1, p.h, p.strips);
        auto load = [&](const SegIter& s, int bi, uint8_t& v0, uint8_t& v1) {
            const int gy = s.ys - 4 + bi, gx = s.strip * G_SW - 4 + xq;
            const uint8_t* src = p.img + (size_t)s.page * p.img_h * p.img_w;
            const bool yin = gy >= 0 && gy < p.img_h;
            v0 = (yin && gx >= 0 && gx < p.img_w) ? __ldg(src + (size_t)gy * p.img_w + gx) : (uint8_t)0;
            v1 = (xq < 8 && yin && gx + 128 >= 0 && gx + 128 < p.img_w) ? __ldg(src + (size_t)gy * p.img_w + gx + 128) : (uint8_t)0;
        };
        // the byte of a row is requested three rows before it is converted: one row of work does not cover a load from L2
        SegIter lit(pos0, pos1, p.h, p.strips);
        bool lhave = lit.next();
        int lbi = 0;
        auto request = [&](uint8_t& v0, uint8_t& v1) {
            v0 = 0; v1 = 0;
            if (!lhave) return;
            load(lit, lbi, v0, v1);
            if (++lbi == lit.rows + 8) { lhave = lit.next(); lbi = 0; }
        };
        uint8_t qa0, qa1, qb0, qb1, qc0, qc1;
        request(qa0, qa1); request(qb0, qb1); request(qc0, qc1);
        while (it.next()) {
            const int nb = it.rows + 8;
            for (int bi = 0; bi < nb; ++bi, ++b) {
                const uint8_t c0 = qa0, c1v = qa1;
                qa0 = qb0; qa1 = qb1; qb0 = qc0; qb1 = qc1;
                request(qc0, qc1);
                T* cvt = s_cvt[b & 1];
                const T v = T((float)c0);
                cvt[xq] = v;
                cvt[G_CVT_COPY + xq + 1] = v;
                if (xq < 8) {
                    const T w = T((float)c1v);
                    cvt[128 + xq] = w;
                    cvt[G_CVT_COPY + 128 + xq + 1] = w;
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const uint32_t* win = reinterpret_cast<const uint32_t*>(cvt + par * G_CVT_COPY) + ((xq + par) >> 1);
                const uint4 unit = make_uint4(win[0], win[1], win[2], win[3]);
                mbar_wait_cnt(&s_eempty[b % G_ER], ((b / G_ER) & 1u) ^ 1u, w_eempty, p.stats);
                *reinterpret_cast<uint4*>(s_e + (size_t)(b % G_ER) * G_PLANE + (size_t)xq * 16) = unit;
                if (b % G_ER == 0) *reinterpret_cast<uint4*>(s_e + (size_t)G_ER * G_PLANE + (size_t)xq * 16) = unit;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_efull[b % G_ER]);
            }
        }
        if (p.stats && warp == 6 && lane == 0) atomicAdd(&g_c12_stats[6], w_eempty);
    } else {
        // ===================== conv2 epilogue: drain one slot per output row (as conv_fold.cu) =====================
        const int quarter = warp & 3, group = (warp - 10) >> 2;
        const int j = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        T* out = reinterpret_cast<T*>(p.out);
        T* pool = reinterpret_cast<T*>(p.pool);
        if (group == 0) {
            for (int s = 0; s < G_SLOTS; ++s) { tmem_st16_zero(t_lane + (uint32_t)(s * G_NPAD)); tmem_st16_zero(t_lane + (uint32_t)(s * G_NPAD + 16)); }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                for (int s = 0; s < G_SLOTS; ++s) mbar_arrive(&s_tempty[s]);
        }
        uint32_t g = 0, w_tfull = 0;
        SegIter it(pos0, pos1, p.h, p.strips);
        while (it.next()) {
            const int page = it.page, ys = it.ys, rows = it.rows;
            const int x = it.strip * G_SW + j;
            const bool xok = j < G_SW && x < p.w;
            for (int o2 = 2 * (int)((G_EG + group - (g >> 1) % G_EG) % G_EG); o2 < rows + 4; o2 += 2 * G_EG) {
                uint32_t kept[G_NPAD / 2];
#pragma unroll
                for (int st = 0; st < 2; ++st) {
                    const uint32_t gg = g + (uint32_t)(o2 + st);
                    const uint32_t slot = gg % G_SLOTS;
                    const int y = ys + o2 + st - 4;
                    const bool real = o2 >= 4;
                    mbar_wait_cnt(&s_tfull[slot], (gg / G_SLOTS) & 1u, w_tfull, p.stats);
                    tc_fence_after();
                    const uint32_t tacc = t_lane + slot * G_NPAD;
                    uint32_t v[G_NPAD];
                    tmem_ld16(tacc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                    tmem_ld16(tacc + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                    tmem_ld_wait();
                    tmem_st16_zero(tacc);
                    tmem_st16_zero(tacc + 16u);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_tempty[slot]);      // drained and re-zeroed: free for the window of input row gg + 4
                    if (!real || (p.dbg & 32)) continue;
                    if (p.plog && !(p.dbg & 16)) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int o = 0; o < G_NPAD; ++o) {
                            const float a = __uint_as_float(v[o]);
                            acc.x = fmaf(a, c_skip_lw12[o * G_NC + 0], acc.x);
                            acc.y = fmaf(a, c_skip_lw12[o * G_NC + 1], acc.y);
                            acc.z = fmaf(a, c_skip_lw12[o * G_NC + 2], acc.z);
                        }
                        if (p.plog_nc > 3) {
#pragma unroll
                            for (int o = 0; o < G_NPAD; ++o) acc.w = fmaf(__uint_as_float(v[o]), c_skip_lw12[o * G_NC + 3], acc.w);
                        }
                        if (xok && y < p.h) p.plog[((size_t)page * p.h + y) * p.w + x] = acc;
                    }
#pragma unroll
                    for (int pl = 0; pl < G_NPAD / 8; ++pl) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            pk[i] = pack2<T>(__uint_as_float(v[pl * 8 + 2 * i]) + p.bias2[pl * 8 + 2 * i],
                                             __uint_as_float(v[pl * 8 + 2 * i + 1]) + p.bias2[pl * 8 + 2 * i + 1]);     // conv2 is linear
                        const int oc = pl * 8;
                        if (out && xok && y < p.h && oc < p.out_cp)
                            *reinterpret_cast<uint4*>(out + act_idx(page, p.out_cp, p.h, p.w, oc, y, x)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (pool) {
                            if (st == 0) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) kept[pl * 4 + i] = pk[i];
                            } else {
                                uint32_t pm[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t mx = max2<T>(pk[i], kept[pl * 4 + i]);
                                    pm[i] = max2<T>(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                                }
                                if (!(lane & 1) && xok && y < p.h && oc < p.pool_cp) {
                                    const int ph = p.h >> 1, pw = p.w >> 1;
                                    *reinterpret_cast<uint4*>(pool + act_idx(page, p.pool_cp, ph, pw, oc, y >> 1, x >> 1)) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
                                }
                            }
                        }
                    }
                }
            }
            g += (uint32_t)(rows + 4);
        }
        if (p.stats && quarter == 0 && lane == 0) atomicAdd(&g_c12_stats[7], w_tfull);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

uint16_t to_operand(float v, int precision) {
    if (precision == PCS_PREC_BF16) { __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b); }
    __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h);
}
float from_operand(uint16_t u, int precision) {
    if (precision == PCS_PREC_BF16) { uint32_t x = (uint32_t)u << 16; float f; memcpy(&f, &x, 4); return f; }
    __half_raw hr; hr.x = u; return __half2float(__half(hr));
}

template <typename T>
int launch_fused_t(pcs_ctx* ctx, FusedParams& p) {
    const size_t smem = std::max<size_t>(36864 + 6144 + (size_t)(G_ER + 1) * G_PLANE + (size_t)G_RING * G_ROW_BYTES + 1024, kSoloSmem);
    static bool set[64] = {};
    if (ctx->device >= 64 || !set[ctx->device]) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(conv12_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (ctx->device < 64) set[ctx->device] = true;
    }
    const int grid = (int)((p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta);
    static const bool stats = getenv("PCSEG_C12_STATS") != nullptr;
    p.stats = stats ? 1 : 0;
    { const char* e = getenv("PCSEG_C12_DEBUG"); p.dbg = e ? atoi(e) : 0; }
    if (stats) {
        unsigned long long z[16] = {};
        PCS_CUDA(ctx, cudaMemcpyToSymbol(g_c12_stats, z, sizeof(z)));
    }
    PCS_CUDA(ctx, launch_kernel_pdl(conv12_fused_kernel<T>, dim3(grid), dim3(G_THREADS), smem, ctx->stream, ctx->pdl, p));
    PCS_LAUNCH_CHECK(ctx, "conv12_fused_kernel");
    if (stats) {
        unsigned long long z[16];
        PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PCS_CUDA(ctx, cudaMemcpyFromSymbol(z, g_c12_stats, sizeof(z)));
        const double rows = (double)std::max<unsigned long long>(1, z[8]);
        fprintf(stderr, "[conv12] cycles waited per row: issuer e_full %.1f t1_empty %.1f full %.1f tempty %.1f | epi1 t1_full %.1f ring-empty %.1f | "
                        "builder e_empty %.1f | epi2 (3 groups) tfull %.1f\n",
                z[0] / rows, z[1] / rows, z[2] / rows, z[3] / rows, z[4] / rows, z[5] / rows, z[6] / rows, z[7] / rows);
    }
    return PCS_OK;
}

}  // namespace

bool conv12_fused_supported(int k1, int cout1, int k2, int cin2, int cout2) {
    return k1 == 5 && cout1 == 20 && k2 == 5 && cin2 == 20 && cout2 <= 32;
}

// conv1 operand image [hi|lo][ks][plane][n < 32][e < 8] as in conv1_umma.cu (value W[dy = 2 ks + plane][dx = e][n]) with the
// pixel-pair columns n = 20..23: channel 16 + (n - 20) of the pixel one to the right, i.e. W[dy][dx = e - 1][16 + n - 20].
size_t conv12_weight_image1(const float* w32 /*[25][1][20]*/, int precision, std::vector<uint16_t>& out) {
    out.assign((size_t)2 * 3 * 2 * 32 * 8, 0);
    for (int ks = 0; ks < 3; ++ks)
        for (int pl = 0; pl < 2; ++pl)
            for (int n = 0; n < 24; ++n)
                for (int e = 0; e < 8; ++e) {
                    const int dy = 2 * ks + pl, ch = n < 20 ? n : 16 + (n - 20), dx = n < 20 ? e : e - 1;
                    if (dy >= 5 || dx < 0 || dx >= 5) continue;
                    const float w = w32[(size_t)(dy * 5 + dx) * 20 + ch];
                    const uint16_t hi = to_operand(w, precision);
                    const uint16_t lo = to_operand(w - from_operand(hi, precision), precision);
                    const size_t idx = (((size_t)ks * 2 + pl) * 32 + n) * 8 + e;
                    out[idx] = hi;
                    out[(size_t)3 * 2 * 32 * 8 + idx] = lo;
                }
    return out.size() * sizeof(uint16_t);
}

// conv2 resident operand image [K step q < 7][K half][row = (4 - dy) * 32 + o][8]:
//   q < 5: input channels 0..15 (half = plane) at horizontal tap q;
//   q = 5: half 0 = channels 16..19 at taps (0, 1), half 1 = at taps (2, 3);   q = 6: half 0 = tap 4 (and a zero tap), half 1 zero.
size_t conv12_weight_image2(const float* w32 /*[25][20][cout]*/, int cout, int precision, std::vector<uint16_t>& out) {
    out.assign((size_t)G_NMMA * 2 * G_NF * 8, 0);
    for (int q = 0; q < G_NMMA; ++q)
        for (int half = 0; half < 2; ++half)
            for (int dy = 0; dy < 5; ++dy)
                for (int o = 0; o < cout; ++o)
                    for (int e = 0; e < 8; ++e) {
                        int ci, dx;
                        if (q < 5) { ci = half * 8 + e; dx = q; }
                        else {
                            ci = 16 + (e & 3);
                            dx = (q == 5 ? 2 * half : 4 + 2 * half) + (e >> 2);
                            if (dx >= 5) continue;
                        }
                        const float v = w32[((size_t)(dy * 5 + dx) * 20 + ci) * cout + o];
                        out[(((size_t)q * 2 + half) * G_NF + (4 - dy) * G_NPAD + o) * 8 + e] = to_operand(v, precision);
                    }
    return out.size() * sizeof(uint16_t);
}

int launch_conv12_fused(pcs_ctx* ctx, const Conv12Args& a) {
    if ((a.h & 3) || (a.w & 1)) return set_err(ctx, PCS_ERR_ARG, "conv12_fused: grid %dx%d must be a multiple of 4 x 2", a.h, a.w);
    FusedParams p{};
    p.img = a.d_image; p.img_h = a.img_h; p.img_w = a.img_w; p.n = a.n; p.h = a.h; p.w = a.w;
    p.w1img = reinterpret_cast<const uint8_t*>(a.w1img); p.w2img = reinterpret_cast<const uint8_t*>(a.w2img);
    for (int i = 0; i < 24; ++i) p.bias1[i] = i < 20 ? a.h_bias1[i] : a.h_bias1[16 + (i - 20)];
    for (int i = 0; i < 32; ++i) p.bias2[i] = i < a.cout2 ? a.h_bias2[i] : 0.f;
    static const int split_env = [] { const char* e = getenv("PCSEG_C1_SPLIT"); return e ? atoi(e) : -1; }();
    const bool bf = ctx->precision == PCS_PREC_BF16;
    p.halves = split_env >= 0 ? (split_env ? 2 : 1) : (bf ? 2 : 1);
    p.out = a.out; p.out_cp = a.out_cp; p.pool = a.pool_out; p.pool_cp = a.pool_cp;
    p.plog = nullptr;
    if (a.plog) {
        if (!a.skip_lw) return set_err(ctx, PCS_ERR_ARG, "conv12_fused: partial logits need the logits rows");
        p.plog = reinterpret_cast<float4*>(a.plog);
        p.plog_nc = ctx->n_classes;
        if (ctx->device >= 64 || g_skip12_owner[ctx->device] != ctx->model_stamp) {
            PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_skip_lw12, a.skip_lw, sizeof(float) * G_LOGC * G_NC, 0, cudaMemcpyDeviceToDevice, ctx->stream));
            if (ctx->device < 64) g_skip12_owner[ctx->device] = ctx->model_stamp;
        }
    }
    p.strips = (a.w + G_SW - 1) / G_SW;
    p.total_rows = (long long)a.n * p.strips * a.h;
    p.rows_per_cta = ((p.total_rows + ctx->sm_count - 1) / ctx->sm_count + 3) / 4 * 4;
    if (p.rows_per_cta < 16) p.rows_per_cta = 16;
    if (bf) return launch_fused_t<__nv_bfloat16>(ctx, p);
    return launch_fused_t<__half>(ctx, p);
}

}  // namespace pcs
