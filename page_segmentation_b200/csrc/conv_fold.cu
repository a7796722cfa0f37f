// "Marching, dy-folded" implicit-GEMM 5x5 'same' convolution for the small-channel encoder layers
// (conv2 ... conv5 of ocr4all_pixel_classifier/lib/model.py:52-60 / :212-217).
//
// With C_out of 30-60 the plain kernel (conv_umma.cu) issues one N=32..64 MMA per tap and re-reads the
// 4 KB A operand from shared memory for each of them: it is shared-memory-read bound at ~40 % tensor
// activity.  Here the five VERTICAL taps are folded into the N dimension.  An input row rho contributes
// to the five output rows y = rho+2-dy, so TMEM holds a ring of 8 output-row accumulators (NPAD columns
// each) and ONE MMA of N' = 5*NPAD adds
//     D[slot(y)][j][o] += sum_c in[rho][x0-2+j+dx][c] * W[dy = rho+2-y][dx][c][o]      for the 5 rows y at once
// (A = the input row shifted by dx, B = [W[4] | W[3] | W[2] | W[1] | W[0]] stacked along N; when the
// 5-slot window wraps around the ring the MMA is split in two).  A is read once per 5 taps, the MMA is
// tensor-bound, the epilogue stays as cheap as in the plain kernel (it drains ONE NPAD-column slot per
// output row, same TMEM lane, no cross-lane traffic) and re-zeroes the slot for its next use.  Further:
//   * all weights of the layer stay resident in shared memory (51-154 KB), loaded once per CTA;
//   * a CTA marches down a 124-pixel strip: every output row needs ONE new input row (one TMA box holding
//     all channel chunks), halo re-read 132/128 instead of 12/8;
//   * the accumulator ring decouples the epilogue from the MMAs by up to three rows;
//   * segments start with four "virtual" output rows (their slots collect the partial sums left over by
//     the previous segment); they are drained and re-zeroed like real rows but never stored.
#include "common.cuh"
#include "umma_ptx.cuh"

namespace pcs {
namespace {
using namespace ptx;

// element-wise maximum of two packed operand pairs (max of rounded values == rounded max: rounding is monotonic)
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
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// NPAD columns of one accumulator slot: whole 16-column pieces, then one of 8 (NPAD is a multiple of 8)
template <int NPAD> __device__ __forceinline__ void tmem_ld_slot(uint32_t taddr, uint32_t (&v)[NPAD]) {
#pragma unroll
    for (int c = 0; c + 16 <= NPAD; c += 16) tmem_ld16(taddr + (uint32_t)c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
    if constexpr (NPAD % 16 == 8) tmem_ld8(taddr + (uint32_t)(NPAD - 8), *reinterpret_cast<uint32_t(*)[8]>(&v[NPAD - 8]));
}
template <int NPAD> __device__ __forceinline__ void tmem_zero_slot(uint32_t taddr) {
#pragma unroll
    for (int c = 0; c + 16 <= NPAD; c += 16) tmem_st16_zero(taddr + (uint32_t)c);
    if constexpr (NPAD % 16 == 8) tmem_st8_zero(taddr + (uint32_t)(NPAD - 8));
}
constexpr int fold_up16(int v) { return (v + 15) / 16 * 16; }

// threads = warp 0 producer, warp 1 MMA issuer, then EG epilogue groups of 4 warps (4 TMEM lane quarters)
constexpr int F_RING_MAX = 20;              // input-row ring entries (one row is consumed per step)
constexpr int F_SLOTS_MAX = 16;             // output-row accumulator slots in TMEM (512 columns / NPAD)
constexpr int F_SW = 124;                   // valid output pixels per strip

// conv2 -> logits partial sums (fcn_skip): concat[deconv5, conv2] -> logits 1x1 (model.py:85-88) is linear
// in conv2, so its share  P[y,x,k] = sum_o conv2[y,x,o] * lw[20+o][k]  is taken here from the fp32
// accumulators (before the bf16 rounding) and the full-resolution conv2 tensor is never written.
constexpr int F_LOGC = 32, F_NC = 4;
__constant__ float4 c_skip_lw[F_LOGC];       // [channel] -> the logits weights of (up to) four classes: ONE 128-bit constant load per channel
int64_t g_skip_owner[64] = {0};

struct FoldParams {
    int n, h, w;
    int strips;                             // 124-pixel strips per page
    long long total_rows, rows_per_cta;     // the (page, strip, row) space is cut into equal contiguous ranges, one per CTA:
                                            // perfect balance, and a strip restart (4 virtual rows) only where a range or a
                                            // strip begins
    const uint8_t* wimg;                    // [chunk][dx][plane][row = (4-dy)*NPAD + o][8]: resident operand image
    float bias[64];                         // by value (zero padded): constant-bank operands of the epilogue's adds
    int cout, relu;
    void* out; int out_cp;
    void* pool; int pool_cp;
    int out_c0;                             // first output channel of this launch (a multiple of 8): layers split along N
    float4* psum_out;                       // [n][NPAD/4][h][w] fp32 partial sums written INSTEAD of the output (K split, first part)
    const float4* psum_in;                  // ... added before bias / activation (K split, last part)
    float4* plog;                           // [n][h][w] partial logits (4 classes, zero padded) or null
    int plog_nc;                            // classes actually present (the 4th FMA chain is skipped for <= 3)
    uint32_t w_bytes;
    int ring;                               // input-row ring depth
};

// K steps of one input row (NPL = 8-channel planes of the source).  Plane pairs (2q, 2q+1) give one
// K=16 MMA per dx (the two K halves are the two planes, LBO = plane pitch).  An odd last plane is paired
// with ITSELF one pixel further: the K halves of one MMA are the taps (dx, dx+1) of that plane, LBO = 16 B
// (one pixel), so 5 taps x 8 channels cost 3 MMAs -- (0,1), (2,3), (3,4) with zero weights for the
// repeated tap 3 -- instead of the 5 that padding the plane count to even would cost.
constexpr int fold_gcd(int a, int b) { return b == 0 ? a : fold_gcd(b, a % b); }

//
// PX: the odd last plane holds at most FOUR channels and arrives as pixel-pair units (unit i = [channels of pixel i - 1 |
// of pixel i], w + 1 units per row, conv1_umma.cu): one 16-byte K half then covers two taps, and the five taps cost TWO
// MMAs -- K halves (0,1) | (2,3) two units apart, then (3,4) | (4,5) with zero weights for everything but tap 4 (every
// unit read lies inside the row's box, so the zero weights meet finite values) -- instead of three.
//
// KS: kernel size, 5 (the FCN layers) or 3 (the U-Net's 64-channel full-resolution layers: N' = 3 * 64 = 192 instead of one
// N = 64 MMA per tap, whose fixed cost caps the plain kernel at 56 % tensor activity); the window of a row is KS slots, a
// strip starts with KS - 1 virtual rows, the halo is KS / 2.  SRC2: the planes are the concatenation of two tensors
// (conv9a: [conv1b, up9]), each loaded with its own tensor map into the same ring entry.
template <int NPL, bool PX = false, int KS = 5> struct FoldK {
    static constexpr int PAIRS = NPL / 2, ODD = NPL & 1;
    static constexpr int NMMA = PAIRS * KS + ODD * (PX ? 2 : 3);
    static_assert(KS == 5 || ODD == 0, "the self-paired odd plane is laid out for five taps");
};

template <typename T, int NPAD, int NPL, int RING, int SLOTS, int EG, bool PX, int KS, bool SRC2>
__global__ void __launch_bounds__(64 + EG * 128, 1)
conv_fold_kernel(const FoldParams p, const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2) {
    static_assert(!PX || (NPL & 1), "the pixel-pair plane is the odd last plane");
    static_assert((KS == 5 || KS == 3) && !(PX && SRC2) && (!SRC2 || (NPL & 1) == 0), "kernel size / source layout");
    // Folded N.  NPAD need not be a multiple of 16 (C_out = 40: five whole planes): the five slots of a window are then
    // 5 * NPAD = 200 accumulator columns and every MMA rounds ITS part of the window up to a legal N.  The extra columns
    // of a part that ends inside the ring belong to the slot after the window and take zero weight rows (the weight
    // image is padded to NF rows): D += 0 there, which is why the issuer also waits for THAT slot to be drained and
    // re-zeroed (a read-modify-write of the tensor pipe racing with the epilogue's zeroing store would undo it).  The
    // extra columns of a part that ends at the ring's end fall behind the ring (columns SLOTS * NPAD ...: unused).
    constexpr int NF = fold_up16(KS * NPAD);
    constexpr bool SPILL = (NPAD % 16) != 0;
    constexpr uint32_t ROW_BYTES = NPL * 2048;                       // one ring entry: all planes of one input row
    constexpr uint32_t WDX_BYTES = 2 * NF * 16;                      // weights of one K step
    constexpr int NMMA = FoldK<NPL, PX, KS>::NMMA;
    constexpr uint32_t IDESC0 = (1u << 4) | ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 7) |
                                ((std::is_same<T, __nv_bfloat16>::value ? 1u : 0u) << 10) | ((uint32_t)(128 >> 4) << 24);
    static_assert(NF <= 256 && NF % 16 == 0, "folded N must be a legal UMMA N");
    static_assert(SLOTS * NPAD + (SPILL ? 16 : 0) <= 512 && SLOTS <= F_SLOTS_MAX && SLOTS >= 8, "accumulator ring must fit TMEM");
    static_assert(NPAD % 8 == 0, "whole planes");
    constexpr int PERIOD = RING / fold_gcd(RING, SLOTS) * SLOTS;     // least common multiple
    static_assert(PERIOD % RING == 0 && PERIOD % SLOTS == 0 && PERIOD <= 20 && RING <= F_RING_MAX,
                  "the issue loop is unrolled over one common period of the input ring and the accumulator ring");

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[F_RING_MAX], s_empty[F_RING_MAX], s_wfull, s_tfull[F_SLOTS_MAX], s_tempty[F_SLOTS_MAX];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = base;
    uint8_t* ring = base + ((p.w_bytes + 1023) / 1024) * 1024;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < RING; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        mbar_init(&s_wfull, 1);
        for (int s = 0; s < SLOTS; ++s) { mbar_init(&s_tfull[s], 1); mbar_init(&s_tempty[s], 4); }
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
    if (threadIdx.x != 0) griddep_wait();          // thread 0 (the producer) first queues the weight loads, see below
    const uint32_t tmem_base = s_tmem_base;
    // this CTA's range of the concatenated (page * strips + strip) * h + y space; segments = its pieces inside one strip
    const long long pos0 = (long long)blockIdx.x * p.rows_per_cta;
    const long long pos1 = pos0 + p.rows_per_cta < p.total_rows ? pos0 + p.rows_per_cta : p.total_rows;

    if (warp == 0) {
        // ===================== producer: resident weights, then the input-row stream =====================
        if (lane == 0) {
            mbar_expect_tx(&s_wfull, p.w_bytes);
            for (uint32_t off = 0; off < p.w_bytes; off += WDX_BYTES) bulk_load(s_w + off, p.wimg + off, WDX_BYTES, &s_wfull);
            griddep_wait();                         // weights are constants; the input rows are the previous layer's output
            uint32_t k = 0;                                           // running input-row counter
            for (long long pos = pos0; pos < pos1;) {
                const int sg = (int)(pos / p.h), ys = (int)(pos - (long long)sg * p.h);
                const int rows = (int)((long long)(p.h - ys) < pos1 - pos ? (long long)(p.h - ys) : pos1 - pos);
                const int page = sg / p.strips, strip = sg - page * p.strips;
                pos += rows;
                const int x0 = strip * F_SW - KS / 2;
                for (int i = 0; i < rows + KS - 1; ++i, ++k) {
                    const uint32_t slot = k % (uint32_t)RING, pass = k / (uint32_t)RING;
                    mbar_wait(&s_empty[slot], (pass & 1u) ^ 1u);
                    mbar_expect_tx(&s_full[slot], ROW_BYTES);
                    // box = 256 u64 (128 px x 16 B) x 1 row x NPL planes (PX: the whole planes, then the pair units from one
                    // unit further: unit j + dx of the box = pixels (x - 2 + dx, x - 1 + dx) of output pixel x = strip start + j)
                    tma_load_4d(ring + (size_t)slot * ROW_BYTES, &tm, &s_full[slot], x0 * 2, ys - KS / 2 + i, 0, page);
                    if constexpr (PX)
                        tma_load_4d(ring + (size_t)slot * ROW_BYTES + (NPL - 1) * 2048, &tm2, &s_full[slot], (x0 + 1) * 2, ys - KS / 2 + i, 0, page);
                    if constexpr (SRC2)
                        tma_load_4d(ring + (size_t)slot * ROW_BYTES + (NPL / 2) * 2048, &tm2, &s_full[slot], x0 * 2, ys - KS / 2 + i, 0, page);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // step k (one input row) accumulates into the output slots k .. k+4 (mod 8) and completes slot k.
        // The issuer does not care about work items, only about the number of input rows of this CTA, and
        // the loop is unrolled over one ring period so that every shared-memory / TMEM address and the
        // split of the window at the ring wrap are compile-time immediates: one thread must keep the
        // tensor pipe fed, ~200 uniform-datapath instructions per row (the first version) did not.
        if (elect_one()) {
            uint32_t total = 0;
            for (long long pos = pos0; pos < pos1;) {
                const int ys = (int)(pos % p.h);
                const int rows = (int)((long long)(p.h - ys) < pos1 - pos ? (long long)(p.h - ys) : pos1 - pos);
                pos += rows;
                total += (uint32_t)(rows + KS - 1);
            }
            const uint32_t hi = (uint32_t)(make_desc(0, 0, 128) >> 32);
            constexpr uint32_t a_lbo_pair = ((2048u >> 4) & 0x3fffu) << 16;               // K halves = two planes
            constexpr uint32_t a_lbo_self = ((16u >> 4) & 0x3fffu) << 16;                 // K halves = two taps of one plane
            constexpr uint32_t a_lbo_two = ((32u >> 4) & 0x3fffu) << 16;                  // K halves = pair units two pixels apart
            constexpr uint32_t b_lbo = (((uint32_t)NF * 16u >> 4) & 0x3fffu) << 16;
            mbar_wait(&s_wfull, 0);
            const uint32_t a_lo0 = (smem_u32(ring) >> 4) & 0x3fffu;
            const uint32_t b_lo0 = ((smem_u32(s_w) >> 4) & 0x3fffu) | b_lbo;
            for (int s = 0; s < SLOTS; ++s) mbar_wait(&s_tempty[s], 0);                    // every slot zeroed once
            tc_fence_after();
            for (uint32_t kk = 0; kk < total; kk += PERIOD) {
#pragma unroll
                for (int u = 0; u < PERIOD; ++u) {
                    if (kk + u >= total) break;
                    mbar_wait(&s_full[u % RING], ((kk + u) / RING) & 1u);
                    // the newest output slot of the window must have been drained and re-zeroed by the epilogue:
                    // tempty phase 0 = initial zeroing, phase n = drain of use n-1: use n waits for phase n
                    mbar_wait(&s_tempty[(u + KS - 1) % SLOTS], ((kk + u + KS - 1) / SLOTS) & 1u);
                    if constexpr (SPILL) mbar_wait(&s_tempty[(u + KS) % SLOTS], ((kk + u + KS) / SLOTS) & 1u);
                    tc_fence_after();
                    // opaque copies: keeps the descriptor arithmetic (base + immediate) next to its MMA instead of
                    // having every one of the RING x NCH x 5 sums hoisted out of the loop into spilled registers
                    uint32_t a_step, b_step, d_step;
                    asm volatile("mov.u32 %0, %3;\n\tmov.u32 %1, %4;\n\tmov.u32 %2, %5;"
                                 : "=r"(a_step), "=r"(b_step), "=r"(d_step) : "r"(a_lo0), "r"(b_lo0), "r"(tmem_base));
                    const int s0 = u % SLOTS;
                    const int n1 = (SLOTS - s0) < KS ? (SLOTS - s0) : KS;         // blocks before the ring wraps
                    const uint32_t idesc1 = IDESC0 | ((uint32_t)(fold_up16(n1 * NPAD) >> 3) << 17);
                    const uint32_t idesc2 = IDESC0 | ((uint32_t)(fold_up16((KS - n1) * NPAD) >> 3) << 17);
#pragma unroll
                    for (int q = 0; q < NMMA; ++q) {
                        // K step q: plane pair (q / 5) at tap q % 5, or the odd last plane at taps {0, 2, 3} (+1)
                        const bool pair = q < FoldK<NPL>::PAIRS * KS;
                        const int t = q - FoldK<NPL>::PAIRS * KS;
                        const int plane = pair ? 2 * (q / KS) : NPL - 1;
                        const int dx = pair ? q % KS : PX ? (t == 0 ? 0 : 3) : (t == 0 ? 0 : t + 1);
                        const uint32_t a_lo = (a_step + (uint32_t)((u % RING) * (ROW_BYTES >> 4) + plane * (2048 >> 4) + dx)) |
                                              (pair ? a_lbo_pair : (PX && t == 0) ? a_lbo_two : a_lbo_self);
                        const uint32_t b_lo = b_step + (uint32_t)(q * (WDX_BYTES >> 4));
                        tc_mma(d_step + (uint32_t)(s0 * NPAD), a_lo, hi, b_lo, hi, idesc1, 1u);   // slots are pre-zeroed
                        if (n1 < KS) tc_mma(d_step, a_lo, hi, b_lo + (uint32_t)(n1 * NPAD), hi, idesc2, 1u);
                    }
                    tc_commit(&s_empty[u % RING]);                                        // input row consumed
                    tc_commit(&s_tfull[s0]);                                              // output slot k is complete
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: drain one slot per output row =====================
        const int quarter = warp & 3, group = (warp - 2) >> 2;
        const int j = quarter * 32 + lane;                           // patch column = output pixel of the strip
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        T* out = reinterpret_cast<T*>(p.out);
        T* pool = reinterpret_cast<T*>(p.pool);
        // group 0 zeroes the whole accumulator ring once
        if (group == 0) {
            for (int s = 0; s < SLOTS; ++s) tmem_zero_slot<NPAD>(t_lane + (uint32_t)(s * NPAD));
            if constexpr (SPILL) tmem_st16_zero(t_lane + (uint32_t)(SLOTS * NPAD));       // defined values behind the ring
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                for (int s = 0; s < SLOTS; ++s) mbar_arrive(&s_tempty[s]);            // completes phase 0 of every slot
        }
        uint32_t g = 0;                                               // running output counter (incl. virtual rows), even
        for (long long pos = pos0; pos < pos1;) {
            const int sg = (int)(pos / p.h), ys = (int)(pos - (long long)sg * p.h);
            const int rows = (int)((long long)(p.h - ys) < pos1 - pos ? (long long)(p.h - ys) : pos1 - pos);
            const int page = sg / p.strips, strip = sg - page * p.strips;
            pos += rows;
            const int x = strip * F_SW + j;
            const bool xok = j < F_SW && x < p.w;
            // row pairs (o = o2 - 4 + st) go round-robin over the epilogue groups, across work items
            for (int o2 = 2 * (int)((EG + group - (g >> 1) % EG) % EG); o2 < rows + KS - 1; o2 += 2 * EG) {
                uint32_t kept[NPAD / 2];
#pragma unroll
                for (int st = 0; st < 2; ++st) {
                    const uint32_t gg = g + (uint32_t)(o2 + st);
                    const uint32_t slot = gg % SLOTS;
                    const int y = ys + o2 + st - (KS - 1);            // < ys: virtual row
                    const bool real = o2 >= KS - 1;
                    // K split, last part: what the first part left is requested BEFORE the wait for this row's accumulator
                    // (the epilogue of a tensor-bound layer spends most of its time in that wait)
                    float4 ps[KS == 5 ? NPAD / 4 : 1];
                    if constexpr (KS == 5) if (p.psum_in && real) {
                        const bool ok = xok && y < p.h;
#pragma unroll
                        for (int c4 = 0; c4 < NPAD / 4; ++c4)
                            ps[c4] = ok ? __ldg(p.psum_in + (((size_t)page * (NPAD / 4) + c4) * p.h + y) * p.w + x)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    mbar_wait(&s_tfull[slot], (gg / SLOTS) & 1u);
                    tc_fence_after();
                    const uint32_t tacc = t_lane + slot * NPAD;
                    uint32_t v[NPAD];
                    tmem_ld_slot<NPAD>(tacc, v);
                    tmem_ld_wait();
                    tmem_zero_slot<NPAD>(tacc);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_tempty[slot]);      // slot free for the window of input row gg+4
                    if (!real) continue;
                    if constexpr (KS == 5) if (p.psum_out) {                // K split, first part: raw fp32 sums, nothing else
                        if (xok && y < p.h) {
#pragma unroll
                            for (int c4 = 0; c4 < NPAD / 4; ++c4)
                                p.psum_out[(((size_t)page * (NPAD / 4) + c4) * p.h + y) * p.w + x] =
                                    make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]),
                                                __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
                        }
                        continue;
                    }
                    if constexpr (KS == 5) if (p.psum_in) {                 // K split, last part: add what the first part left
#pragma unroll
                        for (int c4 = 0; c4 < NPAD / 4; ++c4) {
                            v[4 * c4] = __float_as_uint(__uint_as_float(v[4 * c4]) + ps[c4].x);
                            v[4 * c4 + 1] = __float_as_uint(__uint_as_float(v[4 * c4 + 1]) + ps[c4].y);
                            v[4 * c4 + 2] = __float_as_uint(__uint_as_float(v[4 * c4 + 2]) + ps[c4].z);
                            v[4 * c4 + 3] = __float_as_uint(__uint_as_float(v[4 * c4 + 3]) + ps[c4].w);
                        }
                    }
                    const bool rowok = xok && y < p.h;
                    if constexpr (NPAD == F_LOGC) {
                        if (p.plog) {
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int o = 0; o < NPAD; ++o) {
                                const float a = __uint_as_float(v[o]);
                                const float4 lw = c_skip_lw[o];
                                acc.x = fmaf(a, lw.x, acc.x);
                                acc.y = fmaf(a, lw.y, acc.y);
                                acc.z = fmaf(a, lw.z, acc.z);
                            }
                            if (p.plog_nc > 3) {
#pragma unroll
                                for (int o = 0; o < NPAD; ++o) acc.w = fmaf(__uint_as_float(v[o]), c_skip_lw[o].w, acc.w);
                            }
                            if (rowok) p.plog[((size_t)page * p.h + y) * p.w + x] = acc;
                        }
                    }
                    // bias, activation, pack: one 8-channel plane = one 16-byte unit per pixel.  The activation switch is uniform:
                    // hoisted out of the element loop (a select per element was 6 % of conv2's instructions)
                    uint32_t pk[NPAD / 2];
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < NPAD / 2; ++i)
                            pk[i] = pack2<T>(fmaxf(__uint_as_float(v[2 * i]) + p.bias[2 * i], 0.f), fmaxf(__uint_as_float(v[2 * i + 1]) + p.bias[2 * i + 1], 0.f));
                    } else {
#pragma unroll
                        for (int i = 0; i < NPAD / 2; ++i)
                            pk[i] = pack2<T>(__uint_as_float(v[2 * i]) + p.bias[2 * i], __uint_as_float(v[2 * i + 1]) + p.bias[2 * i + 1]);
                    }
                    if (out && rowok) {
                        T* o0 = out + act_idx(page, p.out_cp, p.h, p.w, p.out_c0, y, x);
                        const size_t plane = (size_t)p.h * p.w * 8;
#pragma unroll
                        for (int pl = 0; pl < NPAD / 8; ++pl)
                            if (p.out_c0 + pl * 8 < p.out_cp)
                                *reinterpret_cast<uint4*>(o0 + pl * plane) = make_uint4(pk[4 * pl], pk[4 * pl + 1], pk[4 * pl + 2], pk[4 * pl + 3]);
                    }
                    if (pool) {
                        if (st == 0) {
#pragma unroll
                            for (int i = 0; i < NPAD / 2; ++i) kept[i] = pk[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < NPAD / 2; ++i) {
                                const uint32_t m = max2<T>(pk[i], kept[i]);
                                pk[i] = max2<T>(m, __shfl_xor_sync(0xffffffffu, m, 1));
                            }
                            if (!(lane & 1) && rowok) {
                                const int ph = p.h >> 1, pw = p.w >> 1;
                                T* q0 = pool + act_idx(page, p.pool_cp, ph, pw, p.out_c0, y >> 1, x >> 1);
                                const size_t plane = (size_t)ph * pw * 8;
#pragma unroll
                                for (int pl = 0; pl < NPAD / 8; ++pl)
                                    if (p.out_c0 + pl * 8 < p.pool_cp)
                                        *reinterpret_cast<uint4*>(q0 + pl * plane) = make_uint4(pk[4 * pl], pk[4 * pl + 1], pk[4 * pl + 2], pk[4 * pl + 3]);
                            }
                        }
                    }
                }
            }
            g += (uint32_t)(rows + KS - 1);
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

template <typename T, int NPAD, int NPL, int RING, int SLOTS, int EG, bool PX = false, int KS = 5, bool SRC2 = false>
int launch_fold_t(pcs_ctx* ctx, const FoldConvArgs& a) {
    constexpr int NF = fold_up16(KS * NPAD);
    FoldParams p{};
    p.n = a.n; p.h = a.h; p.w = a.w;
    p.wimg = reinterpret_cast<const uint8_t*>(a.wimg); p.cout = a.cout; p.relu = a.relu;
    static_assert(NPAD <= 64, "bias travels in the parameter block");
    for (int i = 0; i < 64; ++i) p.bias[i] = i < a.cout ? a.h_bias[i] : 0.f;
    if (KS != 5 && (a.psum_in || a.psum_out || a.plog)) return set_err(ctx, PCS_ERR_ARG, "conv_fold: partial sums / logits share are for the 5x5 kernels");
    p.out = a.out; p.out_cp = a.out_cp; p.pool = a.pool_out; p.pool_cp = a.pool_cp;
    p.out_c0 = a.o0;
    p.psum_out = reinterpret_cast<float4*>(a.psum_out);
    p.psum_in = reinterpret_cast<const float4*>(a.psum_in);
    if (a.o0 & 7) return set_err(ctx, PCS_ERR_ARG, "conv_fold: output channel offset %d is not a whole plane", a.o0);
    p.plog = nullptr;
    if (a.plog) {
        if (NPAD != F_LOGC || !a.skip_lw) return set_err(ctx, PCS_ERR_ARG, "conv_fold: partial logits need the N=32 kernel and weights");
        p.plog = reinterpret_cast<float4*>(a.plog);
        p.plog_nc = ctx->n_classes;
        if (ctx->device >= 64 || g_skip_owner[ctx->device] != ctx->model_stamp) {
            PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_skip_lw, a.skip_lw, sizeof(float) * F_LOGC * F_NC, 0, cudaMemcpyDeviceToDevice, ctx->stream));
            if (ctx->device < 64) g_skip_owner[ctx->device] = ctx->model_stamp;
        }
    }
    p.w_bytes = (uint32_t)FoldK<NPL, PX, KS>::NMMA * 2 * NF * 16;  // [K step][K half][N' rows][16 B]
    p.strips = (a.w + F_SW - 1) / F_SW;
    // equal ranges of whole row quads (row pairs for the pooling, pair rotation over the epilogue groups)
    p.total_rows = (long long)a.n * p.strips * a.h;
    p.rows_per_cta = ((p.total_rows + ctx->sm_count - 1) / ctx->sm_count + 3) / 4 * 4;
    if (p.rows_per_cta < 8) p.rows_per_cta = 8;           // 8 + 4 virtual rows per CTA at least: an 8-page launch of the 1/8-resolution layers still fills the SMs
    if ((a.h & 3) || (a.w & 1)) return set_err(ctx, PCS_ERR_ARG, "conv_fold: grid %dx%d must be a multiple of 4 x 2", a.h, a.w);
    EncodeTiledFn enc = fold_get_encode();
    if (!enc) return set_err(ctx, PCS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t planes = (cuuint64_t)a.src.cp / 8;
    if (SRC2 != (a.src2.p != nullptr) || (SRC2 && (a.src2.cp != a.src.cp || (int)planes * 2 != NPL)))
        return set_err(ctx, PCS_ERR_ARG, "conv_fold: two sources of %d planes each expected", NPL / 2);
    if ((!SRC2 && (int)planes != NPL - (PX ? 1 : 0)) || PX != (a.pair_src != nullptr))
        return set_err(ctx, PCS_ERR_ARG, "conv_fold: source has %d planes%s, kernel expects %d", (int)planes, a.pair_src ? " + pair units" : "", NPL);
    const cuuint64_t dims[4] = {(cuuint64_t)a.w * 2, (cuuint64_t)a.h, planes, (cuuint64_t)a.n};
    const cuuint64_t strides[3] = {(cuuint64_t)a.w * 16, (cuuint64_t)a.h * a.w * 16, planes * a.h * a.w * 16};
    const cuuint32_t box[4] = {256, 1, (cuuint32_t)planes, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a.src.p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(ctx, PCS_ERR_CUDA, "conv_fold: cuTensorMapEncodeTiled failed with %d", (int)r);
    CUtensorMap tm2 = tm;
    if (SRC2) {     // the second tensor of the concatenation: same geometry
        r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a.src2.p), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_err(ctx, PCS_ERR_CUDA, "conv_fold: cuTensorMapEncodeTiled (second source) failed with %d", (int)r);
    }
    if (PX) {       // the pair units: w + 1 per row, one "plane"
        const cuuint64_t pdims[4] = {(cuuint64_t)(a.w + 1) * 2, (cuuint64_t)a.h, 1, (cuuint64_t)a.n};
        const cuuint64_t pstrides[3] = {(cuuint64_t)(a.w + 1) * 16, (cuuint64_t)a.h * (a.w + 1) * 16, (cuuint64_t)a.h * (a.w + 1) * 16};
        const cuuint32_t pbox[4] = {256, 1, 1, 1};
        r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(a.pair_src), pdims, pstrides, pbox, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_err(ctx, PCS_ERR_CUDA, "conv_fold: cuTensorMapEncodeTiled (pair units) failed with %d", (int)r);
    }
    const size_t w_al = ((p.w_bytes + 1023) / 1024) * 1024;
    p.ring = RING;
    const size_t smem = std::max<size_t>(w_al + (size_t)RING * NPL * 2048 + 1024, kSoloSmem);
    if (smem + 2 * 1024 > 227 * 1024) return set_err(ctx, PCS_ERR_ARG, "conv_fold: %zu bytes of shared memory needed", smem);
    static size_t attr_set[64] = {};             // the attribute is per device
    if (ctx->device >= 64 || attr_set[ctx->device] < smem) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(conv_fold_kernel<T, NPAD, NPL, RING, SLOTS, EG, PX, KS, SRC2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (ctx->device < 64) attr_set[ctx->device] = smem;
    }
    const int grid = (int)((p.total_rows + p.rows_per_cta - 1) / p.rows_per_cta);
    PCS_CUDA(ctx, launch_kernel_pdl(conv_fold_kernel<T, NPAD, NPL, RING, SLOTS, EG, PX, KS, SRC2>, dim3(grid), dim3(64 + EG * 128), smem, ctx->stream,
                                    ctx->pdl, p, tm, tm2));
    PCS_LAUNCH_CHECK(ctx, "conv_fold_kernel");
    return PCS_OK;
}

template <typename T>
int launch_fold_dispatch(pcs_ctx* ctx, const FoldConvArgs& a) {
    const int key = a.npad * 10 + a.nplanes;
    if (a.k == 3) {         // U-Net, 64 output channels at full resolution: N' = 192, eight slots of 64 columns, two epilogue groups
        if (a.npad == 64 && a.nplanes == 8 && !a.src2.p) return launch_fold_t<T, 64, 8, 4, 8, 2, false, 3, false>(ctx, a);
        if (a.npad == 64 && a.nplanes == 16 && a.src2.p) return launch_fold_t<T, 64, 16, 2, 8, 2, false, 3, true>(ctx, a);
        if (a.npad == 64 && a.nplanes == 16) return launch_fold_t<T, 64, 16, 2, 8, 2, false, 3, false>(ctx, a);
        return set_err(ctx, PCS_ERR_ARG, "conv_fold: no 3x3 instantiation for N=%d planes=%d", a.npad, a.nplanes);
    }
    if (a.pair_src) {
        if (key == 323) return launch_fold_t<T, 32, 3, 16, 16, 4, true>(ctx, a);    // conv2 <- conv1 with channels 16..19 as pixel-pair units
        return set_err(ctx, PCS_ERR_ARG, "conv_fold: no pixel-pair instantiation for N=%d planes=%d", a.npad, a.nplanes);
    }
    switch (key) {          //               NPAD planes RING SLOTS epilogue groups
        case 323: return launch_fold_t<T, 32, 3, 16, 16, 4>(ctx, a);     // conv2: 20(24) -> 30(32)
        case 484: return launch_fold_t<T, 48, 4, 10, 10, 3>(ctx, a);     // conv3: 30(32) -> 40(48)
        case 485: return launch_fold_t<T, 48, 5, 10, 10, 3>(ctx, a);     // conv4: 40(40) -> 40(48)
        case 325: return launch_fold_t<T, 32, 5, 8, 16, 4>(ctx, a);      // conv5: 40(40) -> 60 as 32 + 28 output channels
        case 328: return launch_fold_t<T, 32, 8, 4, 16, 4>(ctx, a);       // conv6: 60(64) -> 60 as 32 + 28 output channels
        case 488: return launch_fold_t<T, 48, 8, 4, 8, 3>(ctx, a);       // deconv3: 60(64) [+ 60(64)] -> 40(48), one launch per source
        // C_out = 40 as five whole planes (N' = 208 instead of 240: 13 % fewer tensor cycles), twelve slots of 40 columns
        case 404: return launch_fold_t<T, 40, 4, 12, 12, 3>(ctx, a);     // conv3: 30(32) -> 40
        case 405: return launch_fold_t<T, 40, 5, 6, 12, 3>(ctx, a);      // conv4: 40 -> 40
        case 408: return launch_fold_t<T, 40, 8, 4, 12, 3>(ctx, a);      // deconv3: 60(64) [+ 60(64)] -> 40, one launch per source
        default: return set_err(ctx, PCS_ERR_ARG, "conv_fold: no instantiation for N=%d planes=%d", a.npad, a.nplanes);
    }
}

}  // namespace

// (N tile, source planes) pairs with an instantiation; the resident weights + input ring must fit shared memory
bool fold_supported(int k, int npad, int nplanes) {
    if (k == 3) return npad == 64 && (nplanes == 8 || nplanes == 16);       // 16 planes = two concatenated sources of 8
    if (k != 5) return false;
    const int key = npad * 10 + nplanes;
    return key == 323 || key == 484 || key == 485 || key == 325 || key == 328 || key == 488 || key == 404 || key == 405 || key == 408;
}

// Resident operand image [K step][K half][row = (4-dy)*NPAD + o][8] (rows padded with zeros to a multiple of 16; K steps as in FoldK) for the input channels
// [ci0, ci0 + cin) and the output channels [o0, o0 + ncols) of a layer with weights w32[25][cin_total][cout_total]:
// the N blocks run from the oldest output row of the window (dy = 4) to the newest (dy = 0).
size_t fold_weight_image(const float* w32, int cin_total, int cout_total, int ci0, int cin, int o0, int ncols, int npad,
                         int precision, std::vector<uint16_t>& out, bool pairx, int ks) {
    const int npl = pad8(cin) / 8, pairs = npl / 2, odd = npl & 1, nmma = pairs * ks + odd * (pairx ? 2 : 3), nf = (ks * npad + 15) / 16 * 16;
    if (ks != 5 && (odd || pairx)) { out.clear(); return 0; }
    if (pairx && (!odd || cin - (npl - 1) * 8 > 4)) { out.clear(); return 0; }      // pair units carry four channels
    out.assign((size_t)nmma * 2 * nf * 8, 0);
    auto conv = [&](float v) -> uint16_t {
        if (precision == PCS_PREC_BF16) { __nv_bfloat16 b = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&b); }
        __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h);
    };
    for (int q = 0; q < nmma; ++q)
        for (int half = 0; half < 2; ++half) {
            int plane, dx;
            if (q < pairs * ks) { plane = 2 * (q / ks) + half; dx = q % ks; }
            else if (pairx) {
                // K half = one pair unit = [4 channels at tap d | the same channels at tap d + 1]: step 0 holds the taps
                // (0,1) | (2,3), step 1 the taps (3,4) | (4,5) of which only tap 4 of the first half carries weights
                const int t = q - pairs * 5;
                for (int dy = 0; dy < 5; ++dy)
                    for (int o = 0; o < ncols; ++o)
                        for (int e = 0; e < 8; ++e) {
                            const int ci = (npl - 1) * 8 + (e & 3), tap = (t == 0 ? 2 * half : 3 + half) + (e >> 2);
                            if (ci >= cin || tap > 4 || (t == 1 && (half == 1 || tap != 4))) continue;
                            const float v = w32[((size_t)(dy * 5 + tap) * cin_total + ci0 + ci) * cout_total + o0 + o];
                            out[(((size_t)q * 2 + half) * nf + (4 - dy) * npad + o) * 8 + e] = conv(v);
                        }
                continue;
            } else {
                const int t = q - pairs * 5;                  // taps (0,1), (2,3), (3,4); the repeated tap 3 gets zeros
                plane = npl - 1;
                dx = (t == 0 ? 0 : t + 1) + half;
                if (t == 2 && half == 0) continue;
            }
            for (int dy = 0; dy < ks; ++dy)
                for (int o = 0; o < ncols; ++o)
                    for (int e = 0; e < 8; ++e) {
                        const int ci = plane * 8 + e;
                        if (ci >= cin) continue;
                        const float v = w32[((size_t)(dy * ks + dx) * cin_total + ci0 + ci) * cout_total + o0 + o];
                        out[(((size_t)q * 2 + half) * nf + (ks - 1 - dy) * npad + o) * 8 + e] = conv(v);
                    }
        }
    return out.size() * sizeof(uint16_t);
}

int launch_conv_fold(pcs_ctx* ctx, const FoldConvArgs& a) {
    if (!fold_supported(a.k, a.npad, a.nplanes)) return set_err(ctx, PCS_ERR_ARG, "conv_fold: unsupported layer");
    if (ctx->precision == PCS_PREC_BF16) return launch_fold_dispatch<__nv_bfloat16>(ctx, a);
    return launch_fold_dispatch<__half>(ctx, a);
}

}  // namespace pcs
