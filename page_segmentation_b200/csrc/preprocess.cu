// Page preprocessing: line-height-normalised rescale of the grey and the
// binarised page (reference: ocr4all_pixel_classifier/lib/dataset.py:114-150,
// i.e. skimage 0.17.2 rescale(order=0) / resize(order=3, mode='reflect',
// clip=True, preserve_range=True, anti_aliasing=len(unique)>2)).
//
// All interpolation arithmetic is IEEE fp64 with explicit round-to-nearest
// intrinsics in the same association order as the numpy restatement
// (oracle/resize.py) so that results are bit-identical (no FMA contraction).
#include "common.cuh"

namespace pcs {

// ---------------------------------------------------------------------------
// per-page grey-level presence bitmap (256 bits) -> min, max, #levels
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) level_bits_kernel(const uint8_t* __restrict__ src, size_t page_bytes,
                                                         uint32_t* __restrict__ bits /*[n][8]*/,
                                                         const int4* __restrict__ page_lv /* or null: pages with .z set are skipped */) {
    if (page_lv && page_lv[blockIdx.y].z) return;
    __shared__ uint32_t s_bits[8];
    if (threadIdx.x < 8) s_bits[threadIdx.x] = 0;
    __syncthreads();
    const int page = blockIdx.y;
    const uint8_t* p = src + (size_t)page * page_bytes;
    // 16-byte aligned body, scalar head/tail (pages need not start on a 16-B boundary)
    size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > page_bytes) head = page_bytes;
    const size_t nvec = (page_bytes - head) / 16;
    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
    auto mark = [&](uint32_t val) {
        if (!((s_bits[val >> 5] >> (val & 31)) & 1u)) atomicOr(&s_bits[val >> 5], 1u << (val & 31));
    };
    uint32_t last = 0;
    bool have_last = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldg(pv + i);
        uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t wv = words[k];
            if (have_last && wv == last) continue;
            last = wv;
            have_last = true;
#pragma unroll
            for (int b = 0; b < 4; ++b) mark((wv >> (8 * b)) & 0xff);
        }
    }
    if (blockIdx.x == 0) {
        for (size_t i = threadIdx.x; i < head; i += blockDim.x) mark(p[i]);
        for (size_t i = head + nvec * 16 + threadIdx.x; i < page_bytes; i += blockDim.x) mark(p[i]);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_bits[threadIdx.x]) atomicOr(&bits[page * 8 + threadIdx.x], s_bits[threadIdx.x]);
}

__device__ __forceinline__ int reflect_coord(long long c, int dim) {
    // coord_map(dim, coord, 'R') of skimage/_shared/interpolation.pxd
    if (dim == 1) return 0;
    const long long cmax = dim - 1;
    if (c < 0) {
        long long a = -c;
        return (int)(((a / cmax) & 1) ? cmax - (a % cmax) : a % cmax);
    }
    if (c > cmax) return (int)(((c / cmax) & 1) ? cmax - (c % cmax) : c % cmax);
    return (int)c;
}

__device__ __forceinline__ double cubic_rn(double x, double f0, double f1, double f2, double f3) {
    // f1 + 0.5*x*(f2 - f0 + x*(2*f0 - 5*f1 + 4*f2 - f3 + x*(3*(f1 - f2) + f3 - f0)))
    double i3 = __dsub_rn(__dadd_rn(__dmul_rn(3.0, __dsub_rn(f1, f2)), f3), f0);
    double i2 = __dadd_rn(
        __dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, f0), __dmul_rn(5.0, f1)), __dmul_rn(4.0, f2)), f3),
        __dmul_rn(x, i3));
    double i1 = __dadd_rn(__dsub_rn(f2, f0), __dmul_rn(x, i2));
    return __dadd_rn(f1, __dmul_rn(__dmul_rn(0.5, x), i1));
}

__device__ __forceinline__ int level_count(const uint32_t* bits) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) c += __popc(bits[k]);
    return c;
}

// ---------------------------------------------------------------------------
// Fast path for pages with at most two grey levels (binarised scans: the case dataset.py:169-172 makes
// the normal one).  One streaming pass over the page produces a 1-bit-per-pixel plane
// bit(i) = (page[i] != page[0])  and three sums per page that decide exactly whether the page has at most
// two levels; the resampler then never touches the 8.7 MB page again: a source value is
// bit ? other level : page[0].
// ---------------------------------------------------------------------------
// High bit of every byte of x that is not zero.
__device__ __forceinline__ uint32_t nz_flags(uint32_t x) { return (x | ((x | 0x80808080u) - 0x01010101u)) & 0x80808080u; }

// The 16 bytes of v compared with the bytes of ref4: bit i (of 16) = byte i differs.  Flags of two words are
// interleaved at a spacing of four bits and gathered by ONE multiplication (32 partial products, no two on the
// same bit, so no carries): the eight result bits arrive in the top byte of the product.
__device__ __forceinline__ uint32_t ne_bits16(const uint4 v, uint32_t ref4, uint32_t& s1, uint32_t& s2) {
    const uint32_t x0 = v.x ^ ref4, x1 = v.y ^ ref4, x2 = v.z ^ ref4, x3 = v.w ^ ref4;
    s1 = __dp4a(x0, 0x01010101u, s1); s2 = __dp4a(x0, x0, s2);
    s1 = __dp4a(x1, 0x01010101u, s1); s2 = __dp4a(x1, x1, s2);
    s1 = __dp4a(x2, 0x01010101u, s1); s2 = __dp4a(x2, x2, s2);
    s1 = __dp4a(x3, 0x01010101u, s1); s2 = __dp4a(x3, x3, s2);
    const uint32_t p01 = ((nz_flags(x0) >> 4) | nz_flags(x1)) * 0x00204081u;       // top byte: bytes of x0, then of x1
    const uint32_t p23 = ((nz_flags(x2) >> 4) | nz_flags(x3)) * 0x00204081u;
    return __byte_perm(p01, p23, 0x4473) & 0xffffu;                                // byte 0 = p01 >> 24, byte 1 = p23 >> 24
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Per-page sums over x = byte ^ page[0]: [0] = number of non-zero x, [1] = sum of x, [2] = sum of x^2.  The non-zero x
// are all equal (the page has at most two levels) iff  [0] * [2] == [1]^2  (Cauchy-Schwarz, exact in integers).
constexpr int kScanUnroll = 2;

// grid = (blocks, pages); a warp turns 2 KB of page into 64 bitmap words per step: lane l loads the 16-byte groups
// l, 32 + l, 64 + l, 96 + l (fully coalesced), neighbouring lanes exchange half words so that every lane holds whole
// 32-pixel words.  Requires 16-byte aligned pages whose size is a multiple of 32 bytes (checked by the host; other
// shapes take the general kernels).
__global__ void __launch_bounds__(256) scan_pack_kernel(const uint8_t* __restrict__ src, size_t page_bytes,
                                                        unsigned long long* __restrict__ stats /*[n][gridDim.x][3]*/,
                                                        uint32_t* __restrict__ bitmap, size_t bitmap_words /*per page, padded*/) {
    const int page = blockIdx.y;
    const uint8_t* p = src + (size_t)page * page_bytes;
    const uint4* pv = reinterpret_cast<const uint4*>(p);
    uint32_t* bm = bitmap + (size_t)page * bitmap_words;
    const uint32_t ref4 = (uint32_t)__ldg(p) * 0x01010101u;
    const size_t nvec = page_bytes / 16;                  // even
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    uint32_t cnt = 0, s1 = 0, s2 = 0;
    constexpr int U = 2 * kScanUnroll;                    // 16-byte groups per lane and step
    for (size_t base = warp * (32 * U); base < nvec; base += nwarps * (32 * U)) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = base + 32 * u + lane;
            v[u] = i < nvec ? ldg_stream(pv + i) : make_uint4(ref4, ref4, ref4, ref4);
        }
#pragma unroll
        for (int u = 0; u < U; u += 2) {
            const uint32_t h0 = ne_bits16(v[u], ref4, s1, s2), h1 = ne_bits16(v[u + 1], ref4, s1, s2);
            // even lane 2j: word j of the first 32 groups = h0(2j) | h0(2j+1) << 16; odd lane 2j+1: word j of the next 32 groups
            const uint32_t got = __shfl_xor_sync(0xffffffffu, (lane & 1) ? h0 : h1, 1);
            const uint32_t word = (lane & 1) ? (got | (h1 << 16)) : (h0 | (got << 16));
            cnt += __popc(h0) + __popc(h1);
            const size_t wi = (base + 32 * u) / 2 + (lane >> 1) + ((lane & 1) ? 16 : 0);
            if (wi < nvec / 2) bm[wi] = word;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) bm[nvec / 2] = 0;       // pad word read by the funnel shift
    // one partial result per block (plain stores: thousands of atomics on one sector per page serialise in L2)
    __shared__ unsigned long long s_part[8][3];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    s1 = __reduce_add_sync(0xffffffffu, s1);
    unsigned long long q = s2;                           // s2 <= 65025 * 16 * U per lane and step: 64 bits across the warp
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) { s_part[threadIdx.x >> 5][0] = cnt; s_part[threadIdx.x >> 5][1] = s1; s_part[threadIdx.x >> 5][2] = q; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
        stats[((size_t)page * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
    }
}

// A warp per page, a block (16 warps) per group of pages: the verdict of the sums.  lv[page] = {level of a zero bit
// (= page[0]), level of a set bit, 1 iff at most two levels, -}.  Pages with at most two levels get their level bitmap
// here; the others are scanned by level_bits_kernel (gated on lv).  Also the per-group flag "some page of the group
// has more than two levels".
__global__ void __launch_bounds__(512)
levels_from_stats_kernel(const uint8_t* __restrict__ src, size_t page_bytes, const unsigned long long* __restrict__ stats, int parts,
                         int n, int group /* <= 16 */, uint32_t* __restrict__ bits /*[n][8], zeroed*/, int4* __restrict__ lv,
                         int* __restrict__ gflags) {
    __shared__ int s_many[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, pg = blockIdx.x * group + w;
    if (w < 16) s_many[w] = 0;
    __syncthreads();
    if (w < group && pg < n) {
        unsigned long long c = 0, s1 = 0, s2 = 0;
        for (int i = lane; i < parts; i += 32) {
            const unsigned long long* p = stats + ((size_t)pg * parts + i) * 3;
            c += p[0]; s1 += p[1]; s2 += p[2];
        }
        for (int o = 16; o; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            const int a = src[(size_t)pg * page_bytes];
            // c * s2 and s1 * s1 as 128-bit products
            const bool two = c == 0 || (c * s2 == s1 * s1 && __umul64hi(c, s2) == __umul64hi(s1, s1));
            const int b = (two && c) ? a ^ (int)(s1 / c) : a;
            lv[pg] = make_int4(a, b, two ? 1 : 0, 0);
            if (two) {
                atomicOr(&bits[pg * 8 + (a >> 5)], 1u << (a & 31));
                atomicOr(&bits[pg * 8 + (b >> 5)], 1u << (b & 31));
            } else {
                s_many[w] = 1;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int f = 0;
        for (int k = 0; k < 16; ++k) f |= s_many[k];
        gflags[blockIdx.x] = f;
    }
}

__global__ void set_levels_kernel(int4* __restrict__ lv, int n, int l0, int l1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lv[i] = make_int4(l0, l1, 1, 0);
}

// v / 255 correctly rounded without the division sequence: q0 = RN(v * RN(1/255)), exact residual r = v - 255 q0 by one
// FMA, q = RN(q0 + r * RN(1/255)) (Markstein; checked against `/` on 2e9 doubles of [0, 255], incl. subnormals).
__device__ __forceinline__ double div255_rn(double v) {
    const double c = 1.0 / 255.0;
    const double q0 = __dmul_rn(v, c);
    return __fma_rn(__fma_rn(-255.0, q0, v), c, q0);
}
// clip=True, then img = 1.0 - v/255 ; (img*255).astype(uint8)
__device__ __forceinline__ uint8_t finish_image(double v, double vmin, double vmax) {
    v = fmin(fmax(v, vmin), vmax);
    return (uint8_t)(int)__dmul_rn(__dsub_rn(1.0, div255_rn(v)), 255.0);
}

constexpr int R2_TW = 128, R2_TH = 32;                     // output tile: 32 quads of columns x 32 rows
constexpr int R2_MAXF = 4;                                 // largest scale factor (source pixels per output pixel)
constexpr int R2_ROWS = R2_MAXF * (R2_TH - 1) + 6;         // staged source rows: first tap of the first row .. last tap of the last
constexpr int R2_WORDS = (R2_MAXF * (R2_TW - 1) + 5 + 31) / 32 + 2;      // staged words per source row (+1 for the funnel shift)

// grid = (ceil(Ws/128), ceil(tiles_y / tiles_per_block), pages).  Bit-identical to the general kernel:
//   * the horizontal cubic of a two-level row has only 16 possible operand patterns per output column; they are
//     evaluated once per block with the same fp64 operation order (cubic_rn) and looked up;
//   * a pixel whose 4x4 neighbourhood is all one level v gets cubic(v,v,v,v) = v exactly (every intermediate of
//     cubic_rn is an exact small integer or zero);
//   * the vertical cubic, clip and the (1 - v/255) * 255 truncation are otherwise unchanged.
// Source bit rows are staged in LOGICAL coordinates (reflection applied while staging), so that the taps of a pixel are
// four consecutive bits of four consecutive staged rows.  A thread owns a quad of four adjacent output columns: one
// funnel shift per tap row gives the bits of all four neighbourhoods, and a quad whose window is all paper or all ink
// (three of four) is finished with integer work only.  The other quads go to a list in shared memory and are
// evaluated afterwards, one pixel per thread, with every lane busy (phase B).
__global__ void __launch_bounds__(256)
resample_bits_kernel(const uint8_t* __restrict__ bin, int bin_is_grey, int H, int W, int Hs, int Ws, double f_r, double t_r,
                     double f_c, double t_c, const int4* __restrict__ page_lv, const uint32_t* __restrict__ bitmap,
                     size_t bitmap_words, uint8_t* __restrict__ image_out, uint8_t* __restrict__ binary_out, int tiles_per_block) {
    __shared__ double s_lut[16][R2_TW];                   // [pattern][column]
    __shared__ double s_cfrac[R2_TW], s_rfrac[R2_TH];
    __shared__ uint32_t s_bm[R2_ROWS * R2_WORDS];
    __shared__ uint32_t s_list[(R2_TW / 4) * R2_TH * 3];
    __shared__ int s_c0[R2_TW], s_cnn[R2_TW], s_r0[R2_TH], s_rnn[R2_TH];      // floor / round of the sampling position (logical)
    __shared__ int s_count;
    const int page = blockIdx.z;
    const int4 lv = page_lv[page];
    if (!lv.z) return;                                     // more than two levels: the general kernels do this page
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int X0 = blockIdx.x * R2_TW;
    const int va = lv.x, vb = lv.y;
    const double vmin = (double)min(va, vb), vmax = (double)max(va, vb);
    if (tid < R2_TW) {
        const int o = min(X0 + tid, Ws - 1);               // replicate past the edge
        const double pc = __dadd_rn(__dmul_rn(f_c, (double)o), t_c);
        const double pf = floor(pc);
        s_c0[tid] = (int)pf;
        s_cnn[tid] = (int)round(pc);                       // order 0: C round()
        s_cfrac[tid] = __dsub_rn(pc, pf);
    }
    __syncthreads();
    {
        const double fa = (double)va, fb = (double)vb;
        for (int i = tid; i < R2_TW * 16; i += 256) {
            const int c = i & (R2_TW - 1), pat = i / R2_TW;
            s_lut[pat][c] = cubic_rn(s_cfrac[c], (pat & 1) ? fb : fa, (pat & 2) ? fb : fa, (pat & 4) ? fb : fa, (pat & 8) ? fb : fa);
        }
    }
    // the quad of this thread: bit offset of its window in a staged row, offsets of the four neighbourhoods inside the
    // window, positions of the four nearest-neighbour columns inside the window
    const int cl = s_c0[0] - 1;                            // logical source column of staged bit 0
    const int ch = s_c0[R2_TW - 1] + 2;                    // last logical source column needed
    const int q = s_c0[4 * lane] - 1 - cl;
    const int qw = q >> 5, qs = q & 31;
    int off[4], nnb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        off[j] = s_c0[4 * lane + j] - s_c0[4 * lane];
        nnb[j] = s_cnn[4 * lane + j] - (s_c0[4 * lane] - 1);
    }
    const uint32_t span = (2u << (off[3] + 3)) - 1u;       // window bits 0 .. off[3] + 3
    uint32_t out_a4 = finish_image((double)va, vmin, vmax) * 0x01010101u, out_b4 = finish_image((double)vb, vmin, vmax) * 0x01010101u;
    asm volatile("" : "+r"(out_a4), "+r"(out_b4));         // keep the two bytes in registers (not re-derived from doubles per row)
    const uint32_t bin_a = va == 0 ? 1u : 0u, bin_b = vb == 0 ? 1u : 0u;      // bin = (1.0 - NN(binary/255 or binary)).astype(uint8)
    const int x0 = X0 + 4 * lane;
    const int nvalid = min(4, max(0, Ws - x0));            // columns of the quad that exist
    const uint32_t* bm = bitmap + (size_t)page * bitmap_words;
    uint8_t* img_page = image_out + (size_t)page * Hs * Ws;
    uint8_t* bin_page = binary_out ? binary_out + (size_t)page * Hs * Ws : nullptr;
    const int nwq = min(R2_WORDS, (ch - cl + 1 + 31) / 32 + 1);
    if (ch - cl + 1 + 32 > R2_WORDS * 32) __trap();        // host guarantees scale factors <= 4
    auto store4 = [&](uint8_t* p, uint32_t v4) {
        if (nvalid == 4) {
            p[0] = (uint8_t)v4; p[1] = (uint8_t)(v4 >> 8); p[2] = (uint8_t)(v4 >> 16); p[3] = (uint8_t)(v4 >> 24);
        } else {
            for (int j = 0; j < nvalid; ++j) p[j] = (uint8_t)(v4 >> (8 * j));
        }
    };

    const int tiles_y = (Hs + R2_TH - 1) / R2_TH;
    for (int t = 0; t < tiles_per_block; ++t) {
        const int tile_y = blockIdx.y * tiles_per_block + t;
        if (tile_y >= tiles_y) break;
        const int Y0 = tile_y * R2_TH;
        __syncthreads();                                   // phase B of the previous tile is done with the tables
        if (tid < R2_TH) {
            const int o = min(Y0 + tid, Hs - 1);
            const double pr = __dadd_rn(__dmul_rn(f_r, (double)o), t_r);
            const double pf = floor(pr);
            s_r0[tid] = (int)pf;
            s_rnn[tid] = (int)round(pr);
            s_rfrac[tid] = __dsub_rn(pr, pf);
        }
        if (tid == 0) s_count = 0;
        __syncthreads();
        const int rl = s_r0[0] - 1;                        // logical source row of staged row 0
        const int nrows = s_r0[R2_TH - 1] + 2 - rl + 1;
        if (nrows > R2_ROWS) __trap();
        // staging: 16 threads per source row (one word each; a second round for the widest rows), 16 rows per step.  The
        // columns of a word that lie on the page come from one funnel shift (clamped at the page borders)
        for (int r = tid >> 4; r < nrows; r += 16) {
            int prow = rl + r;
            if ((unsigned)prow >= (unsigned)H) prow = reflect_coord((long long)prow, H);
            const uint32_t rowbit = (uint32_t)prow * (uint32_t)W;          // pages have fewer than 2^32 pixels (host check)
            for (int wq = tid & 15; wq < nwq; wq += 16) {
                const int lc = cl + 32 * wq;               // logical column of bit 0 of this word
                uint32_t word = 0;
                if (lc >= 0 && lc + 31 < W) {              // a word inside the page (all of them except in border tiles)
                    const uint32_t b0 = rowbit + (uint32_t)lc;
                    word = __funnelshift_r(__ldg(bm + (b0 >> 5)), __ldg(bm + (b0 >> 5) + 1), b0 & 31u);
                } else {
                    const int lo = max(lc, 0), hi = min(lc + 31, W - 1);
                    if (lo <= hi) {
                        const uint32_t b0 = rowbit + lo;
                        const uint32_t raw = __funnelshift_r(__ldg(bm + (b0 >> 5)), __ldg(bm + (b0 >> 5) + 1), (uint32_t)(b0 & 31));
                        word = (raw & (0xffffffffu >> (31 - (hi - lo)))) << (lo - lc);
                    }
                }
                s_bm[r * R2_WORDS + wq] = word;
            }
        }
        if (cl < 0 || ch > W - 1) {                        // block-uniform: tiles at the left / right page border
            // the (at most two on either side) needed columns beyond the border are reflected one by one, a row per thread
            __syncthreads();
            for (int r = tid; r < nrows; r += 256) {
                int prow = rl + r;
                if ((unsigned)prow >= (unsigned)H) prow = reflect_coord((long long)prow, H);
                const size_t rowbit = (size_t)prow * W;
                for (int c = cl; c < 0; ++c) {
                    const size_t b = rowbit + reflect_coord((long long)c, W);
                    s_bm[r * R2_WORDS + ((c - cl) >> 5)] |= ((__ldg(bm + (b >> 5)) >> (b & 31)) & 1u) << ((c - cl) & 31);
                }
                for (int c = max(W, cl); c <= ch; ++c) {
                    const size_t b = rowbit + reflect_coord((long long)c, W);
                    s_bm[r * R2_WORDS + ((c - cl) >> 5)] |= ((__ldg(bm + (b >> 5)) >> (b & 31)) & 1u) << ((c - cl) & 31);
                }
            }
        }
        __syncthreads();
        // ---- phase A: a quad per thread, a row per warp ----
#pragma unroll
        for (int ty = wrp; ty < R2_TH; ty += 8) {
            const int y = Y0 + ty;
            if (y >= Hs) break;                            // warp-uniform
            const int r0 = s_r0[ty];
            const uint32_t* rowp = s_bm + (r0 - 1 - rl) * R2_WORDS + qw;
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = __funnelshift_r(rowp[k * R2_WORDS], rowp[k * R2_WORDS + 1], qs);
            const uint32_t any = (w[0] | w[1] | w[2] | w[3]) & span, all = w[0] & w[1] & w[2] & w[3] & span;
            const bool pending = any != 0u && all != span;
            const int rowoff = y * Ws + x0;
            const unsigned pm = __ballot_sync(0xffffffffu, pending);
            uint32_t bin4 = (any ? bin_b : bin_a) * 0x01010101u;
            if (pm) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_count, __popc(pm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pending) {
                    const int e = 3 * (base + __popc(pm & ((1u << lane) - 1u)));
                    s_list[e] = __byte_perm(w[0], w[1], 0x5410);
                    s_list[e + 1] = __byte_perm(w[2], w[3], 0x5410);
                    s_list[e + 2] = (uint32_t)(ty << 5 | lane);
                    const uint32_t wn = (s_rnn[ty] != r0) ? w[2] : w[1];        // the nearest row is tap row 1 or 2
                    const uint32_t m = ((wn >> nnb[0]) & 1u) | (((wn >> nnb[1]) & 1u) << 8) | (((wn >> nnb[2]) & 1u) << 16) |
                                       (((wn >> nnb[3]) & 1u) << 24);
                    bin4 = (bin_a * 0x01010101u) ^ (m * (bin_a ^ bin_b));
                }
            }
            if (!pending) store4(img_page + rowoff, any ? out_b4 : out_a4);
            if (bin_page) {
                if (!bin_is_grey) {                        // a binary page of its own: sampled from the uint8 page
                    const int prow = reflect_coord((long long)s_rnn[ty], H);
                    bin4 = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pcol = reflect_coord((long long)s_cnn[min(4 * lane + j, R2_TW - 1)], W);
                        bin4 |= (bin[(size_t)page * H * W + (size_t)prow * W + pcol] == 0 ? 1u : 0u) << (8 * j);
                    }
                }
                store4(bin_page + rowoff, bin4);
            }
        }
        __syncthreads();
        // ---- phase B: the pixels of the listed quads, one per thread ----
        const int npx = 4 * s_count;
        for (int i = tid; i < npx; i += 256) {
            const uint32_t* ent = s_list + 3 * (i >> 2);
            const uint32_t meta = ent[2];
            const int tx = 4 * (int)(meta & 31u) + (i & 3), ty = (int)(meta >> 5);
            if (X0 + tx >= Ws) continue;
            const int o = s_c0[tx] - s_c0[tx & ~3];
            const uint32_t w01 = ent[0] >> o, w23 = ent[1] >> o;         // tap rows 0, 2 at bit 0, rows 1, 3 at bit 16
            const char* lut = reinterpret_cast<const char*>(&s_lut[0][tx]);
            constexpr uint32_t kPat = 15u * R2_TW * 8;                   // pattern index scaled to the byte stride of s_lut
            const double f0 = *reinterpret_cast<const double*>(lut + ((w01 * (R2_TW * 8)) & kPat));
            const double f1 = *reinterpret_cast<const double*>(lut + ((w01 >> 16) * (R2_TW * 8) & kPat));
            const double f2 = *reinterpret_cast<const double*>(lut + ((w23 * (R2_TW * 8)) & kPat));
            const double f3 = *reinterpret_cast<const double*>(lut + ((w23 >> 16) * (R2_TW * 8) & kPat));
            img_page[(Y0 + ty) * Ws + X0 + tx] = finish_image(cubic_rn(s_rfrac[ty], f0, f1, f2, f3), vmin, vmax);
        }
    }
}

static int g_r2_tiles_per_block = 4;
static int launch_resample_bits(pcs_ctx* ctx, const uint8_t* d_bin, int bin_is_grey, int n, int H, int W, int Hs, int Ws, const int4* d_lv,
                                const uint32_t* d_bitmap, size_t bitmap_words, uint8_t* d_image, uint8_t* d_binary) {
    if ((size_t)Hs * Ws >= (size_t)1 << 31 || (size_t)H * W >= (size_t)1 << 32)
        return set_err(ctx, PCS_ERR_ARG, "preprocess: pages of 2^32 pixels (2^31 after scaling) and more are not supported");
    // sampling positions  p = f * o + (0.5 f - 0.5), every operation rounded to nearest (host doubles = the device's)
    const double f_r = (double)H / (double)Hs, f_c = (double)W / (double)Ws;
    const volatile double h_r = 0.5 * f_r, h_c = 0.5 * f_c;
    const double t_r = h_r - 0.5, t_c = h_c - 0.5;
    static const int tpb_env = [] { const char* e = getenv("PCSEG_RESAMPLE_TILES"); return e ? atoi(e) : 0; }();
    const int tpb = tpb_env > 0 ? tpb_env : g_r2_tiles_per_block;
    const int tiles_y = (Hs + R2_TH - 1) / R2_TH;
    dim3 grid((Ws + R2_TW - 1) / R2_TW, (tiles_y + tpb - 1) / tpb, n);
    resample_bits_kernel<<<grid, 256, 0, ctx->stream>>>(d_bin, bin_is_grey, H, W, Hs, Ws, f_r, t_r, f_c, t_c, d_lv, d_bitmap, bitmap_words,
                                                        d_image, d_binary, tpb);
    PCS_LAUNCH_CHECK(ctx, "resample_bits_kernel");
    return PCS_OK;
}

// General resampler: one thread per output pixel, persistent grid over 32x8 tiles of the pages of a group.
// Per page (block-uniform): more than two grey levels => the bicubic source is the Gaussian-
// filtered fp64 plane of that page (anti_aliasing=True, dataset.py:127) and the clip range its
// min/max; otherwise the uint8 page itself and the min/max from the level bitmap.  No host
// decision is involved, so the whole preprocess is asynchronous.
__global__ void __launch_bounds__(256)
resample_kernel(const uint8_t* __restrict__ grey, const uint8_t* __restrict__ bin, int H, int W, int Hs, int Ws,
                const uint32_t* __restrict__ level_bits, const double* __restrict__ aa_planes,
                const unsigned long long* __restrict__ aa_minmax, int page0, uint8_t* __restrict__ image_out,
                uint8_t* __restrict__ binary_out, int skip_two_level, const int* __restrict__ group_flag, int pages,
                double* __restrict__ image_f64 /* [pages][Hs][Ws]: img = 1 - v/255 kept in fp64 (max_width pass) or null */) {
    // skip_two_level: pages with at most two grey levels were already done by resample_bits_kernel; when
    // no page of the group has more (group_flag == 0) the whole persistent grid leaves at once
    if (skip_two_level && group_flag && !*group_flag) return;
    // per-tile coordinate tables: the sampling positions depend on the column (row) only
    __shared__ int s_cols[32][4], s_rows[8][4], s_ci[32], s_ri[8];
    __shared__ double s_xc[32], s_xr[8];
    const int tiles_x = (Ws + 31) / 32, tiles_y = (Hs + 7) / 8;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y * pages; tile += gridDim.x) {
        const int bz = tile / (tiles_x * tiles_y), bxy = tile - bz * (tiles_x * tiles_y);
        const int by = bxy / tiles_x, bx = bxy - by * tiles_x;
        __syncthreads();                               // the coordinate tables of the previous tile are no longer read
        if (skip_two_level && level_count(level_bits + (size_t)(page0 + bz) * 8) <= 2) continue;
        const int tid = threadIdx.y * 32 + threadIdx.x;
        if (tid < 40) {
            const bool is_col = tid < 32;
            const int i = is_col ? tid : tid - 32;
            const int o = is_col ? bx * 32 + i : by * 8 + i;
            const int n_in = is_col ? W : H, n_out = is_col ? Ws : Hs;
            const double f = __ddiv_rn((double)n_in, (double)n_out);
            const double p = __dadd_rn(__dmul_rn(f, (double)o), __dsub_rn(__dmul_rn(0.5, f), 0.5));
            const int nn = reflect_coord((long long)round(p), n_in);        // order 0: C round(), then reflect
            const double pf = floor(p);
            const double frac = __dsub_rn(p, pf);
            const long long p0 = (long long)pf - 1;
            if (is_col) {
                s_ci[i] = nn; s_xc[i] = frac;
                for (int k = 0; k < 4; ++k) s_cols[i][k] = reflect_coord(p0 + k, n_in);
            } else {
                s_ri[i] = nn; s_xr[i] = frac;
                for (int k = 0; k < 4; ++k) s_rows[i][k] = reflect_coord(p0 + k, n_in);
            }
        }
        __syncthreads();
        const int x = bx * 32 + threadIdx.x;
        const int y = by * 8 + threadIdx.y;
        if (x >= Ws || y >= Hs) continue;
        const int page = page0 + bz;
        const size_t src_off = (size_t)page * H * W;
        const size_t dst_off = (size_t)page * Hs * Ws + (size_t)y * Ws + x;

        if (binary_out) {
            const uint8_t v = bin[src_off + (size_t)s_ri[threadIdx.y] * W + s_ci[threadIdx.x]];
            // bin = (1.0 - NN(binary/255 or binary)).astype(uint8): 1 iff v == 0
            binary_out[dst_off] = (v == 0) ? 1 : 0;
        }
        if (image_out || image_f64) {
            const uint32_t* bits = level_bits + (size_t)page * 8;
            const bool aa = aa_planes != nullptr && level_count(bits) > 2;
            double vmin, vmax;
            if (aa) {
                vmin = __longlong_as_double((long long)aa_minmax[2 * bz]);
                vmax = __longlong_as_double((long long)aa_minmax[2 * bz + 1]);
            } else {
                int lo = 0, hi = 255;
                for (int wv = 0; wv < 8; ++wv)
                    if (bits[wv]) { lo = wv * 32 + __ffs(bits[wv]) - 1; break; }
                for (int wv = 7; wv >= 0; --wv)
                    if (bits[wv]) { hi = wv * 32 + 31 - __clz(bits[wv]); break; }
                vmin = (double)lo;
                vmax = (double)hi;
            }
            const double xr = s_xr[threadIdx.y], xc = s_xc[threadIdx.x];
            const int c0 = s_cols[threadIdx.x][0], c1 = s_cols[threadIdx.x][1], c2 = s_cols[threadIdx.x][2], c3 = s_cols[threadIdx.x][3];
            double frow[4];
            if (aa) {
                const double* g = aa_planes + (size_t)bz * H * W;
    #pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const size_t ro = (size_t)s_rows[threadIdx.y][pr] * W;
                    frow[pr] = cubic_rn(xc, g[ro + c0], g[ro + c1], g[ro + c2], g[ro + c3]);
                }
            } else {
                const uint8_t* g = grey + src_off;
    #pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const size_t ro = (size_t)s_rows[threadIdx.y][pr] * W;
                    frow[pr] = cubic_rn(xc, (double)g[ro + c0], (double)g[ro + c1], (double)g[ro + c2], (double)g[ro + c3]);
                }
            }
            double v = cubic_rn(xr, frow[0], frow[1], frow[2], frow[3]);
            v = fmin(fmax(v, vmin), vmax);                       // clip=True
            // img = 1.0 - v/255 ; (img*255).astype(uint8)
            const double img = __dsub_rn(1.0, __ddiv_rn(v, 255.0));
            if (image_f64) image_f64[(size_t)bz * Hs * Ws + (size_t)y * Ws + x] = img;
            else image_out[dst_off] = (uint8_t)(int)__dmul_rn(img, 255.0);      // C truncation
        }
    }
}

// orig_binary = (1 - binary/255).astype(uint8) == (v == 0)
__global__ void __launch_bounds__(256) orig_binary_kernel(const uint8_t* __restrict__ bin, size_t nbytes,
                                                          uint8_t* __restrict__ out) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(bin) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const size_t nvec = aligned ? nbytes / 16 : 0;
    const uint4* pv = reinterpret_cast<const uint4*>(bin);
    uint4* ov = reinterpret_cast<uint4*>(out);
    auto eq0 = [](uint32_t w) -> uint32_t {
        // per-byte (b == 0) ? 1 : 0
        uint32_t t = (w | ((w | 0x80808080u) - 0x01010101u)) & 0x80808080u;   // high bit set iff byte != 0
        return ((~t) & 0x80808080u) >> 7;
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldg(pv + i);
        ov[i] = make_uint4(eq0(v.x), eq0(v.y), eq0(v.z), eq0(v.w));
    }
    for (size_t i = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = bin[i] == 0;
}

// ---------------------------------------------------------------------------
// anti-aliasing Gaussian (scipy.ndimage.gaussian_filter, mode='mirror',
// truncate=4): correlate1d symmetric form  tmp = x0*w0; for j=-R..-1:
// tmp += (x[j] + x[-j]) * w[j]   -- axis 0 then axis 1, fp64.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int mirror_idx(int i, int n) {
    // scipy 'mirror' extension: d c b | a b c d | c b a   (period 2n-2)
    if (n == 1) return 0;
    const int period = 2 * n - 2;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

constexpr int kMaxGaussRadius = 63;   // 2R+1 <= 128: numpy pairwise_sum single block
__constant__ double c_gauss_w[2][kMaxGaussRadius + 1];   // [axis][0..R], w[0] = centre

// Persistent grid (a few blocks per SM) looping over (page, 32x8 tile): pages with <= 2 grey levels are
// skipped, so a batch of binarised pages pays one near-empty launch.  AXIS 0 reads the uint8 page,
// AXIS 1 the fp64 result of axis 0 (or the page when axis 0 is skipped).
// group_flag[g] = 1 iff some page of group g has more than two grey levels: the general-path kernels of a
// group of binarised pages leave on one load instead of scanning the level bitmaps of all its pages
__global__ void group_flags_kernel(const uint32_t* __restrict__ level_bits, int n, int group, int* __restrict__ flags) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * group >= n) return;
    int f = 0;
    for (int pg = g * group; pg < min(n, (g + 1) * group); ++pg) f |= level_count(level_bits + (size_t)pg * 8) > 2;
    flags[g] = f;
}

template <typename SRC, int AXIS>
__global__ void __launch_bounds__(256)
gauss1d_kernel(const SRC* __restrict__ src, size_t src_page_stride, double* __restrict__ dst, int H, int W, int radius,
               const uint32_t* __restrict__ level_bits, int page0, int pages, const int* __restrict__ group_flag) {
    if (!*group_flag) return;
    const double* wts = c_gauss_w[AXIS];
    const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
    for (int pg = 0; pg < pages; ++pg) {
        if (level_count(level_bits + (size_t)(page0 + pg) * 8) <= 2) continue;
        const SRC* sp = src + (size_t)pg * src_page_stride;
        double* dp = dst + (size_t)pg * H * W;
        for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
            const int x = (tile % tiles_x) * 32 + (threadIdx.x & 31);
            const int y = (tile / tiles_x) * 8 + (threadIdx.x >> 5);
            if (x >= W || y >= H) continue;
            auto at = [&](int d) -> double {
                if (AXIS == 0) return (double)sp[(size_t)mirror_idx(y + d, H) * W + x];
                return (double)sp[(size_t)y * W + mirror_idx(x + d, W)];
            };
            double tmp = __dmul_rn(at(0), wts[0]);
            for (int j = radius; j >= 1; --j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(-j), at(j)), wts[j]));
            dp[(size_t)y * W + x] = tmp;
        }
    }
}

// per-page min/max of the filtered plane (values >= 0: the bit patterns order like the doubles)
__global__ void __launch_bounds__(256)
minmax_f64_kernel(const double* __restrict__ planes, size_t n, unsigned long long* __restrict__ out /*[pages][2]*/,
                  const uint32_t* __restrict__ level_bits, int page0, const int* __restrict__ group_flag) {
    if (!*group_flag || level_count(level_bits + (size_t)(page0 + blockIdx.y) * 8) <= 2) return;
    const double* p = planes + (size_t)blockIdx.y * n;
    double lo = 1e300, hi = 0.0;      // idle threads must not win the unsigned-pattern max
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double v = p[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out + 2 * blockIdx.y, (unsigned long long)__double_as_longlong(lo));
        atomicMax(out + 2 * blockIdx.y + 1, (unsigned long long)__double_as_longlong(hi));
    }
}

__global__ void minmax_init_kernel(unsigned long long* out, int pages) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pages) { out[2 * i] = 0x7ff0000000000000ull; out[2 * i + 1] = 0ull; }
}

static int gauss_weights(double sigma, std::vector<double>& w) {
    // scipy _gaussian_kernel1d(sigma, 0, radius): exp(-0.5/sigma^2 * x^2) / sum, radius=int(4*sigma+0.5)
    const int radius = (int)(4.0 * sigma + 0.5);
    std::vector<double> phi(2 * radius + 1);
    const double sigma2 = sigma * sigma;
    double sum = 0.0;
    for (int i = -radius; i <= radius; ++i) {
        phi[i + radius] = exp(-0.5 / sigma2 * (double)(i * i));
    }
    // numpy's pairwise_sum for n <= 128: 8 strided accumulators, tree-combined, then the tail
    const int cnt = 2 * radius + 1;
    if (cnt < 8) {
        for (double v : phi) sum += v;
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = phi[j];
        int i = 8;
        for (; i < cnt - (cnt % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += phi[i + j];
        sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < cnt; ++i) sum += phi[i];
    }
    w.assign(radius + 1, 0.0);
    for (int j = 0; j <= radius; ++j) w[j] = phi[radius + j] / sum;
    return radius;
}

constexpr int kAaGroup = 16;     // pages whose fp64 anti-aliasing planes live in scratch at the same time

// d_image_f64 (internal, max_width pass): the image of the first rescale kept as fp64 `1 - v/255` planes
// [n][Hs][Ws] instead of the truncated uint8 image; n must not exceed kAaGroup then.
static int preprocess_impl(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int Hs,
                           int Ws, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary, double* d_image_f64) {
    if (n <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0) return set_err(ctx, PCS_ERR_ARG, "preprocess: bad shape");
    if (d_image_f64 && (d_image || n > kAaGroup)) return set_err(ctx, PCS_ERR_ARG, "preprocess: fp64 image output is per group");
    const bool want_image = d_image || d_image_f64;
    cudaStream_t st = ctx->stream;
    bool forked = false;
    const size_t page_px = (size_t)H * W;
    dim3 block(32, 8);
    if (want_image || d_binary) {
        // anti-aliasing parameters depend on the shapes only (skimage: sigma = (in/out - 1) / 2 per axis)
        const double fr = (double)H / (double)Hs, fc = (double)W / (double)Ws;
        const double sig[2] = {std::max(0.0, (fr - 1.0) / 2.0), std::max(0.0, (fc - 1.0) / 2.0)};
        const bool may_aa = want_image && (sig[0] > 1e-15 || sig[1] > 1e-15);
        std::vector<double> w0, w1;
        const int r0 = (want_image && sig[0] > 1e-15) ? gauss_weights(sig[0], w0) : -1;
        const int r1 = (want_image && sig[1] > 1e-15) ? gauss_weights(sig[1], w1) : -1;
        if (r0 > kMaxGaussRadius || r1 > kMaxGaussRadius)
            return set_err(ctx, PCS_ERR_ARG, "preprocess: anti-aliasing radius %d/%d exceeds %d", r0, r1, kMaxGaussRadius);
        const int group = std::min(n, kAaGroup);
        // two-level fast path: 16-byte aligned pages of a multiple of 32 bytes, scale factors up to 4
        const bool fast = d_image && (reinterpret_cast<uintptr_t>(d_grey) & 15) == 0 && page_px % 32 == 0 && fr <= 4.0 && fc <= 4.0 &&
                          (size_t)Hs * Ws < (size_t)1 << 31 && page_px < (size_t)1 << 32;
        const size_t bitmap_words = fast ? (page_px / 32 + 1 + 3) / 4 * 4 : 0;
        const int ngroups = (n + group - 1) / group;
        // level bitmaps, group flags, then (fast path) the sums of scan_pack_kernel and its verdict per page
        const size_t head_bits = (((size_t)n * 8 * 4 + (size_t)ngroups * 4 + 255) / 256) * 256;
        const unsigned scan_blocks = fast ? (unsigned)std::min<size_t>((size_t)ctx->sm_count * 4, (page_px / (1024 * kScanUnroll) + 7) / 8 + 1) : 0;
        const size_t head = head_bits + (((size_t)n * (16 + (size_t)scan_blocks * 3 * 8) + 255) / 256) * 256;
        const size_t mm_bytes = (((size_t)group * 16 + 255) / 256) * 256;
        const size_t bm_bytes = (((size_t)n * bitmap_words * 4 + 255) / 256) * 256;
        const size_t plane_bytes = may_aa ? (size_t)group * page_px * sizeof(double) : 0;
        PCS_TRY(scratch_reserve(ctx, head + mm_bytes + bm_bytes + 2 * plane_bytes + 256));
        uint32_t* d_bits = reinterpret_cast<uint32_t*>(ctx->scratch);
        int* d_gflags = reinterpret_cast<int*>(d_bits + (size_t)n * 8);
        int4* d_lv = reinterpret_cast<int4*>(reinterpret_cast<char*>(ctx->scratch) + head_bits);
        unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(d_lv + n);
        unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ctx->scratch) + head);
        uint32_t* d_bitmap = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ctx->scratch) + head + mm_bytes);
        double* t0 = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->scratch) + head + mm_bytes + bm_bytes);
        double* t1 = t0 + (size_t)group * page_px;
        if (want_image) {
            PCS_CUDA(ctx, cudaMemsetAsync(d_bits, 0, head_bits, st));
            if (fast) {
                scan_pack_kernel<<<dim3(scan_blocks, n), 256, 0, st>>>(d_grey, page_px, d_stats, d_bitmap, bitmap_words);
                PCS_LAUNCH_CHECK(ctx, "scan_pack_kernel");
                levels_from_stats_kernel<<<ngroups, 512, 0, st>>>(d_grey, page_px, d_stats, (int)scan_blocks, n, group, d_bits, d_lv, d_gflags);
                PCS_LAUNCH_CHECK(ctx, "levels_from_stats_kernel");
                // From here the two-level resampler runs on the main stream and everything for pages with more than two levels
                // (normally a row of launches that leave at once, ~3 us each) beside it on the side stream.  The two write
                // different pages.
                if (!ctx->aux_stream) {
                    PCS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
                    PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_aux_fork, cudaEventDisableTiming));
                    PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_aux_join, cudaEventDisableTiming));
                }
                PCS_CUDA(ctx, cudaEventRecord(ctx->ev_aux_fork, st));
                PCS_TRY(launch_resample_bits(ctx, d_bin, d_bin == d_grey ? 1 : 0, n, H, W, Hs, Ws, d_lv, d_bitmap, bitmap_words, d_image, d_binary));
                st = ctx->aux_stream;
                forked = true;
                PCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_aux_fork, 0));
                // the level sets of the pages that have more than two levels (every block of the others leaves at once)
                level_bits_kernel<<<dim3(74, n), 256, 0, st>>>(d_grey, page_px, d_bits, d_lv);
                PCS_LAUNCH_CHECK(ctx, "level_bits_kernel");
            } else {
                dim3 grid((unsigned)std::min<size_t>(296, (page_px / 16 + 255) / 256 + 1), n);
                level_bits_kernel<<<grid, 256, 0, st>>>(d_grey, page_px, d_bits, nullptr);
                PCS_LAUNCH_CHECK(ctx, "level_bits_kernel");
                group_flags_kernel<<<(ngroups + 63) / 64, 64, 0, st>>>(d_bits, n, group, d_gflags);
                PCS_LAUNCH_CHECK(ctx, "group_flags_kernel");
            }
            if (may_aa) {
                // pageable host -> constant: staged synchronously, ordered on the stream
                if (r0 >= 0) PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_gauss_w, w0.data(), w0.size() * 8, 0, cudaMemcpyHostToDevice, st));
                if (r1 >= 0)
                    PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_gauss_w, w1.data(), w1.size() * 8, sizeof(double) * (kMaxGaussRadius + 1),
                                                          cudaMemcpyHostToDevice, st));
            }
        }
        for (int p0 = 0; p0 < n; p0 += group) {
            const int m = std::min(group, n - p0);
            const double* planes = nullptr;
            if (may_aa) {
                // pages with <= 2 grey levels leave every one of these kernels in their first instruction
                const unsigned gfull = (unsigned)ctx->sm_count * 8;
                const uint8_t* src = d_grey + (size_t)p0 * page_px;
                if (r0 >= 0) {
                    gauss1d_kernel<uint8_t, 0><<<gfull, 256, 0, st>>>(src, page_px, t0, H, W, r0, d_bits, p0, m, d_gflags + p0 / group);
                    PCS_LAUNCH_CHECK(ctx, "gauss1d<axis 0>");
                    planes = t0;
                }
                if (r1 >= 0) {
                    if (planes) gauss1d_kernel<double, 1><<<gfull, 256, 0, st>>>(t0, page_px, t1, H, W, r1, d_bits, p0, m, d_gflags + p0 / group);
                    else gauss1d_kernel<uint8_t, 1><<<gfull, 256, 0, st>>>(src, page_px, t1, H, W, r1, d_bits, p0, m, d_gflags + p0 / group);
                    PCS_LAUNCH_CHECK(ctx, "gauss1d<axis 1>");
                    planes = t1;
                }
                minmax_init_kernel<<<1, 32, 0, st>>>(d_mm, m);
                PCS_LAUNCH_CHECK(ctx, "minmax_init_kernel");
                minmax_f64_kernel<<<dim3(148, m), 256, 0, st>>>(planes, page_px, d_mm, d_bits, p0, d_gflags + p0 / group);
                PCS_LAUNCH_CHECK(ctx, "minmax_f64_kernel");
            }
            const long long tiles = (long long)((Ws + 31) / 32) * ((Hs + 7) / 8) * m;
            const unsigned rgrid = (unsigned)std::min<long long>(tiles, (long long)ctx->sm_count * 8);
            resample_kernel<<<rgrid, block, 0, st>>>(d_grey, d_bin, H, W, Hs, Ws, d_bits, planes, d_mm, p0, d_image, d_binary, fast ? 1 : 0,
                                                     want_image ? d_gflags + p0 / group : nullptr, m, d_image_f64);
            PCS_LAUNCH_CHECK(ctx, "resample_kernel");
        }
        if (forked) {
            PCS_CUDA(ctx, cudaEventRecord(ctx->ev_aux_join, st));
            st = ctx->stream;
            PCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_aux_join, 0));
        }
    }
    if (d_orig_binary) {
        const size_t nbytes = (size_t)n * page_px;
        orig_binary_kernel<<<(unsigned)std::min<size_t>(148 * 8, (nbytes / 16 + 255) / 256 + 1), 256, 0, st>>>(d_bin, nbytes,
                                                                                                          d_orig_binary);
        PCS_LAUNCH_CHECK(ctx, "orig_binary_kernel");
    }
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// Bit-packed pages.  A binarised page is one bit per pixel by nature; the two-level fast path above packs the uint8 page
// into exactly that form before it resamples.  A caller that already holds the page packed (flat over the page, pixel i
// = bit i & 31 of word i >> 5, i.e. numpy.packbits(..., bitorder='little') viewed as little-endian words; bit 0 = a
// pixel of value level0, bit 1 = level1) skips the 8.7 MB page altogether.  Results are those of pcs_preprocess on the
// uint8 page `bit ? level1 : level0` used as grey and binary page.
// ---------------------------------------------------------------------------
int launch_preprocess_bits(pcs_ctx* ctx, const uint32_t* d_bitmap, size_t bitmap_words, int n, int H, int W, int level0, int level1,
                           int Hs, int Ws, uint8_t* d_image, uint8_t* d_binary) {
    if (n <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: bad shape");
    if (level0 < 0 || level0 > 255 || level1 < 0 || level1 > 255 || level0 == level1)
        return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: the two grey levels must differ and lie in 0..255");
    const size_t page_px = (size_t)H * W;
    if (bitmap_words < page_px / 32 + 1) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: %zu words per page needed (incl. one pad word)", page_px / 32 + 1);
    if ((double)H / Hs > 4.0 || (double)W / Ws > 4.0) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: scale factors above 4 are not supported");
    if (!d_image) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: the image output is required");
    PCS_TRY(scratch_reserve(ctx, (size_t)n * 16 + 256));
    int4* d_lv = reinterpret_cast<int4*>(ctx->scratch);
    set_levels_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_lv, n, level0, level1);
    PCS_LAUNCH_CHECK(ctx, "set_levels_kernel");
    return launch_resample_bits(ctx, nullptr, 1, n, H, W, Hs, Ws, d_lv, d_bitmap, bitmap_words, d_image, d_binary);
}

// uint8 planes <-> flat bit planes (pixel i = bit i & 31 of word i >> 5; a non-zero byte is a set bit, a set bit is byte 1).
// grid = (blocks, pages); words_per_page may exceed ceil(npix / 32) (padding is written as zero).
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t* __restrict__ src, size_t npix, uint32_t* __restrict__ dst, size_t words_per_page) {
    const uint8_t* p = src + (size_t)blockIdx.y * npix;
    uint32_t* o = dst + (size_t)blockIdx.y * words_per_page;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t w0 = warp0 * 32; w0 < words_per_page; w0 += nwarps * 32) {        // a warp packs 32 words = 1024 pixels per step
        uint32_t mine = 0;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const size_t i = (w0 + k) * 32 + lane;
            const uint32_t m = __ballot_sync(0xffffffffu, i < npix && p[i] != 0);
            if (lane == k) mine = m;
        }
        if (w0 + lane < words_per_page) o[w0 + lane] = mine;
    }
}

__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint32_t* __restrict__ src, size_t words_per_page, size_t npix, uint8_t* __restrict__ dst) {
    const uint32_t* p = src + (size_t)blockIdx.y * words_per_page;
    uint8_t* o = dst + (size_t)blockIdx.y * npix;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x)
        o[i] = (uint8_t)((__ldg(p + (i >> 5)) >> (i & 31)) & 1u);
}

int launch_pack_bits(pcs_ctx* ctx, const uint8_t* d_src, int n, size_t npix, uint32_t* d_dst, size_t words_per_page) {
    if (n <= 0 || !npix || words_per_page < (npix + 31) / 32) return set_err(ctx, PCS_ERR_ARG, "pack_bits: bad shape");
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 4, (words_per_page + 255) / 256);
    pack_bits_kernel<<<dim3(blocks, n), 256, 0, ctx->stream>>>(d_src, npix, d_dst, words_per_page);
    PCS_LAUNCH_CHECK(ctx, "pack_bits_kernel");
    return PCS_OK;
}

int launch_unpack_bits(pcs_ctx* ctx, const uint32_t* d_src, int n, size_t words_per_page, size_t npix, uint8_t* d_dst) {
    if (n <= 0 || !npix || words_per_page < (npix + 31) / 32) return set_err(ctx, PCS_ERR_ARG, "unpack_bits: bad shape");
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, (npix + 255) / 256);
    unpack_bits_kernel<<<dim3(blocks, n), 256, 0, ctx->stream>>>(d_src, words_per_page, npix, d_dst);
    PCS_LAUNCH_CHECK(ctx, "unpack_bits_kernel");
    return PCS_OK;
}

int launch_preprocess(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int Hs,
                      int Ws, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary) {
    return preprocess_impl(ctx, d_grey, d_bin, n, H, W, Hs, Ws, d_image, d_binary, d_orig_binary, nullptr);
}

// ---------------------------------------------------------------------------
// max_width second pass (dataset.py:139-143):  bin = scale_binary(bin, n_scale);  img = scale_image(img, bin.shape)
// on the fp64 image `1 - v/255` of the first pass; anti-aliased when that image has more than two distinct
// values (`len(np.unique(img)) > 2`), clipped to the min/max of the image that is warped, then
// (img * 255).astype(uint8).  A rarely used option: plain per-pixel kernels.
// ---------------------------------------------------------------------------
constexpr unsigned long long kNoValue = 0xffffffffffffffffull;     // a NaN pattern: never a pixel value

__global__ void distinct_init_kernel(unsigned long long* slots, int* flags, int pages) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pages) { slots[i] = kNoValue; flags[i] = 0; }
}

// flags[page] = 1 iff the plane holds more than two distinct values
__global__ void __launch_bounds__(256)
distinct_kernel(const double* __restrict__ planes, size_t npx, unsigned long long* __restrict__ slots, int* __restrict__ flags) {
    const int page = blockIdx.y;
    const double* p = planes + (size_t)page * npx;
    const unsigned long long v0 = (unsigned long long)__double_as_longlong(p[0]);
    unsigned long long seen = v0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = (unsigned long long)__double_as_longlong(p[i]);
        if (v == v0 || v == seen) continue;
        seen = v;
        const unsigned long long old = atomicCAS(slots + page, kNoValue, v);
        if (old != kNoValue && old != v) flags[page] = 1;
    }
}

template <int AXIS>
__global__ void __launch_bounds__(256)
gauss_plane_kernel(const double* __restrict__ src, double* __restrict__ dst, int H, int W, int radius,
                   const double* __restrict__ wts, const int* __restrict__ flags) {
    const int page = blockIdx.y;
    if (!flags[page]) return;
    const double* sp = src + (size_t)page * H * W;
    double* dp = dst + (size_t)page * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)H * W; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        auto at = [&](int d) -> double {
            if (AXIS == 0) return sp[(size_t)mirror_idx(y + d, H) * W + x];
            return sp[(size_t)y * W + mirror_idx(x + d, W)];
        };
        double tmp = __dmul_rn(at(0), wts[0]);
        for (int j = radius; j >= 1; --j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(-j), at(j)), wts[j]));
        dp[i] = tmp;
    }
}

// min / max of the plane that is warped (the filtered one when flags[page], else the raw one); values >= 0
__global__ void __launch_bounds__(256)
minmax_plane_kernel(const double* __restrict__ raw, const double* __restrict__ filt, size_t npx, const int* __restrict__ flags,
                    unsigned long long* __restrict__ out /*[pages][2], initialised*/) {
    const int page = blockIdx.y;
    const double* p = (flags[page] ? filt : raw) + (size_t)page * npx;
    double lo = 1e300, hi = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        const double v = p[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out + 2 * page, (unsigned long long)__double_as_longlong(lo));
        atomicMax(out + 2 * page + 1, (unsigned long long)__double_as_longlong(hi));
    }
}

__global__ void __launch_bounds__(256)
resample_plane_kernel(const double* __restrict__ raw, const double* __restrict__ filt, const int* __restrict__ flags,
                      const unsigned long long* __restrict__ mm, int H, int W, int Ho, int Wo, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, page = blockIdx.z;
    if (x >= Wo || y >= Ho) return;
    const double* g = (flags[page] ? filt : raw) + (size_t)page * H * W;
    const double fr = __ddiv_rn((double)H, (double)Ho), fc = __ddiv_rn((double)W, (double)Wo);
    const double pr = __dadd_rn(__dmul_rn(fr, (double)y), __dsub_rn(__dmul_rn(0.5, fr), 0.5));
    const double pc = __dadd_rn(__dmul_rn(fc, (double)x), __dsub_rn(__dmul_rn(0.5, fc), 0.5));
    const double rf = floor(pr), cf = floor(pc);
    const double xr = __dsub_rn(pr, rf), xc = __dsub_rn(pc, cf);
    int cols[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cols[k] = reflect_coord((long long)cf - 1 + k, W);
    double frow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t ro = (size_t)reflect_coord((long long)rf - 1 + k, H) * W;
        frow[k] = cubic_rn(xc, g[ro + cols[0]], g[ro + cols[1]], g[ro + cols[2]], g[ro + cols[3]]);
    }
    double v = cubic_rn(xr, frow[0], frow[1], frow[2], frow[3]);
    v = fmin(fmax(v, __longlong_as_double((long long)mm[2 * page])), __longlong_as_double((long long)mm[2 * page + 1]));
    out[(size_t)page * Ho * Wo + (size_t)y * Wo + x] = (uint8_t)(int)__dmul_rn(v, 255.0);      // (img * 255).astype(uint8)
}

int launch_preprocess_max_width(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int H1, int W1,
                                int H2, int W2, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary) {
    if (n <= 0 || H2 <= 0 || W2 <= 0 || H1 <= 0 || W1 <= 0) return set_err(ctx, PCS_ERR_ARG, "preprocess(max_width): bad shape");
    if (!d_image || !d_binary) return set_err(ctx, PCS_ERR_ARG, "preprocess(max_width): image and binary outputs are required");
    cudaStream_t st = ctx->stream;
    const size_t px1 = (size_t)H1 * W1;
    const int group = std::min(n, kAaGroup);
    const double f2[2] = {(double)H1 / (double)H2, (double)W1 / (double)W2};
    std::vector<double> w[2];
    int rad[2];
    for (int a = 0; a < 2; ++a) {
        const double sg = std::max(0.0, (f2[a] - 1.0) / 2.0);
        rad[a] = sg > 1e-15 ? gauss_weights(sg, w[a]) : -1;
        if (rad[a] > kMaxGaussRadius) return set_err(ctx, PCS_ERR_ARG, "preprocess(max_width): anti-aliasing radius %d", rad[a]);
    }
    // second scratch (the first pass uses ctx->scratch): fp64 planes img1, t0, t1; bin1; small control words
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t need = 3 * al((size_t)group * px1 * 8) + al((size_t)group * px1) + al((size_t)group * 32) + al(2 * 64 * 8) + 256;
    if (need > ctx->scratch2_bytes) {
        PCS_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->scratch2) cudaFree(ctx->scratch2);
        ctx->scratch2 = nullptr; ctx->scratch2_bytes = 0;
        if (cudaMalloc(&ctx->scratch2, need) != cudaSuccess) { cudaGetLastError(); return set_err(ctx, PCS_ERR_NOMEM, "cudaMalloc of %zu bytes failed", need); }
        ctx->scratch2_bytes = need;
    }
    char* q = reinterpret_cast<char*>(ctx->scratch2);
    double* img1 = reinterpret_cast<double*>(q); q += al((size_t)group * px1 * 8);
    double* t0 = reinterpret_cast<double*>(q); q += al((size_t)group * px1 * 8);
    double* t1 = reinterpret_cast<double*>(q); q += al((size_t)group * px1 * 8);
    uint8_t* bin1 = reinterpret_cast<uint8_t*>(q); q += al((size_t)group * px1);
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(q);           // [group] second value, then [group][2] min/max
    unsigned long long* mm = slots + group;
    int* flags = reinterpret_cast<int*>(mm + 2 * group); q += al((size_t)group * 32);
    double* d_w = reinterpret_cast<double*>(q);
    for (int a = 0; a < 2; ++a)
        if (rad[a] >= 0) PCS_CUDA(ctx, cudaMemcpyAsync(d_w + 64 * a, w[a].data(), w[a].size() * 8, cudaMemcpyHostToDevice, st));
    for (int p0 = 0; p0 < n; p0 += group) {
        const int m = std::min(group, n - p0);
        const size_t so = (size_t)p0 * H * W;
        PCS_TRY(preprocess_impl(ctx, d_grey + so, d_bin + so, m, H, W, H1, W1, nullptr, bin1, nullptr, img1));
        PCS_TRY(launch_resize_nearest(ctx, bin1, m, H1, W1, d_binary + (size_t)p0 * H2 * W2, H2, W2));
        distinct_init_kernel<<<1, 64, 0, st>>>(slots, flags, m);
        minmax_init_kernel<<<1, 32, 0, st>>>(mm, m);
        distinct_kernel<<<dim3(64, m), 256, 0, st>>>(img1, px1, slots, flags);
        PCS_LAUNCH_CHECK(ctx, "distinct_kernel");
        const double* filt = img1;
        if (rad[0] >= 0) { gauss_plane_kernel<0><<<dim3(296, m), 256, 0, st>>>(img1, t0, H1, W1, rad[0], d_w, flags); filt = t0; }
        if (rad[1] >= 0) { gauss_plane_kernel<1><<<dim3(296, m), 256, 0, st>>>(filt, t1, H1, W1, rad[1], d_w + 64, flags); filt = t1; }
        PCS_LAUNCH_CHECK(ctx, "gauss_plane_kernel");
        minmax_plane_kernel<<<dim3(148, m), 256, 0, st>>>(img1, filt, px1, flags, mm);
        PCS_LAUNCH_CHECK(ctx, "minmax_plane_kernel");
        resample_plane_kernel<<<dim3((W2 + 31) / 32, (H2 + 7) / 8, m), dim3(32, 8), 0, st>>>(img1, filt, flags, mm, H1, W1, H2, W2,
                                                                                             d_image + (size_t)p0 * H2 * W2);
        PCS_LAUNCH_CHECK(ctx, "resample_plane_kernel");
    }
    if (d_orig_binary) {
        const size_t nbytes = (size_t)n * H * W;
        orig_binary_kernel<<<(unsigned)std::min<size_t>(148 * 8, (nbytes / 16 + 255) / 256 + 1), 256, 0, st>>>(d_bin, nbytes, d_orig_binary);
        PCS_LAUNCH_CHECK(ctx, "orig_binary_kernel");
    }
    return PCS_OK;
}

// preserving_resize (util.py:21-29): order-0 resize of uint8 planes
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst, int Ho, int Wo, double fr,
                      double fc, int vec) {
    // 16 output pixels per thread (one 16-byte store when the row is aligned); fr = H / Ho, fc = W / Wo
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 16;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x0 >= Wo || y >= Ho) return;
    const int page = blockIdx.z;
    const double r = __dadd_rn(__dmul_rn(fr, (double)y), __dsub_rn(__dmul_rn(0.5, fr), 0.5));
    const int ri = reflect_coord((long long)round(r), H);
    const uint8_t* srow = src + (size_t)page * H * W + (size_t)ri * W;
    uint8_t* drow = dst + (size_t)page * Ho * Wo + (size_t)y * Wo;
    const double c_off = __dsub_rn(__dmul_rn(0.5, fc), 0.5);
    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int x = x0 + k;
        if (x < Wo) {
            const double c = __dadd_rn(__dmul_rn(fc, (double)x), c_off);
            const unsigned v = __ldg(srow + reflect_coord((long long)round(c), W));
            if (vec) w[k >> 2] |= v << ((k & 3) * 8);
            else drow[x] = (uint8_t)v;
        }
    }
    if (vec) {
        if (x0 + 16 <= Wo) *reinterpret_cast<uint4*>(drow + x0) = make_uint4(w[0], w[1], w[2], w[3]);
        else
            for (int k = 0; x0 + k < Wo; ++k) drow[x0 + k] = (uint8_t)((w[k >> 2] >> ((k & 3) * 8)) & 0xffu);
    }
}

int launch_resize_nearest(pcs_ctx* ctx, const uint8_t* d_src, int n, int H, int W, uint8_t* d_dst, int Ho, int Wo) {
    if (n <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return set_err(ctx, PCS_ERR_ARG, "resize_nearest: bad shape");
    dim3 grid((Wo + 511) / 512, (Ho + 7) / 8, n), block(32, 8);
    const int vec = (Wo % 16 == 0) && (reinterpret_cast<uintptr_t>(d_dst) % 16 == 0);
    resize_nearest_kernel<<<grid, block, 0, ctx->stream>>>(d_src, H, W, d_dst, Ho, Wo, (double)H / (double)Ho,
                                                           (double)W / (double)Wo, vec);
    PCS_LAUNCH_CHECK(ctx, "resize_nearest_kernel");
    return PCS_OK;
}

}  // namespace pcs
