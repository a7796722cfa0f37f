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
                                                         uint32_t* __restrict__ bits /*[n][8]*/) {
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
// the normal one).  One streaming pass over the page produces the level bitmap AND a 1-bit-per-pixel
// plane  bit(i) = (page[i] != page[0]);  the resampler then never touches the 8.7 MB page again: a
// source value is  bit ? other level : page[0].
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ne_bits4(uint32_t w, uint32_t ref4) {
    const uint32_t x = w ^ ref4;
    const uint32_t t = (x | ((x | 0x80808080u) - 0x01010101u)) & 0x80808080u;      // high bit set iff byte != 0
    return (((t >> 7) * 0x01020408u) >> 24) & 0xfu;                                 // byte k -> bit k
}
__device__ __forceinline__ uint32_t ne_bits16(const uint4 v, uint32_t ref4) {
    return ne_bits4(v.x, ref4) | (ne_bits4(v.y, ref4) << 4) | (ne_bits4(v.z, ref4) << 8) | (ne_bits4(v.w, ref4) << 12);
}

// grid = (blocks, pages); every thread turns 32 page bytes into one bitmap word.  Requires 16-byte aligned
// pages whose size is a multiple of 32 bytes (checked by the host; other shapes take the general kernels).
__global__ void __launch_bounds__(256) scan_pack_kernel(const uint8_t* __restrict__ src, size_t page_bytes,
                                                        uint32_t* __restrict__ bits /*[n][8]*/,
                                                        uint32_t* __restrict__ bitmap, size_t bitmap_words /*per page, padded*/) {
    __shared__ uint32_t s_bits[8];
    if (threadIdx.x < 8) s_bits[threadIdx.x] = 0;
    __syncthreads();
    const int page = blockIdx.y;
    const uint8_t* p = src + (size_t)page * page_bytes;
    const uint4* pv = reinterpret_cast<const uint4*>(p);
    uint32_t* bm = bitmap + (size_t)page * bitmap_words;
    const uint32_t ref4 = (uint32_t)__ldg(p) * 0x01010101u;
    const size_t nwords = page_bytes / 32;
    auto mark = [&](uint32_t val) {
        if (!((s_bits[val >> 5] >> (val & 31)) & 1u)) atomicOr(&s_bits[val >> 5], 1u << (val & 31));
    };
    // high bit of every byte of w that is NOT zero
    auto nz = [](uint32_t w) -> uint32_t { return (w | ((w | 0x80808080u) - 0x01010101u)) & 0x80808080u; };
    // per-thread cache of the second value: words made of {page[0], oth} bytes only need no set update
    uint32_t oth4 = ref4;
    bool have_oth = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(pv + 2 * i), b = __ldg(pv + 2 * i + 1);
        const uint32_t words[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t wv = words[k];
            const uint32_t dif = nz(wv ^ ref4);                         // bytes that differ from page[0]
            out |= ((((dif >> 7) * 0x01020408u) >> 24) & 0xfu) << (4 * k);
            if (dif & nz(wv ^ oth4)) {                                  // some byte is neither page[0] nor oth
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t v = (wv >> (8 * q)) & 0xff;
                    if (v != (ref4 & 0xff)) { mark(v); if (!have_oth) { oth4 = v * 0x01010101u; have_oth = true; } }
                }
            }
        }
        bm[i] = out;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { mark(ref4 & 0xff); bm[nwords] = 0; }       // pad word read by the funnel shift
    __syncthreads();
    if (threadIdx.x < 8 && s_bits[threadIdx.x]) atomicOr(&bits[page * 8 + threadIdx.x], s_bits[threadIdx.x]);
}

constexpr int RB_T = 32, RB_TY = 128;            // output tile of the fast resampler: 32 x 128 (256 threads, 16 rows each): the
                                                 // per-block setup (coordinates, 16-pattern table) is a quarter of the work at 64 rows
constexpr int RB_RW = RB_TY / 32;                // warps that set up the rows
constexpr int RB_MAX_SPAN = 4 * RB_TY + 8;       // staged source rows for scale factors up to 4
constexpr int RB_ROW_WORDS = 8;                  // staged words per source row: (4*32 + 8 + 31 + 31) / 32

// grid = (ceil(Ws/32), ceil(Hs/64), pages).  Pages with more than two grey levels return at once (the
// general kernel below handles them).  Bit-identical to the general kernel:
//   * the horizontal cubic of a two-level row has only 16 possible operand patterns per output column;
//     they are evaluated once per block with the same fp64 operation order (cubic_rn) and looked up;
//   * a pixel whose 4x4 neighbourhood is all one level v gets cubic(v,v,v,v) = v exactly (every
//     intermediate of cubic_rn is an exact small integer or zero), so warps over blank paper or solid ink
//     store the precomputed constant and skip the fp64 arithmetic;
//   * the vertical cubic, clip and the (1 - v/255) * 255 truncation are otherwise unchanged.
__global__ void __launch_bounds__(256)
resample_bits_kernel(const uint8_t* __restrict__ grey, const uint8_t* __restrict__ bin, int bin_is_grey, int H, int W,
                     int Hs, int Ws, const uint32_t* __restrict__ level_bits, const uint32_t* __restrict__ bitmap,
                     size_t bitmap_words, uint8_t* __restrict__ image_out, uint8_t* __restrict__ binary_out, int va_fixed /* >= 0: the level of a
                     zero bit (packed input pages, no uint8 page to read it from) */) {
    __shared__ double s_lut[16][RB_T];            // [pattern][column]: a warp reads 32 consecutive doubles (no bank conflicts)
    __shared__ double s_cfrac[RB_T], s_rfrac[RB_TY];                  // fractional sampling offsets of columns / rows
    __shared__ uint32_t s_bm[RB_MAX_SPAN][RB_ROW_WORDS];
    __shared__ __align__(16) int s_ctap[RB_T][4], s_rtap[RB_TY][4];   // the four (reflected) source columns / rows
    __shared__ int s_cnn[RB_T], s_rnn[RB_TY], s_off[RB_MAX_SPAN], s_rng[2 + 2 * RB_RW];
    // bit address (within the flat staged bitmap) of column 0 of the four tap rows / the nearest row of every output
    // row: the per-pixel work is then one add, one shift and a funnel shift per tap row
    __shared__ __align__(16) int s_tapbit[RB_TY][4];
    __shared__ int s_nnbit[RB_TY];
    const int page = blockIdx.z;
    const uint32_t* bits = level_bits + (size_t)page * 8;
    if (level_count(bits) > 2) return;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    if (wrp < 1 + RB_RW) {                        // warp 0: the 32 columns, warps 1..: the rows of the tile
        const int n_in = wrp == 0 ? W : H, n_out = wrp == 0 ? Ws : Hs;
        const int o = min(wrp == 0 ? blockIdx.x * RB_T + lane : blockIdx.y * RB_TY + (wrp - 1) * 32 + lane, n_out - 1);   // replicate past the edge
        const double f = __ddiv_rn((double)n_in, (double)n_out);
        const double pc = __dadd_rn(__dmul_rn(f, (double)o), __dsub_rn(__dmul_rn(0.5, f), 0.5));
        const int nn = reflect_coord((long long)round(pc), n_in);        // order 0: C round(), then reflect
        const double pf = floor(pc);
        const int ti = wrp == 0 ? lane : (wrp - 1) * 32 + lane;
        (wrp == 0 ? s_cfrac : s_rfrac)[ti] = __dsub_rn(pc, pf);
        (wrp == 0 ? s_cnn : s_rnn)[ti] = nn;
        int lo = nn, hi = nn;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = reflect_coord((long long)pf - 1 + k, n_in);
            (wrp == 0 ? s_ctap : s_rtap)[ti][k] = c;
            lo = min(lo, c); hi = max(hi, c);
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_rng[2 * wrp] = lo; s_rng[2 * wrp + 1] = hi; }
    }
    __syncthreads();
    const int cmin = s_rng[0], cmax = s_rng[1];
    int rmin = s_rng[2], rmax = s_rng[3];
#pragma unroll
    for (int k = 1; k < RB_RW; ++k) { rmin = min(rmin, s_rng[2 + 2 * k]); rmax = max(rmax, s_rng[3 + 2 * k]); }
    const int nrows = rmax - rmin + 1;
    if (nrows > RB_MAX_SPAN || cmax - cmin + 1 + 62 > RB_ROW_WORDS * 32) { __trap(); }      // host guarantees scale <= 4
    const uint32_t* bm = bitmap + (size_t)page * bitmap_words;
    const size_t last_word = (size_t)H * W / 32;
    for (int i = tid; i < nrows * RB_ROW_WORDS; i += 256) {
        const int r = i / RB_ROW_WORDS, wq = i - r * RB_ROW_WORDS;
        const size_t b0 = (size_t)(rmin + r) * W + cmin;
        const size_t w0 = (b0 >> 5) + wq;
        s_bm[r][wq] = w0 <= last_word ? __ldg(bm + w0) : 0u;
        if (wq == 0) s_off[r] = (int)(b0 & 31) - cmin;        // bit (r, c) sits at bit s_off[r] + c of the staged row
    }
    __syncthreads();
    if (tid < RB_TY) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rl = s_rtap[tid][k] - rmin;
            s_tapbit[tid][k] = rl * (RB_ROW_WORDS * 32) + s_off[rl];
        }
        const int rl = s_rnn[tid] - rmin;
        s_nnbit[tid] = rl * (RB_ROW_WORDS * 32) + s_off[rl];
    }
    // the two levels: a = page[0] (bit 0), b = the other one (bit 1)
    int l0 = -1, l1 = -1;
#pragma unroll
    for (int wv = 7; wv >= 0; --wv) {
        const uint32_t m = bits[wv];
        if (m) {
            const int hi = wv * 32 + 31 - __clz(m), lo = wv * 32 + __ffs(m) - 1;
            if (l1 < 0) l1 = hi;
            l0 = lo;
        }
    }
    const int va = va_fixed >= 0 ? va_fixed : (int)__ldg(grey + (size_t)page * H * W);
    const int vb = va == l0 ? l1 : l0;
    const double vmin = (double)l0, vmax = (double)l1;
    if (image_out) {
        const double fa = (double)va, fb = (double)vb;
        for (int i = tid; i < RB_T * 16; i += 256) {
            const int c = i & 31, pat = i >> 5;
            s_lut[pat][c] = cubic_rn(s_cfrac[c], (pat & 1) ? fb : fa, (pat & 2) ? fb : fa, (pat & 4) ? fb : fa, (pat & 8) ? fb : fa);
        }
    }
    __syncthreads();
    const int x = blockIdx.x * RB_T + lane;
    const uint32_t* bmflat = &s_bm[0][0];
    auto bit_at = [&](int rowbit, int c) -> uint32_t {          // rowbit: flat bit address of column 0 of a staged row
        const int b = rowbit + c;
        return (bmflat[b >> 5] >> (b & 31)) & 1u;
    };
    auto finish = [&](double v) -> uint8_t {          // clip=True, then img = 1.0 - v/255 ; (img*255).astype(uint8)
        v = fmin(fmax(v, vmin), vmax);
        return (uint8_t)(int)__dmul_rn(__dsub_rn(1.0, __ddiv_rn(v, 255.0)), 255.0);
    };
    const uint8_t out_a = finish((double)va), out_b = finish((double)vb);
    const int c0 = s_ctap[lane][0], c1 = s_ctap[lane][1], c2 = s_ctap[lane][2], c3 = s_ctap[lane][3];
    const bool consec = c1 == c0 + 1 && c2 == c0 + 2 && c3 == c0 + 3;
    const int nnc = s_cnn[lane];
    const bool xok = x < Ws;
    const uint8_t bin_a = va == 0 ? 1 : 0, bin_b = vb == 0 ? 1 : 0;   // bin = (1.0 - NN(binary/255 or binary)).astype(uint8)
    const int y_first = blockIdx.y * RB_TY + wrp;
    size_t dst = (size_t)page * Hs * Ws + (size_t)y_first * Ws + x;
    const size_t dst_step = (size_t)8 * Ws;
#pragma unroll 1
    for (int ty = wrp; ty < RB_TY && blockIdx.y * RB_TY + ty < Hs; ty += 8, dst += dst_step) {     // warp-uniform bounds
        if (binary_out && xok) {
            uint8_t v;
            if (bin_is_grey) v = bit_at(s_nnbit[ty], nnc) ? bin_b : bin_a;
            else v = bin[(size_t)page * H * W + (size_t)s_rnn[ty] * W + nnc] == 0 ? 1 : 0;
            binary_out[dst] = v;
        }
        if (image_out) {
            const int4 rr = *reinterpret_cast<const int4*>(s_tapbit[ty]);
            const int rb[4] = {rr.x, rr.y, rr.z, rr.w};
            uint32_t pat[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (consec) {
                    const int b = rb[k] + c0;
                    pat[k] = __funnelshift_r(bmflat[b >> 5], bmflat[(b >> 5) + 1], b & 31) & 15u;
                } else {
                    pat[k] = bit_at(rb[k], c0) | (bit_at(rb[k], c1) << 1) | (bit_at(rb[k], c2) << 2) | (bit_at(rb[k], c3) << 3);
                }
            }
            const bool all_a = (pat[0] | pat[1] | pat[2] | pat[3]) == 0u;
            const bool all_b = (pat[0] & pat[1] & pat[2] & pat[3]) == 15u;
            uint8_t o = all_a ? out_a : out_b;
            if (!__all_sync(0xffffffffu, all_a || all_b)) {
                const double v = cubic_rn(s_rfrac[ty], s_lut[pat[0]][lane], s_lut[pat[1]][lane], s_lut[pat[2]][lane], s_lut[pat[3]][lane]);
                o = finish(v);
            }
            if (xok) image_out[dst] = o;
        }
    }
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
        const bool fast = d_image && (reinterpret_cast<uintptr_t>(d_grey) & 15) == 0 && page_px % 32 == 0 && fr <= 4.0 && fc <= 4.0;
        const size_t bitmap_words = fast ? (page_px / 32 + 1 + 3) / 4 * 4 : 0;
        const int ngroups = (n + group - 1) / group;
        const size_t head = (((size_t)n * 8 * 4 + (size_t)ngroups * 4 + 255) / 256) * 256;     // level bitmaps, then group flags
        const size_t mm_bytes = (((size_t)group * 16 + 255) / 256) * 256;
        const size_t bm_bytes = (((size_t)n * bitmap_words * 4 + 255) / 256) * 256;
        const size_t plane_bytes = may_aa ? (size_t)group * page_px * sizeof(double) : 0;
        PCS_TRY(scratch_reserve(ctx, head + mm_bytes + bm_bytes + 2 * plane_bytes + 256));
        uint32_t* d_bits = reinterpret_cast<uint32_t*>(ctx->scratch);
        int* d_gflags = reinterpret_cast<int*>(d_bits + (size_t)n * 8);
        unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ctx->scratch) + head);
        uint32_t* d_bitmap = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ctx->scratch) + head + mm_bytes);
        double* t0 = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->scratch) + head + mm_bytes + bm_bytes);
        double* t1 = t0 + (size_t)group * page_px;
        if (want_image) {
            PCS_CUDA(ctx, cudaMemsetAsync(d_bits, 0, (size_t)n * 8 * sizeof(uint32_t), st));
            if (fast) {
                dim3 grid((unsigned)std::min<size_t>((size_t)ctx->sm_count * 4, (page_px / 32 + 255) / 256), n);
                scan_pack_kernel<<<grid, 256, 0, st>>>(d_grey, page_px, d_bits, d_bitmap, bitmap_words);
                PCS_LAUNCH_CHECK(ctx, "scan_pack_kernel");
                dim3 rgrid((Ws + RB_T - 1) / RB_T, (Hs + RB_TY - 1) / RB_TY, n);
                resample_bits_kernel<<<rgrid, 256, 0, st>>>(d_grey, d_bin, d_bin == d_grey ? 1 : 0, H, W, Hs, Ws, d_bits, d_bitmap,
                                                            bitmap_words, d_image, d_binary, -1);
                PCS_LAUNCH_CHECK(ctx, "resample_bits_kernel");
            } else {
                dim3 grid((unsigned)std::min<size_t>(296, (page_px / 16 + 255) / 256 + 1), n);
                level_bits_kernel<<<grid, 256, 0, st>>>(d_grey, page_px, d_bits);
                PCS_LAUNCH_CHECK(ctx, "level_bits_kernel");
            }
            group_flags_kernel<<<(ngroups + 63) / 64, 64, 0, st>>>(d_bits, n, group, d_gflags);
            PCS_LAUNCH_CHECK(ctx, "group_flags_kernel");
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
__global__ void set_levels_kernel(uint32_t* __restrict__ bits /*[n][8]*/, int n, int l0, int l1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 8) return;
    const int w = i & 7;
    bits[i] = ((l0 >> 5) == w ? 1u << (l0 & 31) : 0u) | ((l1 >> 5) == w ? 1u << (l1 & 31) : 0u);
}

int launch_preprocess_bits(pcs_ctx* ctx, const uint32_t* d_bitmap, size_t bitmap_words, int n, int H, int W, int level0, int level1,
                           int Hs, int Ws, uint8_t* d_image, uint8_t* d_binary) {
    if (n <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: bad shape");
    if (level0 < 0 || level0 > 255 || level1 < 0 || level1 > 255 || level0 == level1)
        return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: the two grey levels must differ and lie in 0..255");
    const size_t page_px = (size_t)H * W;
    if (bitmap_words < page_px / 32 + 1) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: %zu words per page needed (incl. one pad word)", page_px / 32 + 1);
    if ((double)H / Hs > 4.0 || (double)W / Ws > 4.0) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: scale factors above 4 are not supported");
    if (!d_image) return set_err(ctx, PCS_ERR_ARG, "preprocess_bits: the image output is required");
    PCS_TRY(scratch_reserve(ctx, (size_t)n * 8 * 4 + 256));
    uint32_t* d_bits = reinterpret_cast<uint32_t*>(ctx->scratch);
    set_levels_kernel<<<(n * 8 + 255) / 256, 256, 0, ctx->stream>>>(d_bits, n, level0, level1);
    PCS_LAUNCH_CHECK(ctx, "set_levels_kernel");
    dim3 rgrid((Ws + RB_T - 1) / RB_T, (Hs + RB_TY - 1) / RB_TY, n);
    resample_bits_kernel<<<rgrid, 256, 0, ctx->stream>>>(nullptr, nullptr, 1, H, W, Hs, Ws, d_bits, d_bitmap, bitmap_words, d_image, d_binary,
                                                         level0);
    PCS_LAUNCH_CHECK(ctx, "resample_bits_kernel");
    return PCS_OK;
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
