// PNG files assembled on the device: the encoder behind output_data (ocr4all_pixel_classifier/lib/output.py:38-41,
// skimage.io.imsave of the colour / overlay / inverted masks), which is what dominates the wall time of the
// reference's prediction loop once the network is fast (zlib on one host core per image).
//
// A PNG is  signature | IHDR | IDAT(zlib(filtered scanlines)) | IEND.  Everything is data parallel except the entropy
// coder, so the device writes a valid file with the deflate stream in STORED blocks (RFC 1951 section 3.2.4):
//   scanline r      = filter byte 0 + W*C pixel bytes                       (png_body_kernel, one pass over the image)
//   stored block b  = 5-byte header + as many whole scanlines as fit 65535 bytes
//   Adler-32        = per-scanline partial sums, combined in closed form     (png_adler_rows_kernel / png_adler_kernel)
//   CRC-32 of IDAT  = 256-byte CRCs combined by a tree per 32-KB tile, every tile shifted by x^(8 * bytes after it)
//                     mod P and XOR-ed (crc32_combine is linear, so all of it runs in parallel)   (png_crc_kernel)
// Level 0 files are 1.002x the raw image.  Level 1 adds the part of deflate that IS data parallel: every scanline is
// filtered with Sub or Up, whichever costs fewer bits (a run of equal pixels, or a row that repeats the one above,
// becomes a run of zero bytes), every run of equal bytes is coded as one literal
// plus length/distance-1 matches with the FIXED Huffman code (RFC 1951 section 3.2.6), the bit cost of each scanline
// is counted (png_rle_count_kernel), prefix-summed (png_rle_scan_kernel) and the codes are then written at their
// final bit positions by all scanlines at once (png_rle_emit_kernel).  Class-colour masks shrink 30-100x; noise
// grows by at most 1/8.  Either way the masks are written with one memcpy per file instead of ~10 ms of zlib each,
// and any PNG reader decodes them to exactly the mask bytes (tests decode with OpenCV and zlib).
#include "common.cuh"

#include <algorithm>
#include <cstdint>
#include <cstring>

namespace pcs {
namespace {

constexpr uint32_t kCrcPoly = 0xedb88320u;
constexpr size_t kHeadPlain = 8 + 25 + 8;       // signature, IHDR chunk, IDAT length + type
constexpr size_t kMaxHead = 8 + 25 + (12 + 768) + 8;   // ... with a PLTE chunk of up to 256 colours in front of the IDAT

struct PngPlan {
    int H, W, C;
    uint32_t line;                              // bytes per scanline incl. the filter byte
    uint32_t lines_per_block, nblocks;
    uint64_t zlib_bytes, file_bytes, stride;    // stride: bytes between the files of a batch
    uint32_t head_len;                          // bytes in front of the zlib stream (kHeadPlain, or more with a palette)
    uint32_t filter_a;                          // level 1 costs every scanline with this filter (1 Sub, 0 None) and with Up
    uint8_t head[kMaxHead];                     // signature + IHDR (+ PLTE) + IDAT length/type, built on the host
};

__host__ __device__ inline uint32_t crc_step(uint32_t c) {
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
    return c;
}

// (a * b) mod P and x^(n * 2^k) mod P on the reflected representation (zlib's crc32_combine arithmetic)
__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}
__device__ uint32_t x8n_modp(uint64_t n, const uint32_t* x2n /*[32]: x^(2^k)*/) {       // x^(8 n) mod P
    uint32_t p = 1u << 31;
    int k = 3;
    while (n) {
        if (n & 1) p = multmodp(x2n[k & 31], p);
        n >>= 1;
        ++k;
    }
    return p;
}

// Everything in front of the Adler-32: file head, zlib header, stored-block headers and the scanlines.  One thread
// assembles one aligned 32-bit word of the file (four consecutive bytes, decoded once and then advanced byte by byte)
// and stores it whole: the scanlines sit at arbitrary byte offsets, so byte stores would quarter the store throughput.
// grid = (ceil(words / 256), n).  The word that straddles the Adler-32 is completed by png_adler_kernel afterwards.
__global__ void __launch_bounds__(256) png_body_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, const PngPlan pl) {
    const uint32_t kHead = pl.head_len;
    const uint32_t limit = (uint32_t)(kHead + pl.zlib_bytes - 4);                       // first byte of the Adler-32
    const uint32_t o0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
    if (o0 >= limit) return;
    const uint32_t npx = pl.line - 1, pitch = pl.lines_per_block * pl.line + 5;         // bytes from one block header to the next
    const uint8_t* src = img + (uint64_t)blockIdx.y * pl.H * npx;
    // position of byte o0 inside the block structure (only meaningful from the first block header on)
    uint32_t blk = 0, t = 0, row_in = 0, col = 0;
    if (o0 >= kHead + 2) {
        const uint32_t z = o0 - (uint32_t)(kHead + 2);
        blk = z / pitch; t = z - blk * pitch;
        if (t >= 5) { row_in = (t - 5) / pl.line; col = (t - 5) - row_in * pl.line; }
    }
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t o = o0 + k;
        uint32_t b = 0;
        if (o < kHead) b = pl.head[o];
        else if (o < kHead + 2) b = o == kHead ? 0x78u : 0x01u;                          // zlib: deflate, 32K window, no dictionary
        else if (o < limit) {
            if (t < 5) {                                                                 // stored-block header: final flag, LEN, ~LEN
                const uint32_t rows = min(pl.lines_per_block, (uint32_t)pl.H - blk * pl.lines_per_block), len = rows * pl.line;
                b = t == 0 ? (blk + 1 == pl.nblocks ? 1u : 0u) : (t == 1 ? (len & 0xffu) : (t == 2 ? (len >> 8) : (t == 3 ? (~len & 0xffu) : ((~len >> 8) & 0xffu))));
            } else if (col != 0) {
                b = __ldg(src + (uint64_t)(blk * pl.lines_per_block + row_in) * npx + (col - 1));
            }                                                                            // col == 0: filter type 0
            if (++t > 5) { if (++col == pl.line) { col = 0; ++row_in; } }
            if (t == 5) { row_in = 0; col = 0; }
            if (t == pitch) { ++blk; t = 0; }
        }
        word |= b << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(out + (uint64_t)blockIdx.y * pl.stride + o0) = word;
}

// per scanline: A = sum of bytes, B = sum of (line - j) * byte_j  (the filter byte is zero and contributes nothing)
__global__ void __launch_bounds__(256) png_adler_rows_kernel(const uint8_t* __restrict__ img, const PngPlan pl,
                                                             unsigned long long* __restrict__ ab /*[n][H][2]*/) {
    const int row = blockIdx.x, page = blockIdx.y;
    const uint32_t npx = pl.line - 1;
    const uint8_t* src = img + ((uint64_t)page * pl.H + row) * npx;
    unsigned long long a = 0, b = 0;
    for (uint32_t j = threadIdx.x; j < npx; j += 256) {
        const unsigned v = __ldg(src + j);
        a += v;
        b += (unsigned long long)(pl.line - (j + 1)) * v;                                // stream position of pixel byte j is j + 1
    }
    __shared__ unsigned long long sa[8], sb[8];
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { a += sa[k]; b += sb[k]; }
        ab[((uint64_t)page * pl.H + row) * 2] = a;
        ab[((uint64_t)page * pl.H + row) * 2 + 1] = b;
    }
}

// s1 = 1 + sum A_r, s2 = N + sum (B_r + A_r * bytes after row r), both mod 65521; one block per image
__global__ void __launch_bounds__(256) png_adler_kernel(const unsigned long long* __restrict__ ab, uint8_t* __restrict__ out, const PngPlan pl,
                                                        const unsigned long long* __restrict__ zbytes /*[n] or null: level 1 sizes*/) {
    const int page = blockIdx.x;
    const unsigned long long* p = ab + (uint64_t)page * pl.H * 2;
    unsigned long long s1 = 0, s2 = 0;
    for (int r = threadIdx.x; r < pl.H; r += 256) {
        const unsigned long long a = p[2 * r] % 65521ull, b = p[2 * r + 1] % 65521ull;
        const unsigned long long after = ((unsigned long long)(pl.H - 1 - r) * pl.line) % 65521ull;
        s1 += a;
        s2 += b + a * after;
    }
    __shared__ unsigned long long t1[8], t2[8];
    for (int o = 16; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { t1[threadIdx.x >> 5] = s1; t2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { s1 += t1[k]; s2 += t2[k]; }
        const unsigned long long n = ((unsigned long long)pl.H * pl.line) % 65521ull;
        const uint32_t a = (uint32_t)((1 + s1) % 65521ull), b = (uint32_t)((n + s2) % 65521ull);
        uint8_t* f = out + (uint64_t)page * pl.stride + pl.head_len + (zbytes ? zbytes[page] : pl.zlib_bytes) - 4;
        f[0] = (uint8_t)(b >> 8); f[1] = (uint8_t)b; f[2] = (uint8_t)(a >> 8); f[3] = (uint8_t)a;
    }
}

// CRC-32 of the IDAT chunk (type + data).  A block takes 32 KB of it: coalesced copy into shared memory (chunk rows
// padded by one word against bank conflicts), one 256-byte CRC per thread, a 7-level combine tree inside the block
// with the precomputed shifts x^(8 * 256 * 2^j), and one general shift by the bytes that follow the tile.
constexpr int kCrcThreads = 128, kCrcBytes = 256, kCrcTile = kCrcThreads * kCrcBytes, kCrcPitch = kCrcBytes + 4;

__global__ void __launch_bounds__(kCrcThreads) png_crc_kernel(const uint8_t* __restrict__ out, const PngPlan pl,
                                                              const unsigned long long* __restrict__ zbytes /*[n] or null*/,
                                                              uint32_t* __restrict__ crc /*[n], zeroed*/) {
    __shared__ uint32_t tab[256], x2n[32], lvl[7], part[kCrcThreads];
    __shared__ __align__(4) uint8_t tile[kCrcThreads * kCrcPitch];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += kCrcThreads) tab[i] = crc_step((uint32_t)i);
    if (tid == 0) {
        uint32_t p = 1u << 30;                                                           // x^1
        x2n[0] = p;
        for (int k = 1; k < 32; ++k) x2n[k] = p = multmodp(p, p);
    }
    const uint64_t total = 4 + (zbytes ? zbytes[blockIdx.y] : pl.zlib_bytes);            // "IDAT" + data
    const uint64_t tile_begin = (uint64_t)blockIdx.x * kCrcTile;
    if (tile_begin >= total) return;                                                     // level 1: the grid covers the worst case
    const int nbytes = (int)min((uint64_t)kCrcTile, total - tile_begin);
    const uint8_t* src = out + (uint64_t)blockIdx.y * pl.stride + (pl.head_len - 4) + tile_begin;
    for (int i = tid; i < nbytes; i += kCrcThreads) tile[(i >> 8) * kCrcPitch + (i & 255)] = src[i];
    __syncthreads();
    if (tid < 7) lvl[tid] = x8n_modp((uint64_t)kCrcBytes << tid, x2n);
    const int len = max(0, min(kCrcBytes, nbytes - tid * kCrcBytes));
    uint32_t c = 0xffffffffu;
    const uint8_t* mine = tile + tid * kCrcPitch;
    for (int i = 0; i < len; ++i) c = tab[(c ^ mine[i]) & 0xffu] ^ (c >> 8);
    part[tid] = c ^ 0xffffffffu;                                                         // the CRC of an empty chunk is 0
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < 7; ++j) {
        const int stride = 1 << j;
        if ((tid & (2 * stride - 1)) == 0) {
            const int right = tid + stride;
            const int len_r = max(0, min(stride * kCrcBytes, nbytes - right * kCrcBytes));        // bytes under the right subtree
            const uint32_t shift = len_r == stride * kCrcBytes ? lvl[j] : x8n_modp((uint64_t)len_r, x2n);
            part[tid] = multmodp(shift, part[tid]) ^ part[right];
        }
        __syncthreads();
    }
    if (tid == 0) atomicXor(&crc[blockIdx.y], multmodp(x8n_modp(total - (tile_begin + (uint64_t)nbytes), x2n), part[0]));
}

// ---------------------------------------------------------------------------------------------------------------
// level 1: Sub filter + fixed-Huffman deflate whose only matches are runs (distance 1)
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kMaxRleLine = 16384;          // bytes per scanline the run tables hold in shared memory (3 B each)

__device__ __forceinline__ uint32_t lit_bits(uint32_t v) { return v < 144u ? 8u : 9u; }
__device__ __forceinline__ uint32_t match_bits(uint32_t len) {               // length symbol + extra bits + 5-bit distance code
    const uint32_t l = len - 3;
    const uint32_t extra = (len == 258u || l < 8u) ? 0u : (31u - __clz(l)) - 2u;
    return (len >= 115u ? 8u : 7u) + extra + 5u;
}
// a run of n bytes of value v: one literal, then as many 258-byte matches as fit, then one shorter match or 1-2 literals
__device__ __forceinline__ uint32_t run_bits(uint32_t v, uint32_t n) {
    const uint32_t m = n - 1, k = m / 258u, rem = m - 258u * k;
    return lit_bits(v) + 13u * k + (rem >= 3u ? match_bits(rem) : rem * lit_bits(v));
}
// `w` is the scanline's bit buffer in SHARED memory (zeroed, OR-ed by the lanes of the warp), `pos` a bit offset in it
__device__ __forceinline__ void put_bits(uint32_t* w, uint32_t& pos, uint32_t val, uint32_t nbits) {
    const uint32_t word = pos >> 5, sh = pos & 31u;
    atomicOr(&w[word], val << sh);
    if (sh + nbits > 32u) atomicOr(&w[word + 1], val >> (32u - sh));
    pos += nbits;
}
__device__ __forceinline__ void put_literal(uint32_t* w, uint32_t& pos, uint32_t v) {
    const uint32_t nb = lit_bits(v), code = v < 144u ? 0x30u + v : 0x190u + (v - 144u);
    put_bits(w, pos, __brev(code) >> (32u - nb), nb);                         // Huffman codes go in most significant bit first
}
__device__ __forceinline__ void put_match(uint32_t* w, uint32_t& pos, uint32_t len) {
    const uint32_t l = len - 3;
    uint32_t sym, e = 0;
    if (len == 258u) sym = 285u;
    else if (l < 8u) sym = 257u + l;
    else { e = (31u - __clz(l)) - 2u; sym = 261u + 4u * e + ((l >> e) - 4u); }
    const uint32_t nb = sym < 280u ? 7u : 8u, code = sym < 280u ? sym - 256u : 0xc0u + (sym - 280u);
    // symbol, extra bits (least significant bit first), distance symbol 0 = five zero bits
    put_bits(w, pos, (__brev(code) >> (32u - nb)) | ((l & ((1u << e) - 1u)) << nb), nb + e + 5u);
}

// A scanline staged in shared memory with aligned 32-bit loads (a warp reading it byte by byte issues one 32-byte
// request per instruction); byte j of the row is s[mis + j], `mis` (returned) = the row's misalignment.  The first / last
// word of the whole image buffer is assembled from byte loads instead of reaching across the buffer's ends.
__device__ __forceinline__ uint32_t stage_row(const uint8_t* __restrict__ src, uint32_t nbytes, bool guard_head, bool guard_tail,
                                              uint8_t* s) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint32_t mis = (uint32_t)(a & 3);
    const uint32_t* p = reinterpret_cast<const uint32_t*>(a - mis);
    const uint32_t words = (mis + nbytes + 3) >> 2;
    const uint32_t w0 = (guard_head && mis) ? 1u : 0u, w1 = (guard_tail && ((mis + nbytes) & 3u)) ? words - 1 : words;
    uint32_t* sw = reinterpret_cast<uint32_t*>(s);
    for (uint32_t w = w0 + threadIdx.x; w < w1; w += 32) sw[w] = __ldg(p + w);
    if (threadIdx.x == 0) {
        if (w0) for (uint32_t b = mis; b < min(4u, mis + nbytes); ++b) s[b] = __ldg(src + (b - mis));
        if (w1 < words) for (uint32_t b = max(w1 * 4u, mis); b < mis + nbytes; ++b) s[b] = __ldg(src + (b - mis));
    }
    return mis;
}

// Filtered scanline and its run starts in shared memory; returns the number of runs (warp-uniform).  `raw` / `above`
// are the staged rows (shared memory; `above` null on the first row).
// fb[0] is the filter type: 0 (None) fb[i] = raw[i-1];  1 (Sub) fb[i] = raw[i-1] - raw[i-1-C];  2 (Up) fb[i] = raw[i-1] - above[i-1]
// (zeros above row 0).
// One pass: filter, Adler partial sums (ADLER: lane-local, of the FILTERED bytes incl. the filter byte) and run starts;
// the previous byte comes from the neighbouring lane, not from shared memory.
template <bool ADLER>
__device__ uint32_t rle_prepare(const uint8_t* raw, const uint8_t* above, uint32_t ftype, uint32_t line,
                                uint32_t C, uint8_t* fb, uint16_t* starts, unsigned long long& a, unsigned long long& b) {
    const uint32_t lane = threadIdx.x;
    __syncwarp();
    uint32_t count = 0, carry = 0;
    for (uint32_t base = 0; base < line; base += 32) {
        const uint32_t i = base + lane;
        uint32_t v = ftype;
        if (i && i < line) {
            const uint32_t j = i - 1;
            const uint32_t pred = ftype == 1u ? (j >= C ? (uint32_t)raw[j - C] : 0u) : (ftype == 2u && above ? (uint32_t)above[j] : 0u);
            v = ((uint32_t)raw[j] - pred) & 0xffu;
        }
        if (i < line) {
            fb[i] = (uint8_t)v;
            if (ADLER) { a += v; b += (unsigned long long)(line - i) * v; }
        }
        uint32_t prev = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) prev = carry;
        const bool st = i < line && (i == 0 || v != prev);
        const uint32_t m = __ballot_sync(0xffffffffu, st);
        if (st) starts[count + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        count += __popc(m);
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    return count;
}

// one warp per scanline: the scanline is costed with the Sub and with the Up filter (a class-colour mask repeats the row
// above, so Up turns most rows into one long zero run; Sub wins on rows where regions begin), the cheaper one is kept:
// its type, its bit cost and its Adler partial sums (of the FILTERED bytes, filter byte included)
__global__ void __launch_bounds__(32) png_rle_count_kernel(const uint8_t* __restrict__ img, const PngPlan pl,
                                                           unsigned long long* __restrict__ ab, uint32_t* __restrict__ bits,
                                                           uint8_t* __restrict__ ftypes) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* fb = sm;
    uint16_t* starts = reinterpret_cast<uint16_t*>(sm + ((pl.line + 15u) & ~15u));
    const uint32_t row = blockIdx.x, page = blockIdx.y, lane = threadIdx.x, npx = pl.line - 1;
    const uint8_t* graw = img + ((uint64_t)page * pl.H + row) * npx;
    uint8_t* sraw = sm + ((pl.line + 15u) & ~15u) + ((2u * pl.line + 15u) & ~15u);
    uint8_t* sabove = sraw + ((pl.line + 8u + 15u) & ~15u);
    const bool first = page == 0 && row == 0, last = page + 1 == gridDim.y && row + 1 == (uint32_t)pl.H;
    const uint8_t* raw = sraw + stage_row(graw, npx, first, last, sraw);
    const uint8_t* above = row ? sabove + stage_row(graw - npx, npx, page == 0 && row == 1, false, sabove) : nullptr;
    unsigned long long best_a = 0, best_b = 0;
    uint32_t best_t = 0xffffffffu, best_f = pl.filter_a;
    for (uint32_t trial = 0; trial < 2; ++trial) {
        const uint32_t ftype = trial ? 2u : pl.filter_a;
        unsigned long long a = 0, b = 0;
        const uint32_t R = rle_prepare<true>(raw, above, ftype, pl.line, (uint32_t)pl.C, fb, starts, a, b);
        uint32_t t = 0;
        for (uint32_t r = lane; r < R; r += 32) {
            const uint32_t s0 = starts[r], s1 = r + 1 < R ? starts[r + 1] : pl.line;
            t += run_bits(fb[s0], s1 - s0);
        }
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); t += __shfl_xor_sync(0xffffffffu, t, o);
        }
        if (t < best_t) { best_t = t; best_a = a; best_b = b; best_f = ftype; }
    }
    if (lane == 0) {
        ab[((uint64_t)page * pl.H + row) * 2] = best_a;
        ab[((uint64_t)page * pl.H + row) * 2 + 1] = best_b;
        bits[(uint64_t)page * pl.H + row] = best_t;
        ftypes[(uint64_t)page * pl.H + row] = (uint8_t)best_f;
    }
}

// one block per image: first bit of every scanline (absolute bit position in the file), the size of the zlib stream,
// and the fixed bytes in front of the codes (file head with the IDAT length, zlib header, deflate block header)
__global__ void __launch_bounds__(1024) png_rle_scan_kernel(const uint32_t* __restrict__ bits, const PngPlan pl, uint8_t* __restrict__ out,
                                                            unsigned long long* __restrict__ bitbase, unsigned long long* __restrict__ zbytes) {
    __shared__ unsigned long long s_warp[32], s_carry;
    const int page = blockIdx.x, tid = threadIdx.x;
    const uint32_t kHead = pl.head_len;
    if (tid == 0) s_carry = (kHead + 2) * 8ull + 3ull;                                   // after BFINAL = 1, BTYPE = 01
    __syncthreads();
    for (int base = 0; base < pl.H; base += 1024) {
        const int i = base + tid;
        const unsigned long long v = i < pl.H ? bits[(uint64_t)page * pl.H + i] : 0ull;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            unsigned long long w = s_warp[tid];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (tid >= o) w += t;
            }
            s_warp[tid] = w;
        }
        __syncthreads();
        const unsigned long long off = s_carry + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0ull) + incl - v;
        if (i < pl.H) bitbase[(uint64_t)page * pl.H + i] = off;
        __syncthreads();
        if (tid == 1023) s_carry = off + v;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned long long end_bit = s_carry + 7ull;                               // end-of-block symbol: seven zero bits
        const unsigned long long zb = (end_bit + 7ull) / 8ull - kHead + 4ull;            // zlib header + deflate + Adler-32
        zbytes[page] = zb;
        uint8_t* f = out + (uint64_t)page * pl.stride;
        f[kHead - 8] = (uint8_t)(zb >> 24); f[kHead - 7] = (uint8_t)(zb >> 16); f[kHead - 6] = (uint8_t)(zb >> 8); f[kHead - 5] = (uint8_t)zb;
        f[kHead] = 0x78; f[kHead + 1] = 0x01; f[kHead + 2] = 0x03;                       // BFINAL = 1, BTYPE = 01 (fixed Huffman)
    }
    {   // the fixed bytes in front (signature, IHDR, palette, "IDAT"); the IDAT length was written above
        uint8_t* f = out + (uint64_t)page * pl.stride;
        for (uint32_t i = tid; i < kHead; i += 1024)
            if (i < kHead - 8 || i >= kHead - 4) f[i] = pl.head[i];
    }
}

// one warp per scanline: the codes of its runs are assembled in a shared-memory bit buffer (the lanes own consecutive
// runs, i.e. adjacent bit ranges: shared-memory atomics) that starts at the scanline's bit offset inside its first file
// word; whole words are then stored, and only the first and the last word - shared with the neighbouring scanlines -
// are OR-ed into the zero-filled file
__global__ void __launch_bounds__(32) png_rle_emit_kernel(const uint8_t* __restrict__ img, const PngPlan pl,
                                                          const unsigned long long* __restrict__ bitbase,
                                                          const uint8_t* __restrict__ ftypes, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* fb = sm;
    uint16_t* starts = reinterpret_cast<uint16_t*>(sm + ((pl.line + 15u) & ~15u));
    const uint32_t row = blockIdx.x, page = blockIdx.y, lane = threadIdx.x, npx = pl.line - 1;
    const uint8_t* graw = img + ((uint64_t)page * pl.H + row) * npx;
    uint32_t* bitbuf = reinterpret_cast<uint32_t*>(sm + ((pl.line + 15u) & ~15u) + ((2u * pl.line + 15u) & ~15u));
    const uint32_t buf_words = (9u * pl.line + 31u + 31u) / 32u + 1u;                      // worst case 9 bits per byte + the offset
    uint8_t* sraw = reinterpret_cast<uint8_t*>(bitbuf) + ((buf_words * 4u + 15u) & ~15u);
    uint8_t* sabove = sraw + ((pl.line + 8u + 15u) & ~15u);
    const uint32_t ftype = ftypes[(uint64_t)page * pl.H + row];
    const bool first = page == 0 && row == 0, last = page + 1 == gridDim.y && row + 1 == (uint32_t)pl.H;
    const uint8_t* raw = sraw + stage_row(graw, npx, first, last, sraw);
    const uint8_t* above = (row && ftype == 2u) ? sabove + stage_row(graw - npx, npx, page == 0 && row == 1, false, sabove) : nullptr;
    unsigned long long unused_a = 0, unused_b = 0;
    const uint32_t R = rle_prepare<false>(raw, above, ftype, pl.line, (uint32_t)pl.C, fb, starts, unused_a, unused_b);
    for (uint32_t i = lane; i < buf_words; i += 32) bitbuf[i] = 0;
    __syncwarp();
    const unsigned long long base = bitbase[(uint64_t)page * pl.H + row];
    uint32_t carry = (uint32_t)(base & 31ull);                                               // bit offset inside the first file word
    uint32_t* const w = bitbuf;
    for (uint32_t r0 = 0; r0 < R; r0 += 32) {
        const uint32_t r = r0 + lane;
        uint32_t v = 0, n = 0, cost = 0;
        if (r < R) {
            const uint32_t s0 = starts[r], s1 = r + 1 < R ? starts[r + 1] : pl.line;
            v = fb[s0]; n = s1 - s0; cost = run_bits(v, n);
        }
        uint32_t incl = cost;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (r < R) {
            uint32_t pos = carry + incl - cost;
            put_literal(w, pos, v);
            uint32_t m = n - 1;
            for (; m >= 258u; m -= 258u) put_match(w, pos, 258u);
            if (m >= 3u) put_match(w, pos, m);
            else for (; m; --m) put_literal(w, pos, v);
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // carry = offset + bits of the scanline: words [0, nwords) of the buffer go to file words first_word + ...
    uint32_t* fw = reinterpret_cast<uint32_t*>(out + (uint64_t)page * pl.stride) + (base >> 5);
    const uint32_t nwords = (carry + 31u) >> 5;
    for (uint32_t i = lane; i < nwords; i += 32) {
        const uint32_t v = bitbuf[i];
        if (i == 0 || i + 1 == nwords) { if (v) atomicOr(&fw[i], v); }
        else fw[i] = v;
    }
}

__global__ void png_finish_kernel(uint8_t* __restrict__ out, const PngPlan pl, const uint32_t* __restrict__ crc, int n,
                                  const unsigned long long* __restrict__ zbytes, unsigned long long* __restrict__ sizes) {
    const int page = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n) return;
    const uint64_t zb = zbytes ? zbytes[page] : pl.zlib_bytes;
    uint8_t* f = out + (uint64_t)page * pl.stride + pl.head_len + zb;
    const uint32_t c = crc[page];
    const uint8_t tail[16] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c,
                              0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
    for (int i = 0; i < 16; ++i) f[i] = tail[i];
    if (sizes) sizes[page] = pl.head_len + zb + 16;
}

uint32_t host_crc(const uint8_t* p, size_t n) {
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; ++i) c = crc_step((c ^ p[i]) & 0xffu) ^ (c >> 8);
    return c ^ 0xffffffffu;
}
void be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

// depth == 0: 8-bit grey / RGB / RGBA with C channels.  depth in {1, 2, 4, 8}: colour type 3 (indexed), `palette` holds
// ncolors RGB triples, the image is H rows of ceil(W * depth / 8) bytes (leftmost pixel in the high-order bits).
bool make_plan(int H, int W, int C, PngPlan& pl, int depth = 0, const uint8_t* palette = nullptr, int ncolors = 0) {
    if (H <= 0 || W <= 0) return false;
    if (depth == 0 && C != 1 && C != 3 && C != 4) return false;
    if (depth != 0 && (C != 1 || (depth != 1 && depth != 2 && depth != 4 && depth != 8) || ncolors < 1 || ncolors > (1 << depth) || !palette)) return false;
    const uint64_t line = 1 + (depth ? ((uint64_t)W * depth + 7) / 8 : (uint64_t)W * C);
    if (line > 65535) return false;                                                      // one scanline must fit a stored block
    pl.H = H; pl.W = W; pl.C = C;
    pl.filter_a = depth ? 0u : 1u;                                                       // packed indices: None, samples: Sub
    pl.line = (uint32_t)line;
    pl.lines_per_block = (uint32_t)(65535 / line);
    pl.nblocks = ((uint32_t)H + pl.lines_per_block - 1) / pl.lines_per_block;
    pl.zlib_bytes = 2 + (uint64_t)pl.nblocks * 5 + (uint64_t)H * line + 4;
    if (pl.zlib_bytes > 0x7fffffffull) return false;
    pl.head_len = (uint32_t)(kHeadPlain + (depth ? 12 + 3 * (size_t)ncolors : 0));
    pl.file_bytes = pl.head_len + pl.zlib_bytes + 4 + 12;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t* h = pl.head;
    memcpy(h, sig, 8);
    be32(h + 8, 13); memcpy(h + 12, "IHDR", 4);
    be32(h + 16, (uint32_t)W); be32(h + 20, (uint32_t)H);
    h[24] = (uint8_t)(depth ? depth : 8);
    h[25] = (uint8_t)(depth ? 3 : (C == 1 ? 0 : (C == 3 ? 2 : 6))); h[26] = 0; h[27] = 0; h[28] = 0;   // indexed / grey / RGB / RGBA
    be32(h + 29, host_crc(h + 12, 17));
    uint8_t* q = h + 33;
    if (depth) {
        be32(q, 3u * (uint32_t)ncolors); memcpy(q + 4, "PLTE", 4);
        memcpy(q + 8, palette, 3 * (size_t)ncolors);
        be32(q + 8 + 3 * ncolors, host_crc(q + 4, 4 + 3 * (size_t)ncolors));
        q += 12 + 3 * ncolors;
    }
    be32(q, (uint32_t)pl.zlib_bytes); memcpy(q + 4, "IDAT", 4);
    return true;
}

}  // namespace

static size_t plan_file_bytes(const PngPlan& pl, int level) {
    if (level <= 0) return (size_t)pl.file_bytes;
    if (pl.line > kMaxRleLine) return 0;
    // worst case of level 1: every byte a 9-bit literal
    const uint64_t zlib = 2 + (3 + 7 + 9 * (uint64_t)pl.H * pl.line + 7) / 8 + 4;
    return zlib > 0x7fffffffull ? 0 : (size_t)(pl.head_len + zlib + 16);
}

size_t png_file_bytes(int H, int W, int C, int level) {
    PngPlan pl{};
    if (!make_plan(H, W, C, pl)) return 0;
    return plan_file_bytes(pl, level);
}

int png_index_depth(int ncolors) { return ncolors <= 2 ? 1 : (ncolors <= 4 ? 2 : (ncolors <= 16 ? 4 : 8)); }

size_t png_indexed_file_bytes(int H, int W, int ncolors, int level) {
    PngPlan pl{};
    uint8_t pal[768] = {};
    if (ncolors < 1 || ncolors > 256 || !make_plan(H, W, 1, pl, png_index_depth(ncolors), pal, ncolors)) return 0;
    return plan_file_bytes(pl, level);
}

static int encode_with_plan(pcs_ctx* ctx, PngPlan& pl, const uint8_t* d_img, int n, int level, uint8_t* d_out, size_t stride,
                            unsigned long long* d_sizes) {
    const int H = pl.H;
    const size_t bound = plan_file_bytes(pl, level);
    if (!bound) return set_err(ctx, PCS_ERR_ARG, "png_encode: scanlines of %u bytes are too long for level %d", pl.line, level);
    if ((stride & 3) || (reinterpret_cast<uintptr_t>(d_out) & 3)) return set_err(ctx, PCS_ERR_ARG, "png_encode: output and stride must be 4-byte aligned");
    if (stride < ((bound + 3) & ~(size_t)3)) return set_err(ctx, PCS_ERR_ARG, "png_encode: %zu bytes per file needed, stride is %zu", bound, stride);
    if (H > 65535) return set_err(ctx, PCS_ERR_ARG, "png_encode: more than 65535 rows");
    const size_t kHead = pl.head_len;
    pl.stride = stride;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t rows = (size_t)n * H;
    PCS_TRY(scratch_reserve(ctx, al(rows * 16) + al(rows * 4) + al(rows * 8) + al(rows) + al((size_t)n * 8) + al((size_t)n * 4) + 256));
    char* q = reinterpret_cast<char*>(ctx->scratch);
    unsigned long long* ab = reinterpret_cast<unsigned long long*>(q); q += al(rows * 16);
    uint32_t* bits = reinterpret_cast<uint32_t*>(q); q += al(rows * 4);
    unsigned long long* bitbase = reinterpret_cast<unsigned long long*>(q); q += al(rows * 8);
    uint8_t* ftypes = reinterpret_cast<uint8_t*>(q); q += al(rows);
    unsigned long long* zbytes = reinterpret_cast<unsigned long long*>(q); q += al((size_t)n * 8);
    uint32_t* crc = reinterpret_cast<uint32_t*>(q);
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(crc, 0, (size_t)n * 4, st));
    uint64_t zlib_max = pl.zlib_bytes;
    if (level <= 0) {
        const uint64_t words = (kHead + pl.zlib_bytes - 4 + 3) / 4;
        png_body_kernel<<<dim3((unsigned)((words + 255) / 256), n), 256, 0, st>>>(d_img, d_out, pl);
        PCS_LAUNCH_CHECK(ctx, "png_body_kernel");
        png_adler_rows_kernel<<<dim3(H, n), 256, 0, st>>>(d_img, pl, ab);
        PCS_LAUNCH_CHECK(ctx, "png_adler_rows_kernel");
        zbytes = nullptr;
    } else {
        zlib_max = bound - kHead - 16;
        const size_t staged = 2 * (((size_t)pl.line + 8 + 15) & ~(size_t)15);                // the scanline and the one above it
        const size_t smem_count = ((pl.line + 15u) & ~15u) + ((2 * (size_t)pl.line + 15u) & ~(size_t)15u) + staged;
        const size_t smem_emit = ((pl.line + 15u) & ~15u) + ((2 * (size_t)pl.line + 15u) & ~(size_t)15u) +
                                 ((((9 * (size_t)pl.line + 62) / 32 + 1) * 4 + 15) & ~(size_t)15) + staged;
        if (smem_emit > 48 * 1024) {
            static bool set[64] = {};
            if (ctx->device >= 64 || !set[ctx->device]) {
                PCS_CUDA(ctx, cudaFuncSetAttribute(png_rle_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
                PCS_CUDA(ctx, cudaFuncSetAttribute(png_rle_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
                if (ctx->device < 64) set[ctx->device] = true;
            }
        }
        PCS_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)n * stride, st));               // the codes are OR-ed into place
        png_rle_count_kernel<<<dim3(H, n), 32, smem_count, st>>>(d_img, pl, ab, bits, ftypes);
        PCS_LAUNCH_CHECK(ctx, "png_rle_count_kernel");
        png_rle_scan_kernel<<<n, 1024, 0, st>>>(bits, pl, d_out, bitbase, zbytes);
        PCS_LAUNCH_CHECK(ctx, "png_rle_scan_kernel");
        png_rle_emit_kernel<<<dim3(H, n), 32, smem_emit, st>>>(d_img, pl, bitbase, ftypes, d_out);
        PCS_LAUNCH_CHECK(ctx, "png_rle_emit_kernel");
    }
    png_adler_kernel<<<n, 256, 0, st>>>(ab, d_out, pl, zbytes);
    PCS_LAUNCH_CHECK(ctx, "png_adler_kernel");
    const uint64_t tiles = (4 + zlib_max + kCrcTile - 1) / kCrcTile;
    png_crc_kernel<<<dim3((unsigned)tiles, n), kCrcThreads, 0, st>>>(d_out, pl, zbytes, crc);
    PCS_LAUNCH_CHECK(ctx, "png_crc_kernel");
    png_finish_kernel<<<(n + 63) / 64, 64, 0, st>>>(d_out, pl, crc, n, zbytes, d_sizes);
    PCS_LAUNCH_CHECK(ctx, "png_finish_kernel");
    return PCS_OK;
}

int launch_png_encode(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int C, int level, uint8_t* d_out, size_t stride,
                      unsigned long long* d_sizes) {
    PngPlan pl{};
    if (n <= 0 || !make_plan(H, W, C, pl)) return set_err(ctx, PCS_ERR_ARG, "png_encode: unsupported shape %d x %d x %d", H, W, C);
    return encode_with_plan(ctx, pl, d_img, n, level, d_out, stride, d_sizes);
}

// Indexed-colour files (PNG colour type 3) of images whose pixels are palette indices packed `depth` bits each
// (png_index_depth(ncolors); leftmost pixel in the high-order bits, rows of ceil(W * depth / 8) bytes).  What the colour
// masks of a class map are: n_classes + 1 distinct colours, i.e. 2 bits per pixel for the default colour map instead of
// 24 -- a twelfth of the bytes in every pass of the encoder, in the file, across PCIe and on disk; any PNG reader
// expands the palette to the same RGB pixels.
int launch_png_encode_indexed(pcs_ctx* ctx, const uint8_t* d_idx, int n, int H, int W, const uint8_t* h_palette, int ncolors, int level,
                              uint8_t* d_out, size_t stride, unsigned long long* d_sizes) {
    PngPlan pl{};
    if (n <= 0 || ncolors < 1 || ncolors > 256 || !make_plan(H, W, 1, pl, png_index_depth(ncolors), h_palette, ncolors))
        return set_err(ctx, PCS_ERR_ARG, "png_encode_indexed: unsupported shape %d x %d with %d colours", H, W, ncolors);
    return encode_with_plan(ctx, pl, d_idx, n, level, d_out, stride, d_sizes);
}

// Palette indices of the three masks of generate_output_masks (lib/output.py:44-60) straight from the class map and
// data.binary, packed for launch_png_encode_indexed: palette = the n_lut LUT colours + black at index n_lut;
//   color    : label (labels outside the LUT -> black, like ColorMap.to_rgb_array)
//   overlay  : black where (uint8)(1 - binary) == 0      inverted : black where binary == 0
// out: [3][n][H][Wb] with Wb = ceil(W * depth / 8).  One thread per output byte of all three kinds.
__global__ void __launch_bounds__(256) mask_index_kernel(const uint8_t* __restrict__ labels, const uint8_t* __restrict__ binary, int n, int H,
                                                         int W, int n_lut, int depth, int Wb, uint8_t* __restrict__ out) {
    const size_t total = (size_t)n * H * Wb, plane = total;
    const int ppb = 8 / depth;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (size_t)gridDim.x * 256) {
        const size_t rowi = t / Wb;
        const int xb = (int)(t - rowi * Wb);
        const uint8_t* lrow = labels + rowi * W;
        const uint8_t* brow = binary + rowi * W;
        uint32_t c = 0, o = 0, v = 0;
        for (int k = 0; k < ppb; ++k) {
            const int x = xb * ppb + k;
            uint32_t ic = 0, io = 0, iv = 0;
            if (x < W) {
                const uint32_t lab = __ldg(lrow + x), bin = __ldg(brow + x);
                ic = lab < (uint32_t)n_lut ? lab : (uint32_t)n_lut;
                io = ((uint8_t)(1u - bin) != 0) ? ic : (uint32_t)n_lut;
                iv = bin != 0 ? ic : (uint32_t)n_lut;
            }
            const int sh = 8 - depth * (k + 1);
            c |= ic << sh; o |= io << sh; v |= iv << sh;
        }
        out[t] = (uint8_t)c; out[plane + t] = (uint8_t)o; out[2 * plane + t] = (uint8_t)v;
    }
}

int launch_mask_indices(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W, int n_lut, uint8_t* d_out) {
    if (n <= 0 || H <= 0 || W <= 0 || n_lut < 0 || n_lut > 255) return set_err(ctx, PCS_ERR_ARG, "mask_indices: bad argument");
    const int depth = png_index_depth(n_lut + 1), Wb = (W * depth + 7) / 8;
    const size_t total = (size_t)n * H * Wb;
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 16, (total + 255) / 256);
    mask_index_kernel<<<blocks, 256, 0, ctx->stream>>>(d_labels, d_binary, n, H, W, n_lut, depth, Wb, d_out);
    PCS_LAUNCH_CHECK(ctx, "mask_index_kernel");
    return PCS_OK;
}

}  // namespace pcs
