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
// The file is 1.002x the raw image (no compression): the masks are then written with one memcpy per file instead of
// ~10 ms of zlib each.  Any PNG reader decodes them to exactly the mask bytes (tests decode with OpenCV and zlib).
#include "common.cuh"

#include <cstdint>
#include <cstring>

namespace pcs {
namespace {

constexpr uint32_t kCrcPoly = 0xedb88320u;
constexpr size_t kHead = 8 + 25 + 8;            // signature, IHDR chunk, IDAT length + type

struct PngPlan {
    int H, W, C;
    uint32_t line;                              // bytes per scanline incl. the filter byte
    uint32_t lines_per_block, nblocks;
    uint64_t zlib_bytes, file_bytes, stride;    // stride: bytes between the files of a batch
    uint8_t head[kHead];                        // signature + IHDR + IDAT length/type, built on the host
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
__global__ void __launch_bounds__(256) png_adler_kernel(const unsigned long long* __restrict__ ab, uint8_t* __restrict__ out, const PngPlan pl) {
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
        uint8_t* f = out + (uint64_t)page * pl.stride + kHead + pl.zlib_bytes - 4;
        f[0] = (uint8_t)(b >> 8); f[1] = (uint8_t)b; f[2] = (uint8_t)(a >> 8); f[3] = (uint8_t)a;
    }
}

// CRC-32 of the IDAT chunk (type + data).  A block takes 32 KB of it: coalesced copy into shared memory (chunk rows
// padded by one word against bank conflicts), one 256-byte CRC per thread, a 7-level combine tree inside the block
// with the precomputed shifts x^(8 * 256 * 2^j), and one general shift by the bytes that follow the tile.
constexpr int kCrcThreads = 128, kCrcBytes = 256, kCrcTile = kCrcThreads * kCrcBytes, kCrcPitch = kCrcBytes + 4;

__global__ void __launch_bounds__(kCrcThreads) png_crc_kernel(const uint8_t* __restrict__ out, const PngPlan pl,
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
    const uint64_t total = 4 + pl.zlib_bytes;                                            // "IDAT" + data
    const uint64_t tile_begin = (uint64_t)blockIdx.x * kCrcTile;
    const int nbytes = (int)min((uint64_t)kCrcTile, total - tile_begin);
    const uint8_t* src = out + (uint64_t)blockIdx.y * pl.stride + (kHead - 4) + tile_begin;
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

__global__ void png_finish_kernel(uint8_t* __restrict__ out, const PngPlan pl, const uint32_t* __restrict__ crc, int n,
                                  unsigned long long* __restrict__ sizes) {
    const int page = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n) return;
    uint8_t* f = out + (uint64_t)page * pl.stride + kHead + pl.zlib_bytes;
    const uint32_t c = crc[page];
    const uint8_t tail[16] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c,
                              0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
    for (int i = 0; i < 16; ++i) f[i] = tail[i];
    if (sizes) sizes[page] = pl.file_bytes;
}

uint32_t host_crc(const uint8_t* p, size_t n) {
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; ++i) c = crc_step((c ^ p[i]) & 0xffu) ^ (c >> 8);
    return c ^ 0xffffffffu;
}
void be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

bool make_plan(int H, int W, int C, PngPlan& pl) {
    if (H <= 0 || W <= 0 || (C != 1 && C != 3 && C != 4)) return false;
    const uint64_t line = 1 + (uint64_t)W * C;
    if (line > 65535) return false;                                                      // one scanline must fit a stored block
    pl.H = H; pl.W = W; pl.C = C;
    pl.line = (uint32_t)line;
    pl.lines_per_block = (uint32_t)(65535 / line);
    pl.nblocks = ((uint32_t)H + pl.lines_per_block - 1) / pl.lines_per_block;
    pl.zlib_bytes = 2 + (uint64_t)pl.nblocks * 5 + (uint64_t)H * line + 4;
    if (pl.zlib_bytes > 0x7fffffffull) return false;
    pl.file_bytes = kHead + pl.zlib_bytes + 4 + 12;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t* h = pl.head;
    memcpy(h, sig, 8);
    be32(h + 8, 13); memcpy(h + 12, "IHDR", 4);
    be32(h + 16, (uint32_t)W); be32(h + 20, (uint32_t)H);
    h[24] = 8; h[25] = (uint8_t)(C == 1 ? 0 : (C == 3 ? 2 : 6)); h[26] = 0; h[27] = 0; h[28] = 0;      // 8 bit, grey / RGB / RGBA
    be32(h + 29, host_crc(h + 12, 17));
    be32(h + 33, (uint32_t)pl.zlib_bytes); memcpy(h + 37, "IDAT", 4);
    return true;
}

}  // namespace

size_t png_file_bytes(int H, int W, int C) {
    PngPlan pl{};
    return make_plan(H, W, C, pl) ? (size_t)pl.file_bytes : 0;
}

int launch_png_encode(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int C, uint8_t* d_out, size_t stride,
                      unsigned long long* d_sizes) {
    PngPlan pl{};
    if (n <= 0 || !make_plan(H, W, C, pl)) return set_err(ctx, PCS_ERR_ARG, "png_encode: unsupported shape %d x %d x %d", H, W, C);
    if ((stride & 3) || (reinterpret_cast<uintptr_t>(d_out) & 3)) return set_err(ctx, PCS_ERR_ARG, "png_encode: output and stride must be 4-byte aligned");
    if (stride < pl.file_bytes) return set_err(ctx, PCS_ERR_ARG, "png_encode: %zu bytes per file needed, stride is %zu", (size_t)pl.file_bytes, stride);
    if (H > 65535) return set_err(ctx, PCS_ERR_ARG, "png_encode: more than 65535 rows");
    pl.stride = stride;
    const size_t need = ((size_t)n * H * 2 * 8 + 255) / 256 * 256 + (size_t)n * 4 + 256;
    PCS_TRY(scratch_reserve(ctx, need));
    unsigned long long* ab = reinterpret_cast<unsigned long long*>(ctx->scratch);
    uint32_t* crc = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ctx->scratch) + ((size_t)n * H * 2 * 8 + 255) / 256 * 256);
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(crc, 0, (size_t)n * 4, st));
    const uint64_t words = (kHead + pl.zlib_bytes - 4 + 3) / 4;
    png_body_kernel<<<dim3((unsigned)((words + 255) / 256), n), 256, 0, st>>>(d_img, d_out, pl);
    PCS_LAUNCH_CHECK(ctx, "png_body_kernel");
    png_adler_rows_kernel<<<dim3(H, n), 256, 0, st>>>(d_img, pl, ab);
    PCS_LAUNCH_CHECK(ctx, "png_adler_rows_kernel");
    png_adler_kernel<<<n, 256, 0, st>>>(ab, d_out, pl);
    PCS_LAUNCH_CHECK(ctx, "png_adler_kernel");
    const uint64_t tiles = (4 + pl.zlib_bytes + kCrcTile - 1) / kCrcTile;
    png_crc_kernel<<<dim3((unsigned)tiles, n), kCrcThreads, 0, st>>>(d_out, pl, crc);
    PCS_LAUNCH_CHECK(ctx, "png_crc_kernel");
    png_finish_kernel<<<(n + 63) / 64, 64, 0, st>>>(d_out, pl, crc, n, d_sizes);
    PCS_LAUNCH_CHECK(ctx, "png_finish_kernel");
    return PCS_OK;
}

}  // namespace pcs
