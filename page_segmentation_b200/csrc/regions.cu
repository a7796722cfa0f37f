// Region extraction on the device: the pixel work of pc_segmentation.py and xycut.py
// (ocr4all_pixel_classifier/lib/), i.e. everything before the data-dependent recursion / contour tracing.
//
//   segment_masks   pc_segmentation.py:28-33,48,56  cv2.resize(INTER_NEAREST) -> 3x3 dilate -> colour filter,
//                   one fused gather kernel writing one u8 mask per requested colour
//   integral_image  xycut.py:135 np.count_nonzero(image, axis) over arbitrary sub-rectangles of the recursion:
//                   a summed-area table of (mask != 0); the host recursion only takes differences of its rows
//   text_regions    pc_segmentation.py:74-93  cv2.inRange -> close(k1) -> open(k2) -> dilate(k3) -> close(k3) with
//                   rectangular structuring elements, on 1-bit-per-pixel planes (1 MB for an A4 page, L2 resident)
//
// All three are integer work and bit-exact against OpenCV.  HBM-bound parts: inRange reads the RGB page once
// (3 B/px), the unpack writes two u8 planes (2 B/px); the morphology itself touches 1/8 B/px per pass.
#include "common.cuh"

#include <algorithm>
#include <cstdint>

namespace pcs {
namespace {

// ---- find_segments: NN resize + 3x3 dilate + colour filter ---------------------------------------------------
// cv2.resize(INTER_NEAREST): sx = min(floor(x * ifx), W-1) with ifx = 1 / ((double)Wo / W) (resizeNN);
// cv2.dilate(3x3): per-channel max over the neighbourhood, pixels outside the image are ignored.
struct Colours { uint8_t c[8][3]; };

__global__ void segment_masks_kernel(const uint8_t* __restrict__ rgb, int H, int W, int Ho, int Wo, double ify, double ifx,
                                     Colours cols, int m, uint8_t* __restrict__ masks) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Wo || y >= Ho) return;
    int r = 0, g = 0, b = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= Ho) continue;
        const int sy = min((int)floor(__dmul_rn((double)yy, ify)), H - 1);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= Wo) continue;
            const int sx = min((int)floor(__dmul_rn((double)xx, ifx)), W - 1);
            const uint8_t* p = rgb + ((size_t)sy * W + sx) * 3;
            r = max(r, (int)__ldg(p)); g = max(g, (int)__ldg(p + 1)); b = max(b, (int)__ldg(p + 2));
        }
    }
    for (int i = 0; i < m; ++i)
        masks[((size_t)i * Ho + y) * Wo + x] = (r == cols.c[i][0] && g == cols.c[i][1] && b == cols.c[i][2]) ? 1 : 0;
}

// cv2.dilate(image, np.ones((3, 3))) for an interleaved uint8 image [H][W][C] (pc_segmentation.py:63-67)
__global__ void dilate3x3_kernel(const uint8_t* __restrict__ src, int H, int W, int C, uint8_t* __restrict__ dst) {
    const int xc = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (xc >= W * C || y >= H) return;
    const int x = xc / C;
    int v = 0;
    for (int yy = max(0, y - 1); yy <= min(H - 1, y + 1); ++yy)
        for (int dx = -1; dx <= 1; ++dx)
            if (x + dx >= 0 && x + dx < W) v = max(v, (int)__ldg(src + (size_t)yy * W * C + xc + dx * C));
    dst[(size_t)y * W * C + xc] = (uint8_t)v;
}

// ---- summed-area table of (mask != 0): sat[(H+1) x (W+1)], first row / column zero ----------------------------
// pass 1: one warp per row, inclusive scan in 32-pixel chunks with a running carry
__global__ void sat_rows_kernel(const uint8_t* __restrict__ mask, int n, int H, int W, int32_t* __restrict__ sat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n * (H + 1)) return;
    const int img = warp / (H + 1), row = warp % (H + 1);
    int32_t* out = sat + ((size_t)img * (H + 1) + row) * (W + 1);
    if (row == 0) {
        for (int x = lane; x <= W; x += 32) out[x] = 0;
        return;
    }
    const uint8_t* in = mask + ((size_t)img * H + (row - 1)) * W;
    int carry = 0;
    if (lane == 0) out[0] = 0;
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const unsigned bits = __ballot_sync(0xffffffffu, x < W && in[x] != 0);
        const int incl = __popc(bits & (0xffffffffu >> (31 - lane)));
        if (x < W) out[x + 1] = carry + incl;
        carry += __popc(bits);
    }
}
// pass 2: one thread per column, running sum down the rows (coalesced across the warp)
__global__ void sat_cols_kernel(int n, int H, int W, int32_t* __restrict__ sat) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (W + 1)) return;
    const int img = t / (W + 1), x = t % (W + 1);
    int32_t* p = sat + (size_t)img * (H + 1) * (W + 1) + x;
    int acc = 0;
    // eight rows per step: the loads are issued together instead of one dependent load -> add -> store chain per row
    int y = 1;
    for (; y + 7 <= H; y += 8) {
        int v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = p[(size_t)(y + k) * (W + 1)];
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc += v[k]; p[(size_t)(y + k) * (W + 1)] = acc; }
    }
    for (; y <= H; ++y) {
        acc += p[(size_t)y * (W + 1)];
        p[(size_t)y * (W + 1)] = acc;
    }
}

// ---- get_text_contours: inRange + rectangular morphology on bit planes ----------------------------------------
// plane layout: [H][WW] uint32, bit b of word j = pixel 32 j + b; bits at x >= W are always zero.
__global__ void inrange_pack_kernel(const uint8_t* __restrict__ rgb, int H, int W, int WW, uint8_t c0, uint8_t c1, uint8_t c2,
                                    uint32_t* __restrict__ plane) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool hit = false;
    if (x < W) {
        const uint8_t* p = rgb + ((size_t)y * W + x) * 3;
        hit = __ldg(p) == c0 && __ldg(p + 1) == c1 && __ldg(p + 2) == c2;
    }
    const unsigned bits = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < WW) plane[(size_t)y * WW + (x >> 5)] = bits;
}

__device__ __forceinline__ uint32_t valid_bits(int j, int W) {       // mask of the bits of word j that are pixels
    const int left = W - 32 * j;
    return left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : (0xffffffffu >> (32 - left)));
}

// OpenCV erode/dilate with a k x k rectangle anchored at k/2: out(x) = op over i in [0,k) of in(x + i - k/2);
// dilate = OR with zeros outside, erode = AND with ones outside = ~dilate(~in restricted to the image).
template <bool ERODE>
__global__ void morph_h_kernel(const uint32_t* __restrict__ in, int H, int W, int WW, int k, uint32_t* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (j >= WW) return;
    const uint32_t* row = in + (size_t)y * WW;
    const int a = k / 2;
    // words j-q0 .. j+q1 cover the window; fetch lazily through a small lambda
    auto word = [&](int idx) -> uint32_t {
        if (idx < 0 || idx >= WW) return 0u;
        const uint32_t w = row[idx];
        return ERODE ? (~w & valid_bits(idx, W)) : w;
    };
    uint32_t acc = 0;
    // offset d = i - a in [-a, k-1-a]; bit b of result takes bit (32 j + b + d)
    int d = -a;
    const int dmax = k - 1 - a;
    while (d <= dmax) {
        const int q = (d >= 0) ? (d >> 5) : -((-d + 31) >> 5);      // floor(d / 32)
        const int r0 = d - 32 * q;                                  // in [0, 32)
        const uint32_t lo = word(j + q), hi = word(j + q + 1);
        // all offsets with the same q share (lo, hi): r runs from r0 to min(31, dmax - 32 q)
        const int r1 = min(31, dmax - 32 * q);
        if (lo | hi)
            for (int r = r0; r <= r1; ++r) acc |= __funnelshift_r(lo, hi, r);
        d += r1 - r0 + 1;
    }
    acc &= valid_bits(j, W);
    out[(size_t)y * WW + j] = ERODE ? (~acc & valid_bits(j, W)) : acc;
}

template <bool ERODE>
__global__ void morph_v_kernel(const uint32_t* __restrict__ in, int H, int W, int WW, int k, uint32_t* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= WW || y >= H) return;
    const int a = k / 2;
    const int y0 = max(0, y - a), y1 = min(H - 1, y - a + k - 1);
    const uint32_t vb = valid_bits(j, W);
    uint32_t acc = 0;
    for (int yy = y0; yy <= y1; ++yy) {
        const uint32_t w = in[(size_t)yy * WW + j];
        acc |= ERODE ? (~w & vb) : w;
    }
    out[(size_t)y * WW + j] = ERODE ? (~acc & vb) : acc;
}

// bit planes -> u8 images: text_inv = 255 - 255*text (pc_segmentation.py:96), region = 255*region
__global__ void unpack_kernel(const uint32_t* __restrict__ text, const uint32_t* __restrict__ region, int H, int W, int WW,
                              uint8_t* __restrict__ text_inv, uint8_t* __restrict__ region_u8) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t wi = (size_t)y * WW + (x >> 5);
    const uint32_t t = text[wi], r = region[wi];
    const size_t o = (size_t)y * W + x;
    if (text_inv) text_inv[o] = ((t >> (x & 31)) & 1u) ? 0 : 255;
    if (region_u8) region_u8[o] = ((r >> (x & 31)) & 1u) ? 255 : 0;
}

template <bool ERODE>
int morph(pcs_ctx* ctx, const uint32_t* in, uint32_t* tmp, uint32_t* out, int H, int W, int WW, int k) {
    {
        dim3 block(128), grid((WW + 127) / 128, H);
        morph_h_kernel<ERODE><<<grid, block, 0, ctx->stream>>>(in, H, W, WW, k, tmp);
        PCS_LAUNCH_CHECK(ctx, "morph_h_kernel");
    }
    {
        dim3 block(32, 8), grid((WW + 31) / 32, (H + 7) / 8);
        morph_v_kernel<ERODE><<<grid, block, 0, ctx->stream>>>(tmp, H, W, WW, k, out);
        PCS_LAUNCH_CHECK(ctx, "morph_v_kernel");
    }
    return PCS_OK;
}

}  // namespace

int launch_segment_masks(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, int Ho, int Wo, const uint8_t* colours, int m,
                         uint8_t* d_masks) {
    if (m < 1 || m > 8) return set_err(ctx, PCS_ERR_ARG, "segment_masks: 1..8 colours per call");
    Colours cols{};
    for (int i = 0; i < m; ++i)
        for (int c = 0; c < 3; ++c) cols.c[i][c] = colours[i * 3 + c];
    // cv::resize: inv_scale = (double)dsize / ssize; the nearest-neighbour table uses 1 / inv_scale
    const double ifx = 1.0 / ((double)Wo / (double)W), ify = 1.0 / ((double)Ho / (double)H);
    dim3 block(32, 8), grid((Wo + 31) / 32, (Ho + 7) / 8);
    segment_masks_kernel<<<grid, block, 0, ctx->stream>>>(d_rgb, H, W, Ho, Wo, ify, ifx, cols, m, d_masks);
    PCS_LAUNCH_CHECK(ctx, "segment_masks_kernel");
    return PCS_OK;
}

int launch_dilate3x3(pcs_ctx* ctx, const uint8_t* d_src, int H, int W, int C, uint8_t* d_dst) {
    dim3 block(64, 4), grid((W * C + 63) / 64, (H + 3) / 4);
    dilate3x3_kernel<<<grid, block, 0, ctx->stream>>>(d_src, H, W, C, d_dst);
    PCS_LAUNCH_CHECK(ctx, "dilate3x3_kernel");
    return PCS_OK;
}

int launch_integral_image(pcs_ctx* ctx, const uint8_t* d_mask, int n, int H, int W, int32_t* d_sat) {
    {
        const long long warps = (long long)n * (H + 1);
        const int block = 256;
        const long long blocks = (warps * 32 + block - 1) / block;
        sat_rows_kernel<<<(unsigned)blocks, block, 0, ctx->stream>>>(d_mask, n, H, W, d_sat);
        PCS_LAUNCH_CHECK(ctx, "sat_rows_kernel");
    }
    {
        const int t = n * (W + 1);
        sat_cols_kernel<<<(t + 127) / 128, 128, 0, ctx->stream>>>(n, H, W, d_sat);
        PCS_LAUNCH_CHECK(ctx, "sat_cols_kernel");
    }
    return PCS_OK;
}

int launch_text_regions(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, const uint8_t* colour, int k_close, int k_open,
                        int k_region, uint8_t* d_text_inv, uint8_t* d_region) {
    const int WW = (W + 31) / 32;
    const size_t plane = ((size_t)H * WW * 4 + 255) / 256 * 256;
    PCS_TRY(scratch_reserve(ctx, 4 * plane));
    char* base = reinterpret_cast<char*>(ctx->scratch);
    uint32_t* p0 = reinterpret_cast<uint32_t*>(base);
    uint32_t* p1 = reinterpret_cast<uint32_t*>(base + plane);
    uint32_t* p2 = reinterpret_cast<uint32_t*>(base + 2 * plane);
    uint32_t* tmp = reinterpret_cast<uint32_t*>(base + 3 * plane);
    {
        dim3 block(256), grid((WW * 32 + 255) / 256, H);
        inrange_pack_kernel<<<grid, block, 0, ctx->stream>>>(d_rgb, H, W, WW, colour[0], colour[1], colour[2], p0);
        PCS_LAUNCH_CHECK(ctx, "inrange_pack_kernel");
    }
    // closing (pc_segmentation.py:81-82), opening (:83-84)
    PCS_TRY(morph<false>(ctx, p0, tmp, p1, H, W, WW, k_close));
    PCS_TRY(morph<true>(ctx, p1, tmp, p0, H, W, WW, k_close));
    PCS_TRY(morph<true>(ctx, p0, tmp, p1, H, W, WW, k_open));
    PCS_TRY(morph<false>(ctx, p1, tmp, p0, H, W, WW, k_open));          // p0 = `image` after noise removal
    // region_chars = dilate (:92), region_text = close (:93)
    PCS_TRY(morph<false>(ctx, p0, tmp, p1, H, W, WW, k_region));
    PCS_TRY(morph<false>(ctx, p1, tmp, p2, H, W, WW, k_region));
    PCS_TRY(morph<true>(ctx, p2, tmp, p1, H, W, WW, k_region));         // p1 = region_text
    {
        dim3 block(256), grid((W + 255) / 256, H);
        unpack_kernel<<<grid, block, 0, ctx->stream>>>(p0, p1, H, W, WW, d_text_inv, d_region);
        PCS_LAUNCH_CHECK(ctx, "unpack_kernel");
    }
    return PCS_OK;
}

}  // namespace pcs
