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

template <typename SRC>
__device__ __forceinline__ double load_px(const SRC* p, size_t i) { return (double)p[i]; }

// One thread per output pixel.  grid = (ceil(Ws/32), ceil(Hs/8), n).
// SRC = uint8_t (plain page) or double (anti-aliased page, one page per launch).
template <typename SRC>
__global__ void __launch_bounds__(256)
resample_kernel(const SRC* __restrict__ grey, const uint8_t* __restrict__ bin, int H, int W, int Hs, int Ws,
                const uint32_t* __restrict__ level_bits, const int* __restrict__ page_list,
                double fmin_in, double fmax_in, uint8_t* __restrict__ image_out,
                uint8_t* __restrict__ binary_out) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= Ws || y >= Hs) return;
    const int page = page_list ? page_list[blockIdx.z] : blockIdx.z;
    const size_t src_off = (size_t)page * H * W;
    const size_t dst_off = (size_t)page * Hs * Ws + (size_t)y * Ws + x;

    const double fr = __ddiv_rn((double)H, (double)Hs);
    const double fc = __ddiv_rn((double)W, (double)Ws);
    const double r = __dadd_rn(__dmul_rn(fr, (double)y), __dsub_rn(__dmul_rn(0.5, fr), 0.5));
    const double c = __dadd_rn(__dmul_rn(fc, (double)x), __dsub_rn(__dmul_rn(0.5, fc), 0.5));

    if (binary_out) {
        // order 0: C round() (half away from zero), then reflect
        const int ri = reflect_coord((long long)round(r), H);
        const int ci = reflect_coord((long long)round(c), W);
        const uint8_t v = bin[src_off + (size_t)ri * W + ci];
        // bin = (1.0 - NN(binary/255 or binary)).astype(uint8): 1 iff v == 0
        binary_out[dst_off] = (v == 0) ? 1 : 0;
    }
    if (image_out) {
        double vmin, vmax;
        if constexpr (sizeof(SRC) == 1) {
            const uint32_t* bits = level_bits + page * 8;
            int lo = 0, hi = 255;
            for (int wv = 0; wv < 8; ++wv)
                if (bits[wv]) { lo = wv * 32 + __ffs(bits[wv]) - 1; break; }
            for (int wv = 7; wv >= 0; --wv)
                if (bits[wv]) { hi = wv * 32 + 31 - __clz(bits[wv]); break; }
            vmin = (double)lo;
            vmax = (double)hi;
        } else {
            vmin = fmin_in;
            vmax = fmax_in;
        }
        const double r0f = floor(r), c0f = floor(c);
        const double xr = __dsub_rn(r, r0f), xc = __dsub_rn(c, c0f);
        const long long r0 = (long long)r0f - 1, c0 = (long long)c0f - 1;
        int cols[4], rows[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cols[k] = reflect_coord(c0 + k, W);
            rows[k] = reflect_coord(r0 + k, H);
        }
        const SRC* g = grey + (sizeof(SRC) == 1 ? src_off : 0);
        double frow[4];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
            const size_t ro = (size_t)rows[pr] * W;
            frow[pr] = cubic_rn(xc, load_px(g, ro + cols[0]), load_px(g, ro + cols[1]), load_px(g, ro + cols[2]),
                                load_px(g, ro + cols[3]));
        }
        double v = cubic_rn(xr, frow[0], frow[1], frow[2], frow[3]);
        v = fmin(fmax(v, vmin), vmax);                       // clip=True
        // img = 1.0 - v/255 ; (img*255).astype(uint8)
        const double t = __dmul_rn(__dsub_rn(1.0, __ddiv_rn(v, 255.0)), 255.0);
        image_out[dst_off] = (uint8_t)(int)t;                // C truncation
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

template <typename SRC, int AXIS>
__global__ void __launch_bounds__(256)
gauss1d_kernel(const SRC* __restrict__ src, double* __restrict__ dst, int H, int W, int radius) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const double* wts = c_gauss_w[AXIS];
    auto at = [&](int d) -> double {
        if (AXIS == 0) return (double)src[(size_t)mirror_idx(y + d, H) * W + x];
        return (double)src[(size_t)y * W + mirror_idx(x + d, W)];
    };
    double tmp = __dmul_rn(at(0), wts[0]);
    for (int j = radius; j >= 1; --j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(-j), at(j)), wts[j]));
    dst[(size_t)y * W + x] = tmp;
}

__global__ void __launch_bounds__(256) minmax_f64_kernel(const double* __restrict__ p, size_t n,
                                                         unsigned long long* __restrict__ out /*[2]*/) {
    // values are >= 0 here (filtered uint8 levels), so the bit patterns order like the doubles
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
        atomicMin(out, (unsigned long long)__double_as_longlong(lo));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(hi));
    }
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

int launch_preprocess(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int Hs,
                      int Ws, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary) {
    if (n <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0) return set_err(ctx, PCS_ERR_ARG, "preprocess: bad shape");
    cudaStream_t st = ctx->stream;
    const size_t page_bytes = (size_t)H * W;
    PCS_TRY(scratch_reserve(ctx, (size_t)n * 8 * sizeof(uint32_t) + (size_t)n * sizeof(int) + 128));
    uint32_t* d_bits = reinterpret_cast<uint32_t*>(ctx->scratch);
    int* d_list = reinterpret_cast<int*>(d_bits + (size_t)n * 8);
    unsigned long long* d_mm = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<char*>(ctx->scratch) + (((size_t)n * 8 * 4 + (size_t)n * 4 + 15) / 16) * 16);

    std::vector<int> plain, aa;
    std::vector<uint32_t> h_bits((size_t)n * 8);
    if (d_image) {
        PCS_CUDA(ctx, cudaMemsetAsync(d_bits, 0, (size_t)n * 8 * sizeof(uint32_t), st));
        dim3 grid((unsigned)std::min<size_t>(296, (page_bytes / 16 + 255) / 256 + 1), n);
        level_bits_kernel<<<grid, 256, 0, st>>>(d_grey, page_bytes, d_bits);
        PCS_LAUNCH_CHECK(ctx, "level_bits_kernel");
        // the anti-aliasing decision (dataset.py:127) is data dependent: read the 32 B/page back
        PCS_CUDA(ctx, cudaMemcpyAsync(h_bits.data(), d_bits, h_bits.size() * 4, cudaMemcpyDeviceToHost, st));
        PCS_CUDA(ctx, cudaStreamSynchronize(st));
        for (int p = 0; p < n; ++p) {
            int levels = 0;
            for (int k = 0; k < 8; ++k) levels += __builtin_popcount(h_bits[(size_t)p * 8 + k]);
            (levels > 2 ? aa : plain).push_back(p);
        }
    } else {
        for (int p = 0; p < n; ++p) plain.push_back(p);
    }

    dim3 block(32, 8);
    if (!plain.empty()) {
        const int* list = nullptr;
        if ((int)plain.size() != n) {
            PCS_CUDA(ctx, cudaMemcpyAsync(d_list, plain.data(), plain.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            PCS_CUDA(ctx, cudaStreamSynchronize(st));   // plain is a stack vector
            list = d_list;
        }
        dim3 grid((Ws + 31) / 32, (Hs + 7) / 8, (unsigned)plain.size());
        resample_kernel<uint8_t><<<grid, block, 0, st>>>(d_grey, d_bin, H, W, Hs, Ws, d_bits, list, 0.0, 0.0, d_image,
                                                         d_binary);
        PCS_LAUNCH_CHECK(ctx, "resample_kernel<u8>");
    }
    if (!aa.empty()) {
        // per page: gaussian (axis 0, axis 1) into fp64 scratch, min/max, bicubic from fp64
        const double fr = (double)H / (double)Hs, fc = (double)W / (double)Ws;
        const double sig[2] = {std::max(0.0, (fr - 1.0) / 2.0), std::max(0.0, (fc - 1.0) / 2.0)};
        std::vector<double> w0, w1;
        int r0 = sig[0] > 1e-15 ? gauss_weights(sig[0], w0) : -1;
        int r1 = sig[1] > 1e-15 ? gauss_weights(sig[1], w1) : -1;
        if (r0 > kMaxGaussRadius || r1 > kMaxGaussRadius)
            return set_err(ctx, PCS_ERR_ARG, "preprocess: anti-aliasing radius %d/%d exceeds %d", r0, r1, kMaxGaussRadius);
        if (r0 >= 0) PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_gauss_w, w0.data(), w0.size() * 8, 0, cudaMemcpyHostToDevice, st));
        if (r1 >= 0)
            PCS_CUDA(ctx, cudaMemcpyToSymbolAsync(c_gauss_w, w1.data(), w1.size() * 8, sizeof(double) * (kMaxGaussRadius + 1),
                                                  cudaMemcpyHostToDevice, st));
        PCS_CUDA(ctx, cudaStreamSynchronize(st));
        // two fp64 planes of one page behind the bitmap block; growing the scratch drops its
        // contents, so the level bitmaps are re-uploaded from the host copy
        const size_t plane = page_bytes * sizeof(double);
        const size_t head = (((size_t)n * 8 * 4 + (size_t)n * 4 + 15) / 16) * 16 + 64;
        const size_t head_al = (head + 255) / 256 * 256;
        PCS_TRY(scratch_reserve(ctx, head_al + 2 * plane + 256));
        d_bits = reinterpret_cast<uint32_t*>(ctx->scratch);
        d_list = reinterpret_cast<int*>(d_bits + (size_t)n * 8);
        d_mm = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ctx->scratch) + head - 64);
        PCS_CUDA(ctx, cudaMemcpyAsync(d_bits, h_bits.data(), h_bits.size() * 4, cudaMemcpyHostToDevice, st));
        PCS_CUDA(ctx, cudaStreamSynchronize(st));
        double* t0 = reinterpret_cast<double*>(reinterpret_cast<char*>(ctx->scratch) + head_al);
        double* t1 = t0 + page_bytes;
        dim3 gfull((W + 31) / 32, (H + 7) / 8);
        for (int p : aa) {
            const uint8_t* src = d_grey + (size_t)p * page_bytes;
            const double* cur = nullptr;
            if (r0 >= 0) {
                gauss1d_kernel<uint8_t, 0><<<gfull, block, 0, st>>>(src, t0, H, W, r0);
                PCS_LAUNCH_CHECK(ctx, "gauss1d<0>");
                cur = t0;
            }
            if (r1 >= 0) {
                if (cur) gauss1d_kernel<double, 1><<<gfull, block, 0, st>>>(cur, t1, H, W, r1);
                else gauss1d_kernel<uint8_t, 1><<<gfull, block, 0, st>>>(src, t1, H, W, r1);
                PCS_LAUNCH_CHECK(ctx, "gauss1d<1>");
                cur = t1;
            }
            if (!cur) {   // both sigmas zero: plain conversion path
                gauss1d_kernel<uint8_t, 0><<<gfull, block, 0, st>>>(src, t0, H, W, 0);
                PCS_LAUNCH_CHECK(ctx, "gauss1d<copy>");
                cur = t0;
            }
            const unsigned long long init[2] = {0x7ff0000000000000ull, 0ull};
            PCS_CUDA(ctx, cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
            minmax_f64_kernel<<<296, 256, 0, st>>>(cur, page_bytes, d_mm);
            PCS_LAUNCH_CHECK(ctx, "minmax_f64");
            unsigned long long mm[2];
            PCS_CUDA(ctx, cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st));
            PCS_CUDA(ctx, cudaStreamSynchronize(st));
            double vmin, vmax;
            memcpy(&vmin, &mm[0], 8);
            memcpy(&vmax, &mm[1], 8);
            int pidx = p;
            PCS_CUDA(ctx, cudaMemcpyAsync(d_list, &pidx, sizeof(int), cudaMemcpyHostToDevice, st));
            PCS_CUDA(ctx, cudaStreamSynchronize(st));
            dim3 grid((Ws + 31) / 32, (Hs + 7) / 8, 1);
            resample_kernel<double><<<grid, block, 0, st>>>(cur, d_bin, H, W, Hs, Ws, d_bits, d_list, vmin, vmax, d_image,
                                                            d_binary);
            PCS_LAUNCH_CHECK(ctx, "resample_kernel<f64>");
        }
    }
    if (d_orig_binary) {
        const size_t nbytes = (size_t)n * page_bytes;
        orig_binary_kernel<<<(unsigned)std::min<size_t>(148 * 8, (nbytes / 16 + 255) / 256 + 1), 256, 0, st>>>(d_bin, nbytes,
                                                                                                          d_orig_binary);
        PCS_LAUNCH_CHECK(ctx, "orig_binary_kernel");
    }
    return PCS_OK;
}

// preserving_resize (util.py:21-29): order-0 resize of uint8 planes
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const uint8_t* __restrict__ src, int H, int W, uint8_t* __restrict__ dst, int Ho, int Wo) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= Wo || y >= Ho) return;
    const int page = blockIdx.z;
    const double fr = __ddiv_rn((double)H, (double)Ho);
    const double fc = __ddiv_rn((double)W, (double)Wo);
    const double r = __dadd_rn(__dmul_rn(fr, (double)y), __dsub_rn(__dmul_rn(0.5, fr), 0.5));
    const double c = __dadd_rn(__dmul_rn(fc, (double)x), __dsub_rn(__dmul_rn(0.5, fc), 0.5));
    const int ri = reflect_coord((long long)round(r), H);
    const int ci = reflect_coord((long long)round(c), W);
    dst[(size_t)page * Ho * Wo + (size_t)y * Wo + x] = src[(size_t)page * H * W + (size_t)ri * W + ci];
}

int launch_resize_nearest(pcs_ctx* ctx, const uint8_t* d_src, int n, int H, int W, uint8_t* d_dst, int Ho, int Wo) {
    if (n <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return set_err(ctx, PCS_ERR_ARG, "resize_nearest: bad shape");
    dim3 grid((Wo + 31) / 32, (Ho + 7) / 8, n), block(32, 8);
    resize_nearest_kernel<<<grid, block, 0, ctx->stream>>>(d_src, H, W, d_dst, Ho, Wo);
    PCS_LAUNCH_CHECK(ctx, "resize_nearest_kernel");
    return PCS_OK;
}

}  // namespace pcs
