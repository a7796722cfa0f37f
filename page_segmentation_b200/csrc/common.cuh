// Shared declarations for the pcseg_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/pcseg_b200.h"

namespace pcs {

constexpr int kMaxClasses = 16;

// ---------------------------------------------------------------------------
// Activation tensors: plane-major "NC/8HW8": [n][cp/8][h][w][8 channels], i.e. one
// 16-byte unit per (pixel, 8-channel group); channel count padded to a multiple
// of 8 (tensor engine) or 16 (CUDA-core engine) with zeros.  A plane row is contiguous in memory, which is what makes the
// TMA boxes of the tensor-core kernel 2 KB wide and every epilogue store a
// 512-byte coalesced warp store.
// ---------------------------------------------------------------------------
struct Act {
    void* p = nullptr;      // device pointer, element type = model precision (bf16 / fp16)
    int n = 0, h = 0, w = 0;
    int c = 0;              // real channels
    int cp = 0;             // padded channels (stride between pixels, elements)
    size_t bytes() const { return (size_t)n * h * w * cp * 2; }
};

inline int pad16(int c) { return (c + 15) / 16 * 16; }
inline int pad8(int c) { return (c + 7) / 8 * 8; }

// element index of channel c at (page, y, x) of an activation with cp padded channels
__host__ __device__ __forceinline__ size_t act_idx(int page, int cp, int h, int w, int c, int y, int x) {
    return ((((size_t)page * (cp >> 3) + (c >> 3)) * h + y) * w + x) * 8 + (c & 7);
}

// One network layer after host-side weight transformation.
struct Layer {
    std::string name;
    int kind = 0;           // 0 conv (same, s1), 1 deconv s1 (stored flipped as conv), 2 deconv 2x2 s2, 3 logits 1x1
    int k = 0;              // kernel size
    int cin = 0, cout = 0;  // real channel counts (cin over all concatenated sources)
    int relu = 0;
    // fp32 weights in "correlation" form [tap][cin][cout] (deconv s1 flipped + transposed;
    // deconv s2: tap = i*2+j) and bias [cout]; device copies.
    float* d_w32 = nullptr;
    float* d_b32 = nullptr;
    std::vector<float> h_w32, h_b32;
    // tensor-core operand image (precision type): see conv_umma.cu for the layout
    void* d_wmma = nullptr;
    size_t wmma_bytes = 0;
    int npad = 0;           // padded C_out of the UMMA tile
    int co_t = 0;           // deconv s2 on the tensor path: padded channels per tap
    // launches of the dy-folded marching kernel (conv_fold.cu) that together compute this layer: a layer whose
    // resident weights do not fit shared memory is split along N (output channels) or along K (one part per
    // concatenated source, fp32 partial sums handed from the first part to the last)
    struct FoldPart {
        void* d_w = nullptr;      // resident operand image
        void* d_w_px = nullptr;   // ... for a source whose odd last plane arrives as pixel-pair units (conv2 <- conv1)
        int src = 0;              // concatenated source read by this part
        int o0 = 0, ncols = 0;    // output channels [o0, o0 + ncols)
        int npad = 0, nplanes = 0;
        int psum = 0;             // 0 complete, 1 writes partial sums, 2 adds them
        bool both = false;        // reads BOTH concatenated sources in one launch (3x3, U-Net conv9a)
    };
    std::vector<FoldPart> fold;
    void* d_wmma_px = nullptr;    // conv1 (5x5, 20 channels): operand image with the pixel-pair columns 20..23
    void* d_w12 = nullptr;        // conv1 / conv2 of the FCN variants: operand image of the fused kernel (conv12_fused.cu)
    std::vector<float> h_w32_raw; // deconv5, U-Net up*: weights before rounding (composed / summed on the host, rounded once)
    float* d_head_lw = nullptr;   // logits layer: [32][4] rows of the conv2 skip channels, zero padded (fcn_skip)
    float* d_head_lb = nullptr;   // logits layer: [4] bias with the deconv5 / conv2 biases folded in
    int nchunks = 0;        // number of 16-channel K chunks over all sources
};

struct StageTime { std::string name; cudaEvent_t e0, e1; };

}  // namespace pcs

struct pcs_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // model
    int arch = -1, n_classes = 0, precision = PCS_PREC_BF16, engine = PCS_ENGINE_UMMA;
    std::vector<pcs::Layer> layers;
    bool model_ready = false;
    int64_t model_stamp = 0;                // process-unique id of the loaded model (never reused)

    // workspace arena (activations), grown on demand
    char* arena = nullptr;
    size_t arena_bytes = 0;
    size_t arena_used = 0;
    std::map<std::string, pcs::Act> acts;   // activations of the last forward (debug / reuse)

    // small scratch (+ a second one for the max_width pass, which runs the first pass inside)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* scratch2 = nullptr;
    size_t scratch2_bytes = 0;

    // staging buffers of pcs_predict_pages_host
    char* stage = nullptr;
    size_t stage_bytes = 0;
    cudaStream_t copy_streams[2] = {nullptr, nullptr};     // H2D / D2H streams of the host pipeline
    static constexpr int kHostBufs = 4;                    // staging buffers the host pipeline may rotate over
    cudaEvent_t ev_h2d[kHostBufs] = {}, ev_comp[kHostBufs] = {}, ev_d2h[kHostBufs] = {}, ev_sizes[kHostBufs] = {}, ev_fork = nullptr;
    cudaStream_t aux_stream = nullptr;                     // side stream of the preprocess: the (normally empty) general-path launches
    cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;  // run beside the two-level resampler instead of after it
    uint64_t* h_png_sizes = nullptr;                       // pinned [kHostBufs][3][64]: file lengths of the chunk in flight (PNG mode)
    // streaming form of the host-buffer calls (pcs_predict_pages_*_submit / pcs_wait_pages): a submitted call returns once
    // its work is queued; the next submit with the same staging layout CHAINS onto it (the staging buffers keep rotating,
    // ordered by the per-buffer events only), so its upload runs under the kernels of the call before
    static constexpr int kCallRing = 8;
    cudaEvent_t ev_call[kCallRing] = {};                   // "results of submit t are in the caller's buffers" (slot t % kCallRing)
    uint64_t call_seq = 0;                                 // submits so far = the next ticket
    uint64_t host_seq = 0;                                 // chunks queued by the chain so far (staging buffer = seq % nbuf)
    bool host_chain_ok = false;                            // the last host call was a submit that queued completely
    std::vector<size_t> host_layout;                       // staging layout of that call

    void* writer = nullptr;                 // pcs::OutputWriter of pcs_output_pages (output.cu), created on first use

    std::string timings;
    std::vector<pcs::StageTime> stage_times;
    bool timing_enabled = false;
    bool pdl = false;                       // launch the tensor-core kernels with programmatic dependent launch (PCSEG_PDL=1 / pcs_set_pdl);
                                            // measured neutral on B200 (interleaved A/B), so off by default
    bool keep_acts = false;                 // diagnostics: also store activations the fused kernels normally skip

    // fp16 stores saturate at +-65504 (umma_ptx.cuh pack2<__half>): the stored activations of a forward are scanned for
    // saturated values -- mode 1 (default): on the first forward after every model load, 2: on every forward, 0: never
    int sat_mode = 1;
    bool sat_pending = false;               // a model has been loaded and not yet checked
    unsigned long long* d_sat_count = nullptr;   // saturated stored values found since the model was loaded
};

namespace pcs {

int set_err(pcs_ctx* ctx, int code, const char* fmt, ...);

#define PCS_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return pcs::set_err((ctx), PCS_ERR_CUDA, "%s failed: %s (%s:%d)", #call,         \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

#define PCS_LAUNCH_CHECK(ctx, what)                                                          \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return pcs::set_err((ctx), PCS_ERR_CUDA, "launch of %s failed: %s (%s:%d)", what, \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

#define PCS_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != PCS_OK) return rc__; \
    } while (0)

// Dynamic shared memory every tensor-core kernel asks for at least: more than half an SM, so that two of their CTAs
// never share an SM (each allocates all 512 TMEM columns: a second CTA would sit in tcgen05.alloc while other SMs
// idle - which is what the block scheduler does when CTAs trickle in under programmatic dependent launch).
constexpr size_t kSoloSmem = 116 * 1024;

// kernel launch with the programmatic-dependent-launch attribute (see ptx::griddep_wait); `pdl` false = plain launch
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// arena
int arena_reserve(pcs_ctx* ctx, size_t bytes);
void* arena_alloc(pcs_ctx* ctx, size_t bytes);      // 256-B aligned bump allocation; nullptr if full
int scratch_reserve(pcs_ctx* ctx, size_t bytes);

// ---- kernels implemented in the other translation units --------------------
// preprocess.cu
int launch_preprocess(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W,
                      int Hs, int Ws, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary);
int launch_preprocess_max_width(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int H1, int W1,
                                int H2, int W2, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary);
int launch_preprocess_bits(pcs_ctx* ctx, const uint32_t* d_bitmap, size_t bitmap_words, int n, int H, int W, int level0, int level1,
                           int Hs, int Ws, uint8_t* d_image, uint8_t* d_binary);
int launch_pack_bits(pcs_ctx* ctx, const uint8_t* d_src, int n, size_t npix, uint32_t* d_dst, size_t words_per_page);
int launch_unpack_bits(pcs_ctx* ctx, const uint32_t* d_src, int n, size_t words_per_page, size_t npix, uint8_t* d_dst);
int launch_resize_nearest(pcs_ctx* ctx, const uint8_t* d_src, int n, int H, int W, uint8_t* d_dst,
                          int Ho, int Wo);

// conv_direct.cu  (CUDA-core fp32-accumulate kernels)
struct ConvSrc {
    const void* p = nullptr;   // activation (precision type) or uint8 image for the first layer
    int c = 0;                 // real channels taken from this source
    int cp = 0;                // pixel stride in elements
};
struct DirectConvArgs {
    ConvSrc src[2];
    int nsrc = 1;
    int src_u8 = 0;            // src[0] is the uint8 network input (value/255), H,W below are the padded grid
    int img_h = 0, img_w = 0;  // real (unpadded) size of the uint8 input
    int upsample = 0;          // read input at (y/2, x/2): UpSampling2D fused (U-Net up*)
    int fast_first = 0;        // allow the specialised first-layer kernel (constant-bank weights)
    int n = 0, h = 0, w = 0;   // output grid (== input grid unless upsample)
    int k = 0, pad = 0;        // kernel size, pad-before
    const float* w32 = nullptr;   // [k*k][cin][cout]
    const float* b32 = nullptr;
    int cin = 0, cout = 0, relu = 0;
    void* out = nullptr;       // full-resolution output (may be null if only pooled output is needed)
    int out_cp = 0;
    void* pool_out = nullptr;  // optional fused 2x2 max-pool output
    int pool_cp = 0;
};
int launch_conv_direct(pcs_ctx* ctx, const DirectConvArgs& a);

struct DeconvS2Args {
    ConvSrc src[2];
    int nsrc = 1;
    int n = 0, h = 0, w = 0;   // INPUT grid; output is 2h x 2w
    const float* w32 = nullptr;   // [4][cin][cout], tap = i*2+j
    const float* b32 = nullptr;
    int cin = 0, cout = 0, relu = 0;
    void* out = nullptr;
    int out_cp = 0;
};
int launch_deconv_s2_direct(pcs_ctx* ctx, const DeconvS2Args& a);

// epilogue.cu
struct HeadArgs {
    // optional fused stride-2 deconv feeding the logits (FCN deconv5): inputs at half resolution
    int has_deconv = 0;
    ConvSrc dsrc[2];
    int dnsrc = 0;
    const float* dw32 = nullptr;   // [4][dcin][dcout]
    const float* db32 = nullptr;
    int dcin = 0, dcout = 0;
    // optional direct (full-resolution) source concatenated after the deconv output
    ConvSrc skip;
    int has_skip = 0;
    // logits 1x1: [cin_total][n_classes], bias
    const float* lw32 = nullptr;
    const float* lb32 = nullptr;
    int n_classes = 0;
    int n = 0, hp = 0, wp = 0;     // padded full-resolution grid of the sources
    int h = 0, w = 0;              // cropped output size
    const uint8_t* binary = nullptr;
    uint8_t* labels = nullptr;
    float* logits = nullptr;
    float* prob = nullptr;
    const uint8_t* lut = nullptr;  // device copy, n_classes*3
    uint8_t* color = nullptr;
    uint8_t* overlay = nullptr;
    uint8_t* inverted = nullptr;
};
int launch_head(pcs_ctx* ctx, const HeadArgs& a);
int launch_masks(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W,
                 const uint8_t* d_lut, int n_lut, uint8_t* d_color, uint8_t* d_overlay, uint8_t* d_inverted);

int launch_saturation_scan(pcs_ctx* ctx, const void* d_act, size_t n_halves, unsigned long long* d_count);
int launch_eval_counts(pcs_ctx* ctx, const uint8_t* d_pred, const uint8_t* d_mask, const uint8_t* d_bin, size_t n, int n_classes,
                       unsigned long long* d_out);

// ccl.cu
int launch_ccl(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int32_t* d_labels,
               int32_t* d_stats, int max_components, int32_t* d_ncomp);
int launch_cc_majority(pcs_ctx* ctx, uint8_t* d_pred, const uint8_t* d_binary, int n, int H, int W,
                       int n_classes);
int launch_char_height(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int inverse, int32_t* d_out);
int launch_bounding_boxes(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes,
                          uint8_t* d_out);
int launch_class_components(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, int32_t* d_stats,
                            int max_components, int32_t* d_ncomp);

// regions.cu
int launch_segment_masks(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, int Ho, int Wo, const uint8_t* colours, int m,
                         uint8_t* d_masks);
int launch_dilate3x3(pcs_ctx* ctx, const uint8_t* d_src, int H, int W, int C, uint8_t* d_dst);
int launch_integral_image(pcs_ctx* ctx, const uint8_t* d_mask, int n, int H, int W, int32_t* d_sat);
int launch_text_regions(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, const uint8_t* colour, int k_close, int k_open,
                        int k_region, uint8_t* d_text_inv, uint8_t* d_region);

// output.cu
int output_pages(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W, const uint8_t* lut, int n_lut,
                 const char* const* paths);
int output_flush(pcs_ctx* ctx);
void output_writer_destroy(pcs_ctx* ctx);

// png.cu
size_t png_file_bytes(int H, int W, int C, int level);
int launch_png_encode(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int C, int level, uint8_t* d_out, size_t stride,
                      unsigned long long* d_sizes);
int png_index_depth(int ncolors);
size_t png_indexed_file_bytes(int H, int W, int ncolors, int level);
int launch_png_encode_indexed(pcs_ctx* ctx, const uint8_t* d_idx, int n, int H, int W, const uint8_t* h_palette, int ncolors, int level,
                              uint8_t* d_out, size_t stride, unsigned long long* d_sizes);
int launch_mask_indices(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W, int n_lut, uint8_t* d_out);

// train.cu  (fp32 CUDA-core training primitives, planar [C][H][W])
int train_corr2d(pcs_ctx* ctx, const float* x, const float* w, const float* b, float* y, int Ci, int Co, int H, int W, int k, int relu, int acc);
int train_wgrad(pcs_ctx* ctx, const float* x, const float* dy, float* dw, int Ci, int Co, int H, int W, int k);
int train_plane_sum(pcs_ctx* ctx, const float* dy, float* db, int C, size_t plane);
int train_relu_bwd(pcs_ctx* ctx, float* dy, const float* y, size_t n);
int train_maxpool(pcs_ctx* ctx, const float* x, float* y, const float* dy, float* dx, int C, int H, int W, int acc);
int train_deconv2(pcs_ctx* ctx, int mode, const float* x, const float* k2, const float* b, float* y, const float* dy, float* dx, float* dk2,
                  int Ci, int Co, int h, int w, int relu);
int train_softmax_ce(pcs_ctx* ctx, const float* logits, const uint8_t* labels, int C, int H, int W, int Hc, int Wc, float* dlogits, double* loss_sum);
int train_input_plane(pcs_ctx* ctx, const uint8_t* img, int h, int w, float* out, int H, int W);
int train_adam(pcs_ctx* ctx, float* p, const float* g, float* m, float* v, const long long* d_offsets, int nvars, float lr_t, float b1,
               float b2, float eps, float clipnorm, float gscale);

// train_tc.cu  (training step on the tensor cores)
struct TrainTc;
int train_tc_create(pcs_ctx* ctx, int arch, int n_classes, int h, int w, const long long* offsets, int n_offsets, TrainTc** out);
int train_tc_step(pcs_ctx* ctx, TrainTc* t, int phases, const uint8_t* d_img, const uint8_t* d_labels, const float* d_params, float* d_grads,
                  double* d_loss);
int train_tc_destroy(pcs_ctx* ctx, TrainTc* t);
int train_tc_wgrad(pcs_ctx* ctx, const void* d_x, int x_planes, const void* d_dy, int dy_planes, int H, int W, int k, int ci, int co, float* d_dw);

// conv_umma.cu  (tcgen05 / TMEM / TMA implicit GEMM)
struct UmmaHeadArgs {              // fused FCN head epilogue (conv_umma.cu mode 2)
    const void* plog = nullptr;         // device float4 [n][2h][2w]: conv2 share of the logits (fcn_skip) or null
    const float* lb_folded = nullptr;   // device [4]: logits bias with the deconv5 (and conv2) bias folded through
    int n_classes = 0, hs = 0, ws = 0;
    uint8_t* labels = nullptr; float* logits = nullptr; float* prob = nullptr;
};
struct UmmaConvArgs {
    ConvSrc src[2];
    int nsrc = 1;
    int n = 0, h = 0, w = 0;
    int k = 0, pad = 0;
    const void* wmma = nullptr;    // pre-arranged operand image, see conv_umma.cu
    const float* b32 = nullptr;
    int cout = 0, npad = 0, nchunks = 0, relu = 0;
    int mode = 0;                  // 0 = 'same' conv store (+pool), 1 = 2x2 stride-2 transposed conv scatter,
                                   // 2 = fused FCN head (composed deconv5 x logits GEMM)
    const UmmaHeadArgs* head = nullptr;
    int co_t = 0;                  // mode 1: padded channels per tap
    void* out = nullptr;
    int out_cp = 0;
    void* pool_out = nullptr;
    int pool_cp = 0;
};
int launch_conv_umma(pcs_ctx* ctx, const UmmaConvArgs& a);
// host-side operand image builder; returns bytes written (precision: PCS_PREC_*)
size_t umma_weight_image(const float* w32 /*[taps][cin][cout]*/, int taps, const int* src_c, int nsrc,
                         int cout, int npad, int precision, std::vector<uint16_t>& out);
size_t umma_weight_image_deconv(const float* w32 /*[4][cin][cout]*/, const int* src_c, int nsrc, int cout, int co_t,
                                int npad, int precision, std::vector<uint16_t>& out);
size_t umma_weight_image_up2(const float* w32 /*[4][cin][cout]*/, int cin, int cout, int co_t, int npad, int precision,
                             std::vector<uint16_t>& out);
size_t umma_weight_image_head(const double* m /*[4][cin][4]*/, const int* src_c, int nsrc, int precision,
                              std::vector<uint16_t>& out);
bool umma_supported(int k, int npad);

// conv_fold.cu  (marching, dx-folded 5x5 convolution for the small-channel layers)
struct FoldConvArgs {
    ConvSrc src;
    int n = 0, h = 0, w = 0, k = 5;
    const void* wimg = nullptr;
    const float* h_bias = nullptr;     // HOST pointer: the bias travels in the kernel parameter block
    int cout = 0, npad = 0, nplanes = 0, relu = 0;      // nplanes: 8-channel planes of the source
    void* out = nullptr; int out_cp = 0;
    void* pool_out = nullptr; int pool_cp = 0;
    int o0 = 0;                        // first output channel written by this launch (N split; whole planes)
    void* psum_out = nullptr;          // K split: fp32 partial sums [n][npad/4][h][w][4] written instead of the output ...
    const void* psum_in = nullptr;     // ... and added by the last part before bias / activation
    void* plog = nullptr;              // optional float4 [n][h][w]: this layer's share of the logits (fcn_skip conv2)
    const float* skip_lw = nullptr;    // device [32][4]: logits rows of this layer's channels, zero padded
    ConvSrc src2;                      // 3x3 layers on a concatenation: the second tensor (same geometry and plane count as src)
    const void* pair_src = nullptr;    // the source's last (odd) plane as pixel-pair units [n][h][w + 1][8] (conv1_umma.cu): src holds the
                                       // whole planes before it, nplanes counts it, wimg is the image built with pairx
};
bool fold_supported(int k, int npad, int nplanes);
size_t fold_weight_image(const float* w32 /*[25][cin_total][cout_total]*/, int cin_total, int cout_total, int ci0, int cin,
                         int o0, int ncols, int npad, int precision, std::vector<uint16_t>& out, bool pairx = false, int ks = 5);
int launch_conv_fold(pcs_ctx* ctx, const FoldConvArgs& a);

// conv12_fused.cu  (conv1 + conv2 + MaxPool of the FCN variants in one marching kernel; conv1 never leaves the SM)
struct Conv12Args {
    const uint8_t* d_image = nullptr;
    int n = 0, img_h = 0, img_w = 0, h = 0, w = 0;
    const void* w1img = nullptr;       // conv12_weight_image1
    const void* w2img = nullptr;       // conv12_weight_image2
    const float* h_bias1 = nullptr;    // HOST pointers: the biases travel in the kernel parameter block
    const float* h_bias2 = nullptr;
    int cout2 = 0;
    void* out = nullptr; int out_cp = 0;             // full-resolution conv2 (optional)
    void* pool_out = nullptr; int pool_cp = 0;
    void* plog = nullptr;              // optional float4 [n][h][w]: conv2's share of the logits (fcn_skip)
    const float* skip_lw = nullptr;    // device [32][4]
};
bool conv12_fused_supported(int k1, int cout1, int k2, int cin2, int cout2);
size_t conv12_weight_image1(const float* w32 /*[25][1][20]*/, int precision, std::vector<uint16_t>& out);
size_t conv12_weight_image2(const float* w32 /*[25][20][cout]*/, int cout, int precision, std::vector<uint16_t>& out);
int launch_conv12_fused(pcs_ctx* ctx, const Conv12Args& a);

// conv1_umma.cu  (first FCN layer on the tensor cores)
size_t conv1_umma_weight_image(const float* w32 /*[ksz*ksz][1][cout]*/, int ksz, int cout, int precision, std::vector<uint16_t>& out,
                               bool pairx = false);
bool conv1_umma_supported(int ksz, int cout);
int launch_conv1_umma(pcs_ctx* ctx, const uint8_t* d_image, int n, int img_h, int img_w, int h, int w, const void* wimg,
                      const float* h_bias, int ksz, int cout, void* out, int out_cp, void* pair_out = nullptr);

}  // namespace pcs
