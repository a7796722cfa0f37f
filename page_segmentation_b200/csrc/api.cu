// C ABI of pcseg_b200 (see include/pcseg_b200.h): context, model upload, the
// per-architecture forward schedules and the page-batch pipeline.
#include "common.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>

namespace pcs {

int set_err(pcs_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

int arena_reserve(pcs_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->arena_bytes) return PCS_OK;
    if (ctx->arena_used != 0 && ctx->arena) {
        // growing while allocations are live would invalidate them
        return set_err(ctx, PCS_ERR_STATE, "arena grow requested with live allocations (%zu > %zu)", bytes, ctx->arena_bytes);
    }
    cudaStreamSynchronize(ctx->stream);
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    const size_t want = bytes + bytes / 16 + (1u << 20);
    if (cudaMalloc(&ctx->arena, want) != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, PCS_ERR_NOMEM, "cudaMalloc of %zu bytes for the activation arena failed", want);
    }
    ctx->arena_bytes = want;
    return PCS_OK;
}

void* arena_alloc(pcs_ctx* ctx, size_t bytes) {
    const size_t off = (ctx->arena_used + 255) / 256 * 256;
    if (off + bytes > ctx->arena_bytes) return nullptr;
    ctx->arena_used = off + bytes;
    return ctx->arena + off;
}

int scratch_reserve(pcs_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return PCS_OK;
    cudaStreamSynchronize(ctx->stream);
    // only ctx->scratch: scratch2 (the max_width pass, which runs a whole first pass - and therefore this function -
    // while it holds pointers into scratch2) has its own grow path in launch_preprocess_max_width
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, PCS_ERR_NOMEM, "cudaMalloc of %zu bytes of scratch failed", want);
    }
    ctx->scratch_bytes = want;
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// layer tables (Keras creation order; ocr4all_pixel_classifier/lib/model.py)
// ---------------------------------------------------------------------------
struct LayerSpec { const char* name; int kind; int k; int cin; int cout; int relu; };
enum { K_CONV = 0, K_DECONV = 1, K_DECONV_S2 = 2, K_LOGITS = 3 };

static const LayerSpec kFcnSkip[] = {   // model.py:45-92
    {"conv1", K_CONV, 5, 1, 20, 1},      {"conv2", K_CONV, 5, 20, 30, 0},    {"conv3", K_CONV, 5, 30, 40, 1},
    {"conv4", K_CONV, 5, 40, 40, 0},     {"conv5", K_CONV, 5, 40, 60, 1},    {"conv6", K_CONV, 5, 60, 60, 0},
    {"conv7", K_CONV, 5, 60, 80, 1},     {"deconv1", K_DECONV, 5, 80, 80, 1}, {"deconv2", K_DECONV_S2, 2, 80, 60, 1},
    {"deconv3", K_DECONV, 5, 120, 40, 1}, {"deconv4", K_DECONV_S2, 2, 100, 30, 1},
    {"deconv5", K_DECONV_S2, 2, 70, 20, 0}, {"logits", K_LOGITS, 1, 50, -1, 0}};
static const LayerSpec kFcn[] = {       // model.py:206-234
    {"conv1", K_CONV, 5, 1, 20, 1},      {"conv2", K_CONV, 5, 20, 30, 0},    {"conv3", K_CONV, 5, 30, 40, 1},
    {"conv4", K_CONV, 5, 40, 40, 0},     {"conv5", K_CONV, 5, 40, 60, 1},    {"conv6", K_CONV, 5, 60, 60, 0},
    {"conv7", K_CONV, 5, 60, 80, 1},     {"deconv1", K_DECONV, 5, 80, 80, 1}, {"deconv2", K_DECONV_S2, 2, 80, 60, 1},
    {"deconv3", K_DECONV, 5, 60, 40, 1}, {"deconv4", K_DECONV_S2, 2, 40, 30, 1},
    {"deconv5", K_DECONV_S2, 2, 30, 20, 0}, {"logits", K_LOGITS, 1, 20, -1, 0}};
static const LayerSpec kUnet[] = {      // model.py:151-203
    {"conv1a", K_CONV, 3, 1, 64, 1},     {"conv1b", K_CONV, 3, 64, 64, 1},   {"conv2a", K_CONV, 3, 64, 128, 1},
    {"conv2b", K_CONV, 3, 128, 128, 1},  {"conv3a", K_CONV, 3, 128, 256, 1}, {"conv3b", K_CONV, 3, 256, 256, 1},
    {"conv4a", K_CONV, 3, 256, 512, 1},  {"conv4b", K_CONV, 3, 512, 512, 1}, {"conv5a", K_CONV, 3, 512, 1024, 1},
    {"conv5b", K_CONV, 3, 1024, 1024, 1}, {"up6", K_CONV, 2, 1024, 512, 1},  {"conv6a", K_CONV, 3, 1024, 512, 1},
    {"conv6b", K_CONV, 3, 512, 512, 1},  {"up7", K_CONV, 2, 512, 256, 1},    {"conv7a", K_CONV, 3, 512, 256, 1},
    {"conv7b", K_CONV, 3, 256, 256, 1},  {"up8", K_CONV, 2, 256, 128, 1},    {"conv8a", K_CONV, 3, 256, 128, 1},
    {"conv8b", K_CONV, 3, 128, 128, 1},  {"up9", K_CONV, 2, 128, 64, 1},     {"conv9a", K_CONV, 3, 128, 64, 1},
    {"conv9b", K_CONV, 3, 64, 64, 1},    {"logits", K_LOGITS, 1, 64, -1, 0}};

static void arch_table(int arch, const LayerSpec** t, int* n) {
    switch (arch) {
        case PCS_ARCH_FCN_SKIP: *t = kFcnSkip; *n = (int)(sizeof(kFcnSkip) / sizeof(LayerSpec)); break;
        case PCS_ARCH_FCN: *t = kFcn; *n = (int)(sizeof(kFcn) / sizeof(LayerSpec)); break;
        case PCS_ARCH_UNET: *t = kUnet; *n = (int)(sizeof(kUnet) / sizeof(LayerSpec)); break;
        default: *t = nullptr; *n = 0;
    }
}

static void free_layers(pcs_ctx* ctx) {
    for (auto& l : ctx->layers) {
        if (l.d_w32) cudaFree(l.d_w32);
        if (l.d_b32) cudaFree(l.d_b32);
        if (l.d_wmma) cudaFree(l.d_wmma);
        for (auto& f : l.fold) if (f.d_w) cudaFree(f.d_w);
        if (l.d_w12) cudaFree(l.d_w12);
        if (l.d_wmma_px) cudaFree(l.d_wmma_px);
        for (auto& part : l.fold) if (part.d_w_px) cudaFree(part.d_w_px);
        if (l.d_head_lw) cudaFree(l.d_head_lw);
        if (l.d_head_lb) cudaFree(l.d_head_lb);
    }
    ctx->layers.clear();
    ctx->model_ready = false;
}

static Layer* find_layer(pcs_ctx* ctx, const char* name) {
    for (auto& l : ctx->layers)
        if (l.name == name) return &l;
    return nullptr;
}

// ---------------------------------------------------------------------------
// stage timing
// ---------------------------------------------------------------------------
struct StageScope {
    pcs_ctx* ctx; size_t idx; bool on;
    StageScope(pcs_ctx* c, const char* name) : ctx(c), idx(0), on(c->timing_enabled) {
        if (!on) return;
        StageTime st; st.name = name;
        cudaEventCreate(&st.e0); cudaEventCreate(&st.e1);
        cudaEventRecord(st.e0, ctx->stream);
        ctx->stage_times.push_back(st);
        idx = ctx->stage_times.size() - 1;
    }
    ~StageScope() { if (on) cudaEventRecord(ctx->stage_times[idx].e1, ctx->stream); }
};

static void clear_stage_times(pcs_ctx* ctx) {
    for (auto& s : ctx->stage_times) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
    ctx->stage_times.clear();
}

// ---------------------------------------------------------------------------
// forward schedules
// ---------------------------------------------------------------------------
// channel padding: whole 8-channel planes for the tensor engine (its kernels read planes through TMA and
// take an odd plane count in their stride), pairs of planes for the CUDA-core kernels (16-channel tiles)
static int new_act(pcs_ctx* ctx, const char* name, int n, int h, int w, int c, Act* out) {
    Act a; a.n = n; a.h = h; a.w = w; a.c = c; a.cp = ctx->engine == PCS_ENGINE_UMMA ? pad8(c) : pad16(c);
    a.p = arena_alloc(ctx, a.bytes());
    if (!a.p) return set_err(ctx, PCS_ERR_NOMEM, "activation arena exhausted at %s", name);
    ctx->acts[name] = a;
    *out = a;
    return PCS_OK;
}

static ConvSrc src_of(const Act& a) { ConvSrc s; s.p = a.p; s.c = a.c; s.cp = a.cp; return s; }

// conv / deconv-s1 layer ('same', stride 1) on 1 or 2 concatenated sources
static bool fold_disabled() {
    static const bool no_fold = getenv("PCSEG_FOLD") && !strcmp(getenv("PCSEG_FOLD"), "0");
    return no_fold;
}

static int run_conv(pcs_ctx* ctx, const char* lname, const Act* s0, const Act* s1, Act* out, Act* pool_out,
                    bool upsample = false, void* plog = nullptr, const float* skip_lw = nullptr, const void* pair_src = nullptr) {
    Layer* L = find_layer(ctx, lname);
    if (!L) return set_err(ctx, PCS_ERR_STATE, "layer %s missing", lname);
    StageScope ts(ctx, lname);
    const Act& any = out ? *out : *pool_out;
    const int n = any.n;
    const int h = out ? out->h : pool_out->h * 2, w = out ? out->w : pool_out->w * 2;
    if (ctx->engine == PCS_ENGINE_UMMA && !upsample && !L->fold.empty() && !fold_disabled() && (h % 4) == 0) {
        void* psum = nullptr;
        for (const Layer::FoldPart& part : L->fold) {
            const Act* src = part.src == 0 ? s0 : s1;
            if (!src) return set_err(ctx, PCS_ERR_STATE, "layer %s: folded part reads a missing source", lname);
            if (part.psum && !psum) {
                psum = arena_alloc(ctx, (size_t)n * h * w * part.npad * sizeof(float));
                if (!psum) return set_err(ctx, PCS_ERR_NOMEM, "activation arena exhausted at the partial sums of %s", lname);
            }
            FoldConvArgs f;
            f.src = src_of(*src);
            f.n = n; f.h = h; f.w = w; f.k = L->k;
            f.wimg = part.d_w; f.h_bias = L->h_b32.data() + part.o0; f.cout = part.ncols; f.npad = part.npad;
            f.nplanes = src->cp / 8; f.relu = L->relu; f.o0 = part.o0;
            if (part.both) {
                if (!s1) return set_err(ctx, PCS_ERR_STATE, "layer %s: folded part reads a missing second source", lname);
                f.src2 = src_of(*s1); f.nplanes += s1->cp / 8;
            }
            if (pair_src) {
                if (!part.d_w_px) return set_err(ctx, PCS_ERR_STATE, "layer %s: no weight image for a pixel-pair source", lname);
                f.pair_src = pair_src; f.wimg = part.d_w_px; f.nplanes += 1;
            }
            f.out = out ? out->p : nullptr; f.out_cp = out ? out->cp : 0;
            f.pool_out = pool_out ? pool_out->p : nullptr; f.pool_cp = pool_out ? pool_out->cp : 0;
            f.psum_out = part.psum == 1 ? psum : nullptr;
            f.psum_in = part.psum == 2 ? psum : nullptr;
            f.plog = plog; f.skip_lw = skip_lw;
            PCS_TRY(launch_conv_fold(ctx, f));
        }
        return PCS_OK;
    }
    if (plog) return set_err(ctx, PCS_ERR_STATE, "layer %s: partial logits requested off the folded kernel", lname);
    if (pair_src) return set_err(ctx, PCS_ERR_STATE, "layer %s: pixel-pair source off the folded kernel", lname);
    if (ctx->engine == PCS_ENGINE_UMMA && upsample && L->d_wmma && L->k == 2 && !s1 && out) {
        // UpSampling2D(2) + Conv2D(2x2) as a 2x2 convolution on the low-resolution grid, N = (parity, C_out)
        UmmaConvArgs a;
        a.src[0] = src_of(*s0); a.nsrc = 1;
        a.n = n; a.h = s0->h; a.w = s0->w; a.k = 2; a.pad = 0;
        a.wmma = L->d_wmma; a.b32 = L->d_b32; a.cout = L->cout; a.npad = L->npad; a.nchunks = L->nchunks; a.relu = L->relu;
        a.mode = 1; a.co_t = L->co_t;
        a.out = out->p; a.out_cp = out->cp;
        return launch_conv_umma(ctx, a);
    }
    if (ctx->engine == PCS_ENGINE_UMMA && !upsample && L->d_wmma && umma_supported(L->k, L->npad)) {
        UmmaConvArgs a;
        a.src[0] = src_of(*s0); a.nsrc = 1;
        if (s1) { a.src[1] = src_of(*s1); a.nsrc = 2; }
        a.n = n; a.h = h; a.w = w; a.k = L->k; a.pad = (L->k % 2) ? L->k / 2 : 0;
        a.wmma = L->d_wmma; a.b32 = L->d_b32; a.cout = L->cout; a.npad = L->npad; a.nchunks = L->nchunks; a.relu = L->relu;
        a.out = out ? out->p : nullptr; a.out_cp = out ? out->cp : 0;
        a.pool_out = pool_out ? pool_out->p : nullptr; a.pool_cp = pool_out ? pool_out->cp : 0;
        return launch_conv_umma(ctx, a);
    }
    DirectConvArgs a;
    a.src[0] = src_of(*s0); a.nsrc = 1;
    if (s1) { a.src[1] = src_of(*s1); a.nsrc = 2; }
    a.upsample = upsample ? 1 : 0;
    a.n = n; a.h = h; a.w = w; a.k = L->k; a.pad = (L->k % 2) ? L->k / 2 : 0;
    a.w32 = L->d_w32; a.b32 = L->d_b32; a.cin = L->cin; a.cout = L->cout; a.relu = L->relu;
    a.out = out ? out->p : nullptr; a.out_cp = out ? out->cp : 0;
    a.pool_out = pool_out ? pool_out->p : nullptr; a.pool_cp = pool_out ? pool_out->cp : 0;
    return launch_conv_direct(ctx, a);
}

static int run_conv1_u8(pcs_ctx* ctx, const char* lname, const uint8_t* d_image, int n, int hs, int ws, Act* out, void* pair_out = nullptr) {
    Layer* L = find_layer(ctx, lname);
    if (!L) return set_err(ctx, PCS_ERR_STATE, "layer %s missing", lname);
    StageScope ts(ctx, lname);
    if (pair_out)
        return launch_conv1_umma(ctx, d_image, n, hs, ws, out->h, out->w, L->d_wmma_px, L->h_b32.data(), L->k, L->cout, out->p, out->cp, pair_out);
    if (ctx->engine == PCS_ENGINE_UMMA && L->d_wmma && conv1_umma_supported(L->k, L->cout))
        return launch_conv1_umma(ctx, d_image, n, hs, ws, out->h, out->w, L->d_wmma, L->h_b32.data(), L->k, L->cout, out->p, out->cp);
    DirectConvArgs a;
    a.src[0].p = d_image; a.src[0].c = 1; a.src[0].cp = 1; a.nsrc = 1;
    a.src_u8 = 1; a.img_h = hs; a.img_w = ws;
    a.n = n; a.h = out->h; a.w = out->w; a.k = L->k; a.pad = L->k / 2;
    a.w32 = L->d_w32; a.b32 = L->d_b32; a.cin = 1; a.cout = L->cout; a.relu = L->relu;
    a.out = out->p; a.out_cp = out->cp;
    return launch_conv_direct(ctx, a);
}

static int run_deconv_s2(pcs_ctx* ctx, const char* lname, const Act* s0, const Act* s1, Act* out) {
    Layer* L = find_layer(ctx, lname);
    if (!L) return set_err(ctx, PCS_ERR_STATE, "layer %s missing", lname);
    StageScope ts(ctx, lname);
    if (ctx->engine == PCS_ENGINE_UMMA && L->d_wmma) {
        UmmaConvArgs a;
        a.src[0] = src_of(*s0); a.nsrc = 1;
        if (s1) { a.src[1] = src_of(*s1); a.nsrc = 2; }
        a.n = s0->n; a.h = s0->h; a.w = s0->w; a.k = 1; a.pad = 0;
        a.wmma = L->d_wmma; a.b32 = L->d_b32; a.cout = L->cout; a.npad = L->npad; a.nchunks = L->nchunks; a.relu = L->relu;
        a.mode = 1; a.co_t = L->co_t;
        a.out = out->p; a.out_cp = out->cp;
        return launch_conv_umma(ctx, a);
    }
    DeconvS2Args a;
    a.src[0] = src_of(*s0); a.nsrc = 1;
    if (s1) { a.src[1] = src_of(*s1); a.nsrc = 2; }
    a.n = s0->n; a.h = s0->h; a.w = s0->w;
    a.w32 = L->d_w32; a.b32 = L->d_b32; a.cin = L->cin; a.cout = L->cout; a.relu = L->relu;
    a.out = out->p; a.out_cp = out->cp;
    return launch_deconv_s2_direct(ctx, a);
}

struct HeadIO {
    const uint8_t* binary; uint8_t* labels; float* logits; float* prob;
    const uint8_t* d_lut; uint8_t* color; uint8_t* overlay; uint8_t* inverted;
};

static int forward_fcn(pcs_ctx* ctx, bool skip, const uint8_t* d_image, int n, int hs, int ws, const HeadIO& io) {
    const int hp = hs + (32 - hs % 32) % 32, wp = ws + (32 - ws % 32) % 32;    // model.py:10-26
    Act conv1, conv2, pool2, conv3, pool4, conv5, conv6, pool6, conv7, d1, d2, d3, d4;
    // fused head (<= 4 classes, tensor engine): conv2 hands its share of the logits to the head as fp32
    // partial sums and its full-resolution tensor is not stored at all
    Layer* L2 = find_layer(ctx, "conv2");
    Layer* L5 = find_layer(ctx, "deconv5");
    Layer* LL = find_layer(ctx, "logits");
    const bool conv2_folds = L2 && L2->fold.size() == 1 && !fold_disabled();
    const bool fused_head = ctx->engine == PCS_ENGINE_UMMA && L5->d_wmma && LL->d_head_lb && (!skip || (conv2_folds && LL->d_head_lw));
    void* plog = nullptr;
    if (fused_head && skip) {
        plog = arena_alloc(ctx, (size_t)n * hp * wp * sizeof(float4));
        if (!plog) return set_err(ctx, PCS_ERR_NOMEM, "activation arena exhausted at the partial logits");
    }
    const bool store_conv2 = (skip && !fused_head) || ctx->keep_acts;
    if (store_conv2) PCS_TRY(new_act(ctx, "conv2", n, hp, wp, 30, &conv2));
    PCS_TRY(new_act(ctx, "pool2", n, hp / 2, wp / 2, 30, &pool2));
    // conv1 never leaves the SM when the fused kernel applies (tensor engine, folded conv2, activations not kept for inspection)
    // Opt-in (PCSEG_FUSE12=1): measured on B200 the fused kernel takes as long as conv1 + conv2 apart (2.8 ms per 64 A4 pages) --
    // the pair is bound by the issue slots of its epilogues, not by the 94 MB per page of HBM traffic the fusion removes
    // (DESIGN.md section 3.2a).
    static const bool no_fuse12 = !(getenv("PCSEG_FUSE12") && !strcmp(getenv("PCSEG_FUSE12"), "1"));
    static const bool pairplane = !(getenv("PCSEG_PAIRPLANE") && !strcmp(getenv("PCSEG_PAIRPLANE"), "0"));
    Layer* L1 = find_layer(ctx, "conv1");
    if (ctx->engine == PCS_ENGINE_UMMA && !no_fuse12 && !ctx->keep_acts && conv2_folds && L1 && L1->d_w12 && L2->d_w12 && (hp % 4) == 0) {
        StageScope ts(ctx, "conv1+conv2");
        Conv12Args f;
        f.d_image = d_image; f.n = n; f.img_h = hs; f.img_w = ws; f.h = hp; f.w = wp;
        f.w1img = L1->d_w12; f.w2img = L2->d_w12; f.h_bias1 = L1->h_b32.data(); f.h_bias2 = L2->h_b32.data(); f.cout2 = L2->cout;
        f.out = store_conv2 ? conv2.p : nullptr; f.out_cp = store_conv2 ? conv2.cp : 0;
        f.pool_out = pool2.p; f.pool_cp = pool2.cp;
        f.plog = plog; f.skip_lw = plog ? LL->d_head_lw : nullptr;
        PCS_TRY(launch_conv12_fused(ctx, f));
    } else if (pairplane && ctx->engine == PCS_ENGINE_UMMA && !ctx->keep_acts && conv2_folds && L1 && L1->d_wmma_px && L2->fold[0].d_w_px &&
               (hp % 4) == 0) {
        // conv1 hands channels 0..15 over as two whole planes and channels 16..19 as pixel-pair units (w + 1 per row): conv2
        // spends 7 instead of 8 MMAs per input row; the inspection path (keep_acts) keeps three planes
        Act c1; c1.n = n; c1.h = hp; c1.w = wp; c1.c = 16; c1.cp = 16;
        c1.p = arena_alloc(ctx, c1.bytes());
        void* pairs = arena_alloc(ctx, (size_t)n * hp * (wp + 1) * 16);
        if (!c1.p || !pairs) return set_err(ctx, PCS_ERR_NOMEM, "activation arena exhausted at conv1");
        PCS_TRY(run_conv1_u8(ctx, "conv1", d_image, n, hs, ws, &c1, pairs));
        PCS_TRY(run_conv(ctx, "conv2", &c1, nullptr, store_conv2 ? &conv2 : nullptr, &pool2, false, plog, plog ? LL->d_head_lw : nullptr, pairs));
    } else {
        PCS_TRY(new_act(ctx, "conv1", n, hp, wp, 20, &conv1));
        PCS_TRY(run_conv1_u8(ctx, "conv1", d_image, n, hs, ws, &conv1));
        PCS_TRY(run_conv(ctx, "conv2", &conv1, nullptr, store_conv2 ? &conv2 : nullptr, &pool2, false, plog, plog ? LL->d_head_lw : nullptr));
    }
    PCS_TRY(new_act(ctx, "conv3", n, hp / 2, wp / 2, 40, &conv3));
    PCS_TRY(run_conv(ctx, "conv3", &pool2, nullptr, &conv3, nullptr));
    PCS_TRY(new_act(ctx, "pool4", n, hp / 4, wp / 4, 40, &pool4));
    PCS_TRY(run_conv(ctx, "conv4", &conv3, nullptr, nullptr, &pool4));      // conv4 itself is never re-read
    PCS_TRY(new_act(ctx, "conv5", n, hp / 4, wp / 4, 60, &conv5));
    PCS_TRY(run_conv(ctx, "conv5", &pool4, nullptr, &conv5, nullptr));
    PCS_TRY(new_act(ctx, "pool6", n, hp / 8, wp / 8, 60, &pool6));
    if (skip) {
        PCS_TRY(new_act(ctx, "conv6", n, hp / 4, wp / 4, 60, &conv6));
        PCS_TRY(run_conv(ctx, "conv6", &conv5, nullptr, &conv6, &pool6));
    } else {
        PCS_TRY(run_conv(ctx, "conv6", &conv5, nullptr, nullptr, &pool6));
    }
    PCS_TRY(new_act(ctx, "conv7", n, hp / 8, wp / 8, 80, &conv7));
    PCS_TRY(run_conv(ctx, "conv7", &pool6, nullptr, &conv7, nullptr));
    PCS_TRY(new_act(ctx, "deconv1", n, hp / 8, wp / 8, 80, &d1));
    PCS_TRY(run_conv(ctx, "deconv1", &conv7, nullptr, &d1, nullptr));
    PCS_TRY(new_act(ctx, "deconv2", n, hp / 4, wp / 4, 60, &d2));
    PCS_TRY(run_deconv_s2(ctx, "deconv2", &d1, nullptr, &d2));
    PCS_TRY(new_act(ctx, "deconv3", n, hp / 4, wp / 4, 40, &d3));
    PCS_TRY(run_conv(ctx, "deconv3", &d2, skip ? &conv6 : nullptr, &d3, nullptr));     // concat [deconv2, conv6]
    PCS_TRY(new_act(ctx, "deconv4", n, hp / 2, wp / 2, 30, &d4));
    PCS_TRY(run_deconv_s2(ctx, "deconv4", &d3, skip ? &conv5 : nullptr, &d4));         // concat [deconv3, conv5]

    StageScope ts(ctx, "head");
    if (fused_head) {
        UmmaHeadArgs hd;
        hd.plog = plog; hd.lb_folded = LL->d_head_lb;
        hd.n_classes = ctx->n_classes; hd.hs = hs; hd.ws = ws;
        // the class map is always produced (the colour pass reads it); colour masks follow as one vectorised kernel
        uint8_t* labels = io.labels;
        const bool want_masks = io.color || io.overlay || io.inverted;
        if (!labels && want_masks) {
            labels = reinterpret_cast<uint8_t*>(arena_alloc(ctx, (size_t)n * hs * ws));
            if (!labels) return set_err(ctx, PCS_ERR_NOMEM, "activation arena exhausted at labels");
        }
        hd.labels = labels; hd.logits = io.logits; hd.prob = io.prob;
        UmmaConvArgs u;
        u.src[0] = src_of(d4); u.nsrc = 1;
        if (skip) { u.src[1] = src_of(conv3); u.nsrc = 2; }
        u.n = n; u.h = hp / 2; u.w = wp / 2; u.k = 1; u.pad = 0;
        u.wmma = L5->d_wmma; u.b32 = L5->d_b32; u.cout = L5->cout; u.npad = L5->npad; u.nchunks = L5->nchunks; u.relu = 0;
        u.mode = 2; u.co_t = 0; u.head = &hd;
        PCS_TRY(launch_conv_umma(ctx, u));
        if (want_masks)
            return launch_masks(ctx, labels, io.binary, n, hs, ws, io.d_lut, ctx->n_classes, io.color, io.overlay, io.inverted);
        return PCS_OK;
    }
    HeadArgs a;
    a.has_deconv = 1;
    a.dsrc[0] = src_of(d4); a.dnsrc = 1;
    if (skip) { a.dsrc[1] = src_of(conv3); a.dnsrc = 2; }                              // concat [deconv4, conv3]
    a.dw32 = L5->d_w32; a.db32 = L5->d_b32; a.dcin = L5->cin; a.dcout = L5->cout;
    if (skip) { a.skip = src_of(conv2); a.has_skip = 1; }                              // concat [deconv5, conv2]
    a.lw32 = LL->d_w32; a.lb32 = LL->d_b32; a.n_classes = ctx->n_classes;
    a.n = n; a.hp = hp; a.wp = wp; a.h = hs; a.w = ws;
    a.binary = io.binary; a.labels = io.labels; a.logits = io.logits; a.prob = io.prob;
    a.lut = io.d_lut; a.color = io.color; a.overlay = io.overlay; a.inverted = io.inverted;
    return launch_head(ctx, a);
}

static int forward_unet(pcs_ctx* ctx, const uint8_t* d_image, int n, int hs, int ws, const HeadIO& io) {
    const int hp = hs + (32 - hs % 32) % 32, wp = ws + (32 - ws % 32) % 32;
    Act c1a, c1, p1, c2a, c2, p2, c3a, c3, p3, c4a, c4, p4, c5a, c5, u6, c6a, c6, u7, c7a, c7, u8, c8a, c8, u9, c9a, c9;
    PCS_TRY(new_act(ctx, "conv1a", n, hp, wp, 64, &c1a));
    PCS_TRY(run_conv1_u8(ctx, "conv1a", d_image, n, hs, ws, &c1a));
    PCS_TRY(new_act(ctx, "conv1b", n, hp, wp, 64, &c1));
    PCS_TRY(new_act(ctx, "pool1", n, hp / 2, wp / 2, 64, &p1));
    PCS_TRY(run_conv(ctx, "conv1b", &c1a, nullptr, &c1, &p1));
    PCS_TRY(new_act(ctx, "conv2a", n, hp / 2, wp / 2, 128, &c2a));
    PCS_TRY(run_conv(ctx, "conv2a", &p1, nullptr, &c2a, nullptr));
    PCS_TRY(new_act(ctx, "conv2b", n, hp / 2, wp / 2, 128, &c2));
    PCS_TRY(new_act(ctx, "pool2", n, hp / 4, wp / 4, 128, &p2));
    PCS_TRY(run_conv(ctx, "conv2b", &c2a, nullptr, &c2, &p2));
    PCS_TRY(new_act(ctx, "conv3a", n, hp / 4, wp / 4, 256, &c3a));
    PCS_TRY(run_conv(ctx, "conv3a", &p2, nullptr, &c3a, nullptr));
    PCS_TRY(new_act(ctx, "conv3b", n, hp / 4, wp / 4, 256, &c3));
    PCS_TRY(new_act(ctx, "pool3", n, hp / 8, wp / 8, 256, &p3));
    PCS_TRY(run_conv(ctx, "conv3b", &c3a, nullptr, &c3, &p3));
    PCS_TRY(new_act(ctx, "conv4a", n, hp / 8, wp / 8, 512, &c4a));
    PCS_TRY(run_conv(ctx, "conv4a", &p3, nullptr, &c4a, nullptr));
    PCS_TRY(new_act(ctx, "conv4b", n, hp / 8, wp / 8, 512, &c4));                     // drop4 = identity at inference
    PCS_TRY(new_act(ctx, "pool4", n, hp / 16, wp / 16, 512, &p4));
    PCS_TRY(run_conv(ctx, "conv4b", &c4a, nullptr, &c4, &p4));
    PCS_TRY(new_act(ctx, "conv5a", n, hp / 16, wp / 16, 1024, &c5a));
    PCS_TRY(run_conv(ctx, "conv5a", &p4, nullptr, &c5a, nullptr));
    PCS_TRY(new_act(ctx, "conv5b", n, hp / 16, wp / 16, 1024, &c5));
    PCS_TRY(run_conv(ctx, "conv5b", &c5a, nullptr, &c5, nullptr));
    PCS_TRY(new_act(ctx, "up6", n, hp / 8, wp / 8, 512, &u6));
    PCS_TRY(run_conv(ctx, "up6", &c5, nullptr, &u6, nullptr, true));                 // UpSampling2D fused
    PCS_TRY(new_act(ctx, "conv6a", n, hp / 8, wp / 8, 512, &c6a));
    PCS_TRY(run_conv(ctx, "conv6a", &c4, &u6, &c6a, nullptr));                       // concat [drop4, up6]
    PCS_TRY(new_act(ctx, "conv6b", n, hp / 8, wp / 8, 512, &c6));
    PCS_TRY(run_conv(ctx, "conv6b", &c6a, nullptr, &c6, nullptr));
    PCS_TRY(new_act(ctx, "up7", n, hp / 4, wp / 4, 256, &u7));
    PCS_TRY(run_conv(ctx, "up7", &c6, nullptr, &u7, nullptr, true));
    PCS_TRY(new_act(ctx, "conv7a", n, hp / 4, wp / 4, 256, &c7a));
    PCS_TRY(run_conv(ctx, "conv7a", &c3, &u7, &c7a, nullptr));
    PCS_TRY(new_act(ctx, "conv7b", n, hp / 4, wp / 4, 256, &c7));
    PCS_TRY(run_conv(ctx, "conv7b", &c7a, nullptr, &c7, nullptr));
    PCS_TRY(new_act(ctx, "up8", n, hp / 2, wp / 2, 128, &u8));
    PCS_TRY(run_conv(ctx, "up8", &c7, nullptr, &u8, nullptr, true));
    PCS_TRY(new_act(ctx, "conv8a", n, hp / 2, wp / 2, 128, &c8a));
    PCS_TRY(run_conv(ctx, "conv8a", &c2, &u8, &c8a, nullptr));
    PCS_TRY(new_act(ctx, "conv8b", n, hp / 2, wp / 2, 128, &c8));
    PCS_TRY(run_conv(ctx, "conv8b", &c8a, nullptr, &c8, nullptr));
    PCS_TRY(new_act(ctx, "up9", n, hp, wp, 64, &u9));
    PCS_TRY(run_conv(ctx, "up9", &c8, nullptr, &u9, nullptr, true));
    PCS_TRY(new_act(ctx, "conv9a", n, hp, wp, 64, &c9a));
    PCS_TRY(run_conv(ctx, "conv9a", &c1, &u9, &c9a, nullptr));
    PCS_TRY(new_act(ctx, "conv9b", n, hp, wp, 64, &c9));
    PCS_TRY(run_conv(ctx, "conv9b", &c9a, nullptr, &c9, nullptr));

    Layer* LL = find_layer(ctx, "logits");
    StageScope ts(ctx, "head");
    HeadArgs a;
    a.has_deconv = 0;
    a.skip = src_of(c9); a.has_skip = 1;
    a.lw32 = LL->d_w32; a.lb32 = LL->d_b32; a.n_classes = ctx->n_classes;
    a.n = n; a.hp = hp; a.wp = wp; a.h = hs; a.w = ws;
    a.binary = io.binary; a.labels = io.labels; a.logits = io.logits; a.prob = io.prob;
    a.lut = io.d_lut; a.color = io.color; a.overlay = io.overlay; a.inverted = io.inverted;
    return launch_head(ctx, a);
}

static size_t arena_need(int arch, int n, int hs, int ws) {
    const size_t hp = hs + (32 - hs % 32) % 32, wp = ws + (32 - ws % 32) % 32;
    const size_t px = (size_t)n * hp * wp;
    // sum over activations of (pixels at level) * padded channels * 2 bytes (+ slack)
    double ch;   // channel-pixels in units of full-resolution pixels
    if (arch == PCS_ARCH_UNET)
        ch = 64 * 5 + (64 + 128 * 5) / 4.0 + (128 + 256 * 5) / 16.0 + (256 + 512 * 5) / 64.0 + (512 + 1024 * 2) / 256.0;
    else
        ch = 32 + 32 + 8 + (32 + 48 + 32) / 4.0 + (48 + 64 + 64 + 64 + 48 + 96) / 16.0 + (64 + 80 + 80) / 64.0;
    return (size_t)(px * ch * 2.0) + (size_t)64 * 4096 + (1u << 20);
}

}  // namespace pcs

using namespace pcs;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int pcs_abi_version(void) { return PCS_ABI_VERSION; }

int pcs_ctx_create(int device, pcs_ctx** out) {
    if (!out) return PCS_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return PCS_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PCS_ERR_CUDA;
    if (prop.major != 10) return PCS_ERR_DEVICE;      // sm_100a code only; no fallback by design
    if (cudaSetDevice(device) != cudaSuccess) return PCS_ERR_CUDA;
    pcs_ctx* ctx = new pcs_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("PCSEG_PDL")) ctx->pdl = atoi(e) != 0;
    const char* t = getenv("PCSEG_TIMING");
    ctx->timing_enabled = t && t[0] == '1';
    const char* e = getenv("PCSEG_ENGINE");
    if (e && !strcmp(e, "direct")) ctx->engine = PCS_ENGINE_DIRECT;
    *out = ctx;
    return PCS_OK;
}

void pcs_ctx_destroy(pcs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (cudaStream_t cs : ctx->copy_streams)
        if (cs) cudaStreamSynchronize(cs);                 // submitted host calls may still be downloading
    output_writer_destroy(ctx);
    free_layers(ctx);
    clear_stage_times(ctx);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    if (ctx->d_sat_count) cudaFree(ctx->d_sat_count);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->h_png_sizes) cudaFreeHost(ctx->h_png_sizes);
    for (int i = 0; i < 2; ++i)
        if (ctx->copy_streams[i]) cudaStreamDestroy(ctx->copy_streams[i]);
    for (int i = 0; i < pcs_ctx::kHostBufs; ++i)
        if (ctx->ev_h2d[i]) { cudaEventDestroy(ctx->ev_h2d[i]); cudaEventDestroy(ctx->ev_comp[i]); cudaEventDestroy(ctx->ev_d2h[i]); cudaEventDestroy(ctx->ev_sizes[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (cudaEvent_t e : ctx->ev_call)
        if (e) cudaEventDestroy(e);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_aux_fork) cudaEventDestroy(ctx->ev_aux_fork);
    if (ctx->ev_aux_join) cudaEventDestroy(ctx->ev_aux_join);
    delete ctx;
}

const char* pcs_last_error(const pcs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int pcs_set_stream(pcs_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PCS_ERR_ARG;
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return PCS_OK;
}

int pcs_synchronize(pcs_ctx* ctx) {
    if (!ctx) return PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // submitted host calls (pcs_predict_pages_*_submit) do not join the compute stream to their downloads
    if (ctx->call_seq) PCS_CUDA(ctx, cudaEventSynchronize(ctx->ev_call[(ctx->call_seq - 1) % pcs_ctx::kCallRing]));
    return PCS_OK;
}

int64_t pcs_launch_count(const pcs_ctx* ctx) { return ctx ? ctx->launches : -1; }

int pcs_set_engine(pcs_ctx* ctx, int engine) {
    if (!ctx || (engine != PCS_ENGINE_UMMA && engine != PCS_ENGINE_DIRECT)) return PCS_ERR_ARG;
    ctx->engine = engine;
    return PCS_OK;
}

int pcs_model_load(pcs_ctx* ctx, int arch, int n_classes, int precision, const pcs_layer_weights* layers, int n_layers) {
    if (!ctx) return PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    const LayerSpec* table; int nt;
    arch_table(arch, &table, &nt);
    if (!table) return set_err(ctx, PCS_ERR_ARG, "unknown architecture id %d", arch);
    if (n_layers != nt) return set_err(ctx, PCS_ERR_ARG, "architecture %d expects %d weighted layers, got %d", arch, nt, n_layers);
    if (n_classes < 1 || n_classes > kMaxClasses) return set_err(ctx, PCS_ERR_ARG, "n_classes %d out of range 1..%d", n_classes, kMaxClasses);
    if (precision != PCS_PREC_BF16 && precision != PCS_PREC_FP16) return set_err(ctx, PCS_ERR_ARG, "unknown precision %d", precision);
    PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_layers(ctx);
    ctx->arch = arch; ctx->n_classes = n_classes; ctx->precision = precision;
    { static int64_t next_stamp = 0; ctx->model_stamp = ++next_stamp; }
    ctx->sat_pending = true;
    if (ctx->d_sat_count) PCS_CUDA(ctx, cudaMemsetAsync(ctx->d_sat_count, 0, sizeof(unsigned long long), ctx->stream));
    for (int li = 0; li < nt; ++li) {
        const LayerSpec& s = table[li];
        const pcs_layer_weights& w = layers[li];
        const int cout = s.cout < 0 ? n_classes : s.cout;
        const bool transposed = s.kind == K_DECONV || s.kind == K_DECONV_S2;
        const int e2 = transposed ? cout : s.cin, e3 = transposed ? s.cin : cout;
        if (!w.kernel || !w.bias || w.shape[0] != s.k || w.shape[1] != s.k || w.shape[2] != e2 || w.shape[3] != e3) {
            free_layers(ctx);
            return set_err(ctx, PCS_ERR_ARG, "layer %d (%s): kernel shape (%d,%d,%d,%d) != expected (%d,%d,%d,%d)", li, s.name,
                           w.shape[0], w.shape[1], w.shape[2], w.shape[3], s.k, s.k, e2, e3);
        }
        Layer L;
        L.name = s.name; L.kind = s.kind; L.k = s.k; L.cin = s.cin; L.cout = cout; L.relu = s.relu;
        const int taps = s.k * s.k;
        L.h_w32.assign((size_t)taps * s.cin * cout, 0.f);
        for (int i = 0; i < s.k; ++i)
            for (int j = 0; j < s.k; ++j)
                for (int c = 0; c < s.cin; ++c)
                    for (int o = 0; o < cout; ++o) {
                        float v;
                        if (s.kind == K_DECONV) {
                            // gradient-of-conv form -> correlation: K'[i,j,c,o] = K[k-1-i, k-1-j, o, c]
                            v = w.kernel[(((size_t)(s.k - 1 - i) * s.k + (s.k - 1 - j)) * cout + o) * s.cin + c];
                        } else if (s.kind == K_DECONV_S2) {
                            v = w.kernel[(((size_t)i * s.k + j) * cout + o) * s.cin + c];
                        } else {
                            v = w.kernel[(((size_t)i * s.k + j) * s.cin + c) * cout + o];
                        }
                        L.h_w32[(((size_t)i * s.k + j) * s.cin + c) * cout + o] = v;
                    }
        L.h_b32.assign(w.bias, w.bias + cout);
        // every layer except the first and the logits consumes operands rounded to the model precision
        const bool rounded = li != 0 && s.kind != K_LOGITS;
        if (L.name == "deconv5" || (arch == PCS_ARCH_UNET && s.kind == K_CONV && s.k == 2)) L.h_w32_raw = L.h_w32;
        if (rounded) {
            for (float& v : L.h_w32)
                v = precision == PCS_PREC_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
        }
        PCS_CUDA(ctx, cudaMalloc(&L.d_w32, L.h_w32.size() * 4));
        PCS_CUDA(ctx, cudaMalloc(&L.d_b32, L.h_b32.size() * 4));
        PCS_CUDA(ctx, cudaMemcpy(L.d_w32, L.h_w32.data(), L.h_w32.size() * 4, cudaMemcpyHostToDevice));
        PCS_CUDA(ctx, cudaMemcpy(L.d_b32, L.h_b32.data(), L.h_b32.size() * 4, cudaMemcpyHostToDevice));
        ctx->layers.push_back(std::move(L));
    }
    // first layer (C_in = 1; FCN 5x5 -> 20, U-Net 3x3 -> 64): tap-contraction operand image, weights split hi + lo
    if (conv1_umma_supported(ctx->layers[0].k, ctx->layers[0].cout)) {
        Layer& L = ctx->layers[0];
        std::vector<uint16_t> img;
        L.wmma_bytes = conv1_umma_weight_image(L.h_w32.data(), L.k, L.cout, precision, img);
        PCS_CUDA(ctx, cudaMalloc(&L.d_wmma, L.wmma_bytes));
        PCS_CUDA(ctx, cudaMemcpy(L.d_wmma, img.data(), L.wmma_bytes, cudaMemcpyHostToDevice));
        if (L.k == 5 && L.cout == 20) {         // FCN conv1: the image with channels 16..19 of the right neighbour in columns 20..23
            const size_t nb = conv1_umma_weight_image(L.h_w32.data(), L.k, L.cout, precision, img, true);
            PCS_CUDA(ctx, cudaMalloc(&L.d_wmma_px, nb));
            PCS_CUDA(ctx, cudaMemcpy(L.d_wmma_px, img.data(), nb, cudaMemcpyHostToDevice));
        }
    }
    // conv1 + conv2 of the FCN variants as one kernel (conv12_fused.cu)
    if (ctx->layers.size() > 1 && ctx->layers[0].name == "conv1" && ctx->layers[1].name == "conv2" && ctx->layers[1].kind == K_CONV &&
        !ctx->layers[1].relu &&
        conv12_fused_supported(ctx->layers[0].k, ctx->layers[0].cout, ctx->layers[1].k, ctx->layers[1].cin, ctx->layers[1].cout)) {
        Layer &L1 = ctx->layers[0], &L2 = ctx->layers[1];
        std::vector<uint16_t> img;
        size_t nb = conv12_weight_image1(L1.h_w32.data(), precision, img);
        PCS_CUDA(ctx, cudaMalloc(&L1.d_w12, nb));
        PCS_CUDA(ctx, cudaMemcpy(L1.d_w12, img.data(), nb, cudaMemcpyHostToDevice));
        nb = conv12_weight_image2(L2.h_w32.data(), L2.cout, precision, img);
        PCS_CUDA(ctx, cudaMalloc(&L2.d_w12, nb));
        PCS_CUDA(ctx, cudaMemcpy(L2.d_w12, img.data(), nb, cudaMemcpyHostToDevice));
    }
    // tensor-core operand images: stride-1 convolutions with C_in > 1 and the 2x2 stride-2
    // transposed convolutions that are not fused into the head
    for (size_t li = 1; li < ctx->layers.size(); ++li) {
        Layer& L = ctx->layers[li];
        int src_c[2] = {L.cin, 0}, nsrc = 1;
        if (arch == PCS_ARCH_FCN_SKIP) {
            if (L.name == "deconv3") { src_c[0] = 60; src_c[1] = 60; nsrc = 2; }      // [deconv2, conv6]
            if (L.name == "deconv4") { src_c[0] = 40; src_c[1] = 60; nsrc = 2; }      // [deconv3, conv5]
        }
        if (arch == PCS_ARCH_UNET && (L.name == "conv6a" || L.name == "conv7a" || L.name == "conv8a" || L.name == "conv9a")) {
            src_c[0] = L.cin / 2; src_c[1] = L.cin / 2; nsrc = 2;
        }
        std::vector<uint16_t> img;
        if (arch == PCS_ARCH_UNET && L.kind == K_CONV && L.k == 2) {
            // up6..up9: UpSampling2D + 2x2 convolution folded onto the low-resolution grid
            L.co_t = pad16(L.cout);
            L.npad = 128;
            if ((4 * L.co_t) % L.npad) continue;
            L.wmma_bytes = umma_weight_image_up2(L.h_w32_raw.data(), L.cin, L.cout, L.co_t, L.npad, precision, img);
        } else if (L.kind == K_CONV || L.kind == K_DECONV) {
            L.npad = std::min(pad16(L.cout), 128);
            if (!umma_supported(L.k, L.npad) || (L.cout > 128 && L.cout % 128)) continue;
            L.wmma_bytes = umma_weight_image(L.h_w32.data(), L.k * L.k, src_c, nsrc, L.cout, L.npad, precision, img);
            // dy-folded marching kernel: whole layer if its resident weights fit, else split along N (output
            // channels, two launches) or along K (one launch per concatenated source)
            auto add_part = [&](int src, int ci0, int cin, int o0, int ncols, int npad, int psum) -> int {
                std::vector<uint16_t> fimg;
                const size_t fb = fold_weight_image(L.h_w32.data(), L.cin, L.cout, ci0, cin, o0, ncols, npad, precision, fimg, false, L.k);
                Layer::FoldPart part;
                part.src = src; part.o0 = o0; part.ncols = ncols; part.npad = npad; part.nplanes = pad8(cin) / 8; part.psum = psum;
                PCS_CUDA(ctx, cudaMalloc(&part.d_w, fb));
                PCS_CUDA(ctx, cudaMemcpy(part.d_w, fimg.data(), fb, cudaMemcpyHostToDevice));
                if (L.name == "conv2" && cin == 20 && npad == 32 && ci0 == 0 && o0 == 0) {
                    // the same weights for a source whose channels 16..19 arrive as pixel-pair units (7 K steps per row instead of 8)
                    const size_t pb = fold_weight_image(L.h_w32.data(), L.cin, L.cout, ci0, cin, o0, ncols, npad, precision, fimg, true);
                    if (pb) {
                        PCS_CUDA(ctx, cudaMalloc(&part.d_w_px, pb));
                        PCS_CUDA(ctx, cudaMemcpy(part.d_w_px, fimg.data(), pb, cudaMemcpyHostToDevice));
                    }
                }
                L.fold.push_back(part);
                return PCS_OK;
            };
            // C_out of an odd number of planes (40): the folded kernel takes it unpadded
            static const bool fold40 = !(getenv("PCSEG_FOLD40") && !strcmp(getenv("PCSEG_FOLD40"), "0"));
            const int fnp1 = (fold40 && pad8(L.cout) % 16 == 8 && fold_supported(5, pad8(L.cout), pad8(L.cin) / 8)) ? pad8(L.cout) : L.npad;
            const int fnp2 = (fold40 && pad8(L.cout) % 16 == 8 && nsrc == 2 && fold_supported(5, pad8(L.cout), pad8(src_c[0]) / 8) &&
                              fold_supported(5, pad8(L.cout), pad8(src_c[1]) / 8)) ? pad8(L.cout) : L.npad;
            static const bool fold3 = !(getenv("PCSEG_FOLD3") && !strcmp(getenv("PCSEG_FOLD3"), "0"));
            if (fold3 && L.k == 3 && L.kind == K_CONV && L.cout == 64 && ((nsrc == 1 && L.cin == 64) || (nsrc == 2 && src_c[0] == 64 && src_c[1] == 64))) {
                // U-Net conv1b / conv9b / conv9a: 64 output channels at full resolution, dy-folded (N' = 192); the two sources of
                // conv9a's concatenation are read by one launch through two tensor maps
                PCS_TRY(add_part(0, 0, L.cin, 0, 64, 64, 0));
                L.fold.back().both = nsrc == 2;
            } else if (fold3 && L.k == 3 && L.kind == K_CONV && L.cout == 128 && nsrc == 1 && (L.cin == 64 || L.cin == 128)) {
                // conv2a / conv2b / conv8b (half resolution, 128 output channels): two launches of 64 output channels each
                PCS_TRY(add_part(0, 0, L.cin, 0, 64, 64, 0));
                PCS_TRY(add_part(0, 0, L.cin, 64, 64, 64, 0));
            } else if (L.k == 5 && nsrc == 1 && fold_supported(5, fnp1, pad8(L.cin) / 8)) {
                PCS_TRY(add_part(0, 0, L.cin, 0, L.cout, fnp1, 0));
            } else if (fold40 && L.k == 5 && nsrc == 1 && L.cout == 80 && fold_supported(5, 40, pad8(L.cin) / 8)) {
                // conv7 (60 -> 80): two launches of 40 output channels each on the 40-column kernel (N' = 208) instead of the
                // plain kernel, whose N = 80 leaves room for one accumulator stage only (tensor pipe 56 %)
                PCS_TRY(add_part(0, 0, L.cin, 0, 40, 40, 0));
                PCS_TRY(add_part(0, 0, L.cin, 40, 40, 40, 0));
            } else if (L.k == 5 && nsrc == 1 && L.cout > 32 && L.cout <= 64 && fold_supported(5, 32, pad8(L.cin) / 8)) {
                PCS_TRY(add_part(0, 0, L.cin, 0, 32, 32, 0));
                PCS_TRY(add_part(0, 0, L.cin, 32, L.cout - 32, 32, 0));
            } else if (L.k == 5 && nsrc == 2 && fold_supported(5, fnp2, pad8(src_c[0]) / 8) &&
                       fold_supported(5, fnp2, pad8(src_c[1]) / 8)) {
                PCS_TRY(add_part(0, 0, src_c[0], 0, L.cout, fnp2, 1));
                PCS_TRY(add_part(1, src_c[0], src_c[1], 0, L.cout, fnp2, 2));
            }
        } else if (L.kind == K_DECONV_S2 && L.name == "deconv5") {
            // fused head: deconv5 composed with the logits layer on the host (conv_umma.cu EPI_HEAD)
            if (arch == PCS_ARCH_FCN_SKIP) { src_c[0] = 30; src_c[1] = 40; nsrc = 2; }     // [deconv4, conv3]
            if (n_classes > 4) continue;
            const Layer& LG = ctx->layers.back();                                           // logits: [cin][n_classes]
            std::vector<double> m((size_t)4 * L.cin * 4, 0.0);
            for (int t = 0; t < 4; ++t)
                for (int c = 0; c < L.cin; ++c)
                    for (int k = 0; k < n_classes; ++k) {
                        double v = 0.0;
                        for (int o = 0; o < L.cout; ++o)
                            v += (double)L.h_w32_raw[((size_t)t * L.cin + c) * L.cout + o] * (double)LG.h_w32[(size_t)o * n_classes + k];
                        m[((size_t)t * L.cin + c) * 4 + k] = v;
                    }
            L.npad = 32;
            L.wmma_bytes = umma_weight_image_head(m.data(), src_c, nsrc, precision, img);
        } else if (L.kind == K_DECONV_S2) {
            L.co_t = pad16(L.cout);
            L.npad = std::min(4 * L.co_t, 128);
            if (!umma_supported(1, L.npad) || (4 * L.co_t) % L.npad) continue;
            L.wmma_bytes = umma_weight_image_deconv(L.h_w32.data(), src_c, nsrc, L.cout, L.co_t, L.npad, precision, img);
        } else {
            continue;
        }
        L.nchunks = 0;
        for (int s = 0; s < nsrc; ++s) L.nchunks += pad16(src_c[s]) / 16;
        PCS_CUDA(ctx, cudaMalloc(&L.d_wmma, L.wmma_bytes));
        PCS_CUDA(ctx, cudaMemcpy(L.d_wmma, img.data(), L.wmma_bytes, cudaMemcpyHostToDevice));
    }
    if (arch != PCS_ARCH_UNET && n_classes <= 4) {
        // fused head constants: bias lb' = lb + b5 . lw[0:20] (+ b2 . lw[20:50]) and, for fcn_skip, the
        // logits rows of the conv2 skip channels [32][4] that conv2's epilogue applies (conv_fold.cu)
        Layer* L2 = find_layer(ctx, "conv2");
        Layer* L5 = find_layer(ctx, "deconv5");
        Layer* LL = find_layer(ctx, "logits");
        std::vector<float> lw(32 * 4, 0.f), lb(4, 0.f);
        const bool skip = arch == PCS_ARCH_FCN_SKIP;
        for (int k = 0; k < n_classes; ++k) {
            double v = LL->h_b32[k];
            for (int o = 0; o < L5->cout; ++o) v += (double)L5->h_b32[o] * (double)LL->h_w32[(size_t)o * n_classes + k];
            if (skip)
                for (int c = 0; c < L2->cout; ++c) {
                    const float w = LL->h_w32[(size_t)(L5->cout + c) * n_classes + k];
                    lw[(size_t)c * 4 + k] = w;
                    v += (double)L2->h_b32[c] * (double)w;
                }
            lb[k] = (float)v;
        }
        if (skip) {
            PCS_CUDA(ctx, cudaMalloc(&LL->d_head_lw, lw.size() * 4));
            PCS_CUDA(ctx, cudaMemcpy(LL->d_head_lw, lw.data(), lw.size() * 4, cudaMemcpyHostToDevice));
        }
        PCS_CUDA(ctx, cudaMalloc(&LL->d_head_lb, lb.size() * 4));
        PCS_CUDA(ctx, cudaMemcpy(LL->d_head_lb, lb.data(), lb.size() * 4, cudaMemcpyHostToDevice));
    }
    if (getenv("PCSEG_DEBUG_SYNC_LOAD")) PCS_CUDA(ctx, cudaDeviceSynchronize());
    ctx->model_ready = true;
    return PCS_OK;
}

int pcs_preprocess(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int Hs, int Ws,
                   uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary) {
    if (!ctx || !d_bin || (!d_grey && d_image)) return ctx ? set_err(ctx, PCS_ERR_ARG, "preprocess: null input") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->timing_enabled) clear_stage_times(ctx);
    StageScope ts(ctx, "preprocess");
    return launch_preprocess(ctx, d_grey, d_bin, n, H, W, Hs, Ws, d_image, d_binary, d_orig_binary);
}

int pcs_preprocess_bits(pcs_ctx* ctx, const uint32_t* d_bits, size_t words_per_page, int n, int H, int W, int level0, int level1, int Hs, int Ws,
                        uint8_t* d_image, uint8_t* d_binary) {
    if (!ctx || !d_bits) return ctx ? set_err(ctx, PCS_ERR_ARG, "preprocess_bits: null input") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->timing_enabled) clear_stage_times(ctx);
    StageScope ts(ctx, "preprocess");
    return launch_preprocess_bits(ctx, d_bits, words_per_page, n, H, W, level0, level1, Hs, Ws, d_image, d_binary);
}

int pcs_pack_bits(pcs_ctx* ctx, const uint8_t* d_src, int n, size_t n_pixels, uint32_t* d_bits, size_t words_per_page) {
    if (!ctx || !d_src || !d_bits) return ctx ? set_err(ctx, PCS_ERR_ARG, "pack_bits: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_pack_bits(ctx, d_src, n, n_pixels, d_bits, words_per_page);
}

int pcs_unpack_bits(pcs_ctx* ctx, const uint32_t* d_bits, int n, size_t words_per_page, size_t n_pixels, uint8_t* d_dst) {
    if (!ctx || !d_bits || !d_dst) return ctx ? set_err(ctx, PCS_ERR_ARG, "unpack_bits: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_unpack_bits(ctx, d_bits, n, words_per_page, n_pixels, d_dst);
}

int pcs_preprocess_max_width(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin, int n, int H, int W, int H1, int W1,
                             int H2, int W2, uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary) {
    if (!ctx || !d_bin || !d_grey) return ctx ? set_err(ctx, PCS_ERR_ARG, "preprocess: null input") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->timing_enabled) clear_stage_times(ctx);
    StageScope ts(ctx, "preprocess");
    return launch_preprocess_max_width(ctx, d_grey, d_bin, n, H, W, H1, W1, H2, W2, d_image, d_binary, d_orig_binary);
}

int pcs_forward(pcs_ctx* ctx, const uint8_t* d_image, const uint8_t* d_binary, int n, int Hs, int Ws, uint8_t* d_labels,
                float* d_logits, float* d_prob, const uint8_t* lut, uint8_t* d_color, uint8_t* d_overlay,
                uint8_t* d_inverted) {
    if (!ctx) return PCS_ERR_ARG;
    if (!ctx->model_ready) return set_err(ctx, PCS_ERR_STATE, "pcs_forward before pcs_model_load");
    if (!d_image || n <= 0 || Hs <= 0 || Ws <= 0) return set_err(ctx, PCS_ERR_ARG, "forward: bad argument");
    const bool want_masks = d_color || d_overlay || d_inverted;
    if (want_masks && !lut) return set_err(ctx, PCS_ERR_ARG, "forward: colour outputs need a LUT");
    if ((d_overlay || d_inverted) && !d_binary) return set_err(ctx, PCS_ERR_ARG, "forward: overlay outputs need the binary page");
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->timing_enabled && (ctx->stage_times.empty() || ctx->stage_times.back().name != "preprocess"))
        clear_stage_times(ctx);
    ctx->arena_used = 0;
    ctx->acts.clear();
    PCS_TRY(arena_reserve(ctx, arena_need(ctx->arch, n, Hs, Ws)));
    HeadIO io{d_binary, d_labels, d_logits, d_prob, nullptr, d_color, d_overlay, d_inverted};
    if (want_masks) {
        uint8_t* d_lut = reinterpret_cast<uint8_t*>(arena_alloc(ctx, 256 * 3));
        if (!d_lut) return set_err(ctx, PCS_ERR_NOMEM, "arena exhausted (lut)");
        PCS_CUDA(ctx, cudaMemcpyAsync(d_lut, lut, (size_t)ctx->n_classes * 3, cudaMemcpyHostToDevice, ctx->stream));
        io.d_lut = d_lut;
    }
    if (ctx->arch == PCS_ARCH_UNET) PCS_TRY(forward_unet(ctx, d_image, n, Hs, Ws, io));
    else PCS_TRY(forward_fcn(ctx, ctx->arch == PCS_ARCH_FCN_SKIP, d_image, n, Hs, Ws, io));
    if (ctx->precision == PCS_PREC_FP16 && (ctx->sat_mode == 2 || (ctx->sat_mode == 1 && ctx->sat_pending))) {
        // every activation this forward stored (the arena is one contiguous range, untouched parts are whatever an
        // earlier forward stored): saturated fp16 stores show up as +-65504
        if (!ctx->d_sat_count) {
            PCS_CUDA(ctx, cudaMalloc(&ctx->d_sat_count, sizeof(unsigned long long)));
            PCS_CUDA(ctx, cudaMemsetAsync(ctx->d_sat_count, 0, sizeof(unsigned long long), ctx->stream));
        }
        for (const auto& kv : ctx->acts)
            PCS_TRY(launch_saturation_scan(ctx, kv.second.p, kv.second.bytes() / 2, ctx->d_sat_count));
        ctx->sat_pending = false;
    }
    return PCS_OK;
}

int pcs_set_saturation_check(pcs_ctx* ctx, int mode) {
    if (!ctx) return PCS_ERR_ARG;
    if (mode < 0 || mode > 2) return set_err(ctx, PCS_ERR_ARG, "saturation check mode %d (0 never, 1 first forward after a model load, 2 every forward)", mode);
    ctx->sat_mode = mode;
    return PCS_OK;
}

int pcs_saturation_count(pcs_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return ctx ? set_err(ctx, PCS_ERR_ARG, "saturation_count: null argument") : PCS_ERR_ARG;
    *out = 0;
    if (!ctx->d_sat_count) return PCS_OK;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long v = 0;
    PCS_CUDA(ctx, cudaMemcpyAsync(&v, ctx->d_sat_count, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = v;
    return PCS_OK;
}

int pcs_masks(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W, const uint8_t* lut,
              int n_lut, uint8_t* d_color, uint8_t* d_overlay, uint8_t* d_inverted) {
    if (!ctx || !d_labels || !d_binary || !lut) return ctx ? set_err(ctx, PCS_ERR_ARG, "masks: null argument") : PCS_ERR_ARG;
    if (n_lut < 0 || n_lut > 256) return set_err(ctx, PCS_ERR_ARG, "masks: LUT size %d", n_lut);
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    PCS_TRY(scratch_reserve(ctx, 1024));
    uint8_t* d_lut = reinterpret_cast<uint8_t*>(ctx->scratch);
    PCS_CUDA(ctx, cudaMemcpyAsync(d_lut, lut, (size_t)n_lut * 3, cudaMemcpyHostToDevice, ctx->stream));
    StageScope ts(ctx, "masks");
    return launch_masks(ctx, d_labels, d_binary, n, H, W, d_lut, n_lut, d_color, d_overlay, d_inverted);
}

int pcs_resize_nearest(pcs_ctx* ctx, const uint8_t* d_src, int n, int H, int W, uint8_t* d_dst, int Ho, int Wo) {
    if (!ctx || !d_src || !d_dst) return ctx ? set_err(ctx, PCS_ERR_ARG, "resize_nearest: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_resize_nearest(ctx, d_src, n, H, W, d_dst, Ho, Wo);
}

int pcs_ccl(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int32_t* d_labels, int32_t* d_stats,
            int max_components, int32_t* d_ncomp) {
    if (!ctx || !d_img || !d_labels) return ctx ? set_err(ctx, PCS_ERR_ARG, "ccl: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "ccl");
    return launch_ccl(ctx, d_img, n, H, W, d_labels, d_stats, max_components, d_ncomp);
}

int pcs_cc_majority(pcs_ctx* ctx, uint8_t* d_pred, const uint8_t* d_binary, int n, int H, int W, int n_classes) {
    if (!ctx || !d_pred || !d_binary) return ctx ? set_err(ctx, PCS_ERR_ARG, "cc_majority: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "cc_majority");
    return launch_cc_majority(ctx, d_pred, d_binary, n, H, W, n_classes);
}

int pcs_bounding_boxes(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out) {
    if (!ctx || !d_pred || !d_out) return ctx ? set_err(ctx, PCS_ERR_ARG, "bounding_boxes: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "bounding_boxes");
    return launch_bounding_boxes(ctx, d_pred, n, H, W, n_classes, d_out);
}

int pcs_class_components(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, int32_t* d_stats, int max_components,
                         int32_t* d_ncomp) {
    if (!ctx || !d_pred || !d_stats) return ctx ? set_err(ctx, PCS_ERR_ARG, "class_components: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "class_components");
    return launch_class_components(ctx, d_pred, n, H, W, n_classes, d_stats, max_components, d_ncomp);
}

int pcs_char_height(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int inverse, int32_t* d_height) {
    if (!ctx || !d_img || !d_height) return ctx ? set_err(ctx, PCS_ERR_ARG, "char_height: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "char_height");
    return launch_char_height(ctx, d_img, n, H, W, inverse, d_height);
}

int pcs_segment_masks(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, int Ho, int Wo, const uint8_t* colours, int m,
                      uint8_t* d_masks) {
    if (!ctx || !d_rgb || !colours || !d_masks) return ctx ? set_err(ctx, PCS_ERR_ARG, "segment_masks: null argument") : PCS_ERR_ARG;
    if (H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return set_err(ctx, PCS_ERR_ARG, "segment_masks: bad shape");
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "segment_masks");
    return launch_segment_masks(ctx, d_rgb, H, W, Ho, Wo, colours, m, d_masks);
}

int pcs_dilate3x3(pcs_ctx* ctx, const uint8_t* d_src, int H, int W, int C, uint8_t* d_dst) {
    if (!ctx || !d_src || !d_dst) return ctx ? set_err(ctx, PCS_ERR_ARG, "dilate3x3: null argument") : PCS_ERR_ARG;
    if (H <= 0 || W <= 0 || C <= 0 || d_src == d_dst) return set_err(ctx, PCS_ERR_ARG, "dilate3x3: bad shape or in-place call");
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_dilate3x3(ctx, d_src, H, W, C, d_dst);
}

int pcs_integral_image(pcs_ctx* ctx, const uint8_t* d_mask, int n, int H, int W, int32_t* d_sat) {
    if (!ctx || !d_mask || !d_sat) return ctx ? set_err(ctx, PCS_ERR_ARG, "integral_image: null argument") : PCS_ERR_ARG;
    if (n <= 0 || H <= 0 || W <= 0 || (long long)H * W >= (1ll << 31)) return set_err(ctx, PCS_ERR_ARG, "integral_image: bad shape");
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "integral_image");
    return launch_integral_image(ctx, d_mask, n, H, W, d_sat);
}

int pcs_text_regions(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, const uint8_t* colour, int k_close, int k_open,
                     int k_region, uint8_t* d_text_inv, uint8_t* d_region) {
    if (!ctx || !d_rgb || !colour) return ctx ? set_err(ctx, PCS_ERR_ARG, "text_regions: null argument") : PCS_ERR_ARG;
    if (H <= 0 || W <= 0 || H > 65535) return set_err(ctx, PCS_ERR_ARG, "text_regions: bad shape");
    // cv2.getStructuringElement rejects empty elements (char_height < 3 makes int(char_height / 3) zero)
    if (k_close < 1 || k_open < 1 || k_region < 1)
        return set_err(ctx, PCS_ERR_ARG, "text_regions: structuring elements must be at least 1x1 (got %d, %d, %d)", k_close, k_open, k_region);
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "text_regions");
    return launch_text_regions(ctx, d_rgb, H, W, colour, k_close, k_open, k_region, d_text_inv, d_region);
}

// ---- training primitives (train.cu): thin argument checks, everything on the ctx stream
#define PCS_TRAIN_ENTER(ctx, cond, what)                                                        \
    if (!(ctx)) return PCS_ERR_ARG;                                                             \
    if (!(cond)) return set_err((ctx), PCS_ERR_ARG, what ": null or bad argument");             \
    PCS_CUDA((ctx), cudaSetDevice((ctx)->device));

int pcs_train_corr2d(pcs_ctx* ctx, const float* d_x, const float* d_w, const float* d_b, float* d_y, int c_in, int c_out, int H, int W, int k,
                     int relu, int accumulate) {
    PCS_TRAIN_ENTER(ctx, d_x && d_w && d_y, "train_corr2d");
    return train_corr2d(ctx, d_x, d_w, d_b, d_y, c_in, c_out, H, W, k, relu, accumulate);
}
int pcs_train_wgrad(pcs_ctx* ctx, const float* d_x, const float* d_dy, float* d_dw, int c_in, int c_out, int H, int W, int k) {
    PCS_TRAIN_ENTER(ctx, d_x && d_dy && d_dw, "train_wgrad");
    return train_wgrad(ctx, d_x, d_dy, d_dw, c_in, c_out, H, W, k);
}
int pcs_train_bias_grad(pcs_ctx* ctx, const float* d_dy, float* d_db, int channels, size_t plane) {
    PCS_TRAIN_ENTER(ctx, d_dy && d_db && channels > 0 && plane > 0, "train_bias_grad");
    return train_plane_sum(ctx, d_dy, d_db, channels, plane);
}
int pcs_train_relu_bwd(pcs_ctx* ctx, float* d_dy, const float* d_y, size_t n) {
    PCS_TRAIN_ENTER(ctx, d_dy && d_y && n > 0, "train_relu_bwd");
    return train_relu_bwd(ctx, d_dy, d_y, n);
}
int pcs_train_maxpool_fwd(pcs_ctx* ctx, const float* d_x, float* d_y, int channels, int H, int W) {
    PCS_TRAIN_ENTER(ctx, d_x && d_y && channels > 0 && H > 0 && W > 0, "train_maxpool_fwd");
    return train_maxpool(ctx, d_x, d_y, nullptr, nullptr, channels, H, W, 0);
}
int pcs_train_maxpool_bwd(pcs_ctx* ctx, const float* d_x, const float* d_dy, float* d_dx, int channels, int H, int W, int accumulate) {
    PCS_TRAIN_ENTER(ctx, d_x && d_dy && d_dx && channels > 0 && H > 0 && W > 0, "train_maxpool_bwd");
    return train_maxpool(ctx, d_x, nullptr, d_dy, d_dx, channels, H, W, accumulate);
}
int pcs_train_deconv2_fwd(pcs_ctx* ctx, const float* d_x, const float* d_k2, const float* d_b, float* d_y, int c_in, int c_out, int h, int w, int relu) {
    PCS_TRAIN_ENTER(ctx, d_x && d_k2 && d_b && d_y && c_in > 0 && c_out > 0 && h > 0 && w > 0, "train_deconv2_fwd");
    return train_deconv2(ctx, 0, d_x, d_k2, d_b, d_y, nullptr, nullptr, nullptr, c_in, c_out, h, w, relu);
}
int pcs_train_deconv2_bwd_data(pcs_ctx* ctx, const float* d_dy, const float* d_k2, float* d_dx, int c_in, int c_out, int h, int w) {
    PCS_TRAIN_ENTER(ctx, d_dy && d_k2 && d_dx && c_in > 0 && c_out > 0 && h > 0 && w > 0, "train_deconv2_bwd_data");
    return train_deconv2(ctx, 1, nullptr, d_k2, nullptr, nullptr, d_dy, d_dx, nullptr, c_in, c_out, h, w, 0);
}
int pcs_train_deconv2_wgrad(pcs_ctx* ctx, const float* d_x, const float* d_dy, float* d_dk2, int c_in, int c_out, int h, int w) {
    PCS_TRAIN_ENTER(ctx, d_x && d_dy && d_dk2 && c_in > 0 && c_in < 65536 && c_out > 0 && h > 0 && w > 0, "train_deconv2_wgrad");
    return train_deconv2(ctx, 2, d_x, nullptr, nullptr, nullptr, d_dy, nullptr, d_dk2, c_in, c_out, h, w, 0);
}
int pcs_train_softmax_ce(pcs_ctx* ctx, const float* d_logits, const uint8_t* d_labels, int n_classes, int H, int W, int Hc, int Wc,
                         float* d_dlogits, double* d_loss_sum) {
    PCS_TRAIN_ENTER(ctx, d_logits && d_labels && d_dlogits && d_loss_sum && n_classes > 0 && Hc > 0 && Wc > 0 && Hc <= H && Wc <= W, "train_softmax_ce");
    return train_softmax_ce(ctx, d_logits, d_labels, n_classes, H, W, Hc, Wc, d_dlogits, d_loss_sum);
}
int pcs_train_input(pcs_ctx* ctx, const uint8_t* d_image, int h, int w, float* d_plane, int H, int W) {
    PCS_TRAIN_ENTER(ctx, d_image && d_plane && h > 0 && w > 0 && H >= h && W >= w, "train_input");
    return train_input_plane(ctx, d_image, h, w, d_plane, H, W);
}
int pcs_train_adam(pcs_ctx* ctx, float* d_params, const float* d_grads, float* d_m, float* d_v, const int64_t* d_offsets, int n_vars, float lr_t,
                   float beta1, float beta2, float eps, float clipnorm, float grad_scale) {
    PCS_TRAIN_ENTER(ctx, d_params && d_grads && d_m && d_v && d_offsets && n_vars > 0, "train_adam");
    return train_adam(ctx, d_params, d_grads, d_m, d_v, reinterpret_cast<const long long*>(d_offsets), n_vars, lr_t, beta1, beta2, eps, clipnorm, grad_scale);
}

int pcs_train_tc_create(pcs_ctx* ctx, int arch, int n_classes, int h, int w, const int64_t* offsets, int n_offsets, pcs_train_tc** out) {
    PCS_TRAIN_ENTER(ctx, offsets && out, "train_tc_create");
    TrainTc* t = nullptr;
    PCS_TRY(train_tc_create(ctx, arch, n_classes, h, w, reinterpret_cast<const long long*>(offsets), n_offsets, &t));
    *out = reinterpret_cast<pcs_train_tc*>(t);
    return PCS_OK;
}
int pcs_train_tc_step(pcs_ctx* ctx, pcs_train_tc* step, int phases, const uint8_t* d_image, const uint8_t* d_labels, const float* d_params,
                      float* d_grads, double* d_loss_sum) {
    PCS_TRAIN_ENTER(ctx, step && d_params && d_grads && d_loss_sum && (phases & 3) && (!(phases & 1) || (d_image && d_labels)), "train_tc_step");
    return train_tc_step(ctx, reinterpret_cast<TrainTc*>(step), phases, d_image, d_labels, d_params, d_grads, d_loss_sum);
}
int pcs_train_tc_wgrad(pcs_ctx* ctx, const void* d_x, int x_planes, const void* d_dy, int dy_planes, int H, int W, int k, int c_in, int c_out,
                       float* d_dw) {
    PCS_TRAIN_ENTER(ctx, d_x && d_dy && d_dw, "train_tc_wgrad");
    return train_tc_wgrad(ctx, d_x, x_planes, d_dy, dy_planes, H, W, k, c_in, c_out, d_dw);
}
int pcs_train_tc_destroy(pcs_ctx* ctx, pcs_train_tc* step) {
    if (!ctx) return PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    return train_tc_destroy(ctx, reinterpret_cast<TrainTc*>(step));
}

int pcs_eval_counts(pcs_ctx* ctx, const uint8_t* d_pred, const uint8_t* d_mask, const uint8_t* d_bin, size_t n_pixels, int n_classes,
                    uint64_t* d_out) {
    if (!ctx || !d_pred || !d_mask || !d_bin || !d_out || !n_pixels) return ctx ? set_err(ctx, PCS_ERR_ARG, "eval_counts: null or empty argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_eval_counts(ctx, d_pred, d_mask, d_bin, n_pixels, n_classes, reinterpret_cast<unsigned long long*>(d_out));
}

size_t pcs_png_bytes(int H, int W, int channels, int level) { return png_file_bytes(H, W, channels, level); }

int pcs_png_encode(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int channels, int level, uint8_t* d_out, size_t stride,
                   uint64_t* d_sizes) {
    if (!ctx || !d_img || !d_out) return ctx ? set_err(ctx, PCS_ERR_ARG, "png_encode: null argument") : PCS_ERR_ARG;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "png_encode");
    return launch_png_encode(ctx, d_img, n, H, W, channels, level, d_out, stride, reinterpret_cast<unsigned long long*>(d_sizes));
}

int pcs_output_pages(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W, const uint8_t* lut, int n_lut,
                     const char* const* paths) {
    if (!ctx || !d_labels || !d_binary || !lut || !paths) return ctx ? set_err(ctx, PCS_ERR_ARG, "output_pages: null argument") : PCS_ERR_ARG;
    if (n_lut < 0 || n_lut > 256) return set_err(ctx, PCS_ERR_ARG, "output_pages: LUT size %d", n_lut);
    for (int i = 0; i < 3 * n; ++i)
        if (!paths[i]) return set_err(ctx, PCS_ERR_ARG, "output_pages: null path");
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    StageScope ts(ctx, "output_pages");
    return output_pages(ctx, d_labels, d_binary, n, H, W, lut, n_lut, paths);
}

int pcs_output_flush(pcs_ctx* ctx) {
    if (!ctx) return PCS_ERR_ARG;
    return output_flush(ctx);
}

// h_png != nullptr: the three masks leave the device as PNG files (level 1) instead of raw arrays: file (page p, kind k)
// at h_png + (3 p + k) * png_stride, its length in h_png_sizes[3 p + k]; kind 0 = color, 1 = overlay, 2 = inverted.
static int predict_pages_host_body(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                                   int cc_majority, const uint8_t* lut, uint8_t* h_image, uint8_t* h_binary, uint8_t* h_labels,
                                   uint8_t* h_color, uint8_t* h_overlay, uint8_t* h_inverted, uint8_t* h_png, size_t png_stride,
                                   uint64_t* h_png_sizes, int32_t* h_stats = nullptr, int max_components = 0, int32_t* h_ncomp = nullptr,
                                   const uint32_t* h_bits = nullptr, int level0 = 0, int level1 = 0, uint32_t* h_binary_bits = nullptr,
                                   uint64_t* ticket = nullptr) {
    // ticket: the streaming form (pcs_predict_pages_*_submit): return once the work is queued, *ticket names the call for pcs_wait_pages
    // h_bits: the pages arrive bit-packed (pcs_preprocess_bits) instead of as uint8 pages; h_binary_bits: `data.binary` leaves bit-packed
    if (!ctx || (!h_bits && (!h_grey || !h_bin))) return ctx ? set_err(ctx, PCS_ERR_ARG, "predict_pages_host: null input") : PCS_ERR_ARG;
    if (h_stats && (max_components <= 0 || !h_labels)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_segments: max_components and the class map are required");
    if (!ctx->model_ready) return set_err(ctx, PCS_ERR_STATE, "pcs_predict_pages_host before pcs_model_load");
    if (n <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0) return set_err(ctx, PCS_ERR_ARG, "predict_pages_host: bad shape");
    const bool want_png = h_png != nullptr;
    const bool submit = ticket != nullptr;
    if (submit && want_png) return set_err(ctx, PCS_ERR_ARG, "predict_pages: the file call has no streaming form (the host places the files)");
    const bool was_chain = ctx->host_chain_ok;
    ctx->host_chain_ok = false;                 // set again when this call has queued completely
    const bool want_masks = h_color || h_overlay || h_inverted || want_png;
    if (want_masks && !lut) return set_err(ctx, PCS_ERR_ARG, "predict_pages_host: colour outputs need a LUT");
    const size_t png_bound = want_png ? png_file_bytes(Hs, Ws, 3, 1) : 0;
    if (want_png && (!h_png_sizes || !png_bound || png_stride < png_bound))
        return set_err(ctx, PCS_ERR_ARG, "predict_pages_files: %zu bytes per file needed (stride %zu), sizes array required", png_bound, png_stride);
    const size_t dpng_stride = (png_bound + 255) / 256 * 256;
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    // Three-stage pipeline over sub-batches ("chunks") of pages: H2D copy stream -> compute stream
    // (ctx->stream) -> D2H copy stream, rotating over `nbuf` device staging buffers, so that with pinned
    // host memory the PCIe traffic of neighbouring chunks hides behind the kernels.  Per page the three
    // stages cost about the same on a B200 behind PCIe 5 x16 (8.7 MB in, 9.7 MB out, ~0.16 ms of
    // kernels), so what is left outside the overlap is the fill (first H2D) and the drain (last D2H):
    // the schedule therefore starts and ends with small chunks and runs large ones in between.
    if (!ctx->copy_streams[0]) {
        PCS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_streams[0], cudaStreamNonBlocking));
        PCS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_streams[1], cudaStreamNonBlocking));
        for (int i = 0; i < pcs_ctx::kHostBufs; ++i) {
            PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
            PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
            PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
            PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sizes[i], cudaEventDisableTiming));
        }
        PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < pcs_ctx::kCallRing; ++i) PCS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_call[i], cudaEventDisableTiming));
    }
    cudaStream_t s_in = ctx->copy_streams[0], s_out = ctx->copy_streams[1], st = ctx->stream;
    // schedule "head,chunk,tail[,nbuf]": pages of the first chunk, of the steady chunks, of the last chunk
    // With raw masks the three stages of a chunk cost about the same and small chunks (2, 4, 8 ... 8, 4, 2) keep the fill and the
    // drain short.  When few bytes travel (compact / packed transport) or the kernels per page are many (segment
    // extraction) the call is bound by the kernels, which run 15 % faster at 32 pages per launch than at 8: measured on
    // one B200 (64 A4 pages): packed 5 275 -> 5 981 pages/s, segments 1 851 -> 2 343, raw masks 4 241 -> 3 801.
    const bool light = h_bits || h_binary_bits || h_stats || !(h_color || h_overlay || h_inverted || h_png);
    int head = light ? 8 : 2, chunk = light ? 32 : 8, tail = light ? 8 : 2, nbuf = 3;
    // uint8 pages in, compact results out: the 8.7 MB per page going up are the longest stage (0.158 ms per page against 0.145 of
    // kernels), so the upload stream must never wait: short fill, chunks of 8 (4 940 -> 5 100 pages/s on one B200)
    // -- and at 10 pages per launch the kernels keep up with it: 4, 10, 10 ... (5 100 -> 5 240; tools/sweep_chunks.py)
    const bool upload_bound = light && !h_bits && !h_stats;
    if (upload_bound) { head = 4; chunk = 10; tail = 10; }
    if (const char* e = getenv("PCSEG_HOST_SCHED")) sscanf(e, "%d,%d,%d,%d", &head, &chunk, &tail, &nbuf);
    if (const char* e = getenv("PCSEG_HOST_CHUNK")) head = chunk = tail = std::max(1, atoi(e));
    chunk = std::max(1, std::min(chunk, n));
    head = std::max(1, std::min(head, chunk));
    tail = std::max(1, std::min(tail, chunk));
    nbuf = std::max(2, std::min(nbuf, (int)pcs_ctx::kHostBufs));
    const bool steady_only = submit && was_chain && !getenv("PCSEG_HOST_SCHED") && !getenv("PCSEG_HOST_CHUNK");   // a chained call: no fill, no drain
    std::vector<int> first, count;             // first page and page count of every chunk
    std::vector<int> fixed;                    // PCSEG_HOST_CHUNKS=a,b,c,...: an explicit schedule (the last size repeats)
    if (const char* e = getenv("PCSEG_HOST_CHUNKS")) {
        for (const char* q = e; *q;) {
            char* end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q) break;
            if (v > 0) fixed.push_back((int)std::min<long>(v, n));
            q = *end ? end + 1 : end;
        }
        for (int v : fixed) chunk = std::max(chunk, v);
    } else if (upload_bound && !getenv("PCSEG_HOST_SCHED") && !getenv("PCSEG_HOST_CHUNK")) {
        fixed = {std::min(4, n), std::min(10, n)};
        if (submit) {
            // streaming form: a chained call has no fill to keep short and its kernels hide under the upload at 16 pages per
            // launch (10: 5 690, 12: 5 870, 16: 6 020, 20: 5 980, 32: 5 950 pages/s, tools/sweep_stream.sh); the first call of a
            // chain ramps up to that size.  Staging is sized for 16 either way, so that the two layouts agree
            fixed = was_chain ? std::vector<int>{std::min(16, n)} : std::vector<int>{std::min(4, n), std::min(12, n), std::min(16, n)};
            chunk = std::max(chunk, std::min(16, n));
        }
    } else if (h_stats && !h_bits && !want_masks && !getenv("PCSEG_HOST_SCHED") && !getenv("PCSEG_HOST_CHUNK")) {
        // segment call with compact results: kernel-bound (190 us per page against 158 us of upload), so a short fill and then
        // launches large enough for the labelling kernels: 6, 12, 16, 16 ... (4 180 -> 4 310 pages/s; tools/sweep_chunks.py segments)
        fixed = {std::min(6, n), std::min(12, n), std::min(16, n)};
        chunk = std::max(chunk, std::min(16, n));
        if (submit) {
            // streaming form: chained calls are kernel-bound with nothing to fill or drain, so the largest launches win
            // (16: 5 030, 24: 5 120, 32: 5 260, 64 with two buffers: 5 040 pages/s; tools/sweep_stream.sh segments)
            fixed = was_chain ? std::vector<int>{std::min(32, n)} : std::vector<int>{std::min(6, n), std::min(12, n), std::min(16, n), std::min(32, n)};
            chunk = std::max(chunk, std::min(32, n));
        }
    }
    {
        int p = 0;
        auto push = [&](int m) { first.push_back(p); count.push_back(m); p += m; };
        if (!fixed.empty()) {
            for (size_t i = 0; p < n; ++i) push(std::min(fixed[std::min(i, fixed.size() - 1)], n - p));
        } else if (n <= chunk) push(n);
        else if (steady_only) {
            while (p < n) push(std::min(chunk, n - p));
        } else {
            push(head);
            // grow geometrically to the steady size, keep `tail` pages (and a shrinking ramp) for the end
            int m = head;
            std::vector<int> ramp_down;
            for (int t = tail, left = n - head; t < chunk && left - t > 0; t *= 2) { ramp_down.push_back(t); left -= t; }
            int reserve = 0;
            for (int t : ramp_down) reserve += t;
            while (n - p - reserve > 0) {
                m = std::min(chunk, m * 2);
                push(std::min(m, n - p - reserve));
            }
            for (auto it = ramp_down.rbegin(); it != ramp_down.rend(); ++it) push(*it);
        }
    }
    const int nchunks = (int)first.size();
    const size_t src1 = (size_t)H * W, dst1 = (size_t)Hs * Ws;
    const bool same = h_grey == h_bin;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t in_words = (src1 + 31) / 32, bm_words = (src1 / 32 + 1 + 3) / 4 * 4;       // packed page on the host / on the device (padded)
    const size_t bb_words = (dst1 + 31) / 32;                                               // packed `binary` of one page
    const size_t in_bytes = h_bits ? al(bm_words * 4 * chunk) : al(src1 * chunk) * (same ? 1 : 2);
    const size_t png_bytes = want_png ? 3 * (size_t)chunk * dpng_stride + al(3 * (size_t)chunk * 8) : 0;
    const size_t stats1 = h_stats ? (size_t)ctx->n_classes * max_components * 5 * sizeof(int32_t) : 0;      // per page
    const size_t seg_bytes = h_stats ? al(stats1 * chunk) + al((size_t)chunk * ctx->n_classes * sizeof(int32_t)) : 0;
    const size_t bb_bytes = h_binary_bits ? al(bb_words * 4 * chunk) : 0;
    const size_t out_bytes = al(dst1 * chunk) * 3 + al(dst1 * chunk * 3) * 3 + png_bytes + seg_bytes + bb_bytes;
    const size_t need = (size_t)nbuf * (in_bytes + out_bytes) + 4096;
    if (want_png && !ctx->h_png_sizes) PCS_CUDA(ctx, cudaHostAlloc(&ctx->h_png_sizes, pcs_ctx::kHostBufs * 3 * 64 * sizeof(uint64_t), cudaHostAllocDefault));
    if (want_png && chunk > 64) return set_err(ctx, PCS_ERR_ARG, "predict_pages_files: chunks of more than 64 pages");
    if (need > ctx->stage_bytes) {
        PCS_CUDA(ctx, cudaDeviceSynchronize());
        if (ctx->stage) cudaFree(ctx->stage);
        ctx->stage = nullptr; ctx->stage_bytes = 0;
        if (cudaMalloc(&ctx->stage, need) != cudaSuccess) {
            cudaGetLastError();
            return set_err(ctx, PCS_ERR_NOMEM, "cudaMalloc of %zu staging bytes failed", need);
        }
        ctx->stage_bytes = need;
    }
    // A submit chains onto the submit before it when both carve the staging memory the same way: the buffers then keep
    // rotating across the call boundary (buffer = running chunk number % nbuf) and only the per-buffer events order the
    // three streams, so this call's upload runs under the kernels of the call before.  Any other call starts behind
    // everything queued so far (the fork below).
    const std::vector<size_t> layout = {(size_t)reinterpret_cast<uintptr_t>(ctx->stage), (size_t)reinterpret_cast<uintptr_t>(ctx->stream), (size_t)nbuf, (size_t)chunk, src1, dst1, (size_t)same,
                                        (size_t)(h_bits != nullptr), in_bytes, out_bytes, png_bytes, seg_bytes, bb_bytes};
    const bool chained = submit && was_chain && layout == ctx->host_layout;
    const uint64_t seq0 = chained ? ctx->host_seq : 0;
    struct Buf { uint8_t *grey, *bin, *image, *binary, *labels, *color, *overlay, *inverted, *png; uint64_t* png_sizes; int32_t *stats, *ncomp; uint32_t* bbits; } buf[pcs_ctx::kHostBufs];
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(ctx->stage);
        for (int i = 0; i < nbuf; ++i) {
            buf[i].grey = p; p += h_bits ? al(bm_words * 4 * chunk) : al(src1 * chunk);
            buf[i].bin = (same || h_bits) ? buf[i].grey : p; if (!same && !h_bits) p += al(src1 * chunk);
            buf[i].image = p; p += al(dst1 * chunk);
            buf[i].binary = p; p += al(dst1 * chunk);
            buf[i].labels = p; p += al(dst1 * chunk);
            buf[i].color = p; p += al(dst1 * chunk * 3);
            buf[i].overlay = p; p += al(dst1 * chunk * 3);
            buf[i].inverted = p; p += al(dst1 * chunk * 3);
            buf[i].png = p; buf[i].png_sizes = reinterpret_cast<uint64_t*>(p + 3 * (size_t)chunk * dpng_stride); p += png_bytes;
            buf[i].stats = reinterpret_cast<int32_t*>(p); buf[i].ncomp = reinterpret_cast<int32_t*>(p + al(stats1 * chunk)); p += seg_bytes;
            buf[i].bbits = reinterpret_cast<uint32_t*>(p); p += bb_bytes;
        }
    }
    // PCSEG_TRACE_HOST: per-chunk device timeline (timing events on the three streams), printed at the end
    const bool trace = getenv("PCSEG_TRACE_HOST") != nullptr && !submit;
    std::vector<cudaEvent_t> tev;
    cudaEvent_t t0 = nullptr;
    auto mark = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev.push_back(e);
    };
    auto enqueue_h2d = [&](int c) -> int {
        const int b = (int)((seq0 + c) % nbuf), p0 = first[c], m = count[c];
        if (c >= nbuf || chained) PCS_CUDA(ctx, cudaStreamWaitEvent(s_in, ctx->ev_comp[b], 0));   // input buffer consumed by the chunk nbuf before
        mark(s_in);
        if (h_bits) {       // packed rows of in_words words into the padded device pitch; the pad words read as zero
            PCS_CUDA(ctx, cudaMemset2DAsync(buf[b].grey + in_words * 4, bm_words * 4, 0, (bm_words - in_words) * 4, m, s_in));
            PCS_CUDA(ctx, cudaMemcpy2DAsync(buf[b].grey, bm_words * 4, h_bits + (size_t)p0 * in_words, in_words * 4, in_words * 4, m,
                                            cudaMemcpyHostToDevice, s_in));
        } else {
            PCS_CUDA(ctx, cudaMemcpyAsync(buf[b].grey, h_grey + (size_t)p0 * src1, src1 * m, cudaMemcpyHostToDevice, s_in));
            if (!same) PCS_CUDA(ctx, cudaMemcpyAsync(buf[b].bin, h_bin + (size_t)p0 * src1, src1 * m, cudaMemcpyHostToDevice, s_in));
        }
        mark(s_in);
        PCS_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[b], s_in));
        return PCS_OK;
    };
    // order the pipeline after whatever the caller already queued on the compute stream
    if (trace) { cudaEventCreate(&t0); cudaEventRecord(t0, st); }
    if (!chained) {
        // (a submit does not join the compute stream to its downloads: wait for the last one here)
        if (ctx->call_seq) PCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_call[(ctx->call_seq - 1) % pcs_ctx::kCallRing], 0));
        PCS_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        PCS_CUDA(ctx, cudaStreamWaitEvent(s_in, ctx->ev_fork, 0));
        PCS_CUDA(ctx, cudaStreamWaitEvent(s_out, ctx->ev_fork, 0));
    }
    const int ahead = nbuf - 1;                                                        // copies in flight ahead of the compute
    // PNG mode: the file lengths of a chunk reach the host first; once they are there the files themselves are copied
    // (only their bytes, not the worst-case buffers).  The host waits for chunk c-1 after it has queued chunk c.
    auto finish_files = [&](int c) -> int {
        const int b = c % nbuf, p0 = first[c], m = count[c];             // (PNG mode never chains: seq0 = 0)
        PCS_CUDA(ctx, cudaEventSynchronize(ctx->ev_sizes[b]));
        const uint64_t* sz = ctx->h_png_sizes + (size_t)b * 3 * 64;
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < m; ++j) {
                const uint64_t bytes = sz[k * 64 + j];
                if (bytes > png_bound) return set_err(ctx, PCS_ERR_CUDA, "predict_pages_files: encoder reported %llu bytes", (unsigned long long)bytes);
                h_png_sizes[(size_t)(p0 + j) * 3 + k] = bytes;
                PCS_CUDA(ctx, cudaMemcpyAsync(h_png + ((size_t)(p0 + j) * 3 + k) * png_stride, buf[b].png + ((size_t)k * chunk + j) * dpng_stride,
                                              bytes, cudaMemcpyDeviceToHost, s_out));
            }
        PCS_CUDA(ctx, cudaEventRecord(ctx->ev_d2h[b], s_out));
        return PCS_OK;
    };
    for (int c = 0; c < std::min(ahead, nchunks); ++c) PCS_TRY(enqueue_h2d(c));
    for (int c = 0; c < nchunks; ++c) {
        const int b = (int)((seq0 + c) % nbuf), p0 = first[c], m = count[c];
        if (c + ahead < nchunks) PCS_TRY(enqueue_h2d(c + ahead));
        PCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_h2d[b], 0));
        if (c >= nbuf || chained) PCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_d2h[b], 0));       // output buffer drained by the chunk nbuf before
        mark(st);
        if (h_bits) PCS_TRY(pcs_preprocess_bits(ctx, reinterpret_cast<const uint32_t*>(buf[b].grey), bm_words, m, H, W, level0, level1, Hs, Ws,
                                                buf[b].image, buf[b].binary));
        else PCS_TRY(pcs_preprocess(ctx, buf[b].grey, buf[b].bin, m, H, W, Hs, Ws, buf[b].image, buf[b].binary, nullptr));
        if (h_binary_bits) PCS_TRY(launch_pack_bits(ctx, buf[b].binary, m, dst1, buf[b].bbits, bb_words));
        if (cc_majority) {
            PCS_TRY(pcs_forward(ctx, buf[b].image, buf[b].binary, m, Hs, Ws, buf[b].labels, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
            PCS_TRY(pcs_cc_majority(ctx, buf[b].labels, buf[b].binary, m, Hs, Ws, ctx->n_classes));
            if (want_masks)
                PCS_TRY(pcs_masks(ctx, buf[b].labels, buf[b].binary, m, Hs, Ws, lut, ctx->n_classes, (h_color || want_png) ? buf[b].color : nullptr,
                                  (h_overlay || want_png) ? buf[b].overlay : nullptr, (h_inverted || want_png) ? buf[b].inverted : nullptr));
        } else {
            PCS_TRY(pcs_forward(ctx, buf[b].image, buf[b].binary, m, Hs, Ws, buf[b].labels, nullptr, nullptr, want_masks ? lut : nullptr,
                                (h_color || want_png) ? buf[b].color : nullptr, (h_overlay || want_png) ? buf[b].overlay : nullptr,
                                (h_inverted || want_png) ? buf[b].inverted : nullptr));
        }
        if (h_stats)        // segment extraction on the final class map (after the vote, when one runs)
            PCS_TRY(pcs_class_components(ctx, buf[b].labels, m, Hs, Ws, ctx->n_classes, buf[b].stats, max_components, buf[b].ncomp));
        if (want_png) {
            uint8_t* const kinds[3] = {buf[b].color, buf[b].overlay, buf[b].inverted};
            for (int k = 0; k < 3; ++k)
                PCS_TRY(pcs_png_encode(ctx, kinds[k], m, Hs, Ws, 3, 1, buf[b].png + (size_t)k * chunk * dpng_stride, dpng_stride,
                                       buf[b].png_sizes + (size_t)k * chunk));
        }
        mark(st);
        PCS_CUDA(ctx, cudaEventRecord(ctx->ev_comp[b], st));
        PCS_CUDA(ctx, cudaStreamWaitEvent(s_out, ctx->ev_comp[b], 0));
        mark(s_out);
        const size_t o1 = (size_t)p0 * dst1, o3 = o1 * 3;
        if (h_image) PCS_CUDA(ctx, cudaMemcpyAsync(h_image + o1, buf[b].image, dst1 * m, cudaMemcpyDeviceToHost, s_out));
        if (h_binary) PCS_CUDA(ctx, cudaMemcpyAsync(h_binary + o1, buf[b].binary, dst1 * m, cudaMemcpyDeviceToHost, s_out));
        if (h_labels) PCS_CUDA(ctx, cudaMemcpyAsync(h_labels + o1, buf[b].labels, dst1 * m, cudaMemcpyDeviceToHost, s_out));
        if (h_color) PCS_CUDA(ctx, cudaMemcpyAsync(h_color + o3, buf[b].color, dst1 * m * 3, cudaMemcpyDeviceToHost, s_out));
        if (h_overlay) PCS_CUDA(ctx, cudaMemcpyAsync(h_overlay + o3, buf[b].overlay, dst1 * m * 3, cudaMemcpyDeviceToHost, s_out));
        if (h_inverted) PCS_CUDA(ctx, cudaMemcpyAsync(h_inverted + o3, buf[b].inverted, dst1 * m * 3, cudaMemcpyDeviceToHost, s_out));
        if (h_binary_bits) PCS_CUDA(ctx, cudaMemcpyAsync(h_binary_bits + (size_t)p0 * bb_words, buf[b].bbits, bb_words * 4 * m, cudaMemcpyDeviceToHost, s_out));
        if (h_stats) {
            PCS_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(h_stats) + (size_t)p0 * stats1, buf[b].stats, stats1 * m, cudaMemcpyDeviceToHost, s_out));
            if (h_ncomp) PCS_CUDA(ctx, cudaMemcpyAsync(h_ncomp + (size_t)p0 * ctx->n_classes, buf[b].ncomp, (size_t)m * ctx->n_classes * sizeof(int32_t),
                                                       cudaMemcpyDeviceToHost, s_out));
        }
        mark(s_out);
        if (want_png) {
            for (int k = 0; k < 3; ++k)
                PCS_CUDA(ctx, cudaMemcpyAsync(ctx->h_png_sizes + (size_t)b * 3 * 64 + k * 64, buf[b].png_sizes + (size_t)k * chunk, (size_t)m * 8,
                                              cudaMemcpyDeviceToHost, s_out));
            PCS_CUDA(ctx, cudaEventRecord(ctx->ev_sizes[b], s_out));
            if (c >= 1) PCS_TRY(finish_files(c - 1));
        } else {
            PCS_CUDA(ctx, cudaEventRecord(ctx->ev_d2h[b], s_out));
        }
    }
    if (want_png) PCS_TRY(finish_files(nchunks - 1));
    if (submit) {
        // streaming form: nothing waits here.  The call's ticket is an event behind its last download; the compute stream is
        // NOT joined to the downloads (the next chained call's kernels need not wait for them), pcs_wait_pages is the join
        PCS_CUDA(ctx, cudaEventRecord(ctx->ev_call[ctx->call_seq % pcs_ctx::kCallRing], s_out));
        *ticket = ctx->call_seq++;
        ctx->host_seq = seq0 + (uint64_t)nchunks;
        ctx->host_layout = layout;
        ctx->host_chain_ok = true;
        return PCS_OK;
    }
    // the compute stream joins the output stream, so the caller's stream order covers the whole call
    PCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_d2h[(nchunks - 1) % nbuf], 0));
    PCS_CUDA(ctx, cudaStreamSynchronize(s_out));
    PCS_CUDA(ctx, cudaStreamSynchronize(st));
    if (trace) {
        // events were pushed in enqueue order: reconstruct (kind, chunk) by replaying that order
        std::vector<std::pair<char, int>> tag;
        for (int c = 0; c < std::min(ahead, nchunks); ++c) tag.push_back({'i', c});
        for (int c = 0; c < nchunks; ++c) {
            if (c + ahead < nchunks) tag.push_back({'i', c + ahead});
            tag.push_back({'k', c});
            tag.push_back({'o', c});
        }
        fprintf(stderr, "[pcs host] n=%d chunks=%d nbuf=%d  (ms since call start: begin-end)\n", n, nchunks, nbuf);
        for (size_t i = 0; i < tag.size(); ++i) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, t0, tev[2 * i]);
            cudaEventElapsedTime(&b, t0, tev[2 * i + 1]);
            fprintf(stderr, "[pcs host]   %s chunk %2d (%2d pages): %7.3f - %7.3f  (%.3f ms)\n",
                    tag[i].first == 'i' ? "h2d    " : tag[i].first == 'k' ? "kernels" : "d2h    ", tag[i].second,
                    count[tag[i].second], a, b, b - a);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
        cudaEventDestroy(t0);
    }
    return PCS_OK;
}

// A failure in the middle of the pipeline (arena exhaustion on a later chunk, an encoder size check) must not return
// while copies into the caller's host buffers or the rotating staging buffers are still in flight: drain the three
// streams first, so that the caller may free its buffers and the next call starts from idle events.
static int predict_pages_host_impl(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                                   int cc_majority, const uint8_t* lut, uint8_t* h_image, uint8_t* h_binary, uint8_t* h_labels,
                                   uint8_t* h_color, uint8_t* h_overlay, uint8_t* h_inverted, uint8_t* h_png, size_t png_stride,
                                   uint64_t* h_png_sizes, int32_t* h_stats = nullptr, int max_components = 0, int32_t* h_ncomp = nullptr,
                                   const uint32_t* h_bits = nullptr, int level0 = 0, int level1 = 0, uint32_t* h_binary_bits = nullptr,
                                   uint64_t* ticket = nullptr) {
    const int rc = predict_pages_host_body(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, lut, h_image, h_binary, h_labels, h_color,
                                           h_overlay, h_inverted, h_png, png_stride, h_png_sizes, h_stats, max_components, h_ncomp,
                                           h_bits, level0, level1, h_binary_bits, ticket);
    if (rc != PCS_OK && ctx) {
        for (cudaStream_t s : {ctx->copy_streams[0], ctx->copy_streams[1]})
            if (s) cudaStreamSynchronize(s);
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
    }
    return rc;
}

int pcs_predict_pages_host(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                           int cc_majority, const uint8_t* lut, uint8_t* h_image, uint8_t* h_binary, uint8_t* h_labels,
                           uint8_t* h_color, uint8_t* h_overlay, uint8_t* h_inverted) {
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, lut, h_image, h_binary, h_labels, h_color, h_overlay,
                                   h_inverted, nullptr, 0, nullptr);
}

int pcs_predict_pages_files(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                            int cc_majority, const uint8_t* lut, uint8_t* h_labels, uint8_t* h_png, size_t png_stride,
                            uint64_t* h_png_sizes) {
    if (ctx && !h_png) return set_err(ctx, PCS_ERR_ARG, "predict_pages_files: null output");
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, lut, nullptr, nullptr, h_labels, nullptr, nullptr, nullptr,
                                   h_png, png_stride, h_png_sizes);
}

int pcs_predict_pages_segments(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                               int cc_majority, const uint8_t* lut, uint8_t* h_labels, uint8_t* h_color, uint8_t* h_overlay,
                               uint8_t* h_inverted, int32_t* h_stats, int max_components, int32_t* h_ncomp) {
    if (ctx && (!h_stats || !h_labels)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_segments: null output");
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, lut, nullptr, nullptr, h_labels, h_color, h_overlay,
                                   h_inverted, nullptr, 0, nullptr, h_stats, max_components, h_ncomp);
}

int pcs_predict_pages_segments_compact(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                                       int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits, int32_t* h_stats, int max_components,
                                       int32_t* h_ncomp) {
    if (ctx && (!h_stats || !h_labels)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_segments_compact: null output");
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, nullptr, nullptr, nullptr, h_labels, nullptr, nullptr,
                                   nullptr, nullptr, 0, nullptr, h_stats, max_components, h_ncomp, nullptr, 0, 0, h_binary_bits);
}

int pcs_predict_pages_compact(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws, int cc_majority,
                              uint8_t* h_labels, uint32_t* h_binary_bits) {
    if (ctx && !h_labels) return set_err(ctx, PCS_ERR_ARG, "predict_pages_compact: null output");
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, nullptr, nullptr, nullptr, h_labels, nullptr, nullptr, nullptr,
                                   nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, 0, 0, h_binary_bits);
}

int pcs_predict_pages_compact_submit(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                                     int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits, uint64_t* ticket) {
    if (ctx && (!h_labels || !ticket)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_compact_submit: null output");
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, nullptr, nullptr, nullptr, h_labels, nullptr, nullptr, nullptr,
                                   nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, 0, 0, h_binary_bits, ticket);
}

int pcs_predict_pages_segments_compact_submit(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W, int Hs, int Ws,
                                              int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits, int32_t* h_stats,
                                              int max_components, int32_t* h_ncomp, uint64_t* ticket) {
    if (ctx && (!h_stats || !h_labels || !ticket)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_segments_compact_submit: null output");
    return predict_pages_host_impl(ctx, h_grey, h_bin, n, H, W, Hs, Ws, cc_majority, nullptr, nullptr, nullptr, h_labels, nullptr, nullptr,
                                   nullptr, nullptr, 0, nullptr, h_stats, max_components, h_ncomp, nullptr, 0, 0, h_binary_bits, ticket);
}

int pcs_predict_pages_packed_submit(pcs_ctx* ctx, const uint32_t* h_bits, int level0, int level1, int n, int H, int W, int Hs, int Ws,
                                    int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits, uint64_t* ticket) {
    if (ctx && (!h_labels || !h_bits || !ticket)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_packed_submit: null argument");
    return predict_pages_host_impl(ctx, nullptr, nullptr, n, H, W, Hs, Ws, cc_majority, nullptr, nullptr, nullptr, h_labels, nullptr, nullptr, nullptr,
                                   nullptr, 0, nullptr, nullptr, 0, nullptr, h_bits, level0, level1, h_binary_bits, ticket);
}

int pcs_wait_pages(pcs_ctx* ctx, uint64_t ticket) {
    if (!ctx) return PCS_ERR_ARG;
    if (ticket >= ctx->call_seq) return set_err(ctx, PCS_ERR_ARG, "wait_pages: ticket %llu was never issued", (unsigned long long)ticket);
    // a slot re-recorded by a later submit lies behind this one on the (in-order) download stream: waiting on it is enough
    PCS_CUDA(ctx, cudaEventSynchronize(ctx->ev_call[ticket % pcs_ctx::kCallRing]));
    return PCS_OK;
}

int pcs_predict_pages_packed(pcs_ctx* ctx, const uint32_t* h_bits, int level0, int level1, int n, int H, int W, int Hs, int Ws, int cc_majority,
                             uint8_t* h_labels, uint32_t* h_binary_bits) {
    if (ctx && (!h_labels || !h_bits)) return set_err(ctx, PCS_ERR_ARG, "predict_pages_packed: null argument");
    return predict_pages_host_impl(ctx, nullptr, nullptr, n, H, W, Hs, Ws, cc_majority, nullptr, nullptr, nullptr, h_labels, nullptr, nullptr, nullptr,
                                   nullptr, 0, nullptr, nullptr, 0, nullptr, h_bits, level0, level1, h_binary_bits);
}

int pcs_debug_activation(pcs_ctx* ctx, const char* name, float* h_out, size_t capacity_floats, int32_t* shape4) {
    if (!ctx || !name) return PCS_ERR_ARG;
    auto it = ctx->acts.find(name);
    if (it == ctx->acts.end()) return set_err(ctx, PCS_ERR_ARG, "no activation named %s in the last forward", name);
    const Act& a = it->second;
    if (shape4) { shape4[0] = a.n; shape4[1] = a.h; shape4[2] = a.w; shape4[3] = a.c; }
    if (!h_out) return a.c;
    const size_t px = (size_t)a.n * a.h * a.w;
    if (px * a.c > capacity_floats) return set_err(ctx, PCS_ERR_ARG, "debug_activation: buffer too small");
    PCS_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint16_t> raw(px * a.cp);
    PCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PCS_CUDA(ctx, cudaMemcpy(raw.data(), a.p, raw.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < px; ++i)
        for (int c = 0; c < a.c; ++c) {
            const int x = (int)(i % a.w), y = (int)((i / a.w) % a.h), pg = (int)(i / ((size_t)a.w * a.h));
            const uint16_t v = raw[act_idx(pg, a.cp, a.h, a.w, c, y, x)];
            float f;
            if (ctx->precision == PCS_PREC_BF16) {
                uint32_t u = (uint32_t)v << 16;
                memcpy(&f, &u, 4);
            } else {
                __half_raw hr; hr.x = v;
                f = __half2float(__half(hr));
            }
            h_out[i * a.c + c] = f;
        }
    return a.c;
}

int pcs_set_keep_activations(pcs_ctx* ctx, int enabled) {
    if (!ctx) return PCS_ERR_ARG;
    ctx->keep_acts = enabled != 0;
    return PCS_OK;
}

int pcs_set_pdl(pcs_ctx* ctx, int enabled) {
    if (!ctx) return PCS_ERR_ARG;
    ctx->pdl = enabled != 0;
    return PCS_OK;
}

int pcs_set_timing(pcs_ctx* ctx, int enabled) {
    if (!ctx) return PCS_ERR_ARG;
    ctx->timing_enabled = enabled != 0;
    if (!enabled) clear_stage_times(ctx);
    return PCS_OK;
}

const char* pcs_last_timings(pcs_ctx* ctx) {
    if (!ctx) return "";
    ctx->timings.clear();
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->stage_times) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.e0, s.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        char buf[128];
        snprintf(buf, sizeof(buf), "%s:%.4f;", s.name.c_str(), ms);
        ctx->timings += buf;
    }
    return ctx->timings.c_str();
}

}  // extern "C"
