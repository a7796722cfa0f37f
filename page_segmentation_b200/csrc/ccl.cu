// Connected-component labelling (4-connectivity) and the class-map
// post-processors built on it.
//
// Replaces cv2.connectedComponentsWithStats(img, connectivity=4) as used by
// ocr4all_pixel_classifier/lib/postprocess.py:10 (vote_connected_component_class),
// :33 (add_bounding_boxes) and lib/image_ops.py:68 (compute_char_height).
//
// Algorithm: union-find over pixels with warp-ballot run detection.
//   A  init     : every warp owns 32 consecutive pixels of a row; the ballot of
//                 the foreground bits gives each pixel its run start, so all
//                 horizontal merges inside a segment cost no atomics;
//   B  merge    : vertical unions only where a run starts or the upper-left
//                 neighbour is background (one union per touching run pair),
//                 plus one union per run crossing a 32-pixel segment border;
//                 union = atomicMin on the larger root (roots only decrease);
//   C  flatten  : label = root = smallest linear index of the component = its
//                 first pixel in raster order;
//   D  rank     : exclusive scan of the root flags -> OpenCV numbering
//                 (components numbered by raster order of their first pixel).
#include "common.cuh"

#include <climits>

namespace pcs {

constexpr int kBG = INT_MIN;

__device__ __forceinline__ int uf_find(const int* parent, int x) {
    // L2 loads: other SMs re-link nodes concurrently; a stale value would still be a valid
    // ancestor, but reading through L2 keeps the retry count low
    int p = __ldcg(parent + x);
    while (p != x) { x = p; p = __ldcg(parent + x); }
    return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }          // a > b: link a under b
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;                                            // somebody re-linked a meanwhile; retry
    }
}

// grid: (ceil(W/32) * rows_per_block..., H, n) -- one warp per 32-pixel row segment.
// block = 256 threads = 8 segments of one row.
template <bool MATCH_CLASS, bool WRITE_BG = true>
__global__ void __launch_bounds__(256)
ccl_init_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, int* __restrict__ parent,
                int* __restrict__ zero_aux, int aux_stride) {
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int x = seg * 32 + lane;
    const int y = blockIdx.y;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const bool inb = x < W;
    const int idx = y * W + x;
    bool fg = false;
    if (inb) {
        const uint8_t v = img[page_off + idx];
        fg = MATCH_CLASS ? (v == cls) : (v != 0);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, fg);
    if (!inb) return;
    if (!fg) { if (WRITE_BG) parent[page_off + idx] = kBG; return; }      // !WRITE_BG: every later pass tests the image first
    const unsigned below = ~mask & ((1u << lane) - 1u);     // background lanes left of me
    const int start = below ? 32 - __clz(below) : 0;
    parent[page_off + idx] = y * W + seg * 32 + start;
    if (zero_aux && start == lane) {
        // run starts are the only root candidates: clear their accumulators
        int* a = zero_aux + (page_off + idx) * aux_stride;
        for (int k = 0; k < aux_stride; ++k) a[k] = 0;
    }
}

template <bool MATCH_CLASS, bool CONN8>
__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, int* __restrict__ parent) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const uint8_t* im = img + page_off;
    int* par = parent + page_off;
    auto isfg = [&](int yy, int xx) -> bool {
        const uint8_t v = im[yy * W + xx];
        return MATCH_CLASS ? (v == cls) : (v != 0);
    };
    if (!isfg(y, x)) return;
    const int idx = y * W + x;
    const bool left = x > 0 && isfg(y, x - 1);
    if (left && (x & 31) == 0) uf_union(par, idx, idx - 1);          // run crosses a segment border
    if (y > 0 && isfg(y - 1, x)) {
        const bool upleft = x > 0 && isfg(y - 1, x - 1);
        if (!left || !upleft) uf_union(par, idx, idx - W);
    } else if (CONN8 && y > 0) {
        // 8-connectivity: the diagonal neighbours matter only when the pixel above is background (otherwise they
        // are in its run); a diagonal that the horizontal neighbour reaches through ITS upper pixel is skipped
        if (!left && x > 0 && isfg(y - 1, x - 1)) uf_union(par, idx, idx - W - 1);
        if (x + 1 < W && isfg(y - 1, x + 1) && !isfg(y, x + 1)) uf_union(par, idx, idx - W + 1);
    }
}

// `fg` (optional): the image whose non-zero pixels are the foreground.  Nine tenths of a page are background, and
// testing the 1-byte pixel first spares the 4-byte parent read of those pixels (and lets cc_majority leave the
// parents of background pixels unwritten).
__global__ void __launch_bounds__(256) ccl_flatten_kernel(int* __restrict__ parent, size_t page_px, const uint8_t* __restrict__ fg) {
    const size_t page_off = (size_t)blockIdx.y * page_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        if (fg) { if (!fg[page_off + i]) continue; }
        else if (parent[page_off + i] == kBG) continue;
        parent[page_off + i] = uf_find(parent + page_off, (int)i);
    }
}
// NOTE: flatten races are benign: a concurrent writer only replaces a parent by
// another ancestor of the same tree (roots never change after the merge kernel).

constexpr int kScanBlock = 1024;   // pixels per scan block (256 threads x 4)

__global__ void __launch_bounds__(256)
ccl_count_roots_kernel(const int* __restrict__ parent, size_t page_px, int nblocks, int* __restrict__ blocksum) {
    const size_t page_off = (size_t)blockIdx.y * page_px;
    const size_t base = (size_t)blockIdx.x * kScanBlock;
    int cnt = 0;
    for (int k = 0; k < 4; ++k) {
        const size_t i = base + k * 256 + threadIdx.x;
        if (i < page_px) cnt += parent[page_off + i] == (int)i;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += s[k];
        blocksum[(size_t)blockIdx.y * nblocks + blockIdx.x] = t;
    }
}

// one block per page: exclusive scan of blocksum in place; total+1 -> ncomp
__global__ void __launch_bounds__(1024)
ccl_scan_blocks_kernel(int* __restrict__ blocksum, int nblocks, int* __restrict__ ncomp) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    int* bs = blocksum + (size_t)blockIdx.x * nblocks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? bs[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (i < nblocks) bs[i] = carry + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && ncomp) ncomp[blockIdx.x] = s_carry + 1;
}

// rank[root] = 1 + number of roots with a smaller linear index
__global__ void __launch_bounds__(256)
ccl_rank_kernel(const int* __restrict__ parent, size_t page_px, int nblocks, const int* __restrict__ blocksum,
                int* __restrict__ rank) {
    const size_t page_off = (size_t)blockIdx.y * page_px;
    const size_t base = (size_t)blockIdx.x * kScanBlock;
    // thread t owns pixels base + 4t .. base + 4t + 3 (contiguous, keeps raster order)
    int flags[4], cnt = 0;
    for (int k = 0; k < 4; ++k) {
        const size_t i = base + (size_t)threadIdx.x * 4 + k;
        flags[k] = (i < page_px) && parent[page_off + i] == (int)i;
        cnt += flags[k];
    }
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = incl;
    __syncthreads();
    int off = blocksum[(size_t)blockIdx.y * nblocks + blockIdx.x];
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) off += s[k];
    int r = off + incl - cnt;
    for (int k = 0; k < 4; ++k) {
        const size_t i = base + (size_t)threadIdx.x * 4 + k;
        if (flags[k]) rank[page_off + i] = ++r;
    }
}

__global__ void __launch_bounds__(256)
ccl_relabel_kernel(const int* __restrict__ parent, const int* __restrict__ rank, size_t page_px,
                   int32_t* __restrict__ labels) {
    const size_t page_off = (size_t)blockIdx.y * page_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        const int p = parent[page_off + i];
        labels[page_off + i] = (p == kBG) ? 0 : rank[page_off + p];
    }
}

__global__ void __launch_bounds__(256) ccl_stats_init_kernel(int32_t* __restrict__ stats, size_t rows) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
        int32_t* s = stats + i * 5;
        s[0] = INT_MAX; s[1] = INT_MAX; s[2] = -1; s[3] = -1; s[4] = 0;   // min x, min y, max x, max y, area
    }
}

__global__ void __launch_bounds__(256)
ccl_stats_kernel(const int32_t* __restrict__ labels, int H, int W, int32_t* __restrict__ stats, int max_components) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    int32_t* st = stats + (size_t)blockIdx.z * max_components * 5;
    const bool inb = x < W;
    const int l = inb ? labels[page_off + (size_t)y * W + x] : -1;
    // background (label 0): warp-aggregated
    const unsigned bgm = __ballot_sync(0xffffffffu, l == 0);
    if (bgm) {
        const int lane = threadIdx.x & 31;
        if (lane == __ffs(bgm) - 1) {
            const int xb = x - lane;
            atomicMin(&st[0], xb + __ffs(bgm) - 1);
            atomicMax(&st[2], xb + 31 - __clz(bgm));
            atomicMin(&st[1], y);
            atomicMax(&st[3], y);
            atomicAdd(&st[4], __popc(bgm));
        }
    }
    if (l > 0 && l < max_components) {
        int32_t* s = st + (size_t)l * 5;
        atomicMin(&s[0], x); atomicMin(&s[1], y); atomicMax(&s[2], x); atomicMax(&s[3], y); atomicAdd(&s[4], 1);
    }
}

__global__ void __launch_bounds__(256)
ccl_stats_finish_kernel(int32_t* __restrict__ stats, const int32_t* __restrict__ ncomp, int max_components) {
    const int page = blockIdx.y;
    const int nc = min(ncomp[page], max_components);
    int32_t* st = stats + (size_t)page * max_components * 5;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max_components; i += gridDim.x * blockDim.x) {
        int32_t* s = st + (size_t)i * 5;
        if (i < nc && s[4] > 0) { s[2] = s[2] - s[0] + 1; s[3] = s[3] - s[1] + 1; }
        else { s[0] = 0; s[1] = 0; s[2] = 0; s[3] = 0; s[4] = 0; }
    }
}

static int ccl_roots(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int cls, bool match, int* parent,
                     int* zero_aux, int aux_stride, bool conn8 = false, bool fg_only = false) {
    // fg_only: the caller's later passes test the image before they touch a parent, so background parents are not written
    cudaStream_t st = ctx->stream;
    dim3 ginit(((W + 31) / 32 + 7) / 8, H, n);
    if (match) ccl_init_kernel<true><<<ginit, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    else if (fg_only) ccl_init_kernel<false, false><<<ginit, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    else ccl_init_kernel<false><<<ginit, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    PCS_LAUNCH_CHECK(ctx, "ccl_init_kernel");
    dim3 gmerge((W + 255) / 256, H, n);
    if (match) ccl_merge_kernel<true, false><<<gmerge, 256, 0, st>>>(d_img, H, W, cls, parent);
    else if (conn8) ccl_merge_kernel<false, true><<<gmerge, 256, 0, st>>>(d_img, H, W, cls, parent);
    else ccl_merge_kernel<false, false><<<gmerge, 256, 0, st>>>(d_img, H, W, cls, parent);
    PCS_LAUNCH_CHECK(ctx, "ccl_merge_kernel");
    const size_t page_px = (size_t)H * W;
    dim3 gflat((unsigned)std::min<size_t>(2048, (page_px + 255) / 256), n);
    ccl_flatten_kernel<<<gflat, 256, 0, st>>>(parent, page_px, match ? nullptr : d_img);
    PCS_LAUNCH_CHECK(ctx, "ccl_flatten_kernel");
    return PCS_OK;
}

int launch_ccl(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int32_t* d_labels, int32_t* d_stats,
               int max_components, int32_t* d_ncomp) {
    if (n <= 0 || H <= 0 || W <= 0 || (size_t)H * W >= (size_t)INT_MAX) return set_err(ctx, PCS_ERR_ARG, "ccl: bad shape");
    if (d_stats && max_components <= 0) return set_err(ctx, PCS_ERR_ARG, "ccl: max_components must be > 0");
    const size_t page_px = (size_t)H * W, total = page_px * n;
    const int nblocks = (int)((page_px + kScanBlock - 1) / kScanBlock);
    const size_t need = total * 4 * 2 + ((size_t)n * nblocks + n) * 4 + 256;
    PCS_TRY(scratch_reserve(ctx, need));
    int* parent = reinterpret_cast<int*>(ctx->scratch);
    int* rank = parent + total;
    int* blocksum = rank + total;
    int* ncomp_tmp = blocksum + (size_t)n * nblocks;
    cudaStream_t st = ctx->stream;
    PCS_TRY(ccl_roots(ctx, d_img, n, H, W, 0, false, parent, nullptr, 0));
    ccl_count_roots_kernel<<<dim3(nblocks, n), 256, 0, st>>>(parent, page_px, nblocks, blocksum);
    PCS_LAUNCH_CHECK(ctx, "ccl_count_roots_kernel");
    int* ncomp = d_ncomp ? d_ncomp : ncomp_tmp;
    ccl_scan_blocks_kernel<<<n, 1024, 0, st>>>(blocksum, nblocks, ncomp);
    PCS_LAUNCH_CHECK(ctx, "ccl_scan_blocks_kernel");
    ccl_rank_kernel<<<dim3(nblocks, n), 256, 0, st>>>(parent, page_px, nblocks, blocksum, rank);
    PCS_LAUNCH_CHECK(ctx, "ccl_rank_kernel");
    dim3 grel((unsigned)std::min<size_t>(2048, (page_px + 255) / 256), n);
    ccl_relabel_kernel<<<grel, 256, 0, st>>>(parent, rank, page_px, d_labels);
    PCS_LAUNCH_CHECK(ctx, "ccl_relabel_kernel");
    if (d_stats) {
        const size_t rows = (size_t)n * max_components;
        ccl_stats_init_kernel<<<(unsigned)std::min<size_t>(1024, (rows + 255) / 256), 256, 0, st>>>(d_stats, rows);
        PCS_LAUNCH_CHECK(ctx, "ccl_stats_init_kernel");
        ccl_stats_kernel<<<dim3((W + 255) / 256, H, n), 256, 0, st>>>(d_labels, H, W, d_stats, max_components);
        PCS_LAUNCH_CHECK(ctx, "ccl_stats_kernel");
        ccl_stats_finish_kernel<<<dim3((max_components + 255) / 256, n), 256, 0, st>>>(d_stats, ncomp, max_components);
        PCS_LAUNCH_CHECK(ctx, "ccl_stats_finish_kernel");
    }
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// vote_connected_component_class (postprocess.py:9-26)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cc_vote_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ fg, const int* __restrict__ parent, size_t page_px,
               int n_classes, int* __restrict__ hist) {
    const size_t page_off = (size_t)blockIdx.y * page_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        if (!fg[page_off + i]) continue;
        const int p = parent[page_off + i];
        const int c = pred[page_off + i];
        if (c < n_classes) atomicAdd(&hist[(page_off + p) * n_classes + c], 1);
    }
}

__global__ void __launch_bounds__(256)
cc_apply_kernel(uint8_t* __restrict__ pred, const uint8_t* __restrict__ fg, const int* __restrict__ parent, size_t page_px,
                int n_classes, const int* __restrict__ hist) {
    const size_t page_off = (size_t)blockIdx.y * page_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        if (!fg[page_off + i]) continue;
        const int p = parent[page_off + i];
        const int* h = hist + (page_off + p) * n_classes;
        int best = 0, bv = h[0];
        for (int c = 1; c < n_classes; ++c)
            if (h[c] > bv) { bv = h[c]; best = c; }          // ties -> lowest class (np.argmax)
        pred[page_off + i] = (uint8_t)best;
    }
}

int launch_cc_majority(pcs_ctx* ctx, uint8_t* d_pred, const uint8_t* d_binary, int n, int H, int W, int n_classes) {
    if (n <= 0 || H <= 0 || W <= 0 || n_classes <= 0 || n_classes > 255 || (size_t)H * W >= (size_t)INT_MAX)
        return set_err(ctx, PCS_ERR_ARG, "cc_majority: bad argument");
    const size_t page_px = (size_t)H * W, total = page_px * n;
    PCS_TRY(scratch_reserve(ctx, total * 4 * (1 + (size_t)n_classes) + 256));
    int* parent = reinterpret_cast<int*>(ctx->scratch);
    int* hist = parent + total;
    PCS_TRY(ccl_roots(ctx, d_binary, n, H, W, 0, false, parent, hist, n_classes, false, /*fg_only=*/true));
    dim3 grid((unsigned)std::min<size_t>(2048, (page_px + 255) / 256), n);
    cc_vote_kernel<<<grid, 256, 0, ctx->stream>>>(d_pred, d_binary, parent, page_px, n_classes, hist);
    PCS_LAUNCH_CHECK(ctx, "cc_vote_kernel");
    cc_apply_kernel<<<grid, 256, 0, ctx->stream>>>(d_pred, d_binary, parent, page_px, n_classes, hist);
    PCS_LAUNCH_CHECK(ctx, "cc_apply_kernel");
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// add_bounding_boxes (postprocess.py:29-42, evident intent): for c ascending,
// every 4-connected component of (pred == c) paints its bounding box with c.
// Boxes are rasterised through a 2-D difference array + prefix sums.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bbox_accum_kernel(const int* __restrict__ parent, int H, int W, int* __restrict__ box /*[px][4]*/, const uint8_t* __restrict__ fg) {
    // fg (optional): the labelled image; when given, background pixels are recognised by their byte and their
    // (then unwritten) parents are never read
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    if (fg && !fg[page_off + (size_t)y * W + x]) return;
    const int p = parent[page_off + (size_t)y * W + x];
    if (p == kBG) return;
    int* b = box + (page_off + p) * 4;
    // accumulators are zero-initialised at run starts: keep (W - min x, H - min y, max x, max y) as maxima.
    // A large component (a picture block, the page background of add_bounding_boxes) would send every one of
    // its pixels to the same four words: lanes of a warp that share a root first reduce among themselves, and
    // the leader only issues the atomics that can still grow the box (a stale read merely costs an atomic).
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, p);
    const int v0 = __reduce_max_sync(peers, W - x), v1 = __reduce_max_sync(peers, H - y);
    const int v2 = __reduce_max_sync(peers, x), v3 = __reduce_max_sync(peers, y);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        const int4 cur = __ldcg(reinterpret_cast<const int4*>(b));
        if (v0 > cur.x) atomicMax(&b[0], v0);
        if (v1 > cur.y) atomicMax(&b[1], v1);
        if (v2 > cur.z) atomicMax(&b[2], v2);
        if (v3 > cur.w) atomicMax(&b[3], v3);
    }
}

__global__ void __launch_bounds__(256)
bbox_diff_kernel(const int* __restrict__ parent, int H, int W, const int* __restrict__ box, int* __restrict__ diff) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const int idx = y * W + x;
    if (parent[page_off + idx] != idx) return;            // roots only
    const int* b = box + (page_off + idx) * 4;
    const int x0 = W - b[0], y0 = H - b[1], x1 = b[2], y1 = b[3];
    int* d = diff + (size_t)blockIdx.z * (H + 1) * (W + 1);
    atomicAdd(&d[(size_t)y0 * (W + 1) + x0], 1);
    atomicAdd(&d[(size_t)y0 * (W + 1) + x1 + 1], -1);
    atomicAdd(&d[(size_t)(y1 + 1) * (W + 1) + x0], -1);
    atomicAdd(&d[(size_t)(y1 + 1) * (W + 1) + x1 + 1], 1);
}

// in-place inclusive scan along rows: one warp per row
__global__ void __launch_bounds__(256) diff_rowscan_kernel(int* __restrict__ diff, int rows, int cols) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    int* d = diff + ((size_t)blockIdx.y * rows + row) * cols;
    int carry = 0;
    for (int base = 0; base < cols; base += 32) {
        const int i = base + lane;
        int v = i < cols ? d[i] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        v += carry;
        if (i < cols) d[i] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// column scan fused with the paint: coverage > 0 -> out = cls
__global__ void __launch_bounds__(256)
diff_colscan_paint_kernel(const int* __restrict__ diff, int H, int W, int cls, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int* d = diff + (size_t)blockIdx.y * (H + 1) * (W + 1);
    uint8_t* o = out + (size_t)blockIdx.y * H * W;
    int acc = 0;
    for (int y = 0; y < H; ++y) {
        acc += d[(size_t)y * (W + 1) + x];
        if (acc > 0) o[(size_t)y * W + x] = (uint8_t)cls;
    }
}

int launch_bounding_boxes(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out) {
    if (n <= 0 || H <= 0 || W <= 0 || n_classes <= 0 || n_classes > 255 || (size_t)H * W >= (size_t)INT_MAX)
        return set_err(ctx, PCS_ERR_ARG, "bounding_boxes: bad argument");
    const size_t page_px = (size_t)H * W, total = page_px * n;
    const size_t diff_elems = (size_t)n * (H + 1) * (W + 1);
    const size_t total4 = (total + 3) / 4 * 4;                           // keeps the int4 box records 16-byte aligned
    PCS_TRY(scratch_reserve(ctx, total4 * 4 * 5 + diff_elems * 4 + 512));
    int* parent = reinterpret_cast<int*>(ctx->scratch);
    int* box = parent + total4;
    int* diff = box + total4 * 4;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(d_out, 0, total, st));                 // newpred = zeros_like(pred)
    // classes = np.unique(pred) per page in the reference; painting an absent class is a no-op
    for (int c = 0; c < n_classes; ++c) {
        PCS_TRY(ccl_roots(ctx, d_pred, n, H, W, c, true, parent, box, 4));
        dim3 g((W + 255) / 256, H, n);
        bbox_accum_kernel<<<g, 256, 0, st>>>(parent, H, W, box, nullptr);
        PCS_LAUNCH_CHECK(ctx, "bbox_accum_kernel");
        PCS_CUDA(ctx, cudaMemsetAsync(diff, 0, diff_elems * 4, st));
        bbox_diff_kernel<<<g, 256, 0, st>>>(parent, H, W, box, diff);
        PCS_LAUNCH_CHECK(ctx, "bbox_diff_kernel");
        diff_rowscan_kernel<<<dim3((H + 1 + 7) / 8, n), 256, 0, st>>>(diff, H + 1, W + 1);
        PCS_LAUNCH_CHECK(ctx, "diff_rowscan_kernel");
        diff_colscan_paint_kernel<<<dim3((W + 255) / 256, n), 256, 0, st>>>(diff, H, W, c, d_out);
        PCS_LAUNCH_CHECK(ctx, "diff_colscan_paint_kernel");
    }
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// compute_char_height (lib/image_ops.py:58-82): Otsu threshold, connected components, letter-like boxes
// (0.5 < w/h < 2, 10 < h < 60, 5 < w < 50), the height at index len/2 of the sorted valid heights.
// The reference calls cv2.connectedComponentsWithStats(img, 4): the positional 4 lands in the `labels`
// output slot and is ignored, so the components are 8-CONNECTED (cv2 default); restated as such.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hist256_kernel(const uint8_t* __restrict__ img, size_t page_px, unsigned* __restrict__ hist) {
    __shared__ unsigned s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* p = img + (size_t)blockIdx.y * page_px;
    // 16 consecutive pixels per thread, equal neighbours counted as one run: scans are mostly paper, so a thread
    // usually issues one shared-memory atomic per 16 pixels instead of 16 to the same bin
    const size_t chunks = page_px / 16;
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    if (aligned) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (size_t)gridDim.x * blockDim.x) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + i);
            const unsigned w[4] = {q.x, q.y, q.z, q.w};
            unsigned cur = w[0] & 0xffu, run = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned v = (w[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                if (v != cur) { atomicAdd(&s_h[cur], run); cur = v; run = 0; }
                ++run;
            }
            atomicAdd(&s_h[cur], run);
        }
    }
    for (size_t i = (aligned ? chunks * 16 : 0) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px;
         i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&s_h[p[i]], 1u);
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[(size_t)blockIdx.y * 256 + threadIdx.x], s_h[threadIdx.x]);
}

// cv2 getThreshVal_Otsu_8u in its operation order (double precision, no contraction); one thread per page
__global__ void otsu_kernel(const unsigned* __restrict__ hist, size_t page_px, int n, int* __restrict__ thresh) {
    const int page = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n) return;
    const unsigned* h = hist + (size_t)page * 256;
    const double scale = __ddiv_rn(1.0, (double)page_px);
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)h[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        const double p_i = __dmul_rn((double)h[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        const double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < 1.1920928955078125e-07 || fmax(q1, q2) > 1.0 - 1.1920928955078125e-07) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    thresh[page] = max_val;
}

// foreground of the component analysis: THRESH_BINARY (v > t -> 255), inverted unless `inverse`
__global__ void __launch_bounds__(256)
otsu_fg_kernel(const uint8_t* __restrict__ img, size_t page_px, const int* __restrict__ thresh, int inverse, uint8_t* __restrict__ fg) {
    const int t = thresh[blockIdx.y];
    const size_t off = (size_t)blockIdx.y * page_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        const bool above = img[off + i] > t;
        fg[off + i] = (above == (inverse != 0)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
letter_heights_kernel(const int* __restrict__ parent, const uint8_t* __restrict__ fg, int H, int W, const int* __restrict__ box,
                      unsigned* __restrict__ hh /*[n][64]*/) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const int idx = y * W + x;
    if (!fg[page_off + idx] || parent[page_off + idx] != idx) return;            // roots only (background parents are unwritten)
    const int* b = box + (page_off + idx) * 4;
    const int w = b[2] - (W - b[0]) + 1, h = b[3] - (H - b[1]) + 1;
    // 0.5 < w/h < 2  <=>  h < 2w and w < 2h (exact for these small integers)
    if (h < 2 * w && w < 2 * h && h > 10 && h < 60 && w > 5 && w < 50) atomicAdd(&hh[(size_t)blockIdx.z * 64 + h], 1u);
}

// sorted(valid heights)[len / 2], or -1 when there is no valid letter (the reference returns None)
__global__ void median_height_kernel(const unsigned* __restrict__ hh, int n, int32_t* __restrict__ out) {
    const int page = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n) return;
    const unsigned* h = hh + (size_t)page * 64;
    unsigned total = 0;
    for (int i = 0; i < 64; ++i) total += h[i];
    int res = -1;
    if (total) {
        const unsigned k = total / 2;
        unsigned cum = 0;
        for (int i = 0; i < 64; ++i) { cum += h[i]; if (cum > k) { res = i; break; } }
    }
    out[page] = res;
}

int launch_char_height(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int inverse, int32_t* d_out) {
    if (n <= 0 || H <= 0 || W <= 0 || (size_t)H * W >= (size_t)INT_MAX) return set_err(ctx, PCS_ERR_ARG, "char_height: bad shape");
    const size_t page_px = (size_t)H * W, total = page_px * n;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t need = al(total * 4) + al(total * 16) + al(total) + al((size_t)n * 256 * 4) + al((size_t)n * 64 * 4) + al((size_t)n * 4) + 256;
    PCS_TRY(scratch_reserve(ctx, need));
    char* q = reinterpret_cast<char*>(ctx->scratch);
    int* parent = reinterpret_cast<int*>(q); q += al(total * 4);
    int* box = reinterpret_cast<int*>(q); q += al(total * 16);
    uint8_t* fg = reinterpret_cast<uint8_t*>(q); q += al(total);
    unsigned* hist = reinterpret_cast<unsigned*>(q); q += al((size_t)n * 256 * 4);
    unsigned* hh = reinterpret_cast<unsigned*>(q); q += al((size_t)n * 64 * 4);
    int* thresh = reinterpret_cast<int*>(q);
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)n * 256 * 4, st));
    PCS_CUDA(ctx, cudaMemsetAsync(hh, 0, (size_t)n * 64 * 4, st));
    const dim3 gflat((unsigned)std::min<size_t>(1184, (page_px + 255) / 256), n);
    hist256_kernel<<<gflat, 256, 0, st>>>(d_img, page_px, hist);
    PCS_LAUNCH_CHECK(ctx, "hist256_kernel");
    otsu_kernel<<<(n + 63) / 64, 64, 0, st>>>(hist, page_px, n, thresh);
    PCS_LAUNCH_CHECK(ctx, "otsu_kernel");
    otsu_fg_kernel<<<gflat, 256, 0, st>>>(d_img, page_px, thresh, inverse, fg);
    PCS_LAUNCH_CHECK(ctx, "otsu_fg_kernel");
    PCS_TRY(ccl_roots(ctx, fg, n, H, W, 0, false, parent, box, 4, /*conn8=*/true, /*fg_only=*/true));
    const dim3 g((W + 255) / 256, H, n);
    bbox_accum_kernel<<<g, 256, 0, st>>>(parent, H, W, box, fg);
    PCS_LAUNCH_CHECK(ctx, "bbox_accum_kernel");
    letter_heights_kernel<<<g, 256, 0, st>>>(parent, fg, H, W, box, hh);
    PCS_LAUNCH_CHECK(ctx, "letter_heights_kernel");
    median_height_kernel<<<(n + 63) / 64, 64, 0, st>>>(hh, n, d_out);
    PCS_LAUNCH_CHECK(ctx, "median_height_kernel");
    return PCS_OK;
}

}  // namespace pcs
