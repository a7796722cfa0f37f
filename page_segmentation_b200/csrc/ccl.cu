// Connected-component labelling (4-connectivity) and the class-map
// post-processors built on it.
//
// Replaces cv2.connectedComponentsWithStats(img, connectivity=4) as used by
// ocr4all_pixel_classifier/lib/postprocess.py:10 (vote_connected_component_class),
// :33 (add_bounding_boxes) and lib/image_ops.py:68 (compute_char_height).
//
// Algorithm: union-find over pixels; a thread owns a 32-pixel row segment as a bit mask.
//   A  tile     : every 256 x 32 tile is labelled on its own in shared memory: parent =
//                 first pixel of the pixel's run inside its segment (horizontal merges
//                 inside a segment cost nothing), vertical unions only where a run
//                 starts or the upper-left neighbour is background (one union per
//                 touching run pair), plus one union per run crossing a segment border;
//                 the global parent of a pixel is its tile-local root;
//   B  borders  : the same unions for pixel pairs in different tiles, on the global
//                 parents; union = atomicMin on the larger root (roots only decrease);
//   C  flatten  : label = root = smallest linear index of the component = its
//                 first pixel in raster order (one find per run; consumers that walk
//                 runs themselves -- bounding boxes -- skip this pass);
//   D  rank     : exclusive scan of the root flags -> OpenCV numbering
//                 (components numbered by raster order of their first pixel).
#include "common.cuh"

#include "ccl_common.cuh"

namespace pcs {

// A  tile: every 256 x 32 tile is labelled on its own in shared memory (pixels outside the tile count as
// background): parent = first pixel of the run inside the 32-pixel segment, one union per touching run pair
// (vertical unions only where a run starts or the upper-left neighbour is background) with shared-memory atomics,
// one find per run; then every pixel's parent is written as the GLOBAL index of its tile-local root.  The local
// order (row, then x) is the global raster order, so the local root is the tile's first pixel of the component.
// The 32 masks of a warp are handed round so that every store instruction writes 32 consecutive parents.
template <int MODE, bool CONN8, bool WRITE_BG>
__global__ void __launch_bounds__(256)
ccl_tile_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, int* __restrict__ parent,
                int* __restrict__ zero_aux, int aux_stride) {
    __shared__ int lpar[kTileRows * kTileSegs * 32];
    __shared__ unsigned smask[kTileRows][kTileSegs];
    const int tid = threadIdx.x, lane = tid & 31, sx = tid & (kTileSegs - 1), ry = tid / kTileSegs;
    const int ty0 = blockIdx.y * kTileRows, tx0 = blockIdx.x * kTileSegs * 32;
    const int y = ty0 + ry, x0 = tx0 + sx * 32;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const bool valid = y < H && x0 < W;
    const bool last_row = y == H - 1 && blockIdx.z == gridDim.z - 1;
    const unsigned m = valid ? fg_bits<MODE>(img + page_off + (size_t)y * W, x0, W, cls, last_row) : 0u;
    smask[ry][sx] = m;
    const int l0 = tid * 32;                                            // = ry * 256 + sx * 32
    const unsigned nz = __ballot_sync(0xffffffffu, m != 0);            // segments of this warp that hold foreground
    for (unsigned todo = nz; todo; todo &= todo - 1) {                  // warp-uniform loop
        const int j = __ffs(todo) - 1;
        const unsigned mj = __shfl_sync(0xffffffffu, m, j);
        if ((mj >> lane) & 1u) {
            const unsigned below = ~mj & ((1u << lane) - 1u);           // background lanes left of me
            const int lb = (tid - lane + j) * 32;
            lpar[lb + lane] = lb + (below ? 32 - __clz(below) : 0);
        }
    }
    __syncthreads();
    if (m) {
        const unsigned lb = sx > 0 ? smask[ry][sx - 1] >> 31 : 0u;
        if ((m & 1u) && lb) suf_union_h(lpar, l0, l0 - 1);                // run crosses a segment border
        if (ry > 0) {
            const unsigned up = smask[ry - 1][sx];
            const unsigned ulb = sx > 0 ? smask[ry - 1][sx - 1] >> 31 : 0u;
            const unsigned leftm = (m << 1) | lb, upleftm = (up << 1) | ulb;
            unsigned v = m & up & ~(leftm & upleftm);
            while (v) {
                const int k = __ffs(v) - 1;
                v &= v - 1;
                suf_union_h(lpar, l0 + k, l0 + k - 256);
            }
            if (CONN8) {
                // 8-connectivity: the diagonal neighbours matter only when the pixel above is background (otherwise
                // they are in its run); a diagonal that the horizontal neighbour reaches through ITS upper pixel is skipped
                unsigned d1 = m & ~up & ~leftm & upleftm;
                while (d1) {
                    const int k = __ffs(d1) - 1;
                    d1 &= d1 - 1;
                    suf_union_h(lpar, l0 + k, l0 + k - 256 - 1);
                }
                const unsigned rb = sx + 1 < kTileSegs ? smask[ry][sx + 1] & 1u : 0u;
                const unsigned urb = sx + 1 < kTileSegs ? smask[ry - 1][sx + 1] & 1u : 0u;
                unsigned d2 = m & ~up & ((up >> 1) | (urb << 31)) & ~((m >> 1) | (rb << 31));
                while (d2) {
                    const int k = __ffs(d2) - 1;
                    d2 &= d2 - 1;
                    suf_union_h(lpar, l0 + k, l0 + k - 256 + 1);
                }
            }
        }
    }
    __syncthreads();
    {   // one find per run, kept at the run's first pixel (a concurrent walker reads the old or the new ancestor)
        unsigned mm = m;
        int s, len;
        while (next_run(mm, s, len)) {
            const int r = suf_find(lpar, l0 + s);
            lpar[l0 + s] = r;
        }
    }
    __syncthreads();
    int* par = parent + page_off;
    // segments to write: the ones with foreground, or (WRITE_BG) every segment inside the page
    for (unsigned todo = WRITE_BG ? __ballot_sync(0xffffffffu, valid) : nz; todo; todo &= todo - 1) {
        const int j = __ffs(todo) - 1;
        const unsigned mj = __shfl_sync(0xffffffffu, m, j);
        const int tj = tid - lane + j;                                  // the thread that owns segment j: warp-uniform
        const int xj = tx0 + (tj & (kTileSegs - 1)) * 32;
        const int bj = (ty0 + tj / kTileSegs) * W + xj, nj = min(32, W - xj);
        if (lane < nj) {
            if ((mj >> lane) & 1u) {
                const unsigned below = ~mj & ((1u << lane) - 1u);
                const int start = below ? 32 - __clz(below) : 0;
                const int r = lpar[(tid - lane + j) * 32 + start];
                par[bj + lane] = (ty0 + (r >> 8)) * W + tx0 + (r & 255);
                if (zero_aux && start == lane) {
                    // run starts are the only root candidates: clear their accumulators
                    int* z = zero_aux + (page_off + bj + lane) * aux_stride;
                    for (int k = 0; k < aux_stride; ++k) z[k] = 0;
                }
            } else if (WRITE_BG) {
                par[bj + lane] = kBG;
            }
        }
    }
}

// B  borders: the unions between pixels of different tiles, on the global parents.  Rows that start a tile run the
// complete rule set against the row above; elsewhere only the first segment of a tile has a neighbour outside it
// (to the left; with 8-connectivity also the upper-left diagonal of its first pixel, and the upper-right diagonal
// of the last pixel of the tile's last segment).
template <int MODE, bool CONN8, bool COMPRESS>
__global__ void __launch_bounds__(256)
ccl_border_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, int* __restrict__ parent) {
    PCS_SEG_THREAD();
    if (!valid) return;
    const int sx = (x0 >> 5) & (kTileSegs - 1);
    const bool hrow = (y & (kTileRows - 1)) == 0;
    if (!hrow && sx != 0 && !(CONN8 && sx == kTileSegs - 1)) return;
    const uint8_t* im = img + page_off;
    const uint8_t* row = im + (size_t)y * W;
    const unsigned m = fg_bits<MODE>(row, x0, W, cls, last_row);
    if (!m) return;
    auto isfg = [&](const uint8_t* r, int xx) -> unsigned {
        const uint8_t v = __ldg(r + xx);
        return MODE == 1 ? (v == cls) : (v != 0);
    };
    int* par = parent + page_off;
    const int idx0 = y * W + x0;
    const unsigned lb = ((m & 1u) && x0 > 0) ? isfg(row, x0 - 1) : 0u;
    if (lb && sx == 0) uf_union<COMPRESS>(par, idx0, idx0 - 1);         // run crosses a tile border
    if (y == 0 || (!hrow && !CONN8)) return;
    const uint8_t* rup = row - W;
    const unsigned up = fg_bits<MODE>(rup, x0, W, cls, false);
    const unsigned ulb = ((m & 1u) && x0 > 0) ? isfg(rup, x0 - 1) : 0u;
    const unsigned leftm = (m << 1) | lb;                               // bit k: pixel left of k
    const unsigned upleftm = (up << 1) | ulb;                           // bit k: pixel above-left of k
    unsigned v = hrow ? m & up & ~(leftm & upleftm) : 0u;
    while (v) {
        const int k = __ffs(v) - 1;
        v &= v - 1;
        uf_union<COMPRESS>(par, idx0 + k, idx0 + k - W);
    }
    if (CONN8) {
        unsigned d1 = m & ~up & ~leftm & upleftm;
        if (!hrow) d1 &= sx == 0 ? 1u : 0u;
        while (d1) {
            const int k = __ffs(d1) - 1;
            d1 &= d1 - 1;
            uf_union<COMPRESS>(par, idx0 + k, idx0 + k - W - 1);
        }
        const bool edge = (m >> 31) && x0 + 32 < W;
        const unsigned rb = edge ? isfg(row, x0 + 32) : 0u, urb = edge ? isfg(rup, x0 + 32) : 0u;
        unsigned d2 = m & ~up & ((up >> 1) | (urb << 31)) & ~((m >> 1) | (rb << 31));
        if (!hrow) d2 &= sx == kTileSegs - 1 ? 0x80000000u : 0u;
        while (d2) {
            const int k = __ffs(d2) - 1;
            d2 &= d2 - 1;
            uf_union<COMPRESS>(par, idx0 + k, idx0 + k - W + 1);
        }
    }
}

// C  flatten: one find per run (all pixels of a run still point at its first pixel), written to every pixel of it
__global__ void __launch_bounds__(256) ccl_flatten_kernel(const uint8_t* __restrict__ img, int H, int W, int* __restrict__ parent) {
    PCS_SEG_THREAD();
    if (!valid) return;
    unsigned mm = fg_bits<0>(img + page_off + (size_t)y * W, x0, W, 0, last_row);
    int* par = parent + page_off;
    const int base = y * W + x0;
    int s, len;
    while (next_run(mm, s, len)) {
        const int root = uf_find(par, base + s);
        for (int k = 0; k < len; ++k) par[base + s + k] = root;
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


// init + merge (+ flatten).  After the merge every tree's root is the component's first pixel in raster order;
// `flatten` additionally makes every foreground pixel point at it (callers that only walk runs find the root
// themselves and skip that pass).
static int ccl_roots(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int cls, bool match, int* parent,
                     int* zero_aux, int aux_stride, bool conn8 = false, bool fg_only = false, bool flatten = true) {
    // fg_only: the caller's later passes test the image before they touch a parent, so background parents are not written
    cudaStream_t st = ctx->stream;
    const dim3 g = seg_grid(H, W, n);
    const dim3 gt((W + kTileSegs * 32 - 1) / (kTileSegs * 32), (H + kTileRows - 1) / kTileRows, n);
    if (match) ccl_tile_kernel<1, false, true><<<gt, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    else if (conn8 && fg_only) ccl_tile_kernel<0, true, false><<<gt, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    else if (conn8) ccl_tile_kernel<0, true, true><<<gt, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    else if (fg_only) ccl_tile_kernel<0, false, false><<<gt, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    else ccl_tile_kernel<0, false, true><<<gt, 256, 0, st>>>(d_img, H, W, cls, parent, zero_aux, aux_stride);
    PCS_LAUNCH_CHECK(ctx, "ccl_tile_kernel");
    const bool cz = ccl_compress(match);
    if (match) {
        if (cz) ccl_border_kernel<1, false, true><<<g, 256, 0, st>>>(d_img, H, W, cls, parent);
        else ccl_border_kernel<1, false, false><<<g, 256, 0, st>>>(d_img, H, W, cls, parent);
    } else if (conn8) {
        if (cz) ccl_border_kernel<0, true, true><<<g, 256, 0, st>>>(d_img, H, W, cls, parent);
        else ccl_border_kernel<0, true, false><<<g, 256, 0, st>>>(d_img, H, W, cls, parent);
    } else {
        if (cz) ccl_border_kernel<0, false, true><<<g, 256, 0, st>>>(d_img, H, W, cls, parent);
        else ccl_border_kernel<0, false, false><<<g, 256, 0, st>>>(d_img, H, W, cls, parent);
    }
    PCS_LAUNCH_CHECK(ctx, "ccl_border_kernel");
    if (flatten) {
        if (match) return set_err(ctx, PCS_ERR_ARG, "ccl_roots: flatten needs a non-zero-foreground image");
        ccl_flatten_kernel<<<g, 256, 0, st>>>(d_img, H, W, parent);
        PCS_LAUNCH_CHECK(ctx, "ccl_flatten_kernel");
    }
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
// one histogram update per RUN and class: the root is found from the run's first pixel and kept there for the
// apply pass (no flatten pass), the classes of the run's pixels are counted on bit masks
__global__ void __launch_bounds__(256)
cc_vote_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ fg, int H, int W, int* __restrict__ parent,
               int n_classes, int* __restrict__ hist) {
    PCS_SEG_THREAD();
    if (!valid) return;
    const unsigned m = fg_bits<0>(fg + page_off + (size_t)y * W, x0, W, 0, last_row);
    if (!m) return;
    unsigned w[8];
    load_seg32(pred + page_off + (size_t)y * W, x0, W, last_row, 0xffu, w);
    int* par = parent + page_off;
    const int base = y * W + x0;
    unsigned mm = m;
    int s, len;
    while (next_run(mm, s, len)) {                                      // roots first: the class loop re-reads them
        const int root = uf_find(par, base + s);
        par[base + s] = root;
    }
    for (int c = 0; c < n_classes; ++c) {
        const unsigned bc = seg_bits<1>(w, c) & m;
        if (!bc) continue;
        mm = m;
        while (next_run(mm, s, len)) {
            const unsigned run = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
            const int cnt = __popc(bc & run);
            if (cnt) atomicAdd(&hist[(page_off + par[base + s]) * n_classes + c], cnt);
        }
    }
}

__global__ void __launch_bounds__(256)
cc_apply_kernel(uint8_t* __restrict__ pred, const uint8_t* __restrict__ fg, int H, int W, const int* __restrict__ parent,
                int n_classes, const int* __restrict__ hist) {
    PCS_SEG_THREAD();
    if (!valid) return;
    unsigned mm = fg_bits<0>(fg + page_off + (size_t)y * W, x0, W, 0, last_row);
    const int base = y * W + x0;
    int s, len;
    while (next_run(mm, s, len)) {
        const int* h = hist + (page_off + uf_find(parent + page_off, base + s)) * n_classes;   // run start -> (tile root ->) root
        int best = 0, bv = h[0];
        for (int c = 1; c < n_classes; ++c)
            if (h[c] > bv) { bv = h[c]; best = c; }          // ties -> lowest class (np.argmax)
        uint8_t* o = pred + page_off + base + s;
        for (int k = 0; k < len; ++k) o[k] = (uint8_t)best;
    }
}

// ---------------------------------------------------------------------------
// The vote with tile-local histograms (n_classes <= kVoteClasses).  cc_vote_kernel above walks from every run to its root
// through L2 and adds to the root's histogram with global atomics.  Letters rarely leave a 256 x 32 tile, so the tile
// kernel counts the classes of every tile-local component in shared memory (a table indexed by the rank of the root among
// the tile's root candidates = run starts without foreground above; candidates beyond the table take global atomics),
// writes one histogram per tile-local root and a bit mask of those roots; after the border unions the roots that lost
// their status add their histogram to the component's root (ccv_fold_kernel) -- one update per tile and component.
// ---------------------------------------------------------------------------
constexpr int kVoteCap = 1024, kVoteClasses = 4;

struct VoteTileSmem {
    int lpar[kTileRows * kTileSegs * 32];
    int tab[kVoteClasses][kVoteCap];
    unsigned smask[kTileRows][kTileSegs];
    unsigned cand[256];
    int off[256];
    int wsum[8];
};

__global__ void __launch_bounds__(256)
ccv_tile_kernel(const uint8_t* __restrict__ fg, const uint8_t* __restrict__ pred, int H, int W, int n_classes, int* __restrict__ parent,
                unsigned* __restrict__ rootmask, int* __restrict__ hist /*[px][n_classes]*/) {
    extern __shared__ __align__(16) unsigned char ccv_smem_raw[];
    VoteTileSmem& sm = *reinterpret_cast<VoteTileSmem*>(ccv_smem_raw);
    int* lpar = sm.lpar;
    const int tid = threadIdx.x, lane = tid & 31, sx = tid & (kTileSegs - 1), ry = tid / kTileSegs;
    const int ty0 = blockIdx.y * kTileRows, tx0 = blockIdx.x * kTileSegs * 32;
    const int y = ty0 + ry, x0 = tx0 + sx * 32;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const bool valid = y < H && x0 < W;
    const bool last_row = y == H - 1 && blockIdx.z == gridDim.z - 1;
    const unsigned m = valid ? fg_bits<0>(fg + page_off + (size_t)y * W, x0, W, 0, last_row) : 0u;
    sm.smask[ry][sx] = m;
    for (int k = tid; k < kVoteClasses * kVoteCap; k += 256) (&sm.tab[0][0])[k] = 0;
    const int l0 = tid * 32;
    const unsigned nz = __ballot_sync(0xffffffffu, m != 0);
    for (unsigned todo = nz; todo; todo &= todo - 1) {                  // warp-uniform loop
        const int j = __ffs(todo) - 1;
        const unsigned mj = __shfl_sync(0xffffffffu, m, j);
        if ((mj >> lane) & 1u) {
            const unsigned below = ~mj & ((1u << lane) - 1u);
            const int lb = (tid - lane + j) * 32;
            lpar[lb + lane] = lb + (below ? 32 - __clz(below) : 0);
        }
    }
    __syncthreads();
    const unsigned lb = sx > 0 ? sm.smask[ry][sx - 1] >> 31 : 0u;
    const unsigned up = ry > 0 ? sm.smask[ry - 1][sx] : 0u;
    const unsigned ulb = (ry > 0 && sx > 0) ? sm.smask[ry - 1][sx - 1] >> 31 : 0u;
    const unsigned leftm = (m << 1) | lb, upleftm = (up << 1) | ulb;
    const unsigned cand = m & ~leftm & ~up;                             // run starts (tile sense) without foreground above
    sm.cand[tid] = cand;
    const int cnt = __popc(cand);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.wsum[tid >> 5] = incl;
    if (m) {
        if ((m & 1u) && lb) suf_union_h(lpar, l0, l0 - 1);              // run crosses a segment border
        unsigned v = m & up & ~(leftm & upleftm);
        while (v) {
            const int k = __ffs(v) - 1;
            v &= v - 1;
            suf_union_h(lpar, l0 + k, l0 + k - 256);
        }
    }
    __syncthreads();
    int coff = incl - cnt;
    for (int k = 0; k < (tid >> 5); ++k) coff += sm.wsum[k];
    sm.off[tid] = coff;
    auto gidx = [&](int r) { return (ty0 + (r >> 8)) * W + tx0 + (r & 255); };
    if (coff + cnt > kVoteCap) {                                        // candidates without a table slot: histograms in global memory
        unsigned c2 = cand;
        for (int i = coff; c2; ++i, c2 &= c2 - 1) {
            if (i < kVoteCap) continue;
            int* z = hist + (page_off + gidx(l0 + __ffs(c2) - 1)) * n_classes;
            for (int c = 0; c < n_classes; ++c) z[c] = 0;
        }
    }
    {   // one find per run, kept at the run's first pixel (read-only walks)
        unsigned mm = m;
        int s, len;
        while (next_run(mm, s, len)) lpar[l0 + s] = suf_find(lpar, l0 + s);
    }
    __syncthreads();
    if (m) {
        unsigned w[8];
        load_seg32(pred + page_off + (size_t)y * W, x0, W, last_row, 0xffu, w);
        unsigned cb[kVoteClasses];
#pragma unroll
        for (int c = 0; c < kVoteClasses; ++c) cb[c] = c < n_classes ? seg_bits<1>(w, c) & m : 0u;
        unsigned mm = m;
        int s, len;
        while (next_run(mm, s, len)) {
            const unsigned run = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
            const int r = lpar[l0 + s], rt = r >> 5;
            const int ci = sm.off[rt] + __popc(sm.cand[rt] & ((1u << (r & 31)) - 1u));
#pragma unroll
            for (int c = 0; c < kVoteClasses; ++c) {
                const int v = __popc(cb[c] & run);
                if (!v) continue;
                if (ci < kVoteCap) atomicAdd(&sm.tab[c][ci], v);
                else atomicAdd(&hist[(page_off + gidx(r)) * n_classes + c], v);
            }
        }
    }
    __syncthreads();
    int* par = parent + page_off;
    for (unsigned todo = nz; todo; todo &= todo - 1) {                  // every store writes up to 32 consecutive parents
        const int j = __ffs(todo) - 1;
        const unsigned mj = __shfl_sync(0xffffffffu, m, j);
        const int tj = tid - lane + j;
        const int bj = (ty0 + tj / kTileSegs) * W + tx0 + (tj & (kTileSegs - 1)) * 32;
        if ((mj >> lane) & 1u) {
            const unsigned below = ~mj & ((1u << lane) - 1u);
            par[bj + lane] = gidx(lpar[tj * 32 + (below ? 32 - __clz(below) : 0)]);
        }
    }
    if (valid) {
        unsigned roots = 0u, c2 = cand;
        for (int i = coff; c2; ++i, c2 &= c2 - 1) {
            const int k = __ffs(c2) - 1;
            if (lpar[l0 + k] != l0 + k) continue;
            roots |= 1u << k;
            if (i < kVoteCap) {
                int* g = hist + (page_off + gidx(l0 + k)) * n_classes;
                for (int c = 0; c < n_classes; ++c) g[c] = sm.tab[c][i];
            }
        }
        rootmask[((size_t)blockIdx.z * H + y) * ((W + 31) >> 5) + (x0 >> 5)] = roots;
    }
}

__global__ void __launch_bounds__(256)
ccv_fold_kernel(int H, int W, int n_classes, int* __restrict__ parent, const unsigned* __restrict__ rootmask, int* __restrict__ hist) {
    PCS_SEG_THREAD();
    (void)last_row;
    if (!valid) return;
    unsigned rm = rootmask[(size_t)blockIdx.y * H * segs + t];
    int* par = parent + page_off;
    const int base = y * W + x0;
    while (rm) {
        const int k = __ffs(rm) - 1;
        rm &= rm - 1;
        const int r = uf_find_h(par, base + k);                         // no unions any more: only ancestors are stored
        if (r == base + k) continue;
        const int* a = hist + (page_off + base + k) * n_classes;
        int* g = hist + (page_off + r) * n_classes;
        for (int c = 0; c < n_classes; ++c)
            if (a[c]) atomicAdd(&g[c], a[c]);
    }
}

int launch_cc_majority(pcs_ctx* ctx, uint8_t* d_pred, const uint8_t* d_binary, int n, int H, int W, int n_classes) {
    if (n <= 0 || H <= 0 || W <= 0 || n_classes <= 0 || n_classes > 255 || (size_t)H * W >= (size_t)INT_MAX)
        return set_err(ctx, PCS_ERR_ARG, "cc_majority: bad argument");
    const size_t page_px = (size_t)H * W, total = page_px * n;
    const size_t mask_words = (size_t)n * H * ((W + 31) >> 5);
    PCS_TRY(scratch_reserve(ctx, total * 4 * (1 + (size_t)n_classes) + mask_words * 4 + 256));
    int* parent = reinterpret_cast<int*>(ctx->scratch);
    int* hist = parent + total;
    const dim3 grid = seg_grid(H, W, n);
    static const bool vote_global = [] { const char* e = getenv("PCSEG_VOTE_GLOBAL"); return e && atoi(e) != 0; }();   // A/B switch
    if (n_classes <= kVoteClasses && !vote_global) {
        unsigned* rootmask = reinterpret_cast<unsigned*>(hist + total * n_classes);
        cudaStream_t st = ctx->stream;
        static bool attr_set[64] = {};          // the attribute is per device
        if (ctx->device >= 64 || !attr_set[ctx->device]) {
            PCS_CUDA(ctx, cudaFuncSetAttribute(ccv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoteTileSmem)));
            if (ctx->device < 64) attr_set[ctx->device] = true;
        }
        static const bool poison = [] { const char* e = getenv("PCSEG_CCL_POISON"); return e && e[0] == '1'; }();
        if (poison) PCS_CUDA(ctx, cudaMemsetAsync(parent, 0x7f, (total * (1 + (size_t)n_classes) + mask_words) * 4, st));
        const dim3 gt((W + kTileSegs * 32 - 1) / (kTileSegs * 32), (H + kTileRows - 1) / kTileRows, n);
        ccv_tile_kernel<<<gt, 256, sizeof(VoteTileSmem), st>>>(d_binary, d_pred, H, W, n_classes, parent, rootmask, hist);
        PCS_LAUNCH_CHECK(ctx, "ccv_tile_kernel");
        if (ccl_compress(false)) ccl_border_kernel<0, false, true><<<grid, 256, 0, st>>>(d_binary, H, W, 0, parent);
        else ccl_border_kernel<0, false, false><<<grid, 256, 0, st>>>(d_binary, H, W, 0, parent);
        PCS_LAUNCH_CHECK(ctx, "ccl_border_kernel");
        ccv_fold_kernel<<<grid, 256, 0, st>>>(H, W, n_classes, parent, rootmask, hist);
        PCS_LAUNCH_CHECK(ctx, "ccv_fold_kernel");
        cc_apply_kernel<<<grid, 256, 0, st>>>(d_pred, d_binary, H, W, parent, n_classes, hist);
        PCS_LAUNCH_CHECK(ctx, "cc_apply_kernel");
        return PCS_OK;
    }
    PCS_TRY(ccl_roots(ctx, d_binary, n, H, W, 0, false, parent, hist, n_classes, false, /*fg_only=*/true, /*flatten=*/false));
    cc_vote_kernel<<<grid, 256, 0, ctx->stream>>>(d_pred, d_binary, H, W, parent, n_classes, hist);
    PCS_LAUNCH_CHECK(ctx, "cc_vote_kernel");
    cc_apply_kernel<<<grid, 256, 0, ctx->stream>>>(d_pred, d_binary, H, W, parent, n_classes, hist);
    PCS_LAUNCH_CHECK(ctx, "cc_apply_kernel");
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// add_bounding_boxes (postprocess.py:29-42, evident intent): for c ascending,
// every 4-connected component of (pred == c) paints its bounding box with c.
// Boxes are rasterised through a 2-D difference array + prefix sums.
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
bbox_accum_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, const int* __restrict__ parent, int* __restrict__ box /*[px][4]*/) {
    // one update per RUN: the root is found from the run's first pixel (no flatten pass needed), the run
    // contributes (first x, y, last x, y).  Accumulators are zero-initialised at run starts: keep
    // (W - min x, H - min y, max x, max y) as maxima.  A large component (a picture block, the page background of
    // add_bounding_boxes) would send every run to the same four words: the box is read first and only the atomics
    // that can still grow it are issued (a stale read merely costs an atomic).
    PCS_SEG_THREAD();
    if (!valid) return;
    unsigned mm = fg_bits<MODE>(img + page_off + (size_t)y * W, x0, W, cls, last_row);
    const int base = y * W + x0;
    int s, len;
    while (next_run(mm, s, len)) {
        const int p = uf_find(parent + page_off, base + s);
        int* b = box + (page_off + p) * 4;
        const int v0 = W - (x0 + s), v1 = H - y, v2 = x0 + s + len - 1, v3 = y;
        const int4 cur = __ldcg(reinterpret_cast<const int4*>(b));
        if (v0 > cur.x) atomicMax(&b[0], v0);
        if (v1 > cur.y) atomicMax(&b[1], v1);
        if (v2 > cur.z) atomicMax(&b[2], v2);
        if (v3 > cur.w) atomicMax(&b[3], v3);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256)
bbox_diff_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, const int* __restrict__ parent, const int* __restrict__ box,
                 int* __restrict__ diff) {
    PCS_SEG_THREAD();
    if (!valid) return;
    unsigned mm = fg_bits<MODE>(img + page_off + (size_t)y * W, x0, W, cls, last_row);
    const int base = y * W + x0;
    int* d = diff + (size_t)blockIdx.y * (H + 1) * (W + 1);
    int s, len;
    while (next_run(mm, s, len)) {
        const int idx = base + s;
        if (parent[page_off + idx] != idx) continue;        // roots only (a root is the first pixel of its run)
        const int* b = box + (page_off + idx) * 4;
        const int bx0 = W - b[0], by0 = H - b[1], bx1 = b[2], by1 = b[3];
        atomicAdd(&d[(size_t)by0 * (W + 1) + bx0], 1);
        atomicAdd(&d[(size_t)by0 * (W + 1) + bx1 + 1], -1);
        atomicAdd(&d[(size_t)(by1 + 1) * (W + 1) + bx0], -1);
        atomicAdd(&d[(size_t)(by1 + 1) * (W + 1) + bx1 + 1], 1);
    }
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

// column scan in kColBands row bands (a thread walking all H rows of its column leaves the device idle):
// per-band column sums first, then every band scans its rows from the sum of the bands above it.

__global__ void __launch_bounds__(256)
diff_colband_sum_kernel(const int* __restrict__ diff, int H, int W, int* __restrict__ bandsum) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int rows = (H + kColBands - 1) / kColBands;
    const int y0 = blockIdx.y * rows, y1 = min(H, y0 + rows);
    const int* d = diff + (size_t)blockIdx.z * (H + 1) * (W + 1);
    int acc = 0;
    for (int y = y0; y < y1; ++y) acc += d[(size_t)y * (W + 1) + x];
    bandsum[((size_t)blockIdx.z * kColBands + blockIdx.y) * W + x] = acc;
}

// column scan fused with the paint: coverage > 0 -> out = cls
__global__ void __launch_bounds__(256)
diff_colscan_paint_kernel(const int* __restrict__ diff, int H, int W, int cls, const int* __restrict__ bandsum, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int rows = (H + kColBands - 1) / kColBands;
    const int y0 = blockIdx.y * rows, y1 = min(H, y0 + rows);
    const int* d = diff + (size_t)blockIdx.z * (H + 1) * (W + 1);
    uint8_t* o = out + (size_t)blockIdx.z * H * W;
    int acc = 0;
    for (int b = 0; b < (int)blockIdx.y; ++b) acc += bandsum[((size_t)blockIdx.z * kColBands + b) * W + x];
    for (int y = y0; y < y1; ++y) {
        acc += d[(size_t)y * (W + 1) + x];
        if (acc > 0) o[(size_t)y * W + x] = (uint8_t)cls;
    }
}

int ccl_launch_scan_blocks(pcs_ctx* ctx, int* blocksum, int rows, int nblocks, int* ncomp) {
    ccl_scan_blocks_kernel<<<rows, 1024, 0, ctx->stream>>>(blocksum, nblocks, ncomp);
    PCS_LAUNCH_CHECK(ctx, "ccl_scan_blocks_kernel");
    return PCS_OK;
}

// row scan and per-band column sums of `planes` difference arrays of (H + 1) x (W + 1) ints
int ccl_launch_diff_scans(pcs_ctx* ctx, int* diff, int H, int W, int planes, int* bandsum) {
    diff_rowscan_kernel<<<dim3((H + 1 + 7) / 8, planes), 256, 0, ctx->stream>>>(diff, H + 1, W + 1);
    PCS_LAUNCH_CHECK(ctx, "diff_rowscan_kernel");
    diff_colband_sum_kernel<<<dim3((W + 255) / 256, kColBands, planes), 256, 0, ctx->stream>>>(diff, H, W, bandsum);
    PCS_LAUNCH_CHECK(ctx, "diff_colband_sum_kernel");
    return PCS_OK;
}

int launch_bounding_boxes(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out) {
    if (n <= 0 || H <= 0 || W <= 0 || n_classes <= 0 || n_classes > 255 || (size_t)H * W >= (size_t)INT_MAX)
        return set_err(ctx, PCS_ERR_ARG, "bounding_boxes: bad argument");
    if (n_classes <= kMcMaxClasses && !mc_per_class()) return launch_bounding_boxes_mc(ctx, d_pred, n, H, W, n_classes, d_out);
    const size_t page_px = (size_t)H * W, total = page_px * n;
    const size_t diff_elems = (size_t)n * (H + 1) * (W + 1);
    const size_t total4 = (total + 3) / 4 * 4;                           // keeps the int4 box records 16-byte aligned
    const size_t band_elems = (size_t)n * kColBands * W;
    PCS_TRY(scratch_reserve(ctx, total4 * 4 * 5 + (diff_elems + band_elems) * 4 + 512));
    int* parent = reinterpret_cast<int*>(ctx->scratch);
    int* box = parent + total4;
    int* diff = box + total4 * 4;
    int* bandsum = diff + diff_elems;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(d_out, 0, total, st));                 // newpred = zeros_like(pred)
    // classes = np.unique(pred) per page in the reference; painting an absent class is a no-op
    for (int c = 0; c < n_classes; ++c) {
        PCS_TRY(ccl_roots(ctx, d_pred, n, H, W, c, true, parent, box, 4, false, false, /*flatten=*/false));
        const dim3 g = seg_grid(H, W, n);
        bbox_accum_kernel<1><<<g, 256, 0, st>>>(d_pred, H, W, c, parent, box);
        PCS_LAUNCH_CHECK(ctx, "bbox_accum_kernel");
        PCS_CUDA(ctx, cudaMemsetAsync(diff, 0, diff_elems * 4, st));
        bbox_diff_kernel<1><<<g, 256, 0, st>>>(d_pred, H, W, c, parent, box, diff);
        PCS_LAUNCH_CHECK(ctx, "bbox_diff_kernel");
        diff_rowscan_kernel<<<dim3((H + 1 + 7) / 8, n), 256, 0, st>>>(diff, H + 1, W + 1);
        PCS_LAUNCH_CHECK(ctx, "diff_rowscan_kernel");
        const dim3 gc((W + 255) / 256, kColBands, n);
        diff_colband_sum_kernel<<<gc, 256, 0, st>>>(diff, H, W, bandsum);
        PCS_LAUNCH_CHECK(ctx, "diff_colband_sum_kernel");
        diff_colscan_paint_kernel<<<gc, 256, 0, st>>>(diff, H, W, c, bandsum, d_out);
        PCS_LAUNCH_CHECK(ctx, "diff_colscan_paint_kernel");
    }
    return PCS_OK;
}

// ---------------------------------------------------------------------------
// Segment extraction: the labelling add_bounding_boxes runs per class (postprocess.py:31-33,
// cv2.connectedComponentsWithStats(pred == c, connectivity=4)) with its stats table (what cc.py:4-18 indexes) as the
// result instead of a painted map: per page and class the number of labels (background included, like cv2) and the
// rows [left, top, width, height, area], row 0 = the "background" of that labelling (every pixel != c), rows
// 1.. = the components of class c numbered in raster order of their first pixel.
// One accumulator record per root (W - min x, H - min y, max x, max y, area), one update per RUN.
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
cstats_accum_kernel(const uint8_t* __restrict__ img, int H, int W, int cls, const int* __restrict__ parent, int* __restrict__ acc /*[px][5]*/,
                    int* __restrict__ bg /*[n][5]: W - min x, H - min y, max x, max y, count of the pixels != cls*/) {
    PCS_SEG_THREAD();
    const unsigned m = valid ? fg_bits<MODE>(img + page_off + (size_t)y * W, x0, W, cls, last_row) : 0u;
    {   // pixels outside the class: one warp-aggregated update per warp (a warp never straddles two pages)
        const int nx = valid ? min(32, W - x0) : 0;
        const unsigned b = ~m & (nx >= 32 ? 0xffffffffu : ((1u << nx) - 1u));
        const int v0 = b ? W - (x0 + __ffs(b) - 1) : 0, v1 = b ? H - y : 0, v2 = b ? x0 + 31 - __clz(b) : -1, v3 = b ? y : -1;
        const int r0 = __reduce_max_sync(0xffffffffu, v0), r1 = __reduce_max_sync(0xffffffffu, v1);
        const int r2 = __reduce_max_sync(0xffffffffu, v2), r3 = __reduce_max_sync(0xffffffffu, v3);
        const int cnt = __reduce_add_sync(0xffffffffu, __popc(b));
        if ((threadIdx.x & 31) == 0 && cnt) {
            int* g = bg + (size_t)blockIdx.y * 5;
            if (r0 > g[0]) atomicMax(&g[0], r0);
            if (r1 > g[1]) atomicMax(&g[1], r1);
            if (r2 > g[2]) atomicMax(&g[2], r2);
            if (r3 > g[3]) atomicMax(&g[3], r3);
            atomicAdd(&g[4], cnt);
        }
    }
    // the runs of a warp's 32 segments are walked in lock step and runs that end in the same root are combined before
    // the atomics (a page background or a picture block sends every run of a row to ONE record: without this its five
    // words serialise 30 000 updates per page)
    const int base = y * W + x0;
    unsigned mm = m;
    const unsigned lane = threadIdx.x & 31;
    while (__any_sync(0xffffffffu, mm != 0u)) {
        int s = 0, len = 0;
        const bool has = next_run(mm, s, len);
        const int p = has ? uf_find(parent + page_off, base + s) : -1 - (int)lane;       // idle lanes: unique keys
        const unsigned peers = __match_any_sync(0xffffffffu, p);
        int v0 = W - (x0 + s), v1 = H - y, v2 = x0 + s + len - 1, v3 = y, cnt = len;
        if (peers & (peers - 1)) {                                                       // more than one lane on this root
            v0 = __reduce_max_sync(peers, v0); v1 = __reduce_max_sync(peers, v1);
            v2 = __reduce_max_sync(peers, v2); v3 = __reduce_max_sync(peers, v3);
            cnt = __reduce_add_sync(peers, cnt);
        }
        if (has && lane == (unsigned)(__ffs(peers) - 1)) {
            int* a = acc + (page_off + p) * 5;
            if (v0 > __ldcg(a + 0)) atomicMax(&a[0], v0);
            if (v1 > __ldcg(a + 1)) atomicMax(&a[1], v1);
            if (v2 > __ldcg(a + 2)) atomicMax(&a[2], v2);
            if (v3 > __ldcg(a + 3)) atomicMax(&a[3], v3);
            atomicAdd(&a[4], cnt);
        }
    }
}

// stats row of every root (label = rank of the root) and row 0; rows at or beyond the label count stay zero
__global__ void __launch_bounds__(256)
cstats_write_kernel(const int* __restrict__ parent, const int* __restrict__ rank, const int* __restrict__ acc, const int* __restrict__ bg,
                    int H, int W, int32_t* __restrict__ stats, size_t page_stride /*ints between pages*/, int max_components) {
    const size_t page_px = (size_t)H * W, page_off = (size_t)blockIdx.y * page_px;
    int32_t* st = stats + (size_t)blockIdx.y * page_stride;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int* g = bg + (size_t)blockIdx.y * 5;
        if (g[4] > 0) { st[0] = W - g[0]; st[1] = H - g[1]; st[2] = g[2] - (W - g[0]) + 1; st[3] = g[3] - (H - g[1]) + 1; st[4] = g[4]; }
        else { st[0] = 0; st[1] = 0; st[2] = 0; st[3] = 0; st[4] = 0; }     // cv2 leaves an empty label's box at zero extent
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        if (parent[page_off + i] != (int)i) continue;
        const int l = rank[page_off + i];
        if (l >= max_components) continue;
        const int* a = acc + (page_off + i) * 5;
        int32_t* o = st + (size_t)l * 5;
        const int left = W - a[0], top = H - a[1];
        o[0] = left; o[1] = top; o[2] = a[2] - left + 1; o[3] = a[3] - top + 1; o[4] = a[4];
    }
}

__global__ void cstats_ncomp_kernel(const int* __restrict__ ncomp_tmp, int n, int n_classes, int cls, int32_t* __restrict__ d_ncomp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d_ncomp[(size_t)i * n_classes + cls] = ncomp_tmp[i];
}

int launch_class_components(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, int32_t* d_stats, int max_components,
                            int32_t* d_ncomp) {
    if (n <= 0 || H <= 0 || W <= 0 || n_classes <= 0 || n_classes > 255 || (size_t)H * W >= (size_t)INT_MAX || max_components <= 0)
        return set_err(ctx, PCS_ERR_ARG, "class_components: bad argument");
    if (n_classes <= kMcMaxClasses && !mc_per_class())
        return launch_class_components_mc(ctx, d_pred, n, H, W, n_classes, d_stats, max_components, d_ncomp);
    const size_t page_px = (size_t)H * W, total = page_px * n;
    const int nblocks = (int)((page_px + kScanBlock - 1) / kScanBlock);
    PCS_TRY(scratch_reserve(ctx, total * 4 * 7 + ((size_t)n * nblocks + n + (size_t)n * 5) * 4 + 512));
    int* parent = reinterpret_cast<int*>(ctx->scratch);
    int* rank = parent + total;
    int* acc = rank + total;
    int* blocksum = acc + total * 5;
    int* ncomp_tmp = blocksum + (size_t)n * nblocks;
    int* bg = ncomp_tmp + n;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(d_stats, 0, (size_t)n * n_classes * max_components * 5 * sizeof(int32_t), st));
    const dim3 g = seg_grid(H, W, n);
    for (int c = 0; c < n_classes; ++c) {
        PCS_TRY(ccl_roots(ctx, d_pred, n, H, W, c, true, parent, acc, 5, false, false, /*flatten=*/false));
        PCS_CUDA(ctx, cudaMemsetAsync(bg, 0, (size_t)n * 5 * sizeof(int), st));
        cstats_accum_kernel<1><<<g, 256, 0, st>>>(d_pred, H, W, c, parent, acc, bg);
        PCS_LAUNCH_CHECK(ctx, "cstats_accum_kernel");
        ccl_count_roots_kernel<<<dim3(nblocks, n), 256, 0, st>>>(parent, page_px, nblocks, blocksum);
        PCS_LAUNCH_CHECK(ctx, "ccl_count_roots_kernel");
        ccl_scan_blocks_kernel<<<n, 1024, 0, st>>>(blocksum, nblocks, ncomp_tmp);
        PCS_LAUNCH_CHECK(ctx, "ccl_scan_blocks_kernel");
        ccl_rank_kernel<<<dim3(nblocks, n), 256, 0, st>>>(parent, page_px, nblocks, blocksum, rank);
        PCS_LAUNCH_CHECK(ctx, "ccl_rank_kernel");
        cstats_write_kernel<<<dim3((unsigned)std::min<size_t>(512, (page_px + 255) / 256), n), 256, 0, st>>>(
            parent, rank, acc, bg, H, W, d_stats + (size_t)c * max_components * 5, (size_t)n_classes * max_components * 5, max_components);
        PCS_LAUNCH_CHECK(ctx, "cstats_write_kernel");
        if (d_ncomp) {
            cstats_ncomp_kernel<<<(n + 255) / 256, 256, 0, st>>>(ncomp_tmp, n, n_classes, c, d_ncomp);
            PCS_LAUNCH_CHECK(ctx, "cstats_ncomp_kernel");
        }
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
    const bool aligned = ((reinterpret_cast<uintptr_t>(img + off) | reinterpret_cast<uintptr_t>(fg + off)) & 15) == 0;
    const size_t chunks = aligned ? page_px / 16 : 0;
    const unsigned t4 = (unsigned)t * 0x01010101u, flip = inverse ? 0u : 0xffffffffu;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(img + off) + i);
        uint4 r;
        r.x = (__vcmpgtu4(q.x, t4) ^ flip) & 0x01010101u;
        r.y = (__vcmpgtu4(q.y, t4) ^ flip) & 0x01010101u;
        r.z = (__vcmpgtu4(q.z, t4) ^ flip) & 0x01010101u;
        r.w = (__vcmpgtu4(q.w, t4) ^ flip) & 0x01010101u;
        reinterpret_cast<uint4*>(fg + off)[i] = r;
    }
    for (size_t i = chunks * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < page_px; i += (size_t)gridDim.x * blockDim.x) {
        const bool above = img[off + i] > t;
        fg[off + i] = (above == (inverse != 0)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
letter_heights_kernel(const uint8_t* __restrict__ fg, int H, int W, const int* __restrict__ parent, const int* __restrict__ box,
                      unsigned* __restrict__ hh /*[n][64]*/) {
    PCS_SEG_THREAD();
    if (!valid) return;
    unsigned mm = fg_bits<0>(fg + page_off + (size_t)y * W, x0, W, 0, last_row);
    const int base = y * W + x0;
    int s, len;
    while (next_run(mm, s, len)) {
        const int idx = base + s;
        if (parent[page_off + idx] != idx) continue;            // roots only (a root is the first pixel of its run)
        const int* b = box + (page_off + idx) * 4;
        const int w = b[2] - (W - b[0]) + 1, h = b[3] - (H - b[1]) + 1;
        // 0.5 < w/h < 2  <=>  h < 2w and w < 2h (exact for these small integers)
        if (h < 2 * w && w < 2 * h && h > 10 && h < 60 && w > 5 && w < 50) atomicAdd(&hh[(size_t)blockIdx.y * 64 + h], 1u);
    }
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
    PCS_TRY(ccl_roots(ctx, fg, n, H, W, 0, false, parent, box, 4, /*conn8=*/true, /*fg_only=*/true, /*flatten=*/false));
    const dim3 g = seg_grid(H, W, n);
    bbox_accum_kernel<0><<<g, 256, 0, st>>>(fg, H, W, 0, parent, box);
    PCS_LAUNCH_CHECK(ctx, "bbox_accum_kernel");
    letter_heights_kernel<<<g, 256, 0, st>>>(fg, H, W, parent, box, hh);
    PCS_LAUNCH_CHECK(ctx, "letter_heights_kernel");
    median_height_kernel<<<(n + 63) / 64, 64, 0, st>>>(hh, n, d_out);
    PCS_LAUNCH_CHECK(ctx, "median_height_kernel");
    return PCS_OK;
}

}  // namespace pcs
