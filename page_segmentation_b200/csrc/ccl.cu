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

#include <climits>
#include <cstdlib>
#include <cstring>

namespace pcs {

constexpr int kBG = INT_MIN;

__device__ __forceinline__ int uf_find(const int* parent, int x) {
    // L2 loads: other SMs re-link nodes concurrently; a stale value would still be a valid
    // ancestor, but reading through L2 keeps the retry count low
    int p = __ldcg(parent + x);
    while (p != x) { x = p; p = __ldcg(parent + x); }
    return x;
}

// COMPRESS: the two starting nodes are re-linked to the roots found on the first walk (atomicMin towards an ancestor
// of the same tree: pointers only ever decrease, so no cycle can form and no link between two sets is lost)
template <bool COMPRESS>
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    if (COMPRESS) {
        const int ra = uf_find(parent, a), rb = uf_find(parent, b);
        if (ra < a) atomicMin(&parent[a], ra);
        if (rb < b) atomicMin(&parent[b], rb);
        if (ra == rb) return;
        a = ra; b = rb;
    }
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

// the same walk with path halving (see suf_find_h below: every store puts an ANCESTOR into a node that is not a root)
__device__ __forceinline__ int uf_find_h(int* parent, int x) {
    int p = __ldcg(parent + x);
    while (p != x) {
        const int g = __ldcg(parent + p);
        if (g == p) return p;
        __stcg(parent + x, g);
        x = g;
        p = __ldcg(parent + x);
    }
    return x;
}

// union with halving walks; the starting nodes end up one or two links from their roots, so no separate re-linking
__device__ __forceinline__ void uf_union_h(int* parent, int a, int b) {
    while (true) {
        a = uf_find_h(parent, a);
        b = uf_find_h(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// ---------------------------------------------------------------------------------------------------
// Segment-wise passes.  A page is nine tenths background, so the passes over the image do not spend a thread per
// pixel: a thread owns one 32-pixel segment of a row, reads it with two 128-bit loads (32-bit words and a funnel
// shift when the row is not 16-byte aligned), turns it into a bit mask and is done when the mask is empty;
// neighbour relations (left / upper / diagonal) are bit operations between the masks of two rows.
// ---------------------------------------------------------------------------------------------------
// MODE 0: foreground = non-zero byte; MODE 1: foreground = byte equal to `cls`
template <int MODE>
__device__ __forceinline__ unsigned bytes4_to_bits(unsigned w, unsigned cls4) {
    unsigned c = MODE == 1 ? __vcmpeq4(w, cls4) : __vcmpne4(w, 0u);     // 0xff per matching byte
    c &= 0x08040201u;                                                   // byte k keeps bit k of itself
    c |= c >> 8;
    c |= c >> 16;
    return c & 0xfu;
}

// the 32 bytes of a segment as eight words; bytes at or beyond W read as `fill`.
// `last_row`: the row is the last one of the whole buffer (the word path may read 3 bytes past pixel x0 + 31).
__device__ __forceinline__ void load_seg32(const uint8_t* __restrict__ row, int x0, int W, bool last_row, unsigned fill, unsigned (&w)[8]) {
    const uint8_t* p = row + x0;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if (x0 + 32 <= W && (a & 15) == 0) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w; w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
    } else if (x0 + 32 <= W && (!last_row || x0 + 36 <= W)) {
        const unsigned* pa = reinterpret_cast<const unsigned*>(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8;
        unsigned lo = __ldg(pa);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned hi = (k < 7 || sh) ? __ldg(pa + k + 1) : 0u;
            w[k] = __funnelshift_r(lo, hi, sh);
            lo = hi;
        }
    } else {
        const int nx = min(32, W - x0);
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0u;
#pragma unroll
        for (int k = 0; k < 32; ++k) w[k >> 2] |= (unsigned)(k < nx ? __ldg(p + k) : (uint8_t)fill) << ((k & 3) * 8);
    }
}

template <int MODE>
__device__ __forceinline__ unsigned seg_bits(const unsigned (&w)[8], int cls) {
    const unsigned cls4 = (unsigned)cls * 0x01010101u;
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) m |= bytes4_to_bits<MODE>(w[k], cls4) << (4 * k);
    return m;
}

// bit k = pixel x0 + k of this row is foreground; pixels at or beyond W read as background.
template <int MODE>
__device__ __forceinline__ unsigned fg_bits(const uint8_t* __restrict__ row, int x0, int W, int cls, bool last_row) {
    unsigned w[8];
    load_seg32(row, x0, W, last_row, MODE == 1 ? ~(unsigned)cls : 0u, w);
    return seg_bits<MODE>(w, cls);
}

// pops the lowest run of set bits of mm: pixels s .. s + len - 1
__device__ __forceinline__ bool next_run(unsigned& mm, int& s, int& len) {
    if (!mm) return false;
    s = __ffs(mm) - 1;
    len = __ffs(~(mm >> s)) - 1;
    if (len < 0) len = 32;                                              // s == 0 and all 32 pixels set
    mm &= ~((len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s);
    return true;
}

// thread -> (row y, first pixel x0 of its segment); grid = (ceil(segs * H / 256), n pages)
#define PCS_SEG_THREAD()                                                     \
    const int segs = (W + 31) >> 5;                                          \
    const int t = blockIdx.x * 256 + threadIdx.x;                            \
    const bool valid = t < segs * H;                                         \
    const int y = valid ? t / segs : 0;                                      \
    const int x0 = valid ? (t - y * segs) * 32 : 0;                          \
    const size_t page_off = (size_t)blockIdx.y * H * W;                      \
    const bool last_row = y == H - 1 && blockIdx.y == gridDim.y - 1

static inline dim3 seg_grid(int H, int W, int n) { return dim3((unsigned)((((size_t)(W + 31) / 32) * H + 255) / 256), n); }

// union-find inside one tile, in shared memory (local pixel index = thread * 32 + bit)
__device__ __forceinline__ int suf_find(const volatile int* lp, int x) {
    int p = lp[x];
    while (p != x) { x = p; p = lp[x]; }
    return x;
}

__device__ __forceinline__ void suf_union(int* lp, int a, int b) {
    while (true) {
        a = suf_find(lp, a);
        b = suf_find(lp, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(&lp[a], b);
        if (old == a) return;
        a = old;
    }
}

// path halving: every node on the walk is re-linked to its grandparent (plain stores of an ANCESTOR into a node that
// is not a root: roots change through atomicMin only, and a node that stopped being a root never becomes one again)
__device__ __forceinline__ int suf_find_h(volatile int* lp, int x) {
    int p = lp[x];
    while (p != x) {
        const int g = lp[p];
        if (g == p) return p;
        lp[x] = g;
        x = g;
        p = lp[x];
    }
    return x;
}

__device__ __forceinline__ void suf_union_h(int* lp, int a, int b) {
    while (true) {
        a = suf_find_h(lp, a);
        b = suf_find_h(lp, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(&lp[a], b);
        if (old == a) return;
        a = old;
    }
}

constexpr int kTileSegs = 8, kTileRows = 32;        // a block labels a tile of 256 x 32 pixels

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

// Re-linking of union operands: measured on 32 A4 pages it pays where components are huge (the page background of
// add_bounding_boxes: 1.91 -> 1.80 ms) and costs where they are letters (compute_char_height 2.62 -> 2.84 ms,
// cc_majority 0.44 -> 0.46 ms), so it is on for class-match labelling only.  PCSEG_CCL_COMPRESS=0/1 forces it.
static bool ccl_compress(bool match) {
    static const char* e = getenv("PCSEG_CCL_COMPRESS");
    return e ? atoi(e) != 0 : match;
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
        static bool attr_set = false;
        if (!attr_set) {
            PCS_CUDA(ctx, cudaFuncSetAttribute(ccv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoteTileSmem)));
            attr_set = true;
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
constexpr int kColBands = 16;
constexpr int kMcMaxClasses = 8;                    // classes the one-pass labelling (mc_*, below) handles
static int launch_bounding_boxes_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out);
static bool mc_per_class() {                        // A/B switch: one labelling per class instead of one for all
    static const char* e = getenv("PCSEG_SEGMENTS_PER_CLASS");
    return e && atoi(e) != 0;
}

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

// ---------------------------------------------------------------------------
// Segment extraction, every class in ONE labelling (n_classes <= kMcMaxClasses).  The per-class path above labels
// (pred == c) once per class: three tile / border / accumulate / rank rounds per page, each reading the whole class
// map, and one set of five global atomics per run.  Every pixel belongs to exactly one of those labellings, so the
// class map is labelled once with "same byte" as the neighbour relation:
//   mc_tile   : 256 x 32 tiles in shared memory as in ccl_tile_kernel, the run structure taken from byte compares
//               (run start = differs from the left pixel, vertical union = equals the upper pixel where either row
//               starts a run); the statistics of every tile-local component (area, x extent, row mask) are gathered
//               with shared-memory atomics in a table indexed by the rank of the root among the tile's root
//               candidates (run starts without an equal upper pixel; candidates beyond the table go to global
//               atomics), and written once per tile-local root together with a bit mask of those roots;
//   mc_border : unions across tile borders on the global parents;
//   mc_fold   : tile-local roots that lost their root status add their record to the component's root (one set of
//               atomics per tile and component, not per run); roots counted per class and warp (= 1024 pixels in
//               raster order), box and pixel count of every class for row 0 of the tables;
//   scan, mc_write : label = rank of the root among the roots of ITS class (cv2's numbering of that class's
//               labelling), stats row written from the root's record.
// ---------------------------------------------------------------------------
// tile-local root candidates with shared-memory accumulators: 512 records leave room for five tiles per SM (1 024: four)

__device__ __forceinline__ unsigned eq_bits32(const unsigned (&a)[8], const unsigned (&b)[8]) {
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned c = __vcmpeq4(a[k], b[k]) & 0x08040201u;
        c |= c >> 8;
        c |= c >> 16;
        m |= (c & 0xfu) << (4 * k);
    }
    return m;
}

// bit k = pixel k equals its left neighbour; the neighbour of pixel 0 is the low byte of `prev`
__device__ __forceinline__ unsigned eq_left_bits32(const unsigned (&w)[8], unsigned prev) {
    unsigned s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = (w[k] << 8) | (k ? w[k - 1] >> 24 : (prev & 0xffu));
    return eq_bits32(w, s);
}

template <int kMcCap> struct McTileSmem {
    int lpar[kTileRows * kTileSegs * 32];
    union {
        uint4 bytes[256][2];                        // the tile's class bytes (phase 1 only)
        int tab[4][kMcCap];                         // area, 255 - min x, max x, row mask of the tile-local components
    } u;
    unsigned hs[256];                               // run starts (tile sense) of every segment
    unsigned cand[256];                             // root candidates of every segment
    int off[256];                                   // exclusive scan of the candidate counts
    int wsum[8];
};

template <int kMcCap>
__global__ void __launch_bounds__(256)
mc_tile_kernel(const uint8_t* __restrict__ img, int H, int W, int* __restrict__ parent, unsigned* __restrict__ rootmask,
               int* __restrict__ acc /*[px][5]: W - min x, H - min y, max x, max y, area*/) {
    extern __shared__ __align__(16) unsigned char mc_smem_raw[];
    McTileSmem<kMcCap>& sm = *reinterpret_cast<McTileSmem<kMcCap>*>(mc_smem_raw);
    int* lpar = sm.lpar;
    const int tid = threadIdx.x, lane = tid & 31, sx = tid & (kTileSegs - 1), ry = tid / kTileSegs;
    const int ty0 = blockIdx.y * kTileRows, tx0 = blockIdx.x * kTileSegs * 32;
    const int y = ty0 + ry, x0 = tx0 + sx * 32;
    const size_t page_off = (size_t)blockIdx.z * H * W;
    const bool valid = y < H && x0 < W;
    const bool last_row = y == H - 1 && blockIdx.z == gridDim.z - 1;
    const int nx = valid ? min(32, W - x0) : 0;
    const unsigned m = nx >= 32 ? 0xffffffffu : ((1u << nx) - 1u);     // pixels of the segment inside the page
    unsigned w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (valid) load_seg32(img + page_off + (size_t)y * W, x0, W, last_row, 0u, w);
    sm.u.bytes[tid][0] = make_uint4(w[0], w[1], w[2], w[3]);
    sm.u.bytes[tid][1] = make_uint4(w[4], w[5], w[6], w[7]);
    __syncthreads();
    const int l0 = tid * 32;
    unsigned eql = eq_left_bits32(w, sx > 0 ? sm.u.bytes[tid - 1][1].w >> 24 : 0u);
    if (sx == 0) eql &= ~1u;                                            // the tile's first column starts a run
    const unsigned hs_true = ~eql & m;                                  // run starts, tile sense
    const unsigned hs_seg = (hs_true | 1u) & m;                         // run starts, segment sense
    const bool cont0 = (eql & m & 1u) != 0u;                            // the first run continues the left segment's last
    unsigned vsame = 0u;
    if (ry > 0 && valid) {
        const uint4 u0 = sm.u.bytes[tid - kTileSegs][0], u1 = sm.u.bytes[tid - kTileSegs][1];
        const unsigned wu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        vsame = eq_bits32(w, wu) & m;
    }
    const unsigned cand = hs_true & ~vsame;
    sm.hs[tid] = hs_true;
    sm.cand[tid] = cand;
    const unsigned nzv = __ballot_sync(0xffffffffu, m != 0u);
    for (unsigned todo = nzv; todo; todo &= todo - 1) {                 // warp-uniform loop
        const int j = __ffs(todo) - 1;
        const unsigned hj = __shfl_sync(0xffffffffu, hs_seg, j), mj = __shfl_sync(0xffffffffu, m, j);
        if ((mj >> lane) & 1u) {
            const int lb = (tid - lane + j) * 32;
            lpar[lb + lane] = lb + 31 - __clz(hj & (0xffffffffu >> (31 - lane)));
        }
    }
    int cnt = __popc(cand), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.wsum[tid >> 5] = incl;
    __syncthreads();
    if (cont0) suf_union_h(lpar, l0, l0 - 1);
    if (vsame) {
        unsigned v = vsame & (hs_true | sm.hs[tid - kTileSegs]);
        while (v) {
            const int k = __ffs(v) - 1;
            v &= v - 1;
            suf_union_h(lpar, l0 + k, l0 + k - 256);
        }
    }
    int coff = incl - cnt;
    for (int k = 0; k < (tid >> 5); ++k) coff += sm.wsum[k];
    sm.off[tid] = coff;
    for (int k = tid; k < 4 * kMcCap; k += 256) (&sm.u.tab[0][0])[k] = 0;   // the byte copy was last read before the barrier
    auto gidx = [&](int r) { return (ty0 + (r >> 8)) * W + tx0 + (r & 255); };
    if (coff + cnt > kMcCap) {                                          // candidates without a table slot: records in global memory
        unsigned c2 = cand;
        for (int i = coff; c2; ++i, c2 &= c2 - 1) {
            if (i < kMcCap) continue;
            int* z = acc + (page_off + gidx(l0 + __ffs(c2) - 1)) * 5;
            z[0] = 0; z[1] = 0; z[2] = 0; z[3] = 0; z[4] = 0;
        }
    }
    __syncthreads();
    {   // one find per run, kept at the run's first pixel (a concurrent walker reads the old or the new ancestor)
        unsigned mm = hs_seg;
        while (mm) {
            const int s = __ffs(mm) - 1;
            mm &= mm - 1;
            lpar[l0 + s] = suf_find(lpar, l0 + s);                     // read-only walk: a halving store could land after another thread's final one
        }
    }
    __syncthreads();
    {   // statistics.  A thread adds up the runs that share the root of its longest run in registers (the page
        // background or a picture block owns most segments of a tile: its record would serialise thousands of
        // atomics), those sums are combined across the warp; the other runs (letters, specks) update their record.
        auto update = [&](int r, int a, int v0, int v2, unsigned rows) {
            const int rt = r >> 5;
            const int ci = sm.off[rt] + __popc(sm.cand[rt] & ((1u << (r & 31)) - 1u));
            if (ci < kMcCap) {
                atomicAdd(&sm.u.tab[0][ci], a);
                atomicMax(&sm.u.tab[1][ci], v0);
                atomicMax(&sm.u.tab[2][ci], v2);
                atomicOr(reinterpret_cast<unsigned*>(&sm.u.tab[3][ci]), rows);
            } else {
                int* g = acc + (page_off + gidx(r)) * 5;
                atomicMax(&g[0], W - (tx0 + 255 - v0));
                atomicMax(&g[1], H - (ty0 + __ffs(rows) - 1));
                atomicMax(&g[2], tx0 + v2);
                atomicMax(&g[3], ty0 + 31 - __clz(rows));
                atomicAdd(&g[4], a);
            }
        };
        int best_s = 0, best_len = 0;
        for (unsigned mm = hs_seg; mm;) {
            const int s = __ffs(mm) - 1;
            mm &= mm - 1;
            const int len = (mm ? __ffs(mm) - 1 : nx) - s;
            if (len > best_len) { best_len = len; best_s = s; }
        }
        const int rkey = valid ? lpar[l0 + best_s] : -1 - lane;         // idle lanes: unique keys
        int a = 0, v0 = 0, v2 = -1;
        for (unsigned mm = hs_seg; mm;) {
            const int s = __ffs(mm) - 1;
            mm &= mm - 1;
            const int e = mm ? __ffs(mm) - 1 : nx;
            const int r = s == best_s ? rkey : lpar[l0 + s];
            if (r == rkey) { a += e - s; v0 = max(v0, 255 - (sx * 32 + s)); v2 = max(v2, sx * 32 + e - 1); }
            else update(r, e - s, 255 - (sx * 32 + s), sx * 32 + e - 1, 1u << ry);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, rkey);
        unsigned rows = 1u << ry;
        if (peers & (peers - 1)) {
            a = __reduce_add_sync(peers, a);
            v0 = __reduce_max_sync(peers, v0);
            v2 = __reduce_max_sync(peers, v2);
            rows = __reduce_or_sync(peers, rows);
        }
        if (valid && lane == __ffs(peers) - 1) update(rkey, a, v0, v2, rows);
    }
    __syncthreads();
    int* par = parent + page_off;
    // Parents are written where later passes read them: the tile's first and last row and its first and last column
    // (operands of the border unions) and the tile-local roots (the ends of those walks, the pixels mc_fold visits).
    const unsigned edge_rows = __ballot_sync(0xffffffffu, m != 0u && (ry == 0 || ry == kTileRows - 1));
    for (unsigned todo = edge_rows; todo; todo &= todo - 1) {           // every store writes 32 consecutive parents
        const int j = __ffs(todo) - 1;
        const unsigned hj = __shfl_sync(0xffffffffu, hs_seg, j), mj = __shfl_sync(0xffffffffu, m, j);
        const int tj = tid - lane + j;
        const int bj = (ty0 + tj / kTileSegs) * W + tx0 + (tj & (kTileSegs - 1)) * 32;
        if ((mj >> lane) & 1u) par[bj + lane] = gidx(lpar[tj * 32 + 31 - __clz(hj & (0xffffffffu >> (31 - lane)))]);
    }
    if (valid && ry != 0 && ry != kTileRows - 1) {
        if (sx == 0) par[y * W + x0] = gidx(lpar[l0]);
        if (sx == kTileSegs - 1 && nx == 32) par[y * W + x0 + 31] = gidx(lpar[l0 + 31 - __clz(hs_seg)]);
    }
    if (valid) {
        unsigned roots = 0u, c2 = cand;
        for (int i = coff; c2; ++i, c2 &= c2 - 1) {
            const int k = __ffs(c2) - 1;
            if (lpar[l0 + k] != l0 + k) continue;
            roots |= 1u << k;
            par[y * W + x0 + k] = y * W + x0 + k;
            if (i < kMcCap) {
                const unsigned rows = (unsigned)sm.u.tab[3][i];
                int* g = acc + (page_off + gidx(l0 + k)) * 5;
                g[0] = W - (tx0 + 255 - sm.u.tab[1][i]);
                g[1] = H - (ty0 + __ffs(rows) - 1);
                g[2] = tx0 + sm.u.tab[2][i];
                g[3] = ty0 + 31 - __clz(rows);
                g[4] = sm.u.tab[0][i];
            }
        }
        rootmask[((size_t)blockIdx.z * H + y) * ((W + 31) >> 5) + (x0 >> 5)] = roots;
    }
}

// Unions across tile borders.  The walks to the roots are chains of dependent L2 round trips, so the work is spread as
// thinly as possible: one LANE per pixel of a row that starts a tile (a thread per 32-pixel segment spent up to 32
// unions one after the other: 210 us per launch on noise-like maps whatever the number of pages), one thread per row
// and vertical tile border.
template <bool HALVE>
__global__ void __launch_bounds__(256)
mc_border_rows_kernel(const uint8_t* __restrict__ img, int H, int W, int* __restrict__ parent) {
    const int segs = (W + 31) >> 5;
    const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;     // warp -> (tile row, segment)
    const int tr = wid / segs + 1, x = (wid - (tr - 1) * segs) * 32 + lane, y = tr * kTileRows;
    if (y >= H || x >= W) return;
    const size_t page_off = (size_t)blockIdx.y * H * W;
    const uint8_t* row = img + page_off + (size_t)y * W;
    const uint8_t me = __ldg(row + x), up = __ldg(row - W + x);
    if (me != up) return;
    if (x > 0 && __ldg(row + x - 1) == me && __ldg(row - W + x - 1) == up) return;  // the pixel to the left unites the same two runs
    int* par = parent + page_off;
    if (HALVE) uf_union_h(par, y * W + x, (y - 1) * W + x);
    else uf_union<false>(par, y * W + x, (y - 1) * W + x);
}

template <bool HALVE>
__global__ void __launch_bounds__(256)
mc_border_cols_kernel(const uint8_t* __restrict__ img, int H, int W, int* __restrict__ parent) {
    const int ncols = (W - 1) / (kTileSegs * 32);                        // vertical tile borders inside the page
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= ncols * H) return;
    const int y = t / ncols, x = (t - y * ncols + 1) * kTileSegs * 32;
    const size_t page_off = (size_t)blockIdx.y * H * W;
    const uint8_t* row = img + page_off + (size_t)y * W;
    if (__ldg(row + x) != __ldg(row + x - 1)) return;
    if (HALVE) uf_union_h(parent + page_off, y * W + x, y * W + x - 1);
    else uf_union<false>(parent + page_off, y * W + x, y * W + x - 1);
}

// clsbox: [page][n_classes + 1][5] = W - min x, H - min y, max x, max y, pixel count of every class (last: other bytes)
__global__ void __launch_bounds__(256)
mc_fold_kernel(const uint8_t* __restrict__ img, int H, int W, int n_classes, int* __restrict__ parent, unsigned* __restrict__ rootmask,
               int* __restrict__ acc, int* __restrict__ warpcnt /*[page][class][warp]*/, int* __restrict__ clsbox) {
    PCS_SEG_THREAD();
    __shared__ int sbox[(kMcMaxClasses + 1) * 5];
    if (threadIdx.x < (kMcMaxClasses + 1) * 5) sbox[threadIdx.x] = 0;
    __syncthreads();
    const size_t mask_off = (size_t)blockIdx.y * H * segs + t;
    const unsigned rm = valid ? rootmask[mask_off] : 0u;
    int* par = parent + page_off;
    const int base = y * W + x0;
    // The tile roots of a warp's 32 segments are dealt out evenly over its lanes: a segment on a tile border of a noisy map
    // holds a dozen roots that lost their status, each a walk plus a record merge (dependent L2 round trips), its
    // neighbours none -- one thread doing them in turn set the duration of the whole launch.
    __shared__ unsigned short s_list[8][1024];
    __shared__ unsigned s_gm[8][32];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s_gm[wib][lane] = 0u;
    const int cnt = __popc(rm);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    {
        unsigned r2 = rm;
        for (int i = incl - cnt; r2; ++i, r2 &= r2 - 1) s_list[wib][i] = (unsigned short)((lane << 5) | (__ffs(r2) - 1));
    }
    __syncwarp();
    for (int i0 = 0; i0 < total; i0 += 32) {
        const bool has = i0 + lane < total;
        const int e = has ? s_list[wib][i0 + lane] : 0;
        const int k = e & 31, p = __shfl_sync(0xffffffffu, base, e >> 5) + k;
        if (!has) continue;
        const int r = uf_find_h(par, p);                                // no unions any more: only ancestors are stored
        if (r == p) { atomicOr(&s_gm[wib][e >> 5], 1u << k); continue; }
        const int* a = acc + (page_off + p) * 5;
        int* g = acc + (page_off + r) * 5;
        const int v0 = a[0], v1 = a[1], v2 = a[2], v3 = a[3];
        if (v0 > __ldcg(g + 0)) atomicMax(&g[0], v0);
        if (v1 > __ldcg(g + 1)) atomicMax(&g[1], v1);
        if (v2 > __ldcg(g + 2)) atomicMax(&g[2], v2);
        if (v3 > __ldcg(g + 3)) atomicMax(&g[3], v3);
        atomicAdd(&g[4], a[4]);
    }
    __syncwarp();
    const unsigned gm = s_gm[wib][lane];
    if (valid) rootmask[mask_off] = gm;                                 // from here on: the roots of whole components
    const int nx = valid ? min(32, W - x0) : 0;
    const unsigned m = nx >= 32 ? 0xffffffffu : ((1u << nx) - 1u);
    unsigned w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (valid) load_seg32(img + page_off + (size_t)y * W, x0, W, last_row, 0u, w);
    const int nw = gridDim.x * 8, wid = t >> 5;
    unsigned seen = 0u;
    for (int c = 0; c <= n_classes; ++c) {
        const unsigned b = c < n_classes ? seg_bits<1>(w, c) & m : m & ~seen;
        seen |= b;
        if (c < n_classes) {
            const int wc = __reduce_add_sync(0xffffffffu, __popc(gm & b));
            if ((threadIdx.x & 31) == 0) warpcnt[((size_t)blockIdx.y * n_classes + c) * nw + wid] = wc;
        }
        if (!__any_sync(0xffffffffu, b != 0u)) continue;
        const int v0 = b ? W - (x0 + __ffs(b) - 1) : 0, v1 = b ? H - y : 0, v2 = b ? x0 + 31 - __clz(b) : -1, v3 = b ? y : -1;
        const int r0 = __reduce_max_sync(0xffffffffu, v0), r1 = __reduce_max_sync(0xffffffffu, v1);
        const int r2 = __reduce_max_sync(0xffffffffu, v2), r3 = __reduce_max_sync(0xffffffffu, v3);
        const int pc = __reduce_add_sync(0xffffffffu, __popc(b));
        if ((threadIdx.x & 31) == 0) {
            int* s = sbox + c * 5;
            if (r0 > s[0]) atomicMax(&s[0], r0);
            if (r1 > s[1]) atomicMax(&s[1], r1);
            if (r2 > s[2]) atomicMax(&s[2], r2);
            if (r3 > s[3]) atomicMax(&s[3], r3);
            atomicAdd(&s[4], pc);
        }
    }
    __syncthreads();
    if (threadIdx.x < (n_classes + 1) * 5) {
        const int c = threadIdx.x / 5, f = threadIdx.x - c * 5;
        if (sbox[c * 5 + 4] > 0) {
            int* g = clsbox + ((size_t)blockIdx.y * (n_classes + 1) + c) * 5 + f;
            if (f == 4) atomicAdd(g, sbox[threadIdx.x]);
            else if (sbox[threadIdx.x] > __ldcg(g)) atomicMax(g, sbox[threadIdx.x]);
        }
    }
}

__global__ void __launch_bounds__(256)
mc_write_kernel(const uint8_t* __restrict__ img, int H, int W, int n_classes, const unsigned* __restrict__ rootmask, const int* __restrict__ acc,
                const int* __restrict__ warpoff, const int* __restrict__ clsbox, int32_t* __restrict__ stats, int max_components) {
    PCS_SEG_THREAD();
    int32_t* st = stats + (size_t)blockIdx.y * n_classes * max_components * 5;
    if (blockIdx.x == 0 && (int)threadIdx.x < n_classes) {
        // row 0 of class c's table: box and count of every pixel that is NOT c (cv2's label 0 of that labelling)
        int g0 = 0, g1 = 0, g2 = -1, g3 = -1, cnt = 0;
        for (int o = 0; o <= n_classes; ++o) {
            if (o == (int)threadIdx.x) continue;
            const int* b = clsbox + ((size_t)blockIdx.y * (n_classes + 1) + o) * 5;
            if (b[4] <= 0) continue;
            g0 = max(g0, b[0]); g1 = max(g1, b[1]); g2 = max(g2, b[2]); g3 = max(g3, b[3]); cnt += b[4];
        }
        int32_t* o = st + (size_t)threadIdx.x * max_components * 5;
        if (cnt > 0) { o[0] = W - g0; o[1] = H - g1; o[2] = g2 - (W - g0) + 1; o[3] = g3 - (H - g1) + 1; o[4] = cnt; }
        else { o[0] = 0; o[1] = 0; o[2] = 0; o[3] = 0; o[4] = 0; }      // cv2 leaves an empty label's box at zero extent
    }
    const unsigned gm = valid ? rootmask[(size_t)blockIdx.y * H * segs + t] : 0u;
    if (!__any_sync(0xffffffffu, gm != 0u)) return;
    unsigned w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (gm) load_seg32(img + page_off + (size_t)y * W, x0, W, last_row, 0u, w);
    const int nw = gridDim.x * 8, wid = t >> 5, lane = threadIdx.x & 31;
    const int base = y * W + x0;
    for (int c = 0; c < n_classes; ++c) {
        unsigned b = gm ? gm & seg_bits<1>(w, c) : 0u;
        if (!__any_sync(0xffffffffu, b != 0u)) continue;
        const int cnt = __popc(b);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int l = __ldg(warpoff + ((size_t)blockIdx.y * n_classes + c) * nw + wid) + incl - cnt;
        int32_t* sc = st + (size_t)c * max_components * 5;
        while (b) {
            const int k = __ffs(b) - 1;
            b &= b - 1;
            if (++l >= max_components) break;
            const int* a = acc + (page_off + base + k) * 5;
            int32_t* o = sc + (size_t)l * 5;
            const int left = W - a[0], top = H - a[1];
            o[0] = left; o[1] = top; o[2] = a[2] - left + 1; o[3] = a[3] - top + 1; o[4] = a[4];
        }
    }
}

// the labelling shared by segment extraction and add_bounding_boxes: parents, root records, root masks, per-class counts
struct McBuffers { int* parent; int* acc; unsigned* rootmask; int* warpcnt; int* clsbox; int* extra; int nw; };

static int mc_label(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, size_t extra_words, McBuffers& b) {
    const size_t total = (size_t)H * W * n;
    const int segs = (W + 31) >> 5;
    const dim3 g = seg_grid(H, W, n);
    b.nw = (int)g.x * 8;
    const size_t mask_words = (size_t)n * H * segs, cnt_words = (size_t)n * n_classes * b.nw, box_words = (size_t)n * (n_classes + 1) * 5;
    PCS_TRY(scratch_reserve(ctx, (total * 6 + mask_words + cnt_words + box_words + extra_words) * 4 + 512));
    b.parent = reinterpret_cast<int*>(ctx->scratch);
    b.acc = b.parent + total;
    b.rootmask = reinterpret_cast<unsigned*>(b.acc + total * 5);
    b.warpcnt = reinterpret_cast<int*>(b.rootmask + mask_words);
    b.clsbox = b.warpcnt + cnt_words;
    b.extra = b.clsbox + box_words;
    cudaStream_t st = ctx->stream;
    static const bool big_table = [] { const char* e = getenv("PCSEG_MC_CAP"); return e && atoi(e) >= 1024; }();   // A/B switch
    static bool attr_set = false;
    if (!attr_set) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(mc_tile_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(McTileSmem<512>)));
        PCS_CUDA(ctx, cudaFuncSetAttribute(mc_tile_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(McTileSmem<1024>)));
        attr_set = true;
    }
    // diagnosis (PCSEG_CCL_POISON=1): parents / records / root masks start as 0x7f7f7f7f, so a pass that reads one the
    // labelling did not write (parents exist only on tile edges and at roots) walks out of the page instead of finding a
    // plausible stale value from an earlier call
    static const bool poison = [] { const char* e = getenv("PCSEG_CCL_POISON"); return e && e[0] == '1'; }();
    if (poison) PCS_CUDA(ctx, cudaMemsetAsync(b.parent, 0x7f, (total * 6 + mask_words + cnt_words) * 4, st));
    PCS_CUDA(ctx, cudaMemsetAsync(b.clsbox, 0, box_words * sizeof(int), st));
    const dim3 gt((W + kTileSegs * 32 - 1) / (kTileSegs * 32), (H + kTileRows - 1) / kTileRows, n);
    if (big_table) mc_tile_kernel<1024><<<gt, 256, sizeof(McTileSmem<1024>), st>>>(d_pred, H, W, b.parent, b.rootmask, b.acc);
    else mc_tile_kernel<512><<<gt, 256, sizeof(McTileSmem<512>), st>>>(d_pred, H, W, b.parent, b.rootmask, b.acc);
    PCS_LAUNCH_CHECK(ctx, "mc_tile_kernel");
    const int ncols = (W - 1) / (kTileSegs * 32), trows = (H - 1) / kTileRows;     // tile borders inside the page
    const bool halve = ccl_compress(true);
    if (ncols > 0) {
        const dim3 gc((unsigned)(((size_t)ncols * H + 255) / 256), n);
        if (halve) mc_border_cols_kernel<true><<<gc, 256, 0, st>>>(d_pred, H, W, b.parent);
        else mc_border_cols_kernel<false><<<gc, 256, 0, st>>>(d_pred, H, W, b.parent);
        PCS_LAUNCH_CHECK(ctx, "mc_border_cols_kernel");
    }
    if (trows > 0) {
        const dim3 gr((unsigned)(((size_t)trows * segs + 7) / 8), n);
        if (halve) mc_border_rows_kernel<true><<<gr, 256, 0, st>>>(d_pred, H, W, b.parent);
        else mc_border_rows_kernel<false><<<gr, 256, 0, st>>>(d_pred, H, W, b.parent);
        PCS_LAUNCH_CHECK(ctx, "mc_border_rows_kernel");
    }
    mc_fold_kernel<<<g, 256, 0, st>>>(d_pred, H, W, n_classes, b.parent, b.rootmask, b.acc, b.warpcnt, b.clsbox);
    PCS_LAUNCH_CHECK(ctx, "mc_fold_kernel");
    return PCS_OK;
}

// add_bounding_boxes on the one-pass labelling: the box of every component goes into the difference array of ITS class
// (four corner updates per component), the arrays of all classes are scanned together and a pixel takes the highest
// class whose boxes cover it (the reference paints the classes in ascending order over zeros, postprocess.py:29-42)
__global__ void __launch_bounds__(256)
mc_bbox_diff_kernel(const uint8_t* __restrict__ img, int H, int W, int n_classes, const unsigned* __restrict__ rootmask, const int* __restrict__ acc,
                    int* __restrict__ diff /*[page][class][(H + 1) * (W + 1)]*/) {
    PCS_SEG_THREAD();
    if (!valid) return;
    unsigned gm = rootmask[(size_t)blockIdx.y * H * segs + t];
    if (!gm) return;
    const uint8_t* row = img + page_off + (size_t)y * W;
    const int base = y * W + x0;
    const size_t plane = (size_t)(H + 1) * (W + 1);
    while (gm) {
        const int k = __ffs(gm) - 1;
        gm &= gm - 1;
        const int cls = __ldg(row + x0 + k);
        if (cls >= n_classes) continue;
        const int* a = acc + (page_off + base + k) * 5;
        const int bx0 = W - a[0], by0 = H - a[1], bx1 = a[2], by1 = a[3];
        int* d = diff + ((size_t)blockIdx.y * n_classes + cls) * plane;
        atomicAdd(&d[(size_t)by0 * (W + 1) + bx0], 1);
        atomicAdd(&d[(size_t)by0 * (W + 1) + bx1 + 1], -1);
        atomicAdd(&d[(size_t)(by1 + 1) * (W + 1) + bx0], -1);
        atomicAdd(&d[(size_t)(by1 + 1) * (W + 1) + bx1 + 1], 1);
    }
}

__global__ void __launch_bounds__(256)
mc_colscan_paint_kernel(const int* __restrict__ diff, int H, int W, int n_classes, const int* __restrict__ bandsum, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int rows = (H + kColBands - 1) / kColBands;
    const int y0 = blockIdx.y * rows, y1 = min(H, y0 + rows);
    const size_t plane = (size_t)(H + 1) * (W + 1);
    const int* d = diff + (size_t)blockIdx.z * n_classes * plane;
    uint8_t* o = out + (size_t)blockIdx.z * H * W;
    int cov[kMcMaxClasses];
#pragma unroll
    for (int c = 0; c < kMcMaxClasses; ++c) {
        cov[c] = 0;
        if (c < n_classes)
            for (int b = 0; b < (int)blockIdx.y; ++b) cov[c] += bandsum[(((size_t)blockIdx.z * n_classes + c) * kColBands + b) * W + x];
    }
    for (int y = y0; y < y1; ++y) {
        int best = 0;
#pragma unroll
        for (int c = 0; c < kMcMaxClasses; ++c)
            if (c < n_classes) {
                cov[c] += d[(size_t)c * plane + (size_t)y * (W + 1) + x];
                if (cov[c] > 0) best = c;
            }
        o[(size_t)y * W + x] = (uint8_t)best;
    }
}

static int launch_bounding_boxes_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out) {
    const size_t diff_elems = (size_t)n * n_classes * (H + 1) * (W + 1), band_elems = (size_t)n * n_classes * kColBands * W;
    McBuffers b;
    PCS_TRY(mc_label(ctx, d_pred, n, H, W, n_classes, diff_elems + band_elems, b));
    int* diff = b.extra;
    int* bandsum = diff + diff_elems;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(diff, 0, diff_elems * 4, st));
    mc_bbox_diff_kernel<<<seg_grid(H, W, n), 256, 0, st>>>(d_pred, H, W, n_classes, b.rootmask, b.acc, diff);
    PCS_LAUNCH_CHECK(ctx, "mc_bbox_diff_kernel");
    diff_rowscan_kernel<<<dim3((H + 1 + 7) / 8, n * n_classes), 256, 0, st>>>(diff, H + 1, W + 1);
    PCS_LAUNCH_CHECK(ctx, "diff_rowscan_kernel");
    diff_colband_sum_kernel<<<dim3((W + 255) / 256, kColBands, n * n_classes), 256, 0, st>>>(diff, H, W, bandsum);
    PCS_LAUNCH_CHECK(ctx, "diff_colband_sum_kernel");
    mc_colscan_paint_kernel<<<dim3((W + 255) / 256, kColBands, n), 256, 0, st>>>(diff, H, W, n_classes, bandsum, d_out);
    PCS_LAUNCH_CHECK(ctx, "mc_colscan_paint_kernel");
    return PCS_OK;
}

static int launch_class_components_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, int32_t* d_stats,
                                      int max_components, int32_t* d_ncomp) {
    McBuffers b;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(d_stats, 0, (size_t)n * n_classes * max_components * 5 * sizeof(int32_t), st));
    PCS_TRY(mc_label(ctx, d_pred, n, H, W, n_classes, (size_t)n * n_classes, b));
    ccl_scan_blocks_kernel<<<n * n_classes, 1024, 0, st>>>(b.warpcnt, b.nw, d_ncomp ? d_ncomp : b.extra);
    PCS_LAUNCH_CHECK(ctx, "ccl_scan_blocks_kernel");
    mc_write_kernel<<<seg_grid(H, W, n), 256, 0, st>>>(d_pred, H, W, n_classes, b.rootmask, b.acc, b.warpcnt, b.clsbox, d_stats, max_components);
    PCS_LAUNCH_CHECK(ctx, "mc_write_kernel");
    return PCS_OK;
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
