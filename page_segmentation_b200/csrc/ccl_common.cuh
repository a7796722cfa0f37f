// Shared pieces of the connected-component kernels (ccl.cu: labelling, vote, boxes, compute_char_height; ccl_onepass.cu: one
// labelling of a class map for all classes): union-find walks, 32-pixel segments as bit masks, the tile geometry.
#pragma once
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

// Re-linking of union operands: measured on 32 A4 pages it pays where components are huge (the page background of
// add_bounding_boxes: 1.91 -> 1.80 ms) and costs where they are letters (compute_char_height 2.62 -> 2.84 ms,
// cc_majority 0.44 -> 0.46 ms), so it is on for class-match labelling only.  PCSEG_CCL_COMPRESS=0/1 forces it.
static inline bool ccl_compress(bool match) {
    static const char* e = getenv("PCSEG_CCL_COMPRESS");
    return e ? atoi(e) != 0 : match;
}

constexpr int kColBands = 16;                       // row bands of the column scans of the box difference arrays
constexpr int kMcMaxClasses = 8;                    // classes the one-pass labelling (ccl_onepass.cu) handles
inline bool mc_per_class() {                        // A/B switch: one labelling per class instead of one for all
    static const char* e = getenv("PCSEG_SEGMENTS_PER_CLASS");
    return e && atoi(e) != 0;
}

// ccl_onepass.cu
int launch_bounding_boxes_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out);
int launch_class_components_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, int32_t* d_stats,
                               int max_components, int32_t* d_ncomp);
// ccl.cu: launches of its kernels for the other translation unit (a kernel is launched from the unit that defines it)
int ccl_launch_scan_blocks(pcs_ctx* ctx, int* blocksum, int rows, int nblocks, int* ncomp);
int ccl_launch_diff_scans(pcs_ctx* ctx, int* diff, int H, int W, int planes, int* bandsum);

}  // namespace pcs
