// One labelling of a class map for all classes: segment extraction (pcs_class_components, BASELINE configs[3]) and
// add_bounding_boxes (ocr4all_pixel_classifier/lib/postprocess.py:29-42, lib/cc.py:4-18) without a labelling per class.
#include "ccl_common.cuh"

namespace pcs {

// ---------------------------------------------------------------------------
// Segment extraction, every class in ONE labelling (n_classes <= kMcMaxClasses).  The per-class path (ccl.cu) labels
// (pred == c) once per class: three tile / border / accumulate / rank rounds per page, each reading the whole class
// map, and one set of five global atomics per run.  Every pixel belongs to exactly one of those labellings, so the
// class map is labelled once with "same byte" as the neighbour relation:
//   mc_tile   : 256 x 32 tiles in shared memory as in ccl_tile_kernel, the run structure taken from byte compares
//               (run start = differs from the left pixel, vertical union = equals the upper pixel where either row
//               starts a run); the statistics of every tile-local component (area, x extent, row mask) are gathered
//               with shared-memory atomics in a table indexed by the rank of the root among the tile's root
//               candidates (run starts without an equal upper pixel; candidates beyond the table go to global
//               atomics), and written once per tile-local root together with a bit mask of those roots;
//   mc_border : unions across tile borders on the global parents, one lane per pixel of a tile's first row and one
//               thread per row and vertical tile border;
//   mc_fold   : tile-local roots that lost their root status add their record to the component's root (one set of
//               atomics per tile and component, not per run); roots counted per class and warp (= 1024 pixels in
//               raster order), box and pixel count of every class for row 0 of the tables;
//   scan, mc_write : label = rank of the root among the roots of ITS class (cv2's numbering of that class's
//               labelling), stats row written from the root's record.
// The table of tile-local root candidates holds 512 records: room for five tiles per SM (1 024 records: four tiles, 5-8 %
// slower; PCSEG_MC_CAP=1024).
// ---------------------------------------------------------------------------

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
    static bool attr_set[64] = {};          // the attribute is per device
    if (ctx->device >= 64 || !attr_set[ctx->device]) {
        PCS_CUDA(ctx, cudaFuncSetAttribute(mc_tile_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(McTileSmem<512>)));
        PCS_CUDA(ctx, cudaFuncSetAttribute(mc_tile_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(McTileSmem<1024>)));
        if (ctx->device < 64) attr_set[ctx->device] = true;
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

int launch_bounding_boxes_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, uint8_t* d_out) {
    const size_t diff_elems = (size_t)n * n_classes * (H + 1) * (W + 1), band_elems = (size_t)n * n_classes * kColBands * W;
    McBuffers b;
    PCS_TRY(mc_label(ctx, d_pred, n, H, W, n_classes, diff_elems + band_elems, b));
    int* diff = b.extra;
    int* bandsum = diff + diff_elems;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(diff, 0, diff_elems * 4, st));
    mc_bbox_diff_kernel<<<seg_grid(H, W, n), 256, 0, st>>>(d_pred, H, W, n_classes, b.rootmask, b.acc, diff);
    PCS_LAUNCH_CHECK(ctx, "mc_bbox_diff_kernel");
    PCS_TRY(ccl_launch_diff_scans(ctx, diff, H, W, n * n_classes, bandsum));
    mc_colscan_paint_kernel<<<dim3((W + 255) / 256, kColBands, n), 256, 0, st>>>(diff, H, W, n_classes, bandsum, d_out);
    PCS_LAUNCH_CHECK(ctx, "mc_colscan_paint_kernel");
    return PCS_OK;
}

int launch_class_components_mc(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes, int32_t* d_stats,
                                      int max_components, int32_t* d_ncomp) {
    McBuffers b;
    cudaStream_t st = ctx->stream;
    PCS_CUDA(ctx, cudaMemsetAsync(d_stats, 0, (size_t)n * n_classes * max_components * 5 * sizeof(int32_t), st));
    PCS_TRY(mc_label(ctx, d_pred, n, H, W, n_classes, (size_t)n * n_classes, b));
    PCS_TRY(ccl_launch_scan_blocks(ctx, b.warpcnt, n * n_classes, b.nw, d_ncomp ? d_ncomp : b.extra));
    mc_write_kernel<<<seg_grid(H, W, n), 256, 0, st>>>(d_pred, H, W, n_classes, b.rootmask, b.acc, b.warpcnt, b.clsbox, d_stats, max_components);
    PCS_LAUNCH_CHECK(ctx, "mc_write_kernel");
    return PCS_OK;
}

}  // namespace pcs
