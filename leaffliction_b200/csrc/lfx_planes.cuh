// Bit-packed mask planes for one image per thread block: word-parallel morphology, run-based
// union-find connected components, per-component measures.  Shared by lfx_mask.cu (make_mask) and
// lfx_front.cu (Canny, inclusive/enhanced front ends, saliency, brown spots).
#pragma once
#include <math.h>
#include <string.h>

#include "lfx_common.cuh"

namespace {

#ifndef LFX_MT
#define LFX_MT 512
#endif
constexpr int MT = LFX_MT;       // threads per block (a translation unit may choose its own: this header is TU-local)
constexpr int NPLANES = 6;       // P0 raw, PB brown, PR result, T1..T3 temps
constexpr int RCAP_SMEM = 2048;  // runs kept in shared memory; larger tables use global scratch
constexpr int STAGE_BYTES = 6144;

struct FootRow {
    int8_t dy, o1, o2, pad;
};
struct Footprint {
    int n;
    int cross3;  // 1 when this is the 3x3 cross (set by make_ellipse): morph_cross3 fast path
    FootRow r[20];
};

struct MaskParams {
    lfx_mask_cfg cfg;
    Footprint fp_morph, fp_brown, fp_search;
    int H, W, WPR, NW;
    uint32_t lastmask;
    int planes_in_smem;
    int rcap_glob;
    int stage_rows;
    int mode;  // 0 make_mask, 1 postprocess only
    int search_is_e20;  // fp_search is the hard-coded 20x20 ellipse of dilate_ellipse20
    unsigned long long ws_per_block;
};

// Image geometry of the planes: runtime (any shape) or compile-time (the fused kernel's 256x256 fast
// path -- index math becomes shifts by immediates, loops unroll, no address rematerialisation).
struct DynGeom {
    int H, W, WPR, NW;
    uint32_t lastmask;
    int wshift;  // log2(WPR) when WPR is a power of two, else -1
};
template <int H_, int W_>
struct StaticGeom {
    static_assert(W_ % 32 == 0 && ((W_ / 32) & (W_ / 32 - 1)) == 0, "StaticGeom: W must be 32 * 2^k");
    static constexpr int H = H_, W = W_, WPR = W_ / 32, NW = H_ * (W_ / 32);
    static constexpr uint32_t lastmask = 0xFFFFFFFFu;
    static constexpr int wshift = (WPR == 1) ? 0 : (WPR == 2) ? 1 : (WPR == 4) ? 2 : (WPR == 8) ? 3 : (WPR == 16) ? 4 : 5;
};

template <class G>
struct CtxT : G {
    uint32_t* plane[NPLANES];
    int* wbase;
    int* wbase2;  // second per-word base table (background runs of ccl2); nullptr = ccl2 unavailable
    int R1;       // ccl2: number of foreground runs (ids [0,R1) foreground, [R1,R) background)
    // run tables (current selection) + both backing stores
    int* parent;
    uint32_t* geom;
    int* acc;
    uint16_t* ry;
    int *sm_parent, *gl_parent;
    uint32_t *sm_geom, *gl_geom;
    int *sm_acc, *gl_acc;
    uint16_t *sm_ry, *gl_ry;
    int rcap_glob;
    int rcap_smem;
    int R;
    // scratch
    int* s_tmp;                  // [40]
    unsigned long long* s_best;  // [1]
    int* s_bb;                   // [8]
    int* s_hist;                 // [256]
    int status;
    uint32_t* hp[6];     // scratch planes for dilate_ellipse20 (nullptr = use the generic morph)
    unsigned long long* tacc;  // debug phase timing (LFX_CORE_TIMING), nullptr otherwise
    long long tk0;
};

// debug: add the cycles since the last tick to slot `slot`
#define LFX_CTX_TICK(c, slot)                                                     \
    if ((c).tacc && threadIdx.x == 0) {                                           \
        const long long now_ = clock64();                                         \
        atomicAdd(&(c).tacc[slot], (unsigned long long)(now_ - (c).tk0));         \
        (c).tk0 = now_;                                                           \
    }
using Ctx = CtxT<DynGeom>;

// word index -> (row, word-in-row) without an integer division when WPR is a power of two
template <class C>
__device__ __forceinline__ void split_index(const C& c, int i, int& y, int& w) {
    if (c.wshift >= 0) {
        y = i >> c.wshift;
        w = i & (c.WPR - 1);
    } else {
        y = i / c.WPR;
        w = i - y * c.WPR;
    }
}

template <class C>
__device__ __forceinline__ uint32_t valid_mask(const C& c, int w) { return (w == c.WPR - 1) ? c.lastmask : 0xFFFFFFFFu; }

__device__ __forceinline__ void ctx_init_geometry(Ctx& c, int H, int W, int WPR, int NW, uint32_t lastmask) {
    c.H = H; c.W = W; c.WPR = WPR; c.NW = NW; c.lastmask = lastmask;
    c.wshift = ((WPR & (WPR - 1)) == 0) ? (31 - __clz(WPR)) : -1;
#pragma unroll
    for (int k = 0; k < 6; ++k) c.hp[k] = nullptr;
    c.tacc = nullptr;
    c.tk0 = 0;
    c.wbase2 = nullptr;
    c.R1 = 0;
}
template <class C>
__device__ __forceinline__ void ctx_clear_hp(C& c) {
#pragma unroll
    for (int k = 0; k < 6; ++k) c.hp[k] = nullptr;
    c.tacc = nullptr;
    c.tk0 = 0;
    c.wbase2 = nullptr;
    c.R1 = 0;
}

// ---------------------------------------------------------------- block primitives
// Exclusive block scan with ONE barrier: warp scans by shuffle, warp totals through s_tmp[0..15], then every
// warp re-scans the 16 totals itself.  The caller must pass a block barrier before the next call (s_tmp reuse);
// ccl() does (the barrier after run extraction).
__device__ int block_exscan(int v, int* s_tmp, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_tmp[wid] = inc;
    __syncthreads();
    int w = (lane < MT / 32) ? s_tmp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < MT / 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
    }
    total = __shfl_sync(0xffffffffu, winc, MT / 32 - 1);
    const int wexc = __shfl_sync(0xffffffffu, winc - w, wid);
    return wexc + inc - v;
}

template <class C>
__device__ __forceinline__ void plane_zero(uint32_t* p, const C& c) {
    for (int i = threadIdx.x; i < c.NW; i += MT) p[i] = 0;
}
template <class C>
__device__ __forceinline__ void plane_copy(uint32_t* d, const uint32_t* s, const C& c) {
    for (int i = threadIdx.x; i < c.NW; i += MT) d[i] = s[i];
}

// ---------------------------------------------------------------- morphology
template <bool DIL, class C>
__device__ void morph(const uint32_t* in, uint32_t* out, const Footprint& fp, const C& c) {
    for (int i = threadIdx.x; i < c.NW; i += MT) {
        int y, w;
        split_index(c, i, y, w);
        uint32_t acc = DIL ? 0u : 0xFFFFFFFFu;
        for (int k = 0; k < fp.n; ++k) {
            const int yy = y + fp.r[k].dy;
            if (yy < 0 || yy >= c.H) continue;  // outside rows are ignored by both min and max
            const uint32_t* row = in + yy * c.WPR;
            const uint32_t oob = DIL ? 0u : 0xFFFFFFFFu;
            uint32_t cur = row[w];
            if (!DIL && w == c.WPR - 1) cur |= ~c.lastmask;
            uint32_t prev = (w > 0) ? row[w - 1] : oob;
            uint32_t next = oob;
            if (w + 1 < c.WPR) {
                next = row[w + 1];
                if (!DIL && w + 1 == c.WPR - 1) next |= ~c.lastmask;
            }
            // OR / AND of source bits x+o1 .. x+o2 by doubling: T = the bit string re-based at offset o1 (64 bits are enough:
            // the span is at most 32 wide), then T op (T >> 1) op ... in log2(span) steps instead of one funnel shift per offset
            const int o1 = fp.r[k].o1, o2 = fp.r[k].o2, L = o2 - o1 + 1;
            uint32_t tlo, thi;
            if (o1 < 0) {
                tlo = __funnelshift_r(prev, cur, 32 + o1);
                thi = __funnelshift_r(cur, next, 32 + o1);
            } else {
                tlo = __funnelshift_r(cur, next, o1);
                thi = next >> o1;
            }
            unsigned long long r = ((unsigned long long)thi << 32) | tlo;
            int cov = 1;
            while (2 * cov <= L) {
                r = DIL ? (r | (r >> cov)) : (r & (r >> cov));
                cov *= 2;
            }
            if (cov < L) r = DIL ? (r | (r >> (L - cov))) : (r & (r >> (L - cov)));
            acc = DIL ? (acc | (uint32_t)r) : (acc & (uint32_t)r);
        }
        out[i] = acc & valid_mask(c, w);
    }
}

// 3x3 MORPH_ELLIPSE = cross (010/111/010): the footprint of every open/close on the default path
// (morph_kernel = brown_morph_kernel = 3, config.yaml).  5 words in, 2 funnel shifts.
template <bool DIL, class C>
__device__ void morph_cross3(const uint32_t* in, uint32_t* out, const C& c) {
    const uint32_t oob = DIL ? 0u : 0xFFFFFFFFu;
    for (int i = threadIdx.x; i < c.NW; i += MT) {
        int y, w;
        split_index(c, i, y, w);
        const bool last = (w == c.WPR - 1);
        uint32_t cur = in[i];
        uint32_t prev = (w > 0) ? in[i - 1] : oob;
        uint32_t next = oob;
        if (!last) {
            next = in[i + 1];
            if (!DIL && w + 1 == c.WPR - 1) next |= ~c.lastmask;
        }
        if (!DIL && last) cur |= ~c.lastmask;
        uint32_t up = (y > 0) ? in[i - c.WPR] : oob;
        uint32_t dn = (y < c.H - 1) ? in[i + c.WPR] : oob;
        if (!DIL && last) {
            up |= ~c.lastmask;
            dn |= ~c.lastmask;
        }
        const uint32_t l = __funnelshift_r(prev, cur, 31);  // bit x = source bit x-1
        const uint32_t r = __funnelshift_r(cur, next, 1);   // bit x = source bit x+1
        const uint32_t v = DIL ? (cur | l | r | up | dn) : (cur & l & r & up & dn);
        out[i] = v & valid_mask(c, w);
    }
}

static inline bool is_cross3(const Footprint& fp) {
    return fp.n == 3 && fp.r[0].dy == -1 && fp.r[0].o1 == 0 && fp.r[0].o2 == 0 && fp.r[1].dy == 0 && fp.r[1].o1 == -1 &&
           fp.r[1].o2 == 1 && fp.r[2].dy == 1 && fp.r[2].o1 == 0 && fp.r[2].o2 == 0;
}

template <bool DIL, class C>
__device__ __forceinline__ void morph_any(const uint32_t* in, uint32_t* out, const Footprint& fp, const C& c) {
    if (fp.cross3)
        morph_cross3<DIL>(in, out, c);
    else
        morph<DIL>(in, out, fp, c);
}

// Dilation by cv2.getStructuringElement(MORPH_ELLIPSE, (20, 20)) (mask.py:341, anchor (10,10)):
// row spans are nested -- dy -10: [0,0]; +-9: [-4,4]; +-8: [-6,6]; +-7: [-7,7]; +-6: [-8,8]; +-5,+-4: [-9,9];
// -3..3: [-10,9] -- so the six wider horizontal dilations are built incrementally from one
// (prev,cur,next) word triple into six scratch planes (20 funnel shifts per word instead of 325),
// then each output word ORs 20 plane rows.  Needs c.hp[0..5]; ends with a block barrier.
template <class C>
__device__ void dilate_ellipse20(const uint32_t* in, uint32_t* out, const C& c) {
    for (int i = threadIdx.x; i < c.NW; i += MT) {
        int y, w;
        split_index(c, i, y, w);
        const uint32_t cur = in[i];
        const uint32_t prev = (w > 0) ? in[i - 1] : 0u;
        const uint32_t next = (w + 1 < c.WPR) ? in[i + 1] : 0u;
#define LFX_SL(o) __funnelshift_r(cur, next, (o))        /* bit x = source bit x+o */
#define LFX_SR(o) __funnelshift_r(prev, cur, 32 - (o))   /* bit x = source bit x-o */
        uint32_t h = cur | LFX_SL(1) | LFX_SL(2) | LFX_SL(3) | LFX_SL(4) | LFX_SR(1) | LFX_SR(2) | LFX_SR(3) | LFX_SR(4);
        c.hp[0][i] = h;  // [-4,4]
        h |= LFX_SL(5) | LFX_SL(6) | LFX_SR(5) | LFX_SR(6);
        c.hp[1][i] = h;  // [-6,6]
        h |= LFX_SL(7) | LFX_SR(7);
        c.hp[2][i] = h;  // [-7,7]
        h |= LFX_SL(8) | LFX_SR(8);
        c.hp[3][i] = h;  // [-8,8]
        h |= LFX_SL(9) | LFX_SR(9);
        c.hp[4][i] = h;  // [-9,9]
        h |= LFX_SR(10);
        c.hp[5][i] = h;  // [-10,9]
#undef LFX_SL
#undef LFX_SR
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c.NW; i += MT) {
        int y, w;
        split_index(c, i, y, w);
        const int wp = c.WPR;
        uint32_t acc = 0;
        auto row = [&](const uint32_t* p, int dy) {
            const int yy = y + dy;
            if (yy >= 0 && yy < c.H) acc |= p[i + dy * wp];
        };
        row(in, -10);
        row(c.hp[0], -9); row(c.hp[0], 9);
        row(c.hp[1], -8); row(c.hp[1], 8);
        row(c.hp[2], -7); row(c.hp[2], 7);
        row(c.hp[3], -6); row(c.hp[3], 6);
        row(c.hp[4], -5); row(c.hp[4], 5); row(c.hp[4], -4); row(c.hp[4], 4);
#pragma unroll
        for (int dy = -3; dy <= 3; ++dy) row(c.hp[5], dy);
        out[i] = acc & valid_mask(c, w);
    }
    __syncthreads();
}

// true when fp is exactly the 20x20 ellipse dilate_ellipse20 hard-codes
static inline bool is_ellipse20(const Footprint& fp) {
    static const int8_t o1[20] = {0, -4, -6, -7, -8, -9, -9, -10, -10, -10, -10, -10, -10, -10, -9, -9, -8, -7, -6, -4};
    static const int8_t o2[20] = {0, 4, 6, 7, 8, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 8, 7, 6, 4};
    if (fp.n != 20) return false;
    for (int k = 0; k < 20; ++k)
        if (fp.r[k].dy != k - 10 || fp.r[k].o1 != o1[k] || fp.r[k].o2 != o2[k]) return false;
    return true;
}

// ---------------------------------------------------------------- runs + union-find
__device__ __forceinline__ uint32_t starts_of(const uint32_t* m, int idx, int w) {
    const uint32_t cur = m[idx];
    const uint32_t pb = (w > 0) ? (m[idx - 1] >> 31) : 0u;
    return cur & ~((cur << 1) | pb);
}

__device__ __forceinline__ int uf_find(const int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}
__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// Flatten the union-find forest so that parent[r] is the root.  Links always point to smaller ids, so one warp
// can sweep the ids in ascending batches of 32: parents below the batch are already roots (one hop), chains
// inside the batch shrink by pointer jumping (<= 6 rounds).  For large run tables the whole block does
// pointer jumping instead (three hops per round; racing reads see ancestors only).  Ends with a block barrier.
template <class C>
__device__ void uf_flatten(int total, C& c) {
    int* parent = c.parent;
    if (total <= 64) {  // tiny tables only: measured slower than block-wide jumping beyond a couple of batches
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            for (int base = 0; base < total; base += 32) {
                const int r = base + lane;
                int p = (r < total) ? parent[r] : 0;
                for (;;) {
                    int q = p;
                    if (r < total) q = parent[p];
                    const bool moved = (q != p);
                    if (moved) {
                        parent[r] = q;
                        p = q;
                    }
                    if (!__any_sync(0xffffffffu, moved)) break;
                    __syncwarp();
                }
                __syncwarp();
            }
        }
        __syncthreads();
        return;
    }
    for (;;) {
        int moved = 0;
        for (int r = threadIdx.x; r < total; r += MT) {
            const int p1 = parent[r];
            int q = parent[p1];
            if (q != p1) {
                q = parent[q];       // up to four hops per round: the depth shrinks fivefold
                q = parent[q];
                q = parent[q];
                parent[r] = q;
                moved = 1;
            }
        }
        if (!__syncthreads_or(moved)) break;
    }
}

// ---------------------------------------------------------------- ccl2: foreground + background in one pass
// word i of the plane (INV = false) or of its complement restricted to the image (INV = true)
template <bool INV, class C>
__device__ __forceinline__ uint32_t plane_word(const uint32_t* m, int idx, int w, const C& c) {
    return INV ? (~m[idx] & valid_mask(c, w)) : m[idx];
}
template <bool INV, class C>
__device__ __forceinline__ uint32_t starts_of2(const uint32_t* m, int idx, int w, const C& c) {
    const uint32_t cur = plane_word<INV>(m, idx, w, c);
    const uint32_t pb = (w > 0) ? (plane_word<INV>(m, idx - 1, w - 1, c) >> 31) : 0u;
    return cur & ~((cur << 1) | pb);
}

// runs of word i -> tables, ids from `id` upwards; returns the next free id
template <bool INV, class C>
__device__ __forceinline__ int extract_runs(const uint32_t* m, int i, int id, C& c) {
    int y, w;
    split_index(c, i, y, w);
    uint32_t st = starts_of2<INV>(m, i, w, c);
    const uint32_t word = plane_word<INV>(m, i, w, c);
    while (st) {
        const int b = __ffs(st) - 1;
        st &= st - 1;
        const int x0 = w * 32 + b;
        const uint32_t inv = ~(word >> b);
        const int z = __ffs(inv) - 1;  // first zero at/after b (relative); -1 if none
        int x1;
        if (inv != 0 && z < 32 - b) {
            x1 = x0 + z - 1;
        } else {  // the run reaches bit 31: it may continue in the next words
            x1 = w * 32 + 31;
            int ww = w + 1;
            while (ww < c.WPR) {
                const uint32_t nx = plane_word<INV>(m, y * c.WPR + ww, ww, c);
                if (nx == 0xFFFFFFFFu) {
                    x1 += 32;
                    ++ww;
                    continue;
                }
                x1 += __ffs(~nx) - 1;
                break;
            }
        }
        c.geom[id] = (uint32_t)x0 | ((uint32_t)x1 << 16);
        c.ry[id] = (uint16_t)y;
        c.parent[id] = id;
        c.acc[id] = 0;
        ++id;
    }
    return id;
}

// unite run r (row y > 0, columns lo..hi already widened for 8-connectivity) with the runs of the row above
template <bool INV, class C>
__device__ __forceinline__ void unite_up(const uint32_t* m, int r, int y, int lo, int hi, const int* wb, int idbase, C& c) {
    int p = lo;
    while (p <= hi) {
        int wi = p >> 5;
        uint32_t v = plane_word<INV>(m, (y - 1) * c.WPR + wi, wi, c) & (0xFFFFFFFFu << (p & 31));
        while (v == 0) {
            ++wi;
            if (wi * 32 > hi) break;
            v = plane_word<INV>(m, (y - 1) * c.WPR + wi, wi, c);
        }
        if (v == 0) break;
        const int b = __ffs(v) - 1;
        if (wi * 32 + b > hi) break;
        const int idxw = (y - 1) * c.WPR + wi;
        const uint32_t st = starts_of2<INV>(m, idxw, wi, c);
        const int id2 = idbase + wb[idxw] + __popc(st & ((2u << b) - 1u)) - 1;
        uf_unite(c.parent, r, id2);
        p = (int)(c.geom[id2] >> 16) + 1;
    }
}

// id of the foreground run that contains pixel (y, x) of plane m (the pixel must be set); after ccl2
template <class C>
__device__ __forceinline__ int fg_run_at(const uint32_t* m, int y, int x, const C& c) {
    const int wi = x >> 5, idxw = y * c.WPR + wi;
    const uint32_t st = starts_of2<false>(m, idxw, wi, c);
    return c.wbase[idxw] + __popc(st & ((2u << (x & 31)) - 1u)) - 1;
}

// Calls fn(id2, k) for the k-th run of row y-1 (plane m, or its complement when INV) that overlaps columns lo..hi.
// wb = per-word first-run id table of that kind (relative to idbase).
template <bool INV, class C, class F>
__device__ __forceinline__ void for_upper_runs(const uint32_t* m, int y, int lo, int hi, const int* wb, int idbase, const C& c, F fn) {
    int p = lo, k = 0;
    while (p <= hi) {
        int wi = p >> 5;
        uint32_t v = plane_word<INV>(m, (y - 1) * c.WPR + wi, wi, c) & (0xFFFFFFFFu << (p & 31));
        while (v == 0) {
            ++wi;
            if (wi * 32 > hi) break;
            v = plane_word<INV>(m, (y - 1) * c.WPR + wi, wi, c);
        }
        if (v == 0) break;
        const int b = __ffs(v) - 1;
        if (wi * 32 + b > hi) break;
        const int idxw = (y - 1) * c.WPR + wi;
        const uint32_t st = starts_of2<INV>(m, idxw, wi, c);
        const int id2 = idbase + wb[idxw] + __popc(st & ((2u << b) - 1u)) - 1;
        fn(id2, k++);
        p = (int)(c.geom[id2] >> 16) + 1;
    }
}

// Union phase in two sweeps.  (1) LINK: every run points at the FIRST overlapping run of the row above -- a plain
// store, no search, no atomic (ids in the row above are smaller, so links still point to smaller ids); flatten.
// (2) MERGE: only runs with further overlapping neighbours unite the (now shallow) trees with atomicMin; flatten.
// A blob with one run per row -- the usual leaf -- needs no atomic at all.
template <class C, class RANGE>
__device__ void uf_link_merge(const uint32_t* m, int total, C& c, RANGE range) {
    for (int r = threadIdx.x; r < total; r += MT) {
        const int y = c.ry[r];
        if (y == 0) continue;
        range(r, y, [&](int id2, int k) { if (k == 0) c.parent[r] = id2; });
    }
    __syncthreads();
    uf_flatten(total, c);
    int any = 0;
    for (int r = threadIdx.x; r < total; r += MT) {
        const int y = c.ry[r];
        if (y == 0) continue;
        range(r, y, [&](int id2, int k) {
            if (k > 0) {
                uf_unite(c.parent, r, id2);
                any = 1;
            }
        });
    }
    if (__syncthreads_or(any)) uf_flatten(total, c);
}

// Labels the runs of plane m (CONN = 4 or 8).  After return: c.R runs, c.parent[r] = root run id
// (the smallest id of the component = its first run in raster order), c.geom / c.ry, c.acc = 0.
template <int CONN, class C>
__device__ void ccl(const uint32_t* m, C& c) {
    const int per = (c.NW + MT - 1) / MT;
    const int i0 = min(c.NW, (int)threadIdx.x * per), i1 = min(c.NW, i0 + per);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += __popc(starts_of(m, i, i % c.WPR));
    int total;
    int base = block_exscan(cnt, c.s_tmp, total);
    const int base0 = base;
    for (int i = i0; i < i1; ++i) {
        c.wbase[i] = base;
        base += __popc(starts_of(m, i, i % c.WPR));
    }
    c.R = total;
    if (total <= c.rcap_smem) {
        c.parent = c.sm_parent; c.geom = c.sm_geom; c.acc = c.sm_acc; c.ry = c.sm_ry;
    } else {
        c.parent = c.gl_parent; c.geom = c.gl_geom; c.acc = c.gl_acc; c.ry = c.gl_ry;
        c.status |= 2;
    }
    int* parent = c.parent;
    // Run extraction on the thread's own contiguous words (same mapping as the count above: each thread
    // sees a mix of word columns, so speckled borders and clean interiors balance out; ids follow from the
    // thread's scan base, no barrier needed before this loop).
    int id = base0;
    for (int i = i0; i < i1; ++i) {
        int y, w;
        split_index(c, i, y, w);
        uint32_t st = starts_of(m, i, w);
        const uint32_t word = m[i];
        while (st) {
            const int b = __ffs(st) - 1;
            st &= st - 1;
            const int x0 = w * 32 + b;
            const uint32_t inv = ~(word >> b);
            const int z = __ffs(inv) - 1;  // first zero at/after b (relative); -1 if none
            int x1;
            if (inv != 0 && z < 32 - b) {
                x1 = x0 + z - 1;
            } else {
                x1 = w * 32 + 31;
                int ww = w + 1;
                while (ww < c.WPR) {
                    const uint32_t nx = m[y * c.WPR + ww];
                    if (nx == 0xFFFFFFFFu) {
                        x1 += 32;
                        ++ww;
                        continue;
                    }
                    x1 += __ffs(~nx) - 1;
                    break;
                }
            }
            c.geom[id] = (uint32_t)x0 | ((uint32_t)x1 << 16);
            c.ry[id] = (uint16_t)y;
            parent[id] = id;
            c.acc[id] = 0;
            ++id;
        }
    }
    __syncthreads();
    uf_link_merge(m, total, c, [&](int r, int y, auto fn) {
        const uint32_t g = c.geom[r];
        const int x0 = g & 0xFFFF, x1 = g >> 16;
        for_upper_runs<false>(m, y, max(0, x0 - (CONN == 8 ? 1 : 0)), min(c.W - 1, x1 + (CONN == 8 ? 1 : 0)), c.wbase, 0, c, fn);
    });
}

// Labels the 8-connected runs of plane m (ids [0, R1)) AND the 4-connected runs of its complement (ids
// [R1, R)) with one scan / extract / union / flatten sequence -- the two labelings largest_external needs,
// for the latency of one.  Needs c.wbase2.
template <class C>
__device__ void ccl2(const uint32_t* m, C& c) {
    const int per = (c.NW + MT - 1) / MT;
    const int i0 = min(c.NW, (int)threadIdx.x * per), i1 = min(c.NW, i0 + per);
    int cf = 0, cb = 0;
    for (int i = i0; i < i1; ++i) {
        int y, w;
        split_index(c, i, y, w);
        cf += __popc(starts_of2<false>(m, i, w, c));
        cb += __popc(starts_of2<true>(m, i, w, c));
    }
    // both counts ride one scan when each total fits 16 bits (<= 16 runs per word and kind)
    LFX_CTX_TICK(c, 15)
    int tot_f, tot_b, bf, bb;
    if (c.NW <= 2048) {
        int total;
        const unsigned base = (unsigned)block_exscan((int)((unsigned)cf | ((unsigned)cb << 16)), c.s_tmp, total);
        bf = (int)(base & 0xFFFFu); bb = (int)(base >> 16);
        tot_f = (int)((unsigned)total & 0xFFFFu); tot_b = (int)((unsigned)total >> 16);
        // a total of exactly 65536 foreground runs cannot happen (NW * 16 <= 32768)
    } else {
        bf = block_exscan(cf, c.s_tmp, tot_f);
        __syncthreads();
        bb = block_exscan(cb, c.s_tmp, tot_b);
    }
    const int R1 = tot_f, total = tot_f + tot_b;
    c.R1 = R1;
    c.R = total;
    if (total <= c.rcap_smem) {
        c.parent = c.sm_parent; c.geom = c.sm_geom; c.acc = c.sm_acc; c.ry = c.sm_ry;
    } else {
        c.parent = c.gl_parent; c.geom = c.gl_geom; c.acc = c.gl_acc; c.ry = c.gl_ry;
        c.status |= 2;
    }
    int idf = bf, idb = R1 + bb;
    for (int i = i0; i < i1; ++i) {
        c.wbase[i] = idf;
        c.wbase2[i] = idb - R1;
        idf = extract_runs<false>(m, i, idf, c);
        idb = extract_runs<true>(m, i, idb, c);
    }
    __syncthreads();
    LFX_CTX_TICK(c, 10)
    uf_link_merge(m, total, c, [&](int r, int y, auto fn) {
        const uint32_t g = c.geom[r];
        const int x0 = g & 0xFFFF, x1 = g >> 16;
        if (r < R1)
            for_upper_runs<false>(m, y, max(0, x0 - 1), min(c.W - 1, x1 + 1), c.wbase, 0, c, fn);
        else
            for_upper_runs<true>(m, y, x0, x1, c.wbase2, R1, c, fn);
    });
    LFX_CTX_TICK(c, 11)
}

template <class C>
__device__ __forceinline__ void set_run(uint32_t* out, int y, int x0, int x1, const C& c) {
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    for (int w = w0; w <= w1; ++w) {
        uint32_t mk = 0xFFFFFFFFu;
        if (w == w0) mk &= 0xFFFFFFFFu << (x0 & 31);
        if (w == w1) mk &= 0xFFFFFFFFu >> (31 - (x1 & 31));
        atomicOr(&out[y * c.WPR + w], mk);
    }
}

__device__ __forceinline__ int popc_range(const uint32_t* row, int x0, int x1) {
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    int n = 0;
    for (int w = w0; w <= w1; ++w) {
        uint32_t mk = 0xFFFFFFFFu;
        if (w == w0) mk &= 0xFFFFFFFFu << (x0 & 31);
        if (w == w1) mk &= 0xFFFFFFFFu >> (31 - (x1 & 31));
        n += __popc(row[w] & mk);
    }
    return n;
}
template <class C>
__device__ __forceinline__ int get_bit(const uint32_t* m, int y, int x, const C& c) {
    if (y < 0 || y >= c.H || x < 0 || x >= c.W) return 0;
    return (m[y * c.WPR + (x >> 5)] >> (x & 31)) & 1;
}

// acc[root] += pixels
template <class C>
__device__ void measure_area(C& c) {
    for (int r = threadIdx.x; r < c.R; r += MT) {
        const uint32_t g = c.geom[r];
        atomicAdd(&c.acc[c.parent[r]], (int)(g >> 16) - (int)(g & 0xFFFF) + 1);
    }
    __syncthreads();
}

// out = runs whose component has >= min_area pixels (out must be zeroed + synced by the caller)
template <class C>
__device__ void keep_area_ge(uint32_t* out, int min_area, C& c) {
    for (int r = threadIdx.x; r < c.R; r += MT) {
        if (c.acc[c.parent[r]] >= min_area) {
            const uint32_t g = c.geom[r];
            set_run(out, c.ry[r], g & 0xFFFF, g >> 16, c);
        }
    }
    __syncthreads();
}


static inline Footprint make_ellipse(int k) {
    // cv::getStructuringElement(MORPH_ELLIPSE, (k,k)), anchor (k/2, k/2)
    Footprint f;
    f.n = 0;
    const int r = k / 2, cc = k / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; ++i) {
        const int dy = i - r;
        if (abs(dy) > r) continue;
        const int dx = (int)nearbyint(cc * sqrt((r * r - dy * dy) * inv_r2));
        const int j1 = max(cc - dx, 0), j2 = min(cc + dx + 1, k);
        if (j2 <= j1) continue;
        f.r[f.n].dy = (int8_t)dy;
        f.r[f.n].o1 = (int8_t)(j1 - cc);
        f.r[f.n].o2 = (int8_t)(j2 - 1 - cc);
        f.r[f.n].pad = 0;
        ++f.n;
    }
    f.cross3 = is_cross3(f) ? 1 : 0;
    return f;
}


}  // namespace
