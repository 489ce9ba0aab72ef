// Device-side building blocks of make_mask (srcs/transform/filters/mask.py:53-69,335-411,548-582):
// largest filled external contour, _postprocess_mask, Otsu, the RGB pixel passes and plane <-> byte
// conversion.  Shared by lfx_mask.cu (k_make_mask) and lfx_core.cu (the fused core-profile kernel).
#pragma once
#include <float.h>
#include <math.h>

#include "lfx_planes.cuh"

namespace {

// largest_contour + contour_to_mask (Transformation.py:285-299) with ONE labeling pass (ccl2): the 8-connected
// components of `in` and the 4-connected components of its complement are labelled together.  A complement
// component that does not touch the border is a hole; it is enclosed by exactly one foreground component
// (8/4 duality), so hole filling never merges components: each hole run is re-parented to the component of
// the foreground pixel on its left.  2*contourArea = 2N - (P - Q1) - 2 is accumulated per run on the filled
// plane (tA) with run-end terms that do not assume maximal runs.  Same outputs as largest_external.
template <class C>
__device__ bool largest_external2(const uint32_t* in, uint32_t* tA, uint32_t* out, int* info8, C& c) {
    ccl2(in, c);
    const int R1 = c.R1, R = c.R;
    // background components that touch the border are "outside"
    for (int r = R1 + threadIdx.x; r < R; r += MT) {
        const uint32_t g = c.geom[r];
        const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
        if (y == 0 || y == c.H - 1 || x0 == 0 || x1 == c.W - 1) atomicOr(&c.acc[c.parent[r]], 1);
    }
    for (int i = threadIdx.x; i < c.NW; i += MT) {
        tA[i] = in[i];
        out[i] = 0;
    }
    if (threadIdx.x == 0) {
        *c.s_best = 0ull;
        c.s_bb[0] = 0x7fffffff; c.s_bb[1] = 0x7fffffff; c.s_bb[2] = -1; c.s_bb[3] = -1; c.s_bb[4] = 0;
    }
    __syncthreads();
    // holes: fill them in tA and hand each hole run to its enclosing foreground component
    for (int r = R1 + threadIdx.x; r < R; r += MT) {
        if (c.acc[c.parent[r]]) {
            c.parent[r] = -1;  // outside
        } else {
            const uint32_t g = c.geom[r];
            const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
            set_run(tA, y, x0, x1, c);
            c.parent[r] = c.parent[fg_run_at(in, y, x0 - 1, c)];  // x0 > 0: a run starting at the border is outside
        }
    }
    __syncthreads();
    // per-run contribution to 2N - (P - Q1) on the filled plane
    for (int r = threadIdx.x; r < R; r += MT) {
        const int root = c.parent[r];
        if (root < 0) continue;
        const uint32_t g = c.geom[r];
        const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
        const int eL = !get_bit(tA, y, x0 - 1, c), eR = !get_bit(tA, y, x1 + 1, c);  // run ends that are component ends
        int v = -(eL + eR);
        if (y > 0) v += popc_range(tA + (y - 1) * c.WPR, x0, x1);
        if (y < c.H - 1) v += popc_range(tA + (y + 1) * c.WPR, x0, x1);
        v += (eR && !get_bit(tA, y + 1, x1, c) && !get_bit(tA, y + 1, x1 + 1, c));
        v += (eL && !get_bit(tA, y + 1, x0, c) && !get_bit(tA, y + 1, x0 - 1, c));
        v += (eR && !get_bit(tA, y - 1, x1, c) && !get_bit(tA, y - 1, x1 + 1, c));
        v += (eL && !get_bit(tA, y - 1, x0, c) && !get_bit(tA, y - 1, x0 - 1, c));
        atomicAdd(&c.acc[root], v);
    }
    __syncthreads();
    // argmax of (area2, first-pixel order): ties go to the LARGEST root id (reverse discovery order)
    for (int r = threadIdx.x; r < R1; r += MT) {
        if (c.parent[r] == r) {
            const unsigned long long key = ((unsigned long long)(uint32_t)(c.acc[r] - 2 + 1) << 32) | (uint32_t)r;
            atomicMax(c.s_best, key);
        }
    }
    __syncthreads();
    const unsigned long long best = *c.s_best;
    if (best == 0ull) {
        if (threadIdx.x < 8) info8[threadIdx.x] = 0;
        __syncthreads();
        return false;
    }
    const int win = (int)(best & 0xFFFFFFFFu);
    for (int r = threadIdx.x; r < R; r += MT) {
        if (c.parent[r] == win) {
            const uint32_t g = c.geom[r];
            const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
            set_run(out, y, x0, x1, c);
            atomicMin(&c.s_bb[0], x0);
            atomicMin(&c.s_bb[1], y);
            atomicMax(&c.s_bb[2], x1);
            atomicMax(&c.s_bb[3], y);
            atomicAdd(&c.s_bb[4], x1 - x0 + 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        info8[0] = 1;
        info8[1] = c.s_bb[0];
        info8[2] = c.s_bb[1];
        info8[3] = c.s_bb[2] - c.s_bb[0] + 1;
        info8[4] = c.s_bb[3] - c.s_bb[1] + 1;
        info8[5] = (int)(best >> 32) - 1;
        info8[6] = c.s_bb[4];
        info8[7] = (int)(c.geom[win] & 0xFFFF);  // x of the component's first raster pixel (its y is info8[2])
    }
    __syncthreads();
    return true;
}

// largest_contour + contour_to_mask (Transformation.py:285-299) on plane `in`.
// Temps: tA (inverted / filled), out receives the selected filled component (zero when none).
// info8: {found,x,y,w,h,area2,npix,-}.  Returns found (block-uniform).
template <class C>
__device__ bool largest_external(const uint32_t* in, uint32_t* tA, uint32_t* out, int* info8, C& c) {
    if (c.wbase2) return largest_external2(in, tA, out, info8, c);
    // outside = 4-connected background reachable from the border
    for (int i = threadIdx.x; i < c.NW; i += MT) tA[i] = ~in[i] & valid_mask(c, i % c.WPR);
    __syncthreads();
    ccl<4>(tA, c);
    for (int r = threadIdx.x; r < c.R; r += MT) {
        const uint32_t g = c.geom[r];
        const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
        if (y == 0 || y == c.H - 1 || x0 == 0 || x1 == c.W - 1) atomicOr(&c.acc[c.parent[r]], 1);
    }
    __syncthreads();
    // filled = in | enclosed background.  Built in `out`, then moved to tA (tA is still the ccl input).
    plane_copy(out, in, c);
    __syncthreads();
    for (int r = threadIdx.x; r < c.R; r += MT) {
        if (!c.acc[c.parent[r]]) {
            const uint32_t g = c.geom[r];
            set_run(out, c.ry[r], g & 0xFFFF, g >> 16, c);
        }
    }
    __syncthreads();
    plane_copy(tA, out, c);
    __syncthreads();
    ccl<8>(tA, c);
    // per-run contribution to 2N - (P - Q1):  popc(up) + popc(down) - 2 + Q1
    for (int r = threadIdx.x; r < c.R; r += MT) {
        const uint32_t g = c.geom[r];
        const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
        int v = -2;
        if (y > 0) v += popc_range(tA + (y - 1) * c.WPR, x0, x1);
        if (y < c.H - 1) v += popc_range(tA + (y + 1) * c.WPR, x0, x1);
        v += (!get_bit(tA, y + 1, x1, c) && !get_bit(tA, y + 1, x1 + 1, c));
        v += (!get_bit(tA, y + 1, x0, c) && !get_bit(tA, y + 1, x0 - 1, c));
        v += (!get_bit(tA, y - 1, x1, c) && !get_bit(tA, y - 1, x1 + 1, c));
        v += (!get_bit(tA, y - 1, x0, c) && !get_bit(tA, y - 1, x0 - 1, c));
        atomicAdd(&c.acc[c.parent[r]], v);
    }
    if (threadIdx.x == 0) {
        *c.s_best = 0ull;
        c.s_bb[0] = 0x7fffffff; c.s_bb[1] = 0x7fffffff; c.s_bb[2] = -1; c.s_bb[3] = -1; c.s_bb[4] = 0;
    }
    __syncthreads();
    // argmax of (area2, first-pixel order): ties go to the LARGEST root id (reverse discovery order)
    for (int r = threadIdx.x; r < c.R; r += MT) {
        if (c.parent[r] == r) {
            const unsigned long long key = ((unsigned long long)(uint32_t)(c.acc[r] - 2 + 1) << 32) | (uint32_t)r;
            atomicMax(c.s_best, key);
        }
    }
    __syncthreads();
    const unsigned long long best = *c.s_best;
    plane_zero(out, c);
    __syncthreads();
    if (best == 0ull) {
        if (threadIdx.x < 8) info8[threadIdx.x] = 0;
        __syncthreads();
        return false;
    }
    const int win = (int)(best & 0xFFFFFFFFu);
    for (int r = threadIdx.x; r < c.R; r += MT) {
        if (c.parent[r] == win) {
            const uint32_t g = c.geom[r];
            const int y = c.ry[r], x0 = g & 0xFFFF, x1 = g >> 16;
            set_run(out, y, x0, x1, c);
            atomicMin(&c.s_bb[0], x0);
            atomicMin(&c.s_bb[1], y);
            atomicMax(&c.s_bb[2], x1);
            atomicMax(&c.s_bb[3], y);
            atomicAdd(&c.s_bb[4], x1 - x0 + 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        info8[0] = 1;
        info8[1] = c.s_bb[0];
        info8[2] = c.s_bb[1];
        info8[3] = c.s_bb[2] - c.s_bb[0] + 1;
        info8[4] = c.s_bb[3] - c.s_bb[1] + 1;
        info8[5] = (int)(best >> 32) - 1;
        info8[6] = c.s_bb[4];
        info8[7] = (int)(c.geom[win] & 0xFFFF);  // x of the component's first raster pixel (its y is info8[2])
    }
    __syncthreads();
    return true;
}

// _postprocess_mask (mask.py:53-69): raw -> result plane `out`; temps t1..t3.
// When no contour exists, `out` holds the opened mask (reference returns (opened, None)).
template <class C>
__device__ bool postprocess(const uint32_t* raw, uint32_t* out, uint32_t* t1, uint32_t* t2, uint32_t* t3, int* info8,
                            const MaskParams& P, C& c) {
    // pcv.fill: drop 4-connected components with < fill_size pixels
    ccl<4>(raw, c);
    measure_area(c);
    plane_zero(t1, c);
    __syncthreads();
    keep_area_ge(t1, P.cfg.fill_size, c);
    LFX_CTX_TICK(c, 3)
    // close = erode(dilate), open = dilate(erode)
    morph_any<true>(t1, t2, P.fp_morph, c);
    __syncthreads();
    morph_any<false>(t2, t3, P.fp_morph, c);
    __syncthreads();
    morph_any<false>(t3, t2, P.fp_morph, c);
    __syncthreads();
    morph_any<true>(t2, t1, P.fp_morph, c);
    __syncthreads();
    // t1 = opened
    LFX_CTX_TICK(c, 4)
    const bool found = largest_external(t1, t2, out, info8, c);
    LFX_CTX_TICK(c, 5)
    if (!found) {
        plane_copy(out, t1, c);
        __syncthreads();
    }
    return found;
}

// cv::getThreshVal_Otsu_8u on a 256-bin histogram (float64, no FMA contraction).
__device__ int otsu_threshold(const int* h, int n) {
    const double scale = __ddiv_rn(1.0, (double)n);
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)h[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    const double eps = (double)FLT_EPSILON, one_m_eps = 1.0 - (double)FLT_EPSILON;
    for (int i = 0; i < 256; ++i) {
        const double p_i = __dmul_rn((double)h[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        const double q2 = __dadd_rn(1.0, -q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > one_m_eps) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
        const double mu2 = __ddiv_rn(__dadd_rn(mu, -__dmul_rn(q1, mu1)), q2);
        const double dm = __dadd_rn(mu1, -mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), dm), dm);
        if (sigma > max_sigma) {
            max_sigma = sigma;
            max_val = i;
        }
    }
    return max_val;
}

// Pixel pass over the RGB image.  PASS 0: strategy predicate -> p0 (strategies 0/1) and brown
// predicate -> pb.  PASS 1: histogram of HSV channel `chan` into c.s_hist.  PASS 2: p0 = chan > thr
// (light) or chan <= thr (dark).
template <int PASS, class C>
__device__ void pixel_pass(const uint8_t* img, uint8_t* s_stage, const HsvLut* hsv, const LabLut* lab, uint32_t* p0,
                           uint32_t* pb, int chan, int thr, bool dark, const MaskParams& P, C& c) {
    const int rb = c.W * 3;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int y0 = 0; y0 < c.H; y0 += P.stage_rows) {
        const int nrows = min(P.stage_rows, c.H - y0);
        block_load_bytes(s_stage, img + (size_t)y0 * rb, nrows * rb);
        __syncthreads();
        for (int item = wid; item < nrows * c.WPR; item += MT / 32) {
            const int ry = item / c.WPR, w = item - ry * c.WPR;
            const int x = w * 32 + lane;
            bool b0 = false, b1 = false;
            if (x < c.W) {
                const uint8_t* px = s_stage + ry * rb + x * 3;
                const int r = px[0], g = px[1], b = px[2];
                int h, s, v;
                if (PASS == 0) {
                    rgb2hsv(r, g, b, hsv, h, s, v);
                    if (P.cfg.strategy == 0) b0 = (h >= P.cfg.green_lo) && (h <= P.cfg.green_hi) && (s >= 40);
                    int L = 0, A = 0, Bv = 0;
                    if (P.cfg.strategy == 1 || P.cfg.use_lab_brown) rgb2lab(r, g, b, lab, L, A, Bv);
                    if (P.cfg.strategy == 1) b0 = (A <= 135) && (Bv >= 115) && (Bv <= 170);
                    b1 = P.cfg.use_lab_brown
                             ? ((A >= P.cfg.lab_a_min) && (Bv >= P.cfg.lab_b_min))
                             : ((h >= P.cfg.brown_lo) && (h <= P.cfg.brown_hi) && (s >= P.cfg.brown_s_min) &&
                                (v <= P.cfg.brown_v_max));
                } else {
                    // PlantCV's rgb2gray_hsv reads the array as BGR: only the hue channel differs
                    if (chan == 0)
                        rgb2hsv(b, g, r, hsv, h, s, v);
                    else
                        rgb2hsv(r, g, b, hsv, h, s, v);
                    const int val = chan == 0 ? h : (chan == 1 ? s : v);
                    if (PASS == 1) atomicAdd(&c.s_hist[val], 1);
                    if (PASS == 2) b0 = dark ? (val <= thr) : (val > thr);
                }
            }
            if (PASS != 1) {
                const uint32_t m0 = __ballot_sync(0xffffffffu, b0);
                const uint32_t m1 = __ballot_sync(0xffffffffu, b1);
                if (lane == 0) {
                    const int idx = (y0 + ry) * c.WPR + w;
                    if (PASS == 2 || P.cfg.strategy <= 1) p0[idx] = m0;
                    if (PASS == 0) pb[idx] = m1;
                }
            }
        }
        __syncthreads();
    }
}

// raw mask bytes -> bit plane
template <class C>
__device__ void bytes_to_plane(const uint8_t* raw, uint32_t* p, const C& c) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int item = wid; item < c.NW; item += MT / 32) {
        const int y = item / c.WPR, w = item - y * c.WPR;
        const int x = w * 32 + lane;
        const bool on = (x < c.W) && (__ldg(raw + (size_t)y * c.W + x) > 0);
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) p[item] = m;
    }
}

template <class C>
__device__ void plane_to_bytes(const uint32_t* p, uint8_t* mask, const C& c) {
    if ((c.W & 3) == 0) {
        const int gpr = c.W >> 2;
        for (int i = threadIdx.x; i < c.H * gpr; i += MT) {
            const int y = i / gpr, gx = i - y * gpr;
            const int x = gx * 4;
            const uint32_t bits = (p[y * c.WPR + (x >> 5)] >> (x & 31)) & 0xF;
            const uint32_t v = ((bits & 1) ? 0xFFu : 0u) | ((bits & 2) ? 0xFF00u : 0u) | ((bits & 4) ? 0xFF0000u : 0u) |
                               ((bits & 8) ? 0xFF000000u : 0u);
            reinterpret_cast<uint32_t*>(mask)[(size_t)y * gpr + gx] = v;
        }
    } else {
        for (int i = threadIdx.x; i < c.H * c.W; i += MT) {
            const int y = i / c.W, x = i - y * c.W;
            mask[i] = ((p[y * c.WPR + (x >> 5)] >> (x & 31)) & 1) ? 255 : 0;
        }
    }
}


// plane -> 0/255 bytes, 16 pixels (one 128-bit store) per thread; needs W % 16 == 0 and a 16-byte aligned mask.
template <class C>
__device__ void plane_to_bytes16(const uint32_t* p, uint8_t* mask, const C& c) {
    uint4* out = reinterpret_cast<uint4*>(mask);
    for (int i = threadIdx.x; i < c.NW * 2; i += MT) {
        const uint32_t bits = (p[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
        uint4 v;
        // nibble -> 4 bytes: spread bit k to bit 8k (x 0x00204081, disjoint partial products), then x 0xFF
        v.x = (((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        v.y = ((((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        v.z = ((((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
        v.w = (((bits >> 12) * 0x00204081u) & 0x01010101u) * 0xFFu;
        __stcs(&out[i], v);   // streaming store: written once, never re-read by this kernel
    }
}

template <class C>
__device__ void otsu_plane(const uint8_t* img, uint8_t* s_stage, const HsvLut* hsv, uint32_t* p0, int chan, bool dark,
                           const MaskParams& P, C& c) {
    for (int i = threadIdx.x; i < 256; i += MT) c.s_hist[i] = 0;
    __syncthreads();
    pixel_pass<1>(img, s_stage, hsv, nullptr, p0, nullptr, chan, 0, dark, P, c);
    if (threadIdx.x == 0) c.s_tmp[33] = otsu_threshold(c.s_hist, c.H * c.W);
    __syncthreads();
    const int thr = c.s_tmp[33];
    pixel_pass<2>(img, s_stage, hsv, nullptr, p0, nullptr, chan, thr, dark, P, c);
}


// Everything after the raw candidate of make_mask (mask.py:548-582): _postprocess_mask, the Otsu
// fallback of _handle_fallback_and_extension (:495-523) and _extend_mask_with_brown_regions
// (:335-392).  P0 = raw candidate, PB = brown predicate; the result lands in PR / s_info.
// `simg` is the image in global memory (only read by the rare Otsu fallback, staged via s_stage).
template <class C>
__device__ void mask_finish(const uint8_t* simg, uint8_t* s_stage, const HsvLut* s_hsv, uint32_t* P0, uint32_t* PB,
                            uint32_t* PR, uint32_t* T1, uint32_t* T2, uint32_t* T3, int* s_info, int* s_info2,
                            const MaskParams& P, C& c) {
    bool found = postprocess(P0, PR, T1, T2, T3, s_info, P, c);
    if (P.mode == 0) {
        // _find_best_mask rejects a lone candidate only when cnt is None or contourArea <= 1
        if (!found || s_info[5] <= 2) {
            c.status |= 1;
            otsu_plane(simg, s_stage, s_hsv, P0, P.cfg.fallback_channel, false, P, c);
            __syncthreads();
            found = postprocess(P0, PR, T1, T2, T3, s_info, P, c);
        }
        if (P.cfg.extend_brown) {
            // _extend_mask_with_brown_regions (mask.py:335-392)
            // search_area = dilate^2(best) only constrains brown pixels OUTSIDE the best mask (dilation is extensive:
            // best is a subset of search_area).  When every brown-predicate pixel already lies inside the best mask,
            // brown & search_area == brown and the two 20x20 dilations can be skipped.
            int outside = 0;
            for (int i = threadIdx.x; i < c.NW; i += MT) outside |= ((PB[i] & ~PR[i]) != 0u);
            if (!__syncthreads_or(outside)) {
                for (int i = threadIdx.x; i < c.NW; i += MT) T2[i] = 0xFFFFFFFFu;   // stands for search_area in the AND below
                __syncthreads();
            } else if (c.hp[0] && P.search_is_e20) {
                dilate_ellipse20(PR, T1, c);
                dilate_ellipse20(T1, T2, c);
            } else {
                morph_any<true>(PR, T1, P.fp_search, c);
                __syncthreads();
                morph_any<true>(T1, T2, P.fp_search, c);
                __syncthreads();
            }
            LFX_CTX_TICK(c, 6)
            int any = 0;
            for (int i = threadIdx.x; i < c.NW; i += MT) {
                const uint32_t v = PB[i] & T2[i];
                T1[i] = v;
                any |= (v != 0u);
            }
            // No brown pixel near the leaf (or none survives open/close): the extended mask equals the best
            // mask, a single filled component whose contour statistics are already in s_info -- the
            // reference recomputes the same contour (mask.py:383-392), so the result is unchanged.
            if (!__syncthreads_or(any)) return;
            morph_any<false>(T1, T2, P.fp_brown, c);  // open
            __syncthreads();
            morph_any<true>(T2, T1, P.fp_brown, c);
            __syncthreads();
            morph_any<true>(T1, T2, P.fp_brown, c);  // close
            __syncthreads();
            morph_any<false>(T2, T1, P.fp_brown, c);
            // Cleaned brown regions that add no pixel outside the best mask (spots ON the leaf, the usual case) leave
            // the extended mask equal to the best mask whatever the area filter keeps: same early exit as above.
            any = 0;
            for (int i = threadIdx.x; i < c.NW; i += MT) any |= ((T1[i] & ~PR[i]) != 0u);   // own words only: no barrier needed yet
            if (!__syncthreads_or(any)) return;
            LFX_CTX_TICK(c, 7)
            ccl<8>(T1, c);
            measure_area(c);
            plane_copy(T3, PR, c);  // ext = best | filtered brown
            __syncthreads();
            keep_area_ge(T3, P.cfg.brown_min_area_px, c);
            LFX_CTX_TICK(c, 8)
            // contour of the extended mask; the returned mask is the UNFILLED union
            const bool f2 = largest_external(T3, T1, T2, s_info2, c);
            LFX_CTX_TICK(c, 9)
            if (f2) {
                plane_copy(PR, T3, c);
                if (threadIdx.x < 8) s_info[threadIdx.x] = s_info2[threadIdx.x];
            } else {
                if (threadIdx.x < 8) s_info[threadIdx.x] = 0;  // (best_mask, None)
            }
            __syncthreads();
        }
    }
}

}  // namespace
