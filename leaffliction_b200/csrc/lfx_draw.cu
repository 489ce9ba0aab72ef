// Overlay rasterisers: the OpenCV drawing calls of apply_analyze_filter (analyze.py:37-122) and
// apply_roi_filter (roi.py:43-44) on a resident batch, one thread block per image, bit-identical to cv2 4.13
// (drawing.cpp: Line2 / FillConvexPoly / ThickLine / Circle / LineAA / PolyLine in 16.16 fixed point).
//
// Order is part of the result (later primitives overwrite or blend with earlier ones), so every image is drawn in the
// reference's order.  Inside one step the work is spread over the block where the writes cannot collide:
//   * the contour's 2-px segments and the rectangle's edges are one colour -> one thread per segment, any order;
//   * one anti-aliased line touches three distinct pixels per step along its major axis -> one thread per step;
//   * markers, circles, PCA axes: a handful of pixels, drawn by thread 0.
//   * a 2-px line is one colour as well: its outline steps and fill rows are closed forms of OpenCV's incremental DDA /
//     scan conversion, so the block shares them for the long lines (PCA axes, marker).
// HBM traffic is the copy rgb -> overlay (6N bytes per image) plus the edge / mask planes (2N); everything drawn
// afterwards hits lines the block has just written (L2).
#include <math.h>

#include "lfx_common.cuh"

namespace {

typedef long long i64;
constexpr int XY_SHIFT = 16;
constexpr i64 XY_ONE = 1 << XY_SHIFT;

struct Img {
    uint8_t* p;
    int H, W;
};
struct Pt {
    i64 x, y;
};

__constant__ uint8_t c_filter[64] = {
    168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
    254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
    158, 149, 140, 131, 122, 114, 105, 97,  89,  82,  75,  68,  62,  56,  50,  45,
    40,  36,  32,  28,  25,  22,  19,  16,  14,  12,  11,  9,   8,   7,   5,   5};
__constant__ uint8_t c_slope[32] = {181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
                                    203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254};

__device__ __forceinline__ void put(const Img& im, int x, int y, uint32_t col) {
    uint8_t* t = im.p + ((size_t)y * im.W + x) * 3;
    t[0] = (uint8_t)(col & 255);
    t[1] = (uint8_t)((col >> 8) & 255);
    t[2] = (uint8_t)((col >> 16) & 255);
}
__device__ __forceinline__ void put_chk(const Img& im, int x, int y, uint32_t col) {
    if ((unsigned)x < (unsigned)im.W && (unsigned)y < (unsigned)im.H) put(im, x, y, col);
}
__device__ __forceinline__ void hline(const Img& im, int y, int xa, int xb, uint32_t col) {
    for (int x = xa; x <= xb; ++x) put(im, x, y, col);
}

// cv::clipLine(Size2l, Point2l&, Point2l&): the cut points are computed in double and truncated.
__device__ bool clip_line(i64 width, i64 height, Pt& a, Pt& b) {
    const i64 right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
    i64 &x1 = a.x, &y1 = a.y, &x2 = b.x, &y2 = b.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        i64 t;
        if (c1 & 12) {
            t = c1 < 8 ? 0 : bottom;
            x1 += (i64)__ddiv_rn(__dmul_rn((double)(t - y1), (double)(x2 - x1)), (double)(y2 - y1));
            y1 = t;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            t = c2 < 8 ? 0 : bottom;
            x2 += (i64)__ddiv_rn(__dmul_rn((double)(t - y2), (double)(x2 - x1)), (double)(y2 - y1));
            y2 = t;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                t = c1 == 1 ? 0 : right;
                y1 += (i64)__ddiv_rn(__dmul_rn((double)(t - x1), (double)(y2 - y1)), (double)(x2 - x1));
                x1 = t;
                c1 = 0;
            }
            if (c2) {
                t = c2 == 1 ? 0 : right;
                y2 += (i64)__ddiv_rn(__dmul_rn((double)(t - x2), (double)(y2 - y1)), (double)(x2 - x1));
                x2 = t;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// drawing.cpp Line2: 8-connected DDA between 16.16 end points (the outline FillConvexPoly draws for shift != 0).
// Steps t0, t0 + ts, ... are drawn by the caller (a lone thread passes 0, 1; a block passes threadIdx.x, blockDim.x: the
// step-s pixel is the closed form of the DDA, y1 + s * y_step, so the steps are independent).
__device__ __noinline__ void line2(const Img& im, Pt p1, Pt p2, uint32_t col, int t0, int ts) {
    if (!clip_line((i64)im.W << XY_SHIFT, (i64)im.H << XY_SHIFT, p1, p2)) return;
    i64 dx = p2.x - p1.x, dy = p2.y - p1.y;
    const i64 ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    i64 step;
    int ecount;
    if (ax > ay) {
        if (dx < 0) {
            dy = -dy;
            Pt t = p1; p1 = p2; p2 = t;
        }
        step = (dy << XY_SHIFT) / (ax | 1);
        ecount = (int)((p2.x - p1.x) >> XY_SHIFT);
    } else {
        if (dy < 0) {
            dx = -dx;
            Pt t = p1; p1 = p2; p2 = t;
        }
        step = (dx << XY_SHIFT) / (ay | 1);
        ecount = (int)((p2.y - p1.y) >> XY_SHIFT);
    }
    p1.x += XY_ONE >> 1;
    p1.y += XY_ONE >> 1;
    if (t0 == 0) put_chk(im, (int)((p2.x + (XY_ONE >> 1)) >> XY_SHIFT), (int)((p2.y + (XY_ONE >> 1)) >> XY_SHIFT), col);
    if (ax > ay) {
        const int x0 = (int)(p1.x >> XY_SHIFT);
        for (int s = t0; s <= ecount; s += ts) put_chk(im, x0 + s, (int)((p1.y + step * s) >> XY_SHIFT), col);
    } else {
        const int y0 = (int)(p1.y >> XY_SHIFT);
        for (int s = t0; s <= ecount; s += ts) put_chk(im, (int)((p1.x + step * s) >> XY_SHIFT), y0 + s, col);
    }
}

// drawing.cpp FillConvexPoly(LINE_8, shift = XY_SHIFT) for the 4-vertex polygon of a thick segment.  OpenCV walks the rows
// top to bottom and adds each chain's slope once per row; between two vertex events the row-r edge position is
// x + slope * (r - r0), so the rows of one stretch are independent: the caller's threads (t0, ts as in line2) share them.
__device__ void fill_convex_poly4(const Img& im, const Pt* v, uint32_t col, int t0, int ts) {
    constexpr int npts = 4;
    const i64 delta = XY_ONE >> 1;
    i64 xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    int imin = 0;
    Pt p0 = v[npts - 1];
#pragma unroll
    for (int i = 0; i < npts; ++i) {
        const Pt p = v[i];
        if (p.y < ymin) {
            ymin = p.y;
            imin = i;
        }
        ymax = p.y > ymax ? p.y : ymax;
        xmax = p.x > xmax ? p.x : xmax;
        xmin = p.x < xmin ? p.x : xmin;
        line2(im, p0, p, col, t0, ts);
        p0 = p;
    }
    xmin = (xmin + delta) >> XY_SHIFT;
    xmax = (xmax + delta) >> XY_SHIFT;
    ymin = (ymin + delta) >> XY_SHIFT;
    ymax = (ymax + delta) >> XY_SHIFT;
    if ((int)xmax < 0 || (int)ymax < 0 || (int)xmin >= im.W || (int)ymin >= im.H) return;
    if (ymax > im.H - 1) ymax = im.H - 1;
    // the four vertices by selects and the two chains unrolled: nothing here is indexed at run time, so nothing lives in local memory
    auto vx = [&](int k) { return k == 0 ? v[0].x : k == 1 ? v[1].x : k == 2 ? v[2].x : v[3].x; };
    auto vy = [&](int k) { return k == 0 ? v[0].y : k == 1 ? v[1].y : k == 2 ? v[2].y : v[3].y; };
    int e_idx[2] = {imin, imin}, e_ye[2] = {(int)ymin, (int)ymin};
    i64 e_x[2] = {-XY_ONE, -XY_ONE}, e_dx[2] = {0, 0};
    int edges = npts;
    int y = (int)ymin;
    while (true) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (y >= e_ye[i]) {
                int idx0 = e_idx[i];
                const int di = i == 0 ? 1 : npts - 1;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    const int ty = (int)((vy(idx) + delta) >> XY_SHIFT);
                    if (ty > y) {
                        const i64 xs = vx(idx0), xe = vx(idx);
                        e_ye[i] = ty;
                        e_dx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (i64)(ty - y));
                        e_x[i] = xs;
                        e_idx[i] = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        // rows y .. yn-1: no vertex event (both chains end at or after yn)
        int yn = min(e_ye[0], e_ye[1]);
        if (yn <= y) yn = y + 1;
        if (yn > (int)ymax + 1) yn = (int)ymax + 1;
        for (int r = max(y, 0) + t0; r < yn; r += ts) {
            const i64 xa = e_x[0] + e_dx[0] * (r - y), xb = e_x[1] + e_dx[1] * (r - y);
            int xx1 = (int)(((xa > xb ? xb : xa) + delta) >> XY_SHIFT);
            int xx2 = (int)(((xa > xb ? xa : xb) + delta) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < im.W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= im.W) xx2 = im.W - 1;
                hline(im, r, xx1, xx2, col);
            }
        }
        e_x[0] += e_dx[0] * (yn - y);
        e_x[1] += e_dx[1] * (yn - y);
        y = yn;
        if (y > (int)ymax) break;
    }
}

// drawing.cpp Circle(fill = 1): midpoint circle, four clipped spans per step.
__device__ void circle_filled(const Img& im, int cx, int cy, int radius, uint32_t col) {
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    auto span = [&](int y, int xa, int xb) {
        if ((unsigned)y < (unsigned)im.H) hline(im, y, max(xa, 0), min(xb, im.W - 1), col);
    };
    while (dx >= dy) {
        span(cy - dy, cx - dx, cx + dx);
        span(cy + dy, cx - dx, cx + dx);
        span(cy - dx, cx - dy, cx + dy);
        span(cy + dx, cx - dy, cx + dy);
        ++dy;
        err += plus;
        plus += 2;
        const int m = (err <= 0) - 1;
        err -= minus & m;
        dx += m;
        minus -= m & 2;
    }
}

// drawing.cpp ThickLine (LINE_8, shift 0, thickness >= 2) as cv::line / PolyLine reach it: the segment is first clipped to
// the image grown by `thickness`, then drawn as the polygon around it plus round caps (one out-of-line copy: inlined at its
// six call sites the kernel was 13 k instructions and a fifth of its stalls were instruction fetches) (flags bit 0: at p0, bit 1: at p1).
// One colour, so the pixels can be written in any order: a lone thread draws everything (t0 = 0, ts = 1), a block shares the
// polygon's outline steps and rows (t0 = threadIdx.x, ts = blockDim.x; the caps go to thread 0).
__device__ __noinline__ void thick_line(const Img& im, int x0, int y0, int x1, int y1, uint32_t col, int thickness, int flags, int t0 = 0, int ts = 1) {
    Pt a = {(i64)x0 + thickness, (i64)y0 + thickness}, b = {(i64)x1 + thickness, (i64)y1 + thickness};
    if (!clip_line((i64)im.W + 2 * thickness, (i64)im.H + 2 * thickness, a, b)) return;
    Pt q0 = {(a.x - thickness) << XY_SHIFT, (a.y - thickness) << XY_SHIFT};
    Pt q1 = {(b.x - thickness) << XY_SHIFT, (b.y - thickness) << XY_SHIFT};
    const double inv = 1.0 / 65536.0;
    const double dx = (double)(q0.x - q1.x) * inv, dy = (double)(q1.y - q0.y) * inv;
    double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    const int odd = thickness & 1;
    const i64 th = (i64)thickness << (XY_SHIFT - 1);
    if (fabs(r) > 2.220446049250313e-16) {
        r = __ddiv_rn((double)th + odd * 32768.0, __dsqrt_rn(r));
        const i64 dpx = __double2ll_rn(__dmul_rn(dy, r)), dpy = __double2ll_rn(__dmul_rn(dx, r));
        const Pt pt[4] = {{q0.x + dpx, q0.y + dpy}, {q0.x - dpx, q0.y - dpy}, {q1.x - dpx, q1.y - dpy}, {q1.x + dpx, q1.y + dpy}};
        fill_convex_poly4(im, pt, col, t0, ts);
    }
    const int rad = (int)((th + (XY_ONE >> 1)) >> XY_SHIFT);
    if (t0 != 0) return;
    if (flags & 1) circle_filled(im, (int)((q0.x + (XY_ONE >> 1)) >> XY_SHIFT), (int)((q0.y + (XY_ONE >> 1)) >> XY_SHIFT), rad, col);
    if (flags & 2) circle_filled(im, (int)((q1.x + (XY_ONE >> 1)) >> XY_SHIFT), (int)((q1.y + (XY_ONE >> 1)) >> XY_SHIFT), rad, col);
}

__device__ __forceinline__ void blend_aa(const Img& im, int x, int y, int a, int cr, int cg, int cb) {
    uint8_t* t = im.p + ((size_t)y * im.W + x) * 3;
    int v = t[0];
    v += ((cr - v) * a + 127) >> 8;
    v += ((cr - v) * a + 127) >> 8;
    t[0] = (uint8_t)v;
    v = t[1];
    v += ((cg - v) * a + 127) >> 8;
    v += ((cg - v) * a + 127) >> 8;
    t[1] = (uint8_t)v;
    v = t[2];
    v += ((cb - v) * a + 127) >> 8;
    v += ((cb - v) * a + 127) >> 8;
    t[2] = (uint8_t)v;
}

// drawing.cpp LineAA between integer pixel end points, called by EVERY thread of the block (the set-up is recomputed per
// thread, the steps along the major axis are dealt out).  The caller synchronises the block afterwards.
__device__ __noinline__ void line_aa_block(const Img& im, int px0, int py0, int px1, int py1, uint32_t col) {
    Pt p1 = {(i64)px0 << XY_SHIFT, (i64)py0 << XY_SHIFT}, p2 = {(i64)px1 << XY_SHIFT, (i64)py1 << XY_SHIFT};
    if (!clip_line((i64)im.W << XY_SHIFT, (i64)im.H << XY_SHIFT, p1, p2)) return;
    i64 dx = p2.x - p1.x, dy = p2.y - p1.y;
    const i64 ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
    const bool xmajor = ax > ay;
    i64 step, i, j, minor0;
    int ecount, major0, slope;
    if (xmajor) {
        if (dx < 0) {
            dy = -dy;
            Pt t = p1; p1 = p2; p2 = t;
        }
        step = (dy << XY_SHIFT) / (ax | 1);
        p2.x += XY_ONE;
        ecount = (int)((p2.x >> XY_SHIFT) - (p1.x >> XY_SHIFT));
        j = -(p1.x & (XY_ONE - 1));
        p1.y += ((step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        i = (p1.x >> (XY_SHIFT - 7)) & 0x78;
        j = (p2.x >> (XY_SHIFT - 7)) & 0x78;
        major0 = (int)(p1.x >> XY_SHIFT);
        minor0 = p1.y;
    } else {
        if (dy < 0) {
            dx = -dx;
            Pt t = p1; p1 = p2; p2 = t;
        }
        step = (dx << XY_SHIFT) / (ay | 1);
        p2.y += XY_ONE;
        ecount = (int)((p2.y >> XY_SHIFT) - (p1.y >> XY_SHIFT));
        j = -(p1.y & (XY_ONE - 1));
        p1.x += ((step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        i = (p1.y >> (XY_SHIFT - 7)) & 0x78;
        j = (p2.y >> (XY_SHIFT - 7)) & 0x78;
        major0 = (int)(p1.y >> XY_SHIFT);
        minor0 = p1.x;
    }
    slope = (int)((step >> (XY_SHIFT - 5)) & 0x3f);
    slope ^= step < 0 ? 0x3f : 0;
    slope = (slope & 0x20) ? 0x100 : c_slope[slope];
    int ep1, ep2, ep4, ep5, ep6, ep7;   // end-point correction table (entries 0 = 0, 3 = 1, 8 = slope), read through selects
    {
        const int ii = (int)i, jj = (int)j;
        const int t0 = slope << 7, t1 = ((0x78 - ii) | 4) * slope, t2 = (jj | 4) * slope;
        ep1 = (((((jj - ii) & 0x78) | 4) * slope) >> 8) & 0x1ff;
        ep2 = (t1 >> 8) & 0x1ff;
        ep4 = (((((jj - ii) + 0x80) | 4) * slope) >> 8) & 0x1ff;
        ep5 = ((t1 + t0) >> 8) & 0x1ff;
        ep6 = (t2 >> 8) & 0x1ff;
        ep7 = ((t2 + t0) >> 8) & 0x1ff;
    }
    auto ep = [&](int k) {
        return k == 8 ? slope : k == 0 ? 0 : (k == 1 || k == 3) ? ep1 : k == 2 ? ep2 : k == 4 ? ep4 : k == 5 ? ep5 : k == 6 ? ep6 : ep7;
    };
    const int cr = col & 255, cg = (col >> 8) & 255, cb = (col >> 16) & 255;
    const int major_lim = xmajor ? im.W : im.H, minor_lim = xmajor ? im.H : im.W;
    for (int s = threadIdx.x; s <= ecount; s += blockDim.x) {
        const int mj = major0 + s;
        if ((unsigned)mj >= (unsigned)major_lim) continue;
        const i64 mcoord = minor0 + step * s;
        const int scount = s, ec = ecount - s;
        const int m0 = (int)(mcoord >> XY_SHIFT) - 1;
        const int ep_corr = ep((((scount >= 2) + 1) & (scount | 2)) * 3 + (((ec >= 2) + 1) & (ec | 2)));
        const int dist = (int)((mcoord >> (XY_SHIFT - 5)) & 31);
        const int f[3] = {c_filter[dist + 32], c_filter[dist], c_filter[63 - dist]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int a = ((ep_corr * f[k]) >> 8) & 0xff;
            const int mn = m0 + k;
            if ((unsigned)mn < (unsigned)minor_lim) {
                if (xmajor) blend_aa(im, mj, mn, a, cr, cg, cb);
                else blend_aa(im, mn, mj, a, cr, cg, cb);
            }
        }
    }
}

__device__ __forceinline__ void block_copy_image(uint8_t* dst, const uint8_t* src, size_t nbytes) {
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const size_t n16 = nbytes >> 4;
        size_t k = threadIdx.x;
        for (; k + 3 * (size_t)blockDim.x < n16; k += 4 * (size_t)blockDim.x) {   // four loads in flight per thread
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = ld_stream16(src + (k + (size_t)j * blockDim.x) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(dst)[k + (size_t)j * blockDim.x] = v[j];
        }
        for (; k < n16; k += blockDim.x) reinterpret_cast<uint4*>(dst)[k] = ld_stream16(src + k * 16);
        for (size_t k = (n16 << 4) + threadIdx.x; k < nbytes; k += blockDim.x) dst[k] = src[k];
    } else {
        for (size_t k = threadIdx.x; k < nbytes; k += blockDim.x) dst[k] = src[k];
    }
}

__host__ __device__ constexpr uint32_t rgb_u32(int r, int g, int b) { return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16); }

// apply_analyze_filter's overlay (analyze.py:37-122) for one image per block.
__global__ void __launch_bounds__(64, 16) k_analyze_overlay(const uint8_t* __restrict__ rgb, const int32_t* __restrict__ points,
                                                          const int32_t* __restrict__ counts, const int32_t* __restrict__ rec_i,
                                                          const int32_t* __restrict__ hull, const uint8_t* __restrict__ edges,
                                                          const uint8_t* __restrict__ mask, uint8_t* overlay, int H, int W,
                                                          int max_pts, int max_hull) {
    const int img = blockIdx.x;
    const size_t npx = (size_t)H * W;
    Img im = {overlay + (size_t)img * npx * 3, H, W};
    block_copy_image(im.p, rgb + (size_t)img * npx * 3, npx * 3);
    const int32_t* ri = rec_i + (size_t)img * 24;
    const int n = counts[img];
    if (ri[0] == 0 || n <= 0 || n > max_pts) return;   // no contour: the image is returned as it is (the text banner of analyze.py:29 is not drawn)
    __syncthreads();
    const int32_t* p = points + (size_t)img * max_pts * 2;
    // :40  drawContours(thickness 2): PolyLine, segment k joins vertex k-1 -> k, round cap at k
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int kp = k == 0 ? n - 1 : k - 1;
        thick_line(im, p[2 * kp], p[2 * kp + 1], p[2 * k], p[2 * k + 1], rgb_u32(255, 0, 0), 2, 2);
    }
    __syncthreads();
    const int cx = ri[2], cy = ri[3];
    constexpr uint32_t yellow = rgb_u32(255, 255, 0);
    const int t0 = threadIdx.x, ts = blockDim.x;
    // :50-57 drawMarker(MARKER_CROSS, 14, 2); :65-75 the four extreme points: filled circle, then the anti-aliased ray
    // (marker and circles are one colour; a ray blends with what is under it, so it waits for them)
    thick_line(im, cx - 7, cy, cx + 7, cy, yellow, 2, 3, t0, ts);
    thick_line(im, cx, cy - 7, cx, cy + 7, yellow, 2, 3, t0, ts);
    for (int e = 0; e < 4; ++e) {
        const int ex = ri[4 + 2 * e], ey = ri[5 + 2 * e];
        if (threadIdx.x == 0) circle_filled(im, ex, ey, 3, yellow);
        __syncthreads();
        line_aa_block(im, cx, cy, ex, ey, yellow);
        __syncthreads();
    }
    // :77-85 hull polyline, anti-aliased, in cv2.convexHull's vertex order: the device hull (counter-clockwise, starting
    // at the top-most vertex) rotated to start at the hull vertex met LAST along the contour
    const int nh = ri[12];
    if (nh > 0 && nh <= max_hull) {
        const int32_t* hp = hull + (size_t)img * max_hull * 2;
        __shared__ int s_best;
        if (threadIdx.x == 0) s_best = -1;
        __syncthreads();
        if (nh >= 3) {
            for (int k = threadIdx.x; k < nh; k += blockDim.x) {
                const int hx = hp[2 * k], hy = hp[2 * k + 1];
                int last = -1;
                for (int t = n - 1; t >= 0; --t)
                    if (p[2 * t] == hx && p[2 * t + 1] == hy) {
                        last = t;
                        break;
                    }
                if (last >= 0) atomicMax(&s_best, last * 1024 + k);
            }
            __syncthreads();
        }
        const int start = (nh >= 3 && s_best >= 0) ? (s_best & 1023) : 0;
        for (int k = 0; k < nh; ++k) {
            int a = start + k - 1, b = start + k;
            if (k == 0) a = start + nh - 1;
            a %= nh;
            b %= nh;
            line_aa_block(im, hp[2 * a], hp[2 * a + 1], hp[2 * b], hp[2 * b + 1], rgb_u32(0, 255, 0));
            __syncthreads();
        }
    }
    // :88-112 PCA axes, 2 px: each line shared by the block, the second over the first
    thick_line(im, ri[14], ri[15], ri[16], ri[17], yellow, 2, 3, t0, ts);
    __syncthreads();
    thick_line(im, ri[18], ri[19], ri[20], ri[21], rgb_u32(255, 0, 255), 2, 3, t0, ts);
    __syncthreads();
    // :115-122 vein edges inside the mask, cyan.  Edge pixels are sparse: 16 pixels per load, four loads in flight, the mask is
    // only read where an edge byte is set.
    if (edges && mask) {
        const uint8_t* e = edges + (size_t)img * npx;
        const uint8_t* m = mask + (size_t)img * npx;
        auto paint = [&](size_t k) {
            uint8_t* t = im.p + k * 3;
            t[0] = 0;
            t[1] = 255;
            t[2] = 255;
        };
        size_t done = 0;
        if (((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(m)) & 15) == 0) {
            const size_t n16 = npx >> 4;
            for (size_t k0 = threadIdx.x; k0 < n16; k0 += 4 * (size_t)blockDim.x) {
                uint4 ev[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const size_t k = k0 + (size_t)j * blockDim.x;
                    ev[j] = k < n16 ? __ldg(reinterpret_cast<const uint4*>(e) + k) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((ev[j].x | ev[j].y | ev[j].z | ev[j].w) == 0) continue;
                    const size_t k = k0 + (size_t)j * blockDim.x;
                    const uint4 mv = __ldg(reinterpret_cast<const uint4*>(m) + k);
                    const uint32_t ew[4] = {ev[j].x, ev[j].y, ev[j].z, ev[j].w}, mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                    for (int w = 0; w < 4; ++w)
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (((ew[w] >> (8 * b)) & 255) && ((mw[w] >> (8 * b)) & 255)) paint(k * 16 + w * 4 + b);
                }
            }
            done = n16 << 4;
        }
        for (size_t k = done + threadIdx.x; k < npx; k += blockDim.x)
            if (e[k] && m[k]) paint(k);
    }
}

// roi.py:44: vis = rgb.copy(); cv2.rectangle(vis, (x, y), (x + w, y + h), colour, thickness): PolyLine over the four corners.
__global__ void __launch_bounds__(256, 4) k_draw_rectangles(const uint8_t* __restrict__ rgb, const int32_t* __restrict__ info,
                                                          uint8_t* vis, int H, int W, uint32_t col, int thickness) {
    const int img = blockIdx.x;
    const size_t npx = (size_t)H * W;
    Img im = {vis + (size_t)img * npx * 3, H, W};
    block_copy_image(im.p, rgb + (size_t)img * npx * 3, npx * 3);
    const int32_t* bi = info + (size_t)img * 8;
    if (bi[0] == 0 || bi[3] <= 0 || bi[4] <= 0) return;
    __syncthreads();
    // one colour, so the four edges need no order among themselves: each is shared by the whole block
    const int x0 = bi[1], y0 = bi[2], x1 = bi[1] + bi[3], y1 = bi[2] + bi[4];
    const int vx[4] = {x0, x1, x1, x0}, vy[4] = {y0, y0, y1, y1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int kp = (k + 3) & 3;
        thick_line(im, vx[kp], vy[kp], vx[k], vy[k], col, thickness, 2, threadIdx.x, blockDim.x);
    }
}

// A caller-supplied list of primitives per image, drawn in list order (each primitive shared by the block, a barrier between
// two primitives): the rasterisers above as a general op (cv2.line 2 px and wider / anti-aliased, filled cv2.circle,
// cv2.rectangle, cv2.drawMarker cross).
__global__ void __launch_bounds__(64, 16) k_draw_primitives(uint8_t* img, const int32_t* __restrict__ prims, const int32_t* __restrict__ counts,
                                                             int H, int W, int max_prims) {
    const int b = blockIdx.x;
    Img im = {img + (size_t)b * H * W * 3, H, W};
    const int n = min(counts[b], max_prims);
    const int t0 = threadIdx.x, ts = blockDim.x;
    for (int k = 0; k < n; ++k) {
        const int32_t* q = prims + ((size_t)b * max_prims + k) * 8;
        const int kind = q[0], x0 = q[1], y0 = q[2], x1 = q[3], y1 = q[4], size = q[6];
        const uint32_t col = (uint32_t)q[5];
        if (kind == LFX_DRAW_LINE && size >= 2) {
            thick_line(im, x0, y0, x1, y1, col, size, 3, t0, ts);
        } else if (kind == LFX_DRAW_LINE_AA) {
            line_aa_block(im, x0, y0, x1, y1, col);
        } else if (kind == LFX_DRAW_CIRCLE_FILLED && size >= 0) {
            if (t0 == 0) circle_filled(im, x0, y0, size, col);
        } else if (kind == LFX_DRAW_RECTANGLE && size >= 2) {
            const int vx[4] = {x0, x1, x1, x0}, vy[4] = {y0, y0, y1, y1};
#pragma unroll
            for (int e = 0; e < 4; ++e) thick_line(im, vx[(e + 3) & 3], vy[(e + 3) & 3], vx[e], vy[e], col, size, 2, t0, ts);
        } else if (kind == LFX_DRAW_MARKER_CROSS && size >= 2) {
            const int h = x1 / 2;
            thick_line(im, x0 - h, y0, x0 + h, y0, col, size, 3, t0, ts);
            thick_line(im, x0, y0 - h, x0, y0 + h, col, size, 3, t0, ts);
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int lfx_analyze_overlay(const uint8_t* rgb, const int32_t* points, const int32_t* counts, const int32_t* rec_i32,
                                   const int32_t* hull_points, const uint8_t* edges, const uint8_t* mask, uint8_t* overlay,
                                   int B, int H, int W, int max_pts, int max_hull, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(rgb && points && counts && rec_i32 && hull_points && overlay && B > 0 && H > 0 && W > 0 && max_pts > 0 && max_hull > 0,
                LFX_ERR_ARG, "analyze_overlay: bad arguments");
    LFX_REQUIRE((edges == nullptr) == (mask == nullptr), LFX_ERR_ARG, "analyze_overlay: edges and mask go together");
    LFX_REQUIRE(max_hull <= 1023 && H <= 16384 && W <= 16384, LFX_ERR_UNSUPPORTED, "analyze_overlay: max_hull <= 1023, image side <= 16384");
    LFX_REQUIRE(rgb != overlay, LFX_ERR_ARG, "analyze_overlay: in-place operation is not supported");
    // 64 threads per image: the drawing is a chain of short dependent steps, so many small blocks per SM beat few wide ones
    // (4096 x 256^2: 5.5 ms with 256 threads, 2.5 with 128, 2.1 with 64, 2.15 with 32; 64 registers -> 16 blocks per SM: 1.75 ms)
    k_analyze_overlay<<<B, 64, 0, (cudaStream_t)stream>>>(rgb, points, counts, rec_i32, hull_points, edges, mask, overlay, H, W, max_pts,
                                                          max_hull);
    return lfx_check_launch("analyze_overlay");
}

extern "C" int lfx_draw_rectangles(const uint8_t* rgb, const int32_t* info, uint8_t* vis, int B, int H, int W, uint32_t color_rgb,
                                   int thickness, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(rgb && info && vis && B > 0 && H > 0 && W > 0, LFX_ERR_ARG, "draw_rectangles: bad arguments");
    LFX_REQUIRE(thickness >= 2 && thickness <= 64 && H <= 16384 && W <= 16384, LFX_ERR_UNSUPPORTED,
                "draw_rectangles: thickness 2..64, image side <= 16384");
    LFX_REQUIRE(rgb != vis, LFX_ERR_ARG, "draw_rectangles: in-place operation is not supported");
    k_draw_rectangles<<<B, 256, 0, (cudaStream_t)stream>>>(rgb, info, vis, H, W, color_rgb, thickness);
    return lfx_check_launch("draw_rectangles");
}

extern "C" int lfx_draw_primitives(uint8_t* img, const int32_t* prims, const int32_t* counts, int B, int H, int W, int max_prims,
                                   lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(img && prims && counts && B > 0 && H > 0 && W > 0 && max_prims > 0, LFX_ERR_ARG, "draw_primitives: bad arguments");
    LFX_REQUIRE(H <= 16384 && W <= 16384, LFX_ERR_UNSUPPORTED, "draw_primitives: image side <= 16384");
    k_draw_primitives<<<B, 64, 0, (cudaStream_t)stream>>>(img, prims, counts, H, W, max_prims);
    return lfx_check_launch("draw_primitives");
}
