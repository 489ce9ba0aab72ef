// apply_roi_filter canvas (srcs/transform/filters/roi.py:20-46): crop the contour's bounding box
// from the white-masked image, letterbox with cv2.resize(INTER_AREA) into a zero canvas.  A box that fits the canvas is
// upscaled (2-tap fixed-point bilinear with area-mode offsets, OpenCV resize.cpp): k_roi; a box larger than the canvas
// (images bigger than roi_size) is area-averaged down: k_roi_down.
#include "lfx_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int ROI_ROWS = 64;  // canvas rows per block

struct Tap {
    int s;       // source index
    short a, b;  // weights (x2048) for s and s+1
};

// cv::resize area-mode 2-tap coefficients for destination index d (src -> dst upscale).
__device__ __forceinline__ Tap area_tap(int d, int src, int dst) {
    const double inv = __ddiv_rn((double)dst, (double)src);
    const double scale = __ddiv_rn(1.0, inv);
    int sx = (int)floor(__dmul_rn((double)d, scale));
    float fx = __double2float_rn(__dadd_rn((double)(d + 1), -__dmul_rn((double)(sx + 1), inv)));
    fx = fx <= 0.f ? 0.f : __fadd_rn(fx, -floorf(fx));
    if (sx < 0) {
        fx = 0.f;
        sx = 0;
    }
    if (sx >= src - 1) {
        fx = 0.f;
        sx = src - 1;
    }
    Tap t;
    t.s = sx;
    t.a = (short)__float2int_rn(__fmul_rn(__fadd_rn(1.f, -fx), 2048.f));
    t.b = (short)__float2int_rn(__fmul_rn(fx, 2048.f));
    return t;
}

// grid (ceil(RH/ROI_ROWS), B).  One canvas column per thread walking down the block's rows: the horizontal stage of a
// source row is computed once and carried to the next canvas row that uses it (upscale: consecutive canvas rows share
// source rows), the vertical stage is two IMAD.HI; columns / rows outside the resized box store zeros.  Pixels go
// straight to HBM as streaming byte stores (lane-contiguous: L2 merges the sectors).
__global__ void __launch_bounds__(THREADS) k_roi(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                 const int32_t* __restrict__ info, uint8_t* __restrict__ dst, int H, int W,
                                                 int RH, int RW) {
    __shared__ Tap s_yt[ROI_ROWS];
    const int img = blockIdx.y;
    const int r0 = blockIdx.x * ROI_ROWS;
    const int rows = min(ROI_ROWS, RH - r0);
    const int32_t* inf = info + (size_t)img * 8;
    const int found = inf[0], bx = inf[1], by = inf[2], bw = inf[3], bh = inf[4];
    uint8_t* dimg = dst + ((size_t)img * RH + r0) * RW * 3;
    int nw = 0, nh = 0, ox = 0, oy = 0;
    if (found && bw > 0 && bh > 0) {
        // scale = min(W / max(w,1), H / max(h,1)); nw = max(int(w*scale),1)   (roi.py:35-36, Python floats)
        const double sc = fmin(__ddiv_rn((double)RW, (double)max(bw, 1)), __ddiv_rn((double)RH, (double)max(bh, 1)));
        nw = max((int)__dmul_rn((double)bw, sc), 1), nh = max((int)__dmul_rn((double)bh, sc), 1);
        oy = (RH - nh) / 2, ox = (RW - nw) / 2;
        if (nw < bw || nh < bh) return;   // a box larger than the canvas shrinks: k_roi_down writes this image
    }
    for (int i = threadIdx.x; i < rows; i += THREADS) {
        const int d = r0 + i - oy;
        Tap t;
        t.s = -1; t.a = 0; t.b = 0;
        if (d >= 0 && d < nh) t = area_tap(d, bh, nh);
        s_yt[i] = t;
    }
    __syncthreads();
    const uint8_t* simg = src + (size_t)img * H * W * 3;
    const uint8_t* mimg = mask ? mask + (size_t)img * H * W : nullptr;
    for (int cc = threadIdx.x; cc < RW; cc += THREADS) {
        uint8_t* o = dimg + (size_t)cc * 3;
        const int cx = cc - ox;
        if ((unsigned)cx >= (unsigned)nw) {   // left / right band (or nothing found): zeros
            for (int r = 0; r < rows; ++r, o += RW * 3) {
                __stcs(o + 0, (uint8_t)0);
                __stcs(o + 1, (uint8_t)0);
                __stcs(o + 2, (uint8_t)0);
            }
            continue;
        }
        const Tap tx = area_tap(cx, bw, nw);
        const int x0 = bx + tx.s, x1 = bx + min(tx.s + 1, bw - 1);
        const uint32_t xa = (uint32_t)(uint16_t)tx.a, xb = (uint32_t)(uint16_t)tx.b;
        // horizontal stage of source row y (of the white-masked image, Transformation.py:451): HResizeLinear x2048, >> 4
        auto hrow = [&](int y, uint32_t& hr, uint32_t& hg, uint32_t& hb) {
            const size_t ro = (size_t)y * W;
            const bool m0 = !mimg || __ldg(mimg + ro + x0) > 127, m1 = !mimg || __ldg(mimg + ro + x1) > 127;
            const uint8_t* p0 = simg + (ro + x0) * 3;
            const uint8_t* p1 = simg + (ro + x1) * 3;
            const uint32_t r0v = m0 ? __ldg(p0) : 255u, g0v = m0 ? __ldg(p0 + 1) : 255u, b0v = m0 ? __ldg(p0 + 2) : 255u;
            const uint32_t r1v = m1 ? __ldg(p1) : 255u, g1v = m1 ? __ldg(p1 + 1) : 255u, b1v = m1 ? __ldg(p1 + 2) : 255u;
            hr = (r0v * xa + r1v * xb) >> 4;
            hg = (g0v * xa + g1v * xb) >> 4;
            hb = (b0v * xa + b1v * xb) >> 4;
        };
        int prev_s = -4;
        uint32_t h0r = 0, h0g = 0, h0b = 0, h1r = 0, h1g = 0, h1b = 0;
        for (int r = 0; r < rows; ++r, o += RW * 3) {
            const Tap ty = s_yt[r];
            if (ty.s < 0) {   // top / bottom band
                __stcs(o + 0, (uint8_t)0);
                __stcs(o + 1, (uint8_t)0);
                __stcs(o + 2, (uint8_t)0);
                continue;
            }
            if (ty.s != prev_s) {
                if (ty.s == prev_s + 1) {
                    h0r = h1r, h0g = h1g, h0b = h1b;
                } else {
                    hrow(by + ty.s, h0r, h0g, h0b);
                }
                hrow(by + min(ty.s + 1, bh - 1), h1r, h1g, h1b);
                prev_s = ty.s;
            }
            // VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>: ((a * (h0 >> 4)) >> 16) + ((b * (h1 >> 4)) >> 16) + 2 >> 2
            // with (a * h >> 16) taken as the high word of (a << 16) * h; a + b <= 2049 keeps the result in 0..255
            const uint32_t ya = (uint32_t)(uint16_t)ty.a << 16, yb = (uint32_t)(uint16_t)ty.b << 16;
            __stcs(o + 0, (uint8_t)((__umulhi(ya, h0r) + __umulhi(yb, h1r) + 2u) >> 2));
            __stcs(o + 1, (uint8_t)((__umulhi(ya, h0g) + __umulhi(yb, h1g) + 2u) >> 2));
            __stcs(o + 2, (uint8_t)((__umulhi(ya, h0b) + __umulhi(yb, h1b) + 2u) >> 2));
        }
    }
}

// ---- shrink path: cv2.resize(INTER_AREA) of a bounding box LARGER than the canvas (roi.py:35-38 with w > W or h > H;
// images bigger than roi_size).  OpenCV resize.cpp: integer factors in both directions -> ResizeAreaFast (integer block
// sums; 2x2 rounds as (s + 2) >> 2, other factors as cvRound(s * float(1 / area))); every other ratio -> ResizeArea_
// with computeResizeAreaTab's float32 weights and float32 accumulation in table order, cvRound at the end.  One thread
// per canvas pixel; arithmetic order and roundings are the library's (no FMA contraction), so the result is bit-exact.
struct AreaSpan {
    int s_first, s_full0, s_full1, s_last;   // partial first sample (or -1), full samples [s_full0, s_full1), partial last (or -1)
    float a_first, a_full, a_last;
};
__device__ __forceinline__ AreaSpan area_span(int d, int ssize, int dsize) {
    const double scale = __ddiv_rn((double)ssize, (double)dsize);
    const double f1 = __dmul_rn((double)d, scale), f2 = __dadd_rn(f1, scale);
    const double cell = fmin(scale, __dadd_rn((double)ssize, -f1));
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    AreaSpan a;
    a.s_first = -1; a.s_last = -1; a.a_first = 0.f; a.a_last = 0.f;
    if (__dadd_rn((double)s1, -f1) > 1e-3) {
        a.s_first = s1 - 1;
        a.a_first = __double2float_rn(__ddiv_rn(__dadd_rn((double)s1, -f1), cell));
    }
    a.s_full0 = s1; a.s_full1 = s2;
    a.a_full = __double2float_rn(__ddiv_rn(1.0, cell));
    if (__dadd_rn(f2, -(double)s2) > 1e-3) {
        a.s_last = s2;
        a.a_last = __double2float_rn(__ddiv_rn(fmin(fmin(__dadd_rn(f2, -(double)s2), 1.0), cell), cell));
    }
    return a;
}

__global__ void __launch_bounds__(THREADS) k_roi_down(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                      const int32_t* __restrict__ info, uint8_t* __restrict__ dst, int H, int W,
                                                      int RH, int RW) {
    const int img = blockIdx.y;
    const int32_t* inf = info + (size_t)img * 8;
    const int found = inf[0], bx = inf[1], by = inf[2], bw = inf[3], bh = inf[4];
    if (!(found && bw > 0 && bh > 0)) return;
    const double sc = fmin(__ddiv_rn((double)RW, (double)max(bw, 1)), __ddiv_rn((double)RH, (double)max(bh, 1)));
    const int nw = max((int)__dmul_rn((double)bw, sc), 1), nh = max((int)__dmul_rn((double)bh, sc), 1);
    if (!(nw < bw || nh < bh)) return;   // k_roi's image
    const int oy = (RH - nh) / 2, ox = (RW - nw) / 2;
    const int cc = blockIdx.x * 32 + (threadIdx.x & 31), cr = blockIdx.z * 8 + (threadIdx.x >> 5);
    if (cc >= RW || cr >= RH) return;
    uint8_t* o = dst + (((size_t)img * RH + cr) * RW + cc) * 3;
    const int dx = cc - ox, dy = cr - oy;
    if ((unsigned)dx >= (unsigned)nw || (unsigned)dy >= (unsigned)nh) {
        o[0] = 0; o[1] = 0; o[2] = 0;
        return;
    }
    const uint8_t* simg = src + (size_t)img * H * W * 3;
    const uint8_t* mimg = mask ? mask + (size_t)img * H * W : nullptr;
    auto px = [&](int sy, int sx, float& r, float& g, float& b) {   // pixel of the white-masked image (Transformation.py:451)
        const size_t off = (size_t)(by + sy) * W + (bx + sx);
        const bool m = !mimg || __ldg(mimg + off) > 127;
        const uint8_t* p = simg + off * 3;
        r = m ? (float)__ldg(p) : 255.f;
        g = m ? (float)__ldg(p + 1) : 255.f;
        b = m ? (float)__ldg(p + 2) : 255.f;
    };
    const double fx = __ddiv_rn((double)bw, (double)nw), fy = __ddiv_rn((double)bh, (double)nh);
    const int ix = (int)(fx + 0.5), iy = (int)(fy + 0.5);
    if (fabs(fx - ix) < 2.220446049250313e-16 && fabs(fy - iy) < 2.220446049250313e-16) {
        int sr = 0, sg = 0, sb = 0;
        for (int y = 0; y < iy; ++y)
            for (int x = 0; x < ix; ++x) {
                float r, g, b;
                px(dy * iy + y, dx * ix + x, r, g, b);
                sr += (int)r; sg += (int)g; sb += (int)b;
            }
        if (ix == 2 && iy == 2) {
            o[0] = (uint8_t)((sr + 2) >> 2); o[1] = (uint8_t)((sg + 2) >> 2); o[2] = (uint8_t)((sb + 2) >> 2);
        } else {
            const float inv = __fdiv_rn(1.f, (float)(ix * iy));
            o[0] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn((float)sr, inv))));
            o[1] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn((float)sg, inv))));
            o[2] = (uint8_t)min(255, max(0, __float2int_rn(__fmul_rn((float)sb, inv))));
        }
        return;
    }
    const AreaSpan X = area_span(dx, bw, nw), Y = area_span(dy, bh, nh);
    float ar = 0.f, ag = 0.f, ab = 0.f;
    bool first = true;
    auto row = [&](int sy, float beta) {
        float br = 0.f, bg = 0.f, bb = 0.f, r, g, b;
        if (X.s_first >= 0) {
            px(sy, X.s_first, r, g, b);
            br = __fadd_rn(br, __fmul_rn(r, X.a_first)); bg = __fadd_rn(bg, __fmul_rn(g, X.a_first)); bb = __fadd_rn(bb, __fmul_rn(b, X.a_first));
        }
        for (int sx = X.s_full0; sx < X.s_full1; ++sx) {
            px(sy, sx, r, g, b);
            br = __fadd_rn(br, __fmul_rn(r, X.a_full)); bg = __fadd_rn(bg, __fmul_rn(g, X.a_full)); bb = __fadd_rn(bb, __fmul_rn(b, X.a_full));
        }
        if (X.s_last >= 0) {
            px(sy, X.s_last, r, g, b);
            br = __fadd_rn(br, __fmul_rn(r, X.a_last)); bg = __fadd_rn(bg, __fmul_rn(g, X.a_last)); bb = __fadd_rn(bb, __fmul_rn(b, X.a_last));
        }
        if (first) {
            ar = __fmul_rn(beta, br); ag = __fmul_rn(beta, bg); ab = __fmul_rn(beta, bb);
            first = false;
        } else {
            ar = __fadd_rn(ar, __fmul_rn(beta, br)); ag = __fadd_rn(ag, __fmul_rn(beta, bg)); ab = __fadd_rn(ab, __fmul_rn(beta, bb));
        }
    };
    if (Y.s_first >= 0) row(Y.s_first, Y.a_first);
    for (int sy = Y.s_full0; sy < Y.s_full1; ++sy) row(sy, Y.a_full);
    if (Y.s_last >= 0) row(Y.s_last, Y.a_last);
    o[0] = (uint8_t)min(255, max(0, __float2int_rn(ar)));
    o[1] = (uint8_t)min(255, max(0, __float2int_rn(ag)));
    o[2] = (uint8_t)min(255, max(0, __float2int_rn(ab)));
}

}  // namespace

extern "C" int lfx_roi_letterbox(const uint8_t* src, const uint8_t* mask, const int32_t* info, uint8_t* dst, int B, int H,
                                 int W, int RH, int RW, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && info && dst && B >= 0 && H > 0 && W > 0 && RH > 0 && RW > 0 && B <= 65535, LFX_ERR_ARG,
                "roi_letterbox: bad arguments");
    if (B == 0) return LFX_OK;
    dim3 grid(lfx_div_up(RH, ROI_ROWS), B);
    k_roi<<<grid, THREADS, 0, (cudaStream_t)stream>>>(src, mask, info, dst, H, W, RH, RW);
    int rc = lfx_check_launch("roi_letterbox");
    if (rc || (RW >= W && RH >= H)) return rc;
    // the image is larger than the canvas: bounding boxes that have to shrink are written by the INTER_AREA shrink kernel
    LFX_REQUIRE(lfx_div_up(RH, 8) <= 65535, LFX_ERR_UNSUPPORTED, "roi_letterbox: canvas too tall");
    dim3 grid2(lfx_div_up(RW, 32), B, lfx_div_up(RH, 8));
    k_roi_down<<<grid2, THREADS, 0, (cudaStream_t)stream>>>(src, mask, info, dst, H, W, RH, RW);
    return lfx_check_launch("roi_letterbox(shrink)");
}
