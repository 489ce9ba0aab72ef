// External contour of the selected component: cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
// restated as Suzuki-Abe border following from the component's first raster pixel
// (largest_contour, srcs/cli/Transformation.py:285-292).  Border following is inherently
// sequential per contour, so the batch is the parallel axis: one thread per image, the lanes of a
// warp walk their own borders in lock-step.  Also accumulates the polygon's Green-formula sums
// (cv2.moments m00/m10/m01, analyze.py:43) exactly in int64.
#include <limits.h>
#include <math.h>

#include "lfx_common.cuh"

namespace {

// The walk itself, for ONE image, by ONE thread; `on(x, y)` tells whether a pixel belongs to the component's mask.
template <class OnFn>
__device__ __forceinline__ void trace_walk(OnFn on, const int32_t* __restrict__ inf, int32_t* __restrict__ out, int32_t* __restrict__ count,
                                           long long* __restrict__ sm, int H, int W, int max_pts) {
    // direction s = 0..7 (E, NE, N, NW, W, SW, S, SE) -> step; two bits per direction (step + 1) in a constant, so that the
    // dependent chain of the walk carries no table look-up: dx = {1,1,0,-1,-1,-1,0,1}, dy = {0,-1,-1,-1,0,1,1,1}
    struct Dir {
        __device__ __forceinline__ int operator[](int s) const { return (int)((code >> (2 * s)) & 3u) - 1; }
        unsigned code;
    };
    const Dir dx = {0x901Au}, dy = {0xA901u};
    const int x0 = inf[7] >> 8, y0 = inf[2];
    int n = 0;
    long long a00 = 0, a10 = 0, a01 = 0;
    int fx = 0, fy = 0, lx = 0, ly = 0;  // first / previous emitted point
    auto emit = [&](int x, int y) {
        if (n < max_pts) {
            out[n * 2] = x;
            out[n * 2 + 1] = y;
        }
        if (n == 0) {
            fx = x;
            fy = y;
        } else {
            const long long d = (long long)lx * y - (long long)x * ly;
            a00 += d;
            a10 += d * (lx + x);
            a01 += d * (ly + y);
        }
        lx = x;
        ly = y;
        ++n;
    };
    int s = 4;
    const int s_end0 = 4;
    do {
        s = (s - 1) & 7;
    } while (!on(x0 + dx[s], y0 + dy[s]) && s != s_end0);
    if (!on(x0 + dx[s], y0 + dy[s])) {
        emit(x0, y0);  // isolated pixel
    } else {
        const int x1 = x0 + dx[s], y1 = y0 + dy[s];
        int x3 = x0, y3 = y0, prev_s = s ^ 4, px = x0, py = y0;
        const long long guard = 4ll * H * W + 16;  // a closed border never needs more steps
        for (long long it = 0; it < guard; ++it) {
            int x4, y4;
            while (true) {
                ++s;
                x4 = x3 + dx[s & 7];
                y4 = y3 + dy[s & 7];
                if (on(x4, y4)) break;
            }
            s &= 7;
            if (s != prev_s) {  // CHAIN_APPROX_SIMPLE: keep direction changes only
                emit(px, py);
                prev_s = s;
            }
            px += dx[s];
            py += dy[s];
            if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
            x3 = x4;
            y3 = y4;
            s = (s + 4) & 7;
        }
    }
    // close the polygon: term (last -> first)
    if (n > 0) {
        const long long d = (long long)lx * fy - (long long)fx * ly;
        a00 += d;
        a10 += d * (lx + fx);
        a01 += d * (ly + fy);
    }
    *count = n <= max_pts ? n : -n;
    if (sm) {
        sm[0] = a00;
        sm[1] = a10;
        sm[2] = a01;
    }
}

// General shapes: one thread per image, the mask read from global memory (the lanes of a warp walk their own borders in
// lock-step).
__global__ void k_trace_contour(const uint8_t* __restrict__ mask, const int32_t* __restrict__ info,
                                int32_t* __restrict__ points, int32_t* __restrict__ counts,
                                long long* __restrict__ sums, int B, int H, int W, int max_pts) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= B) return;
    const int32_t* inf = info + (size_t)img * 8;
    long long* sm = sums ? sums + (size_t)img * 3 : nullptr;
    if (!inf[0]) {
        counts[img] = 0;
        if (sm) sm[0] = sm[1] = sm[2] = 0;
        return;
    }
    const uint8_t* m = mask + (size_t)img * H * W;
    auto on = [&](int x, int y) { return x >= 0 && x < W && y >= 0 && y < H && __ldg(m + (size_t)y * W + x) != 0; };
    trace_walk(on, inf, points + (size_t)img * max_pts * 2, counts + img, sm, H, W, max_pts);
}

// Images whose bit plane fits shared memory (H * ceil(W / 32) words <= 48 KB: up to 512 x 768): one 32-thread block per image
// packs the mask into a bit plane (32 pixels per word), then lane 0 walks the border out of shared memory.  The walk is a chain
// of dependent neighbour tests (2-5 per border pixel); from global memory each one is an L2 round trip and the 32 walks of a
// warp advance at the pace of the slowest, here it is a shared-memory read and every image has its own warp.
__global__ void __launch_bounds__(32) k_trace_contour_bits(const uint8_t* __restrict__ mask, const int32_t* __restrict__ info,
                                                            int32_t* __restrict__ points, int32_t* __restrict__ counts,
                                                            long long* __restrict__ sums, int H, int W, int max_pts) {
    extern __shared__ uint32_t s_bits[];
    const int img = blockIdx.x;
    const int32_t* inf = info + (size_t)img * 8;
    long long* sm = sums ? sums + (size_t)img * 3 : nullptr;
    if (!inf[0]) {
        if (threadIdx.x == 0) {
            counts[img] = 0;
            if (sm) sm[0] = sm[1] = sm[2] = 0;
        }
        return;
    }
    const uint8_t* m = mask + (size_t)img * H * W;
    const int wpr = (W + 31) >> 5;
    const bool vec = (W & 31) == 0 && (reinterpret_cast<uintptr_t>(m) & 15) == 0;
    for (int w = threadIdx.x; w < H * wpr; w += blockDim.x) {
        const int y = w / wpr, xw = w - y * wpr;
        const uint8_t* row = m + (size_t)y * W + xw * 32;
        uint32_t bits = 0;
        if (vec) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + q);
                const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((ws[j] >> (8 * b)) & 255) bits |= 1u << (q * 16 + j * 4 + b);
            }
        } else {
            const int nb = min(32, W - xw * 32);
            for (int b = 0; b < nb; ++b)
                if (row[b]) bits |= 1u << b;
        }
        s_bits[w] = bits;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto on = [&](int x, int y) { return x >= 0 && x < W && y >= 0 && y < H && ((s_bits[y * wpr + (x >> 5)] >> (x & 31)) & 1u) != 0; };
    trace_walk(on, inf, points + (size_t)img * max_pts * 2, counts + img, sm, H, W, max_pts);
}

// ---- numeric record of apply_analyze_filter (analyze.py:43-98) + convex hull, one thread per image.
// The hull of the contour polygon is the hull of the leftmost / rightmost vertex of every row: those are sorted by
// (y, x) by construction, so Andrew's monotone chain needs no sort.  All predicates are exact in int64.
__device__ __forceinline__ long long cross3(int ax, int ay, int bx, int by, int cx, int cy) {
    return (long long)(bx - ax) * (cy - ay) - (long long)(by - ay) * (cx - ax);
}

// SMEM = false: one thread per image, scratch (row extremes, hull sequence and stack) in the caller's global workspace.
// SMEM = true (6H + 8 words fit shared memory, H <= 2000): one 32-thread block per image, the scratch in shared memory; the record
// is still computed by ONE thread in the same order (bit-identical results) -- the hull's stack and the row extremes are
// read-modify-write chains, an L2 round trip per access from global memory, a shared-memory access here.
template <bool SMEM>
__global__ void k_analyze_record(const int32_t* __restrict__ points, const int32_t* __restrict__ counts,
                                 const long long* __restrict__ sums, int32_t* __restrict__ rec_i, double* __restrict__ rec_f,
                                 int32_t* __restrict__ hull, int32_t* __restrict__ ws, int B, int H, int max_pts, int max_hull) {
    extern __shared__ int32_t s_ws[];
    const int img = SMEM ? (int)blockIdx.x : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (img >= B) return;
    const bool lead = !SMEM || threadIdx.x == 0;
    int32_t* ri = rec_i + (size_t)img * 24;
    double* rf = rec_f + (size_t)img * 12;
    if (lead) {
        for (int k = 0; k < 24; ++k) ri[k] = 0;
        for (int k = 0; k < 12; ++k) rf[k] = 0.0;
    }
    const int n = counts[img];
    if (n <= 0 || n > max_pts) {   // no contour, or the point buffer was too small (count < 0)
        if (lead) ri[1] = n;
        return;
    }
    const int32_t* p = points + (size_t)img * max_pts * 2;
    int32_t* minx = SMEM ? s_ws : ws + (size_t)img * (6 * (size_t)H + 8);
    int32_t* maxx = minx + H;
    int32_t* seq = maxx + H;          // [2H] packed x | y << 16
    int32_t* stk = seq + 2 * H;       // [2H + 8]
    if (SMEM) {
        for (int y = threadIdx.x; y < H; y += blockDim.x) {
            minx[y] = INT_MAX;
            maxx[y] = INT_MIN;
        }
        __syncthreads();
        if (threadIdx.x != 0) return;
    } else {
        for (int y = 0; y < H; ++y) {
            minx[y] = INT_MAX;
            maxx[y] = INT_MIN;
        }
    }
    // ---- extreme points (first argmin / argmax, analyze.py:60-64), row extremes, sums for the PCA
    int lx = p[0], ly = p[1], rx = p[0], ry = p[1], tx = p[0], ty = p[1], bx = p[0], by = p[1];
    double sx = 0.0, sy = 0.0;
    for (int i = 0; i < n; ++i) {
        const int x = p[2 * i], y = p[2 * i + 1];
        if (x < lx) { lx = x; ly = y; }
        if (x > rx) { rx = x; ry = y; }
        if (y < ty) { tx = x; ty = y; }
        if (y > by) { bx = x; by = y; }
        if ((unsigned)y < (unsigned)H) {
            minx[y] = min(minx[y], x);
            maxx[y] = max(maxx[y], x);
        }
        sx += (double)x;
        sy += (double)y;
    }
    // ---- centroid: cv2.moments of the polygon (contourMoments: a00 * 0.5, a10 / 6, sign of a00), analyze.py:43-49
    int cx, cy;
    double m00 = 0.0;
    {
        long long s00, s10, s01;
        if (sums) {
            s00 = sums[(size_t)img * 3]; s10 = sums[(size_t)img * 3 + 1]; s01 = sums[(size_t)img * 3 + 2];
        } else {   // Green-formula sums of the closed polygon, exact in int64 (what lfx_trace_contour accumulates)
            s00 = s10 = s01 = 0;
            int px = p[2 * (n - 1)], py = p[2 * (n - 1) + 1];
            for (int i = 0; i < n; ++i) {
                const int x = p[2 * i], y = p[2 * i + 1];
                const long long d = (long long)px * y - (long long)x * py;
                s00 += d;
                s10 += d * (px + x);
                s01 += d * (py + y);
                px = x;
                py = y;
            }
        }
        const double a00 = (double)s00, a10 = (double)s10, a01 = (double)s01;
        if (fabs(a00) > 1.1920928955078125e-07) {
            const double sg = a00 > 0 ? 1.0 : -1.0;
            m00 = __dmul_rn(a00, 0.5 * sg);
            const double m10 = __dmul_rn(a10, 0.16666666666666666 * sg), m01 = __dmul_rn(a01, 0.16666666666666666 * sg);
            cx = (int)__ddiv_rn(m10, m00);
            cy = (int)__ddiv_rn(m01, m00);
        } else {
            cx = (int)__ddiv_rn(sx, (double)n);
            cy = (int)__ddiv_rn(sy, (double)n);
        }
    }
    // ---- convex hull (analyze.py:77, mask.py:158)
    int ns = 0;
    for (int y = 0; y < H; ++y) {
        if (minx[y] == INT_MAX) continue;
        seq[ns++] = minx[y] | (y << 16);
        if (maxx[y] != minx[y]) seq[ns++] = maxx[y] | (y << 16);
    }
    auto X = [](int q) { return q & 0xFFFF; };
    auto Y = [](int q) { return (int)((unsigned)q >> 16); };
    int nh = 0;
    long long area2 = 0;
    int32_t* ho = hull + (size_t)img * max_hull * 2;
    auto out_pt = [&](int q) {
        if (nh < max_hull) {
            ho[2 * nh] = X(q);
            ho[2 * nh + 1] = Y(q);
        }
        ++nh;
    };
    if (ns <= 2) {
        for (int i = 0; i < ns; ++i) out_pt(seq[i]);
    } else {
        for (int pass = 0; pass < 2; ++pass) {   // the chain over the sequence, then over its reverse
            int k = 0;
            for (int i = 0; i < ns; ++i) {
                const int q = seq[pass ? ns - 1 - i : i];
                while (k >= 2 && cross3(X(stk[k - 2]), Y(stk[k - 2]), X(stk[k - 1]), Y(stk[k - 1]), X(q), Y(q)) <= 0) --k;
                stk[k++] = q;
            }
            for (int i = 0; i + 1 < k; ++i) out_pt(stk[i]);
        }
        if (nh <= max_hull)
            for (int i = 0; i < nh; ++i) {
                const int j = i + 1 == nh ? 0 : i + 1;
                area2 += (long long)ho[2 * i] * ho[2 * j + 1] - (long long)ho[2 * j] * ho[2 * i + 1];
            }
    }
    // ---- PCA of the contour vertices (cv2.PCACompute2, analyze.py:88-98): mean, covariance / n, symmetric 2x2 eigen
    const double mx = sx / n, my = sy / n;
    double cxx = 0.0, cxy = 0.0, cyy = 0.0;
    for (int i = 0; i < n; ++i) {
        const double dx = p[2 * i] - mx, dy = p[2 * i + 1] - my;
        cxx += dx * dx;
        cxy += dx * dy;
        cyy += dy * dy;
    }
    cxx /= n; cxy /= n; cyy /= n;
    const double tr = 0.5 * (cxx + cyy), df = 0.5 * (cxx - cyy), rad = sqrt(df * df + cxy * cxy);
    const double l0 = tr + rad, l1 = tr - rad;
    double v0x, v0y;
    if (rad < 1e-300) {   // isotropic: any basis
        v0x = 1.0; v0y = 0.0;
    } else if (df >= 0) {
        v0x = df + rad; v0y = cxy;
    } else {
        v0x = cxy; v0y = rad - df;
    }
    {
        const double nr = sqrt(v0x * v0x + v0y * v0y);
        if (nr > 0) { v0x /= nr; v0y /= nr; } else { v0x = 1.0; v0y = 0.0; }
    }
    const double v1x = -v0y, v1y = v0x;
    int e[8];
    double pmin0 = 1e300, pmax0 = -1e300, pmin1 = 1e300, pmax1 = -1e300;
    for (int i = 0; i < n; ++i) {
        const double x = p[2 * i], y = p[2 * i + 1];
        const double a = x * v0x + y * v0y, b = x * v1x + y * v1y;
        if (a < pmin0) { pmin0 = a; e[0] = p[2 * i]; e[1] = p[2 * i + 1]; }
        if (a > pmax0) { pmax0 = a; e[2] = p[2 * i]; e[3] = p[2 * i + 1]; }
        if (b < pmin1) { pmin1 = b; e[4] = p[2 * i]; e[5] = p[2 * i + 1]; }
        if (b > pmax1) { pmax1 = b; e[6] = p[2 * i]; e[7] = p[2 * i + 1]; }
    }
    ri[0] = 1; ri[1] = n; ri[2] = cx; ri[3] = cy;
    ri[4] = lx; ri[5] = ly; ri[6] = rx; ri[7] = ry; ri[8] = tx; ri[9] = ty; ri[10] = bx; ri[11] = by;
    ri[12] = nh <= max_hull ? nh : -nh;
    for (int k = 0; k < 8; ++k) ri[14 + k] = e[k];
    rf[0] = m00;
    rf[1] = 0.5 * (double)(area2 < 0 ? -area2 : area2);
    rf[2] = mx; rf[3] = my; rf[4] = v0x; rf[5] = v0y; rf[6] = v1x; rf[7] = v1y; rf[8] = l0; rf[9] = l1;
}

}  // namespace

extern "C" size_t lfx_analyze_workspace(int B, int H) {
    if (B <= 0 || H <= 0) return 0;
    return (size_t)B * (6 * (size_t)H + 8) * sizeof(int32_t);
}

extern "C" int lfx_analyze_record(const int32_t* points, const int32_t* counts, const int64_t* sums, int32_t* rec_i32,
                                  double* rec_f64, int32_t* hull_points, int B, int H, int W, int max_pts, int max_hull,
                                  void* workspace, size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(points && counts && rec_i32 && rec_f64 && hull_points && B > 0 && H > 0 && W > 0 && max_pts > 0 && max_hull > 0,
                LFX_ERR_ARG, "analyze_record: bad arguments");
    LFX_REQUIRE(H <= 32767 && W <= 65535, LFX_ERR_UNSUPPORTED, "analyze_record: image too large for the packed hull points");
    LFX_REQUIRE(workspace && workspace_bytes >= lfx_analyze_workspace(B, H), LFX_ERR_WORKSPACE, "analyze_record: workspace %zu < %zu bytes",
                workspace_bytes, lfx_analyze_workspace(B, H));
    const size_t scratch = (6 * (size_t)H + 8) * sizeof(int32_t);
    if (scratch <= 48 * 1024)
        k_analyze_record<true><<<B, 32, scratch, (cudaStream_t)stream>>>(points, counts, reinterpret_cast<const long long*>(sums), rec_i32,
                                                                       rec_f64, hull_points, nullptr, B, H, max_pts, max_hull);
    else
        k_analyze_record<false><<<lfx_div_up(B, 64), 64, 0, (cudaStream_t)stream>>>(points, counts, reinterpret_cast<const long long*>(sums),
                                                                                   rec_i32, rec_f64, hull_points,
                                                                                   reinterpret_cast<int32_t*>(workspace), B, H, max_pts, max_hull);
    return lfx_check_launch("analyze_record");
}

extern "C" int lfx_trace_contour(const uint8_t* mask, const int32_t* info, int32_t* points, int32_t* counts,
                                 int64_t* sums, int B, int H, int W, int max_pts, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(mask && info && points && counts && B > 0 && H > 0 && W > 0 && max_pts > 0, LFX_ERR_ARG,
                "trace_contour: bad arguments");
    const size_t plane = (size_t)H * ((W + 31) >> 5) * sizeof(uint32_t);
    if (plane <= 48 * 1024)
        k_trace_contour_bits<<<B, 32, plane, (cudaStream_t)stream>>>(mask, info, points, counts, reinterpret_cast<long long*>(sums), H, W,
                                                                   max_pts);
    else
        k_trace_contour<<<lfx_div_up(B, 64), 64, 0, (cudaStream_t)stream>>>(mask, info, points, counts,
                                                                           reinterpret_cast<long long*>(sums), B, H, W, max_pts);
    return lfx_check_launch("trace_contour");
}
