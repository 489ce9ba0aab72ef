// External contour of the selected component: cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
// restated as Suzuki-Abe border following from the component's first raster pixel
// (largest_contour, srcs/cli/Transformation.py:285-292).  Border following is inherently
// sequential per contour, so the batch is the parallel axis: one thread per image, the lanes of a
// warp walk their own borders in lock-step.  Also accumulates the polygon's Green-formula sums
// (cv2.moments m00/m10/m01, analyze.py:43) exactly in int64.
#include "lfx_common.cuh"

namespace {

__device__ __forceinline__ bool on(const uint8_t* m, int H, int W, int x, int y) {
    return x >= 0 && x < W && y >= 0 && y < H && __ldg(m + (size_t)y * W + x) != 0;
}

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
    int32_t* out = points + (size_t)img * max_pts * 2;
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
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
    } while (!on(m, H, W, x0 + dx[s], y0 + dy[s]) && s != s_end0);
    if (!on(m, H, W, x0 + dx[s], y0 + dy[s])) {
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
                if (on(m, H, W, x4, y4)) break;
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
    counts[img] = n <= max_pts ? n : -n;
    if (sm) {
        sm[0] = a00;
        sm[1] = a10;
        sm[2] = a01;
    }
}

}  // namespace

extern "C" int lfx_trace_contour(const uint8_t* mask, const int32_t* info, int32_t* points, int32_t* counts,
                                 int64_t* sums, int B, int H, int W, int max_pts, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(mask && info && points && counts && B > 0 && H > 0 && W > 0 && max_pts > 0, LFX_ERR_ARG,
                "trace_contour: bad arguments");
    k_trace_contour<<<lfx_div_up(B, 64), 64, 0, (cudaStream_t)stream>>>(mask, info, points, counts,
                                                                       reinterpret_cast<long long*>(sums), B, H, W, max_pts);
    return lfx_check_launch("trace_contour");
}
