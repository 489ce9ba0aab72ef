// cv2.resize as the mask path uses it: INTER_CUBIC upscale of the working image (_prepare_working_image,
// srcs/transform/filters/mask.py:29-50) and INTER_NEAREST of the mask back to the original size
// (_resize_results_to_original, :526-545).  OpenCV resize.cpp, 8-bit path: a = -0.75 cubic weights in float32 quantised
// to 11 bits (cvRound), int32 horizontal pass, vertical pass (sum + 2^21) >> 22 with saturation, border taps replicated;
// nearest: sx = min(floor(dx * sw / dw), sw - 1).  The cubic path is the +-1 LSB class (OpenCV's SIMD and scalar code differ
// from each other by 1 LSB on a few per cent of the values); nearest is exact.
#include <math.h>

#include "lfx_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int RC_ROWS = 8;  // output rows per block

// grid (ceil(OH / RC_ROWS), B)
__global__ void __launch_bounds__(THREADS) k_resize_cubic(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                          int OH, int OW, const int32_t* __restrict__ xfirst,
                                                          const int32_t* __restrict__ xw, const int32_t* __restrict__ yfirst,
                                                          const int32_t* __restrict__ yw) {
    const int img = blockIdx.y;
    const int r0 = blockIdx.x * RC_ROWS, rows = min(RC_ROWS, OH - r0);
    const uint8_t* simg = src + (size_t)img * H * W * 3;
    uint8_t* dimg = dst + ((size_t)img * OH + r0) * OW * 3;
    for (int i = threadIdx.x; i < rows * OW; i += THREADS) {
        const int ry = i / OW, ox = i - ry * OW;
        const int oy = r0 + ry;
        const int xs = __ldg(xfirst + ox), ys = __ldg(yfirst + oy);
        const int4 wx = __ldg(reinterpret_cast<const int4*>(xw) + ox), wy = __ldg(reinterpret_cast<const int4*>(yw) + oy);
        const int wxa[4] = {wx.x, wx.y, wx.z, wx.w}, wya[4] = {wy.x, wy.y, wy.z, wy.w};
        int xo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xo[j] = min(max(xs + j, 0), W - 1) * 3;
        int acc[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* row = simg + (size_t)min(max(ys + k, 0), H - 1) * W * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int h = 0;   // HResizeCubic: int32 accumulation of 11-bit weights
#pragma unroll
                for (int j = 0; j < 4; ++j) h += (int)__ldg(row + xo[j] + c) * wxa[j];
                acc[c] += h * wya[k];   // VResizeCubic
            }
        }
        uint8_t* o = dimg + (size_t)i * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(255, max(0, (acc[c] + (1 << 21)) >> 22));   // FixedPtCast<int, uchar, 22>
    }
}

__global__ void __launch_bounds__(THREADS) k_resize_nearest(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                            int C, int OH, int OW) {
    const int img = blockIdx.y;
    const double fx = __ddiv_rn((double)W, (double)OW), fy = __ddiv_rn((double)H, (double)OH);
    const uint8_t* simg = src + (size_t)img * H * W * C;
    uint8_t* dimg = dst + (size_t)img * OH * OW * C;
    for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < (long long)OH * OW; i += (long long)gridDim.x * THREADS) {
        const int oy = (int)(i / OW), ox = (int)(i - (long long)oy * OW);
        const int sx = min((int)floor(__dmul_rn((double)ox, fx)), W - 1), sy = min((int)floor(__dmul_rn((double)oy, fy)), H - 1);
        const uint8_t* s = simg + ((size_t)sy * W + sx) * C;
        uint8_t* o = dimg + (size_t)i * C;
        for (int c = 0; c < C; ++c) o[c] = __ldg(s + c);
    }
}

}  // namespace

// Host: first source index (tap k reads clip(first + k)) and the four 11-bit weights of every destination index.
extern "C" int lfx_cubic_table(int in_size, int out_size, int32_t* first, int32_t* weights) {
    LFX_REQUIRE(in_size > 0 && out_size > 0 && first && weights, LFX_ERR_ARG, "cubic_table: bad arguments");
    const double scale = (double)in_size / (double)out_size;
    const float A = -0.75f;
    for (int d = 0; d < out_size; ++d) {
        float f = (float)(((double)d + 0.5) * scale - 0.5);
        const int s = (int)floorf(f);
        const float x = f - (float)s;
        float c[4];
        c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
        c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
        const float xm = 1.f - x;
        c[2] = ((A + 2.f) * xm - (A + 3.f)) * xm * xm + 1.f;
        c[3] = 1.f - c[0] - c[1] - c[2];
        first[d] = s - 1;
        for (int k = 0; k < 4; ++k) weights[d * 4 + k] = (int32_t)lrintf(c[k] * 2048.f);   // saturate_cast<short>: round half to even
    }
    return LFX_OK;
}

extern "C" int lfx_resize_cubic(const uint8_t* src, uint8_t* dst, int B, int H, int W, int OH, int OW, const int32_t* xfirst,
                                const int32_t* xweights, const int32_t* yfirst, const int32_t* yweights, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && xfirst && xweights && yfirst && yweights && B > 0 && B <= 65535 && H > 0 && W > 0 && OH > 0 && OW > 0,
                LFX_ERR_ARG, "resize_cubic: bad arguments");
    LFX_REQUIRE(((reinterpret_cast<uintptr_t>(xweights) | reinterpret_cast<uintptr_t>(yweights)) & 15) == 0, LFX_ERR_ARG,
                "resize_cubic: weight tables must be 16-byte aligned");
    dim3 grid(lfx_div_up(OH, RC_ROWS), B);
    k_resize_cubic<<<grid, THREADS, 0, (cudaStream_t)stream>>>(src, dst, H, W, OH, OW, xfirst, xweights, yfirst, yweights);
    return lfx_check_launch("resize_cubic");
}

extern "C" int lfx_resize_nearest(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int OH, int OW, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0 && C <= 4 && OH > 0 && OW > 0, LFX_ERR_ARG,
                "resize_nearest: bad arguments");
    dim3 grid((unsigned)min((long long)LFX_NUM_SMS * 8, ((long long)OH * OW + THREADS - 1) / THREADS), B);
    k_resize_nearest<<<grid, THREADS, 0, (cudaStream_t)stream>>>(src, dst, H, W, C, OH, OW);
    return lfx_check_launch("resize_nearest");
}
