// apply_roi_filter canvas (srcs/transform/filters/roi.py:20-46): crop the contour's bounding box
// from the white-masked image, letterbox with cv2.resize(INTER_AREA) (an upscale here: 2-tap
// fixed-point bilinear with area-mode offsets, OpenCV resize.cpp) into a zero canvas.
#include "lfx_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int ROI_ROWS = 16;  // canvas rows per block

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

// grid (ceil(RH/ROI_ROWS), B)
__global__ void __launch_bounds__(THREADS) k_roi(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                 const int32_t* __restrict__ info, uint8_t* __restrict__ dst, int H, int W,
                                                 int RH, int RW) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* s_out = sm;                                              // [ROI_ROWS][RW*3]
    Tap* s_xt = reinterpret_cast<Tap*>(sm + ((ROI_ROWS * RW * 3 + 15) & ~15));  // [RW]
    Tap* s_yt = s_xt + RW;                                             // [ROI_ROWS]
    const int img = blockIdx.y;
    const int r0 = blockIdx.x * ROI_ROWS;
    const int rows = min(ROI_ROWS, RH - r0);
    const int32_t* inf = info + (size_t)img * 8;
    const int found = inf[0], bx = inf[1], by = inf[2], bw = inf[3], bh = inf[4];
    uint8_t* dimg = dst + ((size_t)img * RH + r0) * RW * 3;
    for (int i = threadIdx.x; i < rows * RW * 3; i += THREADS) s_out[i] = 0;
    if (!found || bw <= 0 || bh <= 0) {
        __syncthreads();
        block_store_bytes(dimg, s_out, rows * RW * 3);
        return;
    }
    // scale = min(W / max(w,1), H / max(h,1)); nw = max(int(w*scale),1)   (roi.py:35-36, Python floats)
    const double sc = fmin(__ddiv_rn((double)RW, (double)max(bw, 1)), __ddiv_rn((double)RH, (double)max(bh, 1)));
    const int nw = max((int)__dmul_rn((double)bw, sc), 1), nh = max((int)__dmul_rn((double)bh, sc), 1);
    const int oy = (RH - nh) / 2, ox = (RW - nw) / 2;
    const bool same = (nw == bw && nh == bh);
    for (int i = threadIdx.x; i < nw; i += THREADS) {
        Tap t;
        if (same) {
            t.s = i; t.a = 2048; t.b = 0;
        } else {
            t = area_tap(i, bw, nw);
        }
        s_xt[i] = t;
    }
    for (int i = threadIdx.x; i < rows; i += THREADS) {
        const int d = r0 + i - oy;
        Tap t;
        t.s = -1; t.a = 0; t.b = 0;
        if (d >= 0 && d < nh) {
            if (same) {
                t.s = d; t.a = 2048; t.b = 0;
            } else {
                t = area_tap(d, bh, nh);
            }
        }
        s_yt[i] = t;
    }
    __syncthreads();
    const uint8_t* simg = src + (size_t)img * H * W * 3;
    const uint8_t* mimg = mask ? mask + (size_t)img * H * W : nullptr;
    for (int i = threadIdx.x; i < rows * nw; i += THREADS) {
        const int ry = i / nw, cx = i - ry * nw;
        const Tap ty = s_yt[ry];
        if (ty.s < 0) continue;
        const Tap tx = s_xt[cx];
        const int y0 = by + ty.s, y1 = by + min(ty.s + 1, bh - 1);
        const int x0 = bx + tx.s, x1 = bx + min(tx.s + 1, bw - 1);
        // masked_rgb = apply_mask(rgb, mask, "white")  (Transformation.py:451)
        const bool m00 = !mimg || __ldg(mimg + (size_t)y0 * W + x0) > 127;
        const bool m01 = !mimg || __ldg(mimg + (size_t)y0 * W + x1) > 127;
        const bool m10 = !mimg || __ldg(mimg + (size_t)y1 * W + x0) > 127;
        const bool m11 = !mimg || __ldg(mimg + (size_t)y1 * W + x1) > 127;
        const uint8_t* p00 = simg + ((size_t)y0 * W + x0) * 3;
        const uint8_t* p01 = simg + ((size_t)y0 * W + x1) * 3;
        const uint8_t* p10 = simg + ((size_t)y1 * W + x0) * 3;
        const uint8_t* p11 = simg + ((size_t)y1 * W + x1) * 3;
        uint8_t* o = s_out + (ry * RW + ox + cx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int v00 = m00 ? __ldg(p00 + c) : 255, v01 = m01 ? __ldg(p01 + c) : 255;
            const int v10 = m10 ? __ldg(p10 + c) : 255, v11 = m11 ? __ldg(p11 + c) : 255;
            int res;
            if (same) {
                res = v00;
            } else {
                const int h0 = v00 * tx.a + v01 * tx.b;  // HResizeLinear, x2048
                const int h1 = v10 * tx.a + v11 * tx.b;
                // VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>
                res = ((((int)ty.a * (h0 >> 4)) >> 16) + (((int)ty.b * (h1 >> 4)) >> 16) + 2) >> 2;
                res = min(255, max(0, res));
            }
            o[c] = (uint8_t)res;
        }
    }
    __syncthreads();
    block_store_bytes(dimg, s_out, rows * RW * 3);
}

}  // namespace

extern "C" int lfx_roi_letterbox(const uint8_t* src, const uint8_t* mask, const int32_t* info, uint8_t* dst, int B, int H,
                                 int W, int RH, int RW, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && info && dst && B >= 0 && H > 0 && W > 0 && RH > 0 && RW > 0 && B <= 65535, LFX_ERR_ARG,
                "roi_letterbox: bad arguments");
    LFX_REQUIRE(RW >= W && RH >= H, LFX_ERR_UNSUPPORTED,
                "roi_letterbox: roi_size (%d,%d) smaller than the image (%d,%d) needs the INTER_AREA shrink path", RH, RW, H, W);
    const size_t smem = ((size_t)(ROI_ROWS * RW * 3 + 15) & ~15) + (size_t)(RW + ROI_ROWS) * 8;
    LFX_REQUIRE(smem <= 200 * 1024, LFX_ERR_UNSUPPORTED, "roi_letterbox: roi width %d too large", RW);
    if (B == 0) return LFX_OK;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(k_roi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    dim3 grid(lfx_div_up(RH, ROI_ROWS), B);
    k_roi<<<grid, THREADS, smem, (cudaStream_t)stream>>>(src, mask, info, dst, H, W, RH, RW);
    return lfx_check_launch("roi_letterbox");
}
