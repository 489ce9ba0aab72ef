// Image-dependent terms of _score_mask (srcs/transform/filters/mask.py:143-188) for K post-processed candidate
// masks of every image at once -- what `mask_strategy: auto` (mask.py:435-461) ranks its candidates by:
//   boundary strength : mean over (dilate3x3 ^ erode3x3)(mask) of the min-max normalised Sobel magnitude (:160-170)
//   green fraction    : |green & mask| / |mask| with green = H in green_hue_range & S >= 40            (:172-177)
// The kernel returns exact integer counts, the float32 magnitudes' sum over the boundary (accumulated in fp64) and
// the image-wide min / max of the magnitude; the host finishes the arithmetic (leaffliction_b200/transform.py).
// cv2.Sobel(CV_32F, ksize 3) on the 8-bit grey image: BORDER_REFLECT_101, integer-valued; cv2.magnitude = float32 sqrt
// of an exactly representable sum (<= 2 * 1020^2).  HBM-bound pixel pass; one read of the image serves all K masks.
#include "lfx_common.cuh"

namespace {

constexpr int SC_KMAX = 8;

__device__ __forceinline__ int refl101s(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return n == 1 ? 0 : i;
}

__global__ void __launch_bounds__(256) k_score_features(const uint8_t* __restrict__ src, const uint8_t* __restrict__ masks,
                                                        double* __restrict__ feat, uint32_t* __restrict__ minmax, int B, int H, int W,
                                                        int K, int green_lo, int green_hi, const LfxTables* __restrict__ tab) {
    __shared__ double s_sum[SC_KMAX];
    __shared__ unsigned long long s_cnt[SC_KMAX][3];
    __shared__ uint32_t s_mm[2];
    const int img = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (threadIdx.x < SC_KMAX) {
        s_sum[threadIdx.x] = 0.0;
        s_cnt[threadIdx.x][0] = s_cnt[threadIdx.x][1] = s_cnt[threadIdx.x][2] = 0ull;
    }
    if (threadIdx.x == 0) {
        s_mm[0] = 0u;   // max of the magnitude bits (non-negative floats order like their bit patterns)
        s_mm[1] = 0u;   // max of ~bits = ~min
    }
    __syncthreads();
    const bool inside = x < W && y < H;
    const uint8_t* simg = src + (size_t)img * H * W * 3;
    float mag = 0.f;
    bool green = false;
    if (inside) {
        int g[3][3];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const uint8_t* p = simg + ((size_t)refl101s(y + dy, H) * W + refl101s(x + dx, W)) * 3;
                g[dy + 1][dx + 1] = rgb2gray(__ldg(p), __ldg(p + 1), __ldg(p + 2));
            }
        const int gx = (g[0][2] + 2 * g[1][2] + g[2][2]) - (g[0][0] + 2 * g[1][0] + g[2][0]);
        const int gy = (g[2][0] + 2 * g[2][1] + g[2][2]) - (g[0][0] + 2 * g[0][1] + g[0][2]);
        mag = __fsqrt_rn((float)(gx * gx + gy * gy));
        const uint8_t* p = simg + ((size_t)y * W + x) * 3;
        const int r = __ldg(p), gg = __ldg(p + 1), b = __ldg(p + 2);
        const int v = max(r, max(gg, b)), d = v - min(r, min(gg, b));
        const int s = (d * tab->sdiv[v] + 2048) >> 12;
        int hh = (v == r) ? (gg - b) : (v == gg) ? (b - r + 2 * d) : (r - gg + 4 * d);
        hh = (hh * tab->hdiv[d] + 2048) >> 12;
        const int h = hh < 0 ? hh + 180 : hh;
        green = (h >= green_lo) && (h <= green_hi) && (s >= 40);
        const uint32_t bits = __float_as_uint(mag);
        atomicMax(&s_mm[0], bits);
        atomicMax(&s_mm[1], ~bits);
    }
    for (int k = 0; k < K; ++k) {
        const uint8_t* m = masks + ((size_t)k * B + img) * H * W;
        bool on = false, bnd = false;
        if (inside) {
            // 3x3 MORPH_ELLIPSE = cross; pixels outside the image are ignored by both dilate and erode
            const bool c = __ldg(m + (size_t)y * W + x) != 0;
            bool any = c, all = c;
            if (y > 0) { const bool t = __ldg(m + (size_t)(y - 1) * W + x) != 0; any |= t; all &= t; }
            if (y < H - 1) { const bool t = __ldg(m + (size_t)(y + 1) * W + x) != 0; any |= t; all &= t; }
            if (x > 0) { const bool t = __ldg(m + (size_t)y * W + x - 1) != 0; any |= t; all &= t; }
            if (x < W - 1) { const bool t = __ldg(m + (size_t)y * W + x + 1) != 0; any |= t; all &= t; }
            on = c;
            bnd = any != all;
        }
        const unsigned nb = __popc(__ballot_sync(0xffffffffu, bnd));
        const unsigned no = __popc(__ballot_sync(0xffffffffu, on));
        const unsigned ng = __popc(__ballot_sync(0xffffffffu, on && green));
        double sm = bnd ? (double)mag : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        if ((threadIdx.x & 31) == 0) {
            if (nb) {
                atomicAdd(&s_sum[k], sm);
                atomicAdd(&s_cnt[k][0], (unsigned long long)nb);
            }
            if (no) atomicAdd(&s_cnt[k][1], (unsigned long long)no);
            if (ng) atomicAdd(&s_cnt[k][2], (unsigned long long)ng);
        }
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double* f = feat + ((size_t)threadIdx.x * B + img) * 4;
        if (s_cnt[threadIdx.x][0]) {
            atomicAdd(&f[0], s_sum[threadIdx.x]);
            atomicAdd(&f[1], (double)s_cnt[threadIdx.x][0]);   // counts < 2^53: exact in fp64
        }
        if (s_cnt[threadIdx.x][1]) atomicAdd(&f[2], (double)s_cnt[threadIdx.x][1]);
        if (s_cnt[threadIdx.x][2]) atomicAdd(&f[3], (double)s_cnt[threadIdx.x][2]);
    }
    if (threadIdx.x == 0) {
        atomicMax(&minmax[(size_t)img * 2], s_mm[0]);
        atomicMax(&minmax[(size_t)img * 2 + 1], s_mm[1]);
    }
}

}  // namespace

// feat[K][B][4] (double) = {sum of |grad| over the mask boundary, boundary pixels, mask pixels, green & mask pixels};
// minmax[B][2] = {float bits of max |grad|, ~(float bits of min |grad|)} over the image.  Both buffers are zero-initialised
// here (stream-ordered memsets).
extern "C" int lfx_score_features(const uint8_t* src, const uint8_t* masks, double* feat, uint32_t* minmax, int B, int H, int W, int K,
                                  int green_lo, int green_hi, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0 || K == 0) return LFX_OK;
    LFX_REQUIRE(src && masks && feat && minmax && B > 0 && H > 0 && W > 0 && K > 0 && K <= SC_KMAX && B <= 65535, LFX_ERR_ARG,
                "score_features: bad arguments (K <= %d)", SC_KMAX);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(feat, 0, (size_t)K * B * 4 * sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(minmax, 0, (size_t)B * 2 * sizeof(uint32_t), st);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "score_features memset: %s", cudaGetErrorString(e));
    dim3 grid(lfx_div_up(W, 32), lfx_div_up(H, 8), B);
    k_score_features<<<grid, 256, 0, st>>>(src, masks, feat, minmax, B, H, W, K, green_lo, green_hi, lfx_tables());
    return lfx_check_launch("score_features");
}
