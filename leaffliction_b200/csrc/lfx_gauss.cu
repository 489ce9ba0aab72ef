// cv2.GaussianBlur for uint8 (bit-exact fixed point): separable, 8.8 horizontal pass into a
// shared uint16 tile, 16.16 vertical pass, round-half-up.  BORDER_REFLECT_101.
// Shared-memory halo tiles; both passes fused so the intermediate never touches HBM.
#include "lfx_common.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int G_TH = 32;     // output rows per tile
constexpr int G_TWB = 384;   // output BYTES per tile row (128 RGB pixels / 384 gray pixels)
constexpr int G_MAXK = 15;

struct GaussTaps {
    int k[G_MAXK];
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// grid (col tiles, row tiles, B)
__global__ void __launch_bounds__(THREADS) k_gauss(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                   int C, int ksize, GaussTaps taps) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int r = ksize >> 1;
    const int rowbytes = W * C;
    const int j0 = blockIdx.x * G_TWB;
    const int y0 = blockIdx.y * G_TH;
    const int tw = min(G_TWB, rowbytes - j0);
    const int th = min(G_TH, H - y0);
    const int inw = G_TWB + 2 * r * C;  // padded tile row pitch (bytes)
    const int inh = th + 2 * r;
    uint8_t* s_in = sm;
    uint16_t* s_mid = reinterpret_cast<uint16_t*>(sm + ((G_TH + 2 * G_MAXK) * (G_TWB + 2 * G_MAXK * 3) + 15 & ~15));
    const uint8_t* simg = src + (size_t)blockIdx.z * H * rowbytes;
    uint8_t* dimg = dst + (size_t)blockIdx.z * H * rowbytes;

    // load with reflection at pixel granularity
    const int lw = tw + 2 * r * C;
    for (int i = threadIdx.x; i < inh * lw; i += THREADS) {
        const int ty = i / lw, tj = i - ty * lw;
        const int gy = reflect101(y0 + ty - r, H);
        const int gj = j0 + tj - r * C;  // byte index in the (virtually padded) row
        int px = (gj >= 0) ? gj / C : -((-gj + C - 1) / C);
        const int c = gj - px * C;
        px = reflect101(px, W);
        s_in[ty * inw + tj] = __ldg(simg + (size_t)gy * rowbytes + px * C + c);
    }
    __syncthreads();
    // horizontal pass: 8.8
    for (int i = threadIdx.x; i < inh * tw; i += THREADS) {
        const int ty = i / tw, tj = i - ty * tw;
        const uint8_t* p = s_in + ty * inw + tj;
        int acc = 0;
        for (int t = 0; t < ksize; ++t) acc += p[t * C] * taps.k[t];
        s_mid[ty * G_TWB + tj] = (uint16_t)acc;
    }
    __syncthreads();
    // vertical pass: 16.16, round
    for (int i = threadIdx.x; i < th * tw; i += THREADS) {
        const int ty = i / tw, tj = i - ty * tw;
        const uint16_t* p = s_mid + ty * G_TWB + tj;
        uint32_t acc = 0;
        for (int t = 0; t < ksize; ++t) acc += (uint32_t)p[t * G_TWB] * (uint32_t)taps.k[t];
        dimg[(size_t)(y0 + ty) * rowbytes + j0 + tj] = (uint8_t)((acc + 32768u) >> 16);
    }
}

constexpr size_t G_SMEM = (((G_TH + 2 * G_MAXK) * (G_TWB + 2 * G_MAXK * 3) + 15) & ~15) + (size_t)(G_TH + 2 * G_MAXK) * G_TWB * 2;

}  // namespace

extern "C" int lfx_gauss_u8(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int ksize, double sigma,
                            lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && B >= 0 && H > 0 && W > 0 && (C == 1 || C == 3), LFX_ERR_ARG, "gauss_u8: bad arguments");
    LFX_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= G_MAXK, LFX_ERR_UNSUPPORTED, "gauss_u8: ksize %d (odd <= %d)", ksize,
                G_MAXK);
    LFX_REQUIRE(B <= 65535 && lfx_div_up(H, G_TH) <= 65535, LFX_ERR_UNSUPPORTED, "gauss_u8: grid too large");
    if (B == 0) return LFX_OK;
    GaussTaps taps;
    int32_t k[31];
    const int rc = lfx_gauss_taps(ksize, sigma, k);
    if (rc != LFX_OK) return rc;
    for (int i = 0; i < G_MAXK; ++i) taps.k[i] = i < ksize ? k[i] : 0;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_gauss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
        attr = true;
    }
    dim3 grid(lfx_div_up((long long)W * C, G_TWB), lfx_div_up(H, G_TH), B);
    k_gauss<<<grid, THREADS, G_SMEM, (cudaStream_t)stream>>>(src, dst, H, W, C, ksize, taps);
    return lfx_check_launch("gauss_u8");
}
