// cv2.GaussianBlur for uint8 (bit-exact fixed point): separable, 8.8 horizontal pass into a
// shared uint16 tile, 16.16 vertical pass, round-half-up.  BORDER_REFLECT_101.
// Shared-memory halo tiles; both passes fused so the intermediate never touches HBM.
#include "lfx_common.cuh"

int lfx_gauss_tma_try(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int ksize, const int32_t* taps, cudaStream_t st);

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


// ---------------------------------------------------------------- fast path (K = 5 or 15)
// One thread owns 4 consecutive BYTES of the row (a 32-bit column group) and walks down the rows:
// aligned 32-bit loads of the tap window, horizontal pass as packed 2x16-bit multiply-adds, the last
// K horizontal results kept in registers (no shared memory, no barriers), one 32-bit store per row.
// HBM traffic = read once (+ halo rows from L2) + write once.
template <int C, int K>
__global__ void __launch_bounds__(256) k_gauss_fast(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                    int rows_per_block, GaussTaps taps) {
    constexpr int R = K / 2;
    constexpr int A = ((R * C + 3) / 4) * 4;            // aligned distance from the window start to byte j
    constexpr int NWIN = (A + 4 + R * C + 3) / 4;       // window words
    const int rowbytes = W * C;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j >= rowbytes) return;
    const int y0 = blockIdx.y * rows_per_block;
    const int y_end = min(H, y0 + rows_per_block);      // exclusive
    const uint8_t* simg = src + (size_t)blockIdx.z * H * rowbytes;
    uint8_t* dimg = dst + (size_t)blockIdx.z * H * rowbytes;
    const bool edge = (j - R * C < 0) || (j + 3 + R * C >= rowbytes);
    uint32_t hl[K], hh[K];
#pragma unroll
    for (int t = 0; t < K; ++t) hl[t] = hh[t] = 0;
    const int total = (y_end - y0) + 2 * R;
    for (int base = 0; base < total; base += K) {
#pragma unroll
        for (int sI = 0; sI < K; ++sI) {
            const int step = base + sI;
            if (step < total) {
                const int yy = y0 - R + step;
                const uint8_t* row = simg + (size_t)reflect101(yy, H) * rowbytes;
                uint32_t win[NWIN];
                if (!edge) {
                    const uint32_t* p = reinterpret_cast<const uint32_t*>(row + j - A);
#pragma unroll
                    for (int i = 0; i < NWIN; ++i) win[i] = __ldg(p + i);
                } else {
#pragma unroll
                    for (int i = 0; i < NWIN; ++i) {
                        uint32_t wv = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int g = j - A + i * 4 + b;
                            int px = (g >= 0) ? g / C : -((-g + C - 1) / C);
                            const int c = g - px * C;
                            px = reflect101(px, W);
                            wv |= (uint32_t)__ldg(row + px * C + c) << (8 * b);
                        }
                        win[i] = wv;
                    }
                }
                uint32_t alo = 0, ahi = 0;
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int o = A + (t - R) * C;   // byte offset of this tap's 4 bytes inside the window
                    const int wi = o >> 2, sh = o & 3;
                    const uint32_t v = sh == 0 ? win[wi] : __byte_perm(win[wi], win[wi + 1 < NWIN ? wi + 1 : wi], 0x3210 + 0x1111 * sh);
                    alo += (v & 0x00FF00FFu) * (uint32_t)taps.k[t];
                    ahi += ((v >> 8) & 0x00FF00FFu) * (uint32_t)taps.k[t];
                }
                hl[sI] = alo;
                hh[sI] = ahi;
                if (step >= K - 1) {
                    const int oy = yy - R;
                    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const uint32_t kt = (uint32_t)taps.k[(i - sI - 1 + 2 * K) % K];   // slot i holds window row t
                        s0 += (hl[i] & 0xFFFFu) * kt;
                        s2 += (hl[i] >> 16) * kt;
                        s1 += (hh[i] & 0xFFFFu) * kt;
                        s3 += (hh[i] >> 16) * kt;
                    }
                    const uint32_t o32 = ((s0 + 32768u) >> 16) | (((s1 + 32768u) >> 16) << 8) | (((s2 + 32768u) >> 16) << 16) |
                                         (((s3 + 32768u) >> 16) << 24);
                    *reinterpret_cast<uint32_t*>(dimg + (size_t)oy * rowbytes + j) = o32;
                }
            }
        }
    }
}

template <int C, int K>
void launch_fast(const uint8_t* src, uint8_t* dst, int B, int H, int W, const GaussTaps& taps, cudaStream_t st) {
    const int groups = W * C / 4;
    int block = ((groups + 31) / 32) * 32;
    if (block > 256) block = 256;
    const int rows = K <= 5 ? 32 : 64;
    dim3 grid(lfx_div_up(groups, block), lfx_div_up(H, rows), B);
    k_gauss_fast<C, K><<<grid, block, 0, st>>>(src, dst, H, W, rows, taps);
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
    // TMA-tiled kernel (lfx_gauss_tma.cu) for 5x5 / 15x15 on 16-byte rows; everything else below
    if (lfx_gauss_tma_try(src, dst, B, H, W, C, ksize, k, (cudaStream_t)stream) == 0) return lfx_check_launch("gauss_u8(tma)");
    // fast path: 32-bit aligned rows, the two kernel sizes the reference uses (blur.py:61,72; mask.py:770)
    const bool aligned = ((W * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 4 == 0) &&
                         W > ksize && lfx_div_up(H, 32) <= 65535;
    if (aligned && (ksize == 5 || ksize == 15)) {
        cudaStream_t st = (cudaStream_t)stream;
        if (C == 3 && ksize == 5) launch_fast<3, 5>(src, dst, B, H, W, taps, st);
        else if (C == 1 && ksize == 5) launch_fast<1, 5>(src, dst, B, H, W, taps, st);
        else if (C == 3 && ksize == 15) launch_fast<3, 15>(src, dst, B, H, W, taps, st);
        else launch_fast<1, 15>(src, dst, B, H, W, taps, st);
        return lfx_check_launch("gauss_u8(fast)");
    }
    static bool attr_[LFX_MAX_DEVICES] = {false};
    bool& attr = attr_[lfx_dev()];
    if (!attr) {
        cudaFuncSetAttribute(k_gauss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
        attr = true;
    }
    dim3 grid(lfx_div_up((long long)W * C, G_TWB), lfx_div_up(H, G_TH), B);
    k_gauss<<<grid, THREADS, G_SMEM, (cudaStream_t)stream>>>(src, dst, H, W, C, ksize, taps);
    return lfx_check_launch("gauss_u8");
}
