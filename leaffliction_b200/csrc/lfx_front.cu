// Per-image front ends built on the bit-plane / run-CCL infrastructure (lfx_planes.cuh):
//   lfx_canny            cv2.Canny(gray, lo, hi, aperture 3, L1|L2)          (mask.py:679-680,789; blur.py:30; analyze.py:120)
//   lfx_raw_mask         _create_inclusive_mask / _create_enhanced_mask      (mask.py:727-831, :610-724)
//   lfx_brown_spots      apply_brown_filter numeric core                      (brown.py:21-89)
//   lfx_saliency_blur    apply_blur_filter                                    (blur.py:18-79)
// One thread block per image.  Up to ~256x256 (PlantVillage) the grey image, every mask plane and the run tables stay
// in shared memory, so HBM sees one read of the input and one write of the result; larger images (512x512, 1024x1024)
// keep the planes and the grey image in the block's global scratch (L2-resident) and run the same code on them.
// One block per SM (the whole image lives in shared memory): 1024 threads so that the many short, barrier-separated
// passes over the bit planes have twice the warps to hide their latency.
#define LFX_MT 1024
#include <float.h>

#include "lfx_planes.cuh"

namespace {
// i -> (i / n, i % n) without an integer division when n is a power of two (256-wide images: every n below is)
__device__ __forceinline__ void divmod_fast(int i, int n, int& q, int& r) {
    if ((n & (n - 1)) == 0) {
        q = i >> (31 - __clz(n));
        r = i & (n - 1);
    } else {
        q = i / n;
        r = i - q * n;
    }
}


constexpr int FP_N = 6;       // planes available to the front ends
constexpr int STRIP_BYTES = 36 * 1024;

struct FrontParams {
    int H, W, WPR, NW;
    uint32_t lastmask;
    int stage_rows;
    int mode;  // 0 canny, 1 inclusive, 2 enhanced, 3 brown spots, 4 saliency
    int canny_lo, canny_hi, canny_l2;  // integer thresholds already squared for L2
    int g15[15];                       // gaussian taps 15x15 sigma 0
    int g5[5];                         // gaussian taps 5x5 sigma cfg
    lfx_mask_cfg cfg;
    Footprint fp3, fp5, fp7, fp9, fpb;
    int rcap_glob;
    int planes_in_smem;   // 0: bit planes, grey image and word-base table live in the block's global scratch (large images)
    unsigned long long ws_per_block;
};

struct FrontMem {
    uint8_t* gray;     // [H*W]
    uint8_t* strip;    // STRIP_BYTES scratch
    uint8_t* stage;    // staged RGB rows
    HsvLut* hsv;
    LabLut* lab;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int refl101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// ------------------------------------------------------------------------------ grey / predicates
// Walk the RGB image through the staging buffer; FN(y, x, r, g, b) -> up to 3 predicate bits and
// optionally writes grey.  Bits are ballot-packed into planes.
template <typename FN>
__device__ void rgb_pass(const uint8_t* img, const FrontParams& P, const FrontMem& M, uint32_t* p0, uint32_t* p1,
                         uint32_t* p2, FN fn) {
    const int rb = P.W * 3;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int y0 = 0; y0 < P.H; y0 += P.stage_rows) {
        const int nrows = min(P.stage_rows, P.H - y0);
        block_load_bytes(M.stage, img + (size_t)y0 * rb, nrows * rb);
        __syncthreads();
        for (int item = wid; item < nrows * P.WPR; item += MT / 32) {
            int ry, w;
            divmod_fast(item, P.WPR, ry, w);
            const int x = w * 32 + lane;
            int bits = 0;
            if (x < P.W) {
                const uint8_t* px = M.stage + ry * rb + x * 3;
                bits = fn(y0 + ry, x, (int)px[0], (int)px[1], (int)px[2]);
            }
            const uint32_t m0 = __ballot_sync(0xffffffffu, bits & 1);
            const uint32_t m1 = __ballot_sync(0xffffffffu, bits & 2);
            const uint32_t m2 = __ballot_sync(0xffffffffu, bits & 4);
            if (lane == 0) {
                const int idx = (y0 + ry) * P.WPR + w;
                if (p0) p0[idx] = m0;
                if (p1) p1[idx] = m1;
                if (p2) p2[idx] = m2;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ Canny
// Sobel (BORDER_REPLICATE) -> magnitude -> non-maximum suppression with OpenCV's TG22 fixed point
// -> candidate / strong planes; hysteresis = 8-connected candidate components holding a strong pixel.
__device__ __forceinline__ void sobel_rep(const uint8_t* g, int H, int W, int y, int x, int& dx, int& dy) {
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const int a = g[ym * W + xm], b = g[ym * W + x], c = g[ym * W + xp];
    const int d = g[y * W + xm], f = g[y * W + xp];
    const int h = g[yp * W + xm], i = g[yp * W + x], j = g[yp * W + xp];
    dx = (c + 2 * f + j) - (a + 2 * d + h);
    dy = (h + 2 * i + j) - (a + 2 * b + c);
}

__device__ void canny(const uint8_t* gray, int lo, int hi, bool l2, uint32_t* cand, uint32_t* strong, uint32_t* edges,
                      const FrontParams& P, const FrontMem& M, Ctx& c) {
    const int H = P.H, W = P.W;
    int* mag = reinterpret_cast<int*>(M.strip);
    const int srows = max(1, STRIP_BYTES / (W * 4) - 2);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int ys = 0; ys < H; ys += srows) {
        const int ye = min(H, ys + srows);
        // magnitudes of rows ys-1 .. ye (zero outside the image)
        const int mrows = ye - ys + 2;
        for (int i = threadIdx.x; i < mrows * W; i += MT) {
            int ry, x;
            divmod_fast(i, W, ry, x);
            const int y = ys - 1 + ry;
            int m = 0;
            if (y >= 0 && y < H) {
                int dx, dy;
                sobel_rep(gray, H, W, y, x, dx, dy);
                m = l2 ? dx * dx + dy * dy : abs(dx) + abs(dy);
            }
            mag[i] = m;
        }
        __syncthreads();
        for (int item = wid; item < (ye - ys) * P.WPR; item += MT / 32) {
            int ry, w;
            divmod_fast(item, P.WPR, ry, w);
            const int y = ys + ry, x = w * 32 + lane;
            bool bc = false, bs = false;
            if (x < W) {
                const int* mr = mag + (ry + 1) * W;  // row y
                const int m = mr[x];
                if (m > lo) {
                    int dx, dy;
                    sobel_rep(gray, H, W, y, x, dx, dy);
                    const int ax = abs(dx), ay = abs(dy) << 15;
                    const int tg22x = ax * 13573;
                    bool keep;
                    if (ay < tg22x) {
                        const int l = x > 0 ? mr[x - 1] : 0, r = x < W - 1 ? mr[x + 1] : 0;
                        keep = (m > l) && (m >= r);
                    } else if (ay > tg22x + (ax << 16)) {
                        keep = (m > mr[x - W]) && (m >= mr[x + W]);
                    } else {
                        const int s = ((dx ^ dy) < 0) ? -1 : 1;
                        const int xa = x - s, xb = x + s;
                        const int d1 = (xa >= 0 && xa < W) ? mr[xa - W] : 0;
                        const int d2 = (xb >= 0 && xb < W) ? mr[xb + W] : 0;
                        keep = (m > d1) && (m > d2);
                    }
                    bc = keep;
                    bs = keep && (m > hi);
                }
            }
            const uint32_t mc = __ballot_sync(0xffffffffu, bc), ms = __ballot_sync(0xffffffffu, bs);
            if (lane == 0) {
                cand[y * P.WPR + w] = mc;
                strong[y * P.WPR + w] = ms;
            }
        }
        __syncthreads();
    }
    ccl<8>(cand, c);
    for (int r = threadIdx.x; r < c.R; r += MT) {
        const uint32_t g = c.geom[r];
        if (popc_range(strong + c.ry[r] * P.WPR, g & 0xFFFF, g >> 16)) atomicOr(&c.acc[c.parent[r]], 1);
    }
    plane_zero(edges, c);
    __syncthreads();
    keep_area_ge(edges, 1, c);  // acc[root] >= 1  <=>  the component holds a strong pixel
}

// ------------------------------------------------------------------------------ separable Gaussian on u8 in smem
// dst(y,x) test via callback: FN(y, x, blurred) -> predicate bit, ballot-packed into `plane`.
// K taps (15 or 5), BORDER_REFLECT_101, 8.8 then 16.16 fixed point.
template <int K, typename FN>
__device__ void gauss_gray_pass(const uint8_t* gray, const int* taps, uint32_t* plane, uint8_t* out_u8,
                                const FrontParams& P, const FrontMem& M, FN fn) {
    constexpr int R = K / 2;
    const int H = P.H, W = P.W;
    uint16_t* hbuf = reinterpret_cast<uint16_t*>(M.strip);
    const int srows = max(1, STRIP_BYTES / (W * 2) - 2 * R);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int ys = 0; ys < H; ys += srows) {
        const int ye = min(H, ys + srows);
        const int hrows = ye - ys + 2 * R;
        // horizontal pass, 4 pixels per thread.  Interior groups: aligned 32-bit loads of the byte window, the K
        // taps (all < 256) as dp4a byte weights on funnel-shifted 4-byte windows; groups that touch the left /
        // right border take the reflecting scalar path.
        const int G4 = (W + 3) >> 2;
        const bool vec_ok = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(gray) & 3) == 0);
        uint32_t wq[(K + 3) / 4];
#pragma unroll
        for (int j = 0; j < (K + 3) / 4; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (4 * j + b < K) v |= (uint32_t)taps[4 * j + b] << (8 * b);
            wq[j] = v;
        }
        constexpr int NQ = (K + 3) / 4;                          // dp4a per output
        constexpr int RP = (R + 3) & ~3;
        constexpr int NWORD = (3 + RP - R + 4 * NQ + 3) / 4;     // aligned words covering every 4-byte tap window of the 4 outputs
        // interior groups [gl, gr): the whole aligned window x-RP .. x-RP+4*NWORD-1 lies inside the row.  They get
        // their own loop so that warps stay convergent (a few border lanes would drag every warp down the slow path).
        const int gl = vec_ok ? RP / 4 : 0;
        const int gr = vec_ok ? max(gl, (W + RP - 4 * NWORD) / 4 + 1) : 0;
        const int nint = gr - gl, nedge = G4 - nint;
        for (int i = threadIdx.x; i < hrows * nint; i += MT) {
            const int ry = i / nint, x = (gl + i - ry * nint) * 4;
            const uint8_t* row = gray + refl101(ys - R + ry, H) * W;
            uint16_t* hb = hbuf + ry * W + x;
            const uint32_t* p = reinterpret_cast<const uint32_t*>(row + x - RP);
            uint32_t w[NWORD];
#pragma unroll
            for (int k = 0; k < NWORD; ++k) w[k] = p[k];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int off = e + (RP - R);   // first tap of output x+e inside the window (compile-time after unrolling)
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int bo = off + 4 * j, wi = bo >> 2, sh = (bo & 3) * 8;
                    const uint32_t win = sh ? __funnelshift_r(w[wi], w[wi + 1 < NWORD ? wi + 1 : wi], sh) : w[wi];
                    acc = __dp4a(win, wq[j], acc);
                }
                hb[e] = (uint16_t)acc;
            }
        }
        for (int i = threadIdx.x; i < hrows * nedge; i += MT) {
            const int ry = i / nedge, ge = i - ry * nedge;
            const int x = (ge < gl ? ge : gr + (ge - gl)) * 4;
            const uint8_t* row = gray + refl101(ys - R + ry, H) * W;
            uint16_t* hb = hbuf + ry * W + x;
            for (int e = 0; e < 4 && x + e < W; ++e) {
                int acc = 0;
#pragma unroll
                for (int t = 0; t < K; ++t) acc += row[refl101(x + e + t - R, W)] * taps[t];
                hb[e] = (uint16_t)acc;
            }
        }
        __syncthreads();
        for (int item = wid; item < (ye - ys) * P.WPR; item += MT / 32) {
            int ry, w;
            divmod_fast(item, P.WPR, ry, w);
            const int y = ys + ry, x = w * 32 + lane;
            bool bit = false;
            if (x < W) {
                // OpenCV's fixed-point kernel is symmetric: taps t and K-1-t share a multiply
                const uint16_t* hc = hbuf + ry * W + x;
                uint32_t acc = (uint32_t)hc[R * W] * (uint32_t)taps[R] + 32768u;
#pragma unroll
                for (int t = 0; t < R; ++t) acc += ((uint32_t)hc[t * W] + (uint32_t)hc[(K - 1 - t) * W]) * (uint32_t)taps[t];
                const int bl = (int)(acc >> 16);
                if (out_u8) out_u8[y * W + x] = (uint8_t)bl;
                bit = fn(y, x, bl);
            }
            const uint32_t mb = __ballot_sync(0xffffffffu, bit);
            if (lane == 0 && plane) plane[y * P.WPR + w] = mb;
        }
        __syncthreads();
    }
}

// keep the 8-connected component with most pixels (first label on ties); no-op when empty
__device__ void keep_largest8(uint32_t* m, uint32_t* tmp, Ctx& c) {
    ccl<8>(m, c);
    measure_area(c);
    if (threadIdx.x == 0) *c.s_best = 0ull;
    __syncthreads();
    for (int r = threadIdx.x; r < c.R; r += MT)
        if (c.parent[r] == r)
            atomicMax(c.s_best, ((unsigned long long)(uint32_t)c.acc[r] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)r));
    __syncthreads();
    const unsigned long long best = *c.s_best;
    if (best == 0ull) return;  // block-uniform
    const int win = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFu));
    plane_zero(tmp, c);
    __syncthreads();
    for (int r = threadIdx.x; r < c.R; r += MT)
        if (c.parent[r] == win) {
            const uint32_t g = c.geom[r];
            set_run(tmp, c.ry[r], g & 0xFFFF, g >> 16, c);
        }
    __syncthreads();
    plane_copy(m, tmp, c);
    __syncthreads();
}

__device__ void open_(uint32_t* m, uint32_t* t, const Footprint& fp, Ctx& c) {
    morph_any<false>(m, t, fp, c);
    __syncthreads();
    morph_any<true>(t, m, fp, c);
    __syncthreads();
}
__device__ void close_(uint32_t* m, uint32_t* t, const Footprint& fp, Ctx& c) {
    morph_any<true>(m, t, fp, c);
    __syncthreads();
    morph_any<false>(t, m, fp, c);
    __syncthreads();
}

__device__ void plane_out_bytes(const uint32_t* p, uint8_t* out, int channels, const FrontParams& P) {
    for (int i = threadIdx.x; i < P.H * P.W; i += MT) {
        int y, x;
        divmod_fast(i, P.W, y, x);
        const uint8_t v = ((p[y * P.WPR + (x >> 5)] >> (x & 31)) & 1) ? 255 : 0;
        for (int ch = 0; ch < channels; ++ch) out[(size_t)i * channels + ch] = v;
    }
}
__device__ void bytes_in_plane(const uint8_t* in, uint32_t* p, const FrontParams& P) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int item = wid; item < P.NW; item += MT / 32) {
        int y, w;
        divmod_fast(item, P.WPR, y, w);
        const int x = w * 32 + lane;
        const bool on = (x < P.W) && (__ldg(in + (size_t)y * P.W + x) > 0);
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) p[item] = m;
    }
}

// block-wide min/max of non-negative floats via int atomics on shared slots
__device__ __forceinline__ void fminmax_update(float v, bool ok, int* s_min, int* s_max) {
    float lo = ok ? v : __int_as_float(0x7f800000), hi = ok ? v : 0.f;   // all inputs are >= 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(s_min, __float_as_int(lo));
        atomicMax(s_max, __float_as_int(hi));
    }
}
// cv2.normalize(..., 0, 255, NORM_MINMAX) parameters
__device__ __forceinline__ void norm_params(float smin, float smax, double& scale, double& shift) {
    const double d = (double)smax - (double)smin;
    scale = 255.0 * (d > DBL_EPSILON ? 1.0 / d : 0.0);
    shift = 0.0 - (double)smin * scale;
}
__device__ __forceinline__ float norm_apply(float v, double scale, double shift) {
    return (float)__dadd_rn(__dmul_rn((double)v, scale), shift);
}

__global__ void __launch_bounds__(MT, 1) k_front(const uint8_t* __restrict__ src, const uint8_t* __restrict__ aux,
                                                 uint8_t* __restrict__ out, int32_t* __restrict__ stats,
                                                 uint8_t* __restrict__ ws, int B, const FrontParams P,
                                                 const LfxTables* __restrict__ tab) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ int s_tmp[40];
    __shared__ unsigned long long s_best;
    __shared__ int s_bb[8];
    __shared__ int s_hist[4];
    __shared__ int s_mm[8];

    Ctx c;
    ctx_init_geometry(c, P.H, P.W, P.WPR, P.NW, P.lastmask);
    c.s_tmp = s_tmp; c.s_best = &s_best; c.s_bb = s_bb; c.s_hist = s_hist;
    c.rcap_glob = P.rcap_glob;
    c.rcap_smem = RCAP_SMEM;
    uint8_t* sp = sm;
    uint8_t* gp = ws + (size_t)blockIdx.x * P.ws_per_block;
    auto take = [](uint8_t*& p, size_t bytes) {
        uint8_t* r = p;
        p += (bytes + 15) & ~(size_t)15;
        return r;
    };
    // images up to ~256x256 keep planes, grey image and word-base table in shared memory; larger ones use the block's
    // global scratch for them (L2-resident; every helper below takes plain pointers, block barriers order the accesses)
    uint32_t* PL[FP_N];
    for (int k = 0; k < FP_N; ++k) PL[k] = reinterpret_cast<uint32_t*>(P.planes_in_smem ? take(sp, (size_t)P.NW * 4) : take(gp, (size_t)P.NW * 4));
    for (int k = 0; k < NPLANES; ++k) c.plane[k] = PL[k < FP_N ? k : 0];
    c.wbase = reinterpret_cast<int*>(P.planes_in_smem ? take(sp, (size_t)(P.NW + 1) * 4) : take(gp, (size_t)(P.NW + 1) * 4));
    // run tables: RCAP_SMEM runs in shared memory, worst case (H * ceil(W/2)) in global scratch
    c.sm_parent = reinterpret_cast<int*>(take(sp, RCAP_SMEM * 4));
    c.sm_geom = reinterpret_cast<uint32_t*>(take(sp, RCAP_SMEM * 4));
    c.sm_acc = reinterpret_cast<int*>(take(sp, RCAP_SMEM * 4));
    c.sm_ry = reinterpret_cast<uint16_t*>(take(sp, RCAP_SMEM * 2));
    c.gl_parent = reinterpret_cast<int*>(take(gp, (size_t)P.rcap_glob * 4));
    c.gl_geom = reinterpret_cast<uint32_t*>(take(gp, (size_t)P.rcap_glob * 4));
    c.gl_acc = reinterpret_cast<int*>(take(gp, (size_t)P.rcap_glob * 4));
    c.gl_ry = reinterpret_cast<uint16_t*>(take(gp, (size_t)P.rcap_glob * 2));
    float* fscratch = reinterpret_cast<float*>(take(gp, 2 * (size_t)P.H * P.W * sizeof(float)));
    FrontMem M;
    M.gray = P.planes_in_smem ? take(sp, (size_t)P.H * P.W) : take(gp, (size_t)P.H * P.W);
    M.strip = take(sp, STRIP_BYTES);
    M.stage = take(sp, (size_t)P.stage_rows * P.W * 3);
    M.hsv = reinterpret_cast<HsvLut*>(take(sp, sizeof(HsvLut)));
    M.lab = reinterpret_cast<LabLut*>(take(sp, sizeof(LabLut)));
    load_hsv_lut(M.hsv, tab);
    load_lab_lut(M.lab, tab);
    __syncthreads();
    const size_t npx = (size_t)P.H * P.W;
    const lfx_mask_cfg& cf = P.cfg;

    for (int img = blockIdx.x; img < B; img += gridDim.x) {
        c.status = 0;
        if (P.mode == 0) {
            // ---------------- Canny on a grey image
            block_load_bytes(M.gray, src + img * npx, (int)npx);
            __syncthreads();
            canny(M.gray, P.canny_lo, P.canny_hi, P.canny_l2 != 0, PL[0], PL[1], PL[2], P, M, c);
            plane_out_bytes(PL[2], out + img * npx, 1, P);
        } else if (P.mode == 1) {
            // ---------------- _create_inclusive_mask (mask.py:727-831)
            const int elo = max(0, cf.green_lo - 10), ehi = min(179, cf.green_hi + 15);
            uint8_t* gray = M.gray;
            const int Wd = P.W;
            const HsvLut* hl = M.hsv;
            const LabLut* ll = M.lab;
            rgb_pass(src + img * npx * 3, P, M, PL[0], PL[1], PL[2], [=](int y, int x, int r, int g, int b) {
                int h, s, v, L, A, Bv;
                rgb2hsv(r, g, b, hl, h, s, v);
                rgb2lab(r, g, b, ll, L, A, Bv);
                gray[y * Wd + x] = (uint8_t)rgb2gray(r, g, b);
                const bool strong_green = (h >= elo) && (h <= ehi) && (s >= 30) && (v >= 30);
                // uint8 additions wrap at 256 (mask.py:753-757)
                const bool green_dom = (g > ((r + 15) & 255)) || (g > ((b + 15) & 255)) ||
                                       ((g > ((r + 5) & 255)) && (g > ((b + 5) & 255)) && (s >= 20));
                const bool lab_green = (A <= 125) && (Bv >= 120) && (L >= 20) && (L <= 240);
                const bool bg = ((s <= 25) && (v >= 50) && (v <= 220)) ||
                                ((h >= 120) && (h <= 160) && (s >= 20) && (r > g) && (b > g));
                return (int)(strong_green || green_dom || lab_green) | ((int)bg << 1) | ((int)(s <= 15) << 2);
            });
            // texture_diff = |gray - GaussianBlur(gray,15x15)| < 10  (mask.py:769-771, :784)
            gauss_gray_pass<15>(M.gray, P.g15, PL[3], nullptr, P, M,
                                [=](int y, int x, int bl) { return abs((int)gray[y * Wd + x] - bl) < 10; });
            for (int i = threadIdx.x; i < P.NW; i += MT) PL[1][i] |= (PL[2][i] & PL[3][i]);  // full background plane
            __syncthreads();
            canny(M.gray, 30, 100, false, PL[2], PL[3], PL[4], P, M, c);   // edges -> PL[4]
            morph_any<true>(PL[4], PL[5], P.fp3, c);                             // dilated edges
            __syncthreads();
            for (int i = threadIdx.x; i < P.NW; i += MT) PL[2][i] = (PL[0][i] | PL[5][i]) & ~PL[1][i] & valid_mask(c, i % P.WPR);
            __syncthreads();
            open_(PL[2], PL[3], P.fp3, c);
            close_(PL[2], PL[3], P.fp9, c);
            close_(PL[2], PL[3], P.fp7, c);
            keep_largest8(PL[2], PL[3], c);
            close_(PL[2], PL[3], P.fp5, c);
            plane_out_bytes(PL[2], out + img * npx, 1, P);
        } else if (P.mode == 2) {
            // ---------------- _create_enhanced_mask (mask.py:610-724); the +0.3*edges term never
            // crosses the float32 `> 0.3` threshold, so the mask is the vegetation union
            const HsvLut* hl = M.hsv;
            const LabLut* ll = M.lab;
            rgb_pass(src + img * npx * 3, P, M, PL[0], nullptr, nullptr, [=](int y, int x, int r, int g, int b) {
                int h, s, v, L, A, Bv;
                rgb2hsv(r, g, b, hl, h, s, v);
                rgb2lab(r, g, b, ll, L, A, Bv);
                const bool veg_hsv = (h >= cf.green_lo) && (h <= cf.green_hi) && (s >= 25) && (v >= 20) && (v <= 240);
                const bool veg_lab = (A <= 135) && (Bv >= 105) && (L >= 30) && (L <= 220);
                bool brown;
                if (cf.use_lab_brown)
                    brown = (A >= cf.lab_a_min - 10) && (Bv >= cf.lab_b_min - 10) && (L >= 20);
                else
                    brown = (((h >= cf.brown_lo) && (h <= cf.brown_hi + 20)) || ((h >= 160) && (h <= 180))) &&
                            (s >= cf.brown_s_min - 10) && (v <= cf.brown_v_max + 30);
                return (int)(veg_hsv || veg_lab || brown);
            });
            close_(PL[0], PL[1], P.fp7, c);
            open_(PL[0], PL[1], P.fp3, c);
            close_(PL[0], PL[1], P.fp9, c);
            keep_largest8(PL[0], PL[1], c);
            close_(PL[0], PL[1], P.fp3, c);
            plane_out_bytes(PL[0], out + img * npx, 1, P);
        } else if (P.mode == 3) {
            // ---------------- apply_brown_filter (brown.py:21-89): aux = leaf mask
            bytes_in_plane(aux + img * npx, PL[1], P);
            const HsvLut* hl = M.hsv;
            const LabLut* ll = M.lab;
            rgb_pass(src + img * npx * 3, P, M, PL[0], nullptr, nullptr, [=](int y, int x, int r, int g, int b) {
                int h, s, v;
                rgb2hsv(r, g, b, hl, h, s, v);
                if (cf.use_lab_brown) {
                    int L, A, Bv;
                    rgb2lab(r, g, b, ll, L, A, Bv);
                    return (int)((A >= cf.lab_a_min) && (Bv >= cf.lab_b_min));
                }
                return (int)((h >= cf.brown_lo) && (h <= cf.brown_hi) && (s >= cf.brown_s_min) && (v <= cf.brown_v_max));
            });
            if (threadIdx.x < 4) s_hist[threadIdx.x] = 0;
            __syncthreads();
            int leaf = 0;
            for (int i = threadIdx.x; i < P.NW; i += MT) {
                PL[0][i] &= PL[1][i];
                leaf += __popc(PL[1][i]);
            }
            atomicAdd(&s_hist[0], leaf);
            __syncthreads();
            open_(PL[0], PL[2], P.fpb, c);
            close_(PL[0], PL[2], P.fpb, c);
            ccl<8>(PL[0], c);
            measure_area(c);
            for (int r = threadIdx.x; r < c.R; r += MT)
                if (c.parent[r] == r && c.acc[r] >= cf.brown_min_area_px) {
                    atomicAdd(&s_hist[1], 1);
                    atomicAdd(&s_hist[2], c.acc[r]);
                }
            plane_zero(PL[2], c);
            __syncthreads();
            keep_area_ge(PL[2], cf.brown_min_area_px, c);
            plane_out_bytes(PL[2], out + img * npx, 1, P);
            if (threadIdx.x < 4) stats[(size_t)img * 4 + threadIdx.x] = threadIdx.x == 3 ? c.status : s_hist[threadIdx.x];
        } else {
            // ---------------- apply_blur_filter (blur.py:18-79): aux = leaf mask of make_mask(rgb)
            float* F0 = fscratch;
            float* F1 = F0 + npx;
            const uint8_t* rgb = src + img * npx * 3;
            uint8_t* gray = M.gray;
            const int Wd = P.W, Hd = P.H;
            bytes_in_plane(aux + img * npx, PL[0], P);                       // leaf
            const HsvLut* hl = M.hsv;
            rgb_pass(rgb, P, M, PL[1], nullptr, nullptr, [=](int y, int x, int r, int g, int b) {
                int h, s, v;
                rgb2hsv(r, g, b, hl, h, s, v);
                gray[y * Wd + x] = (uint8_t)rgb2gray(r, g, b);
                return (int)((h >= cf.brown_lo) && (h <= cf.brown_hi) && (s >= cf.brown_s_min) && (v <= cf.brown_v_max));
            });
            for (int i = threadIdx.x; i < P.NW; i += MT) PL[1][i] &= PL[0][i];   // brown & leaf
            __syncthreads();
            close_(PL[1], PL[2], P.fp3, c);
            morph_any<true>(PL[1], PL[2], P.fp3, c);
            __syncthreads();
            morph_any<true>(PL[2], PL[1], P.fp3, c);                              // PL[1] = brown_dilated
            __syncthreads();
            canny(M.gray, P.canny_lo, P.canny_hi, P.canny_l2 != 0, PL[2], PL[3], PL[4], P, M, c);
            morph_any<true>(PL[4], PL[2], P.fp3, c);                              // PL[2] = edges_dilated
            if (threadIdx.x < 8) s_mm[threadIdx.x] = (threadIdx.x & 1) ? 0 : 0x7f800000;   // {min,max} x 3 (+inf / 0)
            __syncthreads();
            // gradient magnitude (Sobel BORDER_REFLECT_101, float32) -> F0, min/max
            for (int i0 = 0; i0 < (int)npx; i0 += MT) {
                const int i = i0 + threadIdx.x;
                float mg = 0.f;
                const bool ok = i < (int)npx;
                if (ok) {
                    const int y = i / Wd, x = i - y * Wd;
                    const int ym = refl101(y - 1, Hd), yp = refl101(y + 1, Hd), xm = refl101(x - 1, Wd), xp = refl101(x + 1, Wd);
                    const int a = gray[ym * Wd + xm], b = gray[ym * Wd + x], cc = gray[ym * Wd + xp];
                    const int d = gray[y * Wd + xm], f = gray[y * Wd + xp];
                    const int h = gray[yp * Wd + xm], ii = gray[yp * Wd + x], j = gray[yp * Wd + xp];
                    const float gx = (float)((cc + 2 * f + j) - (a + 2 * d + h));
                    const float gy = (float)((h + 2 * ii + j) - (a + 2 * b + cc));
                    mg = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
                    F0[i] = mg;
                }
                fminmax_update(mg, ok, &s_mm[0], &s_mm[1]);
            }
            // colour difference: mean_c |rgb - GaussianBlur(rgb,15x15)| (float32) -> F1, min/max.
            // Channels are blurred one at a time through the grey-sized smem plane `chan`.
            for (int i = threadIdx.x; i < (int)npx; i += MT) F1[i] = 0.f;
            __syncthreads();
            for (int ch = 0; ch < 3; ++ch) {
                // channel plane into strip-independent storage: reuse the OUTPUT buffer as scratch u8
                uint8_t* cp = out + img * npx * 3 + (size_t)ch * npx;   // [H*W] bytes inside this image's output slab
                for (int i = threadIdx.x; i < (int)npx; i += MT) cp[i] = __ldg(rgb + (size_t)i * 3 + ch);
                __syncthreads();
                gauss_gray_pass<15>(cp, P.g15, nullptr, nullptr, P, M, [=](int y, int x, int bl) {
                    const int i = y * Wd + x;
                    const float dv = fabsf(__fadd_rn((float)cp[i], -(float)bl));
                    F1[i] = __fadd_rn(F1[i], dv);     // (a + b) + c, float32, channel order
                    return false;
                });
                __syncthreads();
            }
            for (int i0 = 0; i0 < (int)npx; i0 += MT) {
                const int i = i0 + threadIdx.x;
                const bool ok = i < (int)npx;
                float v = 0.f;
                if (ok) {
                    v = __fdiv_rn(F1[i], 3.0f);
                    F1[i] = v;
                }
                fminmax_update(v, ok, &s_mm[2], &s_mm[3]);
            }
            __syncthreads();
            double sc0, sh0, sc1, sh1;
            norm_params(__int_as_float(s_mm[0]), __int_as_float(s_mm[1]), sc0, sh0);
            norm_params(__int_as_float(s_mm[2]), __int_as_float(s_mm[3]), sc1, sh1);
            // saliency = 0.4*edges + 0.3*u8(norm(grad)) + 0.6*brown + 0.2*norm(cdiff)   (float32, that order)
            for (int i0 = 0; i0 < (int)npx; i0 += MT) {
                const int i = i0 + threadIdx.x;
                const bool ok = i < (int)npx;
                float sal = 0.f;
                if (ok) {
                    const int y = i / Wd, x = i - y * Wd;
                    const uint32_t bitpos = x & 31;
                    const int widx = y * P.WPR + (x >> 5);
                    const float e = ((PL[2][widx] >> bitpos) & 1) ? 255.f : 0.f;
                    const float br = ((PL[1][widx] >> bitpos) & 1) ? 255.f : 0.f;
                    const float gn = (float)(int)(uint8_t)(int)norm_apply(F0[i], sc0, sh0);   // .astype(uint8): truncate
                    sal = __fmul_rn(e, 0.4f);
                    sal = __fadd_rn(sal, __fmul_rn(gn, 0.3f));
                    sal = __fadd_rn(sal, __fmul_rn(br, 0.6f));
                    sal = __fadd_rn(sal, __fmul_rn(norm_apply(F1[i], sc1, sh1), 0.2f));
                    F0[i] = sal;
                }
                fminmax_update(sal, ok, &s_mm[4], &s_mm[5]);
            }
            __syncthreads();
            double sc2, sh2;
            norm_params(__int_as_float(s_mm[4]), __int_as_float(s_mm[5]), sc2, sh2);
            for (int i = threadIdx.x; i < (int)npx; i += MT) gray[i] = (uint8_t)(int)norm_apply(F0[i], sc2, sh2);
            __syncthreads();
            // GaussianBlur 5x5 sigma=cfg, zero outside the leaf, grey -> RGB
            uint8_t* o3 = out + img * npx * 3;
            const uint32_t* leafp = PL[0];
            const int WPRd = P.WPR;
            gauss_gray_pass<5>(M.gray, P.g5, nullptr, nullptr, P, M, [=](int y, int x, int bl) {
                const bool lf = (leafp[y * WPRd + (x >> 5)] >> (x & 31)) & 1;
                const uint8_t v = lf ? (uint8_t)bl : 0;
                uint8_t* o = o3 + (size_t)(y * Wd + x) * 3;
                o[0] = v; o[1] = v; o[2] = v;
                return false;
            });
            __syncthreads();
        }
        if (stats && P.mode != 3 && threadIdx.x == 0) stats[img] = c.status;
        __syncthreads();
    }
}

size_t front_plane_bytes(int H, int W) {
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const int WPR = (W + 31) / 32, NW = H * WPR;
    return al((size_t)NW * 4) * FP_N + al((size_t)(NW + 1) * 4) + al((size_t)H * W);
}

size_t front_smem(int H, int W, int stage_rows, bool planes_in_smem) {
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    return (planes_in_smem ? front_plane_bytes(H, W) : 0) + al(RCAP_SMEM * 4) * 3 + al(RCAP_SMEM * 2) + al(STRIP_BYTES) +
           al((size_t)stage_rows * W * 3) + al(sizeof(HsvLut)) + al(sizeof(LabLut));
}

bool front_planes_fit(int H, int W) { return front_smem(H, W, max(1, min(H, 6144 / (W * 3))), true) <= 225 * 1024; }

size_t front_ws_per_block(int H, int W) {
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t rcap = (size_t)H * ((W + 1) / 2);
    return al(rcap * 4) * 3 + al(rcap * 2) + al(2 * (size_t)H * W * sizeof(float)) + (front_planes_fit(H, W) ? 0 : front_plane_bytes(H, W) + 256);
}

int front_launch(const uint8_t* src, const uint8_t* aux, uint8_t* out, int32_t* stats, void* ws, size_t ws_bytes, int B, int H,
                 int W, FrontParams& P, cudaStream_t st, const char* what) {
    LFX_REQUIRE(H >= 3 && W >= 3 && B > 0 && H <= 65535 && W <= 65535, LFX_ERR_ARG, "%s: bad shape", what);
    P.rcap_glob = H * ((W + 1) / 2);
    P.ws_per_block = front_ws_per_block(H, W);
    LFX_REQUIRE(ws && ws_bytes >= P.ws_per_block * (size_t)min(B, LFX_NUM_SMS), LFX_ERR_WORKSPACE,
                "%s: workspace %zu < %zu bytes (lfx_front_workspace)", what, ws_bytes, P.ws_per_block * (size_t)min(B, LFX_NUM_SMS));
    P.H = H; P.W = W; P.WPR = (W + 31) / 32; P.NW = H * P.WPR;
    P.lastmask = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xFFFFFFFFu;
    P.stage_rows = max(1, min(H, 6144 / (W * 3)));
    P.planes_in_smem = front_planes_fit(H, W) ? 1 : 0;
    const size_t smem = front_smem(H, W, P.stage_rows, P.planes_in_smem != 0);
    LFX_REQUIRE(smem <= 225 * 1024 && (size_t)W * 4 * 3 <= STRIP_BYTES && (size_t)W * 2 * 16 <= STRIP_BYTES, LFX_ERR_UNSUPPORTED,
                "%s: image %dx%d: rows of more than %d pixels do not fit the strip buffer", what, H, W, STRIP_BYTES / 32);
    P.fp3 = make_ellipse(3); P.fp5 = make_ellipse(5); P.fp7 = make_ellipse(7); P.fp9 = make_ellipse(9);
    P.fpb = make_ellipse(P.cfg.brown_morph_kernel > 0 ? P.cfg.brown_morph_kernel : 3);
    int32_t t15[31], t5[31];
    lfx_gauss_taps(15, 0.0, t15);
    for (int i = 0; i < 15; ++i) P.g15[i] = t15[i];
    static size_t attr_[LFX_MAX_DEVICES] = {0};
    size_t& attr = attr_[lfx_dev()];
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(k_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "%s smem attr: %s", what, cudaGetErrorString(e));
        attr = smem;
    }
    (void)t5;
    const int grid = min(B, LFX_NUM_SMS);
    k_front<<<grid, MT, smem, st>>>(src, aux, out, stats, (uint8_t*)ws, B, P, lfx_tables());
    return lfx_check_launch(what);
}

// cv::Canny threshold preparation: integer compare thresholds (L2: squared, capped at 32767)
void canny_thresholds(double lo, double hi, int l2, int* ilo, int* ihi) {
    if (lo > hi) {
        const double t = lo;
        lo = hi;
        hi = t;
    }
    if (l2) {
        lo = fmin(32767.0, lo);
        hi = fmin(32767.0, hi);
        if (lo > 0) lo *= lo;
        if (hi > 0) hi *= hi;
    }
    *ilo = (int)floor(lo);
    *ihi = (int)floor(hi);
}

}  // namespace

extern "C" size_t lfx_front_workspace(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return front_ws_per_block(H, W) * (size_t)min(B, LFX_NUM_SMS) + 256;
}

extern "C" int lfx_canny(const uint8_t* gray, uint8_t* edges, int B, int H, int W, double low, double high, int l2gradient,
                         void* workspace, size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(gray && edges, LFX_ERR_ARG, "canny: NULL argument");
    FrontParams P;
    memset(&P, 0, sizeof(P));
    P.mode = 0;
    P.canny_l2 = l2gradient ? 1 : 0;
    canny_thresholds(low, high, l2gradient, &P.canny_lo, &P.canny_hi);
    P.cfg.brown_morph_kernel = 3;
    return front_launch(gray, nullptr, edges, nullptr, workspace, workspace_bytes, B, H, W, P, (cudaStream_t)stream, "canny");
}

extern "C" int lfx_raw_mask(const uint8_t* src, uint8_t* raw, int B, int H, int W, int which, const lfx_mask_cfg* cfg,
                            void* workspace, size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && raw && cfg && (which == 0 || which == 1), LFX_ERR_ARG, "raw_mask: bad arguments");
    FrontParams P;
    memset(&P, 0, sizeof(P));
    P.mode = which == 0 ? 1 : 2;
    P.cfg = *cfg;
    return front_launch(src, nullptr, raw, nullptr, workspace, workspace_bytes, B, H, W, P, (cudaStream_t)stream, "raw_mask");
}

extern "C" int lfx_brown_spots(const uint8_t* src, const uint8_t* mask, uint8_t* spots, int32_t* stats, int B, int H, int W,
                               const lfx_mask_cfg* cfg, void* workspace, size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && mask && spots && stats && cfg, LFX_ERR_ARG, "brown_spots: bad arguments");
    const int bk = cfg->brown_morph_kernel;
    LFX_REQUIRE(bk >= 1 && bk <= 19 && (bk & 1), LFX_ERR_UNSUPPORTED, "brown_spots: morph kernel %d", bk);
    FrontParams P;
    memset(&P, 0, sizeof(P));
    P.mode = 3;
    P.cfg = *cfg;
    return front_launch(src, mask, spots, stats, workspace, workspace_bytes, B, H, W, P, (cudaStream_t)stream, "brown_spots");
}

extern "C" int lfx_saliency_blur(const uint8_t* src, const uint8_t* mask, uint8_t* dst, int B, int H, int W,
                                 double gaussian_sigma, const lfx_mask_cfg* cfg, void* workspace, size_t workspace_bytes,
                                 lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && mask && dst && cfg, LFX_ERR_ARG, "saliency_blur: bad arguments");
    FrontParams P;
    memset(&P, 0, sizeof(P));
    P.mode = 4;
    P.cfg = *cfg;
    P.canny_l2 = 1;
    canny_thresholds(50, 150, 1, &P.canny_lo, &P.canny_hi);   // blur.py:30
    int32_t t5[31];
    const int rc = lfx_gauss_taps(5, gaussian_sigma, t5);
    if (rc) return rc;
    for (int i = 0; i < 5; ++i) P.g5[i] = t5[i];
    return front_launch(src, mask, dst, nullptr, workspace, workspace_bytes, B, H, W, P, (cudaStream_t)stream, "saliency_blur");
}
