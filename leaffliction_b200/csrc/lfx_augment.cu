// Class-balancing augmentations (srcs/preprocessing/image_augmenter.py): flip, rotate (nearest,
// 16.16 fixed point), bicubic affine/perspective warp (Pillow fp64 semantics), crop + Lanczos
// resize (22-bit fixed point, H pass -> u8 -> V pass) and distortion (noise + autocontrast).
// All integer / byte work: HBM-bound, no tensor cores.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "lfx_common.cuh"

namespace {

constexpr int THREADS = 256;

// ------------------------------------------------------------------------------ flip
// grid (row chunks, B); a block moves ROWS rows through shared memory and writes them mirrored.
constexpr int FLIP_SMEM = 32 * 1024;

// Vectorised flip for rows that are a multiple of 16 bytes with W % 4 == 0 (a row is then a whole number of
// 12-byte groups of 4 pixels).  A block moves FV_ROWS rows: 128-bit coalesced loads into shared memory;
// FLIP_TOP_BOTTOM just stores them to the mirrored rows, FLIP_LEFT_RIGHT reverses the pixel order of each row
// on 12-byte groups (3 words in, 4 byte-permutes, 3 words out: P0 P1 P2 P3 -> P3 P2 P1 P0) before the
// 128-bit stores.  Pure data movement: HBM-bound.
constexpr int FV_ROWS = 16;
__global__ void __launch_bounds__(THREADS) k_flip_vec(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                      const int32_t* __restrict__ mode, const int32_t* __restrict__ sidx) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int img = blockIdx.y;
    const int simg_i = sidx ? sidx[img] : img;   // source image of task `img` (dataset_balancer.py:116: random.choice(source_images))
    const int y0 = blockIdx.x * FV_ROWS;
    const int rows = min(FV_ROWS, H - y0);
    const int rb = W * 3, rb16 = rb >> 4;
    const uint4* simg = reinterpret_cast<const uint4*>(src + ((size_t)simg_i * H + y0) * rb);
    uint8_t* dimg = dst + (size_t)img * H * rb;
    const int m = mode[img];
    uint4* s4 = reinterpret_cast<uint4*>(sm);
    if (m != 0) {  // FLIP_TOP_BOTTOM: no shared memory needed
        for (int i = threadIdx.x; i < rows * rb16; i += THREADS) {
            const int y = i / rb16, j = i - y * rb16;
            st_stream16(dimg + (size_t)(H - 1 - (y0 + y)) * rb + (size_t)j * 16, ld_stream16(simg + i));
        }
        return;
    }
    for (int i = threadIdx.x; i < rows * rb16; i += THREADS) s4[i] = ld_stream16(simg + i);
    __syncthreads();
    uint32_t* s_in = reinterpret_cast<uint32_t*>(sm);
    uint32_t* s_out = s_in + (size_t)FV_ROWS * (rb >> 2);
    const int G = W >> 2;  // 12-byte groups per row
    for (int i = threadIdx.x; i < rows * G; i += THREADS) {
        const int y = i / G, g = i - y * G;
        const uint32_t* p = s_in + (size_t)y * (rb >> 2) + 3 * (G - 1 - g);
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        uint32_t* o = s_out + (size_t)y * (rb >> 2) + 3 * g;
        o[0] = __byte_perm(w2, w1, 0x6321);                                // R3 G3 B3 R2
        o[1] = __byte_perm(__byte_perm(w1, w2, 0x0043), w0, 0x3710);       // G2 B2 R1 G1
        o[2] = __byte_perm(w0, w1, 0x2105);                                // B1 R0 G0 B0
    }
    __syncthreads();
    const uint4* o4 = reinterpret_cast<const uint4*>(s_out);
    uint8_t* drow = dimg + (size_t)y0 * rb;
    for (int i = threadIdx.x; i < rows * rb16; i += THREADS) st_stream16(drow + (size_t)i * 16, o4[i]);
}



__global__ void __launch_bounds__(THREADS) k_flip(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                  int rows_per_block, const int32_t* __restrict__ mode,
                                                  const int32_t* __restrict__ sidx) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int img = blockIdx.y;
    const int y0 = blockIdx.x * rows_per_block;
    const int rows = min(rows_per_block, H - y0);
    if (rows <= 0) return;
    const int rb = W * 3;
    uint8_t* s_in = sm;
    uint8_t* s_out = sm + ((rows_per_block * rb + 15) & ~15);
    const uint8_t* simg = src + (size_t)(sidx ? sidx[img] : img) * H * rb;
    uint8_t* dimg = dst + (size_t)img * H * rb;
    block_load_bytes(s_in, simg + (size_t)y0 * rb, rows * rb);
    __syncthreads();
    const int m = mode[img];
    if (m == 0) {  // FLIP_LEFT_RIGHT: rows stay, pixels reverse
        for (int i = threadIdx.x; i < rows * rb; i += THREADS) {
            const int y = i / rb, o = i - y * rb;
            const int x = o / 3, c = o - x * 3;
            s_out[i] = s_in[y * rb + (W - 1 - x) * 3 + c];
        }
        __syncthreads();
        block_store_bytes(dimg + (size_t)y0 * rb, s_out, rows * rb);
    } else {  // FLIP_TOP_BOTTOM: rows reverse; the chunk lands at H - y0 - rows
        for (int i = threadIdx.x; i < rows * rb; i += THREADS) {
            const int y = i / rb, o = i - y * rb;
            s_out[i] = s_in[(rows - 1 - y) * rb + o];
        }
        __syncthreads();
        block_store_bytes(dimg + (size_t)(H - y0 - rows) * rb, s_out, rows * rb);
    }
}

// ------------------------------------------------------------------------------ rotate (nearest)
// A block produces RT_PX consecutive pixels of the FLAT [nh*nw] output array.  Lane-contiguous pixels: the 32 gathers
// of one load instruction walk along a rotated scanline, 3 bytes apart times cos -- a handful of 32-byte sectors per
// request (one thread owning 4 consecutive pixels spread every request over 32 sectors and the kernel was bound by
// L1 sector throughput).  The bytes are assembled in shared memory and leave as 16-byte coalesced stores.
constexpr int RT_STEPS = 8;
constexpr int RT_PX = THREADS * RT_STEPS;

// SMALL: the image has fewer than 2^31 bytes and the output fewer than 2^31 pixels -- every index is 32-bit (the common
// case; the 64-bit instantiation exists for completeness).
template <bool SMALL>
__global__ void __launch_bounds__(THREADS) k_rotate_nn(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       long long dst_stride, int H, int W,
                                                       const int32_t* __restrict__ params, int fill,
                                                       const int32_t* __restrict__ sidx) {
    __shared__ __align__(16) uint8_t s_px[RT_PX * 3];
    const int img = blockIdx.y;
    const int4 pa = *reinterpret_cast<const int4*>(params + img * 8);
    const int4 pb = *reinterpret_cast<const int4*>(params + img * 8 + 4);
    const int a0 = pa.x, a1 = pa.y, a2 = pa.z, a3 = pa.w, a4 = pb.x, a5 = pb.y;
    const int nw = pb.z, nh = pb.w;
    const long long npx = (long long)nw * nh;
    const long long q0 = (long long)blockIdx.x * RT_PX;
    if (q0 >= npx) return;
    const uint8_t* simg = src + (size_t)(sidx ? sidx[img] : img) * H * W * 3;
    uint8_t* dimg = dst + (size_t)img * dst_stride;
    const int nvalid = (int)min((long long)RT_PX, npx - q0);
    int x, y;
    if (SMALL) {
        const int q = (int)q0 + threadIdx.x;
        y = q / nw;
        x = q - y * nw;
    } else {
        const long long q = q0 + threadIdx.x;
        y = (int)(q / nw);
        x = (int)(q - (long long)y * nw);
    }
    // gather phase: all loads of the RT_STEPS pixels are issued before the first use (addresses of outside pixels are
    // redirected to byte 0 of the image and the fill colour is selected afterwards), then the bytes go to shared memory
    uint32_t px[RT_STEPS];   // r | g << 8 | b << 16, or the fill colour
    const uint32_t fill3 = (uint32_t)(fill & 255) * 0x010101u;
    bool in[RT_STEPS];
#pragma unroll
    for (int k = 0; k < RT_STEPS; ++k) {
        // libImaging affine_fixed: 32-bit int arithmetic, arithmetic shift
        const int xin = (a2 + y * a1 + x * a0) >> 16;
        const int yin = (a5 + y * a4 + x * a3) >> 16;
        in[k] = (unsigned)xin < (unsigned)W && (unsigned)yin < (unsigned)H;
        const uint8_t* s;
        if (SMALL) s = simg + (unsigned)(in[k] ? (yin * W + xin) * 3 : 0);
        else s = simg + (in[k] ? ((size_t)yin * W + xin) * 3 : (size_t)0);
        px[k] = (uint32_t)__ldg(s) | ((uint32_t)__ldg(s + 1) << 8) | ((uint32_t)__ldg(s + 2) << 16);
        x += THREADS;
        while (x >= nw) {
            x -= nw;
            ++y;
        }
    }
#pragma unroll
    for (int k = 0; k < RT_STEPS; ++k) {
        const uint32_t v = in[k] ? px[k] : fill3;
        uint8_t* o = s_px + (k * THREADS + threadIdx.x) * 3;   // (the tail past nvalid is written too: it is never stored)
        o[0] = (uint8_t)v;
        o[1] = (uint8_t)(v >> 8);
        o[2] = (uint8_t)(v >> 16);
    }
    __syncthreads();
    uint8_t* d = dimg + q0 * 3;
    const int nbytes = nvalid * 3;
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        const int n16 = nbytes >> 4;
        for (int i = threadIdx.x; i < n16; i += THREADS) st_stream16(d + i * 16, reinterpret_cast<const uint4*>(s_px)[i]);
        for (int i = (n16 << 4) + threadIdx.x; i < nbytes; i += THREADS) d[i] = s_px[i];
    } else {
        for (int i = threadIdx.x; i < nbytes; i += THREADS) d[i] = s_px[i];
    }
}

// ------------------------------------------------------------------------------ bicubic warp
// Pillow Geometry.c bicubic_filter32RGB in fp64 with round-to-nearest mul/add kept separate
// (x86-64 Pillow wheels do not contract to FMA).  A cheap fp32 evaluation decides first; only
// results that land within LFX_WARP_EPS of a truncation boundary are re-evaluated in fp64, so
// the output is bit-identical to the fp64 path at a fraction of its cost.
// fp32 error bound: one Horner evaluation rounds three products / three sums at magnitudes below 2048 (half ulp
// 6.1e-5 each, the first product 3.1e-5) -> < 4e-4; the cubic is linear in its four values, so the row errors reach
// the column evaluation weighted by |w1|+..+|w4| < 1.7 -> 6.8e-4, plus the column's own 4e-4: < 1.1e-3 in total.
// 2^-9 keeps a 1.8x margin and sends 0.4 % of the values down the fp64 path.
// The fp64 re-evaluations are DEFERRED: a risky value is pushed on a per-block queue and the whole block evaluates the
// queue after the band (one value per thread).  Evaluated in place, one risky lane made its whole warp walk the fp64
// code -- 12 % of all warp-level evaluations for 0.4 % of the values.
#define LFX_WARP_EPS 0.001953125f
constexpr int WB_QCAP = 1024;  // queue entries (u16 index of the value inside the band); overflow is evaluated in place

__device__ __forceinline__ double cubic64(double v1, double v2, double v3, double v4, double d) {
    const double p1 = v2;
    const double p2 = __dadd_rn(-v1, v3);
    const double p3 = __dadd_rn(__dadd_rn(__dmul_rn(2.0, __dadd_rn(v1, -v2)), v3), -v4);
    const double p4 = __dadd_rn(__dadd_rn(__dadd_rn(-v1, v2), -v3), v4);
    return __dadd_rn(p1, __dmul_rn(d, __dadd_rn(p2, __dmul_rn(d, __dadd_rn(p3, __dmul_rn(d, p4))))));
}
__device__ __forceinline__ float biased(uint8_t v) { return __uint_as_float(0x4B000000u | v); }  // 2^23 + v: a LOP3, no I2F
// The same cubic in DIFFERENCE form: cubic(v1..v4, d) = v2 + Wa (v1 - v2) + Wb (v3 - v2) + Wc (v4 - v2) with
// Wa = -d + 2d^2 - d^3, Wb = d + d^2 - d^3, Wc = -d^2 + d^3 (collect Pillow's Horner form by tap).  The weights depend on
// the fractional offset only -- one set per output column (horizontal pass) and per output row (vertical pass) -- and an
// evaluation is 3 differences + 3 FMAs instead of 11 operations; differences of biased taps are exact.
// fp32 error: weights carry <= 3 roundings (2e-7 relative, |W| <= 1), a horizontal value < 2e-4 (differences <= 255), a
// vertical one 1.7 x that + 2e-4 < 5.4e-4: inside the 1.1e-3 budget LFX_WARP_EPS was chosen for.  d = 0 gives W = 0 and
// the value v2 exactly, as in the Horner form.
struct CubW {
    float a, b, c;
};
__device__ __forceinline__ CubW cubic_weights(float d) {
    CubW w;
    w.a = d * (-1.f + d * (2.f - d));
    w.b = d * (1.f + d * (1.f - d));
    w.c = d * d * (d - 1.f);
    return w;
}
__device__ __forceinline__ float cubw_b(float v1, float v2, float v3, float v4, const CubW& w) {   // 2^23-biased taps
    return (v2 - 8388608.f) + w.a * (v1 - v2) + w.b * (v3 - v2) + w.c * (v4 - v2);
}
__device__ __forceinline__ float cubw(float v1, float v2, float v3, float v4, const CubW& w) {
    return v2 + w.a * (v1 - v2) + w.b * (v3 - v2) + w.c * (v4 - v2);
}
// Clamped, truncated byte of the fp32 value + whether fp64 must decide (Geometry.c: v <= 0 -> 0, v >= 255 -> 255, else
// (UINT8)v).  With a = f - 0.5, a + 1.5 * 2^23 rounds a to the nearest integer, which is floor(f) unless f is an
// integer (a tie) -- a value nobody trusts anyway: risky <=> f within EPS of an integer (either side), i.e.
// |a - rint(a)| > 0.5 - EPS.  The integer sits in the low bits of the sum (biased by 2^22), so clamping the raw bits to
// [bits(M), bits(M) + 255] is the clamp to 0..255 -- no F2I / FRND (quarter-rate pipe), no select.  Values beyond the
// range by more than EPS are decided here (0 or 255); NaN compares false and goes to fp64.
__device__ __forceinline__ uint8_t warp_trunc(float f, bool& risky) {
    const float a = f - 0.5f;
    const float t = a + 12582912.f;
    const float d = a - (t - 12582912.f);
    risky = !(fabsf(d) <= 0.5f - LFX_WARP_EPS);
    const int bits = min(max(__float_as_int(t), 0x4B400000), 0x4B4000FF);
    return (uint8_t)bits;
}
// The 4 x 3 taps of one row of a PADDED staged rectangle (two replicated pixels left and right of every row, so no
// column clamp): 12 consecutive bytes from byte offset boff = wb + s, fetched as four aligned words, aligned by three
// funnel shifts and turned into 2^23-biased floats by one PRMT each -- 19 instructions instead of 12 byte loads + 12
// adds + the clamps.  t[3 * tap + channel].
__device__ __forceinline__ void taps12_biased(const uint8_t* wp, uint32_t s8, float t[12]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(wp);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    const uint32_t n0 = __funnelshift_r(w0, w1, s8), n1 = __funnelshift_r(w1, w2, s8), n2 = __funnelshift_r(w2, w3, s8);
    constexpr uint32_t C = 0x4B000000u;
    t[0] = __uint_as_float(__byte_perm(n0, C, 0x7540)); t[1] = __uint_as_float(__byte_perm(n0, C, 0x7541));
    t[2] = __uint_as_float(__byte_perm(n0, C, 0x7542)); t[3] = __uint_as_float(__byte_perm(n0, C, 0x7543));
    t[4] = __uint_as_float(__byte_perm(n1, C, 0x7540)); t[5] = __uint_as_float(__byte_perm(n1, C, 0x7541));
    t[6] = __uint_as_float(__byte_perm(n1, C, 0x7542)); t[7] = __uint_as_float(__byte_perm(n1, C, 0x7543));
    t[8] = __uint_as_float(__byte_perm(n2, C, 0x7540)); t[9] = __uint_as_float(__byte_perm(n2, C, 0x7541));
    t[10] = __uint_as_float(__byte_perm(n2, C, 0x7542)); t[11] = __uint_as_float(__byte_perm(n2, C, 0x7543));
}
constexpr int WB_PAD = 16;   // bytes before / after every padded row (the two replicated pixels sit next to the row)
// row pitch of the padded rectangle: a multiple of 128 bytes, so that equal word columns of different rows share a bank and
// the lanes of a warp (24 distinct word columns) conflict only where two neighbours read the same word of different rows
__host__ __device__ constexpr int warp_pad_pitch(int W) { return (W * 3 + 2 * WB_PAD + 127) & ~127; }

// Source coordinates of output pixel (x, y) exactly as Pillow computes them; false = outside (pixel stays 0).
__device__ __forceinline__ bool warp_coords(int x, int y, int H, int W, const double* a, bool is_persp, int& xf, int& yf, double& dx,
                                            double& dy) {
    const double xc = (double)x + 0.5, yc = (double)y + 0.5;
    double xin = __dadd_rn(__dadd_rn(__dmul_rn(a[0], xc), __dmul_rn(a[1], yc)), a[2]);
    double yin = __dadd_rn(__dadd_rn(__dmul_rn(a[3], xc), __dmul_rn(a[4], yc)), a[5]);
    if (is_persp) {
        const double den = __dadd_rn(__dadd_rn(__dmul_rn(a[6], xc), __dmul_rn(a[7], yc)), 1.0);
        xin = __ddiv_rn(xin, den);
        yin = __ddiv_rn(yin, den);
    }
    if (xin < 0.0 || xin >= (double)W || yin < 0.0 || yin >= (double)H) return false;
    xin = __dadd_rn(xin, -0.5);
    yin = __dadd_rn(yin, -0.5);
    xf = (int)floor(xin);
    yf = (int)floor(yin);
    dx = __dadd_rn(xin, -(double)xf);
    dy = __dadd_rn(yin, -(double)yf);
    return true;
}

// Rows 1..3 of the 4x4 window reuse the previous row's VALUE when they fall outside the image (Geometry.c
// BICUBIC_BODY) and row 0 is clamped; the previous row is then itself the clamped row, so the rule equals clamping the
// row index.
// Where the taps come from: the image in global memory (pitch W*3, r0 = cb = 0) or a rectangle of it staged in shared
// memory (rows r0.., byte column cb.. of every row).  Tap (row, col, c) sits at base[(row - r0) * pitch + col*3 - cb + c].
struct WarpView {
    const uint8_t* base;
    int pitch, r0, cb;
};
__device__ __forceinline__ int warp_row_off(const WarpView& v, int yy, int H) { return (min(max(yy, 0), H - 1) - v.r0) * v.pitch - v.cb; }

// The reference value (fp64) of channel c of output pixel (x, y).
__device__ __noinline__ uint8_t bicubic_value64(const uint8_t* base, int pitch, int r0, int cb, int H, int W, int x, int y, int c,
                                                const double* a, bool is_persp) {
    int xf, yf;
    double dx, dy;
    if (!warp_coords(x, y, H, W, a, is_persp, xf, yf, dx, dy)) return 0;
    const WarpView v{base, pitch, r0, cb};
    int xo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) xo[t] = min(max(xf - 1 + t, 0), W - 1) * 3 + c;
    double dv[4];
#pragma unroll
    for (int rj = 0; rj < 4; ++rj) {
        const uint8_t* rp = base + warp_row_off(v, yf - 1 + rj, H);
        dv[rj] = cubic64(rp[xo[0]], rp[xo[1]], rp[xo[2]], rp[xo[3]], dx);
    }
    const double r = cubic64(dv[0], dv[1], dv[2], dv[3], dy);
    return r <= 0.0 ? 0 : (r >= 255.0 ? 255 : (uint8_t)(int)r);
}

// Row coordinate of an a1 = 0 map, yin = (a3*xc + a4*yc) + a5 with the generic path's operations in the generic path's
// order: cubic weights (a, b, c) of the fractional part and, in .w, the bits of floor(yin - 0.5) (INT_MIN: outside).
__device__ __noinline__ float4 warp_row_coords64(double tcol, double trow, double a5, int H) {
    double yin = __dadd_rn(__dadd_rn(tcol, trow), a5);
    const bool yok = !(yin < 0.0 || yin >= (double)H);
    yin = __dadd_rn(yin, -0.5);
    const int yf = (int)floor(yin);
    const CubW w = cubic_weights((float)__dadd_rn(yin, -(double)yf));
    return make_float4(w.a, w.b, w.c, __int_as_float(yok ? yf : INT_MIN));
}

// One output pixel of PIL's transform(..., BICUBIC) by the fp32 fast path: res[] holds the truncated fp32 values,
// the return value has bit c set when channel c must be re-evaluated in fp64.
__device__ __forceinline__ uint32_t bicubic_pixel(const WarpView& v, int H, int W, int x, int y, const double* a, bool is_persp,
                                                  uint8_t res[3]) {
    res[0] = res[1] = res[2] = 0;
    int xf, yf;
    double dx, dy;
    if (!warp_coords(x, y, H, W, a, is_persp, xf, yf, dx, dy)) return 0u;
    int xo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) xo[t] = min(max(xf - 1 + t, 0), W - 1) * 3;
    const CubW wx = cubic_weights((float)dx), wy = cubic_weights((float)dy);
    uint32_t risky = 0u;
    // cubic(v1..v4, 0) == v2 exactly (v2 + 0 * finite), in fp32 and in fp64 alike.  The reference's shear maps one axis
    // with the identity (image_augmenter.py:82: yin = yc or xin = xc), so a whole image has dy == 0 or dx == 0:
    // dy == 0 needs only the row yf (always inside the image), dx == 0 only the column xf of each row.
    if (dy == 0.0) {
        const uint8_t* rp = v.base + warp_row_off(v, yf, H);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float fv = cubw_b(biased(rp[xo[0] + c]), biased(rp[xo[1] + c]), biased(rp[xo[2] + c]), biased(rp[xo[3] + c]), wx);
            bool rk;
            res[c] = warp_trunc(fv, rk);
            risky |= rk ? (1u << c) : 0u;
        }
        return risky;
    }
    const uint8_t* rp0 = v.base + warp_row_off(v, yf - 1, H);
    const uint8_t* rp1 = v.base + warp_row_off(v, yf, H);
    const uint8_t* rp2 = v.base + warp_row_off(v, yf + 1, H);
    const uint8_t* rp3 = v.base + warp_row_off(v, yf + 2, H);
    if (dx == 0.0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // biased taps: every difference is an exact integer, so this equals the cubic on the plain values
            const float fv = cubw_b(biased(rp0[xo[1] + c]), biased(rp1[xo[1] + c]), biased(rp2[xo[1] + c]), biased(rp3[xo[1] + c]), wy);
            bool rk;
            res[c] = warp_trunc(fv, rk);
            risky |= rk ? (1u << c) : 0u;
        }
        return risky;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v0 = cubw_b(biased(rp0[xo[0] + c]), biased(rp0[xo[1] + c]), biased(rp0[xo[2] + c]), biased(rp0[xo[3] + c]), wx);
        const float v1 = cubw_b(biased(rp1[xo[0] + c]), biased(rp1[xo[1] + c]), biased(rp1[xo[2] + c]), biased(rp1[xo[3] + c]), wx);
        const float v2 = cubw_b(biased(rp2[xo[0] + c]), biased(rp2[xo[1] + c]), biased(rp2[xo[2] + c]), biased(rp2[xo[3] + c]), wx);
        const float v3 = cubw_b(biased(rp3[xo[0] + c]), biased(rp3[xo[1] + c]), biased(rp3[xo[2] + c]), biased(rp3[xo[3] + c]), wx);
        const float fv = cubw(v0, v1, v2, v3, wy);
        bool rk;
        res[c] = warp_trunc(fv, rk);
        risky |= rk ? (1u << c) : 0u;
    }
    return risky;
}

// grid (column tiles x bands of WB_ROWS output rows, B): a block owns a WB_ROWS x <= WB_COLS output tile.  For affine maps
// (a6 = a7 = 0: the reference's skew and shear) the source rectangle a tile can touch follows from its four corners;
// it is staged in shared memory with 16-byte loads and the taps come from there.  A rectangle that does not fit
// (tall spans of a vertical shear) is handled by splitting the tile into 2 / 4 / 8 column slices staged one after the
// other; general perspective maps read their taps from global memory.
constexpr int WB_ROWS = 32;
constexpr int WB_COLS = 256;

// source rectangle (rows r0..r1, columns c0..c1) of output pixels [xa, xb) x [ya, yb) under the affine map a.
// Conservative bounds are enough, so this runs in fp32 (coordinates < 2^20: error < 0.1 px, covered by the slack).
__device__ __forceinline__ void warp_src_rect(const double* ad, int xa, int xb, int ya, int yb, int H, int W, int& r0, int& r1, int& c0,
                                              int& c1) {
    const float a[6] = {(float)ad[0], (float)ad[1], (float)ad[2], (float)ad[3], (float)ad[4], (float)ad[5]};
    const float xs0 = (float)xa + 0.5f, xs1 = (float)xb - 0.5f, ys0 = (float)ya + 0.5f, ys1 = (float)yb - 0.5f;
    const float vy00 = a[3] * xs0 + a[4] * ys0 + a[5], vy01 = a[3] * xs0 + a[4] * ys1 + a[5];
    const float vy10 = a[3] * xs1 + a[4] * ys0 + a[5], vy11 = a[3] * xs1 + a[4] * ys1 + a[5];
    const float vx00 = a[0] * xs0 + a[1] * ys0 + a[2], vx01 = a[0] * xs0 + a[1] * ys1 + a[2];
    const float vx10 = a[0] * xs1 + a[1] * ys0 + a[2], vx11 = a[0] * xs1 + a[1] * ys1 + a[2];
    const float ylo = fminf(fminf(vy00, vy01), fminf(vy10, vy11)), yhi = fmaxf(fmaxf(vy00, vy01), fmaxf(vy10, vy11));
    const float xlo = fminf(fminf(vx00, vx01), fminf(vx10, vx11)), xhi = fmaxf(fmaxf(vx00, vx01), fmaxf(vx10, vx11));
    // taps use floor(v - 0.5) - 1 .. + 2 (2.5 below, 2.5 above) + slack; clamped taps of outside coordinates land on the
    // border row / column, which the clamps below keep inside the rectangle
    r0 = (int)fminf(fmaxf(ylo - 3.75f, 0.f), (float)(H - 1));
    r1 = max(r0, (int)fmaxf(fminf(yhi + 3.75f, (float)(H - 1)), 0.f));
    c0 = (int)fminf(fmaxf(xlo - 3.75f, 0.f), (float)(W - 1));
    c1 = max(c0, (int)fmaxf(fminf(xhi + 3.75f, (float)(W - 1)), 0.f));
}

// TILED = false is the W <= WB_COLS instantiation: one tile per band, whole rows staged, no slices (a rectangle that
// does not fit falls back to global taps) -- the column bookkeeping folds away at compile time.  When every row start
// is 16-byte aligned the rows are staged PADDED (pitch W*3 + 32: two replicated border pixels on either side), so the
// specialised paths fetch their taps without column clamps (taps12_biased).
template <bool TILED>
__global__ void __launch_bounds__(THREADS, 3) k_warp_bicubic(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H,
                                                          int W, const double* __restrict__ coef,
                                                          const int32_t* __restrict__ persp, int smem_cap, int ntx,
                                                          const int32_t* __restrict__ sidx, int nsrc, uint32_t m32) {
    extern __shared__ __align__(16) uint8_t s_rows[];
    __shared__ uint16_t s_queue[WB_QCAP];
    __shared__ int s_qn;
    // axis-aligned maps: the cubic weights (a, b, c) of each band row and, in .w, the bits of floor(yin - 0.5) (INT_MIN: row
    // outside the source)
    __shared__ float4 s_wy[WB_ROWS];
    __shared__ double s_trow[WB_ROWS]; // a4 * yc of each band row (maps whose yin also depends on x add their column term)
    const int img = blockIdx.y;
    const int band = TILED ? blockIdx.x / ntx : blockIdx.x, tx = TILED ? blockIdx.x - band * ntx : 0;
    const int y0 = band * WB_ROWS, y1 = min(H, y0 + WB_ROWS);
    const int X0 = TILED ? tx * WB_COLS : 0, X1 = TILED ? min(W, X0 + WB_COLS) : W;
    const double* ap = coef + img * 8;
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ap[i];
    const bool affine = (a[6] == 0.0 && a[7] == 0.0);
    const bool is_persp = (persp[img] != 0) && !affine;   // x / 1.0 == x exactly: the divide is skipped
    const size_t npx = (size_t)H * W;
    const uint8_t* simg = src + (size_t)(sidx ? sidx[img] : img) * npx * 3;
    uint8_t* dimg = dst + (size_t)img * npx * 3;
    const bool al16 = ((W * 3) % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);   // every row start is 16-byte aligned
    if (threadIdx.x == 0) s_qn = 0;
    if (threadIdx.x < WB_ROWS) {   // row coordinates of an axis-aligned map depend on y only: once per band, not per pixel
        const double yc = (double)(y0 + (int)threadIdx.x) + 0.5;
        s_trow[threadIdx.x] = __dmul_rn(a[4], yc);
        double yin = __dadd_rn(__dmul_rn(a[4], yc), a[5]);
        const bool yok = !(yin < 0.0 || yin >= (double)H);
        yin = __dadd_rn(yin, -0.5);
        const int yf = (int)floor(yin);
        const CubW w = cubic_weights((float)__dadd_rn(yin, -(double)yf));
        s_wy[threadIdx.x] = make_float4(w.a, w.b, w.c, __int_as_float(yok ? yf : INT_MIN));
    }
    // number of column slices: the smallest of 1, 2, 4, 8 whose source rectangles all fit in shared memory
    int nsl = 0;
    bool pad = false;
    if (!TILED) {
        if (affine && smem_cap > 0) {
            int r0, r1, c0, c1;
            warp_src_rect(a, 0, W, y0, y1, H, W, r0, r1, c0, c1);
            pad = al16 && (r1 - r0 + 1) * warp_pad_pitch(W) <= smem_cap;
            nsl = (pad || (r1 - r0 + 1) * W * 3 <= smem_cap) ? 1 : 0;
        }
    } else if (affine && smem_cap > 0) {
        for (int n = 1; n <= 8 && !nsl; n *= 2) {
            const int sw = (((X1 - X0 + n - 1) / n) + 3) & ~3;
            bool ok = true;
            for (int xa = X0; xa < X1 && ok; xa += sw) {
                int r0, r1, c0, c1;
                warp_src_rect(a, xa, min(X1, xa + sw), y0, y1, H, W, r0, r1, c0, c1);
                if (al16) c0 &= ~15;
                ok = (r1 - r0 + 1) * (((c1 - c0 + 1) * 3 + 15) & ~15) <= smem_cap;
            }
            if (ok) nsl = n;
        }
    }
    const int sw = (TILED && nsl) ? ((((X1 - X0 + nsl - 1) / nsl) + 3) & ~3) : (X1 - X0);
    for (int x0 = X0; x0 < X1; x0 += sw) {
        const int x1 = TILED ? min(X1, x0 + sw) : W, tw = x1 - x0;
        WarpView v{simg, W * 3, 0, 0};
        if (!TILED) {
            if (nsl) {
                int r0, r1, c0, c1;
                warp_src_rect(a, 0, W, y0, y1, H, W, r0, r1, c0, c1);
                const int rows = r1 - r0 + 1;
                const uint8_t* g = simg + (size_t)r0 * W * 3;
                if (pad) {
                    const int pitch = warp_pad_pitch(W), n16 = (W * 3) >> 4;
                    v = WarpView{s_rows, pitch, r0, -WB_PAD};
                    // the source rows are contiguous: chunk i of the rectangle is row i / n16, chunk i % n16 of that row
                    // (m32 = ceil(2^32 / n16) from the host: the quotient is one multiply-high)
                    for (int i = threadIdx.x; i < rows * n16; i += THREADS) {
                        const int r = (int)__umulhi((uint32_t)i, m32), k = i - r * n16;
                        *reinterpret_cast<uint4*>(s_rows + r * pitch + WB_PAD + k * 16) = ld_stream16(g + (size_t)i * 16);
                    }
                    __syncthreads();
                    if (threadIdx.x < rows) {   // columns -2, -1 repeat column 0; W, W+1 repeat W-1 (Pillow clamps the tap index)
                        uint8_t* d = s_rows + threadIdx.x * pitch + WB_PAD;
                        const uint32_t p = *reinterpret_cast<const uint32_t*>(d) & 0xFFFFFFu;              // pixel 0
                        const uint32_t q = *reinterpret_cast<const uint32_t*>(d + W * 3 - 4) >> 8;         // pixel W-1
                        *reinterpret_cast<uint4*>(d - 16) = make_uint4(0u, 0u, p << 16, (p >> 16) | (p << 8));      // .. p p
                        *reinterpret_cast<uint4*>(d + W * 3) = make_uint4(q | (q << 24), q >> 8, 0u, 0u);           // q q ..
                    }
                } else {
                    v = WarpView{s_rows, W * 3, r0, 0};
                    block_load_bytes(s_rows, g, rows * W * 3);
                }
            }
        } else if (nsl) {
            int r0, r1, c0, c1;
            warp_src_rect(a, x0, x1, y0, y1, H, W, r0, r1, c0, c1);
            if (al16) c0 &= ~15;
            const int nb = (c1 - c0 + 1) * 3, pitch = (nb + 15) & ~15;
            v = WarpView{s_rows, pitch, r0, c0 * 3};
            const uint8_t* g = simg + ((size_t)r0 * W + c0) * 3;
            const int rows = r1 - r0 + 1, lane = threadIdx.x & 31;
            if (al16) {
                const int n16 = pitch >> 4;     // may run up to 15 bytes past c1: still inside the row or the next row / image
                const uint8_t* lim = src + (size_t)nsrc * npx * 3;
                for (int r = threadIdx.x >> 5; r < rows; r += THREADS / 32) {
                    const uint8_t* p = g + (size_t)r * W * 3;
                    uint4* d = reinterpret_cast<uint4*>(s_rows + r * pitch);
                    for (int k = lane; k < n16; k += 32) {
                        uint4 q = make_uint4(0, 0, 0, 0);
                        if (p + k * 16 + 16 <= lim) q = ld_stream16(p + k * 16);
                        else for (int bb = 0; bb < 16; ++bb) if (p + k * 16 + bb < lim) reinterpret_cast<uint8_t*>(&q)[bb] = p[k * 16 + bb];
                        d[k] = q;
                    }
                }
            } else {
                for (int r = threadIdx.x >> 5; r < rows; r += THREADS / 32)
                    for (int k = lane; k < nb; k += 32) s_rows[r * pitch + k] = g[(size_t)r * W * 3 + k];
            }
        }
        __syncthreads();
        // push value vidx of the slice on the queue, or evaluate it in place when the queue is full
        auto defer = [&](int vidx, int x, int y, int c, uint8_t* out) {
            const int slot = atomicAdd(&s_qn, 1);
            if (slot < WB_QCAP)
                s_queue[slot] = (uint16_t)vidx;
            else
                *out = bicubic_value64(v.base, v.pitch, v.r0, v.cb, H, W, x, y, c, a, is_persp);
        };

        if (nsl && (TILED || pad) && a[1] == 0.0) {   // (an unpadded rectangle of the small-image instantiation: general path)
            // ---- maps whose xin depends on x only (a1 = 0): the reference's skew (image_augmenter.py:50-58: a zoom + shift,
            // a3 = 0 too) and its vertical shear ([1, 0, 0, k, 1, 0], :82: xin = xc, so dx = 0 and the horizontal cubic is the
            // tap itself).  The horizontal cubic of source row r at column x serves every output row of that column whose
            // 4-row window contains r: one output column per thread walking down the band with a rolling window of the
            // four row values per channel; the arithmetic (and so the fp64 decision) is the generic path's, value by value.
            // Risky values are noted in one bit mask per channel (bit = row of the strip) and queued after the walk: the
            // walk itself stays free of divergent branches.
            const bool rowonly = (a[3] == 0.0);   // yin depends on y only: row coordinates precomputed per band
            const int ncol = min(tw, THREADS);
            const int strips = max(1, THREADS / ncol);
            const int strip = threadIdx.x / ncol, cx = threadIdx.x - strip * ncol;
            const int rows_per = (y1 - y0 + strips - 1) / strips;
            const int ya = y0 + strip * rows_per, yb_end = min(y1, ya + rows_per);
            const uint32_t wy_addr = (uint32_t)__cvta_generic_to_shared(s_wy);
            auto columns = [&](auto rowonly_t, auto pad_t) {
                constexpr bool ROWONLY = decltype(rowonly_t)::value;
                constexpr bool PADV = decltype(pad_t)::value;
                for (int x = x0 + cx; x < x1; x += ncol) {
                    const double xc = (double)x + 0.5;
                    double xin = __dadd_rn(__dmul_rn(a[0], xc), a[2]);
                    const bool xok = !(xin < 0.0 || xin >= (double)W);
                    xin = __dadd_rn(xin, -0.5);
                    const int xf = xok ? (int)floor(xin) : 0;
                    const float fdx = (float)__dadd_rn(xin, -(double)xf);
                    const CubW wx = cubic_weights(fdx);
                    const double tcol = __dmul_rn(a[3], xc);
                    int xo[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) xo[t] = min(max(xf - 1 + t, 0), W - 1) * 3;
                    // padded rows: tap column xf - 1 starts at byte 3 * (xf - 1) + WB_PAD of its row
                    const int boff = 3 * xf - 3 + WB_PAD;
                    const uint8_t* colp = s_rows + (boff & ~3);
                    const uint32_t s8 = (uint32_t)(boff & 3) * 8u;
                    // The walk is driven by SOURCE rows, four per round of the outer loop, so that the slot a row value lands
                    // in (and the order the vertical cubic reads the four slots in) is known at compile time -- no rotation of
                    // the window registers.  r_next = next source row to push (unclamped); an output row with
                    // floor(yin - 0.5) = yf is emitted when the last four pushes were rows yf - 1 .. yf + 2 (r_next = yf + 3);
                    // a row that needs a window which is not the continuation of the current one restarts at yf - 1.
                    float hw[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                    int r_next = -(1 << 30);
                    int y = ya;
                    uint8_t* dcol = dimg + ((size_t)ya * W + x) * 3;
                    // risky rows of this column, per channel: one bit per row of the strip, the LAST row in bit 0 (the masks
                    // shift left once per row: rm + rm + risky is a single add-with-carry)
                    uint32_t rm0 = 0u, rm1 = 0u, rm2 = 0u;
                    int yf = 0;
                    CubW wy{0.f, 0.f, 0.f};
                    bool pend = false;
                    // Vertical shear (a4 = 1, a5 = 0: yin = a3*xc + yc): along a column yin advances by exactly one per row up to
                    // the rounding of the fp64 sum (< 1e-9 for coordinates below 2^20), so floor(yin - 0.5) advances by one and
                    // the fractional part -- all the fp32 weights see -- is the column's, unless it lies within 1e-6 of 0, 0.5
                    // or 1 (where a floor or the inside test 0 <= yin < H could flip between rows: those columns keep the
                    // per-row arithmetic).  yin >= 0 <=> yf >= 0, or yf = -1 with a fraction >= 0.5; yin < H likewise.
                    bool colfast = false;
                    int yf0 = 0, yf_lo = 0, yf_hi = -1;
                    if (!ROWONLY && xok && a[4] == 1.0 && a[5] == 0.0) {
                        const double t0 = __dadd_rn(__dadd_rn(__dadd_rn(tcol, s_trow[ya - y0]), a[5]), -0.5);
                        const double fl0 = floor(t0), fr = __dadd_rn(t0, -fl0), m = fabs(fr - 0.5);
                        if (m > 1e-6 && m < 0.5 - 1e-6 && fabs(t0) < 1048576.0) {
                            colfast = true;
                            yf0 = (int)fl0;
                            wy = cubic_weights((float)fr);
                            yf_lo = fr >= 0.5 ? -1 : 0;
                            yf_hi = fr >= 0.5 ? H - 2 : H - 1;
                        }
                    }
                    // coordinates of the next output row that has a source pixel (rows without one are stored as zeros here)
                    auto fetch = [&]() {
                        pend = false;
                        while (y < yb_end) {
                            if (ROWONLY) {
                                float4 q;   // s_wy[y - y0] through a hoisted shared-window address (the compiler re-derived it per row)
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(wy_addr + 16u * (uint32_t)(y - y0)));
                                wy.a = q.x, wy.b = q.y, wy.c = q.z;
                                yf = __float_as_int(q.w);
                            } else if (colfast) {   // unit row step: the row index advances by one, the fractional part stays
                                yf = yf0 + (y - ya);
                                if (yf < yf_lo || yf > yf_hi) yf = INT_MIN;
                            } else {   // any other map of this family: per-row fp64 (out of line -- it is the rare case here and
                                       // five inlined copies of it pushed the walk out of the instruction cache)
                                const float4 q = warp_row_coords64(tcol, s_trow[y - y0], a[5], H);
                                wy.a = q.x, wy.b = q.y, wy.c = q.z;
                                yf = __float_as_int(q.w);
                            }
                            if (xok && yf != INT_MIN) {
                                pend = true;
                                break;
                            }
                            dcol[0] = 0, dcol[1] = 0, dcol[2] = 0;
                            rm0 += rm0, rm1 += rm1, rm2 += rm2;
                            ++y, dcol += W * 3;
                        }
                    };
                    fetch();
                    while (pend) {
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            // window (oldest .. newest) = slots p, p+1, p+2, p+3 (mod 4)
                            while (pend && yf + 3 == r_next) {
                                bool rk;
                                dcol[0] = warp_trunc(cubw(hw[0][p], hw[0][(p + 1) & 3], hw[0][(p + 2) & 3], hw[0][(p + 3) & 3], wy), rk);
                                rm0 = rm0 + rm0 + (rk ? 1u : 0u);
                                dcol[1] = warp_trunc(cubw(hw[1][p], hw[1][(p + 1) & 3], hw[1][(p + 2) & 3], hw[1][(p + 3) & 3], wy), rk);
                                rm1 = rm1 + rm1 + (rk ? 1u : 0u);
                                dcol[2] = warp_trunc(cubw(hw[2][p], hw[2][(p + 1) & 3], hw[2][(p + 2) & 3], hw[2][(p + 3) & 3], wy), rk);
                                rm2 = rm2 + rm2 + (rk ? 1u : 0u);
                                ++y, dcol += W * 3;
                                fetch();
                            }
                            if (!pend) break;
                            if (yf + 3 < r_next || yf - 1 > r_next) r_next = yf - 1;
                            const int roff = warp_row_off(v, r_next, H);   // (row - r0) * pitch - cb, row clamped into the image
                            ++r_next;
                            if (PADV) {
                                float t[12];
                                taps12_biased(colp + (roff - WB_PAD), s8, t);   // cb = -WB_PAD is already inside boff
#pragma unroll
                                for (int c = 0; c < 3; ++c)   // dx = 0: the cubic is (v2 - 2^23) + 0 exactly -- the tap itself
                                    hw[c][p] = (!ROWONLY && fdx == 0.f) ? t[3 + c] - 8388608.f : cubw_b(t[c], t[3 + c], t[6 + c], t[9 + c], wx);
                            } else {
                                const uint8_t* rp = s_rows + roff;   // shared loads, not generic
#pragma unroll
                                for (int c = 0; c < 3; ++c)
                                    hw[c][p] = (!ROWONLY && fdx == 0.f) ? biased(rp[xo[1] + c]) - 8388608.f
                                                                        : cubw_b(biased(rp[xo[0] + c]), biased(rp[xo[1] + c]), biased(rp[xo[2] + c]),
                                                                                 biased(rp[xo[3] + c]), wx);
                            }
                        }
                    }
                    const int nrows = yb_end - ya;
                    auto flush = [&](uint32_t m, int c) {
                        while (m) {
                            const int yr = ya + nrows - __ffs(m);   // bit b = row ya + nrows - 1 - b
                            m &= m - 1u;
                            defer(((yr - y0) * tw + (x - x0)) * 3 + c, x, yr, c, dimg + ((size_t)yr * W + x) * 3 + c);
                        }
                    };
                    flush(rm0, 0);
                    flush(rm1, 1);
                    flush(rm2, 2);
                }
            };
            if (strip < strips) {   // one layout per instantiation (padded <=> !TILED): half the code
                if (rowonly)
                    columns(std::true_type{}, std::integral_constant<bool, !TILED>{});
                else
                    columns(std::false_type{}, std::integral_constant<bool, !TILED>{});
            }
        } else if (nsl && (TILED || pad) && (tw & 3) == 0 && a[3] == 0.0 && a[4] == 1.0 && a[5] == 0.0) {
            // ---- identity-y maps (the reference's horizontal shear [1, k, 0, 0, 1, 0], image_augmenter.py:82): yin = yc exactly,
            // so yf = y, dy = 0 and cubic(.., 0) = v2: output row y is a 4-tap filter of source row y.  Same thread layout as
            // the general path (4 consecutive pixels of a row per thread) with the row terms hoisted and no row arithmetic.
            // Risky values are noted in a bit mask (12 bits per round of the loop) and queued after the loop.
            const int band_px = (y1 - y0) * tw;
            const int twsh = ((tw & (tw - 1)) == 0) ? 31 - __clz(tw) : -1;
            auto rows4 = [&](auto pad_t) {
                constexpr bool PADV = decltype(pad_t)::value;
                unsigned long long rlo = 0ull, rhi = 0ull;   // rounds 0-4 / 5-9 of this thread, 12 bits each
                int it = 0;
                for (int q = threadIdx.x * 4; q < band_px; q += THREADS * 4, ++it) {
                    const int qy = twsh >= 0 ? (q >> twsh) : q / tw, qx = q - qy * tw;
                    const int y = y0 + qy;
                    const uint8_t* rp = s_rows + (y - v.r0) * v.pitch - v.cb;
                    const double trow = __dmul_rn(a[1], (double)y + 0.5);
                    uint8_t out[12];
                    uint32_t m12 = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int x = x0 + qx + k;
                        double xin = __dadd_rn(__dadd_rn(__dmul_rn(a[0], (double)x + 0.5), trow), a[2]);
                        uint8_t r3[3] = {0, 0, 0};
                        if (!(xin < 0.0 || xin >= (double)W)) {
                            xin = __dadd_rn(xin, -0.5);
                            const int xf = (int)floor(xin);
                            const CubW wx = cubic_weights((float)__dadd_rn(xin, -(double)xf));
                            float t[12];
                            if (PADV) {
                                const int boff = 3 * xf - 3;   // rp already points at column 0 of the padded row
                                taps12_biased(rp + (boff & ~3), (uint32_t)(boff & 3) * 8u, t);
                            } else {
                                const int xa = min(max(xf - 1, 0), W - 1) * 3, xb = min(max(xf, 0), W - 1) * 3;
                                const int xc2 = min(max(xf + 1, 0), W - 1) * 3, xd = min(max(xf + 2, 0), W - 1) * 3;
#pragma unroll
                                for (int c = 0; c < 3; ++c)
                                    t[c] = biased(rp[xa + c]), t[3 + c] = biased(rp[xb + c]), t[6 + c] = biased(rp[xc2 + c]), t[9 + c] = biased(rp[xd + c]);
                            }
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                bool rk;
                                r3[c] = warp_trunc(cubw_b(t[c], t[3 + c], t[6 + c], t[9 + c], wx), rk);
                                if (rk) m12 |= 1u << (k * 3 + c);
                            }
                        }
                        out[k * 3] = r3[0], out[k * 3 + 1] = r3[1], out[k * 3 + 2] = r3[2];
                    }
                    uint8_t* d = dimg + ((size_t)y * W + x0 + qx) * 3;
                    if ((reinterpret_cast<uintptr_t>(d) & 3) == 0) {
                        uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
                        d32[0] = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
                        d32[1] = out[4] | (out[5] << 8) | (out[6] << 16) | ((uint32_t)out[7] << 24);
                        d32[2] = out[8] | (out[9] << 8) | (out[10] << 16) | ((uint32_t)out[11] << 24);
                    } else {
                        for (int i = 0; i < 12; ++i) d[i] = out[i];
                    }
                    if (it < 5)
                        rlo |= (unsigned long long)m12 << (12 * it);
                    else if (it < 10)
                        rhi |= (unsigned long long)m12 << (12 * (it - 5));
                    else   // (bands of more than 10 rounds do not occur with WB_ROWS x WB_COLS tiles; kept for safety)
                        for (; m12; m12 &= m12 - 1u) {
                            const int bit = __ffs(m12) - 1, k = bit / 3, c = bit - 3 * k;
                            defer((q + k) * 3 + c, x0 + qx + k, y, c, d + k * 3 + c);
                        }
                }
                auto flush = [&](unsigned long long m, int it0) {
                    while (m) {
                        const int bit = __ffsll((long long)m) - 1;
                        m &= m - 1ull;
                        const int r = bit / 12, kc = bit - 12 * r, k = kc / 3, c = kc - 3 * k;
                        const int q = threadIdx.x * 4 + (it0 + r) * THREADS * 4;
                        const int qy = twsh >= 0 ? (q >> twsh) : q / tw, qx = q - qy * tw;
                        const int x = x0 + qx + k, y = y0 + qy;
                        defer((q + k) * 3 + c, x, y, c, dimg + ((size_t)y * W + x) * 3 + c);
                    }
                };
                flush(rlo, 0);
                flush(rhi, 5);
            };
            rows4(std::integral_constant<bool, !TILED>{});
        } else {
            const int band_px = (y1 - y0) * tw;
            for (int q = threadIdx.x * 4; q < band_px; q += THREADS * 4) {
                uint8_t out[12];
                const int qy = q / tw, qx = q - qy * tw;
                const bool one_row = (qx + 4 <= tw);   // the four pixels are contiguous in memory
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    uint8_t res[3] = {0, 0, 0};
                    if (q + k < band_px) {
                        int yy = qy, xx = qx + k;
                        while (xx >= tw) {
                            xx -= tw;
                            ++yy;
                        }
                        const int x = x0 + xx, y = y0 + yy;
                        const uint32_t risky = nsl ? bicubic_pixel(WarpView{s_rows, v.pitch, v.r0, v.cb}, H, W, x, y, a, is_persp, res)
                                                   : bicubic_pixel(WarpView{simg, W * 3, 0, 0}, H, W, x, y, a, is_persp, res);
                        if (risky) {
                            if (risky & 1u) defer((q + k) * 3 + 0, x, y, 0, &res[0]);
                            if (risky & 2u) defer((q + k) * 3 + 1, x, y, 1, &res[1]);
                            if (risky & 4u) defer((q + k) * 3 + 2, x, y, 2, &res[2]);
                        }
                        if (!one_row) {
                            uint8_t* d1 = dimg + ((size_t)y * W + x) * 3;
                            d1[0] = res[0], d1[1] = res[1], d1[2] = res[2];
                        }
                    }
                    out[k * 3] = res[0];
                    out[k * 3 + 1] = res[1];
                    out[k * 3 + 2] = res[2];
                }
                if (one_row) {
                    uint8_t* d = dimg + ((size_t)(y0 + qy) * W + x0 + qx) * 3;
                    if ((reinterpret_cast<uintptr_t>(d) & 3) == 0) {
                        uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
                        d32[0] = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
                        d32[1] = out[4] | (out[5] << 8) | (out[6] << 16) | ((uint32_t)out[7] << 24);
                        d32[2] = out[8] | (out[9] << 8) | (out[10] << 16) | ((uint32_t)out[11] << 24);
                    } else {
                        for (int i = 0; i < 12; ++i) d[i] = out[i];
                    }
                }
            }
        }
        // ---- deferred fp64 values: one per thread; the byte overwrites the fp32 guess stored above (same block, after
        // the barrier, so the two stores to that address are ordered)
        __syncthreads();
        const int nq = min(s_qn, WB_QCAP);
        for (int i = threadIdx.x; i < nq; i += THREADS) {
            const int vidx = s_queue[i];
            const int pix = vidx / 3, c = vidx - pix * 3;
            const int yy = pix / tw, xx = pix - yy * tw;
            dimg[((size_t)(y0 + yy) * W + x0 + xx) * 3 + c] =
                bicubic_value64(v.base, v.pitch, v.r0, v.cb, H, W, x0 + xx, y0 + yy, c, a, is_persp);
        }
        __syncthreads();   // the queue and the staged rectangle are reused by the next slice
        if (threadIdx.x == 0) s_qn = 0;
    }
}

// ------------------------------------------------------------------------------ crop + Lanczos
// grid (column strips, B).  A block owns LZ_TW output columns: horizontal pass over every crop
// row into a shared uint8 strip (Pillow materialises the uint8 intermediate), then the vertical
// pass writes the [OH, LZ_TW] output strip (and its /255 float32 twin when requested).
constexpr int LZ_TW = 16;

__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)min(255, max(0, v)); }

__global__ void __launch_bounds__(THREADS) k_crop_lanczos(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          float* __restrict__ dstf, int H, int W,
                                                          const int32_t* __restrict__ box, int OH, int OW,
                                                          const int32_t* __restrict__ tb, const int32_t* __restrict__ tk,
                                                          int kstride, const int32_t* __restrict__ toff,
                                                          const int32_t* __restrict__ sidx) {
    extern __shared__ __align__(16) uint8_t s_mid[];  // [crop_h][LZ_TW][3]
    const int img = blockIdx.y;
    const int c0 = blockIdx.x * LZ_TW;
    const int ncol = min(LZ_TW, OW - c0);
    const int left = box[img * 4], top = box[img * 4 + 1], cw = box[img * 4 + 2], ch = box[img * 4 + 3];
    const int32_t* xb = tb + (size_t)toff[img * 4 + 0] * 2;
    const int32_t* xk = tk + (size_t)toff[img * 4 + 0] * kstride;
    const int32_t* yb = tb + (size_t)toff[img * 4 + 2] * 2;
    const int32_t* yk = tk + (size_t)toff[img * 4 + 2] * kstride;
    const uint8_t* simg = src + (size_t)(sidx ? sidx[img] : img) * H * W * 3;
    const bool need_h = (cw != OW);
    const bool need_v = (ch != OH);
    // horizontal pass (or plain copy of the crop when widths match -- Pillow skips the pass)
    for (int i = threadIdx.x; i < ch * ncol; i += THREADS) {
        const int y = i / ncol, c = i - y * ncol;
        const int oc = c0 + c;
        const uint8_t* row = simg + ((size_t)(top + y) * W + left) * 3;
        uint8_t* o = s_mid + ((size_t)y * LZ_TW + c) * 3;
        if (need_h) {
            const int xmin = xb[oc * 2], cnt = xb[oc * 2 + 1];
            const int32_t* k = xk + (size_t)oc * kstride;
            int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
            const uint8_t* px = row + (size_t)xmin * 3;
            for (int t = 0; t < cnt; ++t) {
                const int kv = __ldg(k + t);
                s0 += __ldg(px + t * 3) * kv;
                s1 += __ldg(px + t * 3 + 1) * kv;
                s2 += __ldg(px + t * 3 + 2) * kv;
            }
            o[0] = clip8(s0 >> 22);
            o[1] = clip8(s1 >> 22);
            o[2] = clip8(s2 >> 22);
        } else {
            o[0] = __ldg(row + oc * 3);
            o[1] = __ldg(row + oc * 3 + 1);
            o[2] = __ldg(row + oc * 3 + 2);
        }
    }
    __syncthreads();
    uint8_t* dimg = dst + (size_t)img * OH * OW * 3;
    float* fimg = dstf ? dstf + (size_t)img * OH * OW * 3 : nullptr;
    const int rowb = ncol * 3;
    for (int i = threadIdx.x; i < OH * rowb; i += THREADS) {
        const int oy = i / rowb, o = i - oy * rowb;  // o = c*3 + channel inside the strip
        uint8_t v;
        if (need_v) {
            const int ymin = yb[oy * 2], cnt = yb[oy * 2 + 1];
            const int32_t* k = yk + (size_t)oy * kstride;
            int s = 1 << 21;
            for (int t = 0; t < cnt; ++t) s += s_mid[(size_t)(ymin + t) * LZ_TW * 3 + o] * __ldg(k + t);
            v = clip8(s >> 22);
        } else {
            v = s_mid[(size_t)oy * LZ_TW * 3 + o];
        }
        const size_t di = ((size_t)oy * OW + c0) * 3 + o;
        dimg[di] = v;
        if (fimg) fimg[di] = (float)v / 255.0f;
    }
}


// ---- strip version: a block owns LZ_TO output rows of one image at FULL width.  Horizontal pass for the
// crop rows the strip needs (one output column per thread walking down the rows, its <= 8 taps in
// registers) -> u8 intermediate in shared memory (Pillow's ImagingResampleHorizontal_8bpc result), vertical
// pass on 4 output bytes per thread (32-bit shared loads, row coefficients broadcast from shared memory),
// 32-bit / 128-bit coalesced stores.  Needs OW % 4 == 0.
constexpr int LZ_TO = 32;

template <int KMAX>
__global__ void __launch_bounds__(THREADS) k_crop_lanczos_strip(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                float* __restrict__ dstf, int H, int W,
                                                                const int32_t* __restrict__ box, int OH, int OW,
                                                                const int32_t* __restrict__ tb, const int32_t* __restrict__ tk,
                                                                int kstride, const int32_t* __restrict__ toff, int mrows_cap, int lz_to,
                                                                const int32_t* __restrict__ sidx, int nsrc) {
    extern __shared__ __align__(16) uint8_t sm_lz[];
    __shared__ float s_f255[256];   // v / 255.0f (normalize_array, image_utils.py:126-130): correctly rounded division, tabulated
    if (dstf)
        for (int i = threadIdx.x; i < 256; i += THREADS) s_f255[i] = (float)i / 255.0f;
    const int OWB = OW * 3;
    uint8_t* s_mid = sm_lz;                                                               // [mrows_cap][OWB]
    int32_t* s_yk = reinterpret_cast<int32_t*>(sm_lz + (((size_t)mrows_cap * OWB + 15) & ~(size_t)15));  // [LZ_TO][kstride]
    int32_t* s_yb = s_yk + lz_to * kstride;                                               // [lz_to][2]
    int32_t* s_mis = s_yb + lz_to * 2;                                                    // [mrows_cap]
    const int src_pitch = ((W * 3 + 15) & ~15) + 32;                                      // staged crop row pitch
    uint8_t* s_src = reinterpret_cast<uint8_t*>(s_mis + ((mrows_cap + 3) & ~3));          // [mrows_cap][src_pitch]
    const int img = blockIdx.y;
    const int o0 = blockIdx.x * lz_to;   // lz_to output rows per block (LZ_TO, halved for wide images until the strip fits)
    const int nrow = min(lz_to, OH - o0);
    const int left = box[img * 4], top = box[img * 4 + 1], cw = box[img * 4 + 2], ch = box[img * 4 + 3];
    const int32_t* xb = tb + (size_t)toff[img * 4 + 0] * 2;
    const int32_t* xk = tk + (size_t)toff[img * 4 + 0] * kstride;
    const int32_t* yb = tb + (size_t)toff[img * 4 + 2] * 2;
    const int32_t* yk = tk + (size_t)toff[img * 4 + 2] * kstride;
    const uint8_t* simg = src + (size_t)(sidx ? sidx[img] : img) * H * W * 3;
    const bool need_h = (cw != OW), need_v = (ch != OH);
    // crop rows this strip reads
    int m0, m1;
    if (need_v) {
        m0 = yb[o0 * 2];
        m1 = yb[(o0 + nrow - 1) * 2] + yb[(o0 + nrow - 1) * 2 + 1];
        for (int i = threadIdx.x; i < nrow * kstride; i += THREADS) s_yk[i] = yk[(size_t)o0 * kstride + i];
        for (int i = threadIdx.x; i < nrow * 2; i += THREADS) s_yb[i] = yb[o0 * 2 + i];
    } else {
        m0 = o0;
        m1 = o0 + nrow;
    }
    const int mrows = m1 - m0;   // <= mrows_cap by construction of the launch
    // ---- stage the crop rows (columns left .. left+cw-1) in shared memory: 16-byte loads from the aligned span
    const int cwb = cw * 3;
    const size_t g0 = ((size_t)(top + m0) * W + left) * 3;
    for (int r = threadIdx.x >> 5; r < mrows; r += THREADS / 32) {
        const uint8_t* g = simg + g0 + (size_t)r * W * 3;
        uint8_t* d = s_src + (size_t)r * src_pitch;
        const int mis = (int)(reinterpret_cast<uintptr_t>(g) & 15);       // d[mis + i] = g[i]: both sides 16-byte aligned
        const uint4* g16 = reinterpret_cast<const uint4*>(g - mis);
        const int n16 = (mis + cwb + 15) >> 4;
        const uint8_t* img_end = src + (size_t)nsrc * H * W * 3;
        for (int i = threadIdx.x & 31; i < n16; i += 32) {
            if (reinterpret_cast<const uint8_t*>(g16 + i + 1) <= img_end)
                reinterpret_cast<uint4*>(d)[i] = ld_stream16(g16 + i);
            else
                for (int b = 0; b < 16; ++b) {
                    const uint8_t* q = reinterpret_cast<const uint8_t*>(g16 + i) + b;
                    d[i * 16 + b] = q < img_end ? *q : 0;
                }
        }
        if ((threadIdx.x & 31) == 0) s_mis[r] = mis;
    }
    __syncthreads();
    // ---- horizontal pass
    for (int oc = threadIdx.x; oc < OW; oc += THREADS) {
        uint8_t* o = s_mid + oc * 3;
        if (need_h) {
            const int xmin = xb[oc * 2], cnt = xb[oc * 2 + 1];
            int kreg[KMAX];
#pragma unroll
            for (int t = 0; t < KMAX; ++t) kreg[t] = (t < cnt) ? xk[(size_t)oc * kstride + t] : 0;
            // taps beyond cnt have weight 0; clamp their offset to the last valid tap so nothing outside the staged span is read
            int toff3[KMAX];
#pragma unroll
            for (int t = 0; t < KMAX; ++t) toff3[t] = (xmin + min(t, cnt - 1)) * 3;
            for (int r = 0; r < mrows; ++r, o += OWB) {
                const uint8_t* px = s_src + (size_t)r * src_pitch + s_mis[r];
                int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
#pragma unroll
                for (int t = 0; t < KMAX; ++t) {
                    s0 += px[toff3[t]] * kreg[t];
                    s1 += px[toff3[t] + 1] * kreg[t];
                    s2 += px[toff3[t] + 2] * kreg[t];
                }
                o[0] = clip8(s0 >> 22);
                o[1] = clip8(s1 >> 22);
                o[2] = clip8(s2 >> 22);
            }
        } else {
            for (int r = 0; r < mrows; ++r, o += OWB) {
                const uint8_t* px = s_src + (size_t)r * src_pitch + s_mis[r] + oc * 3;
                o[0] = px[0];
                o[1] = px[1];
                o[2] = px[2];
            }
        }
    }
    __syncthreads();
    // ---- vertical pass: 4 bytes per thread
    const int wpr = OWB >> 2;
    uint8_t* dimg = dst + ((size_t)img * OH + o0) * OWB;
    float* fimg = dstf ? dstf + ((size_t)img * OH + o0) * OWB : nullptr;
    const uint32_t* mid32 = reinterpret_cast<const uint32_t*>(s_mid);
    int r = 0, j4 = threadIdx.x;
    for (int i = threadIdx.x; i < nrow * wpr; i += THREADS, j4 += THREADS) {
        while (j4 >= wpr) {
            j4 -= wpr;
            ++r;
        }
        uint32_t outw;
        if (need_v) {
            const int ymin = s_yb[r * 2] - m0, cnt = s_yb[r * 2 + 1];
            const int32_t* k = s_yk + r * kstride;
            int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21, s3 = 1 << 21;
            const uint32_t* p = mid32 + (size_t)ymin * wpr + j4;
#pragma unroll
            for (int t = 0; t < KMAX; ++t) {
                if (t < cnt) {
                    const uint32_t w = p[t * wpr];
                    const int kv = k[t];
                    s0 += (int)(w & 0xFFu) * kv;
                    s1 += (int)((w >> 8) & 0xFFu) * kv;
                    s2 += (int)((w >> 16) & 0xFFu) * kv;
                    s3 += (int)(w >> 24) * kv;
                }
            }
            outw = (uint32_t)clip8(s0 >> 22) | ((uint32_t)clip8(s1 >> 22) << 8) | ((uint32_t)clip8(s2 >> 22) << 16) |
                   ((uint32_t)clip8(s3 >> 22) << 24);
        } else {
            outw = mid32[(size_t)r * wpr + j4];
        }
        reinterpret_cast<uint32_t*>(dimg)[(size_t)r * wpr + j4] = outw;
        if (fimg) {
            float4 f;
            f.x = s_f255[outw & 0xFFu];
            f.y = s_f255[(outw >> 8) & 0xFFu];
            f.z = s_f255[(outw >> 16) & 0xFFu];
            f.w = s_f255[outw >> 24];
            __stcs(reinterpret_cast<float4*>(fimg) + (size_t)r * wpr + j4, f);
        }
    }
}

// ---- dp4a version (the default): the same strip decomposition, both passes on the integer dot-product pipe.
// A 22-bit coefficient splits exactly as k = k0 + 256 k1 + 65536 k2 (k0, k1 unsigned bytes, k2 a signed byte: k <= 2^22),
// so sum(px * k) = dp4a(px, k0) + 256 dp4a(px, k1) + 65536 dp4a(px, k2) over four taps per instruction (IDP.4A, exact in
// int32: |sum| < 2^31).  Four consecutive taps of one channel must share a 32-bit word, so
//   (1) the crop rows are staged raw (16-byte loads), then de-interleaved in shared memory into R / G / B planes;
//   (2) horizontal pass: one output column per thread walking down the rows; the <= 4*NG taps of the column are an
//       unaligned byte window of the plane row = NG+1 aligned words + NG funnel shifts; coefficient words in registers;
//       four consecutive rows of a column are packed into one word of the intermediate M4[row quad][byte column];
//   (3) vertical pass: a thread owns four adjacent byte columns (one 128-bit load per row quad), the window over the rows
//       is again NG+1 words + NG funnel shifts (shift and coefficients are per output row: broadcast loads), the four
//       results leave as one 32-bit store (plus a float4 for the /255 output).
// Identity passes (Pillow skips a pass whose size does not change) are the coefficient 2^22 on one tap: exact.
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {   // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr int LZ4_TO_MAX = 64;   // output rows per strip: chosen by the host (32 by default; LFX_LZ_TO for experiments)
constexpr int LZ4_T = 512;   // threads: 16 warps per block, two blocks per SM

// clamp(v0..v3, 0, 255) packed into one word (v0 = byte 0): two saturating pack instructions (I2IP)
__device__ __forceinline__ uint32_t pack_sat_u8x4(int v0, int v1, int v2, int v3) {
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(hi));
    return d;
}

template <int NG>
__global__ void __launch_bounds__(LZ4_T, 2) k_lanczos_dp4a(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           float* __restrict__ dstf, int H, int W,
                                                           const int32_t* __restrict__ box, int OH, int OW,
                                                           const int32_t* __restrict__ tb, const int32_t* __restrict__ tk,
                                                           int kstride, const int32_t* __restrict__ toff, int mrows_cap,
                                                           const int32_t* __restrict__ sidx, int nsrc, int LZ4_TO) {
    extern __shared__ __align__(16) uint8_t sm_lz[];
    __shared__ float s_f255[256];   // v / 255.0f (normalize_array, image_utils.py:126-130): correctly rounded, tabulated
    constexpr int VKS = (3 * NG + 3) & ~3;                 // coefficient words per output row (padded to 16 bytes)
    constexpr int NWARP = LZ4_T / 32;
    const int OWB = OW * 3;
    const int raw_pitch = ((W * 3 + 15) & ~15) + 32;
    const int mrows4 = (mrows_cap + 3) & ~3;
    const int nquad = mrows4 / 4 + NG + 1;                 // row quads of M4 (the window of the last row may run past the strip)
    const int ppitch = ((W + 3) & ~3) + 4 * NG + 4;        // plane row pitch in bytes (taps past the crop read padding)
    const size_t raw_bytes = (size_t)mrows_cap * raw_pitch, m4_bytes = (size_t)nquad * OWB * 4;
    const size_t uni = ((raw_bytes > m4_bytes ? raw_bytes : m4_bytes) + 15) & ~(size_t)15;
    uint8_t* s_raw = sm_lz;                                // [mrows_cap][raw_pitch]        (dead after the de-interleave)
    uint32_t* s_m4 = reinterpret_cast<uint32_t*>(sm_lz);   // [nquad][OWB] words            (aliases s_raw)
    uint8_t* s_pl = sm_lz + uni;                           // [3][mrows4][ppitch]
    uint32_t* s_vk = reinterpret_cast<uint32_t*>(s_pl + (((size_t)3 * mrows4 * ppitch + 15) & ~(size_t)15));  // [LZ4_TO][VKS]
    int32_t* s_vb = reinterpret_cast<int32_t*>(s_vk + LZ4_TO * VKS);                                          // [LZ4_TO]
    int32_t* s_mis = s_vb + LZ4_TO;                                                                           // [mrows_cap]

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (dstf)
        for (int i = threadIdx.x; i < 256; i += LZ4_T) s_f255[i] = (float)i / 255.0f;
    const int img = blockIdx.y;
    const int o0 = blockIdx.x * LZ4_TO;
    const int nrow = min(LZ4_TO, OH - o0);
    const int left = box[img * 4], top = box[img * 4 + 1], cw = box[img * 4 + 2], ch = box[img * 4 + 3];
    const int32_t* xb = tb + (size_t)toff[img * 4 + 0] * 2;
    const int32_t* xk = tk + (size_t)toff[img * 4 + 0] * kstride;
    const int32_t* yb = tb + (size_t)toff[img * 4 + 2] * 2;
    const int32_t* yk = tk + (size_t)toff[img * 4 + 2] * kstride;
    const uint8_t* simg = src + (size_t)(sidx ? sidx[img] : img) * H * W * 3;
    const bool need_h = (cw != OW), need_v = (ch != OH);
    const int ktaps = min(4 * NG, kstride);

    // coefficient words of one table row: taps 4g..4g+3 of slice s in word [s * NG + g]
    auto split = [&](const int32_t* k, int cnt, uint32_t* out) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int tt = 4 * g + t;
                const int kv = (tt < ktaps && tt < cnt) ? k[tt] : 0;
                w0 |= (uint32_t)(kv & 255) << (8 * t);
                w1 |= (uint32_t)((kv >> 8) & 255) << (8 * t);
                w2 |= (uint32_t)((kv >> 16) & 255) << (8 * t);
            }
            out[g] = w0; out[NG + g] = w1; out[2 * NG + g] = w2;
        }
    };

    // ---- crop rows this strip reads + the vertical coefficient words of its output rows
    int m0, m1;
    if (need_v) {
        m0 = yb[o0 * 2];
        m1 = yb[(o0 + nrow - 1) * 2] + yb[(o0 + nrow - 1) * 2 + 1];
    } else {
        m0 = o0;
        m1 = o0 + nrow;
    }
    const int mrows = m1 - m0;   // <= mrows_cap by construction of the launch
    // ---- (1a) stage the crop rows raw: 16-byte loads from the aligned span
    const int cwb = cw * 3;
    const size_t g0 = ((size_t)(top + m0) * W + left) * 3;
    const uint8_t* img_end = src + (size_t)nsrc * H * W * 3;
    for (int r = wid; r < mrows; r += NWARP) {
        const uint8_t* g = simg + g0 + (size_t)r * W * 3;
        uint8_t* d = s_raw + (size_t)r * raw_pitch;
        const int mis = (int)(reinterpret_cast<uintptr_t>(g) & 15);       // d[mis + i] = g[i]: both sides 16-byte aligned
        const uint4* g16 = reinterpret_cast<const uint4*>(g - mis);
        const int n16 = (mis + cwb + 15) >> 4;
        for (int i = lane; i < n16; i += 32) {
            if (reinterpret_cast<const uint8_t*>(g16 + i + 1) <= img_end)
                reinterpret_cast<uint4*>(d)[i] = ld_stream16(g16 + i);
            else
                for (int b = 0; b < 16; ++b) {
                    const uint8_t* q = reinterpret_cast<const uint8_t*>(g16 + i) + b;
                    d[i * 16 + b] = q < img_end ? *q : 0;
                }
        }
        if (lane == 0) s_mis[r] = mis;
    }
    if (threadIdx.x < nrow) {   // (after the loads are in flight)
        const int r = threadIdx.x;
        uint32_t kw[3 * NG];
        if (need_v) {
            split(yk + (size_t)(o0 + r) * kstride, yb[(o0 + r) * 2 + 1], kw);
            s_vb[r] = yb[(o0 + r) * 2] - m0;
        } else {
            const int32_t one = 1 << 22;
            split(&one, 1, kw);
            s_vb[r] = r;
        }
#pragma unroll
        for (int i = 0; i < 3 * NG; ++i) s_vk[r * VKS + i] = kw[i];
    }
    __syncthreads();
    // ---- (1b) de-interleave: 4 pixels (12 bytes at an arbitrary byte offset) -> one word of each plane; a warp per row
    {
        const int ng4 = (cw + 3) >> 2;
        const int plane = mrows4 * ppitch;
        for (int r = wid; r < mrows; r += NWARP) {
            const int mis = s_mis[r];
            const uint8_t* rawr = s_raw + (size_t)r * raw_pitch;
            uint8_t* dr = s_pl + (size_t)r * ppitch;
            for (int g = lane; g < ng4; g += 32) {
                const int o = mis + 12 * g;
                const uint32_t* q = reinterpret_cast<const uint32_t*>(rawr) + (o >> 2);
                const int sh = (o & 3) * 8;
                const uint32_t q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
                const uint32_t w0 = __funnelshift_r(q0, q1, sh), w1 = __funnelshift_r(q1, q2, sh), w2 = __funnelshift_r(q2, q3, sh);
                // w0 = R0 G0 B0 R1, w1 = G1 B1 R2 G2, w2 = B2 R3 G3 B3
                uint8_t* d = dr + 4 * g;
                *reinterpret_cast<uint32_t*>(d) = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
                *reinterpret_cast<uint32_t*>(d + plane) = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
                *reinterpret_cast<uint32_t*>(d + 2 * plane) = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
            }
        }
    }
    __syncthreads();
    // ---- (2) horizontal pass -> M4.  Thread = (output column, share of the row quads)
    {
        const int nq = (mrows + 3) >> 2;
        const int ncolthreads = min(OW, LZ4_T);
        const int nshare = max(1, LZ4_T / ncolthreads);            // threads per column (row quads split between them)
        const int share = threadIdx.x / ncolthreads;
        const int per = (nq + nshare - 1) / nshare;
        const int rq0 = share * per, rq1 = min(nq, rq0 + per);
        for (int oc = threadIdx.x - share * ncolthreads; oc < OW && share < nshare && rq0 < rq1; oc += ncolthreads) {
            uint32_t kw[3 * NG];
            int xmin;
            if (need_h) {
                xmin = xb[oc * 2];
                split(xk + (size_t)oc * kstride, xb[oc * 2 + 1], kw);
            } else {
                const int32_t one = 1 << 22;
                xmin = oc;
                split(&one, 1, kw);
            }
            const int sh = (xmin & 3) * 8;
            const int plane = mrows4 * ppitch;
            const uint8_t* prow = s_pl + (xmin & ~3) + (size_t)rq0 * 4 * ppitch;
            uint32_t* mo = s_m4 + (size_t)rq0 * OWB + oc * 3;
            for (int rq = rq0; rq < rq1; ++rq, mo += OWB, prow += 4 * ppitch) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t* p = reinterpret_cast<const uint32_t*>(prow + c * plane + i * ppitch);
                        uint32_t w[NG + 1];
#pragma unroll
                        for (int g = 0; g <= NG; ++g) w[g] = p[g];
                        int a0 = 1 << 21, a1 = 0, a2 = 0;
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            const uint32_t a = __funnelshift_r(w[g], w[g + 1], sh);
                            a0 = (int)__dp4a(a, kw[g], (uint32_t)a0);
                            a1 = (int)__dp4a(a, kw[NG + g], (uint32_t)a1);
                            a2 = dp4a_us(a, kw[2 * NG + g], a2);
                        }
                        v[i] = (a2 * 65536 + (a1 * 256 + a0)) >> 22;
                    }
                    mo[c] = pack_sat_u8x4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
    __syncthreads();
    // ---- (3) vertical pass: a warp per output row, 4 byte columns per lane and step
    {
        const int G = OWB >> 2;   // column groups (words) per row
        uint8_t* dimg = dst + ((size_t)img * OH + o0) * OWB;
        float* fimg = dstf ? dstf + ((size_t)img * OH + o0) * OWB : nullptr;
        for (int r = wid; r < nrow; r += NWARP) {
            const int b = s_vb[r];
            const int sh = (b & 3) * 8;
            const uint4* kq = reinterpret_cast<const uint4*>(s_vk + r * VKS);
            uint32_t kw[VKS];
#pragma unroll
            for (int i = 0; i < VKS / 4; ++i) {
                const uint4 q = kq[i];
                kw[4 * i] = q.x; kw[4 * i + 1] = q.y; kw[4 * i + 2] = q.z; kw[4 * i + 3] = q.w;
            }
            const uint4* mrow = reinterpret_cast<const uint4*>(s_m4 + (size_t)(b >> 2) * OWB);
            uint32_t* drow = reinterpret_cast<uint32_t*>(dimg) + (size_t)r * G;
            for (int cg = lane; cg < G; cg += 32) {
                uint4 w[NG + 1];
#pragma unroll
                for (int g = 0; g <= NG; ++g) w[g] = mrow[(size_t)g * G + cg];
                int v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int a0 = 1 << 21, a1 = 0, a2 = 0;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const uint32_t lo = j == 0 ? w[g].x : j == 1 ? w[g].y : j == 2 ? w[g].z : w[g].w;
                        const uint32_t hi = j == 0 ? w[g + 1].x : j == 1 ? w[g + 1].y : j == 2 ? w[g + 1].z : w[g + 1].w;
                        const uint32_t a = __funnelshift_r(lo, hi, sh);
                        a0 = (int)__dp4a(a, kw[g], (uint32_t)a0);
                        a1 = (int)__dp4a(a, kw[NG + g], (uint32_t)a1);
                        a2 = dp4a_us(a, kw[2 * NG + g], a2);
                    }
                    v[j] = (a2 * 65536 + (a1 * 256 + a0)) >> 22;
                }
                const uint32_t outw = pack_sat_u8x4(v[0], v[1], v[2], v[3]);
                drow[cg] = outw;
                if (fimg) {
                    float4 f;
                    f.x = s_f255[outw & 0xFFu];
                    f.y = s_f255[(outw >> 8) & 0xFFu];
                    f.z = s_f255[(outw >> 16) & 0xFFu];
                    f.w = s_f255[outw >> 24];
                    __stcs(reinterpret_cast<float4*>(fimg) + (size_t)r * G + cg, f);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ distortion
// (1) histogram of x = src + noise (mod 256) per image/channel, (2) autocontrast LUT per
// image/channel (ImageOps.autocontrast), (3) dst = lut[x].
constexpr int DREP = 8;

__global__ void __launch_bounds__(THREADS) k_distort_hist(const uint8_t* __restrict__ src, const uint8_t* __restrict__ noise,
                                                          int32_t* __restrict__ hist, int nbytes, int bytes_per_block,
                                                          const int32_t* __restrict__ sidx) {
    __shared__ uint32_t sh[DREP][3 * 256];
    const int img = blockIdx.y;
    const int begin = blockIdx.x * bytes_per_block;
    const int end = min(nbytes, begin + bytes_per_block);
    if (begin >= end) return;
    for (int i = threadIdx.x; i < DREP * 768; i += THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    uint32_t* my = sh[(threadIdx.x >> 5) % DREP];
    const uint8_t* s = src + (size_t)(sidx ? sidx[img] : img) * nbytes;
    const uint8_t* nz = noise + (size_t)img * nbytes;
    // begin is a multiple of 48 (16 pixels): 16-byte vectors keep channel phase = (j % 3)
    const bool vec = (((reinterpret_cast<uintptr_t>(s + begin) | reinterpret_cast<uintptr_t>(nz + begin)) & 15) == 0);
    const int n16 = vec ? (end - begin) >> 4 : 0;
    for (int i = threadIdx.x; i < n16; i += THREADS) {
        const uint4 a = ld_stream16(s + begin + (size_t)i * 16);
        const uint4 b = ld_stream16(nz + begin + (size_t)i * 16);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        int ch = (i * 16) % 3;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t sum = __vadd4(aw[w], bw[w]);  // per-byte wrap-around add
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&my[ch * 256 + ((sum >> (8 * j)) & 255)], 1u);
                ch = ch == 2 ? 0 : ch + 1;
            }
        }
    }
    for (int i = begin + (n16 << 4) + threadIdx.x; i < end; i += THREADS) {
        const int v = (s[i] + nz[i]) & 255;
        atomicAdd(&my[(i % 3) * 256 + v], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += THREADS) {
        uint32_t v = 0;
#pragma unroll
        for (int r = 0; r < DREP; ++r) v += sh[r][i];
        if (v) atomicAdd(&hist[(size_t)img * 768 + i], (int)v);
    }
}

// One warp per (image, channel): sequential cut logic is 256 steps -- lane 0 runs it.
__global__ void k_distort_lut(int32_t* __restrict__ hist, const int32_t* __restrict__ cut_arr, int B) {
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= B * 3) return;
    const int img = idx / 3;
    int32_t* h = hist + (size_t)idx * 256;
    const int lane = threadIdx.x & 31;
    __shared__ int s_lohi[8][2];
    int* lohi = s_lohi[threadIdx.x >> 5];
    if (lane == 0) {
        int cut = cut_arr[img];
        if (cut > 0) {  // ImageOps.autocontrast: remove `cut` pixels from the low end ...
            for (int lo = 0; lo < 256; ++lo) {
                if (cut > h[lo]) {
                    cut -= h[lo];
                    h[lo] = 0;
                } else {
                    h[lo] -= cut;
                    cut = 0;
                }
                if (cut <= 0) break;
            }
            cut = cut_arr[img];  // ... and from the high end
            for (int hi = 255; hi >= 0; --hi) {
                if (cut > h[hi]) {
                    cut -= h[hi];
                    h[hi] = 0;
                } else {
                    h[hi] -= cut;
                    cut = 0;
                }
                if (cut <= 0) break;
            }
        }
        int lo = 0, hi = 255;
        for (lo = 0; lo < 256; ++lo)
            if (h[lo]) break;
        if (lo == 256) lo = 255;  // Python's `for lo in range(256)` leaves lo = 255 on exhaustion
        for (hi = 255; hi >= 0; --hi)
            if (h[hi]) break;
        if (hi < 0) hi = 0;
        lohi[0] = lo;
        lohi[1] = hi;
    }
    __syncwarp();
    const int lo = lohi[0], hi = lohi[1];
    __syncwarp();
    // the LUT replaces the histogram in place (int32 per entry)
    for (int ix = lane; ix < 256; ix += 32) {
        int v = ix;
        if (hi > lo) {
            const double scale = __ddiv_rn(255.0, (double)(hi - lo));
            const double offset = __dmul_rn(-(double)lo, scale);
            const double t = __dadd_rn(__dmul_rn((double)ix, scale), offset);
            v = (int)t;  // Python int(): truncate toward zero
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        h[ix] = v;
    }
}

__global__ void __launch_bounds__(THREADS) k_distort_apply(const uint8_t* __restrict__ src, const uint8_t* __restrict__ noise,
                                                           uint8_t* __restrict__ dst, const int32_t* __restrict__ lut,
                                                           int nbytes, int bytes_per_block, const int32_t* __restrict__ sidx) {
    __shared__ uint8_t sl[768];
    const int img = blockIdx.y;
    const int begin = blockIdx.x * bytes_per_block;
    const int end = min(nbytes, begin + bytes_per_block);
    if (begin >= end) return;
    for (int i = threadIdx.x; i < 768; i += THREADS) sl[i] = (uint8_t)lut[(size_t)img * 768 + i];
    __syncthreads();
    const uint8_t* s = src + (size_t)(sidx ? sidx[img] : img) * nbytes;
    const uint8_t* nz = noise + (size_t)img * nbytes;
    uint8_t* d = dst + (size_t)img * nbytes;
    const bool vec = (((reinterpret_cast<uintptr_t>(s + begin) | reinterpret_cast<uintptr_t>(nz + begin) |
                        reinterpret_cast<uintptr_t>(d + begin)) & 15) == 0);
    const int n16 = vec ? (end - begin) >> 4 : 0;
    for (int i = threadIdx.x; i < n16; i += THREADS) {
        const uint4 a = ld_stream16(s + begin + (size_t)i * 16);
        const uint4 b = ld_stream16(nz + begin + (size_t)i * 16);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        uint32_t ow[4];
        int ch = (i * 16) % 3;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t sum = __vadd4(aw[w], bw[w]);
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o |= (uint32_t)sl[ch * 256 + ((sum >> (8 * j)) & 255)] << (8 * j);
                ch = ch == 2 ? 0 : ch + 1;
            }
            ow[w] = o;
        }
        st_stream16(d + begin + (size_t)i * 16, make_uint4(ow[0], ow[1], ow[2], ow[3]));
    }
    for (int i = begin + (n16 << 4) + threadIdx.x; i < end; i += THREADS)
        d[i] = sl[(i % 3) * 256 + ((s[i] + nz[i]) & 255)];
}

int chunking(int total, int unit, int B, int* per_block) {
    // split `total` items (multiple-of-`unit` chunks) so that the grid has >= 4 blocks per SM
    int chunks = lfx_div_up((long long)LFX_NUM_SMS * 4, B);
    const int max_chunks = max(1, total / (unit * 4));
    chunks = max(1, min(chunks, max_chunks));
    int pb = lfx_div_up(total, chunks);
    pb = lfx_div_up(pb, unit) * unit;
    *per_block = pb;
    return lfx_div_up(total, pb);
}

}  // namespace

extern "C" int lfx_flip(const uint8_t* src, uint8_t* dst, int B, int H, int W, const int32_t* mode, const int32_t* src_index,
                        int n_src, lfx_stream_t stream) {
    (void)n_src;
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && mode && B >= 0 && H > 0 && W > 0 && B <= 65535, LFX_ERR_ARG, "flip: bad arguments");
    const int rb = W * 3;
    if (rb % 16 == 0 && W % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 &&
        (size_t)FV_ROWS * rb * 2 <= 96 * 1024) {
        const size_t smem_v = (size_t)FV_ROWS * rb * 2;
        static size_t attr_v_[LFX_MAX_DEVICES] = {0};
        size_t& attr_v = attr_v_[lfx_dev()];
        if (smem_v > 48 * 1024 && smem_v > attr_v) {
            cudaFuncSetAttribute(k_flip_vec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v);
            attr_v = smem_v;
        }
        dim3 gridv(lfx_div_up(H, FV_ROWS), B);
        k_flip_vec<<<gridv, THREADS, smem_v, (cudaStream_t)stream>>>(src, dst, H, W, mode, src_index);
        return lfx_check_launch("flip(vec)");
    }
    LFX_REQUIRE(W * 3 * 2 + 32 <= FLIP_SMEM, LFX_ERR_UNSUPPORTED, "flip: W > %d unsupported", (FLIP_SMEM - 32) / 6);
    if (B == 0) return LFX_OK;
    int rows = max(1, (FLIP_SMEM - 32) / 2 / rb);
    rows = min(rows, H);
    // keep chunk byte offsets 16-byte aligned when possible (fast vector path)
    if (rows >= 16) rows &= ~15;
    const size_t smem = (size_t)((rows * rb + 15) & ~15) * 2;
    dim3 grid(lfx_div_up(H, rows), B);
    k_flip<<<grid, THREADS, smem, (cudaStream_t)stream>>>(src, dst, H, W, rows, mode, src_index);
    return lfx_check_launch("flip");
}

extern "C" int lfx_rotate_nn(const uint8_t* src, uint8_t* dst, int64_t dst_image_stride, int B, int H, int W,
                             const int32_t* params, int fill, const int32_t* src_index, int n_src, lfx_stream_t stream) {
    (void)n_src;
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && params && B >= 0 && H > 0 && W > 0 && B <= 65535 && dst_image_stride > 0, LFX_ERR_ARG,
                "rotate_nn: bad arguments");
    LFX_REQUIRE(H < 32768 && W < 32768, LFX_ERR_UNSUPPORTED, "rotate_nn: fixed-point path needs sizes < 32768");
    if (B == 0) return LFX_OK;
    const long long max_px = dst_image_stride / 3;
    dim3 grid(lfx_div_up(max_px, RT_PX), B);
    LFX_REQUIRE((reinterpret_cast<uintptr_t>(params) & 15) == 0, LFX_ERR_ARG, "rotate_nn: params must be 16-byte aligned");
    if ((long long)H * W * 3 < (1ll << 31) && max_px < (1ll << 31) - RT_PX)
        k_rotate_nn<true><<<grid, THREADS, 0, (cudaStream_t)stream>>>(src, dst, dst_image_stride, H, W, params, fill, src_index);
    else
        k_rotate_nn<false><<<grid, THREADS, 0, (cudaStream_t)stream>>>(src, dst, dst_image_stride, H, W, params, fill, src_index);
    return lfx_check_launch("rotate_nn");
}

extern "C" int lfx_warp_bicubic(const uint8_t* src, uint8_t* dst, int B, int H, int W, const double* coef,
                                const int32_t* perspective, const int32_t* src_index, int n_src, lfx_stream_t stream) {
    const int nsrc = src_index ? n_src : B;
    LFX_REQUIRE(!src_index || n_src > 0, LFX_ERR_ARG, "warp_bicubic: src_index needs n_src");
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && coef && perspective && B >= 0 && H > 0 && W > 0 && B <= 65535, LFX_ERR_ARG,
                "warp_bicubic: bad arguments");
    if (B == 0) return LFX_OK;
    // staged source rectangle: up to 70 KB (+ 2 KB queue) so that three blocks share an SM
    const int smem = 70 * 1024;
    static bool attr_[LFX_MAX_DEVICES] = {false};
    bool& attr = attr_[lfx_dev()];
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_bicubic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_bicubic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "warp_bicubic smem attr: %s", cudaGetErrorString(e));
        attr = true;
    }
    const int ntx = lfx_div_up(W, WB_COLS);
    dim3 grid(lfx_div_up(H, WB_ROWS) * ntx, B);
    const uint32_t n16 = (uint32_t)max(1, (W * 3) >> 4);
    const uint32_t m32 = (uint32_t)((0x100000000ull + n16 - 1) / n16);   // i / n16 = umulhi(i, m32) for i < 2^32 / n16
    if (ntx == 1)
        k_warp_bicubic<false><<<grid, THREADS, smem, (cudaStream_t)stream>>>(src, dst, H, W, coef, perspective, smem, ntx, src_index, nsrc, m32);
    else
        k_warp_bicubic<true><<<grid, THREADS, smem, (cudaStream_t)stream>>>(src, dst, H, W, coef, perspective, smem, ntx, src_index, nsrc, m32);
    return lfx_check_launch("warp_bicubic");
}

extern "C" int lfx_crop_lanczos(const uint8_t* src, uint8_t* dst, float* dst_f32, int B, int H, int W, const int32_t* box,
                                int OH, int OW, const int32_t* tab_bounds, const int32_t* tab_kk, int kstride,
                                const int32_t* tab_off, const int32_t* src_index, int n_src, lfx_stream_t stream) {
    const int nsrc = src_index ? n_src : B;
    LFX_REQUIRE(!src_index || n_src > 0, LFX_ERR_ARG, "crop_lanczos: src_index needs n_src");
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && box && tab_bounds && tab_kk && tab_off && B >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0 &&
                    kstride > 0 && B <= 65535,
                LFX_ERR_ARG, "crop_lanczos: bad arguments");
    // dp4a kernel: row strips, both passes on IDP.4A (every upscale and downscales up to 2x: <= 12 taps per axis)
    {
        auto taps_bound = [](int in, int out) {   // max taps of one output sample: the window [c - s, c + s) with s = 3 * max(in / out, 1)
            const double sc = (double)in / out;   // holds at most ceil(2 s) + 1 integers (checked against lfx_lanczos_table over 3,000 size pairs)
            return (int)ceil(6.0 * (sc < 1.0 ? 1.0 : sc) - 1e-9) + 1;
        };
        const int kb = max(taps_bound(W, OW), taps_bound(H, OH));
        const int NG = kb <= 8 ? 2 : (kb <= 12 ? 3 : 0);
        // Strip height: taller strips re-filter fewer overlap rows in the horizontal pass (a strip reads its rows + the tap
        // halo), but two blocks must share an SM: the tallest multiple of 4 whose shared memory fits twice (one block per
        // SM for very wide images), evened out over the strips.  Measured on 4096 x 256^2: 32 rows 1.64 ms, 48 rows 1.55 ms.
        static const int lz_env = getenv("LFX_LZ_TO") ? atoi(getenv("LFX_LZ_TO")) : 0;   // experiments only
        const int raw_pitch = ((W * 3 + 15) & ~15) + 32, ppitch = ((W + 3) & ~3) + 4 * NG + 4;
        const int VKS = (3 * NG + 3) & ~3;
        int mrows_cap = 0;
        auto smem_for = [&](int lz) {
            mrows_cap = (int)(((long long)lz * H + OH - 1) / OH) + 4 * NG + 2;
            const int mrows4 = (mrows_cap + 3) & ~3, nquad = mrows4 / 4 + NG + 1;
            const size_t raw_bytes = (size_t)mrows_cap * raw_pitch, m4_bytes = (size_t)nquad * OW * 3 * 4;
            return (((raw_bytes > m4_bytes ? raw_bytes : m4_bytes) + 15) & ~(size_t)15) + (((size_t)3 * mrows4 * ppitch + 15) & ~(size_t)15) +
                   (size_t)lz * VKS * 4 + lz * 4 + (size_t)mrows_cap * 4 + 16;
        };
        int LZ4_TO = 0;
        if (lz_env >= 4 && lz_env <= LZ4_TO_MAX) {
            LZ4_TO = lz_env & ~3;
        } else {
            for (const size_t cap : {(size_t)111 * 1024, (size_t)220 * 1024}) {
                for (int lz = 52; lz >= 4 && !LZ4_TO; lz -= 4)
                    if (smem_for(lz) <= cap) LZ4_TO = lz;
                if (LZ4_TO) break;
            }
            if (LZ4_TO) {   // same number of strips, evened out
                const int strips = lfx_div_up(OH, LZ4_TO);
                LZ4_TO = min(LZ4_TO, (lfx_div_up(OH, strips) + 3) & ~3);
            }
        }
        const size_t smem4 = LZ4_TO ? smem_for(LZ4_TO) : (size_t)1 << 30;
        static const bool no_dp4a = getenv("LFX_LANCZOS_OLD") != nullptr;   // debug: compare against the previous kernel
        const bool ok4 = NG != 0 && LZ4_TO != 0 && !no_dp4a && (OW % 4 == 0) && smem4 <= 220 * 1024 && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) &&
                         (!dst_f32 || (reinterpret_cast<uintptr_t>(dst_f32) & 15) == 0);
        if (ok4) {
            static size_t attr4_[LFX_MAX_DEVICES][2] = {{0}};
            size_t* attr4 = attr4_[lfx_dev()];
            const void* fn = NG == 2 ? (const void*)k_lanczos_dp4a<2> : (const void*)k_lanczos_dp4a<3>;
            if (smem4 > 48 * 1024 && smem4 > attr4[NG - 2]) {
                cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
                LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "crop_lanczos smem attr: %s", cudaGetErrorString(e));
                attr4[NG - 2] = smem4;
            }
            dim3 grid4(lfx_div_up(OH, LZ4_TO), B);
            if (NG == 2)
                k_lanczos_dp4a<2><<<grid4, LZ4_T, smem4, (cudaStream_t)stream>>>(src, dst, dst_f32, H, W, box, OH, OW, tab_bounds, tab_kk,
                                                                                 kstride, tab_off, mrows_cap, src_index, nsrc, LZ4_TO);
            else
                k_lanczos_dp4a<3><<<grid4, LZ4_T, smem4, (cudaStream_t)stream>>>(src, dst, dst_f32, H, W, box, OH, OW, tab_bounds, tab_kk,
                                                                                 kstride, tab_off, mrows_cap, src_index, nsrc, LZ4_TO);
            return lfx_check_launch("crop_lanczos(dp4a)");
        }
    }
    // strip kernel: full-width row strips, taps in registers (every upscale and mild downscale: <= 8 taps per axis)
    {
        int lz_to = LZ_TO, mrows_cap = 0;
        size_t smem2 = 0;
        for (;; lz_to /= 2) {
            mrows_cap = (int)(((long long)lz_to * H + OH - 1) / OH) + kstride + 2;   // crop_h <= H
            smem2 = (((size_t)mrows_cap * OW * 3 + 15) & ~(size_t)15) + (size_t)lz_to * kstride * 4 + lz_to * 8 +
                    (size_t)((mrows_cap + 3) & ~3) * 4 + (size_t)mrows_cap * (((W * 3 + 15) & ~15) + 32);
            if (smem2 <= 200 * 1024 || lz_to <= 8) break;
        }
        const bool ok = (OW % 4 == 0) && kstride <= 16 && smem2 <= 200 * 1024 &&
                        ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) && (!dst_f32 || (reinterpret_cast<uintptr_t>(dst_f32) & 15) == 0);
        if (ok) {
            // taps per output sample: bounded by the whole-image resize (the crop is never larger than the image)
            const int kmax = max(lfx_lanczos_ksize(W, OW), lfx_lanczos_ksize(H, OH));
            const int vi = kmax <= 8 ? 0 : (kmax <= 10 ? 1 : (kmax <= 12 ? 2 : 3));
            static size_t attr2_[LFX_MAX_DEVICES][4] = {{0}};
            size_t* attr2 = attr2_[lfx_dev()];
            const void* fns[4] = {(const void*)k_crop_lanczos_strip<8>, (const void*)k_crop_lanczos_strip<10>,
                                  (const void*)k_crop_lanczos_strip<12>, (const void*)k_crop_lanczos_strip<16>};
            if (smem2 > 48 * 1024 && smem2 > attr2[vi]) {
                cudaFuncSetAttribute(fns[vi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                attr2[vi] = smem2;
            }
            dim3 grid2(lfx_div_up(OH, lz_to), B);
#define LFX_LZ_LAUNCH(K) \
    k_crop_lanczos_strip<K><<<grid2, THREADS, smem2, (cudaStream_t)stream>>>(src, dst, dst_f32, H, W, box, OH, OW, tab_bounds, tab_kk, kstride, \
                                                                         tab_off, mrows_cap, lz_to, src_index, nsrc)
            if (vi == 0) LFX_LZ_LAUNCH(8);
            else if (vi == 1) LFX_LZ_LAUNCH(10);
            else if (vi == 2) LFX_LZ_LAUNCH(12);
            else LFX_LZ_LAUNCH(16);
#undef LFX_LZ_LAUNCH
            return lfx_check_launch("crop_lanczos(strip)");
        }
    }
    const size_t smem = (size_t)H * LZ_TW * 3;
    LFX_REQUIRE(smem <= 200 * 1024, LFX_ERR_UNSUPPORTED, "crop_lanczos: H > %d unsupported", 200 * 1024 / (LZ_TW * 3));
    if (B == 0) return LFX_OK;
    static size_t attr_[LFX_MAX_DEVICES] = {0};
    size_t& attr = attr_[lfx_dev()];
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(k_crop_lanczos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    dim3 grid(lfx_div_up(OW, LZ_TW), B);
    k_crop_lanczos<<<grid, THREADS, smem, (cudaStream_t)stream>>>(src, dst, dst_f32, H, W, box, OH, OW, tab_bounds, tab_kk,
                                                                  kstride, tab_off, src_index);
    return lfx_check_launch("crop_lanczos");
}

extern "C" int lfx_distort(const uint8_t* src, const uint8_t* noise, uint8_t* dst, int B, int H, int W, const int32_t* cut,
                           int32_t* hist_ws, const int32_t* src_index, int n_src, lfx_stream_t stream) {
    (void)n_src;
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && noise && dst && cut && hist_ws && B >= 0 && H > 0 && W > 0 && B <= 65535, LFX_ERR_ARG,
                "distort: bad arguments");
    LFX_REQUIRE((long long)H * W * 3 < (1ll << 31), LFX_ERR_UNSUPPORTED, "distort: image too large");
    if (B == 0) return LFX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nbytes = H * W * 3;
    cudaError_t e = cudaMemsetAsync(hist_ws, 0, (size_t)B * 768 * sizeof(int32_t), st);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "distort memset: %s", cudaGetErrorString(e));
    int pb = 0;
    const int chunks = chunking(nbytes, 48 * 64, B, &pb);
    dim3 grid(chunks, B);
    k_distort_hist<<<grid, THREADS, 0, st>>>(src, noise, hist_ws, nbytes, pb, src_index);
    k_distort_lut<<<lfx_div_up((long long)B * 3, 8), 256, 0, st>>>(hist_ws, cut, B);
    k_distort_apply<<<grid, THREADS, 0, st>>>(src, noise, dst, hist_ws, nbytes, pb, src_index);
    return lfx_check_launch("distort");
}
