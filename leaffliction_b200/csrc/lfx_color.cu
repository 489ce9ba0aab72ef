// Per-pixel colour kernels: cvtColor, fused threshold masks, apply_mask, colour statistics.
// HBM-bound streaming kernels: 16-byte coalesced loads into shared tiles, 4 pixels per thread,
// LUTs in shared memory, warp-private histogram updates merged with atomics.
#include "lfx_common.cuh"

namespace {

constexpr int TILE_PX = 1024;  // pixels per block iteration (256 threads x 4 pixels)
constexpr int THREADS = 256;

// ------------------------------------------------------------------------------ cvtColor
template <int CODE>
__global__ void __launch_bounds__(THREADS) k_cvt_color(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       long long npix, const LfxTables* __restrict__ tab) {
    __shared__ __align__(16) uint8_t s_in[TILE_PX * 3];
    __shared__ __align__(16) uint8_t s_out[TILE_PX * 3];
    __shared__ HsvLut s_hsv;
    __shared__ LabLut s_lab;
    if (CODE == 1) load_hsv_lut(&s_hsv, tab);
    if (CODE == 2) load_lab_lut(&s_lab, tab);
    const long long ntiles = (npix + TILE_PX - 1) / TILE_PX;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TILE_PX;
        const int n = (int)min((long long)TILE_PX, npix - base);
        block_load_bytes(s_in, src + base * 3, n * 3);
        __syncthreads();
        const int p0 = threadIdx.x * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int p = p0 + k;
            if (p < n) {
                const int r = s_in[p * 3], g = s_in[p * 3 + 1], b = s_in[p * 3 + 2];
                if (CODE == 0) {
                    s_out[p] = (uint8_t)rgb2gray(r, g, b);
                } else if (CODE == 1) {
                    int h, s, v;
                    rgb2hsv(r, g, b, &s_hsv, h, s, v);
                    s_out[p * 3] = (uint8_t)h;
                    s_out[p * 3 + 1] = (uint8_t)s;
                    s_out[p * 3 + 2] = (uint8_t)v;
                } else {
                    int L, A, Bv;
                    rgb2lab(r, g, b, &s_lab, L, A, Bv);
                    s_out[p * 3] = (uint8_t)L;
                    s_out[p * 3 + 1] = (uint8_t)A;
                    s_out[p * 3 + 2] = (uint8_t)Bv;
                }
            }
        }
        __syncthreads();
        if (CODE == 0)
            block_store_bytes(dst + base, s_out, n);
        else
            block_store_bytes(dst + base * 3, s_out, n * 3);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ threshold masks
struct ThreshParams {
    int strategy;  // 0 hsv_h, 1 lab
    int green_lo, green_hi;
};

__global__ void __launch_bounds__(THREADS) k_threshold_mask(const uint8_t* __restrict__ src, uint8_t* __restrict__ mask,
                                                            long long npix, ThreshParams prm,
                                                            const LfxTables* __restrict__ tab) {
    __shared__ __align__(16) uint8_t s_in[TILE_PX * 3];
    __shared__ __align__(16) uint8_t s_out[TILE_PX];
    __shared__ HsvLut s_hsv;
    __shared__ LabLut s_lab;
    if (prm.strategy == 0)
        load_hsv_lut(&s_hsv, tab);
    else
        load_lab_lut(&s_lab, tab);
    const long long ntiles = (npix + TILE_PX - 1) / TILE_PX;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TILE_PX;
        const int n = (int)min((long long)TILE_PX, npix - base);
        block_load_bytes(s_in, src + base * 3, n * 3);
        __syncthreads();
        const int p0 = threadIdx.x * 4;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int p = p0 + k;
            bool on = false;
            if (p < n) {
                const int r = s_in[p * 3], g = s_in[p * 3 + 1], b = s_in[p * 3 + 2];
                if (prm.strategy == 0) {
                    int h, s, v;
                    rgb2hsv(r, g, b, &s_hsv, h, s, v);
                    on = (h >= prm.green_lo) && (h <= prm.green_hi) && (s >= 40);  // mask.py:90
                } else {
                    int L, A, Bv;
                    rgb2lab(r, g, b, &s_lab, L, A, Bv);
                    on = (A <= 135) && (Bv >= 115) && (Bv <= 170);  // mask.py:105
                }
            }
            packed |= (on ? 0xFFu : 0u) << (8 * k);
        }
        reinterpret_cast<uint32_t*>(s_out)[threadIdx.x] = packed;
        __syncthreads();
        block_store_bytes(mask + base, s_out, n);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ register-only variants
// 16 pixels per thread straight from global memory (three 16-byte loads -> byte extraction by PRMT at compile-time
// positions -> arithmetic -> packed 16-byte stores): no shared staging, no barriers in the loop.  Used when the buffers
// are 16-byte aligned; the shared-tile kernels above remain the general path and handle the < 16 pixel tail.
template <int K>
__device__ __forceinline__ int px_byte(const uint32_t (&w)[12]) {   // byte K of the 48-byte group
    return (int)__byte_perm(w[K >> 2], 0u, 0x4440u + (K & 3));
}
__device__ __forceinline__ uint32_t pack4u(int a, int b, int c, int d) {
    return __byte_perm(__byte_perm((uint32_t)a, (uint32_t)b, 0x0040), __byte_perm((uint32_t)c, (uint32_t)d, 0x0040), 0x5410);
}
__device__ __forceinline__ void load48(const uint8_t* p, uint32_t (&w)[12]) {
    const uint4 a = ld_stream16(p), b = ld_stream16(p + 16), c = ld_stream16(p + 32);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
    w[8] = c.x, w[9] = c.y, w[10] = c.z, w[11] = c.w;
}

template <int P>
struct PxLoop {   // compile-time loop over the 16 pixels of a group
    template <class F>
    __device__ __forceinline__ static void run(const uint32_t (&w)[12], F&& f) {
        f(P, px_byte<3 * P>(w), px_byte<3 * P + 1>(w), px_byte<3 * P + 2>(w));
        PxLoop<P + 1>::run(w, f);
    }
};
template <>
struct PxLoop<16> {
    template <class F>
    __device__ __forceinline__ static void run(const uint32_t (&)[12], F&&) {}
};

template <int CODE>
__global__ void __launch_bounds__(THREADS) k_cvt_color_vec(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           long long ngroups, const LfxTables* __restrict__ tab) {
    __shared__ HsvLut s_hsv;
    __shared__ LabLut s_lab;
    if (CODE == 1) load_hsv_lut(&s_hsv, tab);
    if (CODE == 2) load_lab_lut(&s_lab, tab);
    __syncthreads();
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < ngroups; g += (long long)gridDim.x * THREADS) {
        uint32_t w[12];
        load48(src + g * 48, w);
        uint8_t o[48];
        PxLoop<0>::run(w, [&](int p, int r, int gg, int b) {
            if (CODE == 0) {
                o[p] = (uint8_t)rgb2gray(r, gg, b);
            } else if (CODE == 1) {
                int h, sv, v;
                rgb2hsv(r, gg, b, &s_hsv, h, sv, v);
                o[p * 3] = (uint8_t)h, o[p * 3 + 1] = (uint8_t)sv, o[p * 3 + 2] = (uint8_t)v;
            } else {
                int L, A, Bv;
                rgb2lab(r, gg, b, &s_lab, L, A, Bv);
                o[p * 3] = (uint8_t)L, o[p * 3 + 1] = (uint8_t)A, o[p * 3 + 2] = (uint8_t)Bv;
            }
        });
        if (CODE == 0) {
            st_stream16(dst + g * 16, make_uint4(pack4u(o[0], o[1], o[2], o[3]), pack4u(o[4], o[5], o[6], o[7]),
                                                 pack4u(o[8], o[9], o[10], o[11]), pack4u(o[12], o[13], o[14], o[15])));
        } else {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                st_stream16(dst + g * 48 + q * 16,
                            make_uint4(pack4u(o[q * 16], o[q * 16 + 1], o[q * 16 + 2], o[q * 16 + 3]),
                                       pack4u(o[q * 16 + 4], o[q * 16 + 5], o[q * 16 + 6], o[q * 16 + 7]),
                                       pack4u(o[q * 16 + 8], o[q * 16 + 9], o[q * 16 + 10], o[q * 16 + 11]),
                                       pack4u(o[q * 16 + 12], o[q * 16 + 13], o[q * 16 + 14], o[q * 16 + 15])));
        }
    }
}

__global__ void __launch_bounds__(THREADS) k_threshold_mask_vec(const uint8_t* __restrict__ src, uint8_t* __restrict__ mask,
                                                                long long ngroups, ThreshParams prm,
                                                                const LfxTables* __restrict__ tab) {
    __shared__ HsvLut s_hsv;
    __shared__ LabLut s_lab;
    if (prm.strategy == 0)
        load_hsv_lut(&s_hsv, tab);
    else
        load_lab_lut(&s_lab, tab);
    __syncthreads();
    const bool direct = prm.green_lo >= 0 && prm.green_lo <= prm.green_hi && prm.green_hi < 150;
    const int lo12 = prm.green_lo << 12, span12 = (prm.green_hi + 1 - prm.green_lo) << 12;
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < ngroups; g += (long long)gridDim.x * THREADS) {
        uint32_t w[12];
        load48(src + g * 48, w);
        uint8_t o[16];
        PxLoop<0>::run(w, [&](int p, int r, int gg, int b) {
            bool on;
            if (prm.strategy == 0 && direct) {
                // lo <= H <= hi  <=>  lo << 12 <= th < (hi + 1) << 12 for ranges below 150 (negative th means H >= 150);
                // S >= 40  <=>  ts >= 40 << 12: neither the shifts nor the hue fix-up of rgb2hsv are needed
                const int v = max(r, max(gg, b)), d = v - min(r, min(gg, b));
                const int ts = d * s_hsv.sdiv[v] + 2048;
                const int hh = (v == r) ? (gg - b) : (v == gg) ? (b - r + 2 * d) : (r - gg + 4 * d);
                const int th = hh * s_hsv.hdiv[d] + 2048;
                on = ((unsigned)(th - lo12) < (unsigned)span12) && (ts >= (40 << 12));  // mask.py:90
            } else if (prm.strategy == 0) {
                int h, sv, v;
                rgb2hsv(r, gg, b, &s_hsv, h, sv, v);
                on = ((unsigned)(h - prm.green_lo) <= (unsigned)(prm.green_hi - prm.green_lo)) && (sv >= 40);  // mask.py:90
            } else {
                int L, A, Bv;
                rgb2lab(r, gg, b, &s_lab, L, A, Bv);
                on = (A <= 135) && (Bv >= 115) && (Bv <= 170);  // mask.py:105
            }
            o[p] = on ? 255 : 0;
        });
        st_stream16(mask + g * 16, make_uint4(pack4u(o[0], o[1], o[2], o[3]), pack4u(o[4], o[5], o[6], o[7]),
                                              pack4u(o[8], o[9], o[10], o[11]), pack4u(o[12], o[13], o[14], o[15])));
    }
}

// ------------------------------------------------------------------------------ apply_mask
__global__ void __launch_bounds__(THREADS) k_apply_mask(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                        uint8_t* __restrict__ dst, long long npix, int color_val) {
    __shared__ __align__(16) uint8_t s_in[TILE_PX * 3];
    __shared__ __align__(16) uint8_t s_m[TILE_PX];
    const long long ntiles = (npix + TILE_PX - 1) / TILE_PX;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * TILE_PX;
        const int n = (int)min((long long)TILE_PX, npix - base);
        block_load_bytes(s_in, src + base * 3, n * 3);
        block_load_bytes(s_m, mask + base, n);
        __syncthreads();
        // 4 pixels per thread: mask bytes > 127 -> 0xFF selectors (bit 7), spread over the 12 RGB bytes
        const uint32_t c32 = (uint32_t)color_val * 0x01010101u;
        const int n4 = n >> 2;
        for (int t = threadIdx.x; t < n4; t += THREADS) {
            const uint32_t m = reinterpret_cast<const uint32_t*>(s_m)[t];
            const uint32_t k = ((m >> 7) & 0x01010101u) * 0xFFu;      // mask_utils.py:68: keep where mask > 127
            uint32_t* p = reinterpret_cast<uint32_t*>(s_in) + 3 * t;
            const uint32_t k0 = __byte_perm(k, k, 0x1000), k1 = __byte_perm(k, k, 0x2211), k2 = __byte_perm(k, k, 0x3332);
            p[0] = (p[0] & k0) | (c32 & ~k0);                          // :76: paint the rest
            p[1] = (p[1] & k1) | (c32 & ~k1);
            p[2] = (p[2] & k2) | (c32 & ~k2);
        }
        for (int i = n4 * 12 + threadIdx.x; i < n * 3; i += THREADS) {
            if (!(s_m[i / 3] > 127)) s_in[i] = (uint8_t)color_val;
        }
        __syncthreads();
        block_store_bytes(dst + base * 3, s_in, n * 3);
        __syncthreads();
    }
}

// apply_mask, 16 pixels per thread in registers (see the register-only variants above)
__global__ void __launch_bounds__(THREADS) k_apply_mask_vec(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                            uint8_t* __restrict__ dst, long long ngroups, int color_val) {
    const uint32_t c32 = (uint32_t)color_val * 0x01010101u;
    for (long long g = (long long)blockIdx.x * THREADS + threadIdx.x; g < ngroups; g += (long long)gridDim.x * THREADS) {
        uint32_t w[12];
        load48(src + g * 48, w);
        const uint4 mq = ld_stream16(mask + g * 16);
        const uint32_t mw[4] = {mq.x, mq.y, mq.z, mq.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t k = ((mw[j] >> 7) & 0x01010101u) * 0xFFu;      // mask_utils.py:68: keep where mask > 127
            const uint32_t k0 = __byte_perm(k, k, 0x1000), k1 = __byte_perm(k, k, 0x2211), k2 = __byte_perm(k, k, 0x3332);
            w[3 * j] = (w[3 * j] & k0) | (c32 & ~k0);                      // :76: paint the rest
            w[3 * j + 1] = (w[3 * j + 1] & k1) | (c32 & ~k1);
            w[3 * j + 2] = (w[3 * j + 2] & k2) | (c32 & ~k2);
        }
        st_stream16(dst + g * 48, make_uint4(w[0], w[1], w[2], w[3]));
        st_stream16(dst + g * 48 + 16, make_uint4(w[4], w[5], w[6], w[7]));
        st_stream16(dst + g * 48 + 32, make_uint4(w[8], w[9], w[10], w[11]));
    }
}

// ------------------------------------------------------------------------------ colour statistics
// One block handles a contiguous pixel range of ONE image; lane-contiguous pixels straight from global memory (the
// mask byte first: unmasked pixels cost one load), twelve 256-bin histograms in shared memory (neighbouring pixels hit
// neighbouring bins = distinct banks, equal bins merge), merged to global with atomicAdd.  The 1 + 8 + 5 category
// counters of hist.py:188,38-65,248-256 come from per-channel LUTs of byte-packed 0/1 flags (AND = joint predicate)
// accumulated in packed 8-bit counters that are flushed every 128 pixels per thread (lfx_core.cu builds the LUT).
constexpr int STATS_FLUSH = 128;

__global__ void __launch_bounds__(THREADS) k_color_stats(const uint8_t* __restrict__ src, const uint8_t* __restrict__ mask,
                                                         int32_t* __restrict__ hist9, int32_t* __restrict__ hsv3,
                                                         int32_t* __restrict__ counters, int HW, int px_per_block,
                                                         const LfxTables* __restrict__ tab, const uint4* __restrict__ cat_lut) {
    __shared__ uint32_t s_hist[12 * 256];
    __shared__ uint32_t s_cnt[16];
    __shared__ uint4 s_cat[3 * 256];
    __shared__ HsvLut s_hsv;
    __shared__ LabLut s_lab;
    const int img = blockIdx.y;
    const int begin = blockIdx.x * px_per_block;
    const int end = min(HW, begin + px_per_block);
    if (begin >= end) return;
    const bool want9 = hist9 != nullptr;
    const bool wantS = (hsv3 != nullptr) || (counters != nullptr);
    for (int i = threadIdx.x; i < 12 * 256; i += THREADS) s_hist[i] = 0;
    if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < 3 * 256; i += THREADS) s_cat[i] = cat_lut[i];
    load_hsv_lut(&s_hsv, tab);
    if (want9) load_lab_lut(&s_lab, tab);
    __syncthreads();
    const uint8_t* simg = src + (size_t)img * HW * 3;
    const uint8_t* mimg = mask ? mask + (size_t)img * HW : nullptr;
    const int lane = threadIdx.x & 31;
    for (int c0 = begin; c0 < end; c0 += STATS_FLUSH * THREADS) {
        const int c1 = min(end, c0 + STATS_FLUSH * THREADS);
        uint32_t cacc[4] = {0u, 0u, 0u, 0u};
        for (int p = c0 + threadIdx.x; p < c1; p += THREADS) {
            const int mk = mimg ? __ldg(mimg + p) : 255;
            if (mk == 0) continue;
            const uint8_t* px = simg + (size_t)p * 3;
            const int r = __ldg(px), g = __ldg(px + 1), b = __ldg(px + 2);
            int h, sv, v;
            rgb2hsv(r, g, b, &s_hsv, h, sv, v);
            if (want9) {
                int L, A, Bv;
                rgb2lab(r, g, b, &s_lab, L, A, Bv);
                atomicAdd(&s_hist[0 * 256 + r], 1u);
                atomicAdd(&s_hist[1 * 256 + g], 1u);
                atomicAdd(&s_hist[2 * 256 + b], 1u);
                atomicAdd(&s_hist[3 * 256 + h], 1u);
                atomicAdd(&s_hist[4 * 256 + sv], 1u);
                atomicAdd(&s_hist[5 * 256 + v], 1u);
                atomicAdd(&s_hist[6 * 256 + L], 1u);
                atomicAdd(&s_hist[7 * 256 + A], 1u);
                atomicAdd(&s_hist[8 * 256 + Bv], 1u);
            }
            // apply_mask binarises at >127 and paints the rest white (s = 0): never in leaf_mask
            if (wantS && mk > 127) {
                const uint4 qh = s_cat[h], qs = s_cat[256 + sv], qv = s_cat[512 + v];
                const uint32_t q0 = qh.x & qs.x & qv.x;
                cacc[0] += q0;
                cacc[1] += qh.y & qs.y & qv.y;
                cacc[2] += qh.z & qs.z & qv.z;
                cacc[3] += qh.w & qs.w & qv.w;
                if (q0 & 1u) {   // leaf_mask (hist.py:188)
                    atomicAdd(&s_hist[9 * 256 + h], 1u);
                    atomicAdd(&s_hist[10 * 256 + sv], 1u);
                    atomicAdd(&s_hist[11 * 256 + v], 1u);
                }
            }
        }
        if (counters) {
#pragma unroll
            for (int k = 0; k < 14; ++k) {
                const uint32_t v = __reduce_add_sync(0xffffffffu, (cacc[k >> 2] >> (8 * (k & 3))) & 0xFFu);
                if (lane == 0 && v) atomicAdd(&s_cnt[k], v);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 12 * 256; i += THREADS) {
        const uint32_t v = s_hist[i];
        if (!v) continue;
        if (i < 9 * 256) {
            if (hist9) atomicAdd(&hist9[(size_t)img * 9 * 256 + i], (int)v);
        } else if (hsv3) {
            atomicAdd(&hsv3[(size_t)img * 3 * 256 + (i - 9 * 256)], (int)v);
        }
    }
    if (counters && threadIdx.x < 14 && s_cnt[threadIdx.x])
        atomicAdd(&counters[(size_t)img * 16 + threadIdx.x], (int)s_cnt[threadIdx.x]);
}


int stream_grid(long long npix) {
    const long long ntiles = (npix + TILE_PX - 1) / TILE_PX;
    const long long cap = (long long)LFX_NUM_SMS * 8;
    return (int)(ntiles < cap ? ntiles : cap);
}

}  // namespace

extern "C" int lfx_cvt_color(const uint8_t* src, uint8_t* dst, int B, int H, int W, int code, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && dst && B >= 0 && H > 0 && W > 0 && code >= 0 && code <= 2, LFX_ERR_ARG, "cvt_color: bad arguments");
    if (B == 0) return LFX_OK;
    const long long npix = (long long)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    // 16-pixel groups by the register-only kernel, the tail (and unaligned buffers) by the shared-tile kernel
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && npix >= 16) {
        const long long ng = npix / 16;
        const int gv = (int)min((ng + THREADS - 1) / THREADS, (long long)LFX_NUM_SMS * 16);
        if (code == 0)
            k_cvt_color_vec<0><<<gv, THREADS, 0, st>>>(src, dst, ng, lfx_tables());
        else if (code == 1)
            k_cvt_color_vec<1><<<gv, THREADS, 0, st>>>(src, dst, ng, lfx_tables());
        else
            k_cvt_color_vec<2><<<gv, THREADS, 0, st>>>(src, dst, ng, lfx_tables());
        done = ng * 16;
    }
    if (done < npix) {
        const long long rem = npix - done;
        const int grid = stream_grid(rem);
        const uint8_t* s2 = src + done * 3;
        uint8_t* d2 = dst + done * (code == 0 ? 1 : 3);
        if (code == 0)
            k_cvt_color<0><<<grid, THREADS, 0, st>>>(s2, d2, rem, lfx_tables());
        else if (code == 1)
            k_cvt_color<1><<<grid, THREADS, 0, st>>>(s2, d2, rem, lfx_tables());
        else
            k_cvt_color<2><<<grid, THREADS, 0, st>>>(s2, d2, rem, lfx_tables());
    }
    return lfx_check_launch("cvt_color");
}

extern "C" int lfx_threshold_mask(const uint8_t* src, uint8_t* mask, int B, int H, int W, const lfx_mask_cfg* cfg,
                                  lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && mask && cfg && B >= 0 && H > 0 && W > 0, LFX_ERR_ARG, "threshold_mask: bad arguments");
    LFX_REQUIRE(cfg->strategy == 0 || cfg->strategy == 1, LFX_ERR_UNSUPPORTED,
                "threshold_mask: strategy %d needs lfx_make_mask (Otsu) or an external front end", cfg->strategy);
    if (B == 0) return LFX_OK;
    const long long npix = (long long)B * H * W;
    ThreshParams prm{cfg->strategy, cfg->green_lo, cfg->green_hi};
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0 && npix >= 16) {
        const long long ng = npix / 16;
        const int gv = (int)min((ng + THREADS - 1) / THREADS, (long long)LFX_NUM_SMS * 16);
        k_threshold_mask_vec<<<gv, THREADS, 0, (cudaStream_t)stream>>>(src, mask, ng, prm, lfx_tables());
        done = ng * 16;
    }
    if (done < npix)
        k_threshold_mask<<<stream_grid(npix - done), THREADS, 0, (cudaStream_t)stream>>>(src + done * 3, mask + done, npix - done, prm,
                                                                                         lfx_tables());
    return lfx_check_launch("threshold_mask");
}

extern "C" int lfx_apply_mask(const uint8_t* src, const uint8_t* mask, uint8_t* dst, int B, int H, int W,
                              int color_val, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && mask && dst && B >= 0 && H > 0 && W > 0 && color_val >= 0 && color_val <= 255, LFX_ERR_ARG,
                "apply_mask: bad arguments");
    if (B == 0) return LFX_OK;
    const long long npix = (long long)B * H * W;
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && npix >= 16) {
        const long long ng = npix / 16;
        const int gv = (int)min((ng + THREADS - 1) / THREADS, (long long)LFX_NUM_SMS * 16);
        k_apply_mask_vec<<<gv, THREADS, 0, (cudaStream_t)stream>>>(src, mask, dst, ng, color_val);
        done = ng * 16;
    }
    if (done < npix)
        k_apply_mask<<<stream_grid(npix - done), THREADS, 0, (cudaStream_t)stream>>>(src + done * 3, mask + done, dst + done * 3, npix - done,
                                                                                     color_val);
    return lfx_check_launch("apply_mask");
}

extern "C" int lfx_color_stats(const uint8_t* src, const uint8_t* mask, int32_t* hist9, int32_t* hsv3,
                               int32_t* counters, int B, int H, int W, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && B >= 0 && H > 0 && W > 0 && (hist9 || hsv3 || counters), LFX_ERR_ARG, "color_stats: bad arguments");
    LFX_REQUIRE((long long)H * W < (1ll << 31), LFX_ERR_UNSUPPORTED, "color_stats: image too large");
    if (B == 0) return LFX_OK;
    const uint4* cat = lfx_cat_lut();
    LFX_REQUIRE(cat != nullptr, LFX_ERR_CUDA, "color_stats: category LUT upload failed");
    const int HW = H * W;
    // enough blocks to fill the GPU (>= 4 per SM) without shrinking chunks below 4 tiles
    int chunks = lfx_div_up((long long)LFX_NUM_SMS * 4, B);
    const int max_chunks = max(1, HW / (TILE_PX * 4));
    chunks = max(1, min(chunks, max_chunks));
    int px_per_block = lfx_div_up(HW, chunks);
    px_per_block = lfx_div_up(px_per_block, TILE_PX) * TILE_PX;
    chunks = lfx_div_up(HW, px_per_block);
    dim3 grid(chunks, B);
    LFX_REQUIRE(B <= 65535, LFX_ERR_UNSUPPORTED, "color_stats: B > 65535, split the batch");
    k_color_stats<<<grid, THREADS, 0, (cudaStream_t)stream>>>(src, mask, hist9, hsv3, counters, HW, px_per_block, lfx_tables(), cat);
    return lfx_check_launch("color_stats");
}
