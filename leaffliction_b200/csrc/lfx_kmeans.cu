// k-means raw mask candidate (SURVEY.md 8a tier C c2, 8f rank 4): `_create_kmeans_mask`, srcs/transform/filters/mask.py:109-140
// -- cv2.setRNGSeed(12345); cv2.kmeans(Z, 3, criteria (EPS+MAX_ITER, 20, 0.5), attempts 1, KMEANS_PP_CENTERS) on the image
// whose longer side is 256 pixels; the cluster picked by hue / bg_bias / saturation becomes the candidate.
//
// cv::kmeans restated exactly (oracle/spec_kmeans.py is checked bit for bit against cv2.kmeans):
//   * cv::RNG multiply-with-carry stream from the pinned seed;
//   * generateCentersPP: squared distances between 8-bit colours are integers, so every sum is exact whatever the
//     order; the sequential `p -= dist[i]` scan = first index whose inclusive prefix sum reaches ceil(p).  The three
//     trials of one centre draw their p from the same sum0, so ONE pass evaluates all three;
//   * Lloyd: float32 distances with each product and sum rounded (no FMA), first strictly smaller wins; the member sums
//     are integers < 2^24 (exact in float32), centre = sum * (1.f / count); stop on iteration 20 or shift^2 <= 0.25;
//     the labels of the last assignment are returned with the new centres; empty clusters as in kmeans.cpp.
// One 1024-thread block per image, the whole image (<= 196,608 bytes) resident in shared memory for the ~5-20 passes.
// Warp w owns the contiguous point range [w*R, (w+1)*R), lane l the points w*R + l + 32 j: byte loads of a warp
// are consecutive (no bank conflicts) and a warp's range sum is a prefix-sum building block.
#include "lfx_common.cuh"

namespace {

constexpr int KT = 1024, KW = KT / 32, KK = 3, KMAX_ITER = 20;
constexpr unsigned long long CV_RNG_COEFF = 4164903690ull;

struct KmState {
    unsigned long long rng;
    int chosen[KK];                 // k-means++ point indices
    int cand[3];                    // the three trial candidates of the centre being drawn
    long long target[3];            // ceil(p) of each trial
    long long wsum[KW];             // range sums of the current k-means++ distance
    long long wnew[3][KW];          // ... for each trial
    float centers[KK][3], old[KK][3];
    int sums[KK][4];                // r, g, b, count of the current assignment
    int ov_idx[2], ov_lab[2], n_ov; // labels overridden by the empty-cluster rule in the last centre update
    unsigned long long far_key;
    int pick, iters, done, empties;
};

__device__ __forceinline__ uint32_t rng_next(unsigned long long& s) {
    s = (unsigned long long)(uint32_t)s * CV_RNG_COEFF + (s >> 32);
    return (uint32_t)s;
}
__device__ __forceinline__ double rng_double(unsigned long long& s) {
    const uint32_t t = rng_next(s);
    const unsigned long long u = ((unsigned long long)t << 32) | rng_next(s);
    return __dmul_rn(__ull2double_rn(u), 5.4210108624275221700372640043497e-20);
}

__device__ __forceinline__ int idist(const uint8_t* px, int i, int cr, int cg, int cb) {
    const int dr = px[3 * i] - cr, dg = px[3 * i + 1] - cg, db = px[3 * i + 2] - cb;
    return dr * dr + dg * dg + db * db;
}

// hal::normL2Sqr_ scalar tail for n = 3: d = 0; d += t*t (each product and sum rounded to float32)
__device__ __forceinline__ float fdist(float r, float g, float b, const float* c) {
    const float t0 = __fsub_rn(r, c[0]), t1 = __fsub_rn(g, c[1]), t2 = __fsub_rn(b, c[2]);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    return __fadd_rn(d, __fmul_rn(t2, t2));
}
__device__ __forceinline__ int label_of(float r, float g, float b, const float (*c)[3]) {
    float m = fdist(r, g, b, c[0]);
    int kb = 0;
    const float d1 = fdist(r, g, b, c[1]);
    if (m > d1) { m = d1; kb = 1; }
    const float d2 = fdist(r, g, b, c[2]);
    if (m > d2) kb = 2;
    return kb;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(KT, 1)
    k_kmeans_raw(const uint8_t* __restrict__ src, uint8_t* __restrict__ raw, float* __restrict__ centers_out,
                 int32_t* __restrict__ kinfo, int B, int H, int W, int green_lo, int green_hi, int bias, uint32_t seed,
                 const LfxTables* __restrict__ tab) {
    extern __shared__ __align__(16) uint8_t s_px[];
    __shared__ KmState S;
    __shared__ HsvLut s_hsv;
    const int N = H * W;
    const int R = (((N + KW - 1) / KW) + 31) & ~31;   // points per warp range, multiple of 32
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int lo = min(N, wid * R), hi = min(N, lo + R);
    load_hsv_lut(&s_hsv, tab);

    for (int img = blockIdx.x; img < B; img += gridDim.x) {
        __syncthreads();
        block_load_bytes(s_px, src + (size_t)img * N * 3, N * 3);
        if (threadIdx.x == 0) {
            S.rng = (unsigned long long)seed;
            S.chosen[0] = (int)(rng_next(S.rng) % (uint32_t)N);
            S.n_ov = 0;
            S.empties = 0;
            S.done = 0;
        }
        __syncthreads();

        // ------------------------------------------------------------------ k-means++ (generateCentersPP, 3 trials)
        {
            const int c0 = S.chosen[0];
            const int cr = s_px[3 * c0], cg = s_px[3 * c0 + 1], cb = s_px[3 * c0 + 2];
            uint32_t acc = 0;
            for (int i = lo + lane; i < hi; i += 32) acc += (uint32_t)idist(s_px, i, cr, cg, cb);
            const long long ws = warp_sum_ll((long long)acc);
            if (lane == 0) S.wsum[wid] = ws;
        }
        __syncthreads();
        for (int k = 1; k < KK; ++k) {
            if (threadIdx.x == 0) {
                long long sum0 = 0;
                for (int w = 0; w < KW; ++w) sum0 += S.wsum[w];
                for (int j = 0; j < 3; ++j) {
                    const double p = __dmul_rn(rng_double(S.rng), __ll2double_rn(sum0));   // sum0 < 2^53: exact
                    S.target[j] = __double2ll_ru(p);
                }
            }
            __syncthreads();
            // the centres chosen so far (k of them)
            int ccr[2], ccg[2], ccb[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c = S.chosen[q < k ? q : 0];
                ccr[q] = s_px[3 * c]; ccg[q] = s_px[3 * c + 1]; ccb[q] = s_px[3 * c + 2];
            }
            auto cur = [&](int i) {
                int d = idist(s_px, i, ccr[0], ccg[0], ccb[0]);
                if (k > 1) d = min(d, idist(s_px, i, ccr[1], ccg[1], ccb[1]));
                return d;
            };
            if (wid < 3) {
                // trial `wid`: first index i < N - 1 whose inclusive prefix sum of the current distance reaches the target
                const long long T = S.target[wid];
                long long run = S.wsum[lane];                     // inclusive scan of the 32 range sums
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long v = __shfl_up_sync(0xffffffffu, run, o);
                    if (lane >= o) run += v;
                }
                const uint32_t reach = __ballot_sync(0xffffffffu, run >= T);
                int ci = N - 1;
                if (reach) {
                    const int w = __ffs(reach) - 1;
                    long long base = __shfl_sync(0xffffffffu, run, w) - __shfl_sync(0xffffffffu, S.wsum[lane], w);
                    const int a = min(N, w * R), b = min(N, a + R);
                    for (int i0 = a; i0 < b; i0 += 32) {
                        const int i = i0 + lane;
                        long long d = (i < b) ? (long long)cur(i) : 0ll;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const long long v = __shfl_up_sync(0xffffffffu, d, o);
                            if (lane >= o) d += v;
                        }
                        const uint32_t hit = __ballot_sync(0xffffffffu, (i < b) && (base + d >= T));
                        if (hit) {
                            ci = min(N - 1, i0 + __ffs(hit) - 1);
                            break;
                        }
                        base += __shfl_sync(0xffffffffu, d, 31);
                    }
                }
                if (lane == 0) S.cand[wid] = ci;
            }
            __syncthreads();
            {
                int tr[3], tg[3], tb[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int c = S.cand[j];
                    tr[j] = s_px[3 * c]; tg[j] = s_px[3 * c + 1]; tb[j] = s_px[3 * c + 2];
                }
                uint32_t a0 = 0, a1 = 0, a2 = 0;
                for (int i = lo + lane; i < hi; i += 32) {
                    const int d = cur(i);
                    a0 += (uint32_t)min(d, idist(s_px, i, tr[0], tg[0], tb[0]));
                    a1 += (uint32_t)min(d, idist(s_px, i, tr[1], tg[1], tb[1]));
                    a2 += (uint32_t)min(d, idist(s_px, i, tr[2], tg[2], tb[2]));
                }
                const long long w0 = warp_sum_ll((long long)a0), w1 = warp_sum_ll((long long)a1), w2 = warp_sum_ll((long long)a2);
                if (lane == 0) { S.wnew[0][wid] = w0; S.wnew[1][wid] = w1; S.wnew[2][wid] = w2; }
            }
            __syncthreads();
            if (wid == 0) {
                long long s0 = warp_sum_ll(S.wnew[0][lane]), s1 = warp_sum_ll(S.wnew[1][lane]), s2 = warp_sum_ll(S.wnew[2][lane]);
                int best = 0;
                long long bs = s0;
                if (s1 < bs) { bs = s1; best = 1; }
                if (s2 < bs) { bs = s2; best = 2; }
                S.wsum[lane] = S.wnew[best][lane];
                if (lane == 0) S.chosen[k] = S.cand[best];
            }
            __syncthreads();
        }
        if (threadIdx.x < KK * 3) {
            const int k = threadIdx.x / 3, j = threadIdx.x - 3 * k;
            S.centers[k][j] = (float)s_px[3 * S.chosen[k] + j];
        }
        if (threadIdx.x == 0) S.iters = 0;
        __syncthreads();

        // ------------------------------------------------------------------ Lloyd iterations
        for (;;) {
            if (threadIdx.x < KK * 4) S.sums[threadIdx.x >> 2][threadIdx.x & 3] = 0;
            __syncthreads();
            {
                float c[KK][3];
#pragma unroll
                for (int k = 0; k < KK; ++k)
#pragma unroll
                    for (int j = 0; j < 3; ++j) c[k][j] = S.centers[k][j];
                // <= 64 points per lane: 16-bit packed partial sums (r | g << 16, b | count << 16) per cluster
                uint32_t rg[KK] = {0, 0, 0}, bn[KK] = {0, 0, 0};
                for (int i = lo + lane; i < hi; i += 32) {
                    const uint32_t r = s_px[3 * i], g = s_px[3 * i + 1], b = s_px[3 * i + 2];
                    const int kb = label_of((float)r, (float)g, (float)b, c);
                    const uint32_t v0 = r | (g << 16), v1 = b | 0x10000u;
#pragma unroll
                    for (int k = 0; k < KK; ++k) {
                        rg[k] += (kb == k) ? v0 : 0u;
                        bn[k] += (kb == k) ? v1 : 0u;
                    }
                }
#pragma unroll
                for (int k = 0; k < KK; ++k) {
                    const int sr = __reduce_add_sync(0xffffffffu, rg[k] & 0xFFFFu), sg = __reduce_add_sync(0xffffffffu, rg[k] >> 16);
                    const int sb = __reduce_add_sync(0xffffffffu, bn[k] & 0xFFFFu), sn = __reduce_add_sync(0xffffffffu, bn[k] >> 16);
                    if (lane == 0) {
                        atomicAdd(&S.sums[k][0], sr); atomicAdd(&S.sums[k][1], sg);
                        atomicAdd(&S.sums[k][2], sb); atomicAdd(&S.sums[k][3], sn);
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                S.n_ov = 0;
                for (int k = 0; k < KK; ++k)
                    for (int j = 0; j < 3; ++j) S.old[k][j] = S.centers[k][j];
            }
            __syncthreads();
            // empty clusters (kmeans.cpp): the farthest member (last on ties) of the biggest cluster moves
            for (int k = 0; k < KK; ++k) {
                if (S.sums[k][3] != 0) continue;          // block-uniform
                __syncthreads();
                int max_k = 0;
                for (int k1 = 1; k1 < KK; ++k1)
                    if (S.sums[max_k][3] < S.sums[k1][3]) max_k = k1;
                const float scale = __fdiv_rn(1.f, (float)S.sums[max_k][3]);
                float base[3];
                for (int j = 0; j < 3; ++j) base[j] = __fmul_rn((float)S.sums[max_k][j], scale);
                if (threadIdx.x == 0) S.far_key = 0ull;
                __syncthreads();
                unsigned long long key = 0ull;
                bool any = false;
                for (int i = lo + lane; i < hi; i += 32) {
                    const float r = (float)s_px[3 * i], g = (float)s_px[3 * i + 1], b = (float)s_px[3 * i + 2];
                    int lab = label_of(r, g, b, S.old);
                    for (int q = 0; q < S.n_ov; ++q)
                        if (S.ov_idx[q] == i) lab = S.ov_lab[q];
                    if (lab != max_k) continue;
                    const unsigned long long kk = ((unsigned long long)__float_as_uint(fdist(r, g, b, base)) << 32) | (uint32_t)i;
                    if (!any || kk > key) key = kk;
                    any = true;
                }
                if (any) atomicMax(&S.far_key, key | (1ull << 63));   // bit 63 marks "a member exists" (distances are >= 0)
                __syncthreads();
                if (threadIdx.x == 0 && (S.far_key >> 63)) {
                    const int far = (int)(uint32_t)S.far_key;
                    S.sums[max_k][3]--; S.sums[k][3]++;
                    for (int j = 0; j < 3; ++j) {
                        S.sums[max_k][j] -= s_px[3 * far + j];
                        S.sums[k][j] += s_px[3 * far + j];
                    }
                    S.ov_idx[S.n_ov] = far; S.ov_lab[S.n_ov] = k; S.n_ov++;
                    S.empties++;
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                double shift = 0.0;
                for (int k = 0; k < KK; ++k) {
                    const float scale = __fdiv_rn(1.f, (float)S.sums[k][3]);
                    double dist = 0.0;
                    for (int j = 0; j < 3; ++j) {
                        const float cj = __fmul_rn((float)S.sums[k][j], scale);
                        S.centers[k][j] = cj;
                        const double t = __dsub_rn((double)cj, (double)S.old[k][j]);
                        dist = __dadd_rn(dist, __dmul_rn(t, t));
                    }
                    shift = fmax(shift, dist);
                }
                // iteration count as kmeans.cpp's `iter`: the k-means++ pass was iteration 0 -> 1
                const int it = S.iters + 2;
                S.iters = S.iters + 1;
                S.done = (it == KMAX_ITER || shift <= 0.25) ? 1 : 0;
            }
            __syncthreads();
            if (S.done) break;
        }

        // ------------------------------------------------------------------ pick the cluster (mask.py:123-136), write the mask
        if (threadIdx.x == 0) {
            int c8[KK][3], hs[KK][3], green[KK];
            bool any_green = false;
            for (int k = 0; k < KK; ++k) {
                for (int j = 0; j < 3; ++j) c8[k][j] = (int)(uint8_t)(int)S.centers[k][j];   // centers.astype(np.uint8)
                rgb2hsv(c8[k][0], c8[k][1], c8[k][2], &s_hsv, hs[k][0], hs[k][1], hs[k][2]);
                green[k] = (hs[k][0] >= green_lo && hs[k][0] <= green_hi && hs[k][1] >= 40) ? 1 : 0;
                any_green |= green[k] != 0;
            }
            int pick = 0;
            auto tot = [&](int k) { return c8[k][0] + c8[k][1] + c8[k][2]; };
            if (bias == 1) {          // dark_bg: brightest centre
                for (int k = 1; k < KK; ++k) if (tot(k) > tot(pick)) pick = k;
            } else if (bias == 2) {   // light_bg: darkest centre
                for (int k = 1; k < KK; ++k) if (tot(k) < tot(pick)) pick = k;
            } else if (any_green) {
                for (int k = KK - 1; k >= 0; --k) if (green[k]) pick = k;
            } else {
                for (int k = 1; k < KK; ++k) if (hs[k][1] > hs[pick][1]) pick = k;
            }
            S.pick = pick;
            if (kinfo) {
                kinfo[(size_t)img * 4 + 0] = pick;
                kinfo[(size_t)img * 4 + 1] = S.iters + 1;
                kinfo[(size_t)img * 4 + 2] = S.empties;
                kinfo[(size_t)img * 4 + 3] = N;
            }
            if (centers_out)
                for (int k = 0; k < KK; ++k)
                    for (int j = 0; j < 3; ++j) centers_out[(size_t)img * 9 + k * 3 + j] = S.centers[k][j];
        }
        __syncthreads();
        {
            const int pick = S.pick;
            uint8_t* out = raw + (size_t)img * N;
            for (int i = lo + lane; i < hi; i += 32) {
                int lab = label_of((float)s_px[3 * i], (float)s_px[3 * i + 1], (float)s_px[3 * i + 2], S.old);
                for (int q = 0; q < S.n_ov; ++q)
                    if (S.ov_idx[q] == i) lab = S.ov_lab[q];
                out[i] = (lab == pick) ? 255 : 0;
            }
        }
    }
}

}  // namespace

extern "C" int lfx_kmeans_raw(const uint8_t* src, uint8_t* raw, float* centers, int32_t* kinfo, int B, int H, int W, int green_lo,
                              int green_hi, int bias, uint32_t seed, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && raw && B > 0 && H > 0 && W > 0, LFX_ERR_ARG, "kmeans_raw: bad argument");
    LFX_REQUIRE(bias >= 0 && bias <= 2, LFX_ERR_ARG, "kmeans_raw: bias %d (0 auto, 1 dark_bg, 2 light_bg)", bias);
    // _create_kmeans_mask resizes so that the longer side is 256 (mask.py:113-118); only that size (no resize) is built
    LFX_REQUIRE(max(H, W) == 256, LFX_ERR_UNSUPPORTED, "kmeans_raw: longer side %d != 256 (the INTER_AREA working copy is not built)",
                max(H, W));
    const int smem = ((H * W * 3 + 15) / 16) * 16;
    static int attr_[LFX_MAX_DEVICES] = {0};
    int& attr = attr_[lfx_dev()];
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(k_kmeans_raw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "kmeans_raw smem attr (%d bytes): %s", smem, cudaGetErrorString(e));
        attr = smem;
    }
    const int grid = min(B, LFX_NUM_SMS);
    k_kmeans_raw<<<grid, KT, smem, (cudaStream_t)stream>>>(src, raw, centers, kinfo, B, H, W, green_lo, green_hi, bias, seed, lfx_tables());
    return lfx_check_launch("kmeans_raw");
}
