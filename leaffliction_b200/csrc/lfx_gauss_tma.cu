// cv2.GaussianBlur for uint8 (bit-exact 8.8 x 8.8 fixed point, BORDER_REFLECT_101) -- TMA-tiled kernel for the two
// sizes the reference uses (5x5: blur.py:72, mask.py:223; 15x15: blur.py:61, mask.py:770), C = 1 or 3 channels.
//
// A block owns TR full-width image rows.  Rows y0-R .. y0+TR+R-1 arrive in shared memory by cp.async.bulk
// (one bulk copy for the interior rows, one per reflected border row) on an mbarrier.  Vertical pass: one
// 32-bit column of bytes per thread walking down the rows with a K-row register window, taps applied to byte
// pairs packed 2 x 16 bit (sums <= 255 * 256 fit, symmetric taps share a multiply), results stored as
// de-interleaved 16-bit planes.  Horizontal pass: dp2a, two taps per instruction, (v + 32768) >> 16 is byte 2
// of the accumulator; 12 output bytes (4 pixels) per thread.  HBM traffic = image in + image out; halo rows are
// L2 hits.  Shapes this kernel does not take (rows not a multiple of 16 bytes, very wide rows) use lfx_gauss.cu.
#include "lfx_common.cuh"

namespace {

constexpr int GT = 256;  // threads

struct GaussTmaParams {
    int H, W, TR, VP;           // VP = plane pitch in elements (W + 2*PAD)
    int t[8];                   // symmetric taps t[0..R]
    uint32_t kev[8], kod[8];    // dp2a tap pairs for even / odd output columns
};

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     g_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(g_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int g_refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

template <int K, int C>
__global__ void __launch_bounds__(GT) k_gauss_tma(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const GaussTmaParams P) {
    constexpr int R = K / 2;
    constexpr int PAD = R + (R & 1);  // even, so that pixel groups of 4 stay 8-byte aligned in the planes
    constexpr int NP = R + 1;         // dp2a pairs per output value
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t s_bar;
    const int H = P.H, W = P.W, TR = P.TR, VP = P.VP, RB = W * C;
    const int y0 = blockIdx.x * TR, nr = min(TR, H - y0);
    const int img = blockIdx.y;
    uint8_t* s_src = sm;                                                                       // [(TR + 2R)][RB]
    uint16_t* s_v = reinterpret_cast<uint16_t*>(sm + (((size_t)(TR + 2 * R) * RB + 127) & ~(size_t)127));  // [C][TR][VP]
    const uint8_t* simg = src + (size_t)img * H * RB;
    uint8_t* dimg = dst + (size_t)img * H * RB;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(g_smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int first = y0 - R, rows = nr + 2 * R;
        const int lo = max(first, 0), hi = min(first + rows - 1, H - 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(&s_bar)), "r"((uint32_t)rows * RB) : "memory");
        g_bulk_g2s(s_src + (size_t)(lo - first) * RB, simg + (size_t)lo * RB, (uint32_t)(hi - lo + 1) * RB, &s_bar);
        for (int t = 0; t < lo - first; ++t) g_bulk_g2s(s_src + (size_t)t * RB, simg + (size_t)g_refl101(first + t, H) * RB, RB, &s_bar);
        for (int t = hi - first + 1; t < rows; ++t) g_bulk_g2s(s_src + (size_t)t * RB, simg + (size_t)g_refl101(first + t, H) * RB, RB, &s_bar);
    }
    __syncthreads();
    {
        uint32_t ok;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok)
                         : "r"(g_smem_u32(&s_bar))
                         : "memory");
        } while (!ok);
    }

    // ---------------- vertical pass: one 32-bit byte column per thread, K-row register window
    const int ncolw = RB >> 2;
    const int pstride = TR * VP;  // plane stride (elements)
    for (int cw = threadIdx.x; cw < ncolw; cw += GT) {
        // destination of the 4 bytes of this column word: byte j = 4cw + b -> pixel j / C, channel j % C
        int dofs[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = 4 * cw + b;
            dofs[b] = (C == 1) ? (j + PAD) : ((j % 3) * pstride + j / 3 + PAD);
        }
        uint32_t lo[K], hi[K];
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(s_src) + cw;
#pragma unroll 1
        for (int base = 0; base < nr + K - 1; base += K) {
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int step = base + s;
                if (step < nr + K - 1) {
                    const uint32_t w = sp[(size_t)step * ncolw];
                    lo[s] = __byte_perm(w, 0u, 0x4240);
                    hi[s] = __byte_perm(w, 0u, 0x4341);
                    if (step >= K - 1) {
                        // window rows step-K+1 .. step sit in slots (s+1)%K .. s; tap i <-> slot (s + 1 + i) % K
                        uint32_t al = lo[(s + 1 + R) % K] * (uint32_t)P.t[R], ah = hi[(s + 1 + R) % K] * (uint32_t)P.t[R];
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            al += (lo[(s + 1 + i) % K] + lo[(s + K - i) % K]) * (uint32_t)P.t[i];
                            ah += (hi[(s + 1 + i) % K] + hi[(s + K - i) % K]) * (uint32_t)P.t[i];
                        }
                        uint16_t* vr = s_v + (size_t)(step - (K - 1)) * VP;
                        if (C == 1) {
                            // 4 consecutive pixels of the single plane: (p0,p1), (p2,p3)
                            uint32_t* v32 = reinterpret_cast<uint32_t*>(vr + dofs[0]);
                            v32[0] = __byte_perm(al, ah, 0x5410);
                            v32[1] = __byte_perm(al, ah, 0x7632);
                        } else {
                            vr[dofs[0]] = (uint16_t)(al & 0xFFFFu);
                            vr[dofs[1]] = (uint16_t)(ah & 0xFFFFu);
                            vr[dofs[2]] = (uint16_t)(al >> 16);
                            vr[dofs[3]] = (uint16_t)(ah >> 16);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---------------- horizontal REFLECT_101 pads: x = -r mirrors x = r, x = W-1+r mirrors x = W-1-r
    for (int i = threadIdx.x; i < C * nr * R * 2; i += GT) {
        const int r = (i % R) + 1;
        const int side = (i / R) & 1;
        const int row = (i / (2 * R)) % nr;
        const int c = i / (2 * R * nr);
        uint16_t* v = s_v + (size_t)c * pstride + (size_t)row * VP + PAD;
        if (side == 0) v[-r] = v[r]; else v[W - 1 + r] = v[W - 1 - r];
    }
    __syncthreads();
    // ---------------- horizontal pass: 4 pixels per thread, NP dp2a per output value
    const int G = W >> 2;
    for (int item = threadIdx.x; item < G * nr; item += GT) {
        const int r = item / G, g = item - r * G;
        uint32_t acc[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            // words covering elements x-PAD .. x+3+PAD (x = 4g); word k holds elements (x-PAD+2k, x-PAD+2k+1)
            const uint2* q = reinterpret_cast<const uint2*>(s_v + (size_t)c * pstride + (size_t)r * VP) + g;
            uint32_t w[PAD + 4];
#pragma unroll
            for (int k = 0; k < (PAD + 4) / 2; ++k) {
                const uint2 t2 = q[k];
                w[2 * k] = t2.x;
                w[2 * k + 1] = t2.y;
            }
            // output x+e (e = 0..3): first tap x+e-R; even offsets start on a word boundary shifted by (PAD-R)
            uint32_t a0 = 32768u, a1 = 32768u, a2 = 32768u, a3 = 32768u;
            constexpr int S = (PAD - R);  // 0 when R even, 1 when R odd: element x-R sits at lane S of word 0
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                // S == 0: x uses "even" pairs starting at word 0, x+1 "odd" pairs starting at word 0 (first weight 0)
                // S == 1: x uses "odd"-style pairs (0,t0) starting at word 0, x+1 even pairs starting at word 1
                a0 = __dp2a_lo(w[k], S ? P.kod[k] : P.kev[k], a0);
                a1 = __dp2a_lo(w[k + S], S ? P.kev[k] : P.kod[k], a1);
                a2 = __dp2a_lo(w[k + 1], S ? P.kod[k] : P.kev[k], a2);
                a3 = __dp2a_lo(w[k + 1 + S], S ? P.kev[k] : P.kod[k], a3);
            }
            acc[c][0] = a0; acc[c][1] = a1; acc[c][2] = a2; acc[c][3] = a3;
        }
        uint8_t* orow = dimg + (size_t)(y0 + r) * RB;
        if (C == 1) {
            const uint32_t lo2 = __byte_perm(acc[0][0], acc[0][1], 0x0062), hi2 = __byte_perm(acc[0][2], acc[0][3], 0x0062);
            reinterpret_cast<uint32_t*>(orow)[g] = __byte_perm(lo2, hi2, 0x5410);
        } else {
            auto pack4 = [](uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
                return __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
            };
            uint32_t* o = reinterpret_cast<uint32_t*>(orow) + 3 * g;
            o[0] = pack4(acc[0][0], acc[1 % C][0], acc[2 % C][0], acc[0][1]);
            o[1] = pack4(acc[1 % C][1], acc[2 % C][1], acc[0][2], acc[1 % C][2]);
            o[2] = pack4(acc[2 % C][2], acc[0][3], acc[1 % C][3], acc[2 % C][3]);
        }
    }
}

template <int K, int C>
int launch_tma(const uint8_t* src, uint8_t* dst, int B, int H, int W, const int32_t* taps, cudaStream_t st) {
    constexpr int R = K / 2, PAD = R + (R & 1);
    const int RB = W * C, VP = W + 2 * PAD;
    // rows per tile: as many as fit ~56 KB (four blocks per SM), at least 4, at most 64
    auto smem_for = [&](int tr) { return (((size_t)(tr + 2 * R) * RB + 127) & ~(size_t)127) + (size_t)C * tr * VP * 2; };
    int TR = 64;
    while (TR > 4 && smem_for(TR) > 56 * 1024) TR -= 4;
    if (smem_for(TR) > 200 * 1024) return 1;  // not for this kernel
    TR = min(TR, ((H + 3) / 4) * 4);
    GaussTmaParams P;
    memset(&P, 0, sizeof(P));
    P.H = H; P.W = W; P.TR = TR; P.VP = VP;
    for (int i = 0; i <= R; ++i) P.t[i] = taps[i];
    // even pairs: (t0,t1),(t2,t3),...,(t_{K-1},0); odd pairs: (0,t0),(t1,t2),...,(t_{K-2},t_{K-1})
    for (int k = 0; k <= R; ++k) {
        const int e0 = 2 * k, e1 = 2 * k + 1, o0 = 2 * k - 1, o1 = 2 * k;
        P.kev[k] = (uint32_t)(e0 < K ? taps[e0] : 0) | ((uint32_t)(e1 < K ? taps[e1] : 0) << 8);
        P.kod[k] = (uint32_t)(o0 >= 0 ? taps[o0] : 0) | ((uint32_t)(o1 < K ? taps[o1] : 0) << 8);
    }
    const size_t smem = smem_for(TR);
    static size_t attr = 0;
    if (smem > 40 * 1024 && smem > attr) {  // static shared memory counts towards the 48 KB default limit
        if (cudaFuncSetAttribute(k_gauss_tma<K, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
        attr = smem;
    }
    dim3 grid((H + TR - 1) / TR, B);
    k_gauss_tma<K, C><<<grid, GT, smem, st>>>(src, dst, P);
    return 0;
}

}  // namespace

// 0 = launched; 1 = shape / taps not handled here (caller falls back to lfx_gauss.cu's kernels)
int lfx_gauss_tma_try(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int ksize, const int32_t* taps, cudaStream_t st) {
    const int R = ksize / 2;
    if ((ksize != 5 && ksize != 15) || (C != 1 && C != 3)) return 1;
    if ((W * C) % 16 != 0 || W % 4 != 0 || H <= R || W <= R || B > 65535) return 1;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) != 0) return 1;
    for (int i = 0; i < ksize; ++i)
        if (taps[i] < 0 || taps[i] > 255 || taps[i] != taps[ksize - 1 - i]) return 1;
    if (ksize == 5) return C == 3 ? launch_tma<5, 3>(src, dst, B, H, W, taps, st) : launch_tma<5, 1>(src, dst, B, H, W, taps, st);
    return C == 3 ? launch_tma<15, 3>(src, dst, B, H, W, taps, st) : launch_tma<15, 1>(src, dst, B, H, W, taps, st);
}
