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
    int H, W, TR, VP;           // VP = plane pitch in elements (TW + 2*PAD)
    int TW, SP;                 // column tile width (pixels; == W: full rows) and staged row pitch (bytes)
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

// FULL = true: tiles of whole rows (x0 = 0, one contiguous bulk copy, every staged byte has a destination): the column
// bookkeeping below folds away at compile time.
template <int K, int C, bool FULL>
__global__ void __launch_bounds__(GT) k_gauss_tma(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const GaussTmaParams P) {
    constexpr int R = K / 2;
    constexpr int PAD = R + (R & 1);  // even, so that pixel groups of 4 stay 8-byte aligned in the planes
    constexpr int NP = R + 1;         // dp2a pairs per output value
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t s_bar;
    const int H = P.H, W = P.W, TR = P.TR, VP = P.VP, RB = W * C, SP = FULL ? RB : P.SP;
    const int y0 = blockIdx.x * TR, nr = min(TR, H - y0);
    const int img = blockIdx.y;
    // column tile: output pixels [x0, x0 + tw); source pixels [pxa, pxb) = the tile + R halo columns inside the image; the
    // staged segment of a row is the 16-byte aligned byte range [as, as + seg) around them (full rows: as = 0, seg = RB)
    const int x0 = FULL ? 0 : blockIdx.z * P.TW, tw = FULL ? W : min(P.TW, W - x0);
    const int pxa = FULL ? 0 : max(x0 - R, 0), pxb = FULL ? W : min(x0 + tw + R, W);
    const int as = FULL ? 0 : (pxa * C) & ~15, seg = FULL ? RB : ((pxb * C + 15) & ~15) - as;
    constexpr bool full = FULL;
    uint8_t* s_src = sm;                                                                       // [(TR + 2R)][SP]
    uint16_t* s_v = reinterpret_cast<uint16_t*>(sm + (((size_t)(TR + 2 * R) * SP + 127) & ~(size_t)127));  // [C][TR][VP]
    const uint8_t* simg = src + (size_t)img * H * RB;
    uint8_t* dimg = dst + (size_t)img * H * RB;

    const int first = y0 - R, rows = nr + 2 * R;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(g_smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(&s_bar)), "r"((uint32_t)rows * seg) : "memory");
        if (full) {
            const int lo = max(first, 0), hi = min(first + rows - 1, H - 1);
            g_bulk_g2s(s_src + (size_t)(lo - first) * RB, simg + (size_t)lo * RB, (uint32_t)(hi - lo + 1) * RB, &s_bar);
            for (int t = 0; t < lo - first; ++t) g_bulk_g2s(s_src + (size_t)t * RB, simg + (size_t)g_refl101(first + t, H) * RB, RB, &s_bar);
            for (int t = hi - first + 1; t < rows; ++t) g_bulk_g2s(s_src + (size_t)t * RB, simg + (size_t)g_refl101(first + t, H) * RB, RB, &s_bar);
        }
    }
    if (!full && threadIdx.x < 32) {   // one bulk copy per staged row segment, spread over the lanes of warp 0
        __syncwarp();                  // the barrier is initialised and expects the bytes before any copy is issued
        for (int t = threadIdx.x; t < rows; t += 32)
            g_bulk_g2s(s_src + (size_t)t * SP, simg + (size_t)g_refl101(first + t, H) * RB + as, (uint32_t)seg, &s_bar);
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
    const int ncolw = SP >> 2;    // words per staged row (pitch)
    const int pstride = TR * VP;  // plane stride (elements)
    for (int cw = threadIdx.x; cw < (seg >> 2); cw += GT) {
        // destination of the 4 bytes of this column word: byte j = as + 4cw + b of the row -> pixel j / C, channel j % C,
        // plane position pixel - x0 + PAD; bytes of pixels outside [pxa, pxb) (alignment slack) have no destination
        int dofs[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = as + 4 * cw + b;
            const int p = (C == 1) ? j : j / 3;
            dofs[b] = (C == 1) ? (p - x0 + PAD) : ((j - p * 3) * pstride + p - x0 + PAD);
            if (!FULL && (p < pxa || p >= pxb)) dofs[b] = -1;
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
                        if (C == 1 && full) {
                            // 4 consecutive pixels of the single plane: (p0,p1), (p2,p3)
                            uint32_t* v32 = reinterpret_cast<uint32_t*>(vr + dofs[0]);
                            v32[0] = __byte_perm(al, ah, 0x5410);
                            v32[1] = __byte_perm(al, ah, 0x7632);
                        } else {
                            if (FULL || dofs[0] >= 0) vr[dofs[0]] = (uint16_t)(al & 0xFFFFu);
                            if (FULL || dofs[1] >= 0) vr[dofs[1]] = (uint16_t)(ah & 0xFFFFu);
                            if (FULL || dofs[2] >= 0) vr[dofs[2]] = (uint16_t)(al >> 16);
                            if (FULL || dofs[3] >= 0) vr[dofs[3]] = (uint16_t)(ah >> 16);
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
        uint16_t* v = s_v + (size_t)c * pstride + (size_t)row * VP + PAD;   // v[0] = pixel x0 of the tile
        if (side == 0) {
            if (x0 == 0) v[-r] = v[r];
        } else if (x0 + tw == W) {
            v[tw - 1 + r] = v[tw - 1 - r];
        }
    }
    __syncthreads();
    // ---------------- horizontal pass: 4 pixels per thread, NP dp2a per output value
    const int G = tw >> 2;
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
        uint8_t* orow = dimg + (size_t)(y0 + r) * RB + (size_t)x0 * C;
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
    const int RB = W * C;
    // rows per tile: as many as fit ~56 KB (four blocks per SM), at least 4, at most 64.  Full-width rows when that leaves at
    // least 16 rows per tile; wider images are cut into 256-pixel column tiles with R halo columns (each staged row segment
    // is the 16-byte aligned range around them), otherwise the 2R halo rows would dominate the work of a thin tile.
    int TW = W, SP = RB;
    auto smem_for = [&](int tr) { return (((size_t)(tr + 2 * R) * SP + 127) & ~(size_t)127) + (size_t)C * tr * (TW + 2 * PAD) * 2; };
    auto rows_for = [&]() {
        int tr = 64;
        while (tr > 4 && smem_for(tr) > 56 * 1024) tr -= 4;
        return tr;
    };
    int TR = rows_for();
    if (TR < 2 * R && W > 256 && C == 3) {   // the 2R halo rows would outnumber the rows of the tile
        TW = 256;
        SP = (((TW + 2 * R) * C + 30) + 15) & ~15;
        TR = rows_for();
    }
    if (smem_for(TR) > 200 * 1024) return 1;  // not for this kernel
    TR = min(TR, ((H + 3) / 4) * 4);
    const int VP = TW + 2 * PAD;
    GaussTmaParams P;
    memset(&P, 0, sizeof(P));
    P.H = H; P.W = W; P.TR = TR; P.VP = VP; P.TW = TW; P.SP = SP;
    for (int i = 0; i <= R; ++i) P.t[i] = taps[i];
    // even pairs: (t0,t1),(t2,t3),...,(t_{K-1},0); odd pairs: (0,t0),(t1,t2),...,(t_{K-2},t_{K-1})
    for (int k = 0; k <= R; ++k) {
        const int e0 = 2 * k, e1 = 2 * k + 1, o0 = 2 * k - 1, o1 = 2 * k;
        P.kev[k] = (uint32_t)(e0 < K ? taps[e0] : 0) | ((uint32_t)(e1 < K ? taps[e1] : 0) << 8);
        P.kod[k] = (uint32_t)(o0 >= 0 ? taps[o0] : 0) | ((uint32_t)(o1 < K ? taps[o1] : 0) << 8);
    }
    const size_t smem = smem_for(TR);
    static size_t attr_[LFX_MAX_DEVICES][2] = {{0}};
    size_t* attr = attr_[lfx_dev()];
    const int v = (TW == W) ? 1 : 0;
    if (smem > 40 * 1024 && smem > attr[v]) {  // static shared memory counts towards the 48 KB default limit
        cudaError_t e = v ? cudaFuncSetAttribute(k_gauss_tma<K, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(k_gauss_tma<K, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return 1;
        attr[v] = smem;
    }
    dim3 grid((H + TR - 1) / TR, B, (W + TW - 1) / TW);
    if (v)
        k_gauss_tma<K, C, true><<<grid, GT, smem, st>>>(src, dst, P);
    else
        k_gauss_tma<K, C, false><<<grid, GT, smem, st>>>(src, dst, P);
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
