// Fused core transform profile (BASELINE config 2: 5x5 Gaussian blur + make_mask + masked ROI letterbox +
// RGB/HSV/LAB histograms) -- one persistent thread block per image, everything between the first read
// of the image and the last output store stays in shared memory.
//
//   phase A  32-row source tiles arrive by TMA bulk copy (cp.async.bulk + mbarrier).  Vertical blur pass
//            on packed byte pairs -> de-interleaved 16-bit planes in shared memory; horizontal pass with
//            dp2a (two taps per instruction) -> blur rows to HBM.  Same tile: RGB->HSV(/Lab) threshold and
//            brown predicates -> bit planes (warp ballots).
//   phase B  mask_finish (lfx_maskops.cuh): _postprocess_mask, Otsu fallback, brown extension on the bit
//            planes; mask bytes + info to HBM.
//   phase C  tiles again (L2 hits): histograms / category counters of the masked pixels (shared-memory
//            atomics, LUT-packed counters), apply_mask(white) in place, INTER_AREA letterbox of the
//            bounding box, canvas rows leave by TMA bulk store.
//
// Reference: srcs/transform/filters/blur.py:72 (GaussianBlur 5x5), mask.py:548-582 (make_mask),
// roi.py:20-46, hist.py:22-67,188,248-256, utils/mask_utils.py:10-83.  Arithmetic identical to the
// stand-alone kernels (lfx_gauss.cu, lfx_mask.cu, lfx_roi.cu, lfx_color.cu), which remain the general path.
#include <stdlib.h>

#include <type_traits>

#include "lfx_maskops.cuh"

namespace {

constexpr int TR = 32;     // image rows per staged tile
constexpr int CHUNK = 4;   // rows of the zero buffer the letterbox bands are bulk-stored from
constexpr int NWARPS = MT / 32;

// Shared-memory byte offsets.  Fixed part: raw / brown / result planes and the colour LUTs; the rest is a
// union over the three phases.  constexpr so that the 256x256 instantiation folds every address.
struct Lay {
    int off_planes, off_hsv, off_lab;
    int off_src, off_v;                              // phase A (inside the union)
    int off_t, off_wbase, off_wbase2, off_runs, off_stage;  // phase B
    int off_src2, off_out, off_hist, off_cat, off_xt, off_yt;  // phase C (off_src shared with A; two tile buffers)
    int smem_bytes;
};
__host__ __device__ constexpr int lay_al(long long b) { return (int)((b + 127) & ~127ll); }
__host__ __device__ constexpr Lay make_lay(int H, int W, int RH, int RW) {
    Lay L{};
    const int NW = H * (W / 32), rb = W * 3;
    const int stage_rows = (STAGE_BYTES / rb) < 1 ? 1 : ((STAGE_BYTES / rb) > H ? H : (STAGE_BYTES / rb));
    int off = 0;
    L.off_planes = off; off += lay_al((long long)NW * 4 * 3);
    L.off_hsv = off; off += lay_al(sizeof(HsvLut));
    L.off_lab = off; off += lay_al(sizeof(LabLut));
    int a = off;
    L.off_src = a; a += lay_al((long long)(TR + 4) * rb + 16);
    L.off_v = a; a += lay_al((long long)3 * TR * (W + 4) * 2);
    int b = off;
    L.off_t = b; b += lay_al((long long)NW * 4 * 3);
    L.off_wbase = b; b += lay_al((long long)(NW + 1) * 4);
    L.off_wbase2 = b; b += lay_al((long long)(NW + 1) * 4);
    L.off_runs = b; b += lay_al((long long)RCAP_SMEM * 14);
    L.off_stage = b; b += lay_al((long long)stage_rows * rb);
    int c3 = L.off_src + lay_al((long long)(TR + 4) * rb + 16);
    L.off_src2 = c3; c3 += lay_al((long long)(TR + 4) * rb + 16);
    L.off_out = c3; c3 += lay_al((long long)CHUNK * RW * 3);
    // the 12 x 256 histogram lives in the raw / brown planes (dead in phase C) when they are big enough
    if ((long long)NW * 8 >= 12 * 256 * 4) {
        L.off_hist = L.off_planes;
    } else {
        L.off_hist = c3; c3 += lay_al(12 * 256 * 4);
    }
    L.off_cat = c3; c3 += lay_al(3 * 256 * 16);
    L.off_xt = c3; c3 += lay_al((long long)RW * 8);
    L.off_yt = c3; c3 += lay_al((long long)RH * 8);
    L.smem_bytes = a > b ? (a > c3 ? a : c3) : (b > c3 ? b : c3);
    return L;
}

struct CoreParams {
    MaskParams M;
    uint32_t K01, K23, K4_, K_0, K12, K34;  // horizontal taps paired for dp2a
    int t0, t1, t2;                         // symmetric vertical taps (t0 = taps[0] = taps[4], ...)
    int RH, RW;
    int need_lab_a;                         // Lab needed in phase A (lab strategy / lab brown)
    int hue_direct;                         // both hue ranges inside [0, 150): predicates compare the unshifted fixed-point values
    int g_lo12, g_span12, b_lo12, b_span12, b_smin12;   // range bounds << 12 for that form
    int timing;                             // debug: accumulate per-phase cycles (LFX_CORE_TIMING=1)
    Lay lay;
    unsigned long long ws_per_block;
};

// ---------------------------------------------------------------- TMA bulk copy + mbarrier (sm_90+ PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
// L2 eviction priorities: the image is read twice (phase A, then phase C ~200 us later) with ~1.6 MB of output
// streamed out per image in between -- keep it (evict_last) after the first read, release it (evict_first) on the second.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// Rows y0-2 .. y0+nr+1 of the image (BORDER_REFLECT_101 above/below) -> s_src, one mbarrier phase.
// Called by ONE thread.  H >= 3.
__device__ void issue_tile_load(const uint8_t* simg, uint8_t* s_src, uint64_t* bar, int y0, int nr, int H, int RB,
                                const void* extra_g, void* extra_s, uint32_t extra_bytes, uint64_t policy) {
    const int first = y0 - 2, rows = nr + 4;
    const int lo = max(first, 0), hi = min(first + rows - 1, H - 1);
    fence_async_smem();
    mbar_expect_tx(bar, (uint32_t)rows * RB + extra_bytes);
    bulk_g2s(s_src + (size_t)(lo - first) * RB, simg + (size_t)lo * RB, (uint32_t)(hi - lo + 1) * RB, bar, policy);
    for (int t = 0; t < lo - first; ++t) bulk_g2s(s_src + (size_t)t * RB, simg + (size_t)refl101(first + t, H) * RB, RB, bar, policy);
    for (int t = hi - first + 1; t < rows; ++t)
        bulk_g2s(s_src + (size_t)t * RB, simg + (size_t)refl101(first + t, H) * RB, RB, bar, policy);
    if (extra_bytes) bulk_g2s(extra_s, extra_g, extra_bytes, bar, l2_policy_evict_last());
}

// ---------------------------------------------------------------- phase A: blur
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// Vertical 5-tap pass on 4 pixels (12 bytes = 3 words) x VRS output rows per thread, packed 2x16-bit
// (sums <= 255*256 fit), results de-interleaved into the 16-bit channel planes s_v[c][row][x+2].
constexpr int VRS = 8;
__device__ __forceinline__ void vpass_item(const uint8_t* s_src, uint16_t* s_v, int g, int r0, int RB, int VP, int G,
                                           const CoreParams& P) {
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(s_src + r0 * RB) + 3 * g;
    const int rw = RB >> 2;
    uint32_t lo[5][3], hi[5][3];
#pragma unroll
    for (int step = 0; step < VRS + 4; ++step) {
        const int slot = step % 5;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t w = sp[step * rw + k];
            lo[slot][k] = prmt(w, 0u, 0x4240);  // bytes 0,2 -> 16-bit lanes
            hi[slot][k] = prmt(w, 0u, 0x4341);  // bytes 1,3
        }
        if (step >= 4) {
            // window rows step-4 .. step live in slots (step-4)%5 .. step%5
            const int s0 = (step + 1) % 5, s1 = (step + 2) % 5, s2 = (step + 3) % 5, s3 = (step + 4) % 5, s4 = slot;
            uint32_t l[3], h[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                l[k] = (lo[s0][k] + lo[s4][k]) * P.t0 + (lo[s1][k] + lo[s3][k]) * P.t1 + lo[s2][k] * P.t2;
                h[k] = (hi[s0][k] + hi[s4][k]) * P.t0 + (hi[s1][k] + hi[s3][k]) * P.t1 + hi[s2][k] * P.t2;
            }
            // bytes: w0 = (R0,G0,B0,R1) w1 = (G1,B1,R2,G2) w2 = (B2,R3,G3,B3);  l = even bytes, h = odd bytes
            const uint32_t R01 = prmt(l[0], h[0], 0x7610), R23 = prmt(l[1], h[2], 0x5432);
            const uint32_t G01 = prmt(h[0], l[1], 0x5410), G23 = prmt(h[1], l[2], 0x7632);
            const uint32_t B01 = prmt(l[0], h[1], 0x5432), B23 = prmt(l[2], h[2], 0x7610);
            const int orow = r0 + step - 4;
            uint32_t* vr = reinterpret_cast<uint32_t*>(s_v + orow * VP) + 2 * g + 1;  // element 4g+2
            const int cs = (TR * VP) >> 1;                                             // channel stride in words
            vr[0] = R01; vr[1] = R23;
            vr[cs] = G01; vr[cs + 1] = G23;
            vr[2 * cs] = B01; vr[2 * cs + 1] = B23;
            if (g == 0) {  // x = -2,-1 mirror x = 2,1
                vr[-1] = prmt(R01, R23, 0x3254);
                vr[cs - 1] = prmt(G01, G23, 0x3254);
                vr[2 * cs - 1] = prmt(B01, B23, 0x3254);
            }
            if (g == G - 1) {  // x = W,W+1 mirror x = W-2,W-3
                vr[2] = prmt(R01, R23, 0x3254);
                vr[cs + 2] = prmt(G01, G23, 0x3254);
                vr[2 * cs + 2] = prmt(B01, B23, 0x3254);
            }
        }
    }
}

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return prmt(prmt(a, b, 0x0062), prmt(c, d, 0x0062), 0x5410);  // byte 2 of each accumulator
}

// Horizontal 5-tap pass for 4 pixels of one row: 3 dp2a per output value, (v + 32768) >> 16 = byte 2.
__device__ __forceinline__ void hpass_item(const uint16_t* s_v, uint8_t* brow, int g, int r, int VP, const CoreParams& P) {
    uint32_t acc[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint2* q = reinterpret_cast<const uint2*>(s_v + (size_t)(c * TR + r) * VP) + g;
        const uint2 q0 = q[0], q1 = q[1];
        const uint32_t w0 = q0.x, w1 = q0.y, w2 = q1.x, w3 = q1.y;
        acc[c][0] = __dp2a_lo(w2, P.K4_, __dp2a_lo(w1, P.K23, __dp2a_lo(w0, P.K01, 32768u)));
        acc[c][1] = __dp2a_lo(w2, P.K34, __dp2a_lo(w1, P.K12, __dp2a_lo(w0, P.K_0, 32768u)));
        acc[c][2] = __dp2a_lo(w3, P.K4_, __dp2a_lo(w2, P.K23, __dp2a_lo(w1, P.K01, 32768u)));
        acc[c][3] = __dp2a_lo(w3, P.K34, __dp2a_lo(w2, P.K12, __dp2a_lo(w1, P.K_0, 32768u)));
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(brow) + 3 * g;
    // streaming stores (evict-first): outputs are written once and must not push the image out of L2
    __stcs(o + 0, pack4(acc[0][0], acc[1][0], acc[2][0], acc[0][1]));
    __stcs(o + 1, pack4(acc[1][1], acc[2][1], acc[0][2], acc[1][2]));
    __stcs(o + 2, pack4(acc[2][2], acc[0][3], acc[1][3], acc[2][3]));
}

// ---------------------------------------------------------------- phase C helpers
// cv::resize area-mode 2-tap coefficients for destination index d (src -> dst upscale); lfx_roi.cu.
// Packed as {source index, a | b << 16} (weights x2048 for s and s+1; a + b in [2047, 2049]).
__device__ __forceinline__ int2 area_tap2(int d, int src, int dst) {
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
    const int a = __float2int_rn(__fmul_rn(__fadd_rn(1.f, -fx), 2048.f));
    const int b = __float2int_rn(__fmul_rn(fx, 2048.f));
    return make_int2(sx, a | (b << 16));
}

struct Geo {
    int found, bx, by, bw, bh, nw, nh, ox, oy;
};

// streaming byte store of the low byte of a 32-bit register (st.u8 truncates: no `& 0xff` in front of it)
__device__ __forceinline__ void st_cs_u8(uint8_t* p, uint32_t v) {
    asm volatile("st.global.cs.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- the kernel
// S = true: compile-time 256x256 image, 256x256 canvas (PlantVillage, BASELINE configs 1-3); S = false: any
// W % 32 == 0, H*W <= 65536 shape with run-time geometry.
// F = true (only with S): the P1 configuration -- strategy hsv_h with both hue ranges below 150 (hue_direct), no Lab, no external
// candidate: the pixel pass carries no run-time strategy selection.
template <bool S, bool F = false>
__global__ void __launch_bounds__(MT, 2)
    k_core(const uint8_t* __restrict__ src, uint8_t* __restrict__ blur, uint8_t* __restrict__ mask, int32_t* __restrict__ info,
           uint8_t* __restrict__ roi, int32_t* __restrict__ hist9, int32_t* __restrict__ hsv3, int32_t* __restrict__ counters,
           int B, const CoreParams P, uint8_t* __restrict__ ws, const LfxTables* __restrict__ tab,
           const uint4* __restrict__ cat_lut, unsigned long long* __restrict__ ds_hist, const uint8_t* __restrict__ raw) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ int s_tmp[40];
    __shared__ unsigned long long s_best;
    __shared__ int s_bb[8];
    __shared__ int s_hist256[256];
    __shared__ int s_info[8];
    __shared__ int s_info2[8];
    __shared__ __align__(8) uint64_t s_bar2[2];
    uint64_t& s_bar = s_bar2[0];
    __shared__ int s_next;
    __shared__ Geo s_geo;
    __shared__ int s_dlo[TR + 2];  // first canvas row of each source tile (at most TR tiles: H <= TR*TR)
    __shared__ uint32_t s_cnt[16];

    const MaskParams& M = P.M;
    const int H = S ? 256 : M.H, W = S ? 256 : M.W, RH = S ? 256 : P.RH, RW = S ? 256 : P.RW;
    const int WPR = W >> 5, NW = H * WPR, RB = W * 3, VP = W + 4, G = W >> 2;
    const Lay L = S ? make_lay(256, 256, 256, 256) : P.lay;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ntiles = (H + TR - 1) / TR;

    using Geom = typename std::conditional<S, StaticGeom<256, 256>, DynGeom>::type;
    CtxT<Geom> c;
    if constexpr (!S) {
        ctx_init_geometry(c, H, W, WPR, NW, M.lastmask);
    } else {
        ctx_clear_hp(c);
    }
    c.s_tmp = s_tmp; c.s_best = &s_best; c.s_bb = s_bb; c.s_hist = s_hist256;
    c.rcap_glob = M.rcap_glob;
    c.rcap_smem = RCAP_SMEM;
    uint32_t* P0 = reinterpret_cast<uint32_t*>(sm + L.off_planes);
    uint32_t* PB = P0 + NW;
    uint32_t* PR = PB + NW;
    uint32_t* T1 = reinterpret_cast<uint32_t*>(sm + L.off_t);
    uint32_t* T2 = T1 + NW;
    uint32_t* T3 = T2 + NW;
    c.plane[0] = P0; c.plane[1] = PB; c.plane[2] = PR; c.plane[3] = T1; c.plane[4] = T2; c.plane[5] = T3;
    c.wbase = reinterpret_cast<int*>(sm + L.off_wbase);
    c.wbase2 = reinterpret_cast<int*>(sm + L.off_wbase2);
    c.sm_parent = reinterpret_cast<int*>(sm + L.off_runs);
    c.sm_geom = reinterpret_cast<uint32_t*>(sm + L.off_runs + RCAP_SMEM * 4);
    c.sm_acc = reinterpret_cast<int*>(sm + L.off_runs + RCAP_SMEM * 8);
    c.sm_ry = reinterpret_cast<uint16_t*>(sm + L.off_runs + RCAP_SMEM * 12);
    {
        uint8_t* gp = ws + 512 + (size_t)blockIdx.x * P.ws_per_block;
        auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
        c.gl_parent = reinterpret_cast<int*>(gp); gp += al((size_t)M.rcap_glob * 4);
        c.gl_geom = reinterpret_cast<uint32_t*>(gp); gp += al((size_t)M.rcap_glob * 4);
        c.gl_acc = reinterpret_cast<int*>(gp); gp += al((size_t)M.rcap_glob * 4);
        c.gl_ry = reinterpret_cast<uint16_t*>(gp);
    }
    if ((size_t)NW * 12 <= (size_t)RCAP_SMEM * 14) {
        c.hp[0] = P0; c.hp[1] = T3; c.hp[2] = reinterpret_cast<uint32_t*>(c.wbase);
        for (int k = 0; k < 3; ++k) c.hp[3 + k] = reinterpret_cast<uint32_t*>(c.sm_parent) + (size_t)k * NW;
    }
    HsvLut* s_hsv = reinterpret_cast<HsvLut*>(sm + L.off_hsv);
    LabLut* s_lab = reinterpret_cast<LabLut*>(sm + L.off_lab);
    uint8_t* s_src = sm + L.off_src;
    uint8_t* const s_srcA = sm + L.off_src;    // phase C tile buffers
    uint8_t* const s_srcB = sm + L.off_src2;
    uint16_t* s_v = reinterpret_cast<uint16_t*>(sm + L.off_v);
    uint8_t* s_stage = sm + L.off_stage;
    uint8_t* s_out = sm + L.off_out;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(sm + L.off_hist);
    const uint4* s_cat = reinterpret_cast<const uint4*>(sm + L.off_cat);
    int2* s_xt = reinterpret_cast<int2*>(sm + L.off_xt);
    int2* s_yt = reinterpret_cast<int2*>(sm + L.off_yt);

    load_hsv_lut(s_hsv, tab);
    load_lab_lut(s_lab, tab);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar2[0], 1);
        mbar_init(&s_bar2[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t par = 0, par1 = 0;   // phase parity of the two tile barriers
    const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
    int* work_counter = reinterpret_cast<int*>(ws);
    const size_t img_px = (size_t)H * W;
    const bool want_stats = hist9 || hsv3 || counters || ds_hist;

    // LFX_CORE_TIMING=1 (debug): per-phase SM-clock cycles summed over images into ws[64..]
    unsigned long long* tacc = P.timing ? reinterpret_cast<unsigned long long*>(ws + 64) : nullptr;
    long long tk0 = 0;
#define LFX_TICK(slot)                                                        \
    if (tacc && threadIdx.x == 0) {                                           \
        const long long now = clock64();                                      \
        atomicAdd(&tacc[slot], (unsigned long long)(now - tk0));              \
        tk0 = now;                                                            \
    }
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(work_counter, 1);
        __syncthreads();
        const int img = s_next;
        if (img >= B) break;
        const uint8_t* simg = src + (size_t)img * img_px * 3;
        if (tacc && threadIdx.x == 0) tk0 = clock64();

        // =========================================================== phase A
        if (threadIdx.x == 0) issue_tile_load(simg, s_src, &s_bar, 0, min(TR, H), H, RB, nullptr, nullptr, 0, pol_keep);
        for (int t = 0; t < ntiles; ++t) {
            const int y0 = t * TR, nr = min(TR, H - y0);
            mbar_wait(&s_bar, par);
            par ^= 1;
            // The vertical blur pass has G * ceil(nr / VRS) items (256 for a full 256-wide tile = 8 warps); the
            // warps it leaves idle take a head start of `extra` pixel-pass items each so both groups finish together.
            const int nvitems = blur ? G * ((nr + VRS - 1) / VRS) : 0;
            const int nvw = min(NWARPS, (nvitems + 31) >> 5);
            if (wid < nvw) {
                for (int item = threadIdx.x; item < nvitems; item += nvw * 32) {
                    const int strip = item / G, g = item - strip * G;
                    vpass_item(s_src, s_v, g, strip * VRS, RB, VP, G, P);
                }
            }
            const int nfree = NWARPS - nvw;                                       // warps without vertical-pass work
            const int npix = nr * WPR;
            const int extra = (nvw > 0 && nfree > 0) ? min(npix / nfree, 8) : 0;  // ~ one vpass item = 8 pixel items
            const int head = extra * nfree;
            // strategy 4: the candidate comes from a front-end kernel (inclusive / enhanced, lfx_raw_mask) as bytes
            const uint8_t* rawt = raw ? raw + (size_t)img * img_px + (size_t)y0 * W : nullptr;
            // one warp item = 32 consecutive pixels of a row: px = this lane's pixel in the staged tile, item = word index in the tile
            auto pixel_px = [&](const uint8_t* px, int item) {
                const int r = px[0], g = px[1], b = px[2];
                bool b0, b1;
                if constexpr (F) {
                    const int v = max(r, max(g, b)), d = v - min(r, min(g, b));
                    const int ts = d * s_hsv->sdiv[v] + 2048;
                    const int hr = g - b, hg = b - r + 2 * d, hb = r - g + 4 * d;      // selects, not branches: the lanes of a warp differ
                    int hh = (v == g) ? hg : hb;
                    hh = (v == r) ? hr : hh;
                    const int th = hh * s_hsv->hdiv[d] + 2048;
                    b0 = ((unsigned)(th - P.g_lo12) < (unsigned)P.g_span12) && (ts >= (40 << 12));   // mask.py:90
                    b1 = ((unsigned)(th - P.b_lo12) < (unsigned)P.b_span12) && (ts >= P.b_smin12) && (v <= M.cfg.brown_v_max);
                } else if (P.hue_direct) {
                    // H = th >> 12 (+180 when negative), S = ts >> 12 (rgb2hsv): lo <= H <= hi  <=>  lo << 12 <= th < (hi + 1) << 12
                    // for ranges below 150 (a negative th maps to H >= 150, so no wrap-around case), S >= T  <=>  ts >= T << 12:
                    // the two shifts and the hue fix-up are not needed for the predicates
                    const int v = max(r, max(g, b)), d = v - min(r, min(g, b));
                    const int ts = d * s_hsv->sdiv[v] + 2048;
                    const int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * d) : (r - g + 4 * d);
                    const int th = hh * s_hsv->hdiv[d] + 2048;
                    b0 = (M.cfg.strategy == 0) && ((unsigned)(th - P.g_lo12) < (unsigned)P.g_span12) && (ts >= (40 << 12));   // mask.py:90
                    b1 = ((unsigned)(th - P.b_lo12) < (unsigned)P.b_span12) && (ts >= P.b_smin12) && (v <= M.cfg.brown_v_max);
                } else if (P.need_lab_a) {
                    int h, s, v;
                    rgb2hsv(r, g, b, s_hsv, h, s, v);
                    int Ll, A, Bv;
                    rgb2lab(r, g, b, s_lab, Ll, A, Bv);
                    b0 = (M.cfg.strategy == 1) ? ((A <= 135) && (Bv >= 115) && (Bv <= 170))
                                               : ((M.cfg.strategy == 0) && (h >= M.cfg.green_lo) && (h <= M.cfg.green_hi) && (s >= 40));
                    b1 = M.cfg.use_lab_brown ? ((A >= M.cfg.lab_a_min) && (Bv >= M.cfg.lab_b_min))
                                             : ((h >= M.cfg.brown_lo) && (h <= M.cfg.brown_hi) && (s >= M.cfg.brown_s_min) &&
                                                (v <= M.cfg.brown_v_max));
                } else {
                    int h, s, v;
                    rgb2hsv(r, g, b, s_hsv, h, s, v);
                    // unsigned range checks: (h - lo) <= (hi - lo)
                    b0 = (M.cfg.strategy == 0) && ((unsigned)(h - M.cfg.green_lo) <= (unsigned)(M.cfg.green_hi - M.cfg.green_lo)) && (s >= 40);  // mask.py:90
                    b1 = ((unsigned)(h - M.cfg.brown_lo) <= (unsigned)(M.cfg.brown_hi - M.cfg.brown_lo)) &&
                         (s >= M.cfg.brown_s_min) && (v <= M.cfg.brown_v_max);
                }
                // strategy 4: the candidate comes from a front-end kernel (inclusive / enhanced, lfx_raw_mask) as bytes
                if constexpr (!F)
                    if (rawt) b0 = __ldg(rawt + item * 32 + lane) != 0;     // item * 32 = ry * W + w * 32
                const uint32_t m0 = __ballot_sync(0xffffffffu, b0);
                const uint32_t m1 = __ballot_sync(0xffffffffu, b1);
                if (lane == 0) {
                    const int idx = y0 * WPR + item;
                    P0[idx] = m0;
                    PB[idx] = m1;
                }
            };
            auto pixel_item = [&](int item) {
                int ry, w;
                split_index(c, item, ry, w);
                pixel_px(s_src + ((ry + 2) * RB + (w * 32 + lane) * 3), item);
            };
            // head start of the warps without vertical-pass work, then every warp: the second loop has a compile-time stride
            // (NWARPS items = NWARPS / WPR rows in the static instantiation), so its pixel address is a plain pointer increment
            if (wid >= nvw && extra != 0)
                for (int item = wid - nvw; item < head; item += nfree) pixel_item(item);
            if constexpr (S) {
                static_assert(NWARPS % 8 == 0, "a warp keeps its word column");
                int item = head + wid;
                const uint8_t* px = s_src + (((item >> 3) + 2) * 768 + ((item & 7) * 32 + lane) * 3);
                for (; item < npix; item += NWARPS, px += (NWARPS / 8) * 768) pixel_px(px, item);
            } else {
                for (int item = head + wid; item < npix; item += NWARPS) pixel_item(item);
            }
            __syncthreads();
            if (threadIdx.x == 0 && t + 1 < ntiles)
                issue_tile_load(simg, s_src, &s_bar, y0 + TR, min(TR, H - y0 - TR), H, RB, nullptr, nullptr, 0, pol_keep);
            if (blur) {
                uint8_t* bimg = blur + (size_t)img * img_px * 3 + (size_t)y0 * RB;
                for (int item = threadIdx.x; item < G * nr; item += MT) {
                    const int r = item / G, g = item - r * G;
                    hpass_item(s_v, bimg + r * RB, g, r, VP, P);
                }
                __syncthreads();
            }
        }

        LFX_TICK(0)
        // =========================================================== phase B
        c.status = 0;
        c.tacc = tacc;
        c.tk0 = tk0;
        if (M.cfg.strategy == 2 || M.cfg.strategy == 3) {
            // hsv_s (Otsu on S, 'light' unless dark_bg) / hsv_v_dark (Otsu on V, 'dark'), mask.py:76-84: the threshold needs
            // the whole histogram, so the candidate plane is built by two more passes over the (L2-resident) image
            const int chan = M.cfg.strategy == 2 ? 1 : 2;
            const bool dark = M.cfg.strategy == 2 ? (M.cfg.bg_dark != 0) : true;
            otsu_plane(simg, s_stage, s_hsv, P0, chan, dark, M, c);
            __syncthreads();
        }
        mask_finish(simg, s_stage, s_hsv, P0, PB, PR, T1, T2, T3, s_info, s_info2, M, c);
        plane_to_bytes16(PR, mask + (size_t)img * img_px, c);
        if (threadIdx.x < 8) {
            int v = s_info[threadIdx.x];
            if (threadIdx.x == 7) v = (c.status & 0xFF) | (v << 8);
            info[(size_t)img * 8 + threadIdx.x] = v;
        }
        if (!roi && !want_stats) {
            __syncthreads();
            continue;
        }

        tk0 = c.tk0;
        LFX_TICK(1)
        // =========================================================== phase C
        if (threadIdx.x == 0) {
            Geo gq;
            gq.found = s_info[0]; gq.bx = s_info[1]; gq.by = s_info[2]; gq.bw = s_info[3]; gq.bh = s_info[4];
            gq.nw = gq.nh = gq.ox = gq.oy = 0;
            if (roi && gq.found && gq.bw > 0 && gq.bh > 0) {
                // scale = min(W / max(w,1), H / max(h,1)); nw = max(int(w*scale),1)   (roi.py:35-36)
                const double sc = fmin(__ddiv_rn((double)RW, (double)max(gq.bw, 1)), __ddiv_rn((double)RH, (double)max(gq.bh, 1)));
                gq.nw = max((int)__dmul_rn((double)gq.bw, sc), 1);
                gq.nh = max((int)__dmul_rn((double)gq.bh, sc), 1);
                gq.ox = (RW - gq.nw) / 2;
                gq.oy = (RH - gq.nh) / 2;
            } else {
                gq.found = 0;
            }
            s_geo = gq;
        }
        for (int i = threadIdx.x; i < 12 * 256; i += MT) s_hist[i] = 0;
        if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0;
        for (int i = threadIdx.x; i < CHUNK * RW * 3 / 16; i += MT) reinterpret_cast<uint4*>(s_out)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();  // also: every phase-B reader of the union region is done
        const Geo geo = s_geo;
        if (threadIdx.x == 0)
            issue_tile_load(simg, s_src, &s_bar, 0, min(TR, H), H, RB, cat_lut, sm + L.off_cat, want_stats ? 3 * 256 * 16 : 0, pol_drop);
        uint8_t* rimg = roi ? roi + (size_t)img * RH * RW * 3 : nullptr;
        if (roi) {
            for (int i = threadIdx.x; i < geo.nw; i += MT) {
                int2 t2 = area_tap2(i, geo.bw, geo.nw);
                t2.x = (geo.bx + t2.x) * 3;  // byte offset of the left source pixel inside a staged row
                s_xt[i] = t2;
            }
            for (int i = threadIdx.x; i < geo.nh; i += MT) s_yt[i] = area_tap2(i, geo.bh, geo.nh);
            // letterbox bands above / below the resized box (the whole canvas when nothing was found)
            if (threadIdx.x == 0) {
                fence_async_smem();
                const int top = geo.found ? geo.oy : RH;
                for (int r0 = 0; r0 < top; r0 += CHUNK)
                    bulk_s2g(rimg + (size_t)r0 * RW * 3, s_out, (uint32_t)min(CHUNK, top - r0) * RW * 3);
                if (geo.found) {
                    // the side bands of the box rows are written by the ROI threads themselves (zero columns below)
                    for (int r0 = geo.oy + geo.nh; r0 < RH; r0 += CHUNK)
                        bulk_s2g(rimg + (size_t)r0 * RW * 3, s_out, (uint32_t)min(CHUNK, RH - r0) * RW * 3);
                }
                bulk_commit();
            }
            __syncthreads();
            // first canvas row whose upper source row lies in tile t (monotone in d): binary search
            if (threadIdx.x <= ntiles) {
                int lo = 0, hi = geo.nh;
                const int ylim = threadIdx.x * TR;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (geo.by + s_yt[mid].x >= ylim) hi = mid; else lo = mid + 1;
                }
                s_dlo[threadIdx.x] = lo;
            }
        }
        // ROI thread mapping for this image: one canvas column per thread over the FULL canvas width (columns left and
        // right of the resized box store zeros), MT / cols row strips (power of two)
        const int roi_cols = max(1, min(RW, MT));
        const int roi_sshift = 31 - __clz(max(1, MT / roi_cols));
        const int roi_strips = 1 << roi_sshift;
        const int roi_strip = threadIdx.x / roi_cols, roi_col0 = threadIdx.x - roi_strip * roi_cols;
        uint32_t cacc[4] = {0u, 0u, 0u, 0u};
        LFX_TICK(16)
        for (int t = 0; t < ntiles; ++t) {
            const int y0 = t * TR, nr = min(TR, H - y0);
            // double-buffered tiles: tile t sits in buffer t & 1; its successor is requested as soon as every thread
            // has left tile t-1 (the barrier below), so the copy overlaps this tile's work
            uint8_t* const s_src = (t & 1) ? s_srcB : s_srcA;
            if (t & 1) {
                mbar_wait(&s_bar2[1], par1);
                par1 ^= 1;
            } else {
                mbar_wait(&s_bar2[0], par);
                par ^= 1;
            }
            __syncthreads();  // s_dlo / taps visible (t == 0); tile t-1 fully consumed
            if (threadIdx.x == 0 && t + 1 < ntiles)
                issue_tile_load(simg, (t & 1) ? s_srcA : s_srcB, &s_bar2[(t + 1) & 1], y0 + TR, min(TR, H - y0 - TR), H, RB, nullptr, nullptr, 0, pol_drop);
            LFX_TICK(17)
            if (want_stats) {
                // contiguous rows per warp: every warp sees all word columns (the leaf sits in the middle ones)
                auto stats_px = [&](const uint8_t* px) {
                    const int r = px[0], g = px[1], b = px[2];
                    int h, s, v, Ll, A, Bv;
                    {   // rgb2hsv with the hue numerator chosen by selects (the lanes of a warp differ: no branch)
                        v = max(r, max(g, b));
                        const int d = v - min(r, min(g, b));
                        s = (d * s_hsv->sdiv[v] + 2048) >> 12;
                        const int hr = g - b, hg = b - r + 2 * d, hb = r - g + 4 * d;
                        int hh = (v == g) ? hg : hb;
                        hh = (v == r) ? hr : hh;
                        hh = (hh * s_hsv->hdiv[d] + 2048) >> 12;
                        h = hh < 0 ? hh + 180 : hh;
                    }
                    rgb2lab(r, g, b, s_lab, Ll, A, Bv);
                    atomicAdd(&s_hist[0 * 256 + r], 1u);
                    atomicAdd(&s_hist[1 * 256 + g], 1u);
                    atomicAdd(&s_hist[2 * 256 + b], 1u);
                    atomicAdd(&s_hist[3 * 256 + h], 1u);
                    atomicAdd(&s_hist[4 * 256 + s], 1u);
                    atomicAdd(&s_hist[5 * 256 + v], 1u);
                    atomicAdd(&s_hist[6 * 256 + Ll], 1u);
                    atomicAdd(&s_hist[7 * 256 + A], 1u);
                    atomicAdd(&s_hist[8 * 256 + Bv], 1u);
                    // leaf mask + 8 categories + 5 hue ranges (hist.py:188,38-65,248-256): per-channel LUTs of
                    // byte-packed 0/1 flags, AND = joint predicate, packed 8-bit counters (<= 128 px / thread)
                    const uint4 qh = s_cat[h], qs = s_cat[256 + s], qv = s_cat[512 + v];
                    const uint32_t q0 = qh.x & qs.x & qv.x;
                    cacc[0] += q0;
                    cacc[1] += qh.y & qs.y & qv.y;
                    cacc[2] += qh.z & qs.z & qv.z;
                    cacc[3] += qh.w & qs.w & qv.w;
                    // hsv3 = H/S/V histograms of the LEAF pixels (hist.py:188) = those of all masked pixels (planes 3..5)
                    // minus those of the masked non-leaf pixels: only the rare non-leaf pixel pays three more atomics
                    if (!(q0 & 1u)) {
                        atomicAdd(&s_hist[9 * 256 + h], 1u);
                        atomicAdd(&s_hist[10 * 256 + s], 1u);
                        atomicAdd(&s_hist[11 * 256 + v], 1u);
                    }
                };
                const int ipw = (nr * WPR + NWARPS - 1) / NWARPS;
                const int iend = min(nr * WPR, (wid + 1) * ipw);
                if constexpr (S) {
                    // rows are contiguous in the staged tile (768 = 8 words x 96 bytes): item -> pixel is one linear map
                    const uint8_t* px = s_src + 2 * 768 + (wid * ipw) * 96 + lane * 3;
                    const uint32_t* pm = PR + y0 * WPR + wid * ipw;
                    for (int k = 0; k < iend - wid * ipw; ++k, px += 96) {
                        const uint32_t m = pm[k];
                        if (m == 0u) continue;
                        if ((m >> lane) & 1u) stats_px(px);
                    }
                } else {
                    for (int item = wid * ipw; item < iend; ++item) {
                        const uint32_t m = PR[y0 * WPR + item];
                        if (m == 0u) continue;
                        if ((m >> lane) & 1u) {
                            int ry, w;
                            split_index(c, item, ry, w);
                            stats_px(s_src + ((ry + 2) * RB + (w * 32 + lane) * 3));
                        }
                    }
                }
            }
            if (roi && geo.found) {
                __syncthreads();
                LFX_TICK(18)
                // apply_mask(rgb, mask, "white") in place (Transformation.py:451) -- only the bounding box is ever
                // sampled: rows by .. by+bh-1 of this tile (+ the next row as lower tap), words covering bx .. bx+bw-1
                {
                    const int ra = max(y0, geo.by), rb2 = min(min(y0 + nr + 1, H), geo.by + geo.bh);
                    const int wa = geo.bx >> 5, nwd = ((geo.bx + geo.bw - 1) >> 5) - wa + 1;
                    for (int y = ra + wid; y < rb2; y += NWARPS) {
                        const uint32_t* prow = PR + y * WPR;
                        uint8_t* srow = s_src + ((y - y0 + 2) * RB + lane * 3);
                        for (int w = wa; w < wa + nwd; ++w) {
                            const uint32_t m = prow[w];
                            if (m == 0xFFFFFFFFu) continue;
                            if (!((m >> lane) & 1u)) {
                                uint8_t* px = srow + w * 96;
                                px[0] = 255; px[1] = 255; px[2] = 255;
                            }
                        }
                    }
                }
                __syncthreads();
                LFX_TICK(19)
                const int dA = s_dlo[t], dB = s_dlo[t + 1];
                const uint8_t* tile0 = s_src + (geo.by - y0 + 2) * RB;  // staged row of source row `by`
                // canvas rows dA .. dB-1 take their upper source row from this tile: one column per thread, the rows split
                // into roi_strips strips; pixels go straight to HBM (3 byte stores per pixel; L2 merges the sectors)
                if (roi_strip < roi_strips && dB > dA) {
                    const int per = (dB - dA + roi_strips - 1) >> roi_sshift;
                    const int da = dA + roi_strip * per, db = min(dB, da + per);
                    for (int cc = roi_col0; cc < RW; cc += roi_cols) {
                        uint8_t* o = rimg + ((size_t)(geo.oy + da) * RW + cc) * 3;
                        const int cx = cc - geo.ox;
                        if ((unsigned)cx >= (unsigned)geo.nw) {   // side band of the letterbox
                            for (int d = da; d < db; ++d, o += RW * 3) {
                                __stcs(o + 0, (uint8_t)0);
                                __stcs(o + 1, (uint8_t)0);
                                __stcs(o + 2, (uint8_t)0);
                            }
                            continue;
                        }
                        const int2 tx = s_xt[cx];
                        const uint32_t xa = tx.y & 0xFFFF, xb = (uint32_t)tx.y >> 16;
                        const uint8_t* colp = tile0 + tx.x;
                        int prev_s = -4;
                        uint32_t h0r = 0, h0g = 0, h0b = 0, h1r = 0, h1g = 0, h1b = 0;
                        for (int d = da; d < db; ++d, o += RW * 3) {
                            const int2 ty = s_yt[d];
                            if (ty.x != prev_s) {
                                const uint8_t* p = colp + ty.x * RB;
                                if (ty.x == prev_s + 1) {
                                    h0r = h1r; h0g = h1g; h0b = h1b;
                                } else {
                                    h0r = (p[0] * xa + p[3] * xb) >> 4;  // HResizeLinear x2048, >> 4 as in VResizeLinear
                                    h0g = (p[1] * xa + p[4] * xb) >> 4;
                                    h0b = (p[2] * xa + p[5] * xb) >> 4;
                                }
                                h1r = (p[RB] * xa + p[RB + 3] * xb) >> 4;
                                h1g = (p[RB + 1] * xa + p[RB + 4] * xb) >> 4;
                                h1b = (p[RB + 2] * xa + p[RB + 5] * xb) >> 4;
                                prev_s = ty.x;
                            }
                            // VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>; a + b <= 2049 keeps it in 0..255.
                            // (ya*h >> 16) as the high word of (ya << 16) * h: one IMAD.HI with the addend fused
                            const uint32_t ya = (uint32_t)ty.y << 16, yb = (uint32_t)ty.y & 0xFFFF0000u;
                            st_cs_u8(o + 0, (__umulhi(ya, h0r) + __umulhi(yb, h1r) + 2u) >> 2);
                            st_cs_u8(o + 1, (__umulhi(ya, h0g) + __umulhi(yb, h1g) + 2u) >> 2);
                            st_cs_u8(o + 2, (__umulhi(ya, h0b) + __umulhi(yb, h1b) + 2u) >> 2);
                        }
                    }
                }
            }
            LFX_TICK(20)
        }
        if (want_stats) {
            if (counters) {
#pragma unroll
                for (int k = 0; k < 14; ++k) {
                    const uint32_t v = __reduce_add_sync(0xffffffffu, (cacc[k >> 2] >> (8 * (k & 3))) & 0xFFu);
                    if (lane == 0 && v) atomicAdd(&s_cnt[k], v);
                }
            }
            __syncthreads();
            if (hist9)
                for (int i = threadIdx.x; i < 9 * 256; i += MT) hist9[(size_t)img * 9 * 256 + i] = (int)s_hist[i];
            // dataset-level colour histogram: reductions without a return value (RED), nothing waits for them
            if (ds_hist)
                for (int i = threadIdx.x; i < 9 * 256; i += MT) {
                    const uint32_t v = s_hist[i];
                    if (v) atomicAdd(&ds_hist[i], (unsigned long long)v);
                }
            if (hsv3)
                for (int i = threadIdx.x; i < 3 * 256; i += MT)
                    hsv3[(size_t)img * 3 * 256 + i] = (int)(s_hist[3 * 256 + i] - s_hist[9 * 256 + i]);
            if (counters && threadIdx.x < 16) counters[(size_t)img * 16 + threadIdx.x] = threadIdx.x < 14 ? (int)s_cnt[threadIdx.x] : 0;
        }
        if (threadIdx.x == 0) bulk_wait_read();  // s_out is reused as part of the phase-A/B union next image
        __syncthreads();
        LFX_TICK(2)
    }
#undef LFX_TICK
    if (threadIdx.x == 0) bulk_wait_all();
}

// general path only: ds[i] += sum over images of hist9[b][i]   (grid (9, chunks), 256 threads = one bin each)
__global__ void k_hist_accum(const int32_t* __restrict__ hist9, int B, unsigned long long* __restrict__ ds) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    unsigned long long acc = 0;
    for (int b = blockIdx.y; b < B; b += gridDim.y) acc += (unsigned long long)(uint32_t)hist9[(size_t)b * 9 * 256 + i];
    if (acc) atomicAdd(&ds[i], acc);
}

// ---------------------------------------------------------------- host side
// Shared-memory plan of the fused kernel; false when this shape / config takes the general path.
bool core_plan(int B, int H, int W, int RH, int RW, const lfx_mask_cfg* cfg, const int32_t* taps, CoreParams* out, int* per_sm) {
    if (W % 32 != 0 || W < 32 || H < 3 || (long long)H * W > 65536 || H > TR * TR) return false;
    if (!cfg || cfg->strategy < 0 || cfg->strategy > 4) return false;   // 2 / 3: Otsu on S / V, one extra pass in phase B; 4: external raw
    if ((RW * 3) % 16 != 0 || RW < W || RH < H || RW > 1024 || RH > 1024) return false;
    if (taps[0] != taps[4] || taps[1] != taps[3]) return false;
    for (int i = 0; i < 5; ++i)
        if (taps[i] < 0 || taps[i] > 255) return false;
    CoreParams& P = *out;
    memset(&P, 0, sizeof(P));
    MaskParams& M = P.M;
    M.cfg = *cfg;
    M.H = H; M.W = W; M.WPR = W / 32; M.NW = H * M.WPR;
    M.lastmask = 0xFFFFFFFFu;
    M.mode = 0;
    M.rcap_glob = H * (W + 2);  // ccl2 labels foreground and background runs together
    M.planes_in_smem = 1;
    const int rb = W * 3;
    M.stage_rows = max(1, min(H, STAGE_BYTES / rb));
    M.fp_morph = make_ellipse(cfg->morph_kernel);
    M.fp_brown = make_ellipse(cfg->brown_morph_kernel > 0 ? cfg->brown_morph_kernel : 3);
    M.fp_search = make_ellipse(20);
    M.search_is_e20 = is_ellipse20(M.fp_search) ? 1 : 0;
    P.t0 = taps[0]; P.t1 = taps[1]; P.t2 = taps[2];
    P.K01 = taps[0] | (taps[1] << 8); P.K23 = taps[2] | (taps[3] << 8); P.K4_ = taps[4];
    P.K_0 = taps[0] << 8; P.K12 = taps[1] | (taps[2] << 8); P.K34 = taps[3] | (taps[4] << 8);
    P.RH = RH; P.RW = RW;
    P.need_lab_a = (cfg->strategy == 1 || cfg->use_lab_brown) ? 1 : 0;
    {
        auto in150 = [](int lo, int hi) { return lo >= 0 && lo <= hi && hi < 150; };
        const bool green_ok = (cfg->strategy != 0) || in150(cfg->green_lo, cfg->green_hi);
        P.hue_direct = (!P.need_lab_a && green_ok && in150(cfg->brown_lo, cfg->brown_hi) && cfg->brown_s_min >= 0 && cfg->brown_s_min <= 255) ? 1 : 0;
        P.g_lo12 = cfg->green_lo << 12;
        P.g_span12 = (cfg->green_hi + 1 - cfg->green_lo) << 12;
        P.b_lo12 = cfg->brown_lo << 12;
        P.b_span12 = (cfg->brown_hi + 1 - cfg->brown_lo) << 12;
        P.b_smin12 = cfg->brown_s_min << 12;
    }
    P.lay = make_lay(H, W, RH, RW);
    auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    P.ws_per_block = al16((size_t)M.rcap_glob * 4) * 3 + al16((size_t)M.rcap_glob * 2);
    if (P.lay.smem_bytes > 226 * 1024) return false;
    *per_sm = max(1, min(2, (228 * 1024) / (P.lay.smem_bytes + 1024 + 1536)));  // + static shared + per-block reserve
    (void)B;
    return true;
}

}  // namespace

// scratch of the fused kernel: image queue counter + per-block run-table spill (H*(W+2) runs: ccl2 labels both kinds)
size_t lfx_core_workspace_bytes(int H, int W) {
    auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t rcap = (size_t)H * (W + 2);
    return 512 + (al16(rcap * 4) * 3 + al16(rcap * 2)) * (size_t)(2 * LFX_NUM_SMS) + 256;
}

extern "C" size_t lfx_pipeline_core_workspace(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return lfx_make_mask_workspace(B, H, W);   // already the maximum of the general and the fused requirement
}

// Fused kernel when the shape and config allow it: LFX_OK = launched, 1 = not eligible (use the general path),
// negative = error.  Also serves lfx_make_mask (blur / roi / stats pointers NULL: phases A-pixel + B only).
int lfx_core_try(const uint8_t* src, uint8_t* blur, uint8_t* mask, int32_t* info, uint8_t* roi, int32_t* hist9, int32_t* hsv3,
                 int32_t* counters, int B, int H, int W, int RH, int RW, double gaussian_sigma, const lfx_mask_cfg* cfg,
                 void* workspace, size_t workspace_bytes, cudaStream_t st, unsigned long long* ds_hist, const uint8_t* raw) {
    int rc;
    int32_t taps[31];
    CoreParams P;
    int per_sm = 1;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(roi) | reinterpret_cast<uintptr_t>(mask)) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(blur) % 4 == 0);
    const int mk = cfg->morph_kernel, bk = cfg->brown_morph_kernel;
    const bool morph_ok = mk >= 1 && mk <= 19 && (mk & 1) && bk >= 1 && bk <= 19 && (bk & 1);
    if ((cfg->strategy == 4) != (raw != nullptr)) return 1;
    if (!(aligned && morph_ok && lfx_gauss_taps(5, gaussian_sigma, taps) == LFX_OK && core_plan(B, H, W, RH, RW, cfg, taps, &P, &per_sm)))
        return 1;
    const uint4* cat_lut = lfx_cat_lut();
    LFX_REQUIRE(cat_lut != nullptr, LFX_ERR_CUDA, "pipeline_core: lfx_init() has not run on the current device");
    const int grid = max(1, min(B, LFX_NUM_SMS * per_sm));
    const size_t need = 512 + (size_t)P.ws_per_block * grid;
    LFX_REQUIRE(workspace && workspace_bytes >= need, LFX_ERR_WORKSPACE, "pipeline_core: workspace %zu < %zu bytes", workspace_bytes, need);
    const bool s256 = (H == 256 && W == 256 && RH == 256 && RW == 256);
    const bool fast = s256 && P.hue_direct && cfg->strategy == 0 && raw == nullptr;
    const int variant = fast ? 2 : (s256 ? 1 : 0);
    static int attr_[LFX_MAX_DEVICES][3] = {{0}};
    int* attr = attr_[lfx_dev()];
    if (P.lay.smem_bytes > attr[variant]) {
        const void* fn = fast ? (const void*)k_core<true, true> : s256 ? (const void*)k_core<true> : (const void*)k_core<false>;
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, P.lay.smem_bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "pipeline_core smem attr (%d bytes): %s", P.lay.smem_bytes, cudaGetErrorString(e));
        attr[variant] = P.lay.smem_bytes;
    }
    cudaError_t e = cudaMemsetAsync(workspace, 0, 512, st);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "pipeline_core memset: %s", cudaGetErrorString(e));
    static const bool timing = getenv("LFX_CORE_TIMING") && atoi(getenv("LFX_CORE_TIMING")) > 0;
    P.timing = timing ? 1 : 0;
    if (fast)
        k_core<true, true><<<grid, MT, P.lay.smem_bytes, st>>>(src, blur, mask, info, roi, hist9, hsv3, counters, B, P,
                                                              (uint8_t*)workspace, lfx_tables(), cat_lut, ds_hist, raw);
    else if (s256)
        k_core<true><<<grid, MT, P.lay.smem_bytes, st>>>(src, blur, mask, info, roi, hist9, hsv3, counters, B, P,
                                                        (uint8_t*)workspace, lfx_tables(), cat_lut, ds_hist, raw);
    else
        k_core<false><<<grid, MT, P.lay.smem_bytes, st>>>(src, blur, mask, info, roi, hist9, hsv3, counters, B, P,
                                                         (uint8_t*)workspace, lfx_tables(), cat_lut, ds_hist, raw);
    if (timing) {  // debug only: synchronises and prints the phase split
        unsigned long long t[32] = {0};
        cudaStreamSynchronize(st);
        cudaMemcpy(t, (uint8_t*)workspace + 64, sizeof(unsigned long long) * 24, cudaMemcpyDeviceToHost);
        double tb = (double)t[1];
        for (int k = 3; k < 16; ++k) tb += (double)t[k];
        double tc = (double)t[2];
        for (int k = 16; k < 24; ++k) tc += (double)t[k];
        const double tot = (double)t[0] + tb + tc;
        fprintf(stderr, "[lfx] k_core B=%d phase cycles/image: A %.0f (%.1f%%)  B %.0f (%.1f%%)  C %.0f (%.1f%%)\n", B, t[0] / (double)B,
                100.0 * t[0] / tot, tb / (double)B, 100.0 * tb / tot, tc / (double)B, 100.0 * tc / tot);
        fprintf(stderr, "[lfx]   C split: prelude %.0f  tma-wait %.0f  stats %.0f  mask-paint %.0f  roi %.0f  tail %.0f\n", t[16] / (double)B,
                t[17] / (double)B, t[18] / (double)B, t[19] / (double)B, t[20] / (double)B, t[2] / (double)B);
        fprintf(stderr, "[lfx]   B split: fill-ccl4 %.0f  close/open %.0f  largest#1 %.0f  dilate20x2 %.0f  brown-morph %.0f  brown-ccl8 %.0f  largest#2 %.0f  tail %.0f\n",
                t[3] / (double)B, t[4] / (double)B, t[5] / (double)B, t[6] / (double)B, t[7] / (double)B, t[8] / (double)B, t[9] / (double)B, t[1] / (double)B);
        fprintf(stderr, "[lfx]   ccl2 (both largest_external calls; NOT included in largest#N above): count %.0f  scan+extract %.0f  union %.0f  flatten %.0f\n",
                t[15] / (double)B, t[10] / (double)B, t[11] / (double)B, t[12] / (double)B);
    }
    return lfx_check_launch("pipeline_core(fused)");
}

extern "C" int lfx_pipeline_core(const uint8_t* src, uint8_t* blur, uint8_t* mask, int32_t* info, uint8_t* roi,
                                 int32_t* hist9, int32_t* hsv3, int32_t* counters, int B, int H, int W, int RH, int RW,
                                 double gaussian_sigma, const lfx_mask_cfg* cfg, void* workspace, size_t workspace_bytes,
                                 int64_t* dataset_hist9, const uint8_t* raw, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && mask && info && cfg, LFX_ERR_ARG, "pipeline_core: NULL argument");
    LFX_REQUIRE(B > 0 && H > 0 && W > 0, LFX_ERR_ARG, "pipeline_core: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* ds_hist = reinterpret_cast<unsigned long long*>(dataset_hist9);
    LFX_REQUIRE(!ds_hist || hist9, LFX_ERR_ARG, "pipeline_core: dataset_hist9 needs hist9");
    LFX_REQUIRE((cfg->strategy == 4) == (raw != nullptr), LFX_ERR_ARG, "pipeline_core: strategy 4 <=> a raw candidate is given");
    int rc = lfx_core_try(src, blur, mask, info, roi, hist9, hsv3, counters, B, H, W, RH, RW, gaussian_sigma, cfg, workspace,
                          workspace_bytes, st, ds_hist, raw);
    if (rc <= 0) return rc;

    // ---- general path: the stand-alone kernels back to back
    if (blur) {
        rc = lfx_gauss_u8(src, blur, B, H, W, 3, 5, gaussian_sigma, stream);
        if (rc) return rc;
    }
    rc = lfx_make_mask(src, raw, mask, info, B, H, W, cfg, workspace, workspace_bytes, stream);
    if (rc) return rc;
    if (roi) {
        rc = lfx_roi_letterbox(src, mask, info, roi, B, H, W, RH, RW, stream);
        if (rc) return rc;
    }
    if (hist9 || hsv3 || counters) {
        cudaError_t e = cudaSuccess;
        if (hist9) e = cudaMemsetAsync(hist9, 0, (size_t)B * 9 * 256 * 4, st);
        if (e == cudaSuccess && hsv3) e = cudaMemsetAsync(hsv3, 0, (size_t)B * 3 * 256 * 4, st);
        if (e == cudaSuccess && counters) e = cudaMemsetAsync(counters, 0, (size_t)B * 16 * 4, st);
        LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "pipeline_core memset: %s", cudaGetErrorString(e));
        rc = lfx_color_stats(src, mask, hist9, hsv3, counters, B, H, W, stream);
        if (rc) return rc;
        if (ds_hist) {
            k_hist_accum<<<dim3(9, 16), 256, 0, st>>>(hist9, B, ds_hist);
            rc = lfx_check_launch("pipeline_core(dataset histogram)");
            if (rc) return rc;
        }
    }
    return LFX_OK;
}
