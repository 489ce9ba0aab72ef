// make_mask / _postprocess_mask on the GPU (srcs/transform/filters/mask.py:53-69,335-411,548-582).
//
// One persistent thread block per image.  Masks are bit-packed (32 px / word) and live in shared
// memory for images up to ~512x512 (global scratch otherwise); morphology is word-parallel funnel
// shifts; connected components are run-based union-find (runs found with bit tricks, ids by
// prefix-sum of run starts, unions with atomicMin); findContours -> max(contourArea) ->
// drawContours(filled) is restated with local counts:
//   outside  = 4-connected background touching the border,  filled = ~outside,
//   2*contourArea(component) = 2N - (P - Q1) - 2  over 8-connected components of `filled`.
#include "lfx_maskops.cuh"

size_t lfx_core_workspace_bytes(int H, int W);
int lfx_core_try(const uint8_t* src, uint8_t* blur, uint8_t* mask, int32_t* info, uint8_t* roi, int32_t* hist9, int32_t* hsv3,
                 int32_t* counters, int B, int H, int W, int RH, int RW, double gaussian_sigma, const lfx_mask_cfg* cfg,
                 void* workspace, size_t workspace_bytes, cudaStream_t st, unsigned long long* ds_hist, const uint8_t* raw);

namespace {

__global__ void __launch_bounds__(MT, 2) k_make_mask(const uint8_t* __restrict__ src, const uint8_t* __restrict__ raw,
                                                     uint8_t* __restrict__ mask, int32_t* __restrict__ info, int B,
                                                     const MaskParams P, uint8_t* __restrict__ ws,
                                                     const LfxTables* __restrict__ tab) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ int s_tmp[40];
    __shared__ unsigned long long s_best;
    __shared__ int s_bb[8];
    __shared__ int s_hist[256];
    __shared__ int s_info[8];
    __shared__ int s_info2[8];

    Ctx c;
    ctx_init_geometry(c, P.H, P.W, P.WPR, P.NW, P.lastmask);
    c.s_tmp = s_tmp; c.s_best = &s_best; c.s_bb = s_bb; c.s_hist = s_hist;
    c.rcap_glob = P.rcap_glob;
    c.rcap_smem = RCAP_SMEM;
    // ---- carve shared memory
    uint8_t* sp = sm;
    uint8_t* gp = ws + (size_t)blockIdx.x * P.ws_per_block;
    auto take = [](uint8_t*& p, size_t bytes) {
        uint8_t* r = p;
        p += (bytes + 15) & ~(size_t)15;
        return r;
    };
    for (int k = 0; k < NPLANES; ++k)
        c.plane[k] = reinterpret_cast<uint32_t*>(P.planes_in_smem ? take(sp, (size_t)P.NW * 4) : take(gp, (size_t)P.NW * 4));
    c.wbase = reinterpret_cast<int*>(P.planes_in_smem ? take(sp, (size_t)(P.NW + 1) * 4) : take(gp, (size_t)(P.NW + 1) * 4));
    c.sm_parent = reinterpret_cast<int*>(take(sp, RCAP_SMEM * 4));
    c.sm_geom = reinterpret_cast<uint32_t*>(take(sp, RCAP_SMEM * 4));
    c.sm_acc = reinterpret_cast<int*>(take(sp, RCAP_SMEM * 4));
    c.sm_ry = reinterpret_cast<uint16_t*>(take(sp, RCAP_SMEM * 2));
    c.gl_parent = reinterpret_cast<int*>(take(gp, (size_t)P.rcap_glob * 4));
    c.gl_geom = reinterpret_cast<uint32_t*>(take(gp, (size_t)P.rcap_glob * 4));
    c.gl_acc = reinterpret_cast<int*>(take(gp, (size_t)P.rcap_glob * 4));
    c.gl_ry = reinterpret_cast<uint16_t*>(take(gp, (size_t)P.rcap_glob * 2));
    uint8_t* s_stage = take(sp, (size_t)P.stage_rows * P.W * 3);
    HsvLut* s_hsv = reinterpret_cast<HsvLut*>(take(sp, sizeof(HsvLut)));
    LabLut* s_lab = reinterpret_cast<LabLut*>(take(sp, sizeof(LabLut)));
    const bool need_lab = (P.mode != 1) && (P.cfg.strategy == 1 || P.cfg.use_lab_brown);
    load_hsv_lut(s_hsv, tab);
    if (need_lab) load_lab_lut(s_lab, tab);
    __syncthreads();

    uint32_t *P0 = c.plane[0], *PB = c.plane[1], *PR = c.plane[2], *T1 = c.plane[3], *T2 = c.plane[4], *T3 = c.plane[5];
    const size_t img_px = (size_t)P.H * P.W;
    // scratch planes of dilate_ellipse20: P0 and T3 are dead by then, the run tables are rebuilt by the next ccl
    if (P.planes_in_smem && (size_t)P.NW * 12 <= (size_t)RCAP_SMEM * 14) {
        c.hp[0] = P0; c.hp[1] = T3; c.hp[2] = reinterpret_cast<uint32_t*>(c.wbase);
        for (int k = 0; k < 3; ++k) c.hp[3 + k] = reinterpret_cast<uint32_t*>(c.sm_parent) + (size_t)k * P.NW;
    }

    for (int img = blockIdx.x; img < B; img += gridDim.x) {
        c.status = 0;
        const uint8_t* simg = src ? src + img * img_px * 3 : nullptr;
        // ---- raw candidate
        if (P.mode == 1 || P.cfg.strategy == 4) {
            bytes_to_plane(raw + img * img_px, P0, c);
            if (P.mode == 0 && P.cfg.extend_brown) {
                // brown predicate still comes from the RGB image; pixel_pass<0> leaves P0 alone
                // when strategy > 1
                __syncthreads();
                pixel_pass<0>(simg, s_stage, s_hsv, s_lab, P0, PB, 0, 0, false, P, c);
            }
        } else if (P.cfg.strategy <= 1) {
            pixel_pass<0>(simg, s_stage, s_hsv, s_lab, P0, PB, 0, 0, false, P, c);
        } else {
            // hsv_s (Otsu on S, 'light' unless dark_bg) / hsv_v_dark (Otsu on V, 'dark')
            const int chan = P.cfg.strategy == 2 ? 1 : 2;
            const bool dark = P.cfg.strategy == 2 ? (P.cfg.bg_dark != 0) : true;
            otsu_plane(simg, s_stage, s_hsv, P0, chan, dark, P, c);
            if (P.cfg.extend_brown) pixel_pass<0>(simg, s_stage, s_hsv, s_lab, P0, PB, 0, 0, false, P, c);
        }
        __syncthreads();
        if (P.mode == 2) {   // raw candidate only (lfx_strategy_raw): the plane goes out as it is
            plane_to_bytes(P0, mask + img * img_px, c);
            __syncthreads();
            continue;
        }

        mask_finish(simg, s_stage, s_hsv, P0, PB, PR, T1, T2, T3, s_info, s_info2, P, c);
        plane_to_bytes(PR, mask + img * img_px, c);
        if (threadIdx.x < 8) {
            int v = s_info[threadIdx.x];
            if (threadIdx.x == 7) v = (c.status & 0xFF) | (v << 8);  // status | start_x << 8
            info[(size_t)img * 8 + threadIdx.x] = v;
        }
        __syncthreads();
    }
}

struct Plan {
    MaskParams P;
    size_t smem;
    size_t ws_per_block;
    int grid;
};

int make_plan(int B, int H, int W, const lfx_mask_cfg* cfg, int mode, Plan* out) {
    MaskParams& P = out->P;
    memset(&P, 0, sizeof(P));
    if (cfg) P.cfg = *cfg;
    P.H = H; P.W = W; P.WPR = (W + 31) / 32; P.NW = H * P.WPR;
    P.lastmask = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xFFFFFFFFu;
    P.mode = mode;
    P.rcap_glob = H * ((W + 1) / 2);
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t plane_bytes = al((size_t)P.NW * 4) * NPLANES + al((size_t)(P.NW + 1) * 4);
    P.planes_in_smem = plane_bytes <= 72 * 1024;
    const int rb = W * 3;
    // staged RGB rows of the pixel passes: 6 KB next to shared-memory planes; large images (planes in global scratch) have the
    // shared memory to stage 48 KB at a time (two blocks per SM still fit) -- 16 rows of a 1024-wide image instead of 2
    P.stage_rows = max(1, (P.planes_in_smem ? STAGE_BYTES : 48 * 1024) / rb);
    if (P.stage_rows > H) P.stage_rows = H;
    size_t smem = al(RCAP_SMEM * 4) * 3 + al(RCAP_SMEM * 2) + al((size_t)P.stage_rows * rb) + al(sizeof(HsvLut)) + al(sizeof(LabLut));
    if (P.planes_in_smem) smem += plane_bytes;
    size_t wsb = al((size_t)P.rcap_glob * 4) * 3 + al((size_t)P.rcap_glob * 2);
    if (!P.planes_in_smem) wsb += plane_bytes;
    P.ws_per_block = wsb;
    out->smem = smem;
    out->ws_per_block = wsb;
    const int per_sm = smem <= 100 * 1024 ? 2 : 1;
    out->grid = max(1, min(B, LFX_NUM_SMS * per_sm));
    if (smem > 220 * 1024) return LFX_ERR_UNSUPPORTED;
    return LFX_OK;
}

int launch(const uint8_t* src, const uint8_t* raw, uint8_t* mask, int32_t* info, int B, int H, int W,
           const lfx_mask_cfg* cfg, int mode, void* ws, size_t ws_bytes, cudaStream_t st) {
    LFX_REQUIRE(H > 0 && W > 0 && H <= 65535 && W <= 65535 && B >= 0, LFX_ERR_ARG, "make_mask: bad shape");
    if (B == 0) return LFX_OK;
    Plan pl;
    int rc = make_plan(B, H, W, cfg, mode, &pl);
    LFX_REQUIRE(rc == LFX_OK, rc, "make_mask: image %dx%d needs %zu bytes of shared memory", H, W, pl.smem);
    LFX_REQUIRE(ws && ws_bytes >= pl.ws_per_block * pl.grid, LFX_ERR_WORKSPACE, "make_mask: workspace %zu < %zu bytes",
                ws_bytes, pl.ws_per_block * pl.grid);
    if (mode == 2) {
        LFX_REQUIRE(src && cfg->strategy >= 0 && cfg->strategy <= 3, LFX_ERR_ARG, "strategy_raw: strategy %d has no raw candidate here",
                    cfg->strategy);
        pl.P.cfg.extend_brown = 0;
    }
    if (mode == 0) {
        const int mk = cfg->morph_kernel, bk = cfg->brown_morph_kernel;
        LFX_REQUIRE(mk >= 1 && mk <= 19 && (mk & 1) && bk >= 1 && bk <= 19 && (bk & 1), LFX_ERR_UNSUPPORTED,
                    "make_mask: morph kernels must be odd <= 19 (got %d, %d)", mk, bk);
        LFX_REQUIRE(cfg->strategy >= 0 && cfg->strategy <= 4, LFX_ERR_UNSUPPORTED, "make_mask: strategy %d", cfg->strategy);
        LFX_REQUIRE(cfg->strategy != 4 || raw, LFX_ERR_ARG, "make_mask: strategy 4 needs a raw mask");
        LFX_REQUIRE(src, LFX_ERR_ARG, "make_mask: src is NULL");
    }
    pl.P.fp_morph = make_ellipse(pl.P.cfg.morph_kernel);
    pl.P.fp_brown = make_ellipse(pl.P.cfg.brown_morph_kernel > 0 ? pl.P.cfg.brown_morph_kernel : 3);
    pl.P.fp_search = make_ellipse(20);
    pl.P.search_is_e20 = is_ellipse20(pl.P.fp_search) ? 1 : 0;
    static size_t attr_[LFX_MAX_DEVICES] = {0};
    size_t& attr = attr_[lfx_dev()];
    if (pl.smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(k_make_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "make_mask smem attr: %s", cudaGetErrorString(e));
        attr = pl.smem;
    }
    k_make_mask<<<pl.grid, MT, pl.smem, st>>>(src, raw, mask, info, B, pl.P, (uint8_t*)ws, lfx_tables());
    return lfx_check_launch("make_mask");
}

}  // namespace

extern "C" size_t lfx_make_mask_workspace(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    Plan pl;
    make_plan(B, H, W, nullptr, 0, &pl);
    const size_t general = pl.ws_per_block * (size_t)min(B, LFX_NUM_SMS * 2) + 256;
    const size_t fused = lfx_core_workspace_bytes(H, W);
    return general > fused ? general : fused;
}

extern "C" int lfx_make_mask(const uint8_t* src, const uint8_t* raw, uint8_t* mask, int32_t* info, int B, int H, int W,
                             const lfx_mask_cfg* cfg, void* workspace, size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(mask && info && cfg, LFX_ERR_ARG, "make_mask: NULL argument");
    if (src && ((!raw && (cfg->strategy == 0 || cfg->strategy == 1)) || (raw && cfg->strategy == 4)) && B > 0 && H > 0 && W > 0) {
        // threshold strategies and external candidates on fused-kernel shapes: k_core without its blur / ROI / statistics phases
        const int rc = lfx_core_try(src, nullptr, mask, info, nullptr, nullptr, nullptr, nullptr, B, H, W, H, W, 1.5, cfg, workspace,
                                    workspace_bytes, (cudaStream_t)stream, nullptr, raw);
        if (rc <= 0) return rc;
    }
    return launch(src, raw, mask, info, B, H, W, cfg, 0, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int lfx_postprocess_mask(const uint8_t* raw, uint8_t* mask, int32_t* info, int B, int H, int W, int fill_size,
                                    int morph_kernel, void* workspace, size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(raw && mask && info, LFX_ERR_ARG, "postprocess_mask: NULL argument");
    LFX_REQUIRE(morph_kernel >= 1 && morph_kernel <= 19 && (morph_kernel & 1), LFX_ERR_UNSUPPORTED,
                "postprocess_mask: morph kernel %d", morph_kernel);
    lfx_mask_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.strategy = 4;
    cfg.fill_size = fill_size;
    cfg.morph_kernel = morph_kernel;
    cfg.brown_morph_kernel = 3;
    return launch(nullptr, raw, mask, info, B, H, W, &cfg, 1, workspace, workspace_bytes, (cudaStream_t)stream);
}

// Raw candidate of one threshold strategy (mask.py:72-106: hsv_s / hsv_v_dark by Otsu, hsv_h, lab), no post-processing:
// what _build_mask_candidates hands to _postprocess_mask, one entry per strategy ("auto" scores several of them).
extern "C" int lfx_strategy_raw(const uint8_t* src, uint8_t* raw, int B, int H, int W, const lfx_mask_cfg* cfg, void* workspace,
                                size_t workspace_bytes, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(src && raw && cfg, LFX_ERR_ARG, "strategy_raw: NULL argument");
    return launch(src, nullptr, raw, nullptr, B, H, W, cfg, 2, workspace, workspace_bytes, (cudaStream_t)stream);
}
