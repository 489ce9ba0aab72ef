// Shared device/host helpers for libleafx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/leafx.h"

#define LFX_NUM_SMS 148  // B200
#define LFX_MAX_DEVICES 16

// Index of the current CUDA device (0 .. LFX_MAX_DEVICES-1), -1 on error: per-device caches (function attributes,
// LUT pointers) are arrays indexed by it.  Defined in lfx_api.cu.
int lfx_device_slot();
static inline int lfx_dev() { const int d = lfx_device_slot(); return d < 0 ? 0 : d; }

// ---- error plumbing (defined in lfx_api.cu) ---------------------------------------------------
void lfx_set_error(const char* fmt, ...);
int lfx_check_launch(const char* what);
bool lfx_ready();

#define LFX_REQUIRE(cond, code, ...)       \
    do {                                   \
        if (!(cond)) {                     \
            lfx_set_error(__VA_ARGS__);    \
            return (code);                 \
        }                                  \
    } while (0)

#define LFX_REQUIRE_READY() LFX_REQUIRE(lfx_ready(), LFX_ERR_CUDA, "lfx_init() has not succeeded on a CUDA device")

// ---- colour LUTs in global memory (uploaded by lfx_init) ----------------------------------------
struct LfxTables {
    int32_t sdiv[256];
    int32_t hdiv[256];
    uint16_t gtab[256];
    uint16_t ctab[3072];
};
const LfxTables* lfx_tables();  // device pointer, valid after lfx_init (lfx_api.cu)
const uint4* lfx_cat_lut();      // [3][256] byte-packed category flags of hist.py, uploaded by lfx_init (device pointer of the current device)

// Shared-memory copies used by the per-pixel device functions.
struct HsvLut {
    int32_t sdiv[256];
    int32_t hdiv[256];
};
struct LabLut {
    uint16_t gtab[256];
    uint16_t ctab[3072];
};

__device__ __forceinline__ void load_hsv_lut(HsvLut* s, const LfxTables* t) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s->sdiv[i] = t->sdiv[i];
        s->hdiv[i] = t->hdiv[i];
    }
}
__device__ __forceinline__ void load_lab_lut(LabLut* s, const LfxTables* t) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s->gtab[i] = t->gtab[i];
    // 3072 uint16 = 1536 words
    const uint32_t* src = reinterpret_cast<const uint32_t*>(t->ctab);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s->ctab);
    for (int i = threadIdx.x; i < 1536; i += blockDim.x) dst[i] = src[i];
}

// ---- per-pixel colour arithmetic (OpenCV 8-bit integer paths) ------------------
__device__ __forceinline__ int rgb2gray(int r, int g, int b) {
    return (r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15;
}

__device__ __forceinline__ void rgb2hsv(int r, int g, int b, const HsvLut* lut, int& h, int& s, int& v) {
    v = max(r, max(g, b));
    const int mn = min(r, min(g, b));
    const int d = v - mn;
    s = (d * lut->sdiv[v] + 2048) >> 12;
    // the hue numerator by selects: the lanes of a warp take different cases, a branch would run all of them in turn
    const int hr = g - b, hg = b - r + 2 * d, hb = r - g + 4 * d;
    int hh = (v == g) ? hg : hb;
    hh = (v == r) ? hr : hh;
    hh = (hh * lut->hdiv[d] + 2048) >> 12;  // arithmetic shift == floor, as in OpenCV
    h = hh < 0 ? hh + 180 : hh;
}

__device__ __forceinline__ int lfx_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void rgb2lab(int r, int g, int b, const LabLut* lut, int& L, int& A, int& Bv) {
    const int R = lut->gtab[r], G = lut->gtab[g], Bl = lut->gtab[b];
    const int fX = lut->ctab[lfx_descale(R * 1777 + G * 1541 + Bl * 778, 12)];
    const int fY = lut->ctab[lfx_descale(R * 871 + G * 2929 + Bl * 296, 12)];
    const int fZ = lut->ctab[lfx_descale(R * 73 + G * 448 + Bl * 3575, 12)];
    L = min(255, max(0, lfx_descale(296 * fY - 1336934, 15)));
    A = min(255, max(0, lfx_descale(500 * (fX - fY) + 128 * 32768, 15)));
    Bv = min(255, max(0, lfx_descale(200 * (fY - fZ) + 128 * 32768, 15)));
}

// ---- memory helpers ----------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w));
}

// Copy nbytes from global to shared; fast 16-byte path when both are 16-byte aligned.
__device__ __forceinline__ void block_load_bytes(uint8_t* smem, const uint8_t* g, int nbytes) {
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(smem)) & 15) == 0) {
        const int n16 = nbytes >> 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x)
            reinterpret_cast<uint4*>(smem)[i] = ld_stream16(g + (size_t)i * 16);
        for (int i = (n16 << 4) + threadIdx.x; i < nbytes; i += blockDim.x) smem[i] = g[i];
    } else {
        for (int i = threadIdx.x; i < nbytes; i += blockDim.x) smem[i] = g[i];
    }
}
__device__ __forceinline__ void block_store_bytes(uint8_t* g, const uint8_t* smem, int nbytes) {
    if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(smem)) & 15) == 0) {
        const int n16 = nbytes >> 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x)
            st_stream16(g + (size_t)i * 16, reinterpret_cast<const uint4*>(smem)[i]);
        for (int i = (n16 << 4) + threadIdx.x; i < nbytes; i += blockDim.x) g[i] = smem[i];
    } else {
        for (int i = threadIdx.x; i < nbytes; i += blockDim.x) g[i] = smem[i];
    }
}

static inline int lfx_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
