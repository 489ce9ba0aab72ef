// Library state, error reporting, LUT upload and the host-only helpers of the C ABI.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "lfx_common.cuh"
#include "lfx_tables.h"

static thread_local char t_err[512] = "";
// Per-device library state (a __device__ symbol has one instance per device; its address is looked up per device).
// Filled by lfx_init under a mutex; every other entry point only reads the slot of the CURRENT device.
__device__ LfxTables g_lfx_tables_storage;
__device__ uint4 g_lfx_cat_lut_storage[3 * 256];
struct LfxDeviceState {
    bool ready;
    const LfxTables* tables;
    const uint4* cat_lut;
};
static LfxDeviceState g_state[LFX_MAX_DEVICES];
static std::mutex g_init_mutex;

int lfx_device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= LFX_MAX_DEVICES) return -1;
    return d;
}

void lfx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int lfx_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        lfx_set_error("%s: %s", what, cudaGetErrorString(e));
        return LFX_ERR_CUDA;
    }
    return LFX_OK;
}

bool lfx_ready() {
    const int d = lfx_device_slot();
    return d >= 0 && g_state[d].ready;
}
const LfxTables* lfx_tables() {
    const int d = lfx_device_slot();
    return d >= 0 ? g_state[d].tables : nullptr;
}
const uint4* lfx_cat_lut() {
    const int d = lfx_device_slot();
    return d >= 0 ? g_state[d].cat_lut : nullptr;
}

// hist.py:38-65 categories and :248-256 hue ranges as per-channel byte flags.  Field k (byte k%4 of word k/4):
// 0 leaf (hist.py:188), 1..8 the categories, 9..13 the hue ranges.
static void build_cat_lut(uint32_t* lut /* [3][256][4] */) {
    memset(lut, 0, 3 * 256 * 16);
    for (int i = 0; i < 256; ++i) {
        const int h = i, s = i, v = i;
        const bool leaf_s = s > 10, leaf_v = v > 15 && v < 245;
        bool fh[14], fs[14], fv[14];
        fh[0] = true; fs[0] = leaf_s; fv[0] = leaf_v;
        fh[1] = h >= 35 && h <= 85; fs[1] = s >= 40; fv[1] = v >= 30;
        fh[2] = h >= 20 && h <= 40; fs[2] = s >= 25; fv[2] = v >= 30;
        fh[3] = h >= 15 && h <= 35; fs[3] = s >= 50; fv[3] = v >= 50;
        fh[4] = h <= 25 || h >= 160; fs[4] = s >= 30; fv[4] = v >= 20;
        fh[5] = (h >= 160 && h <= 180) || h <= 10; fs[5] = s >= 40; fv[5] = v >= 30;
        fh[6] = true; fs[6] = s >= 20; fv[6] = v <= 50;
        fh[7] = true; fs[7] = s <= 30; fv[7] = v >= 200;
        fh[8] = h >= 120 && h <= 160; fs[8] = s >= 20; fv[8] = true;
        fh[9] = h >= 35 && h <= 85; fs[9] = true; fv[9] = true;
        fh[10] = h >= 15 && h <= 35; fs[10] = true; fv[10] = true;
        fh[11] = h <= 15 || h >= 160; fs[11] = true; fv[11] = true;
        fh[12] = h >= 120 && h <= 160; fs[12] = true; fv[12] = true;
        fh[13] = h > 85 && h < 120; fs[13] = true; fv[13] = true;
        for (int k = 0; k < 14; ++k) {
            const uint32_t bit = 1u << (8 * (k & 3));
            if (fh[k]) lut[(0 * 256 + i) * 4 + (k >> 2)] |= bit;
            if (fs[k] && leaf_s) lut[(1 * 256 + i) * 4 + (k >> 2)] |= bit;
            if (fv[k] && leaf_v) lut[(2 * 256 + i) * 4 + (k >> 2)] |= bit;
        }
    }
}

extern "C" int lfx_version(void) { return 100; }

extern "C" const char* lfx_last_error(void) { return t_err; }

extern "C" int lfx_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        lfx_set_error("no CUDA device: %s (libleafx has no CPU fallback)", cudaGetErrorString(e));
        return LFX_ERR_CUDA;
    }
    LFX_REQUIRE(device >= 0 && device < n && device < LFX_MAX_DEVICES, LFX_ERR_ARG, "device %d out of range (0..%d)", device,
                (n < LFX_MAX_DEVICES ? n : LFX_MAX_DEVICES) - 1);
    std::lock_guard<std::mutex> lock(g_init_mutex);
    e = cudaSetDevice(device);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    LFX_REQUIRE(prop.major == 10, LFX_ERR_UNSUPPORTED,
                "libleafx is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    static LfxTables h;
    memcpy(h.sdiv, LFX_HOST_SDIV, sizeof(h.sdiv));
    memcpy(h.hdiv, LFX_HOST_HDIV, sizeof(h.hdiv));
    memcpy(h.gtab, LFX_HOST_GTAB, sizeof(h.gtab));
    memcpy(h.ctab, LFX_HOST_CTAB, sizeof(h.ctab));
    e = cudaMemcpyToSymbol(g_lfx_tables_storage, &h, sizeof(h));
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "LUT upload: %s", cudaGetErrorString(e));
    void* p = nullptr;
    e = cudaGetSymbolAddress(&p, g_lfx_tables_storage);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "cudaGetSymbolAddress: %s", cudaGetErrorString(e));
    static uint32_t hcat[3 * 256 * 4];
    build_cat_lut(hcat);
    e = cudaMemcpyToSymbol(g_lfx_cat_lut_storage, hcat, sizeof(hcat));
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "category LUT upload: %s", cudaGetErrorString(e));
    void* pc = nullptr;
    e = cudaGetSymbolAddress(&pc, g_lfx_cat_lut_storage);
    LFX_REQUIRE(e == cudaSuccess, LFX_ERR_CUDA, "cudaGetSymbolAddress: %s", cudaGetErrorString(e));
    g_state[device].tables = static_cast<const LfxTables*>(p);
    g_state[device].cat_lut = static_cast<const uint4*>(pc);
    g_state[device].ready = true;
    t_err[0] = 0;
    return LFX_OK;
}

// ---- host helpers --------------------------------------------------------------------------------

// OpenCV getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED, 8 fractional bits.
extern "C" int lfx_gauss_taps(int ksize, double sigma, int32_t* taps) {
    LFX_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= 31 && taps, LFX_ERR_ARG, "gauss taps: bad ksize %d", ksize);
    double k[31];
    static const double small[4][7] = {{1.0},
                                       {0.25, 0.5, 0.25},
                                       {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                       {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
    if (sigma <= 0 && ksize <= 7) {
        for (int i = 0; i < ksize; ++i) k[i] = small[ksize >> 1][i];
    } else {
        const double s = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
        const double scale2x = -0.5 / (s * s);
        double sum = 0;
        for (int i = 0; i < ksize; ++i) {
            const double x = i - (ksize - 1) * 0.5;
            k[i] = exp(scale2x * x * x);
            sum += k[i];
        }
        const double inv = 1.0 / sum;
        for (int i = 0; i < ksize; ++i) k[i] *= inv;
    }
    double err = 0.0;
    long long tot = 0;
    for (int i = 0; i < ksize / 2; ++i) {
        const double adj = k[i] * 256.0 + err;
        const long long v = (long long)nearbyint(adj);
        err = adj - (double)v;
        taps[i] = taps[ksize - 1 - i] = (int32_t)v;
        tot += v;
    }
    taps[ksize / 2] = (int32_t)(256 - 2 * tot);
    return LFX_OK;
}

static double lanczos_filter(double x) {
    // libImaging Resample.c: sinc_filter(x) * sinc_filter(x/3) on [-3, 3)
    if (-3.0 <= x && x < 3.0) {
        auto sinc = [](double v) {
            if (v == 0.0) return 1.0;
            v = v * M_PI;
            return sin(v) / v;
        };
        return sinc(x) * sinc(x / 3.0);
    }
    return 0.0;
}

extern "C" int lfx_lanczos_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return LFX_ERR_ARG;
    double filterscale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 3.0 * filterscale;
    return (int)ceil(support) * 2 + 1;
}

extern "C" int lfx_lanczos_table(int in_size, int out_size, int kstride, int32_t* bounds, int32_t* kk) {
    LFX_REQUIRE(in_size > 0 && out_size > 0 && bounds && kk, LFX_ERR_ARG, "lanczos table: bad arguments");
    const int ksize = lfx_lanczos_ksize(in_size, out_size);
    LFX_REQUIRE(kstride >= ksize, LFX_ERR_ARG, "lanczos table: kstride %d < ksize %d", kstride, ksize);
    const double in0 = 0.0, in1 = (double)in_size;
    double scale = (in1 - in0) / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 3.0 * filterscale;
    const double ss = 1.0 / filterscale;
    double* k = new double[ksize];
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int x = 0;
        for (; x < xmax; ++x) {
            const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; ++x) k[x] = 0.0;
        for (x = 0; x < kstride; ++x) {
            int32_t q = 0;
            if (x < ksize) {
                const double v = k[x] * (double)(1 << 22);
                q = (k[x] < 0) ? (int32_t)(-0.5 + v) : (int32_t)(0.5 + v);
            }
            kk[(size_t)xx * kstride + x] = q;
        }
        bounds[xx * 2 + 0] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    delete[] k;
    return ksize;
}
