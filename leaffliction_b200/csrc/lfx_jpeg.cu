// libleafx_jpeg.so: JPEG decode / encode on the GPU through nvJPEG (include/leafx_jpeg.h).
//
// The reference reads every image with Pillow (srcs/utils/image_utils.py:19-47) and writes every output with Pillow
// (image_utils.py:49-59, quality 95) or cv2.imwrite (srcs/cli/Transformation.py:196-205); per image that is ~1 ms decode
// + ~1 ms encode on a host core, more than any arithmetic on the path.  Here a batch of
// bitstreams is decoded straight into the uint8 [B,H,W,3] device batch the kernels take (nvjpegDecodeBatched: Huffman
// on the GPU or on the hardware engine), and result batches are encoded from device memory by `threads` encoder
// states side by side, one CUDA stream each; only bitstreams cross PCIe.
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/leafx_jpeg.h"

namespace {

constexpr int ERR_ARG = -1, ERR_CUDA = -2, ERR_UNSUPPORTED = -3;

thread_local char t_err[512] = "";
char g_err[512] = "";
std::mutex g_err_mu;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(g_err_mu);
    memcpy(g_err, t_err, sizeof(g_err));
}

#define JREQUIRE(cond, code, ...)  \
    do {                           \
        if (!(cond)) {             \
            set_error(__VA_ARGS__); \
            return (code);         \
        }                          \
    } while (0)

struct Encoder {
    nvjpegEncoderState_t state = nullptr;
    nvjpegEncoderParams_t params = nullptr;
    cudaStream_t stream = nullptr;
    int quality = -1, subsampling = -1;
};

struct State {
    bool ready = false;
    int device = -1, backend = -1;
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t dec = nullptr;
    int dec_batch = 0;
    std::vector<Encoder> enc;
    std::mutex mu;   // one decode / encode call at a time
} G;

nvjpegBackend_t backend_of(int b) {
    switch (b) {
        case 1: return NVJPEG_BACKEND_HYBRID;
        case 2: return NVJPEG_BACKEND_GPU_HYBRID;
        case 3: return NVJPEG_BACKEND_HARDWARE;
        default: return NVJPEG_BACKEND_DEFAULT;
    }
}

bool css_of(int code, nvjpegChromaSubsampling_t* out) {
    switch (code) {
        case 420: *out = NVJPEG_CSS_420; return true;
        case 422: *out = NVJPEG_CSS_422; return true;
        case 444: *out = NVJPEG_CSS_444; return true;
        default: return false;
    }
}

int setup_encoder(Encoder& e, int quality, int subsampling) {
    if (e.quality == quality && e.subsampling == subsampling) return 0;
    nvjpegChromaSubsampling_t css;
    JREQUIRE(css_of(subsampling, &css), ERR_ARG, "jpeg_encode: subsampling %d (420, 422 or 444)", subsampling);
    nvjpegStatus_t s = nvjpegEncoderParamsSetQuality(e.params, quality, e.stream);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsSetSamplingFactors(e.params, css, e.stream);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsSetOptimizedHuffman(e.params, 0, e.stream);   // standard tables, as Pillow / OpenCV
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsSetEncoding(e.params, NVJPEG_ENCODING_BASELINE_DCT, e.stream);
    JREQUIRE(s == NVJPEG_STATUS_SUCCESS, ERR_CUDA, "jpeg_encode: nvjpegEncoderParamsSet* failed (%d)", (int)s);
    e.quality = quality;
    e.subsampling = subsampling;
    return 0;
}

}  // namespace

extern "C" const char* lfx_jpeg_last_error(void) { return t_err[0] ? t_err : g_err; }
extern "C" int lfx_jpeg_backend(void) { return G.ready ? G.backend : -1; }

extern "C" void lfx_jpeg_shutdown(void) {
    std::lock_guard<std::mutex> lk(G.mu);
    if (!G.handle) return;
    for (Encoder& e : G.enc) {
        if (e.params) nvjpegEncoderParamsDestroy(e.params);
        if (e.state) nvjpegEncoderStateDestroy(e.state);
        if (e.stream) cudaStreamDestroy(e.stream);
    }
    G.enc.clear();
    if (G.dec) nvjpegJpegStateDestroy(G.dec);
    nvjpegDestroy(G.handle);
    G.handle = nullptr;
    G.dec = nullptr;
    G.dec_batch = 0;
    G.ready = false;
}

extern "C" int lfx_jpeg_init(int device, int backend, int threads) {
    JREQUIRE(backend >= 0 && backend <= 3, ERR_ARG, "jpeg_init: backend %d", backend);
    JREQUIRE(threads >= 1 && threads <= 64, ERR_ARG, "jpeg_init: threads %d (1..64)", threads);
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    JREQUIRE(ce == cudaSuccess && device >= 0 && device < count, ERR_CUDA, "jpeg_init: no CUDA device %d (%s); there is no CPU fallback",
             device, ce == cudaSuccess ? "out of range" : cudaGetErrorString(ce));
    if (G.ready && G.device == device && (int)G.enc.size() == threads) return 0;
    lfx_jpeg_shutdown();
    std::lock_guard<std::mutex> lk(G.mu);
    ce = cudaSetDevice(device);
    JREQUIRE(ce == cudaSuccess, ERR_CUDA, "jpeg_init: cudaSetDevice: %s", cudaGetErrorString(ce));
    nvjpegStatus_t s = nvjpegCreateEx(backend_of(backend), nullptr, nullptr, 0, &G.handle);
    if (s != NVJPEG_STATUS_SUCCESS && backend != 0) {   // e.g. no hardware engine on this part
        backend = 0;
        s = nvjpegCreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, 0, &G.handle);
    }
    JREQUIRE(s == NVJPEG_STATUS_SUCCESS, ERR_CUDA, "jpeg_init: nvjpegCreateEx failed (%d)", (int)s);
    s = nvjpegJpegStateCreate(G.handle, &G.dec);
    JREQUIRE(s == NVJPEG_STATUS_SUCCESS, ERR_CUDA, "jpeg_init: nvjpegJpegStateCreate failed (%d)", (int)s);
    G.enc.resize(threads);
    for (Encoder& e : G.enc) {
        ce = cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking);
        JREQUIRE(ce == cudaSuccess, ERR_CUDA, "jpeg_init: stream: %s", cudaGetErrorString(ce));
        s = nvjpegEncoderStateCreate(G.handle, &e.state, e.stream);
        if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncoderParamsCreate(G.handle, &e.params, e.stream);
        JREQUIRE(s == NVJPEG_STATUS_SUCCESS, ERR_CUDA, "jpeg_init: encoder state (%d)", (int)s);
    }
    G.device = device;
    G.backend = backend;
    G.ready = true;
    return 0;
}

extern "C" int lfx_jpeg_info(const uint8_t* jpeg, size_t len, int* width, int* height, int* components, int* subsampling) {
    JREQUIRE(G.ready, ERR_CUDA, "lfx_jpeg_init() has not succeeded on a CUDA device");
    JREQUIRE(jpeg && len > 0, ERR_ARG, "jpeg_info: empty bitstream");
    int nc = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css;
    nvjpegStatus_t s = nvjpegGetImageInfo(G.handle, jpeg, len, &nc, &css, w, h);
    JREQUIRE(s == NVJPEG_STATUS_SUCCESS, ERR_ARG, "jpeg_info: not a decodable JPEG (%d)", (int)s);
    if (width) *width = w[0];
    if (height) *height = h[0];
    if (components) *components = nc;
    if (subsampling) *subsampling = (int)css;
    return 0;
}

extern "C" int lfx_jpeg_decode_batch(const uint8_t* const* jpeg, const size_t* len, uint8_t* dst, int B, int H, int W,
                                     int32_t* status, lfx_stream_t stream) {
    JREQUIRE(G.ready, ERR_CUDA, "lfx_jpeg_init() has not succeeded on a CUDA device");
    if (B == 0) return 0;
    JREQUIRE(jpeg && len && dst && status && B > 0 && H > 0 && W > 0, ERR_ARG, "jpeg_decode_batch: bad argument");
    std::lock_guard<std::mutex> lk(G.mu);
    std::vector<const unsigned char*> data;
    std::vector<size_t> lens;
    std::vector<nvjpegImage_t> outs;
    data.reserve(B); lens.reserve(B); outs.reserve(B);
    for (int i = 0; i < B; ++i) {
        status[i] = ERR_ARG;
        if (!jpeg[i] || len[i] == 0) continue;
        int nc = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t css;
        if (nvjpegGetImageInfo(G.handle, jpeg[i], len[i], &nc, &css, w, h) != NVJPEG_STATUS_SUCCESS) continue;
        if (w[0] != W || h[0] != H) { status[i] = ERR_UNSUPPORTED; continue; }
        if (nc != 1 && nc != 3) { status[i] = ERR_UNSUPPORTED; continue; }
        nvjpegImage_t o;
        memset(&o, 0, sizeof(o));
        o.channel[0] = dst + (size_t)i * H * W * 3;
        o.pitch[0] = (size_t)W * 3;
        data.push_back(jpeg[i]);
        lens.push_back(len[i]);
        outs.push_back(o);
        status[i] = 0;
    }
    const int n = (int)data.size();
    if (n == 0) return 0;
    if (n != G.dec_batch) {
        unsigned hc = std::thread::hardware_concurrency();
        nvjpegStatus_t s = nvjpegDecodeBatchedInitialize(G.handle, G.dec, n, (int)(hc ? (hc > 16 ? 16 : hc) : 4), NVJPEG_OUTPUT_RGBI);
        JREQUIRE(s == NVJPEG_STATUS_SUCCESS, ERR_CUDA, "jpeg_decode_batch: nvjpegDecodeBatchedInitialize(%d) failed (%d)", n, (int)s);
        G.dec_batch = n;
    }
    nvjpegStatus_t s = nvjpegDecodeBatched(G.handle, G.dec, data.data(), lens.data(), outs.data(), (cudaStream_t)stream);
    if (s != NVJPEG_STATUS_SUCCESS) {
        // one bad stream fails the whole batched call: fall back to one image at a time so that the others survive
        G.dec_batch = 0;
        int k = 0;
        for (int i = 0; i < B; ++i) {
            if (status[i] != 0) continue;
            nvjpegStatus_t si = nvjpegDecode(G.handle, G.dec, data[k], lens[k], NVJPEG_OUTPUT_RGBI, &outs[k], (cudaStream_t)stream);
            if (si != NVJPEG_STATUS_SUCCESS) status[i] = ERR_ARG;
            ++k;
        }
    }
    return 0;
}

extern "C" size_t lfx_jpeg_encode_bound(int H, int W, int quality, int subsampling) {
    (void)quality;
    (void)subsampling;
    if (H <= 0 || W <= 0) return 0;
    // worst case of baseline Huffman coding is ~2 bytes per coefficient + headers; nvjpegEncodeGetBufferSize reports the same order
    return (size_t)H * W * 3 * 2 + 4096;
}

extern "C" int lfx_jpeg_encode_batch(const uint8_t* src, int B, int H, int W, int quality, int subsampling, uint8_t* out, size_t cap,
                                     size_t* out_len, lfx_stream_t stream) {
    JREQUIRE(G.ready, ERR_CUDA, "lfx_jpeg_init() has not succeeded on a CUDA device");
    if (B == 0) return 0;
    JREQUIRE(src && out && out_len && B > 0 && H > 0 && W > 0 && cap > 0, ERR_ARG, "jpeg_encode_batch: bad argument");
    JREQUIRE(quality >= 1 && quality <= 100, ERR_ARG, "jpeg_encode_batch: quality %d", quality);
    std::lock_guard<std::mutex> lk(G.mu);
    cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
    JREQUIRE(ce == cudaSuccess, ERR_CUDA, "jpeg_encode_batch: %s", cudaGetErrorString(ce));
    for (Encoder& e : G.enc) {
        const int rc = setup_encoder(e, quality, subsampling);
        if (rc) return rc;
    }
    const int T = (int)std::min<size_t>(G.enc.size(), (size_t)B);
    std::atomic<int> next(0), failed(0);
    const int device = G.device;
    auto work = [&](int t) {
        cudaSetDevice(device);
        Encoder& e = G.enc[t];
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= B) break;
            out_len[i] = 0;
            nvjpegImage_t im;
            memset(&im, 0, sizeof(im));
            im.channel[0] = const_cast<unsigned char*>(src + (size_t)i * H * W * 3);
            im.pitch[0] = (size_t)W * 3;
            nvjpegStatus_t s = nvjpegEncodeImage(G.handle, e.state, e.params, &im, NVJPEG_INPUT_RGBI, W, H, e.stream);
            size_t n = 0;
            if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegEncodeRetrieveBitstream(G.handle, e.state, nullptr, &n, e.stream);
            if (s == NVJPEG_STATUS_SUCCESS && n <= cap) {
                s = nvjpegEncodeRetrieveBitstream(G.handle, e.state, out + (size_t)i * cap, &n, e.stream);
                if (s == NVJPEG_STATUS_SUCCESS && cudaStreamSynchronize(e.stream) == cudaSuccess) {
                    out_len[i] = n;
                    continue;
                }
            }
            failed.fetch_add(1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (std::thread& x : th) x.join();
    if (failed.load()) set_error("jpeg_encode_batch: %d of %d images failed (out_len 0)", failed.load(), B);
    return 0;
}
