/* leafx_jpeg.h -- C ABI of libleafx_jpeg.so: JPEG decode / encode on the GPU (nvJPEG) for the file boundary of the hot
 * path.
 *
 * Replaces, on the reference side (all paths relative to /root/reference):
 *   ImageLoader.load_pil_image / load_as_array   srcs/utils/image_utils.py:19-47    (Pillow decode, convert("RGB"))
 *   ImageLoader.save_pil_image                   srcs/utils/image_utils.py:49-59    (Pillow encode, quality=95)
 *   imwrite_bgr                                  srcs/cli/Transformation.py:196-205 (cv2.imwrite, default quality 95)
 *
 * Contract: bitstreams live in HOST memory (they come from / go to files), pixels live in DEVICE memory as the
 * uint8 [B,H,W,3] RGB batches every other libleafx entry point takes.  Parity is to JPEG tolerance only (the
 * inverse DCT and the chroma upsampling of nvJPEG and libjpeg-turbo differ by a few LSB; tests/test_gpu_jpeg.py
 * states the bounds).  Unlike libleafx.so this library allocates: nvJPEG owns its device / pinned scratch.
 * A separate shared object so that libleafx.so itself never depends on libnvjpeg.
 */
#ifndef LEAFX_JPEG_H
#define LEAFX_JPEG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* lfx_stream_t; /* cudaStream_t */

/* Creates the nvJPEG handle, decoder state and `threads` encoder states on CUDA device `device`.
 * backend: 0 = nvJPEG default, 1 = hybrid (host Huffman), 2 = GPU hybrid, 3 = hardware engine (falls back to 0 when the
 * engine is not available).  Returns 0, or a negative LFX_ERR_* code (lfx_jpeg_last_error() has the text). */
int lfx_jpeg_init(int device, int backend, int threads);
void lfx_jpeg_shutdown(void);
const char* lfx_jpeg_last_error(void);
/* backend actually in use after lfx_jpeg_init (same numbering), -1 before */
int lfx_jpeg_backend(void);

/* Header probe (host only): width, height, component count and chroma subsampling code (nvjpegChromaSubsampling_t). */
int lfx_jpeg_info(const uint8_t* jpeg, size_t len, int* width, int* height, int* components, int* subsampling);

/* Decodes B bitstreams into dst[B,H,W,3] (device, RGB interleaved; grey-scale and CMYK/YCCK sources are converted as
 * Pillow's convert("RGB") does for grey; 4-component files are rejected).  Images whose header size is not HxW, or
 * that fail to parse, get status[i] < 0 and their destination is left untouched; the call still returns 0.
 * status is a HOST array.  Work is enqueued on `stream`; the bitstreams must stay alive until the stream is synchronised. */
int lfx_jpeg_decode_batch(const uint8_t* const* jpeg, const size_t* len, uint8_t* dst, int B, int H, int W, int32_t* status,
                          lfx_stream_t stream);

/* Upper bound of one encoded HxW bitstream at this quality / subsampling. */
size_t lfx_jpeg_encode_bound(int H, int W, int quality, int subsampling);

/* Encodes src[B,H,W,3] (device, RGB interleaved) as baseline JPEG with the standard Huffman tables (what Pillow's
 * save(quality=q) and cv2.imwrite write): out is a HOST buffer of B slots of `cap` bytes, out_len[i] the stream length
 * (0 when the image failed).  subsampling: 420, 422 or 444 (Pillow / OpenCV default: 420).  The call waits for `stream`
 * first (the pixels must be final) and returns when every bitstream is in `out`. */
int lfx_jpeg_encode_batch(const uint8_t* src, int B, int H, int W, int quality, int subsampling, uint8_t* out, size_t cap,
                          size_t* out_len, lfx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
