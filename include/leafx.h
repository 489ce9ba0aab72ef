/*
 * leafx.h -- C ABI of libleafx.so: the B200 (sm_100a) implementation of leaffliction's
 * per-image preprocessing hot path.
 *
 * The reference (Kiripiro/leaffliction) is pure Python and has no FFI of its own; its boundary
 * for this path is the Python call surface listed below.  Each entry point here is what a
 * ctypes binding for the cited reference function calls (see INTEGRATION.md for the stub a
 * maintainer would add on the reference side).
 *
 * Conventions
 *   - Plain pointers and sizes only.  Unless a parameter is documented "host", every pointer is
 *     a DEVICE pointer; images are contiguous uint8 [B,H,W,3] (HWC, RGB), masks uint8 [B,H,W]
 *     (0/255), histograms int32 [B,9,256].
 *   - `stream` is a cudaStream_t passed as void*.  No entry point allocates device memory or
 *     synchronises with the host; scratch comes from the caller (`*_workspace` queries).
 *   - Return value: 0 on success, negative LFX_ERR_* otherwise; lfx_last_error() returns a
 *     thread-local message.  There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with LFX_ERR_CUDA.
 *   - The augment entry points take `src_index` (device int32[B], may be NULL) and `n_src`: output image i is
 *     computed from source image src_index[i] of the n_src images at `src` (NULL = image i of B).  This is the
 *     balancer's `random.choice(source_images)` (dataset_balancer.py:116) without a gather copy.
 *   - Per-image parameters (angles, coefficients, crop boxes ...) are drawn on the host by the
 *     Python shim with the same `random` / `np.random` calls, in the same order, as the
 *     reference (image_augmenter.py:23,36,48,77,79,101,105,106,121,127), then uploaded.
 */
#ifndef LEAFX_H
#define LEAFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFX_OK 0
#define LFX_ERR_ARG (-1)
#define LFX_ERR_CUDA (-2)
#define LFX_ERR_UNSUPPORTED (-3)
#define LFX_ERR_WORKSPACE (-4)

typedef void* lfx_stream_t;

/* ---- library ------------------------------------------------------------------------------ */
int lfx_version(void);
/* Select `device`, upload the colour LUTs, raise the dynamic shared-memory limits. */
int lfx_init(int device);
const char* lfx_last_error(void);

/* ---- augmentations: srcs/preprocessing/image_augmenter.py ------------------------------------ */

/* ImageAugmenter.flip (image_augmenter.py:20-31; PIL transpose :24,:26).
 * mode[B]: 0 = FLIP_LEFT_RIGHT, 1 = FLIP_TOP_BOTTOM. */
int lfx_flip(const uint8_t* src, uint8_t* dst, int B, int H, int W, const int32_t* mode,
             const int32_t* src_index, int n_src, lfx_stream_t stream);

/* ImageAugmenter.rotate (image_augmenter.py:33-42; PIL rotate NEAREST, expand, white fill :37).
 * params[B][8] = {a0,a1,a2,a3,a4,a5 (16.16 fixed point, libImaging affine_fixed), nw, nh}; 16-byte aligned.
 * Output image i is written at dst + i*dst_image_stride as [nh_i, nw_i, 3] contiguous;
 * pixels that map outside the source get `fill` in every channel. */
int lfx_rotate_nn(const uint8_t* src, uint8_t* dst, int64_t dst_image_stride, int B, int H, int W,
                  const int32_t* params, int fill, const int32_t* src_index, int n_src,
                  lfx_stream_t stream);

/* ImageAugmenter.skew / .shear (image_augmenter.py:44-71, :73-94; PIL transform BICUBIC :61-66,:84-89).
 * coef[B][8] = PIL's inverse-map coefficients a..h (fp64); perspective[B] != 0 selects the
 * PERSPECTIVE divide.  Bit-exact with Pillow's fp64 arithmetic. */
int lfx_warp_bicubic(const uint8_t* src, uint8_t* dst, int B, int H, int W, const double* coef,
                     const int32_t* perspective, const int32_t* src_index, int n_src,
                     lfx_stream_t stream);

/* HOST helpers: Pillow's 8-bit Lanczos coefficient tables (libImaging Resample.c
 * precompute_coeffs + normalize_coeffs_8bpc).  lfx_lanczos_ksize returns the tap count;
 * lfx_lanczos_table fills host arrays bounds[out_size][2] = {first, count} and
 * kk[out_size][kstride] (2^22 fixed point, zero padded), kstride >= ksize. */
int lfx_lanczos_ksize(int in_size, int out_size);
int lfx_lanczos_table(int in_size, int out_size, int kstride, int32_t* bounds, int32_t* kk);

/* ImageAugmenter.crop (image_augmenter.py:96-114; PIL crop + resize LANCZOS :108-109) and
 * ImageTransforms.resize_image (image_utils.py:109-114).
 * box[B][4] = {left, top, crop_w, crop_h}; every image is resized to [OH,OW].
 * tab_bounds / tab_kk: device copies of concatenated tables (lfx_lanczos_table layout, common
 * kstride); tab_off[B][4] = {x_bounds_row, x_ksize, y_bounds_row, y_ksize} where *_row is the
 * first row of that image's table inside the concatenation.
 * If dst_f32 is non-NULL the kernel also writes float32 [B,OH,OW,3] = u8 / 255.0f
 * (ImageTransforms.normalize_array, image_utils.py:117-130; sequence.py:84-88). */
int lfx_crop_lanczos(const uint8_t* src, uint8_t* dst, float* dst_f32, int B, int H, int W,
                     const int32_t* box, int OH, int OW, const int32_t* tab_bounds,
                     const int32_t* tab_kk, int kstride, const int32_t* tab_off,
                     const int32_t* src_index, int n_src, lfx_stream_t stream);

/* ImageAugmenter.distortion (image_augmenter.py:116-133): x = src + noise (uint8 wrap-around,
 * :121-124), per-channel ImageOps.autocontrast(cutoff) (:127).
 * noise[B,H,W,3]: the uint8-cast Gaussian noise; cut[B] = int(H*W*cutoff // 100) computed on the
 * host; hist_ws: device scratch int32 [B][3][256]. */
int lfx_distort(const uint8_t* src, const uint8_t* noise, uint8_t* dst, int B, int H, int W,
                const int32_t* cut, int32_t* hist_ws, const int32_t* src_index, int n_src,
                lfx_stream_t stream);

/* The noise of ImageAugmenter.distortion generated on the device: out[b, 0..n) =
 * np.random.normal(loc, scale, n).astype(np.uint8) after np.random.seed(seeds[b]) -- NumPy's legacy MT19937 +
 * polar-method stream that ImageAugmenter(seed) seeds (image_augmenter.py:16-18,121-123).  seeds[B]: device
 * uint32 (non-zero: seed 0 leaves the reference unseeded).  Feeds lfx_distort's `noise`. */
int lfx_legacy_normal_u8(const uint32_t* seeds, uint8_t* out, int B, int n, double loc, double scale,
                         lfx_stream_t stream);

/* Host-side (no device work, callable without a GPU): the parameters one balancing task draws --
 * `_process_single_transformation` (dataset_balancer.py:201-207) builds ImageAugmenter(seed), which seeds
 * Python's `random` (image_augmenter.py:16-18), and the method consumes that stream in the reference's order
 * (:22 flip, :35 rotate, :50 skew, :79-80 shear, :100-105 crop, :126 distortion).  Restates CPython's
 * MT19937 `random.seed(int)` / `random()` / `choice` / `randint` and PIL's rotate geometry.
 * transform[B]: LFX_AUG_*; seed[B]: task seeds (must be non-zero: seed 0 leaves the reference unseeded);
 * iparams[B][8], dparams[B][8] (HOST pointers), per transform:
 *   FLIP       i0 = lfx_flip mode (0 left-right, 1 top-bottom)
 *   ROTATE     i0..i5 = 16.16 inverse affine, i6 = nw, i7 = nh (lfx_rotate_nn params); d0 = angle
 *   SKEW/SHEAR d0..d7 = lfx_warp_bicubic coefficients, i0 = perspective flag
 *   CROP       i0..i3 = left, top, nw, nh (lfx_crop_lanczos box)
 *   DISTORTION i0 = cut = int(H*W*cutoff // 100) (lfx_distort), d0 = cutoff
 * threads: host threads to use (0 = all). */
enum { LFX_AUG_FLIP = 0, LFX_AUG_ROTATE = 1, LFX_AUG_SKEW = 2, LFX_AUG_SHEAR = 3, LFX_AUG_CROP = 4, LFX_AUG_DISTORTION = 5 };
int lfx_draw_augment_params(const int32_t* transform, const uint32_t* seed, int B, int H, int W,
                            int32_t* iparams, double* dparams, int threads);

/* The seeding half of lfx_draw_augment_params on the device: words[B][nwords] (device) = the first nwords (<= 226) 32-bit
 * outputs of Python's `random` after random.seed(seeds[i]) (image_augmenter.py:16-18; 0 <= seed < 2^32), one thread per
 * task.  lfx_draw_augment_params_words (host) then draws the parameters from a HOST copy of those words, in the same order
 * and with the same results as lfx_draw_augment_params; a task that needs more words is seeded on the host. */
int lfx_seed_words(const uint32_t* seeds, int B, int nwords, uint32_t* words, lfx_stream_t stream);
int lfx_draw_augment_params_words(const int32_t* transform, const uint32_t* seed, const uint32_t* words, int nwords, int B,
                                  int H, int W, int32_t* iparams, double* dparams);

/* ---- transform path: srcs/transform/filters/*.py, srcs/utils/mask_utils.py -------------------- */

/* Host-side: the task list of the balancing pass for an in-memory dataset (dataset_balancer.py:115-129 after
 * random.seed(seed), :31): group g = one (class, transform) pair in plan order with group_count[g] copies drawn from a
 * class of group_class_size[g] images; per task local_index = random.choice's index, task_seed = randint(0, 1000000). */
int lfx_draw_balance_tasks(uint32_t seed, int ngroups, const int32_t* group_count, const int32_t* group_class_size,
                           int32_t* local_index, int32_t* task_seed);

/* cv2.resize as the mask path uses it: INTER_CUBIC upscale of the working image (_prepare_working_image, mask.py:29-50)
 * and INTER_NEAREST of the mask back to the original size (_resize_results_to_original, :526-545).
 * lfx_cubic_table (host): first[out_size] = first source index of every destination index (tap k reads
 * clip(first + k)), weights[out_size][4] = a = -0.75 cubic weights x2048 (OpenCV's 8-bit fixed-point path).
 * lfx_resize_cubic: src [B,H,W,3] -> dst [B,OH,OW,3]; tables on the device (weights 16-byte aligned).  +-1 LSB
 * against cv2 (OpenCV's SIMD and scalar paths already differ by that much); lfx_resize_nearest is exact. */
int lfx_cubic_table(int in_size, int out_size, int32_t* first, int32_t* weights);
int lfx_resize_cubic(const uint8_t* src, uint8_t* dst, int B, int H, int W, int OH, int OW, const int32_t* xfirst,
                     const int32_t* xweights, const int32_t* yfirst, const int32_t* yweights, lfx_stream_t stream);
int lfx_resize_nearest(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int OH, int OW,
                       lfx_stream_t stream);

/* cv2.cvtColor(rgb, COLOR_RGB2{GRAY,HSV,LAB}) (mask.py:87,103; blur.py:27; hist.py:184).
 * code: 0 = GRAY (dst [B,H,W]), 1 = HSV, 2 = LAB (dst [B,H,W,3]). */
int lfx_cvt_color(const uint8_t* src, uint8_t* dst, int B, int H, int W, int code,
                  lfx_stream_t stream);

/* The TransformConfig fields the numeric path reads (Transformation.py:63-93). */
typedef struct lfx_mask_cfg {
    int32_t strategy;        /* 0 hsv_h, 1 lab, 2 hsv_s, 3 hsv_v_dark, 4 external raw mask (lfx_raw_mask) */
    int32_t green_lo, green_hi;          /* green_hue_range */
    int32_t fill_size;                   /* pcv.fill size */
    int32_t morph_kernel;                /* 3,5,7 or 9 */
    int32_t brown_lo, brown_hi;          /* brown_hue_range */
    int32_t brown_s_min, brown_v_max;
    int32_t brown_min_area_px;
    int32_t brown_morph_kernel;
    int32_t use_lab_brown, lab_a_min, lab_b_min;
    int32_t fallback_channel;            /* hsv_channel_for_mask: 0 h, 1 s, 2 v */
    int32_t bg_dark;                     /* bg_bias == "dark_bg" */
    int32_t extend_brown;                /* 1 = run _extend_mask_with_brown_regions (make_mask) */
    int32_t reserved[3];
} lfx_mask_cfg;

/* Threshold strategies _create_hsv_masks.mask_hsv_green / _create_lab_mask (mask.py:86-91,101-106)
 * fused with the colour conversion: raw candidate mask, no post-processing.
 * strategy 0 (hsv_h) or 1 (lab). */
int lfx_threshold_mask(const uint8_t* src, uint8_t* mask, int B, int H, int W,
                       const lfx_mask_cfg* cfg /* host */, lfx_stream_t stream);

/* make_mask (mask.py:548-582) under parity profile P0/P1 (grabcut_refine false, no upscale):
 * strategy mask -> _postprocess_mask (:53-69: fill, close, open, largest filled external contour)
 * -> Otsu fallback (:395-411) -> brown extension (:335-392).  One thread block per image.
 * raw: NULL for strategies 0-3 (computed from src in the kernel); for strategy 4 the raw
 * candidate [B,H,W] produced elsewhere (inclusive / enhanced front ends).
 * info[B][8] = {found, x, y, w, h, 2*contourArea, pixels, status | start_x << 8}; (x,y,w,h) =
 * cv2.boundingRect of the returned contour (roi.py:26); (start_x, y) is the contour's first raster
 * pixel (input of lfx_trace_contour); status bit0 = Otsu fallback used, bit1 = run table spilled
 * to global scratch. */
size_t lfx_make_mask_workspace(int B, int H, int W);
int lfx_make_mask(const uint8_t* src, const uint8_t* raw, uint8_t* mask, int32_t* info, int B,
                  int H, int W, const lfx_mask_cfg* cfg /* host */, void* workspace,
                  size_t workspace_bytes, lfx_stream_t stream);

/* _postprocess_mask alone (mask.py:53-69) on a raw mask. */
int lfx_postprocess_mask(const uint8_t* raw, uint8_t* mask, int32_t* info, int B, int H, int W,
                         int fill_size, int morph_kernel, void* workspace, size_t workspace_bytes,
                         lfx_stream_t stream);

/* largest_contour (Transformation.py:285-292): the external contour of the component selected by
 * lfx_make_mask / lfx_postprocess_mask, as cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
 * returns it.  points[B][max_pts][2] (x,y) int32; counts[B] = number of points (negative = -needed
 * when max_pts is too small; 0 = no contour); sums[B][3] (optional) = Green-formula sums
 * {a00, a10, a01} of the polygon: cv2.moments m00 = |a00|/2, m10 = a10/6, m01 = a01/6 (sign of a00)
 * (analyze.py:43-46). */
int lfx_trace_contour(const uint8_t* mask, const int32_t* info, int32_t* points, int32_t* counts,
                      int64_t* sums, int B, int H, int W, int max_pts, lfx_stream_t stream);

/* apply_analyze_filter's numeric record (analyze.py:43-98) for a batch of traced contours (outputs of
 * lfx_trace_contour): centroid of the contour polygon (cv2.moments, :43-49), extreme points (:60-64), convex hull (:77,
 * also mask.py:158), PCA axes of the contour vertices and the vertices with the extreme projections (:88-98).
 * rec_i32[B][24] = {valid, n_points, cx, cy, left.xy, right.xy, top.xy, bottom.xy, hull_count (negative = -needed when
 *   max_hull is too small), 0, p0_min.xy, p0_max.xy, p1_min.xy, p1_max.xy, 0, 0};
 * rec_f64[B][12] = {m00 (polygon area), hull area, mean.xy, eigenvector0.xy, eigenvector1.xy, eigenvalue0, eigenvalue1, 0, 0};
 * hull_points[B][max_hull][2] (x, y).  `sums` may be NULL (the Green-formula sums are then computed from the points).
 * Exact integer arithmetic for everything but the PCA (fp64). */
size_t lfx_analyze_workspace(int B, int H);
int lfx_analyze_record(const int32_t* points, const int32_t* counts, const int64_t* sums, int32_t* rec_i32,
                       double* rec_f64, int32_t* hull_points, int B, int H, int W, int max_pts, int max_hull,
                       void* workspace, size_t workspace_bytes, lfx_stream_t stream);

/* Overlay drawing, bit-identical to OpenCV 4.13's rasterisers (drawing.cpp: Line2 / FillConvexPoly /
 * ThickLine / Circle / LineAA / PolyLine in 16.16 fixed point), one thread block per image, later primitives over earlier
 * ones in the reference's order.
 * lfx_analyze_overlay = the image apply_analyze_filter returns (analyze.py:37-122): overlay [B,H,W,3] = rgb with the contour
 *   (drawContours, red, 2 px), the centroid cross (drawMarker 14 / 2 px), for the left / right / top / bottom points a filled
 *   circle r = 3 and an anti-aliased ray from the centroid, the anti-aliased convex hull (vertex order of cv2.convexHull),
 *   the two PCA axes (2 px) and, when `edges` (lfx_canny 80 / 160 L2 of the grey image) and `mask` are given, the vein
 *   pixels edges & mask in cyan.  points / counts come from lfx_trace_contour, rec_i32 / hull_points from
 *   lfx_analyze_record (same max_pts / max_hull; hull_count must be >= 0).  Images without a contour (rec_i32[0] == 0) are
 *   copied unchanged (the reference's "Analyze: no object" text banner is not drawn).
 * lfx_draw_rectangles = the `vis` image of apply_roi_filter (roi.py:43-44): cv2.rectangle(vis, (x, y), (x + w, y + h),
 *   colour, thickness) for info[B][8] = {found, x, y, w, h, ...} (lfx_make_mask's layout); color_rgb = r | g << 8 | b << 16.
 * Neither works in place. */
int lfx_analyze_overlay(const uint8_t* rgb, const int32_t* points, const int32_t* counts, const int32_t* rec_i32,
                        const int32_t* hull_points, const uint8_t* edges, const uint8_t* mask, uint8_t* overlay,
                        int B, int H, int W, int max_pts, int max_hull, lfx_stream_t stream);
int lfx_draw_rectangles(const uint8_t* rgb, const int32_t* info, uint8_t* vis, int B, int H, int W, uint32_t color_rgb,
                        int thickness, lfx_stream_t stream);

/* The same rasterisers as a general op: prims[B][max_prims][8] (int32) = {kind, x0, y0, x1, y1, color_rgb, size, 0}, counts[B]
 * primitives per image, drawn IN PLACE on img [B,H,W,3] in list order (later primitives over earlier ones), each bit-identical
 * to the OpenCV call:
 *   LFX_DRAW_LINE           cv2.line(img, (x0,y0), (x1,y1), color, thickness = size >= 2)            (LINE_8)
 *   LFX_DRAW_LINE_AA        cv2.line(img, (x0,y0), (x1,y1), color, 1, cv2.LINE_AA)
 *   LFX_DRAW_CIRCLE_FILLED  cv2.circle(img, (x0,y0), radius = size, color, -1)
 *   LFX_DRAW_RECTANGLE      cv2.rectangle(img, (x0,y0), (x1,y1), color, thickness = size >= 2)
 *   LFX_DRAW_MARKER_CROSS   cv2.drawMarker(img, (x0,y0), color, cv2.MARKER_CROSS, markerSize = x1, thickness = size >= 2)
 * Entries of any other kind (or with a size outside the stated range) are skipped.  End points may lie outside the image. */
enum { LFX_DRAW_NONE = 0, LFX_DRAW_LINE = 1, LFX_DRAW_LINE_AA = 2, LFX_DRAW_CIRCLE_FILLED = 3, LFX_DRAW_RECTANGLE = 4,
       LFX_DRAW_MARKER_CROSS = 5 };
int lfx_draw_primitives(uint8_t* img, const int32_t* prims, const int32_t* counts, int B, int H, int W, int max_prims,
                        lfx_stream_t stream);

/* Raw candidate of one threshold strategy, no post-processing (_build_mask_candidates, mask.py:414-443; strategies 0-3:
 * hsv_h, lab, hsv_s / hsv_v_dark by Otsu): raw [B,H,W] (0/255).  Workspace as lfx_make_mask. */
int lfx_strategy_raw(const uint8_t* src, uint8_t* raw, int B, int H, int W, const lfx_mask_cfg* cfg /* host */,
                     void* workspace, size_t workspace_bytes, lfx_stream_t stream);

/* k-means raw candidate: `_create_kmeans_mask` (srcs/transform/filters/mask.py:109-140) = cv2.setRNGSeed(seed = 12345),
 * cv2.kmeans(K = 3, (EPS + MAX_ITER, 20, 0.5), 1 attempt, KMEANS_PP_CENTERS), cluster chosen by hue / bg_bias / saturation.
 * cv::kmeans and cv::RNG are restated exactly (labels and centres bit-identical to OpenCV 4.13, tests/test_gpu_kmeans.py).
 * Only images whose longer side is 256 (the reference's working size: no INTER_AREA copy) -- others: LFX_ERR_UNSUPPORTED.
 * raw [B,H,W] (0/255); optional centers [B,3,3] (float32) and kinfo [B,4] = {picked cluster, iterations, empty-cluster
 * events, points}; bias: 0 = auto, 1 = dark_bg, 2 = light_bg (TransformConfig.bg_bias). */
int lfx_kmeans_raw(const uint8_t* src, uint8_t* raw, float* centers, int32_t* kinfo, int B, int H, int W, int green_lo,
                   int green_hi, int bias, uint32_t seed, lfx_stream_t stream);

/* Image-dependent terms of _score_mask (mask.py:143-188) for K (<= 8) post-processed candidate masks per image
 * (masks [K][B][H][W]), as `mask_strategy: auto` ranks them (:435-461).
 * feat[K][B][4] (double) = {sum of the Sobel magnitude (float32 values, fp64 sum) over the mask boundary
 *   dilate3x3 ^ erode3x3, boundary pixels, mask pixels, pixels that are mask and green (H in [green_lo, green_hi], S >= 40)};
 * minmax[B][2] = {float bits of max |grad|, ~(float bits of min |grad|)} over the whole image (cv2.normalize NORM_MINMAX). */
int lfx_score_features(const uint8_t* src, const uint8_t* masks, double* feat, uint32_t* minmax, int B, int H, int W,
                       int K, int green_lo, int green_hi, lfx_stream_t stream);

/* apply_mask (mask_utils.py:10-83): dst = mask > 127 ? src : color_val. */
int lfx_apply_mask(const uint8_t* src, const uint8_t* mask, uint8_t* dst, int B, int H, int W,
                   int color_val, lfx_stream_t stream);

/* HOST helper: OpenCV's 8-bit Gaussian taps (8 fractional bits, error-diffused, sum 256). */
int lfx_gauss_taps(int ksize, double sigma, int32_t* taps /* host [ksize] */);

/* cv2.GaussianBlur(img,(k,k),sigma) for uint8, BORDER_REFLECT_101 (blur.py:61,72; mask.py:223,770).
 * C = 1 or 3 interleaved channels; ksize odd <= 15. Bit-exact fixed point (8.8 then 16.16). */
int lfx_gauss_u8(const uint8_t* src, uint8_t* dst, int B, int H, int W, int C, int ksize,
                 double sigma, lfx_stream_t stream);

/* apply_roi_filter canvas (roi.py:20-46): crop info's bounding box from apply_mask(src, mask,
 * white) (mask may be NULL = no masking), letterbox with cv2.resize(INTER_AREA) into a zero
 * [RH,RW,3] canvas -- an upscale when the box fits the canvas, OpenCV's area averaging when it is larger (images bigger
 * than roi_size); both bit-exact.  Images with info.found == 0 get an all-zero canvas. */
int lfx_roi_letterbox(const uint8_t* src, const uint8_t* mask, const int32_t* info, uint8_t* dst,
                      int B, int H, int W, int RH, int RW, lfx_stream_t stream);

/* Colour statistics.  hist9[B][9][256]: histograms of R,G,B,H,S,V,L,a,b over mask > 0 (whole
 * image when mask is NULL) -- PIL histogram inside autocontrast (image_augmenter.py:127), LAB-L
 * percentiles (mask.py:206-219), dataset colour histograms.  hsv3[B][3][256] and
 * counters[B][16]: apply_histogram_filter's numeric core on apply_mask(src, mask, white)
 * (hist.py:188 leaf_mask, :38-65 eight categories, :248-256 five hue ranges):
 * counters = {leaf_px, 8 categories, 5 hue ranges, 0, 0}.  Any output may be NULL.
 * Outputs are ACCUMULATED into (zero them first). */
int lfx_color_stats(const uint8_t* src, const uint8_t* mask, int32_t* hist9, int32_t* hsv3,
                    int32_t* counters, int B, int H, int W, lfx_stream_t stream);

/* ---- per-image front ends (one thread block per image).  Images up to ~256x256 (PlantVillage) keep every bit plane and
 *      the grey image in shared memory; larger ones (512x512, 1024x1024 ...) keep them in the block's global scratch
 *      (lfx_front_workspace accounts for it).  Rows longer than 1152 pixels return LFX_ERR_UNSUPPORTED. ------------------ */

/* Scratch for the four entry points below (run tables that outgrow shared memory + float planes). */
size_t lfx_front_workspace(int B, int H, int W);

/* cv2.Canny(gray, low, high, apertureSize=3, L2gradient) (mask.py:679-680,789; blur.py:30;
 * analyze.py:120): gray [B,H,W] -> edges [B,H,W] (0/255).  Bit-exact. */
int lfx_canny(const uint8_t* gray, uint8_t* edges, int B, int H, int W, double low, double high,
              int l2gradient, void* workspace, size_t workspace_bytes, lfx_stream_t stream);

/* Raw candidate masks of the composite strategies: which = 0 _create_inclusive_mask (mask.py:727-831,
 * the reference's default strategy), which = 1 _create_enhanced_mask (mask.py:610-724).  The result
 * feeds lfx_make_mask with strategy 4. */
int lfx_raw_mask(const uint8_t* src, uint8_t* raw, int B, int H, int W, int which,
                 const lfx_mask_cfg* cfg /* host */, void* workspace, size_t workspace_bytes,
                 lfx_stream_t stream);

/* apply_brown_filter numeric core (brown.py:21-89): brown predicate & leaf mask -> open/close ->
 * 8-connected components with area >= brown_min_area_px.  spots [B,H,W] (0/255);
 * stats[B][4] = {leaf_pixels, spot_count, spot_pixels, status}. */
int lfx_brown_spots(const uint8_t* src, const uint8_t* mask, uint8_t* spots, int32_t* stats, int B,
                    int H, int W, const lfx_mask_cfg* cfg /* host */, void* workspace,
                    size_t workspace_bytes, lfx_stream_t stream);

/* apply_blur_filter (blur.py:18-79): saliency map (Canny 50/150 L2 + Sobel magnitude + brown regions
 * + colour difference to a 15x15 blur, three min-max normalisations) -> GaussianBlur 5x5
 * (gaussian_sigma) -> zero outside `mask` -> grey replicated to RGB.  dst [B,H,W,3].
 * Integer stages are exact; the float32 normalisations make this a +-1 LSB op. */
int lfx_saliency_blur(const uint8_t* src, const uint8_t* mask, uint8_t* dst, int B, int H, int W,
                      double gaussian_sigma, const lfx_mask_cfg* cfg /* host */, void* workspace,
                      size_t workspace_bytes, lfx_stream_t stream);

/* Fused core transform profile (BASELINE config 2: blur + mask + ROI + histograms), equivalent to
 * lfx_gauss_u8(5x5) + lfx_make_mask + lfx_roi_letterbox + lfx_color_stats in one submission.
 * dataset_hist9 (optional, device int64 [9][256], needs hist9): the batch's histograms are ADDED to it -- the
 * per-rank partial of the dataset-level colour histogram that one allreduce merges.
 * raw (NULL unless cfg->strategy == 4): the raw candidate [B,H,W] of a composite strategy (lfx_raw_mask: inclusive, the
 * reference's default, or enhanced) -- the default-strategy profile is then lfx_raw_mask + this one call. */
size_t lfx_pipeline_core_workspace(int B, int H, int W);
int lfx_pipeline_core(const uint8_t* src, uint8_t* blur, uint8_t* mask, int32_t* info,
                      uint8_t* roi, int32_t* hist9, int32_t* hsv3, int32_t* counters, int B, int H,
                      int W, int RH, int RW, double gaussian_sigma, const lfx_mask_cfg* cfg /* host */,
                      void* workspace, size_t workspace_bytes, int64_t* dataset_hist9, const uint8_t* raw,
                      lfx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LEAFX_H */
