/*
 * pcseg_b200 - C ABI of the B200-native (sm_100a) OCR4All pixel-classifier
 * inference hot path.
 *
 * The reference (ocr4all_pixel_classifier 0.6.5) is pure Python and has no FFI;
 * its hot path is the Python API of lib/predictor.py, lib/network.py,
 * lib/dataset.py, lib/output.py and lib/postprocess.py.  Every entry point
 * below names the reference function (file:line under
 * ocr4all_pixel_classifier/) whose arithmetic it replaces.  The Python shim in
 * page_segmentation_b200/lib/ binds these symbols with ctypes and re-exposes
 * the reference's class / function names on top of them.
 *
 * Conventions
 *   - every function returns 0 on success or a negative pcs_status;
 *     pcs_last_error(ctx) returns a message for the last failure on that ctx;
 *   - no exceptions and no ownership cross the ABI: all in/out buffers are
 *     caller-allocated; activations/workspace are owned by the ctx;
 *   - pointers named d_* are DEVICE pointers on the ctx's device, pointers
 *     named h_* are HOST pointers (pinned memory makes the copies async);
 *   - all work is enqueued on the ctx's stream (pcs_set_stream); functions are
 *     asynchronous with respect to that stream unless stated otherwise;
 *   - one ctx per GPU per host thread (thread-compatible, not thread-safe);
 *   - images are row-major, pages of a batch are contiguous ([n][H][W]...).
 */
#ifndef PCSEG_B200_H
#define PCSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCS_ABI_VERSION 1

#if defined(__GNUC__)
#define PCS_API __attribute__((visibility("default")))
#else
#define PCS_API
#endif

typedef struct pcs_ctx pcs_ctx;

typedef enum {
    PCS_OK = 0,
    PCS_ERR_ARG = -1,       /* bad argument / unsupported configuration */
    PCS_ERR_CUDA = -2,      /* CUDA runtime or driver error (see pcs_last_error) */
    PCS_ERR_STATE = -3,     /* call order: e.g. forward before model load */
    PCS_ERR_NOMEM = -4,     /* device allocation failed */
    PCS_ERR_DEVICE = -5,    /* not an sm_100 device */
    PCS_ERR_IO = -6         /* a file of pcs_output_pages could not be written */
} pcs_status;

/* Architecture.value strings of lib/architecture.py:6-11 that are in scope. */
typedef enum { PCS_ARCH_FCN_SKIP = 0, PCS_ARCH_FCN = 1, PCS_ARCH_UNET = 2 } pcs_arch;

/* Operand type of the tensor-core convolutions (accumulation is always fp32). */
typedef enum { PCS_PREC_BF16 = 0, PCS_PREC_FP16 = 1 } pcs_precision;

/* Which convolution engine runs the body: hand-written tcgen05/TMEM/TMA
 * implicit GEMM (default) or the CUDA-core direct kernels (numerics twin). */
typedef enum { PCS_ENGINE_UMMA = 0, PCS_ENGINE_DIRECT = 1 } pcs_engine;

/* One layer's weights exactly as Keras stores them (lib/model.py:45-92,
 * :151-203, :206-234): Conv2D kernel (kh,kw,C_in,C_out), Conv2DTranspose
 * kernel (kh,kw,C_out,C_in), both float32 row-major; bias (C_out). */
typedef struct {
    const float* kernel;
    const float* bias;
    int32_t shape[4];
} pcs_layer_weights;

/* ---- context -------------------------------------------------------- */
PCS_API int pcs_abi_version(void);
PCS_API int pcs_ctx_create(int device, pcs_ctx** out);
PCS_API void pcs_ctx_destroy(pcs_ctx* ctx);
PCS_API const char* pcs_last_error(const pcs_ctx* ctx);
PCS_API int pcs_set_stream(pcs_ctx* ctx, void* cuda_stream);
PCS_API int pcs_synchronize(pcs_ctx* ctx);
/* number of kernels this library launched on ctx since creation */
PCS_API int64_t pcs_launch_count(const pcs_ctx* ctx);

/* ---- model: replaces Network.__init__ weight loading, lib/network.py:75-107
 * (the .h5 container is parsed on the host side; this uploads the tensors in
 * Keras layer order, pre-transformed for the kernels).  Synchronous. */
PCS_API int pcs_model_load(pcs_ctx* ctx, int arch, int n_classes, int precision,
                   const pcs_layer_weights* layers, int n_layers);
PCS_API int pcs_set_engine(pcs_ctx* ctx, int engine);

/* ---- preprocess: replaces prepare_images, lib/dataset.py:131-150 (with
 * scale_binary :114-119 and scale_image :122-128, i.e. skimage rescale order 0
 * / resize order 3, mode='reflect', clip, preserve_range).
 *   d_grey, d_bin : [n][H][W] uint8 source pages (may alias, dataset.py:169-172)
 *   Hs, Ws        : np.round(scale * (H, W)) computed by the caller
 *   d_image       : [n][Hs][Ws] uint8  `data.image`  (inverted, truncated)
 *   d_binary      : [n][Hs][Ws] uint8  `data.binary` in {0,1}, 1 = ink
 *   d_orig_binary : [n][H][W] uint8 `data.orig_binary` or NULL
 * The anti-aliasing Gaussian of resize() is applied per page when the page has
 * more than two grey levels, as dataset.py:127 decides. */
PCS_API int pcs_preprocess(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin,
                   int n, int H, int W, int Hs, int Ws,
                   uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary);

/* ---- prepare_images for pages that arrive BIT-PACKED.  A binarised page is one bit per pixel by nature (the reference
 * decodes it into a uint8 array, lib/dataset.py:169-172); a caller that holds it packed moves 1.1 MB per A4 page across
 * PCIe instead of 8.7.  Layout: flat over the page, pixel i = bit (i & 31) of word (i >> 5) - numpy.packbits(page.ravel()
 * != level0, bitorder='little') viewed as little-endian uint32 - words_per_page >= H * W / 32 + 1 (one readable pad
 * word); a clear bit is a pixel of grey level `level0`, a set bit one of `level1`.  Results are exactly those of
 * pcs_preprocess on the uint8 page `bit ? level1 : level0` used as grey and binary page (scale factors up to 4). */
PCS_API int pcs_preprocess_bits(pcs_ctx* ctx, const uint32_t* d_bits, size_t words_per_page, int n, int H, int W,
                        int level0, int level1, int Hs, int Ws, uint8_t* d_image, uint8_t* d_binary);
/* uint8 planes <-> flat bit planes of the same layout: non-zero byte <-> set bit <-> byte 1; n planes of n_pixels */
PCS_API int pcs_pack_bits(pcs_ctx* ctx, const uint8_t* d_src, int n, size_t n_pixels, uint32_t* d_bits, size_t words_per_page);
PCS_API int pcs_unpack_bits(pcs_ctx* ctx, const uint32_t* d_bits, int n, size_t words_per_page, size_t n_pixels, uint8_t* d_dst);

/* ---- prepare_images with max_width (lib/dataset.py:139-143): after the first
 * rescale to (H1, W1) = np.round(scale * (H, W)), `n_scale = max_width / W1 < 1`
 * triggers a second one to (H2, W2) = np.round(n_scale * (H1, W1)): order 0 for
 * the binary, order 3 for the fp64 image (anti-aliased when it holds more than
 * two distinct values), then (img * 255).astype(uint8).  The caller computes the
 * two shapes.  d_image, d_binary: [n][H2][W2]; d_orig_binary: [n][H][W] or NULL. */
PCS_API int pcs_preprocess_max_width(pcs_ctx* ctx, const uint8_t* d_grey, const uint8_t* d_bin,
                             int n, int H, int W, int H1, int W1, int H2, int W2,
                             uint8_t* d_image, uint8_t* d_binary, uint8_t* d_orig_binary);

/* ---- network body + head: replaces Network.predict_single_data,
 * lib/network.py:248-260 = default_preprocess x/255 (architecture.py:67-68),
 * the Keras graph (model.py:45-92 / :206-234 / :151-203 incl. pad :20-26 and
 * crop :29-42), scipy softmax and np.argmax; optionally fused with the colour
 * epilogue generate_output_masks, lib/output.py:44-60.
 *   d_image  : [n][Hs][Ws] uint8 network input (`data.image`)
 *   d_binary : [n][Hs][Ws] uint8 {0,1} (only read when masks are requested)
 *   d_labels : [n][Hs][Ws] uint8 argmax class (ties -> lowest index)
 *   d_logits, d_prob : [n][Hs][Ws][n_classes] float32 or NULL
 *   lut      : n_classes x 3 uint8 HOST array (ColorMap label -> rgb) or NULL
 *   d_color, d_overlay, d_inverted : [n][Hs][Ws][3] uint8 or NULL */
PCS_API int pcs_forward(pcs_ctx* ctx, const uint8_t* d_image, const uint8_t* d_binary,
                int n, int Hs, int Ws,
                uint8_t* d_labels, float* d_logits, float* d_prob,
                const uint8_t* lut, uint8_t* d_color, uint8_t* d_overlay, uint8_t* d_inverted);

/* ---- colour epilogue on an existing class map: replaces
 * generate_output_masks, lib/output.py:44-60 (fg_color_mask == inverted). */
PCS_API int pcs_masks(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary,
              int n, int H, int W, const uint8_t* lut, int n_lut,
              uint8_t* d_color, uint8_t* d_overlay, uint8_t* d_inverted);

/* ---- nearest-neighbour resize: replaces preserving_resize, lib/util.py:21-29
 * as used by scale_to_original_shape, lib/output.py:63-79 (uint8 planes). */
PCS_API int pcs_resize_nearest(pcs_ctx* ctx, const uint8_t* d_src, int n, int H, int W,
                       uint8_t* d_dst, int Ho, int Wo);

/* ---- connected components, 4-connectivity: replaces
 * cv2.connectedComponentsWithStats(img, connectivity=4) as called at
 * lib/postprocess.py:10,33 and lib/image_ops.py:68.  Labels are numbered in
 * raster order of each component's first pixel (label 0 = background).
 *   d_img     : [n][H][W] uint8, nonzero = foreground
 *   d_labels  : [n][H][W] int32
 *   d_stats   : [n][max_components][5] int32 (left, top, width, height, area),
 *               row 0 = background, or NULL
 *   d_ncomp   : [n] int32 number of labels incl. background */
PCS_API int pcs_ccl(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W,
            int32_t* d_labels, int32_t* d_stats, int max_components, int32_t* d_ncomp);

/* ---- replaces vote_connected_component_class, lib/postprocess.py:9-26:
 * every 4-connected component of d_binary gets the most frequent class of
 * d_pred inside it (ties -> lowest class); in place on d_pred. */
PCS_API int pcs_cc_majority(pcs_ctx* ctx, uint8_t* d_pred, const uint8_t* d_binary,
                    int n, int H, int W, int n_classes);

/* ---- replaces add_bounding_boxes, lib/postprocess.py:29-42 (evident intent:
 * per class c ascending, every component of pred==c paints its bbox with c). */
PCS_API int pcs_bounding_boxes(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W,
                       int n_classes, uint8_t* d_out);

/* ---- segment extraction: the per-class labelling of add_bounding_boxes, lib/postprocess.py:31-33
 * (cv2.connectedComponentsWithStats(pred == c, connectivity=4) for every class c), returning the stats table that
 * lib/cc.py:4-18 (cc_bbox / cc_bbox_func) indexes instead of a painted map: BASELINE configs[3]'s "connected-component
 * segment extraction".
 *   d_pred  : [n][H][W] uint8 class map
 *   d_stats : [n][n_classes][max_components][5] int32 (left, top, width, height, area); row 0 = the labelling's
 *             background (every pixel != c), rows 1.. = the components of class c in raster order of their first
 *             pixel, rows at or beyond the label count are zero; components beyond max_components are dropped
 *   d_ncomp : [n][n_classes] int32 number of labels incl. background (may exceed max_components), or NULL */
PCS_API int pcs_class_components(pcs_ctx* ctx, const uint8_t* d_pred, int n, int H, int W, int n_classes,
                         int32_t* d_stats, int max_components, int32_t* d_ncomp);

/* ---- replaces compute_char_height, lib/image_ops.py:58-82 (the producer of
 * `line_height_px`): cv2 Otsu threshold of the grey page, inversion unless
 * `inverse`, connected components (8-connected: the reference's positional `4`
 * lands in cv2's `labels` slot, so cv2's default connectivity applies), boxes
 * with 0.5 < w/h < 2, 10 < h < 60, 5 < w < 50, and the height at index len/2 of
 * the sorted valid heights.
 *   d_img    : [n][H][W] uint8 grey pages
 *   d_height : [n] int32, -1 where the reference returns None (no valid box) */
PCS_API int pcs_char_height(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W,
                    int inverse, int32_t* d_height);

/* ---- whole page batch through the pipeline with HOST buffers (the copies are
 * part of the call): prepare_images -> predict_single_data -> optional
 * cc_majority -> generate_output_masks.  Synchronous.  Any output may be NULL.
 *   h_grey/h_bin  : [n][H][W] uint8;  h_labels : [n][Hs][Ws] uint8;
 *   h_color/h_overlay/h_inverted : [n][Hs][Ws][3] uint8 */
PCS_API int pcs_predict_pages_host(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin,
                           int n, int H, int W, int Hs, int Ws, int cc_majority,
                           const uint8_t* lut,
                           uint8_t* h_image, uint8_t* h_binary, uint8_t* h_labels,
                           uint8_t* h_color, uint8_t* h_overlay, uint8_t* h_inverted);

/* The same pipeline with the three masks leaving the device as PNG files (pcs_png_encode, level 1) instead of raw
 * arrays - Predictor.predict + output_data (lib/predictor.py:27-42, lib/output.py:20-41) for a batch, minus the
 * file write.  The device-to-host traffic drops from 9.7 MB to about 1.3 MB per A4 page.
 *   h_labels    : [n][Hs][Ws] uint8 or NULL
 *   h_png       : file (page p, kind k) starts at h_png + (3 p + k) * png_stride; kind 0 = color, 1 = overlay,
 *                 2 = inverted; png_stride >= pcs_png_bytes(Hs, Ws, 3, 1)
 *   h_png_sizes : [n][3] uint64 file lengths */
PCS_API int pcs_predict_pages_files(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin,
                            int n, int H, int W, int Hs, int Ws, int cc_majority, const uint8_t* lut,
                            uint8_t* h_labels, uint8_t* h_png, size_t png_stride, uint64_t* h_png_sizes);

/* The same pipeline with COMPACT results: what Predictor.predict yields per page (lib/predictor.py:27-42) is the class
 * map; the colour masks are pure functions of (class map, binary page, LUT) and are produced on request (pcs_unpack_bits
 * + pcs_masks) instead of crossing PCIe for every page: 1.1 MB per A4 page leave the device instead of 9.7.
 *   h_labels      : [n][Hs][Ws] uint8
 *   h_binary_bits : [n][ceil(Hs * Ws / 32)] uint32, `data.binary` bit-packed (layout of pcs_preprocess_bits), or NULL
 * pcs_predict_pages_packed additionally takes the PAGES bit-packed ([n][ceil(H * W / 32)] words, levels as in
 * pcs_preprocess_bits): 1.1 MB per page in, 1.1 MB out. */
PCS_API int pcs_predict_pages_compact(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W,
                              int Hs, int Ws, int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits);
PCS_API int pcs_predict_pages_packed(pcs_ctx* ctx, const uint32_t* h_bits, int level0, int level1, int n, int H, int W,
                             int Hs, int Ws, int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits);

/* STREAMING form of the compact calls: Predictor.predict is a generator (lib/predictor.py:27-30), a caller with a
 * long page list feeds it batch after batch.  `_submit` queues the whole call and returns; *ticket names it.  A submit
 * that follows a submit of the same shapes is chained onto it: its upload runs under the kernels of the call before, so
 * the fill and the drain of the three-stage pipeline are paid once per run of submits instead of once per call.
 * pcs_wait_pages(ticket) returns when the results of that submit are in the caller's host buffers (which, like the
 * input pages, must stay untouched until then); pcs_synchronize waits for everything.  Results are those of the
 * blocking calls, bit for bit. */
PCS_API int pcs_predict_pages_compact_submit(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin, int n, int H, int W,
                              int Hs, int Ws, int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits, uint64_t* ticket);
PCS_API int pcs_predict_pages_segments_compact_submit(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin,
                               int n, int H, int W, int Hs, int Ws, int cc_majority,
                               uint8_t* h_labels, uint32_t* h_binary_bits,
                               int32_t* h_stats, int max_components, int32_t* h_ncomp, uint64_t* ticket);
PCS_API int pcs_predict_pages_packed_submit(pcs_ctx* ctx, const uint32_t* h_bits, int level0, int level1, int n, int H, int W,
                             int Hs, int Ws, int cc_majority, uint8_t* h_labels, uint32_t* h_binary_bits, uint64_t* ticket);
PCS_API int pcs_wait_pages(pcs_ctx* ctx, uint64_t ticket);

/* The same pipeline followed by segment extraction (pcs_class_components on the final class map): BASELINE configs[3],
 * "normalization rescale + FCN + connected-component segment extraction".  The masks are optional.
 *   h_stats : [n][n_classes][max_components][5] int32;  h_ncomp : [n][n_classes] int32 or NULL */
PCS_API int pcs_predict_pages_segments(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin,
                               int n, int H, int W, int Hs, int Ws, int cc_majority, const uint8_t* lut,
                               uint8_t* h_labels, uint8_t* h_color, uint8_t* h_overlay, uint8_t* h_inverted,
                               int32_t* h_stats, int max_components, int32_t* h_ncomp);

/* The segment call with the compact transport of pcs_predict_pages_compact: uint8 pages in; class map, bit-packed
 * `binary` (or NULL), stats tables and label counts out -- 1.3 MB per page instead of 9.9 MB with the three masks,
 * which are a function of the class map, the binary and the colour table (pcs_unpack_bits + pcs_masks). */
PCS_API int pcs_predict_pages_segments_compact(pcs_ctx* ctx, const uint8_t* h_grey, const uint8_t* h_bin,
                               int n, int H, int W, int Hs, int Ws, int cc_majority,
                               uint8_t* h_labels, uint32_t* h_binary_bits,
                               int32_t* h_stats, int max_components, int32_t* h_ncomp);

/* ---- image files: the encoder of output_data, lib/output.py:38-41 (skimage.io.imsave of the three masks).
 * Builds n complete PNG files on the device from [n][H][W][channels] uint8 images (channels 1 = grey, 3 = RGB,
 * 4 = RGBA): signature, IHDR, one IDAT, Adler-32, CRC-32, IEND.  level 0: stored deflate blocks, the file is the
 * raw image + 0.2 %; level 1: Sub-filtered scanlines, fixed-Huffman deflate with run-length matches (class-colour
 * masks shrink 30-100x, scanlines up to 16384 bytes).  File i starts at d_out + i * stride (stride a multiple of 4
 * and >= pcs_png_bytes rounded up to 4; at level 1 the buffer is zero-filled first); its length is written to
 * d_sizes[i] (device, may be NULL at level 0 where it equals pcs_png_bytes).  Any PNG reader decodes the files to
 * exactly the input bytes. */
PCS_API size_t pcs_png_bytes(int H, int W, int channels, int level);   /* exact (level 0) / upper bound (level 1); 0 = unsupported */
PCS_API int pcs_png_encode(pcs_ctx* ctx, const uint8_t* d_img, int n, int H, int W, int channels, int level,
                   uint8_t* d_out, size_t stride, uint64_t* d_sizes);

/* ---- output_data, lib/output.py:20-41, for n pages in one asynchronous call: generate_output_masks (pcs_masks) + the
 * three PNG files of every page (pcs_png_encode, level 1; level 0 for scanlines beyond its limit).  d_labels /
 * d_binary: [n][H][W] uint8 device pointers (class map, data.binary); lut: HOST [n_lut][3]; paths: HOST array of
 * 3 n C strings, page-major (page p: paths[3 p] = color, [3 p + 1] = overlay, [3 p + 2] = inverted), copied by the call.
 * The kernels run on the context's stream into buffers the library owns (the inputs may be released once work queued
 * on the stream after this call has run); worker threads of the library fetch only the bytes of the files and write
 * them.  1 <= n <= 64.  pcs_output_flush returns once every file handed over so far is on disk and reports the first
 * write error (PCS_ERR_IO, also reported by the next pcs_output_pages); pcs_ctx_destroy flushes. */
PCS_API int pcs_output_pages(pcs_ctx* ctx, const uint8_t* d_labels, const uint8_t* d_binary, int n, int H, int W,
                             const uint8_t* lut, int n_lut, const char* const* paths);
PCS_API int pcs_output_flush(pcs_ctx* ctx);

/* ---- region extraction (downstream consumer of the `inverted` colour image):
 * the pixel work of lib/pc_segmentation.py and lib/xycut.py; the data-dependent
 * parts (XY-cut recursion, contour tracing) stay on the host. -------------- */

/* find_segments, lib/pc_segmentation.py:28-33 + :48/:56: cv2.resize(image,
 * (Wo, Ho), INTER_NEAREST) -> dilate with a 3x3 rectangle (:63-67) ->
 * color_map.filter_label for each of the m <= 8 colours.
 *   d_rgb    : [H][W][3] uint8 colour image
 *   colours  : HOST [m][3] uint8
 *   d_masks  : [m][Ho][Wo] uint8 in {0,1} */
PCS_API int pcs_segment_masks(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, int Ho, int Wo,
                      const uint8_t* colours, int m, uint8_t* d_masks);

/* dilate, lib/pc_segmentation.py:63-67: cv2.dilate with a 3x3 rectangle of an
 * interleaved uint8 image [H][W][C]; d_dst must not alias d_src. */
PCS_API int pcs_dilate3x3(pcs_ctx* ctx, const uint8_t* d_src, int H, int W, int C, uint8_t* d_dst);

/* projection profiles of do_xy_cut / recursive_cut, lib/xycut.py:95-161: the
 * recursion evaluates np.count_nonzero(sub_image, axis) on nested sub-rectangles
 * (:135); all of them are differences of rows / columns of the summed-area table
 *   d_sat : [n][H+1][W+1] int32, sat[y][x] = #{mask[y'<y][x'<x] != 0}. */
PCS_API int pcs_integral_image(pcs_ctx* ctx, const uint8_t* d_mask, int n, int H, int W, int32_t* d_sat);

/* get_text_contours, lib/pc_segmentation.py:70-96: cv2.inRange(image, colour,
 * colour) -> MORPH_CLOSE with a k_close x k_close rectangle -> MORPH_OPEN with
 * k_open -> dilate with k_region -> MORPH_CLOSE with k_region (OpenCV anchors,
 * borders and even-sized elements).  Outputs, both [H][W] uint8 or NULL:
 *   d_text_inv : 255 - image after the opening (:96, the canvas the contours are drawn on)
 *   d_region   : region_text (:93, the input of the first cv2.findContours) */
PCS_API int pcs_text_regions(pcs_ctx* ctx, const uint8_t* d_rgb, int H, int W, const uint8_t* colour,
                     int k_close, int k_open, int k_region, uint8_t* d_text_inv, uint8_t* d_region);

/* ---- evaluation counts: the reductions of fgpa and fgoverlap_per_class, lib/image_ops.py:8-55.  Over the pixels with
 * bin != 0: d_out[0] = their number, d_out[1] = those with pred != mask, d_out[2 + p * (n_classes + 2) + m] = the
 * confusion matrix of (pred, mask), class values above n_classes folded into the last bucket.  uint8 class maps. */
PCS_API int pcs_eval_counts(pcs_ctx* ctx, const uint8_t* d_pred, const uint8_t* d_mask, const uint8_t* d_bin,
                    size_t n_pixels, int n_classes, uint64_t* d_out);

/* ---- training step primitives (first version: fp32 on the CUDA cores): Network.train_dataset with batch 1,
 * lib/network.py:151-161,167-242; metrics.loss, lib/metrics.py:8-9; Keras Adam with per-variable clipnorm as compiled
 * at lib/network.py:91-103.  Tensors are planar float32 [C][H][W] on the device; the host side (lib/trainer.py)
 * walks the graph of lib/model.py:45-92 / :206-234.
 *   corr2d      : 'same' correlation, k in {1, 5}: y = act(b + x (*) w), w[c_out][c_in][k][k].  Serves Conv2D, the
 *                 stride-1 Conv2DTranspose (flipped kernel), both input gradients (transposed / flipped kernel) and
 *                 the logits; accumulate = 1 adds to y (gradients meeting at a skip connection).
 *   wgrad       : dw[c_out][c_in][k][k] = sum over pixels of dy * shifted x
 *   deconv2_*   : Conv2DTranspose(2x2, stride 2), k2[tap = 2i + j][c_out][c_in]
 *   softmax_ce  : mean sparse cross entropy from logits over the Hc x Wc crop; d_loss_sum receives the SUM over pixels
 *   adam        : one variable per offsets[i] .. offsets[i+1]: g *= grad_scale, clip_by_norm(g, clipnorm) when
 *                 clipnorm > 0, m / v update, p -= lr_t * m / (sqrt(v) + eps) with lr_t computed by the caller */
PCS_API int pcs_train_input(pcs_ctx* ctx, const uint8_t* d_image, int h, int w, float* d_plane, int H, int W);
PCS_API int pcs_train_corr2d(pcs_ctx* ctx, const float* d_x, const float* d_w, const float* d_b, float* d_y,
                     int c_in, int c_out, int H, int W, int k, int relu, int accumulate);
PCS_API int pcs_train_wgrad(pcs_ctx* ctx, const float* d_x, const float* d_dy, float* d_dw, int c_in, int c_out, int H, int W, int k);
PCS_API int pcs_train_bias_grad(pcs_ctx* ctx, const float* d_dy, float* d_db, int channels, size_t plane);
PCS_API int pcs_train_relu_bwd(pcs_ctx* ctx, float* d_dy, const float* d_y, size_t n);
PCS_API int pcs_train_maxpool_fwd(pcs_ctx* ctx, const float* d_x, float* d_y, int channels, int H, int W);
PCS_API int pcs_train_maxpool_bwd(pcs_ctx* ctx, const float* d_x, const float* d_dy, float* d_dx, int channels, int H, int W, int accumulate);
PCS_API int pcs_train_deconv2_fwd(pcs_ctx* ctx, const float* d_x, const float* d_k2, const float* d_b, float* d_y,
                          int c_in, int c_out, int h, int w, int relu);
PCS_API int pcs_train_deconv2_bwd_data(pcs_ctx* ctx, const float* d_dy, const float* d_k2, float* d_dx, int c_in, int c_out, int h, int w);
PCS_API int pcs_train_deconv2_wgrad(pcs_ctx* ctx, const float* d_x, const float* d_dy, float* d_dk2, int c_in, int c_out, int h, int w);
PCS_API int pcs_train_softmax_ce(pcs_ctx* ctx, const float* d_logits, const uint8_t* d_labels, int n_classes, int H, int W,
                         int Hc, int Wc, float* d_dlogits, double* d_loss_sum);
PCS_API int pcs_train_adam(pcs_ctx* ctx, float* d_params, const float* d_grads, float* d_m, float* d_v, const int64_t* d_offsets,
                   int n_vars, float lr_t, float beta1, float beta2, float eps, float clipnorm, float grad_scale);

/* ---- training step on the tensor cores (csrc/train_tc.cu): the same step as the primitives above compose -- forward
 * with kept activations, mean sparse cross entropy from logits (metrics.py:8-9), backward -- for ONE page of the
 * fcn_skip / fcn graph (network.py:151-161 trains with batch 1), in mixed precision: bf16 activations and activation
 * gradients, fp32 accumulation, fp32 master weights / gradients.  `offsets` (host, 27 entries) are the element offsets of
 * kernel_0, bias_0, kernel_1, ... and the end in the flat parameter / gradient buffers, kernels in the layouts of the
 * primitives above (w[C_out][C_in][k][k]; stride-1 transposed convolutions flipped; 2x2 stride-2 ones k2[tap][C_out][C_in]).
 * pcs_train_tc_step fills d_grads (zeroed by phase 1) and *d_loss_sum (un-normalised: divide by h * w).  `phases`: bit 0 =
 * forward + loss + backward of logits .. deconv1, bit 1 = backward of conv7 .. conv1 (two calls let the caller start the
 * data-parallel all-reduce of the first half while the second runs); 3 = the whole step. */
typedef struct pcs_train_tc pcs_train_tc;
PCS_API int pcs_train_tc_create(pcs_ctx* ctx, int arch, int n_classes, int h, int w, const int64_t* offsets, int n_offsets,
                                pcs_train_tc** out);
PCS_API int pcs_train_tc_step(pcs_ctx* ctx, pcs_train_tc* step, int phases, const uint8_t* d_image, const uint8_t* d_labels,
                              const float* d_params, float* d_grads, double* d_loss_sum);
PCS_API int pcs_train_tc_destroy(pcs_ctx* ctx, pcs_train_tc* step);
/* the weight-gradient kernel of that step on its own: d_dw[c_out][c_in][k][k] += sum over pixels of x[pixel + tap][c_in] *
 * dy[pixel][c_out] ('same' zero border), k = 5 or 1, for bf16 tensors in the plane-major activation layout
 * [planes][H][W][8 channels] (x_planes <= 16, k * max(32, c_out rounded up to 16) <= 512); fp32 accumulation. */
PCS_API int pcs_train_tc_wgrad(pcs_ctx* ctx, const void* d_x, int x_planes, const void* d_dy, int dy_planes, int H, int W, int k,
                               int c_in, int c_out, float* d_dw);

/* ---- diagnostics ------------------------------------------------------ */
/* fp16 models (PCS_PREC_FP16, the default) store activations with a saturating conversion: a value beyond the fp16 range
 * is stored as +-65504 instead of inf.  The stored activations of the FIRST forward after every pcs_model_load (mode 1,
 * the default; 2 = of every forward, 0 = never) are scanned for such values; pcs_saturation_count returns how many were
 * found since the model was loaded (synchronises the stream).  A non-zero count means the model needs PCS_PREC_BF16 for
 * parity with the reference's fp32 arithmetic (keras models: network.py:75-84). */
PCS_API int pcs_set_saturation_check(pcs_ctx* ctx, int mode);
PCS_API int pcs_saturation_count(pcs_ctx* ctx, uint64_t* out);
/* copies one named internal activation of the last pcs_forward to a float32
 * NHWC host buffer (real channels only); returns the channel count or <0. */
PCS_API int pcs_debug_activation(pcs_ctx* ctx, const char* name, float* h_out, size_t capacity_floats,
                         int32_t* shape4);
/* 1: the next pcs_forward calls also store the activations that the fused kernels
 * normally never write (fcn_skip conv2 at full resolution), so that
 * pcs_debug_activation can return them; 0 (default): production schedule */
PCS_API int pcs_set_keep_activations(pcs_ctx* ctx, int enabled);
/* 1 (or PCSEG_PDL=1 at context creation; default 0, measured neutral on B200): launch the tensor-core kernels with programmatic dependent
 * launch, so that a layer's prologue (barrier init, TMEM allocation, resident weights) overlaps the tail of
 * the previous layer; 0: plain stream-ordered launches */
PCS_API int pcs_set_pdl(pcs_ctx* ctx, int enabled);
/* enable (1) / disable (0) CUDA-event timing of every stage of the next calls */
PCS_API int pcs_set_timing(pcs_ctx* ctx, int enabled);
/* device time in ms of the stages since the last pcs_forward began ("name:ms;"...) */
PCS_API const char* pcs_last_timings(pcs_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PCSEG_B200_H */
