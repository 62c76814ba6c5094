/*
 * qcgraft.h -- C ABI of the B200-native inspection hot path (libqcgraft.so).
 *
 * The reference (JayDS22/Industrial-Quality-Control-Computer-Vision-System-) is pure
 * Python and has no FFI today; its seam is the Python method surface
 * (inference/detector.py:72, models/ensemble.py:71, models/yolo_model.py:84,
 * inference/segmentation.py:39).  Each entry point below replaces the library
 * routine(s) the reference executes at the cited call site; the ctypes binding a
 * maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative qc_status otherwise;
 *     qc_last_error(ctx) gives the message (the Python shim turns it into the
 *     reference's {"error": str} convention, inference/detector.py:124-126).
 *   - all data pointers are DEVICE pointers owned by the caller; the library
 *     allocates nothing on the data path, never synchronises the stream and
 *     never touches host memory (one documented exception: qc_jpeg_decode reads the
 *     compressed byte streams where they are, on the host).  Scratch space comes from
 *     the caller via the *_workspace_bytes() queries.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - no CPU fallback exists: every entry point launches sm_100a kernels.
 */
#ifndef QCGRAFT_H
#define QCGRAFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qc_ctx qc_ctx;

typedef enum {
    QC_OK = 0,
    QC_ERR_INVALID = -1,   /* bad argument */
    QC_ERR_CUDA = -2,      /* CUDA runtime error (message holds cudaGetErrorString) */
    QC_ERR_WORKSPACE = -3, /* workspace too small */
    QC_ERR_UNSUPPORTED = -4
} qc_status;

/* ---- context ------------------------------------------------------------
 * qc_create allocates 8 KB of device memory on `device` (a ring of zeroed work counters for the kernels that hand
 * out work items dynamically; a call takes the next 128-byte line and leaves it at zero); qc_destroy frees it.     */
int qc_create(int device, qc_ctx** out);
void qc_destroy(qc_ctx* ctx);
const char* qc_last_error(const qc_ctx* ctx);
const char* qc_version(void);
int qc_sm_count(const qc_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py "gpu_launches") */
int64_t qc_launch_count(const qc_ctx* ctx);

/* Optional per-kernel timing (measurement only; off by default).  While enabled every entry point records one CUDA
 * event on the caller's stream at entry and one after each kernel it launches (up to 4096 per collection; stream
 * capture must not be active).  qc_profile_collect() waits for the recorded events, writes the kernels' names (static
 * strings) and durations in launch order into names[0..cap) / ms[0..cap), clears the list and returns the number of
 * kernels recorded.  A kernel's duration runs from the previous event on its stream to its own, so it includes the
 * launch gap in front of it -- microseconds for back-to-back launches. */
int qc_profile_enable(qc_ctx* ctx, int on);
int qc_profile_collect(qc_ctx* ctx, const char** names, float* ms, int cap);
/* sizeof() of the ABI structs, for bindings to self-check their layout:
 * 0 qc_pre_plan, 1 qc_yolo_params, 2 qc_seg_region, 3 qc_seg_params, 4 qc_jpeg_info,
 * 5 qc_rules_row, 6 qc_rules_refined, 7 qc_rules_out */
size_t qc_sizeof(int which);

/* ---- K1: letterbox / stretch resize + normalise + NCHW pack -------------
 * Replaces: cv2.cvtColor(BGR2RGB) inference/detector.py:233-234 and
 * models/yolo_model.py:79-80 (-> swap_rb); cv2.resize inference/detector.py:240-242
 * (-> QC_PRE_STRETCH); ultralytics LetterBox + [...,::-1] + NCHW + float()/255 run by
 * models/yolo_model.py:105-110 (-> QC_PRE_LETTERBOX*); and, with a mean/std LUT, the
 * ToTensor+Normalize of models/resent_model.py:59-66.
 */
typedef enum { QC_PRE_STRETCH = 0, QC_PRE_LETTERBOX = 1, QC_PRE_LETTERBOX_AUTO = 2 } qc_pre_mode;
typedef enum { QC_DT_F32 = 0, QC_DT_BF16 = 1 } qc_dtype;

typedef struct {
    int32_t src_h, src_w;       /* source image (uint8, 3 channels, HWC)             */
    int32_t new_h, new_w;       /* size after the bilinear resize ("new_unpad")      */
    int32_t top, left;          /* border added before the resized image             */
    int32_t dst_h, dst_w;       /* output plane size (new + borders)                 */
    float gain;                 /* min(H/h0, W/w0) as float32 (scale_boxes)          */
    int32_t pad_x, pad_y;       /* scale_boxes' pad: round((W-w0*gain)/2-0.1), ...   */
} qc_pre_plan;

/* Host-side geometry, exactly ultralytics' LetterBox / scale_boxes arithmetic
 * (Python round() == round-half-even).  mode STRETCH: new = dst = (H,W), no border. */
int qc_pre_make_plan(int mode, int src_h, int src_w, int H, int W, int stride, qc_pre_plan* plan);

/* src: n images, uint8 HWC (3 ch), row pitch src_row_stride bytes, image pitch
 * src_img_stride bytes.  lut: device float32 [3][256], lut[c][v] = the normalised value of
 * byte v in OUTPUT channel c (the caller builds it with the reference's own float32 ops, so
 * the result is bit-exact; bf16 output = RNE of that float32).  dst: [n,3,dst_h,dst_w]
 * contiguous, dtype out_dtype.  pad_value: border byte (114).  flags: QC_PRE_SWAP_RB =
 * output channel c reads source channel 2-c; QC_PRE_LUT_UNIFORM = the caller asserts the three
 * LUT rows are identical (plain /255), which lets the kernel keep one shared-memory table.  */
#define QC_PRE_SWAP_RB 1
#define QC_PRE_LUT_UNIFORM 2
int qc_preprocess(qc_ctx* ctx, const uint8_t* src, int n, int64_t src_img_stride,
                  int32_t src_row_stride, const qc_pre_plan* plan, const float* lut,
                  int pad_value, int flags, int out_dtype, void* dst, void* stream);

/* ---- K2: YOLOv8 DFL decode + confidence filter + class-aware NMS ----------
 * Replaces ultralytics Detect._inference + ops.non_max_suppression (torchvision.ops.nms on
 * class-offset boxes) + ops.scale_boxes, all run inside the model call at
 * models/yolo_model.py:105-110; outputs are what models/yolo_model.py:133-135 reads.
 */
#define QC_MAX_LEVELS 4
typedef struct {
    int32_t n;                  /* images                                             */
    int32_t nc;                 /* classes (<= 80)                                     */
    int32_t reg_max;            /* DFL bins (16)                                       */
    int32_t num_levels;         /* <= QC_MAX_LEVELS                                    */
    int32_t level_h[QC_MAX_LEVELS], level_w[QC_MAX_LEVELS];
    float level_stride[QC_MAX_LEVELS];                 /* 8,16,32                      */
    int64_t level_batch_stride[QC_MAX_LEVELS];         /* elements between images      */
    int64_t level_chan_stride[QC_MAX_LEVELS];          /* elements between channels    */
    int32_t in_dtype;           /* qc_dtype of the head tensors                        */
    double conf_thres, iou_thres;  /* compared as float32 / double like the reference  */
    int32_t max_det, max_nms;   /* 300, 30000                                          */
    float max_wh;               /* 7680 class offset                                   */
    /* scale_boxes: network input (H,W) -> original (h0,w0) */
    float gain; int32_t pad_x, pad_y; int32_t orig_h, orig_w;
    int32_t scale_boxes;        /* 0: leave boxes in network coordinates               */
} qc_yolo_params;

size_t qc_yolo_workspace_bytes(const qc_yolo_params* p);

/* levels[l]: device pointer to level l, [n, 4*reg_max+nc, h_l, w_l] addressed through
 * the strides above.  out_det [n][max_det][6] float32 rows (x1,y1,x2,y2,conf,cls) in
 * descending score order; out_anchor [n][max_det] int32 anchor index of each kept box;
 * out_count [n] int32 kept boxes; out_ncand [n] int32 candidates that passed conf
 * (may be NULL).  Rows >= count are zero.                                              */
int qc_yolo_decode_nms(qc_ctx* ctx, const void* const* levels, const qc_yolo_params* p,
                       void* workspace, size_t workspace_bytes, float* out_det,
                       int32_t* out_anchor, int32_t* out_count, int32_t* out_ncand, void* stream);

/* NMS-only entry (the torchvision.ops.nms + [:max_det] step on caller-supplied candidates):
 * boxes [n][cap][4] xyxy float32 (un-offset), scores [n][cap], cls [n][cap] int32,
 * counts [n].  out_keep [n][max_det] int32 = candidate indices kept, score-descending
 * (ties: lower index first); out_count [n].                                             */
size_t qc_nms_workspace_bytes(int n, int cap);
int qc_nms(qc_ctx* ctx, const float* boxes, const float* scores, const int32_t* cls,
           const int32_t* counts, int n, int cap, double iou_thres, float max_wh, int max_det,
           void* workspace, size_t workspace_bytes, int32_t* out_keep, int32_t* out_count,
           void* stream);

/* ---- K3: ROI-align crop-resize into classifier inputs ----------------------
 * Implements the hand-off models/ensemble.py:129-142,181-195 only mocks: each kept box is
 * cropped/resampled to [3,out_h,out_w] with torchvision.ops.roi_align semantics
 * (spatial_scale=1) and normalised through the same LUT mechanism as K1.
 * imgs: n images uint8 HWC; rois: device float32 [K][5] = (img, x1, y1, x2, y2).
 * sampling_ratio 1..2, out_w % 4 == 0, out_w <= 224: two launches per call -- k3_roi_win (boxes whose bins are <= 1.95
 * source pixels with every sample inside the frame) and k3_roi_tma (whatever is left); other shapes: k3_roi_sep /
 * k3_roi_align.  Results do not depend on which kernel took a box beyond float32 rounding (<= 1e-3 absolute).
 * Work items are handed out in table order: when the boxes differ much in size, list the biggest first (with
 * qc_roi_align_rows every ROI carries its destination row, so the order of the table is free).                     */
int qc_roi_align(qc_ctx* ctx, const uint8_t* imgs, int n, int h, int w, int64_t img_stride,
                 int32_t row_stride, const float* rois, int K, int out_h, int out_w,
                 int sampling_ratio, int aligned, int swap_rb, const float* norm,
                 int out_dtype, void* dst, void* stream);
/* Same, with a destination row per ROI: ROI k is written to dst row dst_rows[k] (device int32 [K]; NULL = k), so one
 * launch can fill the static input buffers of several classifier graphs that sit at fixed places of one arena.     */
int qc_roi_align_rows(qc_ctx* ctx, const uint8_t* imgs, int n, int h, int w, int64_t img_stride,
                      int32_t row_stride, const float* rois, const int32_t* dst_rows, int K, int out_h, int out_w,
                      int sampling_ratio, int aligned, int swap_rb, const float* norm, int out_dtype, void* dst,
                      void* stream);
/* norm: device float32 [3][3] rows (div, mean, std) per OUTPUT channel:
 * out = (roi_align_value / div - mean) / std in float32, i.e. /255 then Normalize.
 * dst: [K,3,out_h,out_w] contiguous.                                                    */

/* ---- K3b: whole-frame classifier input, PIL-exact ------------------------------
 * Replaces models/resent_model.py:57-66,114-130: Image.fromarray -> transforms.Resize((224,224))
 * (Pillow's antialiased BILINEAR resampler: 22-bit integer coefficients, horizontal then vertical
 * pass, uint8 intermediates) -> ToTensor -> Normalize.  The resized bytes are bit-exact with
 * Pillow; the float32 output is bit-exact with torchvision's transform (bf16: its RNE rounding).
 * Tables depend only on the geometry: build them once on the host, keep one copy on the device.
 *   qc_pil_resize_table_bytes   size of the int32 table for (src_h,src_w)->(dst_h,dst_w)
 *   qc_pil_resize_make_tables   fills a HOST buffer of that size (Pillow's precompute_coeffs +
 *                               normalize_coeffs_8bpc, float64)
 *   qc_pil_resize               imgs: n images uint8 HWC (strides in bytes); dev_tables = device copy
 *                               of host_tables; norm as for qc_roi_align; dst [n,3,dst_h,dst_w].   */
size_t qc_pil_resize_table_bytes(int src_h, int src_w, int dst_h, int dst_w);
int qc_pil_resize_make_tables(int src_h, int src_w, int dst_h, int dst_w, int32_t* host_tables);
int qc_pil_resize(qc_ctx* ctx, const uint8_t* imgs, int n, int h, int w, int64_t img_stride,
                  int32_t row_stride, const int32_t* dev_tables, const int32_t* host_tables,
                  int swap_rb, const float* norm, int out_dtype, void* dst, void* stream);

/* ---- K4: Otsu + 3x3 cross morphology + union-find CCL + stats ---------------
 * Replaces, per region: cv2.cvtColor(RGB2GRAY) inference/segmentation.py:66-69; the Otsu
 * threshold + polarity compare of :212-234; cv2.morphologyEx CLOSE/OPEN with the 3x3 cross
 * :284-286 (iters=2 == skimage disk(2), :458-460); cv2.connectedComponents + bincount +
 * isin of :464-484; and np.sum(mask>0) :98 (sum of kept areas).
 */
typedef struct {
    /* set by the caller */
    int32_t img;                /* image index                                         */
    int32_t x, y, w, h;         /* region rectangle inside the image (w,h >= 1)        */
    int32_t polarity;           /* 1: light defects (v > T) ; 0: dark defects           */
    int32_t min_area;           /* components with area < min_area are removed          */
    int32_t max_comp;           /* stats rows reserved for this region (incl. row 0)    */
    double thr_scale;           /* 1.0: cv2 BINARY / BINARY_INV (v > t | v <= t);
                                   else the reference's float compare of :226-234,
                                   v > t*s (light) | v < t*s (dark), s = class factor   */
    /* filled by qc_segment_plan() -- offsets of this region in outputs / workspace      */
    int64_t out_offset;         /* first pixel in out_mask / out_labels (pitch = w)     */
    int32_t stats_offset;       /* first row in out_stats                               */
    int32_t band_offset, chunk_offset, bits_offset, blk_offset;   /* internal           */
} qc_seg_region;

typedef struct {
    /* set by the caller */
    int32_t n, h, w, channels;  /* image batch: uint8 [n,h,w] gray or [n,h,w,3] RGB      */
    int64_t img_stride; int32_t row_stride;            /* bytes                         */
    int32_t morph_iters;        /* 0..4: close then open, `iters` iterations each       */
    int32_t swap_rb;            /* channels==3: 1 = pixels are stored B,G,R relative to what
                                   the reference's RGB2GRAY sees (inference/detector.py:234
                                   swaps before inference/segmentation.py:67)            */
    /* filled by qc_segment_plan() */
    int32_t num_regions;
    int64_t total_pixels;       /* size of out_mask / out_labels (elements)             */
    int32_t total_stats_rows;   /* rows of out_stats                                    */
    int32_t total_bits_words, total_blocks, total_bands, total_chunks;   /* internal    */
    int32_t max_region_w, max_region_h;
} qc_seg_params;

/* Host-side: validates the regions against p->{n,h,w} and fills every offset/total above.
 * The caller then uploads `regions` to the device and sizes outputs/workspace from `p`.  */
int qc_segment_plan(qc_seg_region* regions, int num_regions, qc_seg_params* p);

size_t qc_segment_workspace_bytes(const qc_seg_params* p);

/* regions: DEVICE array [num_regions] (planned on the host, then uploaded).  out_mask uint8
 * {0,255} and out_labels int32, both [total_pixels], region r occupying
 * [out_offset, out_offset + w*h) row-major with pitch w.  Labels follow
 * cv2.connectedComponents' numbering (8-connectivity, ordered by first 2x2 block); removed
 * (small) components are zeroed but the numbering is NOT compacted.
 * out_ncomp [num_regions]: number of labels incl. background (cv2's return value).
 * out_stats [total_stats_rows][5] int32 rows (x,y,w,h,area) at stats_offset + label
 * (label >= 1; row 0 of each region is zero; rows at and past out_ncomp are not written); labels >= max_comp are
 * counted but not recorded (and never removed by min_area).  out_thr [num_regions] int32 Otsu threshold;
 * out_area [num_regions] int64 kept foreground pixels (== np.sum(mask>0)).
 * out_mask / out_labels may be NULL to skip those writes.                               */
int qc_segment(qc_ctx* ctx, const uint8_t* imgs, const qc_seg_params* p,
               const qc_seg_region* regions, void* workspace, size_t workspace_bytes,
               uint8_t* out_mask, int32_t* out_labels, int32_t* out_ncomp, int32_t* out_stats,
               int32_t* out_thr, int64_t* out_area, void* stream);

/* ---- K7: the reference's own per-class segmentation methods (row f3) -----------------
 * qc_segment() above is the north-star composite (Otsu -> morphology -> CCL).  qc_segment_ref() runs, on the same
 * region tables / workspace / outputs, the two methods inference/segmentation.py:185-195 actually routes to:
 *
 *   QC_SEG_REF_THRESHOLD  _threshold_segmentation :197-246 + _clean_mask :449-462, literally:
 *       skimage gaussian(sigma=1): img_as_float (v * (1/255)) then scipy's 9-tap float64 correlate1d along axis 0,
 *       then axis 1, mode='nearest' (operation order of ni_filters.c, no fma)
 *       -> threshold_otsu: np.histogram(256 bins over [min,max]) with numpy's edge-corrected binning, float32
 *          counts/weights, float64 means, first argmax, bin CENTRE; a constant image returns its value
 *       -> x region.thr_scale (class factor) -> blurred < T (polarity 0) | blurred > T (polarity 1)
 *       -> components with area < region.min_area removed (cv2.connectedComponents, 8-conn; reference: 10)
 *       -> scipy.ndimage.binary_fill_holes (4-connected background reachable from the ROI border stays)
 *       -> closing then opening with the 3x3 cross x morph_iters (skimage disk(2) == 2 iterations).
 *     Outputs: out_area = np.sum(mask>0); out_mask; out_labels = 1 on the mask (the reference does not label the
 *     final mask); out_ncomp = 2 if the mask is non-empty else 1; stats row 1 = extent (x,y,w,h) + area of the WHOLE
 *     mask (what inference/postprocess.py:784-790 reads); out_thr = index of the Otsu bin.
 *   QC_SEG_REF_ADAPTIVE   _adaptive_segmentation :248-298:
 *       cv2.adaptiveThreshold(ADAPTIVE_THRESH_GAUSSIAN_C, polarity 0: BINARY_INV | 1: BINARY,
 *       block = max(11, min(h,w)/10) | 1, C = adaptive_c): mean = cvRound(float32 separable Gaussian, replicate
 *       border) with cv2's summation order (fused multiply-adds of its AVX2 path: columns < w&~7 fused in both
 *       passes, the next group of 4 fused in the row pass only, the last w%4 columns unfused), then
 *       src - mean > -C (<= -C for INV) -> 3x3 cross close/open x morph_iters (reference: 1) -> min-area filter
 *       (reference: 20) with the full K4 outputs (labels in cv2 numbering, per-component stats).
 *     akern: device float32 table of cv2.getGaussianKernel(bs, -1, CV_32F), row (bs-11)/2 holding the bs taps of
 *     block size bs, row pitch akern_pitch floats, rows up to akern_max_block (the caller fills it with cv2's own
 *     numbers, like the K1 LUT; only needed for QC_SEG_REF_ADAPTIVE).
 *
 * out_sums (may be NULL): int64 [num_regions][4].  QC_SEG_REF_THRESHOLD: (sum of gray over the mask, sum of gray, sum of
 * gray^2, pixels of the ROI) -- the integers behind inference/segmentation.py:527-555's confidence (means and std of the
 * ROI).  QC_SEG_REF_ADAPTIVE: (Canny edge pixels on the mask, Canny edge pixels, 0, pixels of the ROI) with
 * cv2.Canny(roi, 50, 150) computed on device bit for bit (K9) -- the integers behind :557-578's edge-overlap confidence.
 * out_contours (may be NULL): int64 [num_regions][4] = (n_axis, n_diag, n_contours, 0) of the final mask's external
 * contours with cv2.contourArea >= 10 -- inference/segmentation.py:486-525 without tracing a contour: perimeter
 * (sum of cv2.arcLength) = n_axis + sqrt(2) * n_diag, contour count = n_contours (K8, csrc/k7_refseg.cuh).
 * workspace: qc_segment_ref_workspace_bytes(p, method).                                                    */
typedef enum { QC_SEG_COMPOSITE = 0, QC_SEG_REF_THRESHOLD = 1, QC_SEG_REF_ADAPTIVE = 2 } qc_seg_method;

size_t qc_segment_ref_workspace_bytes(const qc_seg_params* p, int method);

int qc_segment_ref(qc_ctx* ctx, int method, const uint8_t* imgs, const qc_seg_params* p,
                   const qc_seg_region* regions, const float* akern, int akern_pitch, int akern_max_block,
                   int adaptive_c, void* workspace, size_t workspace_bytes, uint8_t* out_mask,
                   int32_t* out_labels, int32_t* out_ncomp, int32_t* out_stats, int32_t* out_thr,
                   int64_t* out_area, int64_t* out_sums, int64_t* out_contours, void* stream);

/* ---- SPC per-batch statistics (payload of the one NCCL all-reduce) ---------
 * Device-side reduction of K2's outputs into the 9-vector derived from
 * analytics/sec_analysis.py:105-147: [n_images, sum defects, sum defects^2, sum conf,
 * n_detections, n_critical, n_major, n_minor, sum bbox area] (float64).  Severity follows
 * models/yolo_model.py:162-172 on int-truncated boxes.  out9 is overwritten.            */
int qc_spc_stats(qc_ctx* ctx, const float* det, const int32_t* count, int n, int max_det,
                 double* out9, void* stream);

/* ---- K6: same-class overlap clustering of the post-NMS rule engine (row f1) ------
 * The quadratic step of inference/postprocess.py:233-303: per image, detections of one class are linked when
 * `1.0 - IoU <= 1.0 - overlap_threshold` (IoU = integer intersection / integer union in float64, :859-877) and
 * labelled like DBSCAN(metric='precomputed', eps = 1 - overlap_threshold, min_samples = 1) labels `1 - iou_matrix`
 * with its zero-IoU diagonal: -1 for rows without a partner (they all share the noise label), otherwise the index
 * of the connected component among the components of that class, in order of first member.
 * box [n][cap][4] int32 (x1,y1,x2,y2; width/height = differences), cls [n][cap] int32 (>= 0),
 * keep [n][cap] uint8 or NULL (rows filtered out before clustering), count [n]; out_label [n][cap] int32
 * (-2 for rows >= count or not kept).  cap <= 2048; box 16-byte aligned.                                      */
int qc_rules_cluster(qc_ctx* ctx, const int32_t* box, const int32_t* cls, const uint8_t* keep,
                     const int32_t* count, int n, int cap, double overlap_threshold, int32_t* out_label,
                     void* stream);

/* ---- K6 (whole chain): refine -> filter -> merge -> validate of the post-NMS rule engine (row f1) ----
 * Replaces the detection chain of PostProcessor.process_results, inference/postprocess.py:64-81:
 * _refine_detections :111-184 (+ _refine_bbox_from_mask :764-812, _refine_severity :814-840),
 * _filter_detections :186-231, _merge_overlapping_detections :233-303, _merge_detection_cluster
 * :305-358, _validate_detections :360-410, _calculate_validation_score :412-464 -- for a batch of
 * images, one table of `cap` rows per image.  float64 results are bit-identical with the reference's
 * Python / numpy arithmetic (np.mean in pairwise order, IEEE divisions).
 *   rows[n][cap]         input: what the detector, the ensemble and the segmentation produced per detection
 *   count[n], shapes[n][2] = (height, width) of each image
 *   elongated_classes    bit c set = class id c is one of the reference's ('crack', 'scratch') (:833)
 *   out_refined[n][cap]  per input row: refined bbox / severity / area and whether the filter kept it
 *   out_rows[n][cap]     the merged list in the reference's order (out_count[n] rows); flags bit 1 =
 *                        validation_score > 0.3 (the row survives), bit 0 = area_consistency_warning
 *   out_member[n][cap]   per input row: index of the out row it went into, -1 = filtered out
 */
typedef struct {
    int32_t x1, y1, x2, y2;    /* detection bbox (ints of models/yolo_model.py:139)                    */
    int32_t cls, severity;     /* class id; 1 minor, 2 major, 3 critical                               */
    int32_t has_seg, reserved;
    int32_t ext[4];            /* mask extent y_min, y_max, x_min, x_max in frame coordinates; ext[0] < 0 = empty mask */
    int64_t area_pixels;       /* segmentation area (has_seg)                                          */
    double conf;               /* ensemble_confidence                                                  */
    double seg_conf, area_pct, compactness;
} qc_rules_row;
typedef struct {
    int32_t x1, y1, x2, y2, severity, kept;
    int64_t area_pixels;
} qc_rules_refined;
typedef struct {
    int32_t base, merged_from; /* input row whose other fields the row inherits; number of members      */
    int32_t x1, y1, x2, y2, severity, flags;
    int64_t area_pixels;
    double conf, validation_score;
} qc_rules_out;
int qc_rules_apply(qc_ctx* ctx, const qc_rules_row* rows, const int32_t* count, const int32_t* shapes, int n,
                   int cap, double min_conf, double max_area_pct, uint64_t elongated_classes,
                   double overlap_threshold, qc_rules_refined* out_refined, qc_rules_out* out_rows,
                   int32_t* out_member, int32_t* out_count, void* stream);

/* ---- K10: batched baseline-JPEG decode (row f4, input side) ---------------------------------
 * Replaces: cv2.imdecode(buf, cv2.IMREAD_COLOR) at inference/detector.py:214,229 and
 * models/yolo_model.py:74-76 (the 1-D uint8 upload buffer of app.py:190-194), plus -- with rgb=1 --
 * the cv2.cvtColor(BGR2RGB) that follows it (inference/detector.py:233-234).  Output bytes are
 * identical to cv2's (libjpeg-turbo 3.x defaults: islow IDCT, fancy upsampling, integer colour
 * tables).  THE ONE EXCEPTION to the "device pointers only" convention: the JPEG byte streams,
 * `infos`, `dst_offsets`, `dst_row_strides` and `host_staging` are HOST memory; the Huffman stage
 * runs on `host_threads` host threads inside the call (it is serial per restart interval), then one
 * cudaMemcpyAsync moves the coefficient blocks and two kernels rebuild the pixels.  `host_staging`
 * (pinned for an asynchronous copy) must stay untouched until the stream has passed the copy.
 * Supported: 8-bit sequential Huffman (SOF0/SOF1), grayscale or YCbCr 4:4:4 / 4:2:2 / 4:2:0, one
 * interleaved scan, restart intervals.  Anything else (progressive, arithmetic, CMYK / RGB-coded,
 * other samplings, EXIF orientation != 1) -> QC_ERR_UNSUPPORTED; the library has no other decoder.
 */
typedef enum { QC_JPEG_GRAY = 0, QC_JPEG_444 = 1, QC_JPEG_422 = 2, QC_JPEG_420 = 3 } qc_jpeg_subsampling;
typedef struct {
    int32_t width, height;
    int32_t components;        /* 1 or 3 (the decoded frame always has 3 channels, like IMREAD_COLOR) */
    int32_t subsampling;       /* qc_jpeg_subsampling */
    int32_t restart_interval;  /* MCUs between RSTn markers, 0 = none */
    int32_t reserved;
    int64_t blocks;            /* 8x8 coefficient blocks incl. MCU padding (sizes staging / workspace) */
} qc_jpeg_info;
/* host only; ctx may be NULL (then no message is kept) */
int qc_jpeg_info_host(qc_ctx* ctx, const uint8_t* data, size_t len, qc_jpeg_info* info);
/* host only: the entropy stage of ONE image on its own -- coefs[blocks*64] int16 in natural (row-major) order, not
 * dequantised, laid out [component][block row][block column][64] over the MCU-padded planes (Y has
 * mcus_x*h_samp x mcus_y*v_samp blocks); quant[components*64] in natural order.  What qc_jpeg_decode stages. */
int qc_jpeg_entropy_host(qc_ctx* ctx, const uint8_t* data, size_t len, int16_t* coefs, uint16_t* quant);
size_t qc_jpeg_staging_bytes(const qc_jpeg_info* infos, int n);     /* host staging buffer            */
size_t qc_jpeg_workspace_bytes(const qc_jpeg_info* infos, int n);   /* device workspace               */
/* image i is written as packed 3-byte pixels at dst + dst_offsets[i] with dst_row_strides[i] bytes per
 * row (>= 3*width): BGR like cv2.imdecode, or RGB when rgb != 0 */
int qc_jpeg_decode(qc_ctx* ctx, const uint8_t* const* jpegs, const size_t* lens, int n,
                   const qc_jpeg_info* infos, int host_threads, void* host_staging, size_t staging_bytes,
                   void* workspace, size_t workspace_bytes, uint8_t* dst, const int64_t* dst_offsets,
                   const int32_t* dst_row_strides, int rgb, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QCGRAFT_H */
