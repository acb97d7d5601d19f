/* libvcb200 -- C ABI of the B200-native detect + ReID hot path for kaylode/vehicle-counting.
 *
 * The reference has no FFI: its boundary is the Python call surface
 *   networks/yolo.py:68-99            YoloBackbone.detect      (detector forward + NMS)
 *   networks/detector.py:36-38        Detector.inference_step
 *   networks/deepsort/deep/feature_extractor.py:26-47   Extractor._preprocess / __call__
 *   networks/deepsort/deep_sort.py:119-129              DeepSort._get_features (ROI crops)
 * Every numeric library call those functions make (cuDNN conv via torch, ATen pointwise,
 * torchvision.ops.nms, cv2.resize) is replaced by one entry point below.  The Python mirror of the
 * reference classes (vehicle_counting_b200/networks/...) binds them with ctypes; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; all pointers are DEVICE pointers owned by the caller
 * unless a parameter is documented as host memory; work is enqueued on the caller's stream, no hidden
 * synchronisation; return 0 on success, a negative VCB_ERR_* otherwise (vcb_last_error_string() has
 * the detail); nothing throws across the ABI; sm_100 only -- any other device is VCB_ERR_ARCH (there
 * is no fallback path).  Activations are NHWC fp16 with an explicit channel pitch so that several
 * producers can write disjoint channel slices of one buffer (concat-free).
 */
#ifndef VCB200_H_
#define VCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vcb_stream_t; /* cudaStream_t */

enum {
  VCB_OK = 0,
  VCB_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  VCB_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
  VCB_ERR_ARCH = -3,        /* device is not sm_100 */
  VCB_ERR_FAULT = -4        /* a kernel recorded a pipeline fault (see vcb_last_fault) */
};

/* VCB_ACT_SILU_TANH: SiLU evaluated as h + h*tanh(h), h = x/2, with tanh.approx.f32 (relative error 2^-11, i.e. the
 * size of the fp16 rounding of the stored result); half the special-function work of VCB_ACT_SILU (ex2 + rcp) */
enum { VCB_ACT_NONE = 0, VCB_ACT_SILU = 1, VCB_ACT_RELU = 2, VCB_ACT_SILU_TANH = 3 };
enum { VCB_RES_NONE = 0, VCB_RES_AFTER_ACT = 1, VCB_RES_BEFORE_ACT = 2 };
enum { VCB_F16 = 0, VCB_F32 = 1 };
/* VCB_A_ROWWIN (stems: 3x3 / stride 1 / pad 1 over cin = cin_pitch = 16): x is W-PADDED, fp16 [n][h][w + 2][16] with pixel
 * (y, x) at column x + 1, zero pad columns and 16 zero elements after the last row.  One K chunk = one filter row = the 64
 * contiguous elements of 4 neighbouring pixels (the 4th has zero weights), fetched by ONE tiled TMA box per filter row through
 * a tensor map whose pixel stride (32 B) is smaller than its 128-byte inner extent: 3 loads per tile instead of 9. */
enum { VCB_A_AUTO = 0, VCB_A_IM2COL_TMA = 1, VCB_A_GATHER = 2, VCB_A_C4 = 3, VCB_A_ROWWIN = 4 };

/* ---- library state ------------------------------------------------------------------------ */
int vcb_init(int device);                 /* selects device, checks sm_100, resolves driver entry points.  One process per GPU:
                                            * a second, different device in the same process is VCB_ERR_INVALID */
const char* vcb_last_error_string(void);  /* thread-local, never NULL */
int vcb_last_fault(int32_t out4[4]);      /* host: {code, block, info0, info1} of the last kernel fault */
int vcb_version(void);
/* library options (process-wide).  "pdl" = 1: convolution launches carry the programmatic-dependent-launch attribute, so the
 * set-up of one conv kernel overlaps the tail of the previous kernel on the stream (the kernels order their global-memory
 * accesses with griddepcontrol.wait); default 1 (measured: +6-7 % at 1 ... 8 frames per call, neutral at 64), $VCB_PDL=0 at vcb_init()
 * turns it off.  "l2_hint" = 1: activation TMA loads carry the
 * evict-first L2 policy and weight loads evict-last ($VCB_L2_HINT).  Unknown names: VCB_ERR_INVALID / -1. */
int vcb_set_option(const char* name, int32_t value);
/* Frames that already live in page-locked host memory go to the device IN PLACE (no gather into a staging buffer): n sources of
 * bytes_each bytes -> dst[i * bytes_each], asynchronous on `stream`; neighbouring sources that are contiguous in host memory are
 * merged into one copy.  Returns 1 when every source is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) and the
 * copies were queued, 0 when at least one is pageable (nothing is queued: the caller gathers into its own pinned buffer, as
 * the reference's lists of cv2 frames need, modules/datasets.py:72-76), negative on error.  The sources must stay untouched until
 * the stream has passed the copies. */
int vcb_h2d_frames_inplace(void* dst, const void* const* srcs, int32_t n, int64_t bytes_each, vcb_stream_t stream);
int vcb_get_option(const char* name);
/* development aid: "prof" = 1 zeroes and enables per-role cycle counters inside the conv kernel (summed over CTAs: CTA
 * lifetime, set-up, producer / MMA / epilogue waits and totals, CTAs, tiles); vcb_read_prof synchronises and copies them
 * to host memory.  Off by default; costs a few clock reads per tile when on. */
int vcb_read_prof(uint64_t out16[16]);

/* ---- K1: implicit-GEMM convolution on tcgen05 (replaces torch Conv2d+BN+SiLU / ReLU, reference
 *      call sites networks/yolo.py:70 [upstream DetectionModel] and deepsort/deep/model.py:83-95) --- */
typedef struct VcbConvDesc {
  int32_t n, h, w;            /* input batch, height, width */
  int32_t cin, cin_pitch;     /* logical input channels; channel pitch of x in elements */
  int32_t cout, cout_pitch;   /* logical output channels; channel pitch of y in elements */
  int32_t kh, kw, stride, pad;
  int32_t act;                /* VCB_ACT_* applied after bias */
  int32_t res_mode;           /* VCB_RES_*: residual tensor has the output's shape */
  int32_t res_pitch;          /* channel pitch of the residual (elements) */
  int32_t out_dtype;          /* VCB_F16 or VCB_F32 */
  int32_t a_mode;             /* VCB_A_*: how the im2col operand reaches shared memory */
  int32_t block_n;            /* 0 = auto; N tile (multiple of 16, <= 256) */
  int32_t stages;             /* 0 = auto; smem pipeline depth */
  int32_t reserved[4];        /* all zero = automatic.  Tuning / debug switches (every non-zero [0] selects the generic epilogue):
                               * [0] 1: direct global stores; 3: skip the epilogue (timing only, output undefined); 4: one accumulator
                               *     stage; 5: no resident weights; 7: one CTA per SM
                               * [1] 1: 8-byte C4 gather; 2: im2col-mode TMA even for 1x1; 3: one accumulation chain (patch mode);
                               *     4: K steps dealt to several accumulators (128-row kernel); 6/7: main-loop timing experiments
                               *     bit 8 (| 0x100), any kernel: walk the tiles from the last to the first, so that a layer starts on the
                               *     part of its input that the previous layer wrote last (still in L2); results are identical
                               * [2] 16/32/64: force the K chunk (swizzle) width of the TMA path
                               * [3] kernel: 1 = 128-row tiles, two CTAs per SM; 2 / 4 = CTA pairs (cta_group::2) with one / two clusters
                               *     per SM pair; 3 = 256-row tiles, two accumulators; 5 = patch mode (3x3/s1/p1, 64-channel chunks).
                               *     Automatic: CTA pairs x2 for 3x3 layers with N >= 128, patch mode for 3x3/s1 layers with N <= 64,
                               *     128-row tiles otherwise. */
} VcbConvDesc;

/* element counts of the packed fp16 weight blob and the padded fp32 bias for this descriptor */
int vcb_conv_packed_sizes(const VcbConvDesc* d, int64_t* weight_halfs, int64_t* bias_floats);
/* w_oihw: fp32 [cout][cin][kh][kw] with BN already folded; bias may be NULL (zeros) */
int vcb_conv_pack_weights(const VcbConvDesc* d, const float* w_oihw, const float* bias, void* w_packed,
                          float* bias_packed, vcb_stream_t stream);
int vcb_conv2d_fwd(const VcbConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                   const void* residual, void* y, vcb_stream_t stream);
/* The same convolution (act = VCB_ACT_NONE, no residual, fp16 output: the pre-BatchNorm tensor of a train-mode BN layer) that ALSO
 * accumulates, from its epilogue, the per-(segment, channel) sum and sum of squares of the stored outputs into
 * sums[seg_of_image[n]][cout][2] (double; the caller zeroes it): vcb_bn_seg_stats_f16's result without reading the tensor again.
 * seg_of_image: int32 [d->n], non-decreasing.  VCB_ERR_INVALID when the geometry does not run the split epilogue (the caller then
 * uses vcb_conv2d_fwd + vcb_bn_seg_stats_f16). */
int vcb_conv2d_fwd_stats(const VcbConvDesc* d, const void* x, const void* w_packed, const float* bias_packed, void* y,
                         const int32_t* seg_of_image, double* sums, vcb_stream_t stream);
/* output spatial size for a descriptor */
int vcb_conv_out_hw(const VcbConvDesc* d, int32_t* ho, int32_t* wo);

/* ---- K2: data-movement / pooling kernels (upstream Upsample, SPPF max-pools, ReID MaxPool2d) --- */
/* uint8 RGB/BGR HWC frames -> fp16 NHWC4 (4th channel zero), value/255 (AutoShape: x/255) */
int vcb_frames_to_f16c4(const uint8_t* frames, void* out, int32_t n, int32_t h, int32_t w, vcb_stream_t stream);
/* uint8 HWC3 frames -> fp16 space-to-depth NHWC16: out[n][y/2][x/2][(dy*2+dx)*3+c] = in/255, channels 12..15 zero
 * (h, w even).  Turns the 6x6/s2/p2 YOLOv5 stem into a 3x3/s1/p1 convolution that the im2col TMA can feed. */
int vcb_frames_to_f16_s2d(const uint8_t* frames, void* out, int32_t n, int32_t h, int32_t w, vcb_stream_t stream);
/* the same pixels in the W-padded layout VCB_A_ROWWIN reads: out fp16 [n][h/2][w/2 + 2][16], pixel (y, x) at column x + 1.
 * Columns 0 and w/2 + 1 and the 16 elements after the last row are never written: the caller zeroes the buffer once. */
int vcb_frames_to_f16_s2d_wpad(const uint8_t* frames, void* out, int32_t n, int32_t h, int32_t w, vcb_stream_t stream);
/* letterbox with an exact 2x reduction ([upstream] AutoShape: cv2.resize(INTER_LINEAR) + copyMakeBorder(114), reached from
 * networks/yolo.py:70): dst uint8 [n][h1][w1][3] = 2x2 box mean with round-half-up of src uint8 [n][h0][w0][3] (h0, w0 even)
 * placed at (top, left), pad_value elsewhere; bit-identical to cv2 for this ratio (1280x720 -> 640x360 inside 384x640) */
int vcb_letterbox_half_u8(const uint8_t* src, int32_t n, int32_t h0, int32_t w0, uint8_t* dst, int32_t h1, int32_t w1,
                          int32_t top, int32_t left, int32_t pad_value, vcb_stream_t stream);
/* letterbox with ANY ratio, bit-identical to cv2.resize(INTER_LINEAR) on uint8 + copyMakeBorder: OpenCV's 11-bit fixed-point
 * bilinear.  The resized image is new_h x new_w at (top, left) of the h1 x w1 frame; xtab / ytab are DEVICE int32 [new_w | new_h][4]
 * = source index 0, source index 1, weight 0, weight 1 (weights scaled by 2048, built on the host the way OpenCV builds them --
 * vehicle_counting_b200/networks/yolo.py cv2_linear_table).  Same-size frame batches of any resolution then need no host resize. */
int vcb_letterbox_bilinear_u8(const uint8_t* src, int32_t n, int32_t h0, int32_t w0, uint8_t* dst, int32_t h1, int32_t w1, int32_t top,
                              int32_t left, int32_t new_h, int32_t new_w, const int32_t* xtab, const int32_t* ytab, int32_t pad_value,
                              vcb_stream_t stream);
/* nearest x2 upsample of a channel slice into a channel slice (nn.Upsample(None, 2, 'nearest')) */
int vcb_upsample2x(const void* src, int32_t src_pitch, void* dst, int32_t dst_pitch, int32_t n, int32_t h,
                   int32_t w, int32_t c, vcb_stream_t stream);
/* SPPF: buf holds x in channels [0,c) of pitch `pitch`; writes mp5(x), mp5^2(x), mp5^3(x) into
 * channels [c,2c), [2c,3c), [3c,4c) (stride 1, pad 2, -inf padding) */
int vcb_sppf_pool(void* buf, int32_t pitch, int32_t n, int32_t h, int32_t w, int32_t c, vcb_stream_t stream);
/* MaxPool2d(k, s, p) NHWC fp16 */
int vcb_maxpool(const void* src, int32_t src_pitch, void* dst, int32_t dst_pitch, int32_t n, int32_t h, int32_t w,
                int32_t c, int32_t k, int32_t s, int32_t p, vcb_stream_t stream);

/* ---- K3: Detect decode + confidence filter + best class (upstream Detect.forward inference branch and
 *      the filtering half of non_max_suppression) ---------------------------------------------- */
typedef struct VcbDetectLevel {
  const void* logits;      /* fp16 or fp32 [n][ny][nx][pitch], channel = anchor*no + field */
  int32_t pitch;
  int32_t ny, nx;
  float stride;
  float anchor_w[3], anchor_h[3];   /* pixels */
} VcbDetectLevel;

typedef struct VcbDetectDesc {
  int32_t n;               /* frames */
  int32_t nc;              /* classes (no = nc + 5) */
  int32_t num_levels;      /* <= 4 */
  int32_t logits_dtype;    /* VCB_F16 / VCB_F32 */
  float conf_thres;
  int32_t max_candidates;  /* capacity per frame of the candidate arrays */
  VcbDetectLevel level[4];
  /* optional class filter (upstream non_max_suppression `classes=`, set by networks/yolo.py:64): a prediction whose BEST class is
   * not in the mask is dropped before the max_nms / max_det cuts, as upstream does.  Bit c of class_mask[c / 32]; classes >= 256
   * are never filtered.  use_class_mask = 0: keep every class. */
  int32_t use_class_mask;
  uint32_t class_mask[8];
} VcbDetectDesc;

/* cand_box: float [n][max_candidates][4] xyxy (inference-image pixels); cand_score: float; cand_cls: int32;
 * cand_index: int32 global prediction index (level order, then anchor, y, x); cand_count: int32 [n]
 * (zeroed by this call).  Candidate order within a frame is unspecified; vcb_nms sorts. */
int vcb_detect_decode(const VcbDetectDesc* d, float* cand_box, float* cand_score, int32_t* cand_cls,
                      int32_t* cand_index, int32_t* cand_count, vcb_stream_t stream);

/* ---- K4: class-aware greedy NMS (upstream non_max_suppression tail + torchvision.ops.nms) ------ */
typedef struct VcbNmsDesc {
  int32_t n;               /* frames */
  int32_t max_candidates;  /* stride of the candidate arrays */
  int32_t max_det;         /* 300 */
  float iou_thres;         /* suppress iff IoU > thr */
  float max_wh;            /* class offset (4096 in v6.0) */
  int32_t max_nms;         /* 30000: keep only the top-scoring max_nms candidates */
  /* scale_coords: (x - pad_x) / gain, clip to [0, w0] x [0, h0]; per frame arrays of length n, may be NULL */
  const float* gain;
  const float* pad_x;
  const float* pad_y;
  const float* w0;
  const float* h0;
} VcbNmsDesc;

/* bytes of the sort workspace vcb_nms needs for n frames of max_candidates slots */
int64_t vcb_nms_workspace_bytes(int32_t n, int32_t max_candidates);
/* sort_ws: workspace of vcb_nms_workspace_bytes() bytes (only touched when a frame has more than 8192
 * candidates; smaller frames sort in shared memory).  det: float [n][max_det][6] = x1,y1,x2,y2,conf,cls
 * sorted by descending score (ties: lower prediction index first); det_count: int32 [n]. */
int vcb_nms(const VcbNmsDesc* d, const float* cand_box, const float* cand_score, const int32_t* cand_cls,
            const int32_t* cand_index, const int32_t* cand_count, uint64_t* sort_ws, float* det,
            int32_t* det_count, vcb_stream_t stream);

/* ---- K5: ROI crop + bilinear resize + normalise (deep_sort.py:119-129 + feature_extractor.py:26-39) --- */
typedef struct VcbRoiDesc {
  int32_t num_rois;
  int32_t out_size;        /* 50 */
  float mean[3], inv_std[3];   /* applied to channel 0,1,2 of the stored frame order */
  int32_t out_channels;    /* channel pitch of `out`: 4 (default when 0), 8 or 16; channels >= 3 are written as zero */
  int32_t num_frames;      /* frames behind `frames` (0 = unchecked): a ROI whose frame index is outside [0, num_frames) yields zeros */
} VcbRoiDesc;
/* frames: uint8 [*][fh][fw][3]; rois: int32 [num_rois][5] = frame, x1, y1, x2, y2 (already int-truncated and
 * clipped, end exclusive); out: fp16 [num_rois][out][out][out_channels] */
int vcb_roi_resize_norm(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois,
                        void* out, vcb_stream_t stream);
/* Fused ReID stem for the folded-BN path (feature_extractor.py:26-39 + model.py:52-60: crop preprocessing, Conv3x3(3->64)+BN+ReLU,
 * MaxPool2d(3, 2, padding=1)) in two launches:
 *   vcb_roi_stem_patches: same crop/resize/normalise arithmetic as vcb_roi_resize_norm (out_size must be 50), written as the
 *     stem's im2col operand: patches fp16 [num_rois][25 blocks][128 rows][32], block (by, bx) = 5x5 pooled pixels, row
 *     ti*11+tj = conv output (10*by-1+ti, 10*bx-1+tj), element (r*3+s)*3+c = input (y+r-1, x+s-1, c); padding zero.
 *   vcb_reid_stem_pool: w_packed fp16 [64][32] (same K order, BN folded, zero padded), bias fp32 [64] ->
 *     out fp16 [num_rois][25][25][64] = maxpool3x3s2p1(relu(conv + bias)); tcgen05 GEMM + pooling in shared memory. */
int vcb_roi_stem_patches(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois,
                         void* patches, vcb_stream_t stream);
int vcb_reid_stem_pool(const void* patches, const void* w_packed, const float* bias, void* out, int32_t num_rois,
                       vcb_stream_t stream);
/* The same stem fed straight from the frames (no im2col operand in HBM): crop + resize + normalise (the arithmetic of
 * vcb_roi_resize_norm, out_size must be 50) happen in shared memory inside the tcgen05 kernel, every CTA owning a contiguous range
 * of ROIs.  w_packed: fp16 [64][32], k = (r*3+s)*3+c for k < 27, k = 27 / 28 hold the conv bias split into an fp16 head and tail
 * (the kernel feeds 1.0 there), k >= 29 zero.  out fp16 [num_rois][25][25][64].
 *   vcb_reid_stem_direct        maxpool3x3s2p1(relu(conv + bias))                                 (BatchNorm folded into w / bias)
 *   vcb_reid_stem_direct_stats  sums[seg][64][2] += sum / sum of squares of conv + bias           (train-mode BN, pass 1)
 *   vcb_reid_stem_direct_bn     maxpool3x3s2p1(relu((conv + bias) * scale[seg] + shift[seg]))     (pass 2; affine from
 *                               vcb_bn_seg_finalize with bias = NULL: the bias is already inside the statistics) */
int vcb_reid_stem_direct(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, const void* w_packed,
                         void* out, vcb_stream_t stream);
int vcb_reid_stem_direct_stats(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, const void* w_packed,
                               const int32_t* seg_of_crop, double* sums, vcb_stream_t stream);
int vcb_reid_stem_direct_bn(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, const void* w_packed,
                            const float* affine, const int32_t* seg_of_crop, void* out, vcb_stream_t stream);
/* float64 xyxy boxes -> the reference's integer crop rectangle (deep_sort.py:78-95), on device.
 * boxes: double [num][4]; frame_of: int32 [num]; rois out: int32 [num][5] */
int vcb_boxes_to_rois(const double* boxes_xyxy, const int32_t* frame_of, int32_t num, int32_t fw, int32_t fh,
                      int32_t* rois, vcb_stream_t stream);

/* ---- K6/K7: ReID tails (model.py:71, :93-95) and train-mode BatchNorm (feature_extractor.py:10-22) ---- */
/* x: fp16 [n][hw][pitch] -> out fp32 [n][c]: mean over hw positions then L2 normalise */
int vcb_avgpool_l2norm(const void* x, int32_t pitch, int32_t n, int32_t hw, int32_t c, float* out,
                       vcb_stream_t stream);
/* per-(segment, channel) batch statistics of fp32 x [rows][c]; seg_row_start: int32 [num_seg+1] row
 * offsets (rows = crops*hw).  Writes scale/shift fp32 [num_seg][c] so that y = x*scale + shift equals
 * BatchNorm2d in training mode (biased variance, eps) with affine gamma/beta. */
int vcb_bn_train_stats(const float* x, int32_t c, const int32_t* seg_row_start, int32_t num_seg,
                       const float* gamma, const float* beta, float eps, float* scale, float* shift,
                       vcb_stream_t stream);
/* y(fp16, pitch) = act(x*scale[seg]+shift[seg] (+ residual)), x fp32 [rows][c]; row_seg: int32 [rows] */
int vcb_bn_apply(const float* x, int32_t c, int32_t rows, const int32_t* row_seg, const float* scale,
                 const float* shift, const void* residual, int32_t res_pitch, int32_t act, void* y,
                 int32_t y_pitch, vcb_stream_t stream);

/* Train-mode BatchNorm over fp16 pre-BN tensors with DEVICE-resident segment tables (replayable inside a CUDA graph): crops are
 * grouped in segments = one reference Extractor call each (all detections of one class in one frame, feature_extractor.py:42-47).
 *   seg_of_crop: int32 [n], values in [0, num_seg]; num_seg marks padding crops (kept out of the real statistics)
 *   seg_crops:   int32 [num_seg + 1], crops per segment
 *   sums:        double [num_seg + 1][c][2] = per (segment, channel) sum and sum of squares; the caller zeroes it before
 *                vcb_bn_seg_stats_f16, which accumulates into it (x: fp16 [n][hw][c], c = 8 * power of two <= 512)
 * vcb_bn_seg_finalize (below): sums -> affine fp32 [num_seg + 1][c][2] = scale, shift; scale = gamma / sqrt(var + eps) (biased
 *                variance), shift = beta - mean * scale (+ an optional conv bias that was NOT part of x: see the fused stem)
 * vcb_bn_seg_apply_f16: y = act(x * scale[seg] + shift[seg] (+ residual)); act = VCB_ACT_NONE | VCB_ACT_RELU; pool != 0
 * additionally applies MaxPool2d(3, 2, padding=1) (model.py:57) to the activated h x w map, y then holds
 * [n][ceil(h/2)][ceil(w/2)] pixels (no residual in that case). */
/* The fused stem under train-mode BatchNorm, in two passes over the same im2col patches (the 50x50x64 pre-BN map never exists):
 *   vcb_reid_stem_stats    accumulates sum / sum of squares of conv + bias per (segment, channel) into sums[num_seg + 1][64][2]
 *   vcb_bn_seg_finalize    sums -> affine fp32 [num_seg + 1][c][2] = scale, shift with the conv bias and the mean folded into the shift
 *   vcb_reid_stem_pool_bn  maxpool3x3s2p1(relu(conv * scale[seg] + shift[seg])) -> out fp16 [num_rois][25][25][64] */
int vcb_reid_stem_stats(const void* patches, const void* w_packed, const float* bias, int32_t num_rois, const int32_t* seg_of_crop,
                        double* sums, vcb_stream_t stream);
int vcb_bn_seg_finalize(const double* sums, const int32_t* seg_crops, int32_t num_seg_plus1, int32_t c, int32_t hw, const float* gamma,
                        const float* beta, const float* bias, float eps, float* affine, vcb_stream_t stream);
int vcb_reid_stem_pool_bn(const void* patches, const void* w_packed, const float* affine, const int32_t* seg_of_crop, void* out,
                          int32_t num_rois, vcb_stream_t stream);
int vcb_bn_seg_stats_f16(const void* x, int32_t c, int32_t hw, int32_t n, const int32_t* seg_of_crop, double* sums, vcb_stream_t stream);
/* vcb_bn_seg_finalize + vcb_bn_seg_apply_f16 in one launch: scale / shift are derived in the kernel from `sums` (same arithmetic), and
 * the residual may itself be a pre-BN tensor -- the downsample branch of a BasicBlock (model.py:19-25, :33-37) -- normalised on the
 * fly with its own statistics / gamma / beta (res_sums != NULL), which removes that branch's apply pass:
 *   y = act(BN(x) + (res_sums ? BN_r(residual) : residual));  x: fp16 [n][hw][c], c = 8 * power of two <= 512; no pooling form. */
int vcb_bn_seg_apply_fused_f16(const void* x, int32_t c, int32_t hw, int32_t n, const int32_t* seg_of_crop, const int32_t* seg_crops,
                               const double* sums, const float* gamma, const float* beta, float eps, const void* residual, int32_t res_pitch,
                               const double* res_sums, const float* res_gamma, const float* res_beta, int32_t act, void* y, int32_t y_pitch,
                               vcb_stream_t stream);
int vcb_bn_seg_apply_f16(const void* x, int32_t c, int32_t h, int32_t w, int32_t n, const int32_t* seg_of_crop, const float* affine,
                         const void* residual, int32_t res_pitch, int32_t act, int32_t pool, void* y, int32_t y_pitch, vcb_stream_t stream);

/* ---- whole-path executor: the calls above, issued once on a capturing stream, become one CUDA
 *      graph that is replayed per batch (tensor maps and shapes are baked in as kernel parameters) --- */
typedef struct VcbGraph VcbGraph;
int vcb_graph_begin(vcb_stream_t stream);                 /* cudaStreamBeginCapture (stream must not be 0) */
int vcb_graph_end(vcb_stream_t stream, VcbGraph** out);   /* end capture + instantiate */
int vcb_graph_launch(VcbGraph* g, vcb_stream_t stream);
int vcb_graph_num_kernels(const VcbGraph* g);             /* kernel nodes in the captured graph */
int vcb_graph_destroy(VcbGraph* g);

#ifdef __cplusplus
}
#endif
#endif /* VCB200_H_ */
