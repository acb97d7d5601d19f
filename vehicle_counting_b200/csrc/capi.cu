// extern "C" surface of libvcb200.so (declared in include/vcb200.h) + library state + CUDA-graph capture.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "vcb_internal.h"

namespace vcb {

// other translation units
int frames_to_f16c4(const uint8_t*, void*, int, int, int, cudaStream_t);
int frames_to_f16_s2d(const uint8_t*, void*, int, int, int, cudaStream_t);
int frames_to_f16_s2d_wpad(const uint8_t*, void*, int, int, int, cudaStream_t);
int letterbox_half(const uint8_t*, int, int, int, uint8_t*, int, int, int, int, int, cudaStream_t);
int letterbox_bilinear(const uint8_t*, int, int, int, uint8_t*, int, int, int, int, int, int, const int*, const int*, int, cudaStream_t);
int upsample2x(const void*, int, void*, int, int, int, int, int, cudaStream_t);
int sppf_pool(void*, int, int, int, int, int, cudaStream_t);
int maxpool(const void*, int, void*, int, int, int, int, int, int, int, int, cudaStream_t);
int avgpool_l2norm(const void*, int, int, int, int, float*, cudaStream_t);
int bn_train_stats(const float*, int, const int*, int, const float*, const float*, float, float*, float*, cudaStream_t);
int bn_apply(const float*, int, int, const int*, const float*, const float*, const void*, int, int, void*, int, cudaStream_t);
int bn_seg_stats_f16(const void*, int, int, int, const int*, double*, cudaStream_t);
int bn_seg_apply_f16(const void*, int, int, int, int, const int*, const float*, const void*, int, int, int, void*, int, cudaStream_t);
int bn_seg_apply_fused_f16(const void*, int, int, int, const int*, const int*, const double*, const float*, const float*, float, const void*, int,
                           const double*, const float*, const float*, int, void*, int, cudaStream_t);
int detect_decode(const VcbDetectDesc&, float*, float*, int*, int*, int*, cudaStream_t);
int nms(const VcbNmsDesc&, const float*, const float*, const int*, const int*, const int*, unsigned long long*, float*, int*,
        cudaStream_t);
long long nms_workspace_bytes(int n, int max_candidates);
int roi_resize_norm(const VcbRoiDesc&, const uint8_t*, int, int, const int*, void*, cudaStream_t);
int boxes_to_rois(const double*, const int*, int, int, int, int*, cudaStream_t);
int roi_stem_patches(const VcbRoiDesc&, const uint8_t*, int, int, const int*, void*, cudaStream_t);
int reid_stem_pool(const void*, const void*, const float*, void*, int, cudaStream_t);
int reid_stem_stats(const void*, const void*, const float*, int, const int*, double*, cudaStream_t);
int reid_stem_pool_bn(const void*, const void*, const float*, const int*, void*, int, cudaStream_t);
int reid_stem_direct(const VcbRoiDesc&, const uint8_t*, int, int, const int*, const void*, void*, cudaStream_t);
int reid_stem_direct_stats(const VcbRoiDesc&, const uint8_t*, int, int, const int*, const void*, const int*, double*, cudaStream_t);
int reid_stem_direct_bn(const VcbRoiDesc&, const uint8_t*, int, int, const int*, const void*, const float*, const int*, void*, cudaStream_t);
int bn_seg_finalize(const double*, const int*, int, int, int, const float*, const float*, const float*, float, float*, cudaStream_t);

static thread_local char g_err[512] = "";

State& state() {
  static State s;
  return s;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VCB_OK;
  return set_error(VCB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int require_init() {
  if (!state().initialised) return set_error(VCB_ERR_INVALID, "vcb_init() has not been called");
  return VCB_OK;
}

}  // namespace vcb

using namespace vcb;

struct VcbGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int num_kernels = 0;
};

extern "C" {

int vcb_version(void) { return 100; }

const char* vcb_last_error_string(void) { return g_err; }

int vcb_init(int device) {
  State& s = state();
  if (s.initialised && s.device == device) return VCB_OK;
  // one process per GPU: function attributes (opt-in shared memory), the SM count and the fault record are per device, and
  // the library keeps one copy of each -- a second device in the same process is refused instead of half-working
  if (s.initialised) return set_error(VCB_ERR_INVALID, "libvcb200 is bound to device %d in this process (one process per GPU); cannot switch to device %d", s.device, device);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return check_cuda(e, "cudaSetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDeviceProperties");
  if (prop.major != 10) return set_error(VCB_ERR_ARCH, "device %d is sm_%d%d; libvcb200 runs on sm_100 only (no fallback path)", device, prop.major, prop.minor);
  s.num_sms = prop.multiProcessorCount;
  cudaDriverGetVersion(&s.driver_version);
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  s.encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not found");
  s.encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  if (s.fault_host == nullptr) {
    e = cudaHostAlloc(reinterpret_cast<void**>(&s.fault_host), sizeof(KernelFault), cudaHostAllocMapped);
    if (e != cudaSuccess) return check_cuda(e, "cudaHostAlloc(fault record)");
    memset(s.fault_host, 0, sizeof(KernelFault));
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&s.fault_dev), s.fault_host, 0);
    if (e != cudaSuccess) return check_cuda(e, "cudaHostGetDevicePointer");
  }
  s.device = device;
  s.initialised = true;
  if (const char* e_pdl = getenv("VCB_PDL")) s.pdl = atoi(e_pdl) != 0 ? 1 : 0;
  if (const char* e_l2 = getenv("VCB_L2_HINT")) s.l2_hint = atoi(e_l2) != 0 ? 1 : 0;
  if (const char* e_sp = getenv("VCB_EPI_SPLIT")) s.epi_split = atoi(e_sp) != 0 ? 1 : 0;
  return VCB_OK;
}

int vcb_h2d_frames_inplace(void* dst, const void* const* srcs, int32_t n, int64_t bytes_each, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  if (!dst || !srcs || n < 0 || bytes_each <= 0) return set_error(VCB_ERR_INVALID, "h2d_frames_inplace: bad argument");
  for (int i = 0; i < n; ++i) {
    if (!srcs[i]) return set_error(VCB_ERR_INVALID, "h2d_frames_inplace: null source");
    cudaPointerAttributes a;
    const cudaError_t e = cudaPointerGetAttributes(&a, srcs[i]);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    if (a.type != cudaMemoryTypeHost) return 0;                 // pageable (cudaMemoryTypeUnregistered) or not host memory at all
  }
  for (int i = 0; i < n;) {
    int j = i + 1;                                               // [i, j): one contiguous run of sources
    while (j < n && static_cast<const char*>(srcs[j]) == static_cast<const char*>(srcs[j - 1]) + bytes_each) ++j;
    const cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)i * bytes_each, srcs[i], (size_t)(j - i) * bytes_each,
                                          cudaMemcpyHostToDevice, (cudaStream_t)st);
    if (e != cudaSuccess) return check_cuda(e, "cudaMemcpyAsync(h2d_frames_inplace)");
    i = j;
  }
  return 1;
}

int vcb_set_option(const char* name, int32_t value) {
  if (!name) return set_error(VCB_ERR_INVALID, "null option name");
  if (strcmp(name, "pdl") == 0) { state().pdl = value != 0 ? 1 : 0; return VCB_OK; }
  if (strcmp(name, "l2_hint") == 0) { state().l2_hint = value != 0 ? 1 : 0; return VCB_OK; }
  if (strcmp(name, "epi_split") == 0) { state().epi_split = value != 0 ? 1 : 0; return VCB_OK; }
  if (strcmp(name, "prof") == 0) {
    State& s = state();
    if (value && s.prof_dev == nullptr) {
      const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s.prof_dev), 16 * sizeof(unsigned long long));
      if (e != cudaSuccess) return check_cuda(e, "cudaMalloc(prof)");
    }
    if (s.prof_dev) cudaMemset(s.prof_dev, 0, 16 * sizeof(unsigned long long));
    s.prof_on = value != 0 ? 1 : 0;
    return VCB_OK;
  }
  return set_error(VCB_ERR_INVALID, "unknown option '%s'", name);
}
int vcb_read_prof(uint64_t out16[16]) {
  State& s = state();
  if (!out16) return set_error(VCB_ERR_INVALID, "null output");
  if (!s.prof_dev) { memset(out16, 0, 16 * sizeof(uint64_t)); return VCB_OK; }
  return check_cuda(cudaMemcpy(out16, s.prof_dev, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost), "cudaMemcpy(prof)");
}
int vcb_get_option(const char* name) {
  if (name && strcmp(name, "pdl") == 0) return state().pdl;
  if (name && strcmp(name, "l2_hint") == 0) return state().l2_hint;
  if (name && strcmp(name, "epi_split") == 0) return state().epi_split;
  return -1;
}

int vcb_last_fault(int32_t out4[4]) {
  State& s = state();
  if (!out4) return set_error(VCB_ERR_INVALID, "null output");
  if (!s.fault_host) { out4[0] = out4[1] = out4[2] = out4[3] = 0; return VCB_OK; }
  out4[0] = s.fault_host->code; out4[1] = s.fault_host->block; out4[2] = s.fault_host->info0; out4[3] = s.fault_host->info1;
  return VCB_OK;
}

#define VCB_GUARD(d) do { if ((d) == nullptr) return set_error(VCB_ERR_INVALID, "null descriptor"); const int rc_ = require_init(); if (rc_ != VCB_OK) return rc_; } while (0)

int vcb_conv_packed_sizes(const VcbConvDesc* d, int64_t* weight_halfs, int64_t* bias_floats) {
  if (!d) return set_error(VCB_ERR_INVALID, "null descriptor");
  return conv_packed_sizes(*d, weight_halfs, bias_floats);
}
int vcb_conv_out_hw(const VcbConvDesc* d, int32_t* ho, int32_t* wo) {
  if (!d) return set_error(VCB_ERR_INVALID, "null descriptor");
  return conv_out_hw(*d, ho, wo);
}
int vcb_conv_pack_weights(const VcbConvDesc* d, const float* w, const float* bias, void* wp, float* bp, vcb_stream_t st) {
  VCB_GUARD(d);
  if (!w || !wp || !bp) return set_error(VCB_ERR_INVALID, "conv_pack_weights: null pointer");
  return conv_pack_weights(*d, w, bias, wp, bp, (cudaStream_t)st);
}
int vcb_conv2d_fwd(const VcbConvDesc* d, const void* x, const void* wp, const float* bp, const void* res, void* y, vcb_stream_t st) {
  VCB_GUARD(d);
  return conv2d_fwd(*d, x, wp, bp, res, y, (cudaStream_t)st);
}
int vcb_conv2d_fwd_stats(const VcbConvDesc* d, const void* x, const void* wp, const float* bp, void* y, const int32_t* seg_of_image,
                         double* sums, vcb_stream_t st) {
  VCB_GUARD(d);
  if (!seg_of_image || !sums) return set_error(VCB_ERR_INVALID, "conv2d_fwd_stats: null segment table / sums");
  return conv2d_fwd(*d, x, wp, bp, nullptr, y, (cudaStream_t)st, seg_of_image, sums);
}
int vcb_frames_to_f16c4(const uint8_t* frames, void* out, int32_t n, int32_t h, int32_t w, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return frames_to_f16c4(frames, out, n, h, w, (cudaStream_t)st);
}
int vcb_frames_to_f16_s2d(const uint8_t* frames, void* out, int32_t n, int32_t h, int32_t w, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return frames_to_f16_s2d(frames, out, n, h, w, (cudaStream_t)st);
}
int vcb_frames_to_f16_s2d_wpad(const uint8_t* frames, void* out, int32_t n, int32_t h, int32_t w, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return frames_to_f16_s2d_wpad(frames, out, n, h, w, (cudaStream_t)st);
}
int vcb_letterbox_half_u8(const uint8_t* src, int32_t n, int32_t h0, int32_t w0, uint8_t* dst, int32_t h1, int32_t w1, int32_t top,
                          int32_t left, int32_t pad_value, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return letterbox_half(src, n, h0, w0, dst, h1, w1, top, left, pad_value, (cudaStream_t)st);
}
int vcb_letterbox_bilinear_u8(const uint8_t* src, int32_t n, int32_t h0, int32_t w0, uint8_t* dst, int32_t h1, int32_t w1, int32_t top,
                              int32_t left, int32_t new_h, int32_t new_w, const int32_t* xtab, const int32_t* ytab, int32_t pad_value,
                              vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return letterbox_bilinear(src, n, h0, w0, dst, h1, w1, top, left, new_h, new_w, xtab, ytab, pad_value, (cudaStream_t)st);
}
int vcb_upsample2x(const void* src, int32_t sp, void* dst, int32_t dp, int32_t n, int32_t h, int32_t w, int32_t c, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return upsample2x(src, sp, dst, dp, n, h, w, c, (cudaStream_t)st);
}
int vcb_sppf_pool(void* buf, int32_t pitch, int32_t n, int32_t h, int32_t w, int32_t c, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return sppf_pool(buf, pitch, n, h, w, c, (cudaStream_t)st);
}
int vcb_maxpool(const void* src, int32_t sp, void* dst, int32_t dp, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t s,
                int32_t p, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return maxpool(src, sp, dst, dp, n, h, w, c, k, s, p, (cudaStream_t)st);
}
int vcb_detect_decode(const VcbDetectDesc* d, float* cb, float* cs, int32_t* cc, int32_t* ci, int32_t* cnt, vcb_stream_t st) {
  VCB_GUARD(d);
  return detect_decode(*d, cb, cs, cc, ci, cnt, (cudaStream_t)st);
}
int64_t vcb_nms_workspace_bytes(int32_t n, int32_t max_candidates) { return nms_workspace_bytes(n, max_candidates); }
int vcb_nms(const VcbNmsDesc* d, const float* cb, const float* cs, const int32_t* cc, const int32_t* ci, const int32_t* cnt,
            uint64_t* ws, float* det, int32_t* det_count, vcb_stream_t st) {
  VCB_GUARD(d);
  return nms(*d, cb, cs, cc, ci, cnt, reinterpret_cast<unsigned long long*>(ws), det, det_count, (cudaStream_t)st);
}
int vcb_roi_resize_norm(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, void* out,
                        vcb_stream_t st) {
  VCB_GUARD(d);
  return roi_resize_norm(*d, frames, fh, fw, rois, out, (cudaStream_t)st);
}
int vcb_roi_stem_patches(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, void* patches,
                         vcb_stream_t st) {
  VCB_GUARD(d);
  return roi_stem_patches(*d, frames, fh, fw, rois, patches, (cudaStream_t)st);
}
int vcb_reid_stem_pool(const void* patches, const void* w_packed, const float* bias, void* out, int32_t num_rois, vcb_stream_t st) {
  return reid_stem_pool(patches, w_packed, bias, out, num_rois, (cudaStream_t)st);
}
int vcb_reid_stem_direct(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, const void* w_packed,
                         void* out, vcb_stream_t st) {
  VCB_GUARD(d);
  return reid_stem_direct(*d, frames, fh, fw, rois, w_packed, out, (cudaStream_t)st);
}
int vcb_reid_stem_direct_stats(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, const void* w_packed,
                               const int32_t* seg_of_crop, double* sums, vcb_stream_t st) {
  VCB_GUARD(d);
  return reid_stem_direct_stats(*d, frames, fh, fw, rois, w_packed, seg_of_crop, sums, (cudaStream_t)st);
}
int vcb_reid_stem_direct_bn(const VcbRoiDesc* d, const uint8_t* frames, int32_t fh, int32_t fw, const int32_t* rois, const void* w_packed,
                            const float* affine, const int32_t* seg_of_crop, void* out, vcb_stream_t st) {
  VCB_GUARD(d);
  return reid_stem_direct_bn(*d, frames, fh, fw, rois, w_packed, affine, seg_of_crop, out, (cudaStream_t)st);
}
int vcb_boxes_to_rois(const double* boxes, const int32_t* frame_of, int32_t num, int32_t fw, int32_t fh, int32_t* rois, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return boxes_to_rois(boxes, frame_of, num, fw, fh, rois, (cudaStream_t)st);
}
int vcb_avgpool_l2norm(const void* x, int32_t pitch, int32_t n, int32_t hw, int32_t c, float* out, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return avgpool_l2norm(x, pitch, n, hw, c, out, (cudaStream_t)st);
}
int vcb_bn_train_stats(const float* x, int32_t c, const int32_t* seg, int32_t num_seg, const float* gamma, const float* beta, float eps,
                       float* scale, float* shift, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return bn_train_stats(x, c, seg, num_seg, gamma, beta, eps, scale, shift, (cudaStream_t)st);
}
int vcb_bn_apply(const float* x, int32_t c, int32_t rows, const int32_t* row_seg, const float* scale, const float* shift,
                 const void* residual, int32_t res_pitch, int32_t act, void* y, int32_t y_pitch, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return bn_apply(x, c, rows, row_seg, scale, shift, residual, res_pitch, act, y, y_pitch, (cudaStream_t)st);
}

int vcb_reid_stem_stats(const void* patches, const void* w_packed, const float* bias, int32_t num_rois, const int32_t* seg_of_crop,
                        double* sums, vcb_stream_t st) {
  return reid_stem_stats(patches, w_packed, bias, num_rois, seg_of_crop, sums, (cudaStream_t)st);
}
int vcb_reid_stem_pool_bn(const void* patches, const void* w_packed, const float* affine, const int32_t* seg_of_crop, void* out,
                          int32_t num_rois, vcb_stream_t st) {
  return reid_stem_pool_bn(patches, w_packed, affine, seg_of_crop, out, num_rois, (cudaStream_t)st);
}
int vcb_bn_seg_apply_fused_f16(const void* x, int32_t c, int32_t hw, int32_t n, const int32_t* seg_of_crop, const int32_t* seg_crops,
                               const double* sums, const float* gamma, const float* beta, float eps, const void* residual, int32_t res_pitch,
                               const double* res_sums, const float* res_gamma, const float* res_beta, int32_t act, void* y, int32_t y_pitch,
                               vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return bn_seg_apply_fused_f16(x, c, hw, n, seg_of_crop, seg_crops, sums, gamma, beta, eps, residual, res_pitch, res_sums, res_gamma, res_beta,
                                act, y, y_pitch, (cudaStream_t)st);
}
int vcb_bn_seg_finalize(const double* sums, const int32_t* seg_crops, int32_t num_seg_plus1, int32_t c, int32_t hw, const float* gamma,
                        const float* beta, const float* bias, float eps, float* affine, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return bn_seg_finalize(sums, seg_crops, num_seg_plus1, c, hw, gamma, beta, bias, eps, affine, (cudaStream_t)st);
}
int vcb_bn_seg_stats_f16(const void* x, int32_t c, int32_t hw, int32_t n, const int32_t* seg_of_crop, double* sums, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return bn_seg_stats_f16(x, c, hw, n, seg_of_crop, sums, (cudaStream_t)st);
}
int vcb_bn_seg_apply_f16(const void* x, int32_t c, int32_t h, int32_t w, int32_t n, const int32_t* seg_of_crop, const float* affine,
                         const void* residual, int32_t res_pitch, int32_t act, int32_t pool, void* y, int32_t y_pitch, vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  return bn_seg_apply_f16(x, c, h, w, n, seg_of_crop, affine, residual, res_pitch, act, pool, y, y_pitch, (cudaStream_t)st);
}

// ---- CUDA graph capture ------------------------------------------------------------------------
int vcb_graph_begin(vcb_stream_t st) {
  const int rc = require_init(); if (rc) return rc;
  if (st == nullptr) return set_error(VCB_ERR_INVALID, "graph capture needs a non-default stream");
  return check_cuda(cudaStreamBeginCapture((cudaStream_t)st, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
}
int vcb_graph_end(vcb_stream_t st, VcbGraph** out) {
  if (!out) return set_error(VCB_ERR_INVALID, "null output");
  VcbGraph* g = new VcbGraph();
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)st, &g->graph);
  if (e != cudaSuccess || g->graph == nullptr) { delete g; return check_cuda(e != cudaSuccess ? e : cudaErrorUnknown, "cudaStreamEndCapture"); }
  size_t n = 0;
  cudaGraphGetNodes(g->graph, nullptr, &n);
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) cudaGraphGetNodes(g->graph, nodes.data(), &n);
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++g->num_kernels;
  }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(g->graph); delete g; return check_cuda(e, "cudaGraphInstantiate"); }
  *out = g;
  return VCB_OK;
}
int vcb_graph_launch(VcbGraph* g, vcb_stream_t st) {
  if (!g || !g->exec) return set_error(VCB_ERR_INVALID, "null graph");
  return check_cuda(cudaGraphLaunch(g->exec, (cudaStream_t)st), "cudaGraphLaunch");
}
int vcb_graph_num_kernels(const VcbGraph* g) { return g ? g->num_kernels : 0; }
int vcb_graph_destroy(VcbGraph* g) {
  if (!g) return VCB_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return VCB_OK;
}

}  // extern "C"
