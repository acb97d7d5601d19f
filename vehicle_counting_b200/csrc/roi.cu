// K5 -- ROI crop + bilinear resize + normalise into one dense NHWC4 fp16 batch for the ReID CNN.
//
// Replaces the reference's per-crop host loop:
//   crop rule       /root/reference/networks/deepsort/deep_sort.py:78-95, :119-129
//                   (centre/size in float64, int() truncation, clip to [0,W-1]x[0,H-1], exclusive end)
//   preprocessing   /root/reference/networks/deepsort/deep/feature_extractor.py:26-39
//                   (im/255 -> cv2.resize INTER_LINEAR to 50x50 -> (x-mean)/std, RGB statistics applied
//                    to the stored B,G,R order)
// cv2.resize semantics restated: source coordinate (d+0.5)*scale-0.5 in double, floor, fraction as
// float, index clamped to the edge with fraction 0 (oracle/reid.py resize_bilinear is the CPU twin).
// HBM-bound: reads at most 4 taps x 3 B per output pixel, writes 8 B per output pixel.
#include "vcb_internal.h"

namespace vcb {

template <int OC>   // output channel pitch: 4, 8 or 16 halves per pixel (channels >= 3 are zero)
__global__ void roi_resize_norm_kernel(const VcbRoiDesc d, const uint8_t* __restrict__ frames, int fh, int fw,
                                       const int* __restrict__ rois, uint2* __restrict__ out) {
  constexpr int U = OC / 4;   // uint2 words per pixel
  const int r = blockIdx.x;
  const int f = rois[r * 5 + 0], x1 = rois[r * 5 + 1], y1 = rois[r * 5 + 2], x2 = rois[r * 5 + 3], y2 = rois[r * 5 + 4];
  const int cw = x2 - x1, chh = y2 - y1;
  const int S = d.out_size;
  uint2* o = out + (long long)r * S * S * U;
  if (cw <= 0 || chh <= 0 || x1 < 0 || y1 < 0 || x2 > fw || y2 > fh || f < 0 || (d.num_frames > 0 && f >= d.num_frames)) {   // the reference would raise inside cv2.resize
    for (int i = threadIdx.x; i < S * S * U; i += blockDim.x) o[i] = make_uint2(0u, 0u);
    return;
  }
  const uint8_t* img = frames + (long long)f * fh * fw * 3;
  const double sx = (double)cw / (double)S, sy = (double)chh / (double)S;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const int oy = i / S, ox = i - oy * S;
    const double fy_d = ((double)oy + 0.5) * sy - 0.5, fx_d = ((double)ox + 0.5) * sx - 0.5;
    int iy = (int)floor(fy_d), ix = (int)floor(fx_d);
    float fy = (float)(fy_d - (double)iy), fx = (float)(fx_d - (double)ix);
    if (iy < 0) { iy = 0; fy = 0.f; }
    if (ix < 0) { ix = 0; fx = 0.f; }
    int iy1 = iy + 1, ix1 = ix + 1;
    if (iy >= chh - 1) { iy = chh - 1; iy1 = chh - 1; fy = 0.f; }
    if (ix >= cw - 1) { ix = cw - 1; ix1 = cw - 1; fx = 0.f; }
    const uint8_t* p00 = img + ((long long)(y1 + iy) * fw + (x1 + ix)) * 3;
    const uint8_t* p01 = img + ((long long)(y1 + iy) * fw + (x1 + ix1)) * 3;
    const uint8_t* p10 = img + ((long long)(y1 + iy1) * fw + (x1 + ix)) * 3;
    const uint8_t* p11 = img + ((long long)(y1 + iy1) * fw + (x1 + ix1)) * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = (float)p00[c] / 255.0f, b = (float)p01[c] / 255.0f;
      const float e = (float)p10[c] / 255.0f, g = (float)p11[c] / 255.0f;
      const float top = a * (1.0f - fx) + b * fx;
      const float bot = e * (1.0f - fx) + g * fx;
      const float val = top * (1.0f - fy) + bot * fy;
      v[c] = (val - d.mean[c]) * d.inv_std[c];
    }
    const __half2 lo = __floats2half2_rn(v[0], v[1]);
    const __half2 hi = __floats2half2_rn(v[2], 0.0f);
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t*>(&lo);
    w.y = *reinterpret_cast<const uint32_t*>(&hi);
    o[(long long)i * U] = w;
#pragma unroll
    for (int u = 1; u < U; ++u) o[(long long)i * U + u] = make_uint2(0u, 0u);
  }
}

int roi_resize_norm(const VcbRoiDesc& d, const uint8_t* frames, int fh, int fw, const int* rois, void* out, cudaStream_t st) {
  if (d.num_rois < 0 || d.out_size <= 0 || !frames || !rois || !out || fh <= 0 || fw <= 0 || ((uintptr_t)out & 7))
    return set_error(VCB_ERR_INVALID, "roi_resize_norm: bad argument");
  if (d.num_rois == 0) return VCB_OK;
  const int oc = d.out_channels == 0 ? 4 : d.out_channels;
  if (oc == 4) roi_resize_norm_kernel<4><<<d.num_rois, 256, 0, st>>>(d, frames, fh, fw, rois, reinterpret_cast<uint2*>(out));
  else if (oc == 8) roi_resize_norm_kernel<8><<<d.num_rois, 256, 0, st>>>(d, frames, fh, fw, rois, reinterpret_cast<uint2*>(out));
  else if (oc == 16) roi_resize_norm_kernel<16><<<d.num_rois, 256, 0, st>>>(d, frames, fh, fw, rois, reinterpret_cast<uint2*>(out));
  else return set_error(VCB_ERR_INVALID, "roi_resize_norm: out_channels must be 4, 8 or 16");
  return check_cuda(cudaGetLastError(), "roi_resize_norm launch");
}

// deep_sort.py:78-95 in float64: w=x2-x1; cx=x1+w/2; x1i=max(int(cx-w/2),0); x2i=min(int(cx+w/2),W-1) ...
__global__ void boxes_to_rois_kernel(const double* __restrict__ boxes, const int* __restrict__ frame_of, int num, int fw, int fh,
                                     int* __restrict__ rois) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num) return;
  const double x1 = boxes[i * 4 + 0], y1 = boxes[i * 4 + 1], x2 = boxes[i * 4 + 2], y2 = boxes[i * 4 + 3];
  const double w = __dsub_rn(x2, x1), h = __dsub_rn(y2, y1);
  const double cx = __dadd_rn(x1, w / 2.0), cy = __dadd_rn(y1, h / 2.0);
  int ix1 = (int)__dsub_rn(cx, w / 2.0), ix2 = (int)__dadd_rn(cx, w / 2.0);
  int iy1 = (int)__dsub_rn(cy, h / 2.0), iy2 = (int)__dadd_rn(cy, h / 2.0);
  ix1 = ix1 > 0 ? ix1 : 0;
  iy1 = iy1 > 0 ? iy1 : 0;
  ix2 = ix2 < fw - 1 ? ix2 : fw - 1;
  iy2 = iy2 < fh - 1 ? iy2 : fh - 1;
  rois[i * 5 + 0] = frame_of ? frame_of[i] : 0;
  rois[i * 5 + 1] = ix1;
  rois[i * 5 + 2] = iy1;
  rois[i * 5 + 3] = ix2;
  rois[i * 5 + 4] = iy2;
}

int boxes_to_rois(const double* boxes, const int* frame_of, int num, int fw, int fh, int* rois, cudaStream_t st) {
  if (num < 0 || !boxes || !rois || fw <= 0 || fh <= 0) return set_error(VCB_ERR_INVALID, "boxes_to_rois: bad argument");
  if (num == 0) return VCB_OK;
  boxes_to_rois_kernel<<<(num + 127) / 128, 128, 0, st>>>(boxes, frame_of, num, fw, fh, rois);
  return check_cuda(cudaGetLastError(), "boxes_to_rois launch");
}

}  // namespace vcb
