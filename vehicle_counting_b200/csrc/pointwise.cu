// K2 / K6 / K7 -- HBM-bound data-movement kernels around the convolutions: frame ingest, nearest
// upsample into a concat slice, the SPPF max-pool cascade, generic max-pool, the ReID average-pool +
// L2-norm tail and the train-mode BatchNorm statistics / apply pair.  All are 16-byte vectorised over
// the NHWC channel dimension; none reuses data enough to want shared-memory staging except SPPF.
#include "vcb_internal.h"

namespace vcb {

static inline int grid_for(long long work, int block, int max_blocks = 148 * 16) {
  long long b = (work + block - 1) / block;
  if (b < 1) b = 1;
  return (int)(b < max_blocks ? b : max_blocks);
}

// ---------------------------------------------------------------- frames (uint8 HWC3) -> fp16 NHWC4
// [upstream AutoShape.forward: x = torch.from_numpy(x).to(device).type_as(p) / 255]
__global__ void frames_to_f16c4_kernel(const uint8_t* __restrict__ in, uint2* __restrict__ out, long long pixels) {
  const long long quads = pixels >> 2;   // 4 pixels = 12 input bytes = 3 aligned words
  const float k = 1.0f / 255.0f;
  (void)k;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + q * 3;
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    const uint8_t b[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                           (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                           (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const __half2 lo = __floats2half2_rn((float)b[px * 3] / 255.0f, (float)b[px * 3 + 1] / 255.0f);
      const __half2 hi = __floats2half2_rn((float)b[px * 3 + 2] / 255.0f, 0.0f);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      out[q * 4 + px] = o;
    }
  }
  // tail pixels (pixels % 4)
  const long long tail0 = quads << 2;
  for (long long px = tail0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; px < pixels; px += (long long)gridDim.x * blockDim.x) {
    const uint8_t* s = in + px * 3;
    const __half2 lo = __floats2half2_rn((float)s[0] / 255.0f, (float)s[1] / 255.0f);
    const __half2 hi = __floats2half2_rn((float)s[2] / 255.0f, 0.0f);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo);
    o.y = *reinterpret_cast<const uint32_t*>(&hi);
    out[px] = o;
  }
}

int frames_to_f16c4(const uint8_t* frames, void* out, int n, int h, int w, cudaStream_t st) {
  if (!frames || !out || n <= 0 || h <= 0 || w <= 0) return set_error(VCB_ERR_INVALID, "frames_to_f16c4: bad argument");
  if (((uintptr_t)frames & 3) || ((uintptr_t)out & 7)) return set_error(VCB_ERR_INVALID, "frames_to_f16c4: misaligned pointer");
  const long long pixels = (long long)n * h * w;
  frames_to_f16c4_kernel<<<grid_for(pixels / 4 + 1, 256), 256, 0, st>>>(frames, reinterpret_cast<uint2*>(out), pixels);
  return check_cuda(cudaGetLastError(), "frames_to_f16c4 launch");
}

// ---------------------------------------------------------------- frames (uint8 HWC3) -> fp16 space-to-depth NHWC16
// out[n][y/2][x/2][(dy*2+dx)*3 + c] = in[n][y][x][c] / 255, channels 12..15 zero.  With this layout the 6x6/s2/p2 stem
// of YOLOv5 v6.0 is exactly a 3x3/s1/p1 convolution over 12 (+4 zero) channels: w'[a][b][(dy,dx,c)] = w[2a+dy][2b+dx][c].
// `opitch` = pixels per output row, `xoff` = column of the first pixel: (w/2, 0) for the dense layout, (w/2 + 2, 1) for the
// W-padded layout of the row-window stem (pad columns are never written: the caller zeroes the buffer once).
__global__ void frames_to_f16_s2d_kernel(const uint8_t* __restrict__ in, uint4* __restrict__ out, int n, int h, int w, int opitch, int xoff) {
  const int h2 = h >> 1, w2 = w >> 1;
  const long long total = (long long)n * h2 * w2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x2 = (int)(i % w2);
    long long t = i / w2;
    const int y2 = (int)(t % h2);
    const int b = (int)(t / h2);
    const uint8_t* r0 = in + (((long long)b * h + 2 * y2) * w + 2 * x2) * 3;   // two pixels = 6 contiguous bytes per row
    const uint8_t* r1 = r0 + (long long)w * 3;
    float v[16];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      v[k] = (float)__ldg(r0 + k) / 255.0f;       // (dy=0, dx=0..1, c)
      v[6 + k] = (float)__ldg(r1 + k) / 255.0f;   // (dy=1, dx=0..1, c)
    }
    v[12] = v[13] = v[14] = v[15] = 0.0f;
    __half2 hv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) hv[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    const long long o = ((long long)b * h2 + y2) * opitch + xoff + x2;
    out[o * 2] = *reinterpret_cast<const uint4*>(hv);
    out[o * 2 + 1] = *reinterpret_cast<const uint4*>(hv + 4);
  }
}

// w % 8 == 0: one thread converts FOUR output pixels = 24 contiguous bytes of two frame rows (three 8-byte loads each)
// into 128 contiguous output bytes -- 8x fewer, wider memory instructions than the per-pixel kernel above.
__global__ void frames_to_f16_s2d_x4_kernel(const uint8_t* __restrict__ in, uint4* __restrict__ out, int n, int h, int w, int opitch, int xoff) {
  const int h2 = h >> 1, w8 = w >> 3;
  const long long total = (long long)n * h2 * w8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(i % w8);
    long long t = i / w8;
    const int y2 = (int)(t % h2);
    const int b = (int)(t / h2);
    const uint8_t* r0 = in + (((long long)b * h + 2 * y2) * w + 8 * xq) * 3;
    const uint2* p0 = reinterpret_cast<const uint2*>(r0);
    const uint2* p1 = reinterpret_cast<const uint2*>(r0 + (long long)w * 3);
    uint32_t wd[2][6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint2 a = __ldg(p0 + k), c = __ldg(p1 + k);
      wd[0][2 * k] = a.x; wd[0][2 * k + 1] = a.y;
      wd[1][2 * k] = c.x; wd[1][2 * k + 1] = c.y;
    }
    const long long o = ((long long)b * h2 + y2) * opitch + xoff + 4 * xq;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float v[16];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int j = px * 6 + k;                  // byte j of the 24-byte row segment (constant after unrolling)
        v[k] = (float)((wd[0][j >> 2] >> (8 * (j & 3))) & 0xffu) / 255.0f;
        v[6 + k] = (float)((wd[1][j >> 2] >> (8 * (j & 3))) & 0xffu) / 255.0f;
      }
      v[12] = v[13] = v[14] = v[15] = 0.0f;
      __half2 hv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hv[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
      out[(o + px) * 2] = *reinterpret_cast<const uint4*>(hv);
      out[(o + px) * 2 + 1] = *reinterpret_cast<const uint4*>(hv + 4);
    }
  }
}

static int frames_to_f16_s2d_impl(const uint8_t* frames, void* out, int n, int h, int w, int wpad, cudaStream_t st) {
  if (!frames || !out || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || ((uintptr_t)out & 15))
    return set_error(VCB_ERR_INVALID, "frames_to_f16_s2d: bad argument (h, w must be even, out 16-byte aligned)");
  const int opitch = w / 2 + (wpad ? 2 : 0), xoff = wpad ? 1 : 0;
  if ((w & 7) == 0 && ((uintptr_t)frames & 7) == 0) {
    const long long total = (long long)n * (h / 2) * (w / 8);
    frames_to_f16_s2d_x4_kernel<<<grid_for(total, 256), 256, 0, st>>>(frames, reinterpret_cast<uint4*>(out), n, h, w, opitch, xoff);
  } else {
    const long long total = (long long)n * (h / 2) * (w / 2);
    frames_to_f16_s2d_kernel<<<grid_for(total, 256), 256, 0, st>>>(frames, reinterpret_cast<uint4*>(out), n, h, w, opitch, xoff);
  }
  return check_cuda(cudaGetLastError(), "frames_to_f16_s2d launch");
}
int frames_to_f16_s2d(const uint8_t* frames, void* out, int n, int h, int w, cudaStream_t st) {
  return frames_to_f16_s2d_impl(frames, out, n, h, w, 0, st);
}
int frames_to_f16_s2d_wpad(const uint8_t* frames, void* out, int n, int h, int w, cudaStream_t st) {
  return frames_to_f16_s2d_impl(frames, out, n, h, w, 1, st);
}

// ---------------------------------------------------------------- letterbox, exact 2x reduction
// upstream AutoShape letterboxes on the host: cv2.resize(INTER_LINEAR) + copyMakeBorder(114) (reached from
// /root/reference/networks/yolo.py:70).  For an exact 2x reduction of uint8 data cv2's fixed-point bilinear equals the 2x2
// box mean with round-half-up, (a+b+c+d+2)>>2 (checked bit for bit in tests/test_kernels_gpu.py), so a
// 1280x720 frame becomes its 640x360 image inside the 384x640 inference frame on the device; one thread per output pixel.
__global__ void letterbox_half_kernel(const uint8_t* __restrict__ src, int n, int h0, int w0, uint8_t* __restrict__ dst, int h1, int w1,
                                      int top, int left, int pad) {
  const int hh = h0 >> 1, wh = w0 >> 1;
  const long long total = (long long)n * h1 * w1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w1);
    long long t = i / w1;
    const int y = (int)(t % h1);
    const int b = (int)(t / h1);
    const int sy = y - top, sx = x - left;
    uint8_t* o = dst + i * 3;
    if ((unsigned)sy < (unsigned)hh && (unsigned)sx < (unsigned)wh) {
      const uint8_t* p0 = src + (((long long)b * h0 + 2 * sy) * w0 + 2 * sx) * 3;
      const uint8_t* p1 = p0 + (long long)w0 * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = (uint8_t)(((int)p0[c] + (int)p0[3 + c] + (int)p1[c] + (int)p1[3 + c] + 2) >> 2);
    } else {
      o[0] = o[1] = o[2] = (uint8_t)pad;
    }
  }
}

int letterbox_half(const uint8_t* src, int n, int h0, int w0, uint8_t* dst, int h1, int w1, int top, int left, int pad, cudaStream_t st) {
  if (!src || !dst || n <= 0 || h0 <= 0 || w0 <= 0 || (h0 & 1) || (w0 & 1) || h1 <= 0 || w1 <= 0 || top < 0 || left < 0 ||
      top + h0 / 2 > h1 || left + w0 / 2 > w1 || pad < 0 || pad > 255)
    return set_error(VCB_ERR_INVALID, "letterbox_half: bad argument (even source size; the halved image must fit at (top, left))");
  const long long total = (long long)n * h1 * w1;
  letterbox_half_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, n, h0, w0, dst, h1, w1, top, left, pad);
  return check_cuda(cudaGetLastError(), "letterbox_half launch");
}

// ---------------------------------------------------------------- letterbox, any ratio: cv2.resize(INTER_LINEAR) on uint8, bit for bit
// OpenCV resizes 8-bit images in fixed point (resize.cpp: INTER_RESIZE_COEF_BITS = 11): per destination column / row two source
// indices and two 11-bit weights (computed on the host exactly as OpenCV does, in float32: networks/yolo.py cv2_linear_table),
// a horizontal pass in int32, and the vertical pass ((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2.  xtab / ytab:
// int32 [new_w | new_h][4] = index0, index1, weight0, weight1.  One thread per output pixel; pixels outside the resized image
// get the pad value (copyMakeBorder(114)).  Replaces the host cv2 pass of upstream's letterbox for same-size frame batches.
__global__ void letterbox_bilinear_kernel(const uint8_t* __restrict__ src, int n, int h0, int w0, uint8_t* __restrict__ dst, int h1, int w1,
                                          int top, int left, int nh, int nw, const int4* __restrict__ xtab, const int4* __restrict__ ytab,
                                          int pad) {
  const long long total = (long long)n * h1 * w1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w1);
    long long t = i / w1;
    const int y = (int)(t % h1);
    const int b = (int)(t / h1);
    const int sy = y - top, sx = x - left;
    uint8_t* o = dst + i * 3;
    if ((unsigned)sy < (unsigned)nh && (unsigned)sx < (unsigned)nw) {
      const int4 xt = __ldg(xtab + sx), yt = __ldg(ytab + sy);
      const uint8_t* img = src + (long long)b * h0 * w0 * 3;
      const uint8_t* r0 = img + (long long)yt.x * w0 * 3;
      const uint8_t* r1 = img + (long long)yt.y * w0 * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int a0 = (int)r0[xt.x * 3 + c] * xt.z + (int)r0[xt.y * 3 + c] * xt.w;
        const int a1 = (int)r1[xt.x * 3 + c] * xt.z + (int)r1[xt.y * 3 + c] * xt.w;
        int v = (((yt.z * (a0 >> 4)) >> 16) + ((yt.w * (a1 >> 4)) >> 16) + 2) >> 2;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        o[c] = (uint8_t)v;
      }
    } else {
      o[0] = o[1] = o[2] = (uint8_t)pad;
    }
  }
}

int letterbox_bilinear(const uint8_t* src, int n, int h0, int w0, uint8_t* dst, int h1, int w1, int top, int left, int nh, int nw,
                       const int* xtab, const int* ytab, int pad, cudaStream_t st) {
  if (!src || !dst || !xtab || !ytab || n <= 0 || h0 <= 0 || w0 <= 0 || h1 <= 0 || w1 <= 0 || nh <= 0 || nw <= 0 || top < 0 || left < 0 ||
      top + nh > h1 || left + nw > w1 || pad < 0 || pad > 255 || ((uintptr_t)xtab & 15) || ((uintptr_t)ytab & 15))
    return set_error(VCB_ERR_INVALID, "letterbox_bilinear: bad argument (the resized image must fit at (top, left); tables 16-byte aligned)");
  const long long total = (long long)n * h1 * w1;
  letterbox_bilinear_kernel<<<grid_for(total, 256), 256, 0, st>>>(src, n, h0, w0, dst, h1, w1, top, left, nh, nw,
                                                                   reinterpret_cast<const int4*>(xtab), reinterpret_cast<const int4*>(ytab), pad);
  return check_cuda(cudaGetLastError(), "letterbox_bilinear launch");
}

// ---------------------------------------------------------------- nearest x2 upsample into a channel slice
__global__ void upsample2x_kernel(const uint4* __restrict__ src, int src_pitch8, uint4* __restrict__ dst, int dst_pitch8, int n,
                                  int h, int w, int c8) {
  const long long total = (long long)n * (2 * h) * (2 * w) * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8);
    long long t = i / c8;
    const int ox = (int)(t % (2 * w));
    t /= (2 * w);
    const int oy = (int)(t % (2 * h));
    const int b = (int)(t / (2 * h));
    const uint4 v = __ldg(src + ((long long)(b * h + (oy >> 1)) * w + (ox >> 1)) * src_pitch8 + cc);
    dst[((long long)(b * 2 * h + oy) * (2 * w) + ox) * dst_pitch8 + cc] = v;
  }
}

int upsample2x(const void* src, int src_pitch, void* dst, int dst_pitch, int n, int h, int w, int c, cudaStream_t st) {
  if (!src || !dst || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 7) || (src_pitch & 7) || (dst_pitch & 7) ||
      ((uintptr_t)src & 15) || ((uintptr_t)dst & 15))
    return set_error(VCB_ERR_INVALID, "upsample2x: bad argument (channels/pitches must be multiples of 8, pointers 16B aligned)");
  const long long total = (long long)n * 4 * h * w * (c / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), src_pitch / 8,
                                                           reinterpret_cast<uint4*>(dst), dst_pitch / 8, n, h, w, c / 8);
  return check_cuda(cudaGetLastError(), "upsample2x launch");
}

// ---------------------------------------------------------------- SPPF cascade: y1 = mp5(x), y2 = mp5(y1), y3 = mp5(y2)
__device__ __forceinline__ uint4 hmax8(const uint4& a, const uint4& b) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// one CTA per (frame, 8-channel group): the whole h*w map of that group lives in shared memory
__global__ void sppf_pool_kernel(uint4* __restrict__ buf, int pitch8, int h, int w, int c8) {
  extern __shared__ uint4 sm[];
  uint4* A = sm;
  uint4* B = sm + h * w;
  const int b = blockIdx.x / c8, cc = blockIdx.x % c8;
  uint4* base = buf + (long long)b * h * w * pitch8 + cc;
  const int hw = h * w;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) A[i] = base[(long long)i * pitch8];
  __syncthreads();
  for (int round = 1; round <= 3; ++round) {
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {       // horizontal 5-max A -> B
      const int y = i / w, x = i - y * w;
      uint4 m = A[i];
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (dx != 0 && xx >= 0 && xx < w) m = hmax8(m, A[y * w + xx]);
      }
      B[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < hw; i += blockDim.x) {       // vertical 5-max B -> A, and out
      const int y = i / w, x = i - y * w;
      uint4 m = B[i];
      for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (dy != 0 && yy >= 0 && yy < h) m = hmax8(m, B[yy * w + x]);
      }
      A[i] = m;
      base[(long long)i * pitch8 + (long long)round * c8] = m;
    }
    __syncthreads();
  }
}

int sppf_pool(void* buf, int pitch, int n, int h, int w, int c, cudaStream_t st) {
  if (!buf || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 7) || (pitch & 7) || pitch < 4 * c || ((uintptr_t)buf & 15))
    return set_error(VCB_ERR_INVALID, "sppf_pool: bad argument");
  const size_t smem = (size_t)h * w * 16 * 2;
  if (smem > 200 * 1024) return set_error(VCB_ERR_INVALID, "sppf_pool: feature map too large for the shared-memory cascade");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    const cudaError_t e = cudaFuncSetAttribute(sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(sppf)");
    configured = 200 * 1024;
  }
  sppf_pool_kernel<<<n * (c / 8), 256, smem, st>>>(reinterpret_cast<uint4*>(buf), pitch / 8, h, w, c / 8);
  return check_cuda(cudaGetLastError(), "sppf_pool launch");
}

// ---------------------------------------------------------------- generic MaxPool2d(k, s, p), NHWC fp16
__global__ void maxpool_kernel(const uint4* __restrict__ src, int src_pitch8, uint4* __restrict__ dst, int dst_pitch8, int n, int h,
                               int w, int c8, int k, int s, int p, int ho, int wo) {
  const long long total = (long long)n * ho * wo * c8;
  const __half2 ninf = __float2half2_rn(-65504.0f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8);
    long long t = i / c8;
    const int ox = (int)(t % wo);
    t /= wo;
    const int oy = (int)(t % ho);
    const int b = (int)(t / ho);
    uint4 m;
    __half2* pm = reinterpret_cast<__half2*>(&m);
    pm[0] = pm[1] = pm[2] = pm[3] = ninf;
    for (int dy = 0; dy < k; ++dy) {
      const int y = oy * s - p + dy;
      if (y < 0 || y >= h) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int x = ox * s - p + dx;
        if (x < 0 || x >= w) continue;
        m = hmax8(m, __ldg(src + ((long long)(b * h + y) * w + x) * src_pitch8 + cc));
      }
    }
    dst[((long long)(b * ho + oy) * wo + ox) * dst_pitch8 + cc] = m;
  }
}

int maxpool(const void* src, int src_pitch, void* dst, int dst_pitch, int n, int h, int w, int c, int k, int s, int p,
            cudaStream_t st) {
  if (!src || !dst || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 7) || (src_pitch & 7) || (dst_pitch & 7) || k <= 0 || s <= 0 ||
      p < 0 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15))
    return set_error(VCB_ERR_INVALID, "maxpool: bad argument");
  const int ho = (h + 2 * p - k) / s + 1, wo = (w + 2 * p - k) / s + 1;
  if (ho <= 0 || wo <= 0) return set_error(VCB_ERR_INVALID, "maxpool: empty output");
  const long long total = (long long)n * ho * wo * (c / 8);
  maxpool_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), src_pitch / 8,
                                                        reinterpret_cast<uint4*>(dst), dst_pitch / 8, n, h, w, c / 8, k, s, p, ho, wo);
  return check_cuda(cudaGetLastError(), "maxpool launch");
}

// ---------------------------------------------------------------- AvgPool(hw) + L2 normalise (model.py:71, :93-95)
__global__ void avgpool_l2norm_kernel(const __half* __restrict__ x, int pitch, int hw, int c, float* __restrict__ out) {
  extern __shared__ float vals[];           // c floats + 32 for the reduction
  float* red = vals + c;
  const __half* xb = x + (long long)blockIdx.x * hw * pitch;
  float ss = 0.0f;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.0f;
    for (int i = 0; i < hw; ++i) s += __half2float(xb[(long long)i * pitch + ch]);
    s /= (float)hw;
    vals[ch] = s;
    ss += s * s;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[0] = sqrtf(t);
  }
  __syncthreads();
  const float nrm = red[0];
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) out[(long long)blockIdx.x * c + ch] = vals[ch] / nrm;
}

int avgpool_l2norm(const void* x, int pitch, int n, int hw, int c, float* out, cudaStream_t st) {
  if (!x || !out || n <= 0 || hw <= 0 || c <= 0 || pitch < c) return set_error(VCB_ERR_INVALID, "avgpool_l2norm: bad argument");
  avgpool_l2norm_kernel<<<n, 256, (size_t)(c + 32) * sizeof(float), st>>>(reinterpret_cast<const __half*>(x), pitch, hw, c, out);
  return check_cuda(cudaGetLastError(), "avgpool_l2norm launch");
}

// ---------------------------------------------------------------- train-mode BatchNorm (reference Extractor never calls .eval())
// grid = (num_seg, ceil(c/32)); block = 32 channels x 8 row lanes; double accumulation.
__global__ void bn_train_stats_kernel(const float* __restrict__ x, int c, const int* __restrict__ seg_row_start,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  __shared__ double s1[8][33], s2[8][33];
  const int seg = blockIdx.x;
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const int r0 = seg_row_start[seg], r1 = seg_row_start[seg + 1];
  double a = 0.0, b = 0.0;
  if (ch < c) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const double v = (double)x[(long long)r * c + ch];
      a += v;
      b += v * v;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    const double cnt = (double)(r1 - r0);
    const double mean = cnt > 0 ? a / cnt : 0.0;
    double var = cnt > 0 ? b / cnt - mean * mean : 0.0;
    if (var < 0) var = 0;
    const float sc = gamma[ch] * (float)(1.0 / sqrt(var + (double)eps));
    scale[(long long)seg * c + ch] = sc;
    shift[(long long)seg * c + ch] = beta[ch] - (float)mean * sc;
  }
}

int bn_train_stats(const float* x, int c, const int* seg_row_start, int num_seg, const float* gamma, const float* beta, float eps,
                   float* scale, float* shift, cudaStream_t st) {
  if (!x || !seg_row_start || !gamma || !beta || !scale || !shift || c <= 0 || num_seg <= 0)
    return set_error(VCB_ERR_INVALID, "bn_train_stats: bad argument");
  dim3 grid(num_seg, (c + 31) / 32), block(32, 8);
  bn_train_stats_kernel<<<grid, block, 0, st>>>(x, c, seg_row_start, gamma, beta, eps, scale, shift);
  return check_cuda(cudaGetLastError(), "bn_train_stats launch");
}

__global__ void bn_apply_kernel(const float* __restrict__ x, int c, long long rows, const int* __restrict__ row_seg,
                                const float* __restrict__ scale, const float* __restrict__ shift, const __half* __restrict__ residual,
                                int res_pitch, int act, __half* __restrict__ y, int y_pitch) {
  const int c4 = c >> 2;
  const long long total = rows * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int ch = (int)(i - r * c4) * 4;
    const int seg = row_seg[r];
    const float4 v = *reinterpret_cast<const float4*>(x + r * c + ch);
    const float4 sc = *reinterpret_cast<const float4*>(scale + (long long)seg * c + ch);
    const float4 sh = *reinterpret_cast<const float4*>(shift + (long long)seg * c + ch);
    float f[4] = {v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w};
    if (residual != nullptr) {
      const __half2* rp = reinterpret_cast<const __half2*>(residual + r * res_pitch + ch);
      const float2 r0 = __half22float2(rp[0]), r1 = __half22float2(rp[1]);
      f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
    }
    if (act == VCB_ACT_RELU) {
      for (int k = 0; k < 4; ++k) f[k] = fmaxf(f[k], 0.0f);
    } else if (act == VCB_ACT_SILU) {
      for (int k = 0; k < 4; ++k) f[k] = f[k] / (1.0f + __expf(-f[k]));
    }
    __half2* yp = reinterpret_cast<__half2*>(y + r * y_pitch + ch);
    yp[0] = __floats2half2_rn(f[0], f[1]);
    yp[1] = __floats2half2_rn(f[2], f[3]);
  }
}

int bn_apply(const float* x, int c, int rows, const int* row_seg, const float* scale, const float* shift, const void* residual,
             int res_pitch, int act, void* y, int y_pitch, cudaStream_t st) {
  if (!x || !row_seg || !scale || !shift || !y || c <= 0 || (c & 3) || rows <= 0 || (y_pitch & 3) || (residual && (res_pitch & 3)))
    return set_error(VCB_ERR_INVALID, "bn_apply: bad argument");
  const long long total = (long long)rows * (c / 4);
  bn_apply_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, c, rows, row_seg, scale, shift, reinterpret_cast<const __half*>(residual),
                                                         res_pitch, act, reinterpret_cast<__half*>(y), y_pitch);
  return check_cuda(cudaGetLastError(), "bn_apply launch");
}

// ---------------------------------------------------------------- train-mode BatchNorm, graph-friendly fast path (round 2)
// The reference's Extractor never calls .eval() (feature_extractor.py:10-22, :42-47): every BatchNorm2d of model.py normalises
// with the statistics of ONE call = the crops of one (frame, class).  A batch of many calls is a list of SEGMENTS of crops
// (seg_of_crop[crop] in [0, num_seg]; the value num_seg marks padding crops of a bucketed batch, kept out of every statistic).
// Pre-BN convolution outputs are stored as fp16 (x), so a layer costs conv -> bn_seg_stats -> bn_seg_apply with 2 + 2 + 2(+2)
// bytes per element instead of the 4 + 4 + 4 + 2 of the fp32 path above; both kernels take device-resident segment tables, so
// the whole pass is a replayable CUDA graph.  sums: double [num_seg + 1][c][2] (sum, sum of squares), zeroed by the caller.
// grid = crops; block = 256 threads = (c / 8 channel vectors) x row lanes.
__global__ void __launch_bounds__(256) bn_seg_stats_f16_kernel(const uint4* __restrict__ x, int c8, int hw, int n, int cpb,
                                                                const int* __restrict__ seg_of_crop, double* __restrict__ sums) {
  // A block walks `cpb` consecutive crops (>= ~128 KiB of data) and keeps its partial sums in registers while the segment stays
  // the same: one flush (shared-memory reduction + c * 2 fp64 atomics) per segment per block instead of per crop -- the deep
  // layers (16-49 pixels per crop, 256-512 channels) were bound by their 4 M atomics per launch (profiles/r02_train_bn.md).
  __shared__ float red[256 * 16];
  const int lanes = 256 / c8;                       // row lanes (c8 is a power of two <= 64)
  const int cv = threadIdx.x % c8, rl = threadIdx.x / c8;
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = 0.f; b[k] = 0.f; }
  const int crop0 = blockIdx.x * cpb, crop1 = min(n, crop0 + cpb);
  int cur = crop0 < n ? seg_of_crop[crop0] : -1;
  auto flush = [&](int seg) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[threadIdx.x * 16 + k] = a[k]; red[threadIdx.x * 16 + 8 + k] = b[k]; a[k] = 0.f; b[k] = 0.f; }
    __syncthreads();
    for (int t = threadIdx.x; t < c8 * 16; t += 256) {      // value t of the block: (channel vector, slot) = (t / 16, t % 16)
      const int v = t >> 4, slot = t & 15;
      double acc = 0.0;
      for (int l = 0; l < lanes; ++l) acc += (double)red[(l * c8 + v) * 16 + slot];
      atomicAdd(sums + ((long long)seg * (c8 * 8) + v * 8 + (slot & 7)) * 2 + (slot >> 3), acc);
    }
  };
  for (int crop = crop0; crop < crop1; ++crop) {
    const int seg = seg_of_crop[crop];
    if (seg != cur) { flush(cur); cur = seg; }
    const uint4* base = x + (long long)crop * hw * c8;
    for (int r0 = rl; r0 < hw; r0 += 4 * lanes) {           // four independent 16-byte loads in flight per thread (eight: 0.79 -> 1.19 ms)
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * lanes;
        v[u] = r < hw ? __ldg(base + (long long)r * c8 + cv) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h[k]);
          a[2 * k] += f.x; a[2 * k + 1] += f.y;
          b[2 * k] = fmaf(f.x, f.x, b[2 * k]); b[2 * k + 1] = fmaf(f.y, f.y, b[2 * k + 1]);
        }
      }
    }
  }
  if (cur >= 0) flush(cur);
}

// y(fp16, pitch) = act(x * scale[seg] + shift[seg] (+ residual)), scale / shift from the table vcb_bn_seg_finalize wrote
// (BatchNorm2d training mode: biased variance, eps).  A block walks `cpb` consecutive crops and reloads the table row into shared
// memory only when the segment changes.  pool != 0: the 3x3 / stride 2 / pad 1 max-pool of model.py:57 applied to the normalised,
// activated map (h x w -> ceil(h/2) x ceil(w/2)).
__global__ void __launch_bounds__(256) bn_seg_apply_f16_kernel(const uint4* __restrict__ x, int c8, int h, int w, int n, int cpb,
                                                                const int* __restrict__ seg_of_crop, const float2* __restrict__ affine,
                                                                const uint4* __restrict__ residual, int res_pitch8, int act, int pool,
                                                                uint4* __restrict__ y, int y_pitch8) {
  __shared__ float sc[512], sh[512];
  const int c = c8 * 8, hw = h * w;
  const int crop0 = blockIdx.x * cpb, crop1 = min(n, crop0 + cpb);
  int cur = -1, cur_regs = -1;
  float rs[8], rh[8];                                 // scale / shift of this thread's 8 channels (non-pool path)
  auto norm8 = [&](const uint4& v, int cv, float* f) {
    const __half2* hh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = __half22float2(hh[k]);
      f[2 * k] = fmaf(t.x, sc[cv * 8 + 2 * k], sh[cv * 8 + 2 * k]);
      f[2 * k + 1] = fmaf(t.y, sc[cv * 8 + 2 * k + 1], sh[cv * 8 + 2 * k + 1]);
    }
  };
  for (int crop = crop0; crop < crop1; ++crop) {
    const int seg = seg_of_crop[crop];
    if (seg != cur) {                                     // block-uniform
      __syncthreads();
      for (int ch = threadIdx.x; ch < c; ch += 256) {
        const float2 k = __ldg(affine + (long long)seg * c + ch);
        sc[ch] = k.x; sh[ch] = k.y;
      }
      __syncthreads();
      cur = seg;
    }
    const uint4* xb = x + (long long)crop * hw * c8;
    if (!pool) {
      // c8 divides 256 (power of two <= 64): a thread always meets the same 8 channels, so their scale / shift live in registers for
      // the whole segment and the loop is loads -> 8 FMAs -> store, four independent rows in flight per thread
      const int cv = threadIdx.x & (c8 - 1), rl = threadIdx.x / c8, lanes = 256 / c8;
      if (seg != cur_regs) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { rs[k] = sc[cv * 8 + k]; rh[k] = sh[cv * 8 + k]; }
        cur_regs = seg;
      }
      const long long row0 = (long long)crop * hw;
      for (int r0 = rl; r0 < hw; r0 += 4 * lanes) {
        uint4 v[4], rv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * lanes;
          v[u] = r < hw ? __ldg(xb + (long long)r * c8 + cv) : make_uint4(0u, 0u, 0u, 0u);
          if (residual != nullptr) rv[u] = r < hw ? __ldg(residual + (row0 + r) * res_pitch8 + cv) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u * lanes;
          if (r >= hw) break;
          const __half2* hh = reinterpret_cast<const __half2*>(&v[u]);
          float f[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 t = __half22float2(hh[k]);
            f[2 * k] = fmaf(t.x, rs[2 * k], rh[2 * k]);
            f[2 * k + 1] = fmaf(t.y, rs[2 * k + 1], rh[2 * k + 1]);
          }
          if (residual != nullptr) {
            const __half2* rr = reinterpret_cast<const __half2*>(&rv[u]);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 t = __half22float2(rr[k]); f[2 * k] += t.x; f[2 * k + 1] += t.y; }
          }
          if (act == VCB_ACT_RELU) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
          }
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
          y[(row0 + r) * y_pitch8 + cv] = o;
        }
      }
    } else {
      const int ho = (h + 1) / 2, wo = (w + 1) / 2;
      const int total = ho * wo * c8;
      for (int i = threadIdx.x; i < total; i += 256) {
        const int cv = i % c8, pq = i / c8;
        const int oy = pq / wo, ox = pq - oy * wo;
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = -3.0e38f;
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = 2 * oy + dy;
          if ((unsigned)yy >= (unsigned)h) continue;
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = 2 * ox + dx;
            if ((unsigned)xx >= (unsigned)w) continue;
            float f[8];
            norm8(__ldg(xb + (long long)(yy * w + xx) * c8 + cv), cv, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], act == VCB_ACT_RELU ? fmaxf(f[k], 0.0f) : f[k]);
          }
        }
        // rounding to fp16 is monotonic, so the rounded maximum equals the maximum of the rounded (stored) activations
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(m[2 * k], m[2 * k + 1]);
        y[((long long)crop * ho * wo + pq) * y_pitch8 + cv] = o;
      }
    }
  }
}

// Train-mode BatchNorm apply with the scale / shift derived IN the kernel from the statistics (no vcb_bn_seg_finalize launch, no affine
// table) and, optionally, a residual that is itself a pre-BN tensor (the downsample branch of a BasicBlock, model.py:19-25,33-37):
//   y = act( x * k + (beta - mean * k)  +  [ r * kr + (beta_r - mean_r * kr)  |  r ] ),   k = gamma / sqrt(var + eps)
// mean / var per (segment, channel) from sums[seg][c][2] over seg_crops[seg] * hw values, computed once per block and segment in
// shared memory with the arithmetic of bn_seg_finalize_kernel; a thread then keeps the 8 channels it always meets in registers.
__global__ void __launch_bounds__(256, 3) bn_seg_apply_fused_f16_kernel(const uint4* __restrict__ x, int c8, int hw, int n, int cpb,
                                                                      const int* __restrict__ seg_of_crop, const int* __restrict__ seg_crops,
                                                                      const double* __restrict__ sums, const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, float eps,
                                                                      const uint4* __restrict__ residual, int res_pitch8,
                                                                      const double* __restrict__ res_sums, const float* __restrict__ res_gamma,
                                                                      const float* __restrict__ res_beta, int act, uint4* __restrict__ y, int y_pitch8) {
  __shared__ float sc[512], sh[512], rc[512];
  const int c = c8 * 8;
  const int crop0 = blockIdx.x * cpb, crop1 = min(n, crop0 + cpb);
  const int cv = threadIdx.x & (c8 - 1), rl = threadIdx.x / c8, lanes = 256 / c8;
  int cur = -1;
  float rs[8], rh[8], rq[8];
  for (int crop = crop0; crop < crop1; ++crop) {
    const int seg = seg_of_crop[crop];
    if (seg != cur) {                                     // block-uniform
      __syncthreads();
      const double cnt = (double)seg_crops[seg] * (double)hw;
      for (int ch = threadIdx.x; ch < c; ch += 256) {
        const long long i = (long long)seg * c + ch;
        const double mean = cnt > 0 ? sums[i * 2] / cnt : 0.0;
        double var = cnt > 0 ? sums[i * 2 + 1] / cnt - mean * mean : 0.0;
        if (var < 0) var = 0;
        const float k = gamma[ch] * (float)(1.0 / sqrt(var + (double)eps));
        float shift = beta[ch] + (0.0f - (float)mean) * k;
        float kr = 1.0f;
        if (res_sums != nullptr) {
          const double mr = cnt > 0 ? res_sums[i * 2] / cnt : 0.0;
          double vr = cnt > 0 ? res_sums[i * 2 + 1] / cnt - mr * mr : 0.0;
          if (vr < 0) vr = 0;
          kr = res_gamma[ch] * (float)(1.0 / sqrt(vr + (double)eps));
          shift += res_beta[ch] + (0.0f - (float)mr) * kr;
        }
        sc[ch] = k; sh[ch] = shift; rc[ch] = kr;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) { rs[k] = sc[cv * 8 + k]; rh[k] = sh[cv * 8 + k]; rq[k] = rc[cv * 8 + k]; }
      cur = seg;
    }
    const uint4* xb = x + (long long)crop * hw * c8;
    const long long row0 = (long long)crop * hw;
    for (int r0 = rl; r0 < hw; r0 += 4 * lanes) {
      uint4 v[4], rv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * lanes;
        v[u] = r < hw ? __ldg(xb + (long long)r * c8 + cv) : make_uint4(0u, 0u, 0u, 0u);
        if (residual != nullptr) rv[u] = r < hw ? __ldg(residual + (row0 + r) * res_pitch8 + cv) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * lanes;
        if (r >= hw) break;
        const __half2* hh = reinterpret_cast<const __half2*>(&v[u]);
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t = __half22float2(hh[k]);
          f[2 * k] = fmaf(t.x, rs[2 * k], rh[2 * k]);
          f[2 * k + 1] = fmaf(t.y, rs[2 * k + 1], rh[2 * k + 1]);
        }
        if (residual != nullptr) {
          const __half2* rr = reinterpret_cast<const __half2*>(&rv[u]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 t = __half22float2(rr[k]);
            f[2 * k] = fmaf(t.x, rq[2 * k], f[2 * k]);
            f[2 * k + 1] = fmaf(t.y, rq[2 * k + 1], f[2 * k + 1]);
          }
        }
        if (act == VCB_ACT_RELU) {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
        y[(row0 + r) * y_pitch8 + cv] = o;
      }
    }
  }
}

static int crops_per_block(int n, int hw, int c, int target_bytes) {
  long long per_crop = (long long)hw * c * 2;
  int cpb = (int)(target_bytes / (per_crop > 0 ? per_crop : 1));
  if (cpb < 1) cpb = 1;
  if (cpb > 32) cpb = 32;
  while (cpb > 1 && (n + cpb - 1) / cpb < 2 * 148) cpb >>= 1;       // keep at least two blocks per SM
  return cpb;
}

int bn_seg_stats_f16(const void* x, int c, int hw, int n, const int* seg_of_crop, double* sums, cudaStream_t st) {
  const int c8 = c / 8;
  if (!x || !seg_of_crop || !sums || n <= 0 || hw <= 0 || c <= 0 || (c & 7) || c8 > 64 || (c8 & (c8 - 1)) || ((uintptr_t)x & 15))
    return set_error(VCB_ERR_INVALID, "bn_seg_stats_f16: bad argument (c must be 8 * a power of two, <= 512)");
  // ~384 KiB per block: the flush (two barriers, a shared-memory reduction, c * 2 fp64 atomics) stays small against the streaming
  // time (measured on 4096 crops: 0.885 -> 0.79 ms over the 20 layers against 128 KiB)
  const int cpb = crops_per_block(n, hw, c, 393216);
  bn_seg_stats_f16_kernel<<<(n + cpb - 1) / cpb, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), c8, hw, n, cpb, seg_of_crop, sums);
  return check_cuda(cudaGetLastError(), "bn_seg_stats_f16 launch");
}

int bn_seg_apply_f16(const void* x, int c, int h, int w, int n, const int* seg_of_crop, const float* affine, const void* residual,
                     int res_pitch, int act, int pool, void* y, int y_pitch, cudaStream_t st) {
  if (!x || !seg_of_crop || !affine || !y || n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c & 7) || c > 512 || ((c / 8) & (c / 8 - 1)) || (y_pitch & 7) ||
      (residual && (res_pitch & 7)) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15) || ((uintptr_t)residual & 15) || ((uintptr_t)affine & 7) ||
      (pool && residual) || (act != VCB_ACT_NONE && act != VCB_ACT_RELU))
    return set_error(VCB_ERR_INVALID, "bn_seg_apply_f16: bad argument (c must be 8 * a power of two, <= 512)");
  const int cpb = crops_per_block(n, h * w, c, 131072);      // more, smaller blocks: the apply pass wants loads in flight (384 KiB: 1.31 -> 1.46 ms)
  bn_seg_apply_f16_kernel<<<(n + cpb - 1) / cpb, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), c / 8, h, w, n, cpb, seg_of_crop,
                                                              reinterpret_cast<const float2*>(affine), reinterpret_cast<const uint4*>(residual),
                                                              res_pitch / 8, act, pool, reinterpret_cast<uint4*>(y), y_pitch / 8);
  return check_cuda(cudaGetLastError(), "bn_seg_apply_f16 launch");
}

int bn_seg_apply_fused_f16(const void* x, int c, int hw, int n, const int* seg_of_crop, const int* seg_crops, const double* sums,
                           const float* gamma, const float* beta, float eps, const void* residual, int res_pitch, const double* res_sums,
                           const float* res_gamma, const float* res_beta, int act, void* y, int y_pitch, cudaStream_t st) {
  if (!x || !seg_of_crop || !seg_crops || !sums || !gamma || !beta || !y || n <= 0 || hw <= 0 || c <= 0 || (c & 7) || c > 512 ||
      ((c / 8) & (c / 8 - 1)) || (y_pitch & 7) || (residual && (res_pitch & 7)) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15) ||
      ((uintptr_t)residual & 15) || (res_sums && (!residual || !res_gamma || !res_beta)) || (act != VCB_ACT_NONE && act != VCB_ACT_RELU))
    return set_error(VCB_ERR_INVALID, "bn_seg_apply_fused_f16: bad argument (c must be 8 * a power of two, <= 512)");
  const int cpb = crops_per_block(n, hw, c, 131072);
  bn_seg_apply_fused_f16_kernel<<<(n + cpb - 1) / cpb, 256, 0, st>>>(
      reinterpret_cast<const uint4*>(x), c / 8, hw, n, cpb, seg_of_crop, seg_crops, sums, gamma, beta, eps, reinterpret_cast<const uint4*>(residual),
      res_pitch / 8, res_sums, res_gamma, res_beta, act, reinterpret_cast<uint4*>(y), y_pitch / 8);
  return check_cuda(cudaGetLastError(), "bn_seg_apply_fused_f16 launch");
}

}  // namespace vcb
