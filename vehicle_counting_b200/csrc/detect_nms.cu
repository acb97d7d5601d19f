// K3 / K4 -- anchor decode + confidence filter + best class, and class-aware greedy NMS.
//
// Restates, for the GPU, upstream YOLOv5 v6.0 `Detect.forward` (inference branch) and
// `non_max_suppression` + `torchvision.ops.nms`, which the reference reaches through
// /root/reference/networks/yolo.py:70 (AutoShape call) with the settings of :62-66
// (conf, iou, classes=None, multi_label=False, max_det).  Both kernels are HBM/latency bound:
// decode reads P*85 fp16 logits per frame once (coalesced over the 85-field record by a warp) and
// writes only the surviving candidates; NMS is one CTA per frame (sort in shared memory, then a
// blocked greedy scan with O(n) memory instead of torchvision's O(n^2/64) mask).
#include <float.h>

#include "vcb_internal.h"

namespace vcb {

struct DecodeParams {
  VcbDetectDesc d;
  int level_pred_start[5];   // prefix sums of 3*ny*nx over levels
  int total_groups;          // warps of work: sum over (frame, level, anchor) of ceil(ny*nx/32)
  int groups_per_frame;
  int level_group_start[5];  // within a frame: prefix of 3*ceil(cells/32)
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T>
__device__ __forceinline__ float load_logit(const void* base, long long idx) {
  return (float)reinterpret_cast<const T*>(base)[idx];
}
template <>
__device__ __forceinline__ float load_logit<__half>(const void* base, long long idx) {
  return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}

template <typename T>
__global__ void detect_decode_kernel(const DecodeParams prm, float* __restrict__ cand_box, float* __restrict__ cand_score,
                                     int* __restrict__ cand_cls, int* __restrict__ cand_index, int* __restrict__ cand_count) {
  const VcbDetectDesc& d = prm.d;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int no = d.nc + 5;
  for (int g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < prm.total_groups; g += gridDim.x * warps_per_block) {
    const int b = g / prm.groups_per_frame;
    int gi = g - b * prm.groups_per_frame;
    int l = 0;
    while (l + 1 < d.num_levels && gi >= prm.level_group_start[l + 1]) ++l;
    gi -= prm.level_group_start[l];
    const VcbDetectLevel& L = d.level[l];
    const int cells = L.ny * L.nx;
    const int groups_per_anchor = (cells + 31) >> 5;
    const int a = gi / groups_per_anchor;
    const int cell0 = (gi - a * groups_per_anchor) << 5;
    const long long frame_base = (long long)b * cells * L.pitch;

    // phase 1: objectness of 32 consecutive cells
    const int cell = cell0 + lane;
    bool pass = false;
    if (cell < cells) {
      const float obj = sigmoidf_acc(load_logit<T>(L.logits, frame_base + (long long)cell * L.pitch + a * no + 4));
      pass = obj > d.conf_thres;
    }
    unsigned mask = __ballot_sync(0xffffffffu, pass);
    // phase 2: the warp decodes each surviving cell together (fields spread over lanes)
    while (mask) {
      const int src_lane = __ffs(mask) - 1;
      mask &= mask - 1;
      const int cl = cell0 + src_lane;
      const long long rec = frame_base + (long long)cl * L.pitch + a * no;
      float f0 = 0.f, best = -1.0f;
      int best_j = 0x7fffffff;
      float head[5];
      // lanes 0..4 fetch x,y,w,h,obj; every lane scans classes lane, lane+32, ...
      if (lane < 5) f0 = sigmoidf_acc(load_logit<T>(L.logits, rec + lane));
#pragma unroll
      for (int i = 0; i < 5; ++i) head[i] = __shfl_sync(0xffffffffu, f0, i);
      const float obj = head[4];
      for (int j = lane; j < d.nc; j += 32) {
        const float s = sigmoidf_acc(load_logit<T>(L.logits, rec + 5 + j)) * obj;   // x[:, 5:] *= x[:, 4:5]
        if (s > best) { best = s; best_j = j; }
      }
      for (int o = 16; o > 0; o >>= 1) {           // argmax, first index wins ties
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
      }
      const bool cls_ok = !d.use_class_mask || best_j >= 256 || ((d.class_mask[best_j >> 5] >> (best_j & 31)) & 1u);
      if (lane == 0 && best > d.conf_thres && cls_ok) {
        const int gy = cl / L.nx, gx = cl - gy * L.nx;
        const float cx = (head[0] * 2.0f - 0.5f + (float)gx) * L.stride;
        const float cy = (head[1] * 2.0f - 0.5f + (float)gy) * L.stride;
        const float tw = head[2] * 2.0f, th = head[3] * 2.0f;
        const float bw = tw * tw * L.anchor_w[a], bh = th * th * L.anchor_h[a];
        const int slot = atomicAdd(cand_count + b, 1);
        if (slot < d.max_candidates) {
          const long long o = (long long)b * d.max_candidates + slot;
          float4 bx;
          bx.x = cx - bw / 2.0f;     // xywh2xyxy
          bx.y = cy - bh / 2.0f;
          bx.z = cx + bw / 2.0f;
          bx.w = cy + bh / 2.0f;
          reinterpret_cast<float4*>(cand_box)[o] = bx;
          cand_score[o] = best;
          cand_cls[o] = best_j;
          cand_index[o] = prm.level_pred_start[l] + a * cells + cl;
        }
      }
    }
  }
}

int detect_decode(const VcbDetectDesc& d, float* cand_box, float* cand_score, int* cand_cls, int* cand_index, int* cand_count,
                  cudaStream_t st) {
  if (d.n <= 0 || d.nc <= 0 || d.num_levels <= 0 || d.num_levels > 4 || d.max_candidates <= 0 || !cand_box || !cand_score ||
      !cand_cls || !cand_index || !cand_count)
    return set_error(VCB_ERR_INVALID, "detect_decode: bad argument");
  if ((uintptr_t)cand_box & 15) return set_error(VCB_ERR_INVALID, "detect_decode: cand_box must be 16-byte aligned");
  DecodeParams prm;
  prm.d = d;
  int pred = 0, grp = 0;
  for (int l = 0; l < d.num_levels; ++l) {
    const VcbDetectLevel& L = d.level[l];
    if (!L.logits || L.ny <= 0 || L.nx <= 0 || L.pitch < 3 * (d.nc + 5)) return set_error(VCB_ERR_INVALID, "detect_decode: bad level");
    prm.level_pred_start[l] = pred;
    prm.level_group_start[l] = grp;
    pred += 3 * L.ny * L.nx;
    grp += 3 * ((L.ny * L.nx + 31) / 32);
  }
  prm.level_pred_start[d.num_levels] = pred;
  prm.level_group_start[d.num_levels] = grp;
  prm.groups_per_frame = grp;
  prm.total_groups = grp * d.n;
  cudaError_t e = cudaMemsetAsync(cand_count, 0, sizeof(int) * d.n, st);
  if (e != cudaSuccess) return check_cuda(e, "detect_decode memset");
  const int block = 256, wpb = block / 32;
  int grid = (prm.total_groups + wpb - 1) / wpb;
  if (grid > 148 * 8) grid = 148 * 8;
  if (d.logits_dtype == VCB_F32)
    detect_decode_kernel<float><<<grid, block, 0, st>>>(prm, cand_box, cand_score, cand_cls, cand_index, cand_count);
  else
    detect_decode_kernel<__half><<<grid, block, 0, st>>>(prm, cand_box, cand_score, cand_cls, cand_index, cand_count);
  return check_cuda(cudaGetLastError(), "detect_decode launch");
}

// ================================================================================================
// NMS: one CTA per frame
// ================================================================================================
constexpr int kNmsThreads = 1024;
constexpr int kNmsChunk = 64;
constexpr int kSmemSortMax = 8192;     // candidates sorted in shared memory (12 B each)

__host__ __device__ inline long long nms_ws_u64_per_frame(int max_candidates) {
  long long p2 = 1;
  while (p2 < max_candidates) p2 <<= 1;
  return p2 + p2 / 2 + 2;      // npad keys (8 B) + npad vals (4 B)
}
long long nms_workspace_bytes(int n, int max_candidates) { return (long long)n * nms_ws_u64_per_frame(max_candidates) * 8; }

__device__ __forceinline__ void bitonic_sort_pairs(unsigned long long* keys, int* vals, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b; keys[ixj] = a;
            const int t = vals[i]; vals[i] = vals[ixj]; vals[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

// IoU test with torchvision's float32 arithmetic: areas (x2-x1)*(y2-y1), inter / (a + b - inter) > thr
__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ovr > thr;
}

struct NmsSmemLayout {
  int kept_box, kept_area, kept_slot, ch_box, ch_area, ch_slot, ch_supp, ch_mask, nkept, sort, total;
};
__host__ __device__ inline NmsSmemLayout nms_smem_layout(int max_det) {
  NmsSmemLayout l;
  int o = 0;
  auto take = [&](int bytes) { const int r = o; o = (o + bytes + 15) & ~15; return r; };
  l.kept_box = take(max_det * 16);
  l.kept_area = take(max_det * 4);
  l.kept_slot = take(max_det * 4);
  l.ch_box = take(kNmsChunk * 16);
  l.ch_area = take(kNmsChunk * 4);
  l.ch_slot = take(kNmsChunk * 4);
  l.ch_supp = take(kNmsChunk * 4);
  l.ch_mask = take(kNmsChunk * 8);
  l.nkept = take(16);
  l.sort = take(kSmemSortMax * 12);
  l.total = o;
  return l;
}

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_kernel(const VcbNmsDesc d, const float* __restrict__ cand_box, const float* __restrict__ cand_score,
           const int* __restrict__ cand_cls, const int* __restrict__ cand_index, const int* __restrict__ cand_count,
           unsigned long long* __restrict__ ws, float* __restrict__ det, int* __restrict__ det_count) {
  extern __shared__ unsigned char nms_smem[];
  const int b = blockIdx.x;
  int n = cand_count[b];
  if (n > d.max_candidates) n = d.max_candidates;
  const long long base = (long long)b * d.max_candidates;

  // shared-memory layout (byte offsets, every array 16-byte aligned)
  const NmsSmemLayout lay = nms_smem_layout(d.max_det);
  float4* kept_box = reinterpret_cast<float4*>(nms_smem + lay.kept_box);                 // [max_det] offset boxes
  float* kept_area = reinterpret_cast<float*>(nms_smem + lay.kept_area);                 // [max_det]
  int* kept_slot = reinterpret_cast<int*>(nms_smem + lay.kept_slot);                     // [max_det]
  float4* ch_box = reinterpret_cast<float4*>(nms_smem + lay.ch_box);                     // [64]
  float* ch_area = reinterpret_cast<float*>(nms_smem + lay.ch_area);                     // [64]
  int* ch_slot = reinterpret_cast<int*>(nms_smem + lay.ch_slot);                         // [64]
  int* ch_supp = reinterpret_cast<int*>(nms_smem + lay.ch_supp);                         // [64]
  unsigned long long* ch_mask = reinterpret_cast<unsigned long long*>(nms_smem + lay.ch_mask);   // [64]
  int* s_nkept = reinterpret_cast<int*>(nms_smem + lay.nkept);
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(nms_smem + lay.sort);
  int npad = 1;
  while (npad < n) npad <<= 1;
  unsigned long long* keys;
  int* vals;
  if (npad <= kSmemSortMax) {
    keys = s_keys;
    vals = reinterpret_cast<int*>(s_keys + npad);
  } else {   // rare: spill the sort to the global workspace (12 B x next_pow2(max_candidates) per frame)
    keys = ws + (long long)b * nms_ws_u64_per_frame(d.max_candidates);
    vals = reinterpret_cast<int*>(keys + npad);
  }
  if (threadIdx.x == 0) s_nkept[0] = 0;
  // key: descending score, then ascending prediction index
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    if (i < n) {
      const unsigned sbits = __float_as_uint(cand_score[base + i]);          // scores are positive
      keys[i] = ((unsigned long long)(0xffffffffu - sbits) << 32) | (unsigned)cand_index[base + i];
      vals[i] = i;
    } else {
      keys[i] = ~0ull;
      vals[i] = -1;
    }
  }
  __syncthreads();
  bitonic_sort_pairs(keys, vals, npad);
  if (n > d.max_nms) n = d.max_nms;                                           // x[x[:, 4].argsort(descending=True)[:max_nms]]

  int nkept = 0;
  for (int c0 = 0; c0 < n && nkept < d.max_det; c0 += kNmsChunk) {
    const int cn = min(kNmsChunk, n - c0);
    if (threadIdx.x < kNmsChunk) {
      const int t = threadIdx.x;
      if (t < cn) {
        const int slot = vals[c0 + t];
        float4 bx = reinterpret_cast<const float4*>(cand_box)[base + slot];
        const float off = __fmul_rn((float)cand_cls[base + slot], d.max_wh);   // c = x[:, 5:6] * max_wh
        bx.x = __fadd_rn(bx.x, off); bx.y = __fadd_rn(bx.y, off);
        bx.z = __fadd_rn(bx.z, off); bx.w = __fadd_rn(bx.w, off);
        ch_box[t] = bx;
        ch_area[t] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        ch_slot[t] = slot;
      }
      ch_supp[t] = (t < cn) ? 0 : 1;
    }
    __syncthreads();
    // phase A: suppression by boxes kept in earlier chunks
    for (int i = threadIdx.x; i < cn * nkept; i += blockDim.x) {
      const int c = i % cn, k = i / cn;
      if (iou_gt(kept_box[k], kept_area[k], ch_box[c], ch_area[c], d.iou_thres)) ch_supp[c] = 1;
    }
    // phase B: pairwise mask inside the chunk (j > c)
    if (threadIdx.x < kNmsChunk) {
      const int c = threadIdx.x;
      unsigned long long m = 0;
      if (c < cn) {
        const float4 bc = ch_box[c];
        const float ac = ch_area[c];
        for (int j = c + 1; j < cn; ++j)
          if (iou_gt(bc, ac, ch_box[j], ch_area[j], d.iou_thres)) m |= 1ull << j;
      }
      ch_mask[c] = m;
    }
    __syncthreads();
    // phase C: serial greedy resolve of the chunk
    if (threadIdx.x == 0) {
      unsigned long long supp = 0;
      int nk = nkept;
      for (int c = 0; c < cn && nk < d.max_det; ++c) {
        if (ch_supp[c] || ((supp >> c) & 1ull)) continue;
        kept_box[nk] = ch_box[c];
        kept_area[nk] = ch_area[c];
        kept_slot[nk] = ch_slot[c];
        ++nk;
        supp |= ch_mask[c];
      }
      s_nkept[0] = nk;
    }
    __syncthreads();
    nkept = s_nkept[0];
  }

  // output rows (x1,y1,x2,y2,conf,cls), scale_coords + clip when requested
  const bool rescale = d.gain != nullptr;
  float gain = 1.f, padx = 0.f, pady = 0.f, w0 = FLT_MAX, h0 = FLT_MAX;
  if (rescale) { gain = d.gain[b]; padx = d.pad_x[b]; pady = d.pad_y[b]; w0 = d.w0[b]; h0 = d.h0[b]; }
  for (int i = threadIdx.x; i < nkept; i += blockDim.x) {
    const int slot = kept_slot[i];
    float4 bx = reinterpret_cast<const float4*>(cand_box)[base + slot];
    if (rescale) {
      bx.x = __fdiv_rn(__fsub_rn(bx.x, padx), gain); bx.z = __fdiv_rn(__fsub_rn(bx.z, padx), gain);
      bx.y = __fdiv_rn(__fsub_rn(bx.y, pady), gain); bx.w = __fdiv_rn(__fsub_rn(bx.w, pady), gain);
      bx.x = fminf(fmaxf(bx.x, 0.f), w0); bx.z = fminf(fmaxf(bx.z, 0.f), w0);
      bx.y = fminf(fmaxf(bx.y, 0.f), h0); bx.w = fminf(fmaxf(bx.w, 0.f), h0);
    }
    float* o = det + ((long long)b * d.max_det + i) * 6;
    o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
    o[4] = cand_score[base + slot];
    o[5] = (float)cand_cls[base + slot];
  }
  if (threadIdx.x == 0) det_count[b] = nkept;
}

int nms(const VcbNmsDesc& d, const float* cand_box, const float* cand_score, const int* cand_cls, const int* cand_index,
        const int* cand_count, unsigned long long* ws, float* det, int* det_count, cudaStream_t st) {
  if (d.n <= 0 || d.max_candidates <= 0 || d.max_det <= 0 || d.max_det > 4096 || d.max_nms <= 0 || !cand_box || !cand_score ||
      !cand_cls || !cand_index || !cand_count || !ws || !det || !det_count)
    return set_error(VCB_ERR_INVALID, "nms: bad argument");
  if (d.gain && (!d.pad_x || !d.pad_y || !d.w0 || !d.h0)) return set_error(VCB_ERR_INVALID, "nms: incomplete rescale arrays");
  const size_t smem = (size_t)nms_smem_layout(d.max_det).total;
  static bool attr = false;
  if (!attr) {
    const cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(nms)");
    attr = true;
  }
  if (smem > 200 * 1024) return set_error(VCB_ERR_INVALID, "nms: max_det too large");
  nms_kernel<<<d.n, kNmsThreads, smem, st>>>(d, cand_box, cand_score, cand_cls, cand_index, cand_count, ws, det, det_count);
  return check_cuda(cudaGetLastError(), "nms launch");
}

}  // namespace vcb
