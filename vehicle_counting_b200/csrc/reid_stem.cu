// K5b/K6b -- fused ReID stem: ROI crop+resize+normalise -> Conv3x3(3->64, bias, folded BN) -> ReLU -> MaxPool(3, 2, 1).
//
// Replaces, for the folded-BN ("eval") path,
//   /root/reference/networks/deepsort/deep/feature_extractor.py:26-39  (crop preprocessing, per crop on the CPU)
//   /root/reference/networks/deepsort/deep/model.py:52-60             (stem conv + BN + ReLU + MaxPool2d(3, 2, padding=1))
// The unfused path (roi_resize_norm -> conv_umma 16->64 -> maxpool) writes the 50x50x64 stem map to HBM and reads it
// back (2.6 GB per 4096 crops) and feeds the tensor cores with nine 4-KiB TMA boxes per 128-pixel tile; measured
// 0.14 + 0.56 + 0.41 ms per 4096 crops.  Here:
//   1. roi_stem_patches_kernel (one CTA per ROI) resizes the crop exactly like roi_resize_norm (cv2 INTER_LINEAR
//      semantics, fp16 rounding of the normalised pixel) into shared memory and writes the stem's im2col operand:
//      the 25x25 pooled map is cut into 5x5 blocks of 5x5 pooled pixels; a block needs the 11x11 conv outputs
//      (rows 10*by-1 .. 10*by+9), one GEMM row each, K = 27 (tap-major, channel-minor) padded to 32:
//      patches[roi][block 0..24][row 0..127][32] fp16, rows >= 121 and out-of-map taps zero.
//   2. reid_stem_pool_kernel (persistent, tcgen05): per block ONE 8-KiB TMA load, two tcgen05.mma (M=128, N=64, K=16),
//      epilogue TMEM -> +bias -> ReLU -> fp16 -> shared memory, then the 3x3/s2 max over the staged 11x11x64 tile
//      and 16-byte stores of the 5x5x64 pooled block.  Conv outputs of row/column -1 (blocks on the top/left edge)
//      are padding for the pool: they are zeroed, which cannot win a max over ReLU outputs (every window holds at
//      least one in-map value).
#include "vcb_internal.h"
#include "vcb_ptx.cuh"

namespace vcb {

constexpr int kStemBlocks = 25;         // 5 x 5 blocks per crop
constexpr int kStemRows = 128;          // GEMM rows per block (121 used)
constexpr int kStemK = 32;              // 27 used
constexpr int kStemN = 64;
constexpr int kStemTileBytes = kStemRows * kStemK * 2;   // 8 KiB
constexpr int kStemStages = 8;
constexpr int kStemAccs = 4;
constexpr int kStemThreads = 320;       // warps 0-7 epilogue + pooling, warp 8 TMA producer, warp 9 MMA issuer

__global__ void __launch_bounds__(256) roi_stem_patches_kernel(const VcbRoiDesc d, const uint8_t* __restrict__ frames, int fh, int fw,
                                                               const int* __restrict__ rois, uint4* __restrict__ out) {
  constexpr int S = 50;
  __shared__ __half crop[S * S * 3];
  const int r = blockIdx.x;
  const int f = rois[r * 5 + 0], x1 = rois[r * 5 + 1], y1 = rois[r * 5 + 2], x2 = rois[r * 5 + 3], y2 = rois[r * 5 + 4];
  const int cw = x2 - x1, chh = y2 - y1;
  uint4* o = out + (long long)r * kStemBlocks * kStemRows * (kStemK * 2 / 16);
  if (cw <= 0 || chh <= 0 || x1 < 0 || y1 < 0 || x2 > fw || y2 > fh || f < 0 || (d.num_frames > 0 && f >= d.num_frames)) {   // the reference would raise inside cv2.resize
    for (int i = threadIdx.x; i < kStemBlocks * kStemRows * 4; i += blockDim.x) o[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const uint8_t* img = frames + (long long)f * fh * fw * 3;
  const double sx = (double)cw / (double)S, sy = (double)chh / (double)S;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {      // identical arithmetic to roi_resize_norm_kernel
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
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = (float)p00[c] / 255.0f, b = (float)p01[c] / 255.0f;
      const float e = (float)p10[c] / 255.0f, g = (float)p11[c] / 255.0f;
      const float top = a * (1.0f - fx) + b * fx;
      const float bot = e * (1.0f - fx) + g * fx;
      const float val = top * (1.0f - fy) + bot * fy;
      crop[i * 3 + c] = __float2half_rn((val - d.mean[c]) * d.inv_std[c]);
    }
  }
  __syncthreads();
  // im2col rows: one thread per (block, row): the 27 taps are three runs of nine consecutive halves of the staged crop
  // (k = (r*3+s)*3 + c), written as four 16-byte stores (64 contiguous bytes per thread, consecutive rows across the warp)
  const unsigned short* cs = reinterpret_cast<const unsigned short*>(crop);
  for (int i = threadIdx.x; i < kStemBlocks * kStemRows; i += blockDim.x) {
    const int row = i & (kStemRows - 1), blk = i >> 7;
    const int by = blk / 5, bx = blk - by * 5;
    const int ti = row / 11, tj = row - ti * 11;
    const int cy = 10 * by - 1 + ti, cx = 10 * bx - 1 + tj;          // conv output position of this row
    const bool row_ok = row < 121 && cy >= 0 && cx >= 0;
    unsigned short v[32];
#pragma unroll
    for (int e = 27; e < 32; ++e) v[e] = 0;
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      const int yy = cy + rr - 1;
      const bool y_ok = row_ok && (unsigned)yy < (unsigned)S;
#pragma unroll
      for (int ss = 0; ss < 3; ++ss) {
        const int xx = cx + ss - 1;
        const bool ok = y_ok && (unsigned)xx < (unsigned)S;
        const int base = ok ? (yy * S + xx) * 3 : 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(rr * 3 + ss) * 3 + c] = ok ? cs[base + c] : (unsigned short)0;
      }
    }
    uint4 w4[4];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      w4[qd].x = (uint32_t)v[8 * qd + 0] | ((uint32_t)v[8 * qd + 1] << 16);
      w4[qd].y = (uint32_t)v[8 * qd + 2] | ((uint32_t)v[8 * qd + 3] << 16);
      w4[qd].z = (uint32_t)v[8 * qd + 4] | ((uint32_t)v[8 * qd + 5] << 16);
      w4[qd].w = (uint32_t)v[8 * qd + 6] | ((uint32_t)v[8 * qd + 7] << 16);
    }
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) o[(long long)i * 4 + qd] = w4[qd];
  }
}

// weights fp16 [64][32] (K-major, k = (r*3+s)*3+c, zero padded), bias fp32 [64]; out fp16 [num_rois][25][25][64]
// MODE 0: BatchNorm folded into the weights / bias (eval statistics): conv + bias -> ReLU -> max-pool.
// Train-mode BatchNorm (the reference as shipped: statistics of one Extractor call = one SEGMENT of crops) runs the SAME kernel
// twice instead of writing the 50x50x64 pre-BN map (1.3 GB per 4096 crops) to HBM and reading it back twice:
// MODE 1: statistics only -- per-(segment, channel) sum and sum of squares of conv + bias over the 50x50 outputs (each counted
//         once: the interior 10x10 of a block's 11x11 tile), accumulated in registers across the CTA's tiles and flushed with
//         double atomics when the segment changes; nothing else is written.
// MODE 2: conv -> per-segment scale / shift (vcb_bn_seg_finalize: gamma / sqrt(var + eps), bias and mean folded into the shift)
//         -> ReLU -> max-pool.  The stem conv is 1 % of the network's FLOPs, so computing it twice is cheaper than the traffic.
// MODE 1 / 2 give every CTA a CONTIGUOUS range of tiles (crops), so a segment change is rare.
template <int MODE>
__global__ void __launch_bounds__(kStemThreads, MODE == 1 ? 1 : 2)
reid_stem_pool_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ bias,
                      __half* __restrict__ out, int num_tiles, KernelFault* fault, const int* __restrict__ seg_of_crop,
                      double* __restrict__ sums, const float* __restrict__ affine) {
  // tile walk: strided (MODE 0) or a contiguous range per CTA
  const int per_cta = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t_begin = MODE == 0 ? (int)blockIdx.x : (int)blockIdx.x * per_cta;
  const int t_end = MODE == 0 ? num_tiles : min(num_tiles, t_begin + per_cta);
  const int t_step = MODE == 0 ? (int)gridDim.x : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [8 x 8 KiB A stages][4 KiB weights][2 x 16 KiB conv-out staging][barriers][tmem slot][bias]
  const uint32_t w_smem = smem_base + kStemStages * kStemTileBytes;
  const uint32_t stage0 = w_smem + 4096u;
  const uint32_t bars = stage0 + 2u * 16384u;
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(kStemStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kStemStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kStemStages + kStemAccs + a); };
  const uint32_t w_bar = bars + 8u * (2 * kStemStages + 2 * kStemAccs);
  const uint32_t tmem_slot = w_bar + 8u;
  const uint32_t tail_off = (uint32_t)(kStemStages * kStemTileBytes + 4096 + 2 * 16384 + 8 * (2 * kStemStages + 2 * kStemAccs + 1));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + tail_off);
  float* bias_s = reinterpret_cast<float*>(smem_gen + tail_off + 16);
  uint8_t* stage_gen = smem_gen + kStemStages * kStemTileBytes + 4096;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (MODE != 2 && threadIdx.x < kStemN) bias_s[threadIdx.x] = __ldg(bias + threadIdx.x);      // MODE 2: the bias lives in the shift
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStemStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < kStemAccs; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_w); }
  if (warp == 9) { tmem_alloc(tmem_slot, 256u); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 8) {
    if (lane == 0) {       // ---- TMA producer: the weights once, then one 8-KiB patch tile per block
      mbar_arrive_expect_tx(w_bar, 4096u);
      tma_load_2d(&tmap_w, w_bar, w_smem, 0, 0);
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
        const int st = it % kStemStages;
        mbar_wait(empty_bar(st), ((it / kStemStages) & 1u) ^ 1u, fault, FAULT_EMPTY_WAIT, 500 + st);
        mbar_arrive_expect_tx(full_bar(st), (uint32_t)kStemTileBytes);
        tma_load_2d(&tmap_a, full_bar(st), smem_base + (uint32_t)st * kStemTileBytes, 0, tile * kStemRows);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {       // ---- MMA issuer: D[128 x 64] = A[128 x 32] * W[64 x 32]^T, two K=16 steps
      const uint32_t idesc = umma_idesc_f16((uint32_t)kStemN);
      const uint64_t desc_hi = umma_desc_kmajor(0, 8u * kStemK * 2u, 4u);      // SWIZZLE_64B, 8-row groups 512 B apart
      const uint64_t w_desc = desc_hi | (uint64_t)((w_smem & 0x3FFFF) >> 4);
      mbar_wait(w_bar, 0u, fault, FAULT_FULL_WAIT, 510);
      uint32_t it = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
        const int st = it % kStemStages;
        const uint32_t acc = it % kStemAccs;
        mbar_wait(tempty_bar(acc), ((it / kStemAccs) & 1u) ^ 1u, fault, FAULT_TMEM_EMPTY_WAIT, 520 + (int)acc);
        mbar_wait(full_bar(st), (it / kStemStages) & 1u, fault, FAULT_FULL_WAIT, 530 + st);
        tcgen05_fence_after();
        const uint64_t a_desc = desc_hi | (uint64_t)(((smem_base + (uint32_t)st * kStemTileBytes) & 0x3FFFF) >> 4);
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)kStemN;
        umma_f16(d_tmem, a_desc, w_desc, idesc, 0u);
        umma_f16(d_tmem, a_desc + 2u, w_desc + 2u, idesc, 1u);
        umma_commit(empty_bar(st));
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ---- epilogue (256 threads)
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                       // GEMM row = position ti*11 + tj of the 11x11 conv tile
    const int ti = row / 11, tj = row - ti * 11;
    const int sw = row & 7;
    float* scratch = reinterpret_cast<float*>(stage_gen);                // 32 KiB: [128 rows][64 channels] fp32 (MODE 1 flush)
    // MODE 2: [64][2] scale, shift of the current segment, read as float4 -> 16-byte aligned (the slack is in kStemSmemBytes)
    float* aff_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bias_s + kStemN) + 15) & ~(uintptr_t)15);
    float bv[MODE == 2 ? 1 : 32];
    if (MODE != 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) bv[i] = bias_s[half * 32 + i];
    }
    float s1[MODE == 1 ? 32 : 1], s2[MODE == 1 ? 32 : 1];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    }
    int cur_seg = -1;
    // MODE 1: add this CTA's per-thread partial sums of segment `seg` to the global table (all 256 epilogue threads call it)
    auto flush = [&](int seg) {
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) scratch[row * 64 + ((half * 32 + i + row) & 63)] = pass == 0 ? s1[i] : s2[i];   // skewed: conflict-free
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (threadIdx.x < 64) {
          const int ch = threadIdx.x;
          double a = 0.0;
          for (int r = 0; r < 128; ++r) a += (double)scratch[r * 64 + ((ch + r) & 63)];
          atomicAdd(sums + ((long long)seg * kStemN + ch) * 2 + pass, a);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    };
    uint32_t it = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
      const uint32_t acc = it % kStemAccs;
      const int roi = tile / kStemBlocks, blk = tile - roi * kStemBlocks;
      const int by = blk / 5, bx = blk - by * 5;
      if (MODE != 0) {
        const int seg = __ldg(seg_of_crop + roi);
        if (seg != cur_seg) {                              // CTA-uniform: every thread sees the same tile sequence
          if (MODE == 1) {
            if (cur_seg >= 0) flush(cur_seg);
          } else {
            asm volatile("bar.sync 2, 256;" ::: "memory");      // everyone is done with the previous segment's table
            if (threadIdx.x < 128) aff_s[threadIdx.x] = __ldg(affine + (long long)seg * 128 + threadIdx.x);
            asm volatile("bar.sync 2, 256;" ::: "memory");
          }
          cur_seg = seg;
        }
      }
      mbar_wait(tfull_bar(acc), (it / kStemAccs) & 1u, fault, FAULT_TMEM_FULL_WAIT, 540 + (int)acc);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + acc * (uint32_t)kStemN + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16);
      uint32_t v0[16], v1[16];
      tmem_ld_x16(t_row, v0);
      tmem_ld_x16(t_row + 16u, v1);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (MODE == 1) {
        // every conv output of the crop exactly once: rows 1..10 x columns 1..10 of the 11x11 tile (row 0 / column 0 belong to the
        // neighbouring block, or are padding)
        if (row < 121 && ti >= 1 && tj >= 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = __uint_as_float(v0[i]) + bv[i], b = __uint_as_float(v1[i]) + bv[16 + i];
            s1[i] += a; s2[i] = fmaf(a, a, s2[i]);
            s1[16 + i] += b; s2[16 + i] = fmaf(b, b, s2[16 + i]);
          }
        }
        continue;
      }
      // rows that are padding for the pool (conv row/column -1, rows >= 121) contribute zeros
      const bool valid = row < 121 && !(by == 0 && ti == 0) && !(bx == 0 && tj == 0);
      uint32_t h2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a, b;
        const uint32_t* v = i < 8 ? v0 : v1;
        const int c0 = (i < 8 ? 0 : 16) + 2 * (i & 7);
        if (MODE == 2) {
          const float4 k = *reinterpret_cast<const float4*>(aff_s + (half * 32 + c0) * 2);     // scale, shift, scale, shift (broadcast)
          a = fmaf(__uint_as_float(v[2 * (i & 7)]), k.x, k.y);
          b = fmaf(__uint_as_float(v[2 * (i & 7) + 1]), k.z, k.w);
        } else {
          a = __uint_as_float(v[2 * (i & 7)]) + bv[MODE == 2 ? 0 : c0];
          b = __uint_as_float(v[2 * (i & 7) + 1]) + bv[MODE == 2 ? 0 : c0 + 1];
        }
        const __half2 t = __floats2half2_rn(valid ? fmaxf(a, 0.0f) : 0.0f, valid ? fmaxf(b, 0.0f) : 0.0f);
        h2[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
      // staged tile: row pitch 128 B (64 channels), 16-byte chunk c stored at (c ^ (row & 7)): conflict-free both ways
      const uint32_t buf = stage0 + (it & 1u) * 16384u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t dst = buf + (uint32_t)row * 128u + (uint32_t)(((half * 4 + i) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h2[4 * i]), "r"(h2[4 * i + 1]), "r"(h2[4 * i + 2]),
                     "r"(h2[4 * i + 3]) : "memory");
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      // 3x3/s2 max-pool of the 11x11 tile -> 5x5 pooled pixels x 8 chunks of 8 channels = 200 work items
      if (threadIdx.x < 200) {
        const int pp = threadIdx.x >> 3, ch = threadIdx.x & 7;
        const int pi = pp / 5, pj = pp - pi * 5;
        const uint8_t* bg = stage_gen + (it & 1u) * 16384u;
        __half2 m[4];
        bool first = true;
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
          for (int dj = 0; dj < 3; ++dj) {
            const int rr = (2 * pi + di) * 11 + (2 * pj + dj);
            const uint4 x = *reinterpret_cast<const uint4*>(bg + rr * 128 + ((ch ^ (rr & 7)) << 4));
            const __half2* xh = reinterpret_cast<const __half2*>(&x);
            if (first) { m[0] = xh[0]; m[1] = xh[1]; m[2] = xh[2]; m[3] = xh[3]; first = false; }
            else { m[0] = __hmax2(m[0], xh[0]); m[1] = __hmax2(m[1], xh[1]); m[2] = __hmax2(m[2], xh[2]); m[3] = __hmax2(m[3], xh[3]); }
          }
        __half* o = out + (((long long)roi * 25 + (5 * by + pi)) * 25 + (5 * bx + pj)) * kStemN + ch * 8;
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(m);
      }
    }
    if (MODE == 1 && cur_seg >= 0) flush(cur_seg);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 256u);
}

constexpr size_t kStemSmemBytes = 1024 + kStemStages * kStemTileBytes + 4096 + 2 * 16384 + 8 * (2 * kStemStages + 2 * kStemAccs + 1) + 16 +
                                  kStemN * 4 + kStemN * 8 + 64;

int roi_stem_patches(const VcbRoiDesc& d, const uint8_t* frames, int fh, int fw, const int* rois, void* patches, cudaStream_t st) {
  if (d.num_rois < 0 || d.out_size != 50 || !frames || !rois || !patches || fh <= 0 || fw <= 0 || ((uintptr_t)patches & 15))
    return set_error(VCB_ERR_INVALID, "roi_stem_patches: bad argument (out_size must be 50)");
  if (d.num_rois == 0) return VCB_OK;
  roi_stem_patches_kernel<<<d.num_rois, 256, 0, st>>>(d, frames, fh, fw, rois, reinterpret_cast<uint4*>(patches));
  return check_cuda(cudaGetLastError(), "roi_stem_patches launch");
}

// scale / shift table of the stem's BatchNorm for MODE 2: sums hold the statistics of conv + bias over hw outputs per crop;
// y = gamma * (conv + bias - mean) / sqrt(var + eps) + beta = conv * scale + (beta + (bias - mean) * scale)
__global__ void bn_seg_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ seg_crops, int num_seg1, int c, int hw,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ bias, float eps,
                                       float* __restrict__ affine) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_seg1 * c) return;
  const int seg = i / c, ch = i - seg * c;
  const double cnt = (double)seg_crops[seg] * (double)hw;
  const double mean = cnt > 0 ? sums[(long long)i * 2] / cnt : 0.0;
  double var = cnt > 0 ? sums[(long long)i * 2 + 1] / cnt - mean * mean : 0.0;
  if (var < 0) var = 0;
  const float k = gamma[ch] * (float)(1.0 / sqrt(var + (double)eps));
  affine[(long long)i * 2] = k;
  affine[(long long)i * 2 + 1] = beta[ch] + ((bias ? bias[ch] : 0.0f) - (float)mean) * k;
}

int bn_seg_finalize(const double* sums, const int* seg_crops, int num_seg1, int c, int hw, const float* gamma, const float* beta,
                    const float* bias, float eps, float* affine, cudaStream_t st) {
  if (!sums || !seg_crops || !gamma || !beta || !affine || num_seg1 <= 0 || c <= 0 || hw <= 0)
    return set_error(VCB_ERR_INVALID, "bn_seg_finalize: bad argument");
  const int total = num_seg1 * c;
  bn_seg_finalize_kernel<<<(total + 255) / 256, 256, 0, st>>>(sums, seg_crops, num_seg1, c, hw, gamma, beta, bias, eps, affine);
  return check_cuda(cudaGetLastError(), "bn_seg_finalize launch");
}

template <int MODE>
static int launch_stem(const void* patches, const void* w_packed, const float* bias, void* out, int num_rois, const int* seg_of_crop,
                       double* sums, const float* affine, cudaStream_t st) {
  int rc = require_init();
  if (rc != VCB_OK) return rc;
  if (num_rois < 0 || !patches || !w_packed || ((uintptr_t)patches & 15) || ((uintptr_t)w_packed & 15) || ((uintptr_t)out & 15) ||
      (MODE != 2 && !bias) || (MODE != 1 && !out) || (MODE != 0 && !seg_of_crop) || (MODE == 1 && !sums) || (MODE == 2 && !affine))
    return set_error(VCB_ERR_INVALID, "reid_stem: bad argument");
  if (num_rois == 0) return VCB_OK;
  if ((long long)num_rois * kStemBlocks * kStemRows > 0x7fffff00LL) return set_error(VCB_ERR_INVALID, "reid_stem: too many crops");
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(reid_stem_pool_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStemSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(reid stem)");
    attr_set = true;
  }
  alignas(64) CUtensorMap ta, tw;
  const int num_tiles = num_rois * kStemBlocks;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kStemK, (cuuint64_t)num_tiles * kStemRows};
    const cuuint64_t strides[1] = {(cuuint64_t)kStemK * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kStemK, (cuuint32_t)kStemRows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(patches), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(stem patches) failed: %d", (int)r);
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kStemK, (cuuint64_t)kStemN};
    const cuuint64_t strides[1] = {(cuuint64_t)kStemK * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kStemK, (cuuint32_t)kStemN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(stem weights) failed: %d", (int)r);
  }
  const int max_ctas = state().num_sms * (MODE == 1 ? 1 : 2);
  const int grid = num_tiles < max_ctas ? num_tiles : max_ctas;
  reid_stem_pool_kernel<MODE><<<grid, kStemThreads, kStemSmemBytes, st>>>(ta, tw, bias, reinterpret_cast<__half*>(out), num_tiles,
                                                                         state().fault_dev, seg_of_crop, sums, affine);
  return check_cuda(cudaGetLastError(), "reid_stem launch");
}

int reid_stem_pool(const void* patches, const void* w_packed, const float* bias, void* out, int num_rois, cudaStream_t st) {
  return launch_stem<0>(patches, w_packed, bias, out, num_rois, nullptr, nullptr, nullptr, st);
}
int reid_stem_stats(const void* patches, const void* w_packed, const float* bias, int num_rois, const int* seg_of_crop, double* sums,
                    cudaStream_t st) {
  return launch_stem<1>(patches, w_packed, bias, nullptr, num_rois, seg_of_crop, sums, nullptr, st);
}
int reid_stem_pool_bn(const void* patches, const void* w_packed, const float* affine, const int* seg_of_crop, void* out, int num_rois,
                      cudaStream_t st) {
  return launch_stem<2>(patches, w_packed, nullptr, out, num_rois, seg_of_crop, nullptr, affine, st);
}

}  // namespace vcb
