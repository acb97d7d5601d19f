// K5c -- the fused ReID stem fed straight from the frames: ROI crop + resize + normalise -> Conv3x3(3->64) (+BN) -> ReLU ->
// MaxPool(3, 2, 1) in ONE kernel, with no im2col operand in HBM.
//
// Replaces the pair roi_stem_patches_kernel + reid_stem_pool_kernel (reid_stem.cu), which wrote 776 MB of im2col patches per 4096
// crops and read them back (twice under train-mode BatchNorm) for 450 MB of algorithmic traffic.  Same reference lines:
//   /root/reference/networks/deepsort/deep/feature_extractor.py:26-39  (crop preprocessing)
//   /root/reference/networks/deepsort/deep/model.py:52-60             (stem conv + BN + ReLU + MaxPool2d(3, 2, padding=1))
// Every CTA owns a contiguous range of crops.  Eight producer warps (two groups of four; group g builds the tiles it % 2 == g) keep
// the current crop, resized exactly like roi_resize_norm (cv2 INTER_LINEAR semantics, fp16 rounding of the normalised pixel), in
// shared memory as a zero-bordered [52][52][3] fp16 image and build, per block of 5x5 pooled pixels, the 128 x 32 im2col tile (one
// GEMM row per thread: three runs of nine consecutive halves of the staged crop) directly in the 64-byte-swizzled layout the tensor
// core reads; the NEXT crop is resized into the other buffer one hundred pixels per tile, so its global loads hide behind the row
// building (the two taps of a source row arrive as <= 3 aligned 32-bit words; u8 / 255 is an exact FMA sequence).  Thread 0 of a
// producer group issues the two tcgen05.mma of its tile once its group has arrived.  The conv bias rides in the GEMM: K slots
// 27 / 28 of every row hold 1.0 and the packed weights hold the bias split into an fp16 head and tail there, so the accumulator is
// conv + bias and the epilogue keeps no per-channel registers.  Epilogue (warps 0-7): TMEM -> fp16 -> staged 11x11x64 tile ->
// 3x3/s2 max (pool padding = the window centre re-read) -> ReLU -> 16-byte stores.  MODE 0: BatchNorm folded; MODE 1: statistics
// only, read through the 16x256b accumulator fragment (a thread owns fixed columns: sums stay in registers across tiles); MODE 2:
// per-segment scale / shift before the pool.  Measurements and the ncu findings that shaped it: profiles/r02_stem_direct.md.
#include "vcb_internal.h"
#include "vcb_ptx.cuh"

namespace vcb {

namespace {

constexpr int kS = 50;                        // crop size (feature_extractor.py:18)
constexpr int kPad = 52;                      // staged crop: one zero pixel all round
constexpr int kCropBytes = 16256;             // 52 * 52 * 3 halves = 16224 B, rounded up to a multiple of 128
constexpr int kBlocks = 25;                   // 5 x 5 blocks of 5 x 5 pooled pixels per crop
constexpr int kRows = 128;                    // GEMM rows per block: the 11 x 11 conv outputs a block's pool windows touch (121 used)
constexpr int kK = 32;                        // 27 taps + 2 bias slots + 3 zeros
constexpr int kN = 64;
constexpr int kTileBytes = kRows * kK * 2;    // 8 KiB
constexpr int kStages = 4;
constexpr int kAccs = 4;
constexpr int kEpiThreads = 256;              // warps 0-7
constexpr int kProdThreads = 256;             // warps 8-15: two groups of 128, group g builds (and issues the MMAs of) tiles it % 2 == g
constexpr int kThreads = kEpiThreads + kProdThreads;

// shared-memory map (offsets from the 1024-byte aligned base)
constexpr uint32_t kOffW = kStages * kTileBytes;                 // 32768: weights [64][32] fp16, 64-byte swizzle
constexpr uint32_t kOffStage = kOffW + 4096;                     // 2 x 16 KiB staged conv tiles
constexpr uint32_t kOffCrop = kOffStage + 2 * 16384;             // 2 crops
constexpr uint32_t kOffTab = kOffCrop + 2 * kCropBytes;          // 2 x 100 x {int off0, int off1, float frac, int unused}
constexpr uint32_t kOffImg = kOffTab + 2 * 100 * 16;             // 2 x {const uint8_t* frame, int valid, int unused}
constexpr uint32_t kOffBars = kOffImg + 2 * 16;                  // full[4], empty[4], tfull[4], tempty[4], w_bar
constexpr uint32_t kOffTmemSlot = kOffBars + 8 * (2 * kStages + 2 * kAccs + 1);
constexpr uint32_t kOffAff = (kOffTmemSlot + 16 + 15) & ~15u;    // MODE 2: [64][2] scale, shift of the current segment
constexpr uint32_t kSmemBytes = 1024 + kOffAff + kN * 8 + 64;

struct TabEntry { int off0, off1; float frac; int unused; };
struct ImgEntry { const uint8_t* frame; int valid; int unused; };

}  // namespace

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
reid_stem_direct_kernel(const __grid_constant__ CUtensorMap tmap_w, const VcbRoiDesc d, const uint8_t* __restrict__ frames, int fh, int fw,
                        const int* __restrict__ rois, __half* __restrict__ out, int num_rois, KernelFault* fault,
                        const int* __restrict__ seg_of_crop, double* __restrict__ sums, const float* __restrict__ affine) {
  // crops [c_begin, c_end) of this CTA: the first `rem` CTAs take one more
  const int q_ = num_rois / (int)gridDim.x, rem_ = num_rois - q_ * (int)gridDim.x;
  const int c_begin = (int)blockIdx.x * q_ + min((int)blockIdx.x, rem_);
  const int c_end = c_begin + q_ + ((int)blockIdx.x < rem_ ? 1 : 0);
  const int my_tiles = (c_end - c_begin) * kBlocks;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kOffBars;
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kStages + kAccs + a); };
  const uint32_t w_bar = bars + 8u * (2 * kStages + 2 * kAccs);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemSlot);
  uint8_t* stage_gen = smem_gen + kOffStage;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < kAccs; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiThreads); }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    if (lane == 0) tma_prefetch_desc(&tmap_w);
    tmem_alloc(smem_base + kOffTmemSlot, 256u);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp >= 8) {
    // ---- producers (256 threads = two groups): thread r of group g builds GEMM row r of the tiles it % 2 == g; thread 0 of a group
    // also issues that tile's two tcgen05.mma (M = 128, N = 64, K = 16) once its group has arrived
    const int pt = (int)threadIdx.x - kEpiThreads;
    const int grp = pt >> 7, r_ = pt & 127;
    if (pt == 0) {
      mbar_arrive_expect_tx(w_bar, 4096u);
      tma_load_2d(&tmap_w, w_bar, smem_base + kOffW, 0, 0);
    }
    const uint32_t idesc = umma_idesc_f16((uint32_t)kN);
    const uint64_t desc_hi = umma_desc_kmajor(0, 8u * kK * 2u, 4u);      // SWIZZLE_64B, 8-row groups 512 B apart
    const uint64_t w_desc = desc_hi | (uint64_t)(((smem_base + kOffW) & 0x3FFFF) >> 4);
    auto issue_tile = [&](uint32_t it) {
      const int st = it % kStages;
      const uint32_t acc = it % kAccs;
      mbar_wait_parked(full_bar(st), (it / kStages) & 1u, fault, FAULT_FULL_WAIT, 630 + st, 500u);
      mbar_wait_parked(tempty_bar(acc), ((it / kAccs) & 1u) ^ 1u, fault, FAULT_TMEM_EMPTY_WAIT, 620 + (int)acc);
      tcgen05_fence_after();
      const uint64_t a_desc = desc_hi | (uint64_t)(((smem_base + (uint32_t)st * kTileBytes) & 0x3FFFF) >> 4);
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)kN;
      umma_f16(d_tmem, a_desc, w_desc, idesc, 0u);
      umma_f16(d_tmem, a_desc + 2u, w_desc + 2u, idesc, 1u);
      umma_commit(empty_bar(st));
      umma_commit(tfull_bar(acc));
    };
    const int rr = min(r_, 120);                          // rows 121..127 are never read by the epilogue: any finite content
    const int ti = rr / 11, tj = rr - ti * 11;
    const uint32_t row_smem = (uint32_t)r_ * 64u;
    const uint32_t swz = (uint32_t)((r_ >> 1) & 3);
    const uint32_t crop_smem = smem_base + kOffCrop;
    TabEntry* tabs = reinterpret_cast<TabEntry*>(smem_gen + kOffTab);
    ImgEntry* imgs = reinterpret_cast<ImgEntry*>(smem_gen + kOffImg);
    auto prod_sync = [&]() { asm volatile("bar.sync 3, 256;" ::: "memory"); };

    // resize tables of crop c into table set b: entries 0..49 = columns, 50..99 = rows (the arithmetic of roi_resize_norm_kernel)
    auto build_tables = [&](int c, int b) {
      if (pt < 100) {
        const int f = __ldg(rois + c * 5 + 0), x1 = __ldg(rois + c * 5 + 1), y1 = __ldg(rois + c * 5 + 2), x2 = __ldg(rois + c * 5 + 3),
                  y2 = __ldg(rois + c * 5 + 4);
        const int cw = x2 - x1, chh = y2 - y1;
        const bool ok = !(cw <= 0 || chh <= 0 || x1 < 0 || y1 < 0 || x2 > fw || y2 > fh || f < 0 || (d.num_frames > 0 && f >= d.num_frames));
        const bool is_y = pt >= 50;
        const int o = is_y ? pt - 50 : pt;
        const int len = is_y ? chh : cw, org = is_y ? y1 : x1;
        const double sc = (double)len / (double)kS;
        const double fd = ((double)o + 0.5) * sc - 0.5;
        int i0 = (int)floor(fd);
        float fr = (float)(fd - (double)i0);
        if (i0 < 0) { i0 = 0; fr = 0.f; }
        int i1 = i0 + 1;
        if (i0 >= len - 1) { i0 = len - 1; i1 = len - 1; fr = 0.f; }
        const int unit = is_y ? fw * 3 : 3;
        TabEntry e;
        e.off0 = ok ? (org + i0) * unit : 0; e.off1 = ok ? (org + i1) * unit : 0; e.frac = fr; e.unused = 0;
        tabs[b * 100 + pt] = e;
        if (pt == 0) {
          ImgEntry ie;
          ie.frame = frames + (long long)(ok ? f : 0) * fh * fw * 3; ie.valid = ok ? 1 : 0; ie.unused = 0;
          imgs[b] = ie;
        }
      }
    };
    // one resized, normalised pixel (oy, ox) of the crop described by table set b -> crop buffer b.  Split in two so that the global
    // loads are in flight while the caller builds its im2col row.  The two taps of a source row are 6 contiguous bytes (3 when the
    // column is clamped): they arrive as up to three ALIGNED 32-bit words (a word is loaded only if it holds a needed byte, so every
    // load stays inside the frame) instead of six byte gathers -- the byte loads' sector lookups were saturating L1.
    // u8 -> float / 255 exactly as cv2 (im.astype(float32) / 255.): q0 = i * r, q = fma(fma(-q0, 255, i), r, q0) with r = RN(1 / 255)
    // is the correctly rounded quotient for every i in 0..255 (checked exhaustively).
    struct PixLoad { uint32_t t0, t1, t2, b0, b1, b2, sft; float fx, fy; bool same_x; };
    auto row_load = [&](const uint8_t* pa, bool same_x, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
      const uintptr_t u = reinterpret_cast<uintptr_t>(pa);
      const uint32_t o = (uint32_t)(u & 3u), nb = same_x ? 3u : 6u;
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
      w0 = __ldg(wp);
      w1 = (o + nb > 4u) ? __ldg(wp + 1) : 0u;
      w2 = (o + nb > 8u) ? __ldg(wp + 2) : 0u;
    };
    auto pixel_load = [&](int oy, int ox, int b, const uint8_t* frame, PixLoad& L_) {
      const TabEntry ex = tabs[b * 100 + ox], ey = tabs[b * 100 + 50 + oy];
      L_.fx = ex.frac; L_.fy = ey.frac; L_.same_x = ex.off1 == ex.off0;
      const uint8_t* pt_ = frame + ey.off0 + ex.off0;
      const uint8_t* pb_ = frame + ey.off1 + ex.off0;
      // both rows start at the same offset modulo 4 only if the row pitch is a multiple of 4: keep the shift per row
      L_.sft = (uint32_t)(reinterpret_cast<uintptr_t>(pt_) & 3u) | ((uint32_t)(reinterpret_cast<uintptr_t>(pb_) & 3u) << 8);
      row_load(pt_, L_.same_x, L_.t0, L_.t1, L_.t2);
      row_load(pb_, L_.same_x, L_.b0, L_.b1, L_.b2);
    };
    auto unpack_row = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t o, bool same_x, uint32_t& A, uint32_t& B) {
      const uint32_t s_ = o * 8u;
      const uint32_t x0 = __funnelshift_r(w0, w1, s_), x1 = __funnelshift_r(w1, w2, s_);
      A = x0;                                                  // bytes 0..2 = the left tap
      B = same_x ? x0 : ((x0 >> 24) | (x1 << 8));              // bytes 0..2 = the right tap
    };
    auto u8_over_255 = [&](uint32_t packed, int c) {
      const float i = __uint_as_float(__byte_perm(packed, 0x4B000000u, 0x7650u | (uint32_t)c)) - 8388608.0f;
      const float r = __uint_as_float(0x3b808081u);
      const float q0 = i * r;
      return fmaf(fmaf(-q0, 255.0f, i), r, q0);
    };
    auto pixel_store = [&](int oy, int ox, int b, bool ok, const PixLoad& L_) {
      __half* dst = reinterpret_cast<__half*>(smem_gen + kOffCrop + b * kCropBytes) + ((oy + 1) * kPad + (ox + 1)) * 3;
      const float fx = L_.fx, fy = L_.fy;
      uint32_t p00, p01, p10, p11;
      unpack_row(L_.t0, L_.t1, L_.t2, L_.sft & 3u, L_.same_x, p00, p01);
      unpack_row(L_.b0, L_.b1, L_.b2, (L_.sft >> 8) & 3u, L_.same_x, p10, p11);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = u8_over_255(p00, c), b2 = u8_over_255(p01, c);
        const float e = u8_over_255(p10, c), g = u8_over_255(p11, c);
        const float top = a * (1.0f - fx) + b2 * fx;
        const float bot = e * (1.0f - fx) + g * fx;
        const float val = top * (1.0f - fy) + bot * fy;
        dst[c] = ok ? __float2half_rn((val - d.mean[c]) * d.inv_std[c]) : __float2half_rn(0.f);
      }
    };

    if (my_tiles > 0) {
      // both crop buffers start as zeros: the one-pixel border is never written again
      for (int i = pt; i < 2 * kCropBytes / 16; i += kProdThreads)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(crop_smem + (uint32_t)i * 16u), "r"(0u) : "memory");
      build_tables(c_begin, 0);
      prod_sync();
      {
        const ImgEntry ie = imgs[0];
        for (int px = pt; px < kS * kS; px += kProdThreads) {
          const int oy = px / kS, ox = px - oy * kS;
          PixLoad pl;
          pixel_load(oy, ox, 0, ie.frame, pl);
          pixel_store(oy, ox, 0, ie.valid != 0, pl);
        }
      }
      prod_sync();
      if (r_ == 0) mbar_wait(w_bar, 0u, fault, FAULT_FULL_WAIT, 610);
      // pixels of the NEXT crop this thread resizes while the tiles of the current one are built: tile blk -> rows 2*blk, 2*blk+1
      const bool px_thread = r_ < 100;
      const int ox_n = r_ < 50 ? r_ : r_ - 50, oy_n = r_ < 50 ? 0 : 1;
      uint32_t it = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int cur = (c - c_begin) & 1;
        const bool has_next = c + 1 < c_end;
        if (has_next) build_tables(c + 1, cur ^ 1);
        prod_sync();
        const ImgEntry ie = imgs[cur ^ 1];                     // garbage when !has_next: never used then
        const bool do_px = has_next && px_thread;
        const uint32_t crop_cur = crop_smem + (uint32_t)cur * kCropBytes;
        int by = 0, bx = 0;
        for (int blk = 0; blk < kBlocks; ++blk, ++it) {
          if ((int)(it & 1u) != grp) {                        // the other group's tile
            if (++bx == 5) { bx = 0; ++by; }
            continue;
          }
          PixLoad pl;
          if (do_px) pixel_load(2 * blk + oy_n, ox_n, cur ^ 1, ie.frame, pl);
          // ---- im2col row: conv output (cy, cx) = (10*by - 1 + ti, 10*bx - 1 + tj); output row / column -1 is pool padding (never
          // read by the pool), addressed as 0 to stay inside the buffer
          const int cyc = max(10 * by - 1 + ti, 0), cxc = max(10 * bx - 1 + tj, 0);
          const int h0 = (cyc * kPad + cxc) * 3;                 // first half of tap row 0 in the bordered crop
          const uint32_t sh = (uint32_t)(h0 & 1) * 16u;          // same parity for the three tap rows (52 is even)
          const uint32_t a0 = crop_cur + (uint32_t)(h0 >> 1) * 4u;
          uint32_t t[3][5];
#pragma unroll
          for (int r3 = 0; r3 < 3; ++r3) {
            uint32_t w[5];
#pragma unroll
            for (int i = 0; i < 5; ++i)
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"(a0 + (uint32_t)(r3 * (kPad * 3 * 2) + i * 4)));
#pragma unroll
            for (int i = 0; i < 4; ++i) t[r3][i] = __funnelshift_r(w[i], w[i + 1], sh);
            t[r3][4] = w[4] >> sh;
          }
          // 27 halves: A0..A8 | B0..B8 | C0..C8, then 1.0, 1.0 (bias slots), zeros
          uint32_t o[16];
          o[0] = t[0][0]; o[1] = t[0][1]; o[2] = t[0][2]; o[3] = t[0][3];
          o[4] = (t[0][4] & 0xffffu) | (t[1][0] << 16);
          o[5] = __funnelshift_r(t[1][0], t[1][1], 16);
          o[6] = __funnelshift_r(t[1][1], t[1][2], 16);
          o[7] = __funnelshift_r(t[1][2], t[1][3], 16);
          o[8] = __funnelshift_r(t[1][3], t[1][4], 16);
          o[9] = t[2][0]; o[10] = t[2][1]; o[11] = t[2][2]; o[12] = t[2][3];
          o[13] = (t[2][4] & 0xffffu) | 0x3C000000u;
          o[14] = 0x00003C00u; o[15] = 0u;
          const int st = it % kStages;
          mbar_wait_parked(empty_bar(st), ((it / kStages) & 1u) ^ 1u, fault, FAULT_EMPTY_WAIT, 600 + st, 500u);
          const uint32_t dst = smem_base + (uint32_t)st * kTileBytes + row_smem;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((((uint32_t)j) ^ swz) << 4)), "r"(o[4 * j]), "r"(o[4 * j + 1]),
                         "r"(o[4 * j + 2]), "r"(o[4 * j + 3]) : "memory");
          fence_proxy_async_smem();
          mbar_arrive(full_bar(st));
          if (do_px) pixel_store(2 * blk + oy_n, ox_n, cur ^ 1, ie.valid != 0, pl);
          if (r_ == 0) issue_tile(it);
          if (++bx == 5) { bx = 0; ++by; }
        }
        prod_sync();        // the next crop is complete and nobody reads this one any more
      }
    }
  } else if (MODE == 1) {
    // ---- statistics epilogue (256 threads): per-(segment, channel) sum / sum of squares of conv + bias over every conv output of a
    // crop exactly once = rows 1..10 x columns 1..10 of each 11x11 tile (row 0 / column 0 belong to the neighbouring block, or are
    // padding).  The accumulator is read in the 16x256b fragment shape: a thread owns 4 FIXED rows and 8 FIXED columns of every tile,
    // so the sums stay in 16 registers across the CTA's tiles and meet the other rows only when a segment ends.
    const int q = warp & 3, half = warp >> 2;
    float vf[4];                                               // 1.0 for the rows that count
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = q * 32 + (k >> 1) * 16 + (lane >> 2) + (k & 1) * 8;
      const int ti = row / 11, tj = row - ti * 11;
      vf[k] = (row < 121 && ti >= 1 && tj >= 1) ? 1.0f : 0.0f;
    }
    float s1[8], s2[8];                                        // columns half*32 + 8j + 2(lane%4) + e  ->  index 2j + e
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    int cur_seg = -1;
    auto flush = [&](int seg) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) {
          s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], m);
          s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], m);
        }
      }
      if (lane < 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ch = half * 32 + 8 * (i >> 1) + 2 * lane + (i & 1);
          atomicAdd(sums + ((long long)seg * kN + ch) * 2 + 0, (double)s1[i]);
          atomicAdd(sums + ((long long)seg * kN + ch) * 2 + 1, (double)s2[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    };
    int roi = c_begin, blk = 0;
    for (uint32_t it = 0; it < (uint32_t)my_tiles; ++it) {
      const uint32_t acc = it % kAccs;
      if (blk == 0) {
        const int seg = __ldg(seg_of_crop + roi);
        if (seg != cur_seg) {
          if (cur_seg >= 0) flush(cur_seg);
          cur_seg = seg;
        }
      }
      if (lane == 0) mbar_wait_parked(tfull_bar(acc), (it / kAccs) & 1u, fault, FAULT_TMEM_FULL_WAIT, 640 + (int)acc);   // one poller per warp: 256 spinning
      __syncwarp();                                                                                                // threads saturate the shared-memory pipe
      tcgen05_fence_after();
      const uint32_t t_addr = tmem_base + acc * (uint32_t)kN + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16);
      uint32_t v0[16], v1[16];
      tmem_ld_16x256b_x4(t_addr, v0);
      tmem_ld_16x256b_x4(t_addr + (16u << 16), v1);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float a0 = __uint_as_float(v0[4 * j + e]) * vf[0], a1 = __uint_as_float(v0[4 * j + 2 + e]) * vf[1];
          const float a2 = __uint_as_float(v1[4 * j + e]) * vf[2], a3 = __uint_as_float(v1[4 * j + 2 + e]) * vf[3];
          s1[2 * j + e] += (a0 + a1) + (a2 + a3);
          s2[2 * j + e] = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, s2[2 * j + e]))));
        }
      if (++blk == kBlocks) { blk = 0; ++roi; }
    }
    if (cur_seg >= 0) flush(cur_seg);
  } else {
    // ---- epilogue (256 threads): conv + bias (MODE 2: per-segment scale / shift) -> fp16 -> staged 11x11x64 tile -> 3x3/s2 max ->
    // ReLU (after the max: both are monotone) -> 16-byte stores of the 5x5x64 pooled block
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                       // GEMM row = position ti*11 + tj of the 11x11 conv tile
    const int sw = row & 7;
    const uint32_t stage0 = smem_base + kOffStage;
    float* aff_s = reinterpret_cast<float*>(smem_gen + kOffAff);
    // pooling work item of this thread: pooled pixel (pi, pj) of the block, channels 8*ch .. 8*ch+7
    const int pp = threadIdx.x >> 3, ch = threadIdx.x & 7;
    const int pi = pp / 5, pj = pp - pi * 5;
    int cur_seg = -1;
    int roi = c_begin, blk = 0, by = 0, bx = 0;
    for (uint32_t it = 0; it < (uint32_t)my_tiles; ++it) {
      const uint32_t acc = it % kAccs;
      if (MODE == 2 && blk == 0) {
        const int seg = __ldg(seg_of_crop + roi);
        if (seg != cur_seg) {                              // CTA-uniform: every thread sees the same tile sequence
          asm volatile("bar.sync 2, 256;" ::: "memory");      // everyone is done with the previous segment's table
          if (threadIdx.x < 128) aff_s[threadIdx.x] = __ldg(affine + (long long)seg * 128 + threadIdx.x);
          asm volatile("bar.sync 2, 256;" ::: "memory");
          cur_seg = seg;
        }
      }
      if (lane == 0) mbar_wait_parked(tfull_bar(acc), (it / kAccs) & 1u, fault, FAULT_TMEM_FULL_WAIT, 640 + (int)acc);   // one poller per warp: 256 spinning
      __syncwarp();                                                                                                // threads saturate the shared-memory pipe
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + acc * (uint32_t)kN + (uint32_t)(half * 32) + ((uint32_t)(q * 32) << 16);
      uint32_t v0[16], v1[16];
      tmem_ld_x16(t_row, v0);
      tmem_ld_x16(t_row + 16u, v1);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
      uint32_t h2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t* v = i < 8 ? v0 : v1;
        float a = __uint_as_float(v[2 * (i & 7)]), b = __uint_as_float(v[2 * (i & 7) + 1]);
        if (MODE == 2) {
          const int c0 = (i < 8 ? 0 : 16) + 2 * (i & 7);
          const float4 k = *reinterpret_cast<const float4*>(aff_s + (half * 32 + c0) * 2);     // scale, shift, scale, shift (broadcast)
          a = fmaf(a, k.x, k.y);
          b = fmaf(b, k.z, k.w);
        }
        const __half2 t = __floats2half2_rn(a, b);
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
      // 3x3/s2 max-pool of the 11x11 tile -> 5x5 pooled pixels x 8 chunks of 8 channels = 200 work items.  Tile row 0 (column 0) of a
      // block on the top (left) edge is conv output -1 = the pool's own padding: skipped; rows >= 121 are never touched.
      if (threadIdx.x < 200) {
        const uint8_t* bg = stage_gen + (it & 1u) * 16384u;
        const bool skip_r = by == 0 && pi == 0, skip_c = bx == 0 && pj == 0;
        __half2 m[4];
        {
          const int r9 = (2 * pi + 1) * 11 + (2 * pj + 1);
          const uint4 x = *reinterpret_cast<const uint4*>(bg + r9 * 128 + ((ch ^ (r9 & 7)) << 4));
          const __half2* xh = reinterpret_cast<const __half2*>(&x);
          m[0] = xh[0]; m[1] = xh[1]; m[2] = xh[2]; m[3] = xh[3];
        }
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
          for (int dj = 0; dj < 3; ++dj) {
            if (di == 1 && dj == 1) continue;
            // a skipped tap re-reads the window centre's row / column instead (branch-free; a duplicate cannot change a max)
            const int r9 = (2 * pi + ((di == 0 && skip_r) ? 1 : di)) * 11 + (2 * pj + ((dj == 0 && skip_c) ? 1 : dj));
            const uint4 x = *reinterpret_cast<const uint4*>(bg + r9 * 128 + ((ch ^ (r9 & 7)) << 4));
            const __half2* xh = reinterpret_cast<const __half2*>(&x);
            m[0] = __hmax2(m[0], xh[0]); m[1] = __hmax2(m[1], xh[1]); m[2] = __hmax2(m[2], xh[2]); m[3] = __hmax2(m[3], xh[3]);
          }
        const __half2 z = __float2half2_rn(0.f);
        m[0] = __hmax2(m[0], z); m[1] = __hmax2(m[1], z); m[2] = __hmax2(m[2], z); m[3] = __hmax2(m[3], z);
        __half* o = out + (((long long)roi * 25 + (5 * by + pi)) * 25 + (5 * bx + pj)) * kN + ch * 8;
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(m);
      }
      if (++bx == 5) { bx = 0; ++by; }
      if (++blk == kBlocks) { blk = 0; by = 0; bx = 0; ++roi; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256u);
}

template <int MODE>
static int launch_stem_direct(const VcbRoiDesc& d, const uint8_t* frames, int fh, int fw, const int* rois, const void* w_packed, void* out,
                              const int* seg_of_crop, double* sums, const float* affine, cudaStream_t st) {
  int rc = require_init();
  if (rc != VCB_OK) return rc;
  if (d.num_rois < 0 || d.out_size != kS || !frames || !rois || !w_packed || fh <= 0 || fw <= 0 || ((uintptr_t)w_packed & 15) ||
      ((uintptr_t)out & 15) || (MODE != 1 && !out) || (MODE != 0 && !seg_of_crop) || (MODE == 1 && !sums) || (MODE == 2 && !affine))
    return set_error(VCB_ERR_INVALID, "reid_stem_direct: bad argument (out_size must be 50)");
  if ((long long)fh * fw * 3 > 0x7fffffffLL) return set_error(VCB_ERR_INVALID, "reid_stem_direct: frame too large");
  if (d.num_rois == 0) return VCB_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(reid_stem_direct_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(reid stem direct)");
    attr_set[dev] = true;
  }
  alignas(64) CUtensorMap tw;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)kN};
    const cuuint64_t strides[1] = {(cuuint64_t)kK * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kK, (cuuint32_t)kN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(stem weights) failed: %d", (int)r);
  }
  const int max_ctas = state().num_sms * 2;
  const int grid = d.num_rois < max_ctas ? d.num_rois : max_ctas;
  reid_stem_direct_kernel<MODE><<<grid, kThreads, kSmemBytes, st>>>(tw, d, frames, fh, fw, rois, reinterpret_cast<__half*>(out), d.num_rois,
                                                                    state().fault_dev, seg_of_crop, sums, affine);
  return check_cuda(cudaGetLastError(), "reid_stem_direct launch");
}

int reid_stem_direct(const VcbRoiDesc& d, const uint8_t* frames, int fh, int fw, const int* rois, const void* w_packed, void* out,
                     cudaStream_t st) {
  return launch_stem_direct<0>(d, frames, fh, fw, rois, w_packed, out, nullptr, nullptr, nullptr, st);
}
int reid_stem_direct_stats(const VcbRoiDesc& d, const uint8_t* frames, int fh, int fw, const int* rois, const void* w_packed,
                           const int* seg_of_crop, double* sums, cudaStream_t st) {
  return launch_stem_direct<1>(d, frames, fh, fw, rois, w_packed, nullptr, seg_of_crop, sums, nullptr, st);
}
int reid_stem_direct_bn(const VcbRoiDesc& d, const uint8_t* frames, int fh, int fw, const int* rois, const void* w_packed, const float* affine,
                        const int* seg_of_crop, void* out, cudaStream_t st) {
  return launch_stem_direct<2>(d, frames, fh, fw, rois, w_packed, out, seg_of_crop, nullptr, affine, st);
}

}  // namespace vcb
