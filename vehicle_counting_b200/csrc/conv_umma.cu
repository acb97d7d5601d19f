// K1 -- implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, fp16 in / fp32 TMEM
// accumulate) with bias + SiLU/ReLU + residual fused into the epilogue.
//
// Replaces the cuDNN convolutions the reference reaches through torch.nn.Conv2d:
//   YOLOv5 Conv = Conv2d + BatchNorm2d(folded) + SiLU   (/root/reference/networks/yolo.py:70 ->
//                 [upstream models/common.py Conv])
//   ReID  conv/BN/ReLU/residual blocks                  (/root/reference/networks/deepsort/deep/model.py:5-37)
//
// GEMM view: D[M = n*ho*wo, N = cout] = A[M, K = kh*kw*cin] * B[N, K]^T.  A is never materialised:
// one K-step is one filter tap (r, s) x 64 input channels, i.e. a 128-pixel x 128-byte tile that is
// fetched straight from the NHWC activation tensor.  Three A producers share the rest of the kernel:
//   A_TMA    one im2col-mode TMA per K-step (cp.async.bulk.tensor.4d...im2col), zero fill by the TMA unit
//   A_GATHER 128 producer threads issue 16-byte cp.async with zero-fill predicates (any geometry with
//            cin % 8 == 0; also the bring-up / cross-check path for A_TMA)
//   A_C4     same, 8-byte granules over a 4-channel (RGBx) input: the 6x6/s2 YOLO stem and the 3x3 ReID
//            stem, where one K-step is 16 taps x 4 channels
// B (weights, packed [cout_pad][K_pad] fp16, K-major) always arrives by tiled TMA.  Both operands land
// in the 128-byte-swizzled K-major layout the UMMA shared-memory descriptor expects.
//
// CTA = warps 0-7 epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2), warp 8 TMA producer,
// warp 9 MMA issuer (+ TMEM allocator), warps 10-13 gather producers (A_GATHER / A_C4 only).
// Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ...; the fp32 accumulator is double-buffered
// in TMEM when it fits so the epilogue of tile i overlaps the main loop of tile i+1.  The epilogue converts 128 x 64
// (fp16) / 128 x 32 (fp32) sub-tiles into a 128-byte-swizzled staging buffer and hands them to the TMA
// store unit (coalesced, asynchronous, clipped to the tensor bounds); residual tiles arrive the same way.
//
// Kernels in this file (host side: conv_geometry picks one per layer; VcbConvDesc.reserved[3] forces one):
//   conv_umma_kernel<A_MODE, BK, false>  128-pixel tiles, two persistent CTAs per SM (default)
//   conv_umma_2cta_kernel<BK>            cta_group::2 CTA pairs (M = 256 over two SMs, each CTA loads half of every weight tile),
//                                        two co-resident clusters per SM pair: 3x3 layers with N >= 128
//   conv_patch_kernel                    3x3/s1/p1, N <= 64: one input patch per 64-channel chunk in shared memory, the nine taps
//                                        through shifted UMMA descriptors, one MMA-issuing warp per accumulator
//   conv_umma_kernel<A_TMA, BK, true>    256-pixel tiles in two accumulators (opt-in experiment)
// Epilogues: conv_epilogue_fast<ACT, RES, F32OUT, M256, TWO_CTA, PATCH> (specialised, all TMA modes) and conv_epilogue
// (generic, runtime flags: gather / debug modes).  Measured facts that shaped the design are in DESIGN.md section 3.
#include "vcb_internal.h"
#include "vcb_ptx.cuh"

namespace vcb {

// A/B switch (-DVCB_EPI_PARK=1): epilogue warps wait for an accumulator with ONE polling lane per warp and a suspend-time hint on the
// try_wait (the hardware parks the thread) instead of 256 spinning threads.  Measured on B200 (tools/r2_call18.sh): YOLOv5m B=64
// 5.58 -> 5.64 ms, ReID 3.82 -> 3.91 ms -- the spinning warps do not take issue slots anyone needs, and the parked ones wake late.  Off.
#ifndef VCB_EPI_PARK
#define VCB_EPI_PARK 0
#endif
__device__ __forceinline__ void epi_wait_full(uint32_t bar, uint32_t parity, KernelFault* f, int info) {
#if VCB_EPI_PARK
  if ((threadIdx.x & 31) == 0) mbar_wait_parked(bar, parity, f, FAULT_TMEM_FULL_WAIT, info, 1000u);
  __syncwarp();
#else
  mbar_wait(bar, parity, f, FAULT_TMEM_FULL_WAIT, info);
#endif
}

enum { A_TMA = 0, A_GATHER = 1, A_C4 = 2 };

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // fp16 elements = 128 bytes = one swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kGatherLag = 2;                     // cp.async groups in flight per gather thread
constexpr int kNumEpilogueThreads = 256;
constexpr int kEpilogueWarps = 8;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kGatherWarp0 = 10;
constexpr int kGatherThreads = 128;
constexpr int kStageOutBytes = kBlockM * 128;        // one staged output sub-tile (128 rows x 128 B)
constexpr int kThreadsTma = (kMmaWarp + 1) * 32;     // 320
constexpr int kThreadsGather = kThreadsTma + kGatherThreads;   // 448

struct ConvParams {
  // geometry
  int N, H, W, Cin, cin_pitch;
  int P, Q, PQ, M;
  int kh, kw, stride, pad;
  int Cout;            // logical output channels
  int cout_store;      // channels actually written (Cout rounded up to 8, <= out_pitch)
  int out_pitch;
  int chunks_per_tap;  // ceil(Cin / bk)           (A_TMA, A_GATHER)
  int num_k_iters;     // pipeline stages consumed per tile = ceil(total_chunks / chunks_per_stage)
  int bk;              // K elements per chunk: 64 (128B swizzle), 32 (64B), 16 (32B); gather modes use 64
  int chunks_per_stage;   // 64 / bk: a stage always carries up to 64 K elements
  int total_chunks;    // taps * chunks_per_tap
  int block_n, n_tiles, m_tiles, num_tiles;
  int num_pair_tiles;  // cta_group::2 kernel: ceil(m_tiles / 2) * n_tiles
  int num_stages, acc_stages, tmem_cols;
  int act, res_mode, res_pitch, out_fp32;
  int out_stage_bufs;  // 1 or 2 staged sub-tiles in flight
  int out_stage_bytes; // output staging area: 8 warps x 2 boxes x (2 KiB fp16 | 4 KiB fp32)
  int epi_direct;      // 1: per-thread 16-byte global stores (debug / cross-check); 0: staged TMA store
  int c4_wide;         // A_C4: 16-byte granules (two taps) instead of 8-byte ones
  int dbg_skip_epilogue;   // timing experiment: epilogue only hands the accumulator back (results are garbage)
  int dbg_swap;            // timing experiment (256-row tiles): ONE tcgen05.mma with the weights as A (M=128) and 256 pixels as N
  int b_resident;      // A_TMA single-CTA, one N tile, small weights: the whole packed B stays in shared memory
  int b_res_bytes;     // bytes of the resident weight region (multiple of 1024)
  int split_b;         // A_TMA single-CTA: weight tiles are issued by a second producer warp
  int a_tiled;         // A_TMA, 1x1/s1/p0: A is the plain [M][C] matrix -> tiled-mode TMA instead of im2col mode
  int epi_kind;        // 0: generic epilogue; 1..8: specialised instance (epi_kind_of)
  // patch mode (3x3/s1/p1): an M tile is R output rows x Xs output columns of one image; its input patch of (R+2) x (Xs+2)
  // pixels sits in shared memory once per 64-channel chunk and the nine taps read it through shifted descriptors
  int patch_R, patch_Xs, patch_Lp, patch_xsegs, patch_ytiles;
  int patch_box_bytes;   // bytes one patch TMA box delivers = (R+2) * Lp * 128
  int patch_stage_bytes; // shared memory per patch stage (>= (128 + 2*Lp + 2) rows, multiple of 1024)
  int patch_out_bytes;   // bytes of one output / residual box = R * Xs * 128
  int a_stages, b_stages;
  int kchains;         // 1, 2 or 4 accumulators per tile: K steps are dealt round-robin to independent accumulation chains
                       // (dependent tcgen05.mma on ONE accumulator issue ~200 clk apart), the epilogue sums them
  int cout_pad;        // n_tiles * block_n = length of the packed bias
  int epi_split;       // split epilogue (two independent four-warp groups, conv_epilogue_split)
  int epi_empty_count; // arrivals that free an accumulator stage: epilogue threads that read each tile (x2 for CTA pairs)
  int a_rowwin;        // VCB_A_ROWWIN: one tiled 4-D TMA box per filter ROW over the W-padded input (patch_R x patch_Xs = 128 pixels)
  int ksteps_lim;      // K = 16 steps issued per 64-element chunk (3 in row-window mode: the 4th pixel has zero weights)
  int tile_rev;        // walk the tiles from the last to the first: a layer that starts where its producer stopped finds the
                       // most recently written part of its input still in L2 (the engine alternates the direction per layer)
  unsigned long long a_policy, b_policy;   // L2 cache policies of the activation / weight TMA loads (kL2Evict*)
  // train-mode BatchNorm statistics from the epilogue (vcb_conv2d_fwd_stats): per-(segment, channel) sum / sum of squares of the
  // STORED (fp16-rounded) outputs, accumulated into stat_sums[seg][Cout][2]; image n of the batch belongs to segment stat_seg[n]
  const int* stat_seg;
  double* stat_sums;
  const __half* x;
  const float* bias;
  const __half* residual;
  void* out;
  KernelFault* fault;
  unsigned long long* prof;   // optional role timers (vcb_set_option("prof", 1)): cycles summed over CTAs, see PROF_* below
};

// role timers (development aid, off unless p.prof != nullptr): where each warp role of the conv kernel waits
enum { PROF_CTA_TOTAL = 0, PROF_SETUP, PROF_PROD_WAIT_EMPTY, PROF_PROD_TOTAL, PROF_MMA_WAIT_FULL, PROF_MMA_WAIT_TMEM, PROF_MMA_TOTAL,
       PROF_EPI_WAIT_TMEM, PROF_EPI_SYNC_STORE, PROF_EPI_TOTAL, PROF_CTAS, PROF_TILES, PROF_N };
__device__ __forceinline__ long long prof_clock(const unsigned long long* prof) { return prof ? clock64() : 0; }
__device__ __forceinline__ void prof_add(unsigned long long* prof, int slot, long long v) {
  if (prof) atomicAdd(prof + slot, (unsigned long long)v);
}

struct RowInfo {     // one output pixel of the current M tile (gather modes)
  int img_base;      // n * H * W
  short h0, w0;      // top-left input coordinate of the receptive field (may be negative)
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VCB_ACT_SILU) return __fdividef(v, 1.0f + __expf(-v));   // ex2.approx + rcp.approx: ~1e-6 relative
  if (act == VCB_ACT_SILU_TANH) {      // x*sigmoid(x) = h + h*tanh(h), h = x/2: ONE MUFU op per element instead of two
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  if (act == VCB_ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogue (warps 0-7): TMEM -> registers -> bias/activation/residual -> swizzled smem sub-tiles -> TMA store.
// kTwoCta: the CTA is one half of a cta_group::2 pair; it drains its own 128 accumulator rows of the 256-row pair
// tile and releases the accumulator on the LEADER's tmem_empty barrier (remote arrive for the peer CTA).
// ---------------------------------------------------------------------------------------------------------------
template <bool kTwoCta>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, const CUtensorMap* tmap_out_ptr, uint32_t tmem_base,
                                              uint32_t out_stage, uint32_t tmem_full0, uint32_t tmem_empty0, int first_tile,
                                              int tile_step, int cta_rank) {
  // 8 warps: warp (q = warp & 3, half = warp >> 2) reads accumulator rows [32q, 32q+32) and one half of the columns of
  // each staged sub-tile (128 rows x 128 bytes); one TMA store per sub-tile, two sub-tiles in flight.  (A variant with
  // one small TMA store per warp and no CTA-wide barrier measured 20-50 % slower: few large TMA boxes win.)
  const CUtensorMap& tmap_out = *tmap_out_ptr;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto tmem_full_bar = [&](int a) { return tmem_full0 + 8u * (uint32_t)a; };
  auto tmem_empty_bar = [&](int a) { return tmem_empty0 + 8u * (uint32_t)a; };
    // ======================= epilogue: TMEM -> registers -> (smem -> TMA store | global) =======================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int half = warp >> 2;                  // which half of a sub-tile's columns
    const int row_in_tile = q * 32 + lane;
    const int sub_cols = p.out_fp32 ? 32 : 64;   // columns per staged sub-tile (128 bytes per row)
    const int my_cols = sub_cols >> 1;           // 32 (fp16) or 16 (fp32) columns per thread per sub-tile
    const int num_sub = (p.block_n + sub_cols - 1) / sub_cols;
    const bool issuer = threadIdx.x == 0;
    uint32_t tile_iter = 0, sub_count = 0;
    for (int tile_seq = first_tile; tile_seq < (kTwoCta ? p.num_pair_tiles : p.num_tiles); tile_seq += tile_step, ++tile_iter) {
      const int tile = p.tile_rev ? (kTwoCta ? p.num_pair_tiles : p.num_tiles) - 1 - tile_seq : tile_seq;
      const int m_tile = kTwoCta ? 2 * (tile / p.n_tiles) + cta_rank : tile / p.n_tiles;
      const int n_tile = tile % p.n_tiles;
      const uint32_t acc = (p.acc_stages == 2) ? (tile_iter & 1u) : 0u;
      const uint32_t acc_ph = (p.acc_stages == 2) ? ((tile_iter >> 1) & 1u) : (tile_iter & 1u);
      mbar_wait(tmem_full_bar(acc), acc_ph, p.fault, FAULT_TMEM_FULL_WAIT, (int)acc);
      tcgen05_fence_after();
      if (p.dbg_skip_epilogue) {
        tcgen05_fence_before();
        if (kTwoCta && cta_rank != 0) mbar_arrive_remote(tmem_empty_bar(acc), 0);
        else mbar_arrive(tmem_empty_bar(acc));
        continue;
      }
      const int row = m_tile * kBlockM + row_in_tile;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + acc * (uint32_t)p.block_n + ((uint32_t)(q * 32) << 16);
      const size_t out_row = (size_t)row * (size_t)p.out_pitch;
      const size_t res_row = (size_t)row * (size_t)p.res_pitch;
      for (int sub = 0; sub < num_sub; ++sub, ++sub_count) {
        const uint32_t stage_buf = out_stage + ((p.out_stage_bufs == 2) ? (sub_count & 1u) : 0u) * kStageOutBytes;
        if (!p.epi_direct && p.out_stage_bufs == 1) {          // single staging buffer: drain the previous store first
          if (issuer) tma_store_wait_read<0>();
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        const int col_base = sub * sub_cols + half * my_cols;            // first column (within the N tile) of this thread
        const int ngroups = (col_base >= p.block_n) ? 0 : ((my_cols == 32 && col_base + 16 < p.block_n) ? 2 : 1);   // warp-uniform
        // both TMEM loads first, bias (and residual) loads while they are in flight, one wait
        uint32_t v[2][16];
        if (ngroups > 0) tmem_ld_x16(t_row + (uint32_t)col_base, v[0]);
        if (ngroups > 1) tmem_ld_x16(t_row + (uint32_t)(col_base + 16), v[1]);
        const int n_first = n_tile * p.block_n + col_base;
        if (ngroups > 0) tmem_ld_wait();
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {                               // up to two 16-column groups
          if (g16 >= ngroups) continue;
          const int n0 = n_first + g16 * 16;
          float fg[16];
          {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(bp + i);
              fg[4 * i + 0] = __uint_as_float(v[g16][4 * i + 0]) + b4.x;
              fg[4 * i + 1] = __uint_as_float(v[g16][4 * i + 1]) + b4.y;
              fg[4 * i + 2] = __uint_as_float(v[g16][4 * i + 2]) + b4.z;
              fg[4 * i + 3] = __uint_as_float(v[g16][4 * i + 3]) + b4.w;
            }
          }
          if (p.res_mode != VCB_RES_NONE) {
            float rr[16];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (row_ok && n0 + hh * 8 < p.cout_store) {
                const uint4 rv = __ldcg(reinterpret_cast<const uint4*>(p.residual + res_row + n0 + hh * 8));
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 t = __half22float2(rh[i]);
                  rr[hh * 8 + 2 * i] = t.x;
                  rr[hh * 8 + 2 * i + 1] = t.y;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) rr[hh * 8 + i] = 0.f;
              }
            }
            if (p.res_mode == VCB_RES_BEFORE_ACT) {
#pragma unroll
              for (int i = 0; i < 16; ++i) fg[i] = apply_act(fg[i] + rr[i], p.act);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) fg[i] = apply_act(fg[i], p.act) + rr[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) fg[i] = apply_act(fg[i], p.act);
          }
          if (p.epi_direct) {
            if (row_ok) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int nn = n0 + hh * 8;
                if (nn >= p.cout_store) continue;
                if (p.out_fp32) {
                  float* o = reinterpret_cast<float*>(p.out) + out_row + nn;
                  *reinterpret_cast<float4*>(o) = make_float4(fg[hh * 8 + 0], fg[hh * 8 + 1], fg[hh * 8 + 2], fg[hh * 8 + 3]);
                  *reinterpret_cast<float4*>(o + 4) = make_float4(fg[hh * 8 + 4], fg[hh * 8 + 5], fg[hh * 8 + 6], fg[hh * 8 + 7]);
                } else {
                  __half2 h2[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) h2[i] = __floats2half2_rn(fg[hh * 8 + 2 * i], fg[hh * 8 + 2 * i + 1]);
                  __half* o = reinterpret_cast<__half*>(p.out) + out_row + nn;
                  *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(h2);
                }
              }
            }
          } else {
            // staged row = 128 bytes; this thread owns 16-byte chunks [half*4, half*4+4); 128-byte swizzle
            const uint32_t row_addr = stage_buf + (uint32_t)row_in_tile * 128u;
            const int sw = row_in_tile & 7;
            if (p.out_fp32) {            // 16 floats = 4 chunks
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t dst = row_addr + (uint32_t)(((half * 4 + i) ^ sw) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(fg[4 * i]), "f"(fg[4 * i + 1]),
                             "f"(fg[4 * i + 2]), "f"(fg[4 * i + 3]) : "memory");
              }
            } else {                     // 16 halves = 2 chunks per 16-column group
              uint32_t h2[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __half2 t = __floats2half2_rn(fg[2 * i], fg[2 * i + 1]);
                h2[i] = *reinterpret_cast<const uint32_t*>(&t);
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint32_t dst = row_addr + (uint32_t)(((half * 4 + g16 * 2 + i) ^ sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h2[4 * i]), "r"(h2[4 * i + 1]),
                             "r"(h2[4 * i + 2]), "r"(h2[4 * i + 3]) : "memory");
              }
            }
          }
        }
        if (!p.epi_direct) {
          fence_proxy_async_smem();                         // staged writes -> visible to the TMA (async proxy)
          // two buffers: draining every earlier store BEFORE this barrier tells all threads that the other buffer (next
          // sub-tile's target) is free, so one barrier per sub-tile suffices
          if (issuer && p.out_stage_bufs == 2) tma_store_wait_read<0>();
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (issuer) {
            tma_store_2d(&tmap_out, stage_buf, n_tile * p.block_n + sub * sub_cols, m_tile * kBlockM);
            tma_store_commit();
          }
        }
      }
      tcgen05_fence_before();
      if (kTwoCta && cta_rank != 0) mbar_arrive_remote(tmem_empty_bar(acc), 0);
      else mbar_arrive(tmem_empty_bar(acc));
    }
    if (!p.epi_direct && issuer) tma_store_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------------------------
// Specialised epilogue (single-CTA TMA kernel, staged TMA store): activation / residual mode / output type are
// template parameters, the bias comes from shared memory (broadcast reads), SiLU is 5 (ex2 + rcp) or 3 (tanh)
// instructions per element with no range fix-ups -- ~6 thread instructions per output instead of ~23 in the generic
// epilogue (ncu, profiles/r01_ncu_*): small-K layers (1x1 convs, stems) are bound by this loop, not by the MMAs.
// ---------------------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_fast(float v) {
  if (ACT == VCB_ACT_SILU) {             // v * 1/(1 + 2^(-v*log2e)); 1+e is in [1, inf] so rcp needs no scaling
#ifdef VCB_SILU_NR
    // A/B build (tools/ab_build.sh nr -DVCB_SILU_NR): ONE special-function operation per element.  The reciprocal of d = 1 + e runs on
    // the FMA pipe: seed from the exponent trick (as_float(0x7EF127EA - as_int(d)), |rel err| <= 5.1e-2), two Newton steps
    // r <- r * (2 - d * r) (2.6e-3, then 6.6e-6 -- far below the fp16 rounding of the stored result).  The exponent is clamped at
    // 126 so that d stays finite (v < -87: the result is < 1e-36 either way).
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v * -1.4426950408889634f, 126.0f)));
    const float d = 1.0f + e;
    float r = __int_as_float(0x7EF127EA - __float_as_int(d));
    r = r * fmaf(-d, r, 2.0f);
    r = r * fmaf(-d, r, 2.0f);
    return v * r;
#else
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
#endif
  }
  if (ACT == VCB_ACT_SILU_TANH) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  if (ACT == VCB_ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

// Four activations at once.  SiLU = v / (1 + 2^(-v*log2e)) needs one reciprocal per element; the special-function unit (16 lanes
// per clock per SM against 128 for the FMA pipe) is the busiest pipe of the 1x1 / stem epilogues (ncu: XU 50 % on 1x1 96->96,
// profiles/r01_ncu_conv_kernels.md).  One rcp.approx serves four elements: r = 1/(d0*d1*d2*d3), 1/d0 = r*d1*(d2*d3), ... -- 1.25
// MUFU operations per element instead of 2, at 3.25 extra FMULs.  The exponent is clamped at 31 so that the product of four
// denominators stays below 2^124 (no overflow, no denormal reciprocal); below v = -21.5 the exact result is under half an fp16
// subnormal (|v * sigmoid(v)| < 1.1e-8), and so is the clamped one.
template <int ACT>
__device__ __forceinline__ void act_fast4(float& v0, float& v1, float& v2, float& v3) {
#ifdef VCB_SILU_BATCH4     // A/B build only (tools/ab_build.sh batch4 -DVCB_SILU_BATCH4).  Measured on B200 (round 2, call 1): every 1x1 SiLU layer
                           // got 8-12 % SLOWER (192->192 at M=102400: 28.7 -> 32.8 us) -- the epilogue is bound by its dependent-latency chain, not by
                           // MUFU throughput, and the shared reciprocal lengthens that chain.  Default: one ex2 + one rcp per element.
  if (ACT == VCB_ACT_SILU) {
    float e0, e1, e2, e3, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(v0 * -1.4426950408889634f, 31.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(v1 * -1.4426950408889634f, 31.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fminf(v2 * -1.4426950408889634f, 31.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fminf(v3 * -1.4426950408889634f, 31.0f)));
    const float d0 = 1.0f + e0, d1 = 1.0f + e1, d2 = 1.0f + e2, d3 = 1.0f + e3;
    const float p01 = d0 * d1, p23 = d2 * d3;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    const float r01 = r * p23, r23 = r * p01;        // 1/(d0*d1), 1/(d2*d3)
    v0 *= r01 * d1; v1 *= r01 * d0; v2 *= r23 * d3; v3 *= r23 * d2;
  } else
#endif
  {
    v0 = act_fast<ACT>(v0); v1 = act_fast<ACT>(v1); v2 = act_fast<ACT>(v2); v3 = act_fast<ACT>(v3);
  }
}

template <int ACT, int RES, bool F32OUT, bool M256, bool TWO_CTA = false, bool PATCH = false>
__device__ __forceinline__ void conv_epilogue_fast(const ConvParams& p, const CUtensorMap* tmap_out_ptr, const CUtensorMap* tmap_res_ptr,
                                                   uint32_t tmem_base, uint32_t out_stage, uint32_t bias_smem, uint32_t res_bar0,
                                                   uint32_t tmem_full0, uint32_t tmem_empty0, int first_tile = (int)blockIdx.x,
                                                   int tile_step = (int)gridDim.x, int cta_rank = 0) {
  // TWO_CTA: this CTA is one half of a cta_group::2 pair; it drains its own 128 accumulator rows of the 256-row pair tile
  // (tiles are numbered per cluster) and releases the accumulator on the LEADER's tmem_empty barrier.
  const int total_tiles = TWO_CTA ? p.num_pair_tiles : p.num_tiles;
  // Residual (RES != NONE): the residual sub-tile is fetched by TMA INTO the staging buffer (same 128-byte-swizzled layout
  // as the output sub-tile), every thread reads its own 16-byte chunks, adds, and writes the result back in place.  Per-thread
  // global loads of a row-per-lane layout touch 32 lines per request and, with no L1 left beside ~220 KiB of shared memory,
  // cost ~17k cycles per 128x192 tile (role timers, profiles/r01_role_timers.md); the TMA fetch is coalesced and asynchronous.
  // M256 = false: the CTA tile is 128 rows; warp (q = warp & 3, half = warp >> 2) drains rows [32q, 32q+32) and one
  //               half of the columns of each staged sub-tile.
  // M256 = true:  the CTA tile is 256 rows in two accumulators (chains); warps 0-3 drain chain 0 (rows 0-127), warps 4-7
  //               chain 1 (rows 128-255), every thread all columns of the sub-tile in two passes of 32 (fp16) / 16 (fp32).
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q = warp & 3;                      // TMEM lane quarter this warp may read
  const int hi = warp >> 2;                    // column half (M256 = false) or accumulator chain (M256 = true)
  constexpr int TILE_M = M256 ? 256 : 128;
  const int row_in_tile = (M256 ? hi * 128 : 0) + q * 32 + lane;
  constexpr int SUB_COLS = F32OUT ? 32 : 64;   // columns per staged sub-tile (128 bytes per row)
  constexpr int MY_COLS = SUB_COLS / 2;        // columns per thread per pass: 32 (fp16) or 16 (fp32)
  constexpr int NG = MY_COLS / 16;             // 16-column groups per pass
  constexpr int PASSES = M256 ? 2 : 1;
  constexpr uint32_t STAGE_BYTES = (uint32_t)TILE_M * 128u;
  const int num_sub = (p.block_n + SUB_COLS - 1) / SUB_COLS;
  const bool issuer = threadIdx.x == 0;
  // PATCH: accumulator row m is lattice point (i, j) = (m / Lp, m % Lp) of the R x (Xs+2) output lattice; columns j >= Xs and
  // rows i >= R are junk.  Valid points are staged densely (row i*Xs + j) so that one 4-D TMA box {64, Xs, R, 1} stores them.
  const int lat_i = PATCH ? row_in_tile / p.patch_Lp : 0;
  const int lat_j = PATCH ? row_in_tile - lat_i * p.patch_Lp : 0;
  const bool lat_ok = !PATCH || (lat_i < p.patch_R && lat_j < p.patch_Xs);
  const int stage_row = PATCH ? lat_i * p.patch_Xs + lat_j : row_in_tile;
  const uint32_t row_off = (uint32_t)stage_row * 128u;
  const int sw = stage_row & 7;
  const uint32_t res_bytes = PATCH ? (uint32_t)p.patch_out_bytes : STAGE_BYTES;
  auto issue_res_load = [&](uint32_t bar, uint32_t dst, int c0, int mt) {
    mbar_arrive_expect_tx(bar, res_bytes);
    if (PATCH) {
      const int per_img = p.patch_ytiles * p.patch_xsegs;
      const int ni = mt / per_img, rem = mt - ni * per_img;
      const int yt = rem / p.patch_xsegs, xs = rem - yt * p.patch_xsegs;
      tma_load_4d(tmap_res_ptr, bar, dst, c0, xs * p.patch_Xs, yt * p.patch_R, ni);
    } else {
      tma_load_2d(tmap_res_ptr, bar, dst, c0, mt * TILE_M);
    }
  };
  auto issue_store = [&](uint32_t src, int c0, int mt) {
    if (PATCH) {
      const int per_img = p.patch_ytiles * p.patch_xsegs;
      const int ni = mt / per_img, rem = mt - ni * per_img;
      const int yt = rem / p.patch_xsegs, xs = rem - yt * p.patch_xsegs;
      tma_store_4d(tmap_out_ptr, src, c0, xs * p.patch_Xs, yt * p.patch_R, ni);
    } else {
      tma_store_2d(tmap_out_ptr, src, c0, mt * TILE_M);
    }
  };
  const bool two_bufs = p.out_stage_bufs == 2;
  // TMEM layout of one accumulator stage: [K chain][row chain (M256 only)][block_n columns]
  const uint32_t kstride = (uint32_t)p.block_n * (M256 ? 2u : 1u);
  const int kchains = p.kchains;
  const uint32_t acc_stride = kstride * (uint32_t)kchains;
  uint32_t tile_iter = 0, sub_count = 0;
  long long t_wait = 0, t_sync = 0;
  const long long t_begin = prof_clock(issuer ? p.prof : nullptr);
  auto map_tile = [&](int seq) { return p.tile_rev ? total_tiles - 1 - seq : seq; };
  if (RES != VCB_RES_NONE && issuer && first_tile < total_tiles) {      // residual of the first sub-tile
    const int ft = map_tile(first_tile);
    const int pm0 = ft / p.n_tiles, nt0 = ft - pm0 * p.n_tiles;
    const int mt0 = TWO_CTA ? 2 * pm0 + cta_rank : pm0;
    issue_res_load(res_bar0, out_stage, nt0 * p.block_n, mt0);
  }
  for (int tile_seq = first_tile; tile_seq < total_tiles; tile_seq += tile_step, ++tile_iter) {
    const int tile = map_tile(tile_seq);
    const int pm_tile = tile / p.n_tiles;
    const int n_tile = tile - pm_tile * p.n_tiles;
    const int m_tile = TWO_CTA ? 2 * pm_tile + cta_rank : pm_tile;
    const uint32_t acc = (p.acc_stages == 2) ? (tile_iter & 1u) : 0u;
    const uint32_t acc_ph = (p.acc_stages == 2) ? ((tile_iter >> 1) & 1u) : (tile_iter & 1u);
    const long long tw0 = prof_clock(issuer ? p.prof : nullptr);
    epi_wait_full(tmem_full0 + 8u * acc, acc_ph, p.fault, (int)acc);
    t_wait += prof_clock(issuer ? p.prof : nullptr) - tw0;
    tcgen05_fence_after();
    if (p.dbg_skip_epilogue) {
      tcgen05_fence_before();
      mbar_arrive(tmem_empty0 + 8u * acc);
      continue;
    }
    const int row = m_tile * TILE_M + row_in_tile;
    const bool row_ok = row < p.M;
    const uint32_t t_row = tmem_base + acc * acc_stride + (M256 ? (uint32_t)(hi * p.block_n) : 0u) + ((uint32_t)(q * 32) << 16);
    const int n_base = n_tile * p.block_n;
    (void)row; (void)row_ok;
    for (int sub = 0; sub < num_sub; ++sub, ++sub_count) {
      const uint32_t buf_idx = two_bufs ? (sub_count & 1u) : 0u;
      const uint32_t stage_buf = out_stage + buf_idx * STAGE_BYTES;
      if (RES != VCB_RES_NONE) {
        // the residual sub-tile was requested one sub-tile ago (the request waited for the buffer's previous store)
        mbar_wait(res_bar0 + 8u * buf_idx, (two_bufs ? (sub_count >> 1) : sub_count) & 1u, p.fault, FAULT_FULL_WAIT, 400 + (int)buf_idx);
      } else if (!two_bufs) {                            // single staging buffer: drain the previous store first
        if (issuer) tma_store_wait_read<0>();
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
#pragma unroll
      for (int pass = 0; pass < PASSES; ++pass) {
        const int colhalf = M256 ? pass : hi;
        const int col_base = sub * SUB_COLS + colhalf * MY_COLS;         // first column (within the N tile) of this pass
        const int ngroups = (col_base >= p.block_n) ? 0 : ((NG == 2 && col_base + 16 < p.block_n) ? 2 : 1);   // warp-uniform
        uint32_t v[NG][16];
        if (ngroups > 0) tmem_ld_x16(t_row + (uint32_t)col_base, v[0]);
        if (NG == 2 && ngroups > 1) tmem_ld_x16(t_row + (uint32_t)(col_base + 16), v[NG - 1]);
        if (kchains > 1 && ngroups > 0) {                  // sum the K chains (fp32)
          for (int c = 1; c < kchains; ++c) {
            uint32_t w[NG][16];
            tmem_ld_x16(t_row + (uint32_t)c * kstride + (uint32_t)col_base, w[0]);
            if (NG == 2 && ngroups > 1) tmem_ld_x16(t_row + (uint32_t)c * kstride + (uint32_t)(col_base + 16), w[NG - 1]);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < NG; ++g)
              if (g < ngroups) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[g][i] = __float_as_uint(__uint_as_float(v[g][i]) + __uint_as_float(w[g][i]));
              }
          }
        }
        uint4 rv[NG][2];
        if (RES != VCB_RES_NONE) {                         // this thread's own chunks of the staged residual sub-tile
#pragma unroll
          for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t src = stage_buf + row_off + (uint32_t)(((colhalf * 4 + g * 2 + hh) ^ sw) << 4);
              if (g < ngroups && lat_ok)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv[g][hh].x), "=r"(rv[g][hh].y), "=r"(rv[g][hh].z), "=r"(rv[g][hh].w) : "r"(src));
              else
                rv[g][hh] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (ngroups > 0) tmem_ld_wait();
        if (sub == num_sub - 1 && pass == PASSES - 1) {    // accumulator fully read: hand it back before the math / stores
          tcgen05_fence_before();
          if (TWO_CTA && cta_rank != 0) mbar_arrive_remote(tmem_empty0 + 8u * acc, 0);
          else mbar_arrive(tmem_empty0 + 8u * acc);
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (g >= ngroups) continue;
          const uint32_t b_addr = bias_smem + (uint32_t)(n_base + col_base + g * 16) * 4u;
          float f[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(b_addr + 16u * i));
            f[4 * i + 0] = __uint_as_float(v[g][4 * i + 0]) + b0;
            f[4 * i + 1] = __uint_as_float(v[g][4 * i + 1]) + b1;
            f[4 * i + 2] = __uint_as_float(v[g][4 * i + 2]) + b2;
            f[4 * i + 3] = __uint_as_float(v[g][4 * i + 3]) + b3;
          }
          if (RES != VCB_RES_NONE) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const __half2* rh = reinterpret_cast<const __half2*>(&rv[g][hh]);
#pragma unroll
              for (int i = 0; i < 4; i += 2) {                    // four elements = two half2 of the staged residual
                const float2 ta = __half22float2(rh[i]), tb = __half22float2(rh[i + 1]);
                float& f0 = f[hh * 8 + 2 * i]; float& f1 = f[hh * 8 + 2 * i + 1];
                float& f2 = f[hh * 8 + 2 * i + 2]; float& f3 = f[hh * 8 + 2 * i + 3];
                if (RES == VCB_RES_BEFORE_ACT) { f0 += ta.x; f1 += ta.y; f2 += tb.x; f3 += tb.y; }
                act_fast4<ACT>(f0, f1, f2, f3);
                if (RES == VCB_RES_AFTER_ACT) { f0 += ta.x; f1 += ta.y; f2 += tb.x; f3 += tb.y; }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i += 4) act_fast4<ACT>(f[i], f[i + 1], f[i + 2], f[i + 3]);
          }
          // staged row = 128 bytes; this pass owns 16-byte chunks [colhalf*4, colhalf*4+4); 128-byte swizzle
          const uint32_t row_addr = stage_buf + row_off;
          if (!lat_ok) {
            // junk lattice point (patch mode): nothing to stage
          } else if (F32OUT) {            // 16 floats = 4 chunks
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t dst = row_addr + (uint32_t)(((colhalf * 4 + i) ^ sw) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f[4 * i]), "f"(f[4 * i + 1]), "f"(f[4 * i + 2]),
                           "f"(f[4 * i + 3]) : "memory");
            }
          } else {                 // 16 halves = 2 chunks per 16-column group
            uint32_t h2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
              h2[i] = *reinterpret_cast<const uint32_t*>(&t);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t dst = row_addr + (uint32_t)(((colhalf * 4 + g * 2 + i) ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h2[4 * i]), "r"(h2[4 * i + 1]), "r"(h2[4 * i + 2]),
                           "r"(h2[4 * i + 3]) : "memory");
            }
          }
        }
      }
      fence_proxy_async_smem();                         // staged writes -> visible to the TMA (async proxy)
      const long long ts0 = prof_clock(issuer ? p.prof : nullptr);
      if (issuer && two_bufs) tma_store_wait_read<0>();   // the other buffer (next sub-tile's target) is free after the barrier
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (issuer) {
        issue_store(stage_buf, n_base + sub * SUB_COLS, m_tile);
        tma_store_commit();
        if (RES != VCB_RES_NONE) {                        // request the residual of the next sub-tile (maybe of the next tile)
          int nseq = tile_seq, nsub = sub + 1;
          if (nsub == num_sub) { nsub = 0; nseq = tile_seq + tile_step; }
          if (nseq < total_tiles) {
            if (!two_bufs) tma_store_wait_read<0>();        // same buffer: the store just issued must have read it
            const int nt = map_tile(nseq);
            const int npm = nt / p.n_tiles, nn = nt - npm * p.n_tiles;
            const int nm = TWO_CTA ? 2 * npm + cta_rank : npm;
            const uint32_t nb = two_bufs ? ((sub_count + 1u) & 1u) : 0u;
            issue_res_load(res_bar0 + 8u * nb, out_stage + nb * STAGE_BYTES, nn * p.block_n + nsub * SUB_COLS, nm);
          }
        }
      }
      t_sync += prof_clock(issuer ? p.prof : nullptr) - ts0;
    }
  }
  if (issuer) tma_store_wait_all<0>();
  if (issuer && p.prof) {
    prof_add(p.prof, PROF_EPI_WAIT_TMEM, t_wait);
    prof_add(p.prof, PROF_EPI_SYNC_STORE, t_sync);
    prof_add(p.prof, PROF_EPI_TOTAL, clock64() - t_begin);
    prof_add(p.prof, PROF_TILES, tile_iter);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Split epilogue (round 2).  ncu + role timers on the 1x1 / stem layers (profiles/r02_epilogue_analysis.md): these layers are
// bound by the LATENCY of the epilogue's dependent chain per sub-tile (accumulator wait -> tcgen05.ld -> bias -> ex2 -> rcp ->
// cvt -> st.shared -> fence -> CTA-wide barrier -> TMA store), not by MUFU or issue throughput: eight warps in lockstep leave
// the SM with two such chains (two CTAs).  Here the eight warps are TWO independent groups of four (one TMEM lane quarter per
// warp, 128 rows per group); a group takes every other work item -- whole tiles when the layer has one 64-column sub-tile
// (each group then owns one accumulator stage), else every other (tile, sub-tile) -- and handles all 64 columns of it in two
// halves.  Each group has its own staging buffer, 128-thread named barrier, store-issuing thread and residual barrier, so the
// SM runs four epilogue chains out of phase and the per-item fixed costs are paid once per 64 columns instead of once per 32.
// The buffer-free handshake (previous store has read the staging buffer) sits after the first half's arithmetic, where it is
// normally already satisfied.  Requires two staging buffers, one K chain and, in whole-tile mode, two accumulator stages.
// ---------------------------------------------------------------------------------------------------------------
template <int ACT, int RES, bool F32OUT, bool TWO_CTA>
__device__ __forceinline__ void conv_epilogue_split(const ConvParams& p, const CUtensorMap* tmap_out_ptr, const CUtensorMap* tmap_res_ptr,
                                                    uint32_t tmem_base, uint32_t out_stage, uint32_t bias_smem, uint32_t res_bar0,
                                                    uint32_t tmem_full0, uint32_t tmem_empty0, int first_tile, int tile_step, int cta_rank,
                                                    float* stat_scratch) {
  const int total_tiles = TWO_CTA ? p.num_pair_tiles : p.num_tiles;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q = warp & 3;                      // TMEM lane quarter
  const int grp = warp >> 2;                   // epilogue group 0 / 1
  const int row_in_tile = q * 32 + lane;
  constexpr int SUB_COLS = F32OUT ? 32 : 64;   // columns per staged sub-tile (128 bytes per row)
  constexpr int HALF = SUB_COLS / 2;           // columns per half: 32 (fp16) or 16 (fp32)
  constexpr int NG = HALF / 16;                // 16-column TMEM loads per half
  const int num_sub = (p.block_n + SUB_COLS - 1) / SUB_COLS;
  const bool issuer = (threadIdx.x & 127) == 0;
  const uint32_t stage_buf = out_stage + (uint32_t)grp * kStageOutBytes;
  const uint32_t res_bar = res_bar0 + 8u * (uint32_t)grp;
  const uint32_t row_addr = stage_buf + (uint32_t)row_in_tile * 128u;
  const int sw = row_in_tile & 7;
  const uint32_t bar_id = 2u + (uint32_t)grp;
  const bool two_acc = p.acc_stages == 2;
  const int n_tiles = p.n_tiles;
  const int per_img = p.patch_ytiles * p.patch_xsegs;      // row-window mode only
  if (p.stat_sums != nullptr) stat_scratch[threadIdx.x] = 0.f;      // 256 floats: two groups x 64 columns x 2 statistics

  // (tile ordinal, sub-tile) of this group's current item; items alternate between the groups
  int t_idx = (num_sub == 1) ? grp : 0;
  int sub = (num_sub == 1) ? 0 : grp;
  int waited = -1;                                          // tile ordinal whose accumulator this thread has waited for
  uint32_t my_items = 0;

  auto decode = [&](int tidx, int& m_tile, int& n_base) {
    const int seq = first_tile + tidx * tile_step;
    const int tile = p.tile_rev ? total_tiles - 1 - seq : seq;
    const int pm = (n_tiles == 1) ? tile : tile / n_tiles;
    n_base = (n_tiles == 1) ? 0 : (tile - pm * n_tiles) * p.block_n;
    m_tile = TWO_CTA ? 2 * pm + cta_rank : pm;
  };
  auto issue_res_load = [&](int tidx, int sb) {
    int mt, nb;
    decode(tidx, mt, nb);
    mbar_arrive_expect_tx(res_bar, (uint32_t)kStageOutBytes);
    tma_load_2d(tmap_res_ptr, res_bar, stage_buf, nb + sb * SUB_COLS, mt * kBlockM);
  };
  if (RES != VCB_RES_NONE && issuer && first_tile + t_idx * tile_step < total_tiles) issue_res_load(t_idx, sub);

  while (first_tile + t_idx * tile_step < total_tiles) {
    int m_tile, n_base;
    decode(t_idx, m_tile, n_base);
    const uint32_t acc = two_acc ? ((uint32_t)t_idx & 1u) : 0u;
    if (t_idx != waited) {
      const uint32_t acc_ph = two_acc ? (((uint32_t)t_idx >> 1) & 1u) : ((uint32_t)t_idx & 1u);
      epi_wait_full(tmem_full0 + 8u * acc, acc_ph, p.fault, (int)acc);
      tcgen05_fence_after();
      waited = t_idx;
    }
    const bool last_in_tile = sub + 2 >= num_sub;           // this group's last item of the tile
    const uint32_t t_row = tmem_base + acc * (uint32_t)p.block_n + ((uint32_t)(q * 32) << 16);
    const int col0 = sub * SUB_COLS;
    const int ncols = min(SUB_COLS, p.block_n - col0);     // multiple of 16
    if (RES != VCB_RES_NONE) mbar_wait(res_bar, my_items & 1u, p.fault, FAULT_FULL_WAIT, 400 + grp);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cb = col0 + h * HALF;                        // first column (within the N tile) of this half
      int ng = (ncols - h * HALF) >> 4;                      // warp-uniform
      ng = ng < 0 ? 0 : (ng > NG ? NG : ng);
      uint32_t v[NG][16];
      if (ng > 0) tmem_ld_x16(t_row + (uint32_t)cb, v[0]);
      if (NG == 2 && ng > 1) tmem_ld_x16(t_row + (uint32_t)(cb + 16), v[NG - 1]);
      uint4 rv[NG][2];
      if (RES != VCB_RES_NONE) {                             // this thread's own chunks of the staged residual sub-tile
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t src = row_addr + (uint32_t)(((h * 4 + g * 2 + hh) ^ sw) << 4);
            if (g < ng)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv[g][hh].x), "=r"(rv[g][hh].y), "=r"(rv[g][hh].z), "=r"(rv[g][hh].w) : "r"(src));
            else
              rv[g][hh] = make_uint4(0u, 0u, 0u, 0u);
          }
      }
      if (ng > 0) tmem_ld_wait();
      if (h == 1 && last_in_tile) {                          // this thread is done with the accumulator
        tcgen05_fence_before();
        if (TWO_CTA && cta_rank != 0) mbar_arrive_remote(tmem_empty0 + 8u * acc, 0);
        else mbar_arrive(tmem_empty0 + 8u * acc);
      }
      uint32_t packed[NG][F32OUT ? 16 : 8];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (g >= ng) continue;
        const uint32_t b_addr = bias_smem + (uint32_t)(n_base + cb + g * 16) * 4u;
        float f[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float b0, b1, b2, b3;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(b_addr + 16u * i));
          f[4 * i + 0] = __uint_as_float(v[g][4 * i + 0]) + b0;
          f[4 * i + 1] = __uint_as_float(v[g][4 * i + 1]) + b1;
          f[4 * i + 2] = __uint_as_float(v[g][4 * i + 2]) + b2;
          f[4 * i + 3] = __uint_as_float(v[g][4 * i + 3]) + b3;
        }
        if (RES != VCB_RES_NONE) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const __half2* rh = reinterpret_cast<const __half2*>(&rv[g][hh]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 t = __half22float2(rh[i]);
              if (RES == VCB_RES_BEFORE_ACT) {
                f[hh * 8 + 2 * i] = act_fast<ACT>(f[hh * 8 + 2 * i] + t.x);
                f[hh * 8 + 2 * i + 1] = act_fast<ACT>(f[hh * 8 + 2 * i + 1] + t.y);
              } else {
                f[hh * 8 + 2 * i] = act_fast<ACT>(f[hh * 8 + 2 * i]) + t.x;
                f[hh * 8 + 2 * i + 1] = act_fast<ACT>(f[hh * 8 + 2 * i + 1]) + t.y;
              }
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = act_fast<ACT>(f[i]);
        }
        if (F32OUT) {
#pragma unroll
          for (int i = 0; i < 16; ++i) packed[g][i] = __float_as_uint(f[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            packed[g][i] = *reinterpret_cast<const uint32_t*>(&t);
          }
        }
      }
      if (RES == VCB_RES_NONE && h == 0) {                   // the group's previous store must have read the staging buffer
        if (issuer) tma_store_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (g >= ng) continue;
        constexpr int CH = F32OUT ? 4 : 2;                   // 16-byte chunks per 16-column group
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const uint32_t dst = row_addr + (uint32_t)(((h * 4 + g * CH + i) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[g][4 * i]), "r"(packed[g][4 * i + 1]),
                       "r"(packed[g][4 * i + 2]), "r"(packed[g][4 * i + 3]) : "memory");
        }
      }
    }
    fence_proxy_async_smem();                                // staged writes -> visible to the TMA (async proxy)
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    // next item of this group (two items ahead in the CTA's sequence)
    int nt = t_idx, ns = sub + 2;
    if (ns >= num_sub) { ns -= num_sub; ++nt; if (ns >= num_sub) { ns -= num_sub; ++nt; } }
    if (issuer) {
      if (p.a_rowwin) {
        const int ni = m_tile / per_img, rem = m_tile - ni * per_img;
        const int yt = rem / p.patch_xsegs, xs = rem - yt * p.patch_xsegs;
        tma_store_4d(tmap_out_ptr, stage_buf, n_base + col0, xs * p.patch_Xs, yt * p.patch_R, ni);
      } else {
        tma_store_2d(tmap_out_ptr, stage_buf, n_base + col0, m_tile * kBlockM);
      }
      tma_store_commit();
      if (RES != VCB_RES_NONE && first_tile + nt * tile_step < total_tiles) {
        tma_store_wait_read<0>();                            // same buffer: the store just issued must have read it
        issue_res_load(nt, ns);
      }
    }
    if (ACT == VCB_ACT_NONE && RES == VCB_RES_NONE && !F32OUT && p.stat_sums != nullptr) {
      // BatchNorm statistics of this 128 x 64 sub-tile, read back from the staged fp16 tile (the values the normalisation pass will
      // see): thread (column pair cp, row quarter rq) sums 32 rows, the four row quarters meet in shared memory, then ONE fp64 atomic
      // per (column, statistic).  Tiles that straddle two segments (rare: segment boundaries only) take the per-row path.
      const int t = threadIdx.x & 127;
      const int cp = t & 31, rq = t >> 5;
      const int m0 = m_tile * kBlockM;
      const int rows_valid = min(kBlockM, p.M - m0);
      const bool col_ok = 2 * cp < ncols && n_base + col0 + 2 * cp < p.Cout;
      const uint32_t cbase = stage_buf + (uint32_t)(cp & 3) * 4u;
      const uint32_t chunk = (uint32_t)(cp >> 2);
      int seg_a = 0, seg_b = 0;
      if (rows_valid > 0) {
        seg_a = __ldg(p.stat_seg + m0 / p.PQ);
        seg_b = __ldg(p.stat_seg + (m0 + rows_valid - 1) / p.PQ);
      }
      const int r_lo = rq * 32, r_hi = min(r_lo + 32, rows_valid);
      float* scratch = stat_scratch + grp * 128;
      if (seg_a == seg_b) {
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        if (col_ok) {
          for (int r = r_lo; r < r_hi; ++r) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(cbase + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
            s0 += f.x; s1 += f.y;
            q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
          }
          atomicAdd(scratch + 4 * cp + 0, s0); atomicAdd(scratch + 4 * cp + 1, q0);
          atomicAdd(scratch + 4 * cp + 2, s1); atomicAdd(scratch + 4 * cp + 3, q1);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        {
          const int c_ = t >> 1, ch = n_base + col0 + c_;          // slot t = column t / 2, statistic t % 2
          if (rows_valid > 0 && c_ < ncols && ch < p.Cout) atomicAdd(p.stat_sums + ((long long)seg_a * p.Cout + ch) * 2 + (t & 1), (double)scratch[t]);
          scratch[t] = 0.f;
        }
      } else if (col_ok) {
        const int ch = n_base + col0 + 2 * cp;
        for (int r = r_lo; r < r_hi; ++r) {
          uint32_t w;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(cbase + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
          double* dst = p.stat_sums + ((long long)__ldg(p.stat_seg + (m0 + r) / p.PQ) * p.Cout + ch) * 2;
          atomicAdd(dst + 0, (double)f.x); atomicAdd(dst + 1, (double)f.x * (double)f.x);
          atomicAdd(dst + 2, (double)f.y); atomicAdd(dst + 3, (double)f.y * (double)f.y);
        }
      }
    }
    t_idx = nt; sub = ns;
    ++my_items;
  }
  if (issuer) tma_store_wait_all<0>();
}

// epilogue variants with a specialised instance; anything else runs the generic conv_epilogue
__host__ __device__ constexpr int epi_kind_of(int act, int res, int f32) {
  return f32 ? ((act == VCB_ACT_NONE && res == VCB_RES_NONE) ? 7 : 0)
             : (act == VCB_ACT_SILU ? (res == VCB_RES_NONE ? 1 : (res == VCB_RES_AFTER_ACT ? 2 : 0))
                : act == VCB_ACT_SILU_TANH ? (res == VCB_RES_NONE ? 3 : (res == VCB_RES_AFTER_ACT ? 4 : 0))
                : act == VCB_ACT_RELU ? (res == VCB_RES_NONE ? 5 : (res == VCB_RES_BEFORE_ACT ? 6 : 0))
                : (res == VCB_RES_NONE ? 8 : 0));
}

// Shared-memory carve-up (all offsets from a 1024-byte aligned base):
//   [num_stages x (A tile 16 KiB | B tile block_n*128 B)] [2 x 16 KiB output staging] [barriers] [tmem slot] [row table]
// M256 (A_TMA only): the CTA tile is 256 output pixels x N held in TWO accumulators ("chains": rows 0-127 and 128-255).
// Every weight tile that reaches shared memory feeds both chains, so the L2 -> SM weight traffic per output is halved
// (the whole path is bound by the ~6300 B/clk chip-wide L2 throughput), and the two independent
// accumulation chains keep the tensor pipe busy from ONE resident CTA that owns all of the SM's shared memory.
template <int A_MODE, int BK, bool M256>
__global__ void __launch_bounds__(A_MODE == A_TMA ? kThreadsTma : kThreadsGather, (A_MODE == A_TMA && !M256) ? 2 : 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, const ConvParams p) {
  static_assert(!M256 || A_MODE == A_TMA, "256-row tiles exist for the TMA path only");
  constexpr int kTileM = M256 ? 2 * kBlockM : kBlockM;
  constexpr uint32_t kATile = (uint32_t)kTileM * kBlockK * 2;        // bytes of A per pipeline stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t stage_bytes = kATile + (p.b_resident ? 0u : b_tile_bytes);
  const uint32_t b_res = smem_base + (uint32_t)p.num_stages * stage_bytes;              // resident weights (may be empty)
  const uint32_t out_stage = b_res + (uint32_t)p.b_res_bytes;                             // 1024-aligned
  const uint32_t bars = out_stage + (uint32_t)p.out_stage_bytes;                          // 8-byte aligned
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(kMaxStages + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kMaxStages + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 4);
  const uint32_t bres_bar = tmem_slot + 8u;          // second half of the 16-byte slot region
  const uint32_t res_bar0 = tmem_slot + 16u;         // 2 residual-arrival barriers (TMA path: the gather row table is unused)
  // generic pointers to the same locations (for plain loads/stores)
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* tail_gen = smem_gen + (size_t)p.num_stages * stage_bytes + p.b_res_bytes + p.out_stage_bytes + 8 * (2 * kMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail_gen);
  RowInfo* rows = reinterpret_cast<RowInfo*>(tail_gen + 16);
  const uint32_t bias_smem = tmem_slot + 16u + (uint32_t)(kBlockM * sizeof(RowInfo));    // fp32 [cout_pad], 16-byte aligned
  float* bias_gen = reinterpret_cast<float*>(tail_gen + 16 + kBlockM * sizeof(RowInfo));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.epi_kind != 0)        // the packed bias is a constant of the layer: staged once, read as smem broadcasts
    for (int i = threadIdx.x; i < p.cout_pad; i += blockDim.x) bias_gen[i] = __ldg(p.bias + i);

  // Programmatic dependent launch: the next kernel of the stream may start its own set-up (barrier init, TMEM
  // allocation, tensor-map prefetch) on SMs this grid no longer fills; it blocks in griddep_wait() until this grid
  // has completed and flushed.  Both instructions are no-ops for a normally serialised launch.
  griddep_launch_dependents();
  const long long t_cta0 = prof_clock(p.prof);

  if (threadIdx.x == 0) {
    const uint32_t full_count = (A_MODE == A_TMA) ? (p.split_b ? 2u : 1u) : (1u + kGatherThreads);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar(s), full_count);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), (uint32_t)p.epi_empty_count);
    }
    mbar_init(bres_bar, 1);
    if (A_MODE == A_TMA) { mbar_init(res_bar0, 1); mbar_init(res_bar0 + 8u, 1); }
    fence_mbar_init();
  }
  if (warp == kProducerWarp && lane == 0) {
    if (A_MODE == A_TMA) tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (!p.epi_direct) tma_prefetch_desc(&tmap_out);
    if (A_MODE == A_TMA && p.epi_kind != 0 && p.res_mode != VCB_RES_NONE) tma_prefetch_desc(&tmap_res);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_wait();      // everything below may read what earlier kernels of the stream wrote (and overwrite what they read)
  if (threadIdx.x == 0 && p.prof) { prof_add(p.prof, PROF_SETUP, clock64() - t_cta0); prof_add(p.prof, PROF_CTAS, 1); }

  if (warp == kProducerWarp) {
    // ======================= TMA producer (one thread) =======================
    if (lane == 0) {
      long long t_w = 0;
      const long long t_b = prof_clock(p.prof);
      uint32_t it = 0, ring_st = 0, ring_ph = 0;
      if (A_MODE == A_TMA && p.b_resident && (int)blockIdx.x < p.num_tiles) {
        // small layers: every packed weight chunk is loaded once per CTA and stays put
        const uint32_t bc = (uint32_t)(p.block_n * BK * 2);
        mbar_arrive_expect_tx(bres_bar, (uint32_t)p.total_chunks * bc);
        for (int kc = 0; kc < p.total_chunks; ++kc) tma_load_2d(&tmap_b, bres_bar, b_res + (uint32_t)kc * bc, kc * BK, 0, p.b_policy);
      }
      for (int tile_seq = blockIdx.x; tile_seq < p.num_tiles; tile_seq += gridDim.x) {
        const int tile = p.tile_rev ? p.num_tiles - 1 - tile_seq : tile_seq;
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        int cw = 0, ch = 0, cn = 0;
        if (A_MODE == A_TMA && p.a_rowwin) {        // tile = patch_R rows x patch_Xs columns of image cn, top-left (ch, cw)
          const int per_img = p.patch_ytiles * p.patch_xsegs;
          cn = m_tile / per_img;
          const int rem = m_tile - cn * per_img;
          const int yt = rem / p.patch_xsegs;
          ch = yt * p.patch_R;
          cw = (rem - yt * p.patch_xsegs) * p.patch_Xs;
        } else if (A_MODE == A_TMA) {
          const int m0 = m_tile * kTileM;
          cn = m0 / p.PQ;
          const int rem = m0 - cn * p.PQ;
          const int p0 = rem / p.Q, q0 = rem - p0 * p.Q;
          cw = q0 * p.stride - p.pad;
          ch = p0 * p.stride - p.pad;
        }
        int r = 0, s = 0, c = 0, kidx = 0;
        constexpr uint32_t a_chunk = (uint32_t)(kTileM * BK * 2);
        const uint32_t b_chunk = (uint32_t)(p.block_n * BK * 2);
        constexpr int G = kBlockK / BK;
        for (int kit = 0; kit < p.num_k_iters; ++kit, ++it) {
          const int st = (int)ring_st;
          const uint32_t ph = ring_ph;
          if (++ring_st == (uint32_t)p.num_stages) { ring_st = 0; ring_ph ^= 1u; }      // no integer division in the issue loops
          const long long tw = prof_clock(p.prof);
          mbar_wait(empty_bar(st), ph ^ 1u, p.fault, FAULT_EMPTY_WAIT, st);
          t_w += prof_clock(p.prof) - tw;
          const uint32_t a_dst = smem_base + (uint32_t)st * stage_bytes;
          const uint32_t b_dst = a_dst + kATile;
          const int nch = (G == 1) ? 1 : min(G, p.total_chunks - kit * G);
          if (A_MODE == A_TMA) {
            mbar_arrive_expect_tx(full_bar(st), (uint32_t)nch * ((p.split_b || p.b_resident) ? a_chunk : a_chunk + b_chunk));
            for (int g = 0; g < nch; ++g, ++kidx) {
              if (p.a_rowwin) tma_load_4d(&tmap_a, full_bar(st), a_dst, 0, cw, ch - 1 + kit, cn, p.a_policy);
              else if (p.a_tiled) tma_load_2d(&tmap_a, full_bar(st), a_dst + (uint32_t)g * a_chunk, c * BK, m_tile * kTileM, p.a_policy);
              else tma_load_im2col_4d(&tmap_a, full_bar(st), a_dst + (uint32_t)g * a_chunk, c * BK, cw, ch, cn, (uint16_t)s, (uint16_t)r, p.a_policy);
              if (!p.split_b && !p.b_resident)
                tma_load_2d(&tmap_b, full_bar(st), b_dst + (uint32_t)g * b_chunk, kidx * BK, n_tile * p.block_n, p.b_policy);
              if (++c == p.chunks_per_tap) { c = 0; if (++s == p.kw) { s = 0; ++r; } }
            }
          } else {
            mbar_arrive_expect_tx(full_bar(st), b_tile_bytes);
            tma_load_2d(&tmap_b, full_bar(st), b_dst, kit * kBlockK, n_tile * p.block_n, p.b_policy);
          }
        }
      }
      if (p.prof) { prof_add(p.prof, PROF_PROD_WAIT_EMPTY, t_w); prof_add(p.prof, PROF_PROD_TOTAL, clock64() - t_b); }
    }
  } else if (warp == kMmaWarp) {
    // ======================= MMA issuer (one thread) =======================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16((uint32_t)p.block_n);
      constexpr uint32_t a_chunk = (uint32_t)(kTileM * BK * 2);
      constexpr uint32_t a_chain = (uint32_t)(kBlockM * BK * 2);                     // chain 1 = rows 128..255 of each chunk
      const uint32_t b_chunk = (uint32_t)(p.block_n * BK * 2);
      constexpr uint32_t sbo = (uint32_t)(8 * BK * 2);                               // 8 rows of one swizzle row each
      constexpr uint32_t layout_type = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);       // SWIZZLE_128B / 64B / 32B
      constexpr int ksteps = BK / 16;
      constexpr int G = kBlockK / BK;                                                // chunks per stage
      // descriptor high words are loop invariants; only the 14-bit start-address field changes
      const uint64_t desc_hi = umma_desc_kmajor(0, sbo, layout_type);
      uint32_t it = 0, tile_iter = 0, ring_st = 0, ring_ph = 0;
      const int klim = p.ksteps_lim;
      long long t_wf = 0, t_wt = 0;
      const long long t_b = prof_clock(p.prof);
      if (p.b_resident && (int)blockIdx.x < p.num_tiles) mbar_wait(bres_bar, 0u, p.fault, FAULT_FULL_WAIT, 300);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
        const uint32_t acc = (p.acc_stages == 2) ? (tile_iter & 1u) : 0u;
        const uint32_t acc_ph = (p.acc_stages == 2) ? ((tile_iter >> 1) & 1u) : (tile_iter & 1u);
        const long long tw1 = prof_clock(p.prof);
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u, p.fault, FAULT_TMEM_EMPTY_WAIT, (int)acc);
        t_wt += prof_clock(p.prof) - tw1;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.block_n * (M256 ? 2u : (uint32_t)p.kchains);
        uint32_t ks = 0;                                  // K=16 steps issued for this tile (K chains, opt-in)
        const uint32_t chain_mask = (uint32_t)p.kchains - 1u;
        uint32_t accumulate = 0u;
        uint64_t bres_desc = desc_hi | (uint64_t)((b_res & 0x3FFFF) >> 4);      // resident weights: chunk kc at + kc * b_chunk
        for (int kit = 0; kit < p.num_k_iters; ++kit, ++it) {
          // this loop is ONE thread: every integer instruction here delays the next tcgen05.mma (tools/umma_rate_probe.cu), so the
          // ring position is a counter and the descriptors advance by additions
          const int st = (int)ring_st;
          const long long tw2 = prof_clock(p.prof);
          mbar_wait(full_bar(st), ring_ph, p.fault, FAULT_FULL_WAIT, st);
          t_wf += prof_clock(p.prof) - tw2;
          tcgen05_fence_after();
          const uint32_t a_addr = smem_base + ring_st * stage_bytes;
          const uint64_t a_desc0 = desc_hi | (uint64_t)((a_addr & 0x3FFFF) >> 4);
          const uint64_t b_desc0 = p.b_resident ? bres_desc : (a_desc0 + (uint64_t)(kATile >> 4));
          const int nch = (G == 1) ? 1 : min(G, p.total_chunks - kit * G);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (g < nch) {
              const uint64_t a_desc = a_desc0 + (uint64_t)(((uint32_t)g * a_chunk) >> 4);
              const uint64_t b_desc = b_desc0 + (uint64_t)(((uint32_t)g * b_chunk) >> 4);
#pragma unroll
              for (int k = 0; k < ksteps; ++k) {
                if (k >= klim) continue;
                // +32 bytes per UMMA_K step inside the swizzle row: +2 in the (addr >> 4) field
                if (M256 && p.dbg_swap) {   // experiment: D^T[128 x 256 pixels] = W[128 x 16] * X[256 x 16]^T, one instruction per K step
                  umma_f16(tmem_base + acc * 256u, b_desc + (uint64_t)(2 * k), a_desc + (uint64_t)(2 * k), umma_idesc_f16(256u), accumulate);
                } else if (M256) {    // two row chains: rows 0..127 and 128..255 of the same chunk against the same weight tile
                  umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                  umma_f16(d_tmem + (uint32_t)p.block_n, a_desc + (uint64_t)((a_chain >> 4) + 2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                } else if (chain_mask != 0u) {       // K chains (opt-in): step ks accumulates into accumulator ks % kchains
                  umma_f16(d_tmem + (ks & chain_mask) * (uint32_t)p.block_n, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                           ks > chain_mask ? 1u : 0u);
                  ++ks;
                } else {
                  umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accumulate);
                }
                accumulate = 1u;
              }
            }
          }
          if (p.b_resident) bres_desc += (uint64_t)(((uint32_t)nch * b_chunk) >> 4);
          umma_commit(empty_bar(st));
          if (kit == p.num_k_iters - 1) umma_commit(tmem_full_bar(acc));
          if (++ring_st == (uint32_t)p.num_stages) { ring_st = 0; ring_ph ^= 1u; }
        }
      }
      if (p.prof) { prof_add(p.prof, PROF_MMA_WAIT_FULL, t_wf); prof_add(p.prof, PROF_MMA_WAIT_TMEM, t_wt); prof_add(p.prof, PROF_MMA_TOTAL, clock64() - t_b); }
    }
  } else if (warp < kEpilogueWarps) {
#define VCB_EPI_CASE(K, ACT, RES, F32) \
    case K: \
      if (!M256 && p.epi_split) conv_epilogue_split<ACT, RES, F32, false>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tmem_full_bar(0), tmem_empty_bar(0), blockIdx.x, gridDim.x, 0, bias_gen + p.cout_pad); \
      else conv_epilogue_fast<ACT, RES, F32, M256>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tmem_full_bar(0), tmem_empty_bar(0)); \
      break;
    if (A_MODE == A_TMA && !M256 && p.a_rowwin && !p.epi_split) {   // 2-D pixel blocks: the patch-mode staging / 4-D store with a dense lattice (Lp = Xs)
      if (p.epi_kind == 1)
        conv_epilogue_fast<VCB_ACT_SILU, VCB_RES_NONE, false, false, false, true>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tmem_full_bar(0), tmem_empty_bar(0));
      else
        conv_epilogue_fast<VCB_ACT_RELU, VCB_RES_NONE, false, false, false, true>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tmem_full_bar(0), tmem_empty_bar(0));
    } else
    switch (A_MODE == A_TMA ? p.epi_kind : 0) {
      VCB_EPI_CASE(1, VCB_ACT_SILU, VCB_RES_NONE, false)
      VCB_EPI_CASE(2, VCB_ACT_SILU, VCB_RES_AFTER_ACT, false)
      VCB_EPI_CASE(3, VCB_ACT_SILU_TANH, VCB_RES_NONE, false)
      VCB_EPI_CASE(4, VCB_ACT_SILU_TANH, VCB_RES_AFTER_ACT, false)
      VCB_EPI_CASE(5, VCB_ACT_RELU, VCB_RES_NONE, false)
      VCB_EPI_CASE(6, VCB_ACT_RELU, VCB_RES_BEFORE_ACT, false)
      VCB_EPI_CASE(7, VCB_ACT_NONE, VCB_RES_NONE, true)
      VCB_EPI_CASE(8, VCB_ACT_NONE, VCB_RES_NONE, false)
      default:
        if (!M256) conv_epilogue<false>(p, &tmap_out, tmem_base, out_stage, tmem_full_bar(0), tmem_empty_bar(0), blockIdx.x, gridDim.x, 0);
    }
#undef VCB_EPI_CASE
  } else if (A_MODE == A_TMA) {
    // ======================= optional second TMA producer (warp 10): weight tiles =======================
    if (!M256 && p.split_b && warp == kGatherWarp0 && lane == 0) {
      const uint32_t b_chunk = (uint32_t)(p.block_n * BK * 2);
      constexpr int G = kBlockK / BK;
      uint32_t ring_st = 0, ring_ph = 0, it = 0;
      for (int tile_seq = blockIdx.x; tile_seq < p.num_tiles; tile_seq += gridDim.x) {
        const int n_tile = (p.tile_rev ? p.num_tiles - 1 - tile_seq : tile_seq) % p.n_tiles;
        int kidx = 0;
        for (int kit = 0; kit < p.num_k_iters; ++kit, ++it) {
          const int st = (int)ring_st;
          const uint32_t ph = ring_ph;
          if (++ring_st == (uint32_t)p.num_stages) { ring_st = 0; ring_ph ^= 1u; }      // no integer division in the issue loops
          mbar_wait(empty_bar(st), ph ^ 1u, p.fault, FAULT_EMPTY_WAIT, 200 + st);
          const uint32_t b_dst = smem_base + (uint32_t)st * stage_bytes + kATileBytes;
          const int nch = (G == 1) ? 1 : min(G, p.total_chunks - kit * G);
          mbar_arrive_expect_tx(full_bar(st), (uint32_t)nch * b_chunk);
          for (int g = 0; g < nch; ++g, ++kidx)
            tma_load_2d(&tmap_b, full_bar(st), b_dst + (uint32_t)g * b_chunk, kidx * BK, n_tile * p.block_n);
        }
      }
    }
  } else {
    // ======================= gather producers (warps 10-13) =======================
    const int gtid = threadIdx.x - kGatherWarp0 * 32;
    uint32_t ring_st = 0, ring_ph = 0, it = 0;          // K-steps issued by this thread (same sequence in every gather thread)
    uint32_t arrived = 0;     // K-steps already signalled on their full barrier
    for (int tile_seq = blockIdx.x; tile_seq < p.num_tiles; tile_seq += gridDim.x) {
      const int m_tile = (p.tile_rev ? p.num_tiles - 1 - tile_seq : tile_seq) / p.n_tiles;
      asm volatile("bar.sync 1, 128;" ::: "memory");     // everyone is done reading the previous table
      {
        const int m = m_tile * kBlockM + gtid;
        RowInfo ri;
        if (m < p.M) {
          const int n = m / p.PQ;
          const int rem = m - n * p.PQ;
          const int pp = rem / p.Q, qq = rem - pp * p.Q;
          ri.img_base = n * p.H * p.W;
          ri.h0 = (short)(pp * p.stride - p.pad);
          ri.w0 = (short)(qq * p.stride - p.pad);
        } else {
          ri.img_base = 0;
          ri.h0 = (short)-16384;      // every tap fails the bounds test -> zero fill
          ri.w0 = (short)-16384;
        }
        rows[gtid] = ri;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int r = 0, s = 0, c = 0;
      for (int kit = 0; kit < p.num_k_iters; ++kit, ++it) {
        const int st = (int)ring_st;
        const uint32_t ph = ring_ph;
        if (++ring_st == (uint32_t)p.num_stages) { ring_st = 0; ring_ph ^= 1u; }
        mbar_wait(empty_bar(st), ph ^ 1u, p.fault, FAULT_EMPTY_WAIT, 100 + st);
        const uint32_t a_dst = smem_base + (uint32_t)st * stage_bytes;
        if (A_MODE == A_GATHER) {
          const int j = gtid & 7;                         // 16-byte chunk inside the 128-byte row
          const int cbase = c * kBlockK + j * 8;
          const bool c_ok = cbase < p.Cin;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (gtid >> 3) + 16 * i;
            const RowInfo ri = rows[rr];
            const int h = ri.h0 + r, w = ri.w0 + s;
            const bool ok = c_ok && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
            const __half* src = ok ? p.x + (size_t)(ri.img_base + h * p.W + w) * (size_t)p.cin_pitch + cbase : p.x;
            cp_async_16(a_dst + (uint32_t)rr * 128u + (uint32_t)((j ^ (rr & 7)) << 4), src, ok);
          }
          if (++c == p.chunks_per_tap) { c = 0; if (++s == p.kw) { s = 0; ++r; } }
        } else if (p.c4_wide) {   // A_C4, 16-byte granules: chunk j of the row = taps (2u, 2u+1), u = kit*8 + j
          const int j = gtid & 7;
          const int t0 = (kit * 8 + j) * 2;
          const int tr = t0 / p.kw, ts = t0 - tr * p.kw;      // ts is even (kw even), so both taps share the row
          const bool t_ok = t0 < p.kh * p.kw;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (gtid >> 3) + 16 * i;
            const RowInfo ri = rows[rr];
            const int h = ri.h0 + tr, w = ri.w0 + ts;         // w even and W even: the pixel pair is in or out together
            const bool ok = t_ok && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
            const __half* src = ok ? p.x + (size_t)(ri.img_base + h * p.W + w) * 4 : p.x;
            cp_async_16(a_dst + (uint32_t)rr * 128u + (uint32_t)((j ^ (rr & 7)) << 4), src, ok);
          }
        } else {   // A_C4: 16 taps x 4 channels per K-step, 8-byte granules
          const int u = gtid & 15;
          const int t = kit * 16 + u;
          const int tr = t / p.kw, ts = t - tr * p.kw;
          const bool t_ok = t < p.kh * p.kw;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int rr = (gtid >> 4) + 8 * i;
            const RowInfo ri = rows[rr];
            const int h = ri.h0 + tr, w = ri.w0 + ts;
            const bool ok = t_ok && (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
            const __half* src = ok ? p.x + (size_t)(ri.img_base + h * p.W + w) * 4 : p.x;
            cp_async_8(a_dst + (uint32_t)rr * 128u + (uint32_t)(((u >> 1) ^ (rr & 7)) << 4) + (uint32_t)((u & 1) << 3),
                       src, ok);
          }
        }
        cp_async_commit();
        if (it + 1 - arrived > (uint32_t)kGatherLag) {     // oldest outstanding group is complete
          cp_async_wait<kGatherLag>();
          fence_proxy_async_smem();
          mbar_arrive(full_bar(arrived % p.num_stages));
          ++arrived;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (; arrived < it; ++arrived) mbar_arrive(full_bar(arrived % p.num_stages));
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 0 && p.prof) prof_add(p.prof, PROF_CTA_TOTAL, clock64() - t_cta0);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster own one 256 x N tile.  Each CTA loads its own 128
// pixel rows of A and HALF of the weight tile (N/2 rows); one tcgen05.mma issued by the leader CTA spans both SMs
// (M = 256) and reads the B halves from both shared memories, so the L2 -> SM weight traffic and the number of MMA
// instructions per FLOP are halved.  TMA loads of both CTAs signal the leader's full barrier; the leader's
// tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; both epilogues release the
// accumulator on the leader's tmem_empty barrier.  A operand: im2col TMA only.
// ---------------------------------------------------------------------------------------------------------------
template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsTma, 2)
conv_umma_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (uint32_t)p.block_n * 64u;                 // this CTA's half: N/2 rows x 128 B
  const uint32_t stage_bytes = kATileBytes + b_tile_bytes;
  const uint32_t out_stage = smem_base + (uint32_t)p.num_stages * stage_bytes;
  const uint32_t bars = out_stage + (uint32_t)p.out_stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(kMaxStages + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kMaxStages + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* tail_gen = smem_gen + (size_t)p.num_stages * stage_bytes + p.out_stage_bytes + 8 * (2 * kMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail_gen);
  const uint32_t res_bar0 = tmem_slot + 16u;         // same tail layout as the single-CTA kernel: slot, (row table ->) 2 residual barriers, bias
  const uint32_t bias_smem = tmem_slot + 16u + (uint32_t)(kBlockM * sizeof(RowInfo));
  float* bias_gen = reinterpret_cast<float*>(tail_gen + 16 + kBlockM * sizeof(RowInfo));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.epi_kind != 0)
    for (int i = threadIdx.x; i < p.cout_pad; i += blockDim.x) bias_gen[i] = __ldg(p.bias + i);
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  griddep_launch_dependents();

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar(s), 2);                        // one producer arrival per CTA (leader's copy is the live one)
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), (uint32_t)p.epi_empty_count);   // both CTAs' epilogues (leader's copy is the live one)
    }
    mbar_init(res_bar0, 1);
    mbar_init(res_bar0 + 8u, 1);
    fence_mbar_init();
  }
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (p.epi_kind != 0 && p.res_mode != VCB_RES_NONE) tma_prefetch_desc(&tmap_res);
  }
  if (warp == kMmaWarp) {
    tmem_alloc_2cta(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish_2cta();
  }
  tcgen05_fence_before();
  cluster_sync_all();                                   // peers' barriers exist before any remote arrive / TMA signal
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_wait();

  if (warp == kProducerWarp) {
    if (lane == 0) {
      constexpr uint32_t a_chunk = (uint32_t)(kBlockM * BK * 2);
      const uint32_t b_chunk = (uint32_t)((p.block_n >> 1) * BK * 2);
      constexpr int G = kBlockK / BK;
      uint32_t ring_st = 0, ring_ph = 0, it = 0;
      for (int tile_seq = cluster_id; tile_seq < p.num_pair_tiles; tile_seq += num_clusters) {
        const int tile = p.tile_rev ? p.num_pair_tiles - 1 - tile_seq : tile_seq;
        const int pm = tile / p.n_tiles, n_tile = tile - pm * p.n_tiles;
        const int m0 = (2 * pm + rank) * kBlockM;      // may lie past M for the last pair: the TMA zero-fills
        const int cn = m0 / p.PQ;
        const int rem = m0 - cn * p.PQ;
        const int p0 = rem / p.Q, q0 = rem - p0 * p.Q;
        const int cw = q0 * p.stride - p.pad, ch = p0 * p.stride - p.pad;
        int r = 0, s = 0, c = 0, kidx = 0;
        for (int kit = 0; kit < p.num_k_iters; ++kit, ++it) {
          const int st = (int)ring_st;
          const uint32_t ph = ring_ph;
          if (++ring_st == (uint32_t)p.num_stages) { ring_st = 0; ring_ph ^= 1u; }      // no integer division in the issue loops
          mbar_wait(empty_bar(st), ph ^ 1u, p.fault, FAULT_EMPTY_WAIT, st);
          const uint32_t a_dst = smem_base + (uint32_t)st * stage_bytes;
          const uint32_t b_dst = a_dst + kATileBytes;
          const int nch = (G == 1) ? 1 : min(G, p.total_chunks - kit * G);
          const uint32_t lead_full = mapa_shared(full_bar(st), 0);
          if (leader) mbar_arrive_expect_tx(full_bar(st), 2u * (uint32_t)nch * (a_chunk + b_chunk));
          else mbar_arrive_remote(full_bar(st), 0);
          for (int g = 0; g < nch; ++g, ++kidx) {
            if (p.a_tiled) tma_load_2d_2cta(&tmap_a, lead_full, a_dst + (uint32_t)g * a_chunk, c * BK, m0, p.a_policy);
            else tma_load_im2col_4d_2cta(&tmap_a, lead_full, a_dst + (uint32_t)g * a_chunk, c * BK, cw, ch, cn, (uint16_t)s, (uint16_t)r, p.a_policy);
            tma_load_2d_2cta(&tmap_b, lead_full, b_dst + (uint32_t)g * b_chunk, kidx * BK,
                             n_tile * p.block_n + rank * (p.block_n >> 1), p.b_policy);
            if (++c == p.chunks_per_tap) { c = 0; if (++s == p.kw) { s = 0; ++r; } }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_f16((uint32_t)p.block_n, 256u);
      constexpr uint32_t a_chunk = (uint32_t)(kBlockM * BK * 2);
      const uint32_t b_chunk = (uint32_t)((p.block_n >> 1) * BK * 2);
      constexpr uint32_t sbo = (uint32_t)(8 * BK * 2);
      constexpr uint32_t layout_type = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);
      constexpr int ksteps = BK / 16;
      constexpr int G = kBlockK / BK;
      const uint64_t desc_hi = umma_desc_kmajor(0, sbo, layout_type);
      uint32_t ring_st = 0, ring_ph = 0, it = 0, tile_iter = 0;
      for (int tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters, ++tile_iter) {
        const uint32_t acc = (p.acc_stages == 2) ? (tile_iter & 1u) : 0u;
        const uint32_t acc_ph = (p.acc_stages == 2) ? ((tile_iter >> 1) & 1u) : (tile_iter & 1u);
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u, p.fault, FAULT_TMEM_EMPTY_WAIT, (int)acc);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.block_n;
        for (int kit = 0; kit < p.num_k_iters; ++kit, ++it) {
          const int st = (int)ring_st;
          const uint32_t ph = ring_ph;
          if (++ring_st == (uint32_t)p.num_stages) { ring_st = 0; ring_ph ^= 1u; }      // no integer division in the issue loops
          mbar_wait(full_bar(st), ph, p.fault, FAULT_FULL_WAIT, st);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_base + (uint32_t)st * stage_bytes;
          const int nch = (G == 1) ? 1 : min(G, p.total_chunks - kit * G);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (g < nch) {
              const uint64_t a_desc = desc_hi | (uint64_t)(((a_addr + (uint32_t)g * a_chunk) & 0x3FFFF) >> 4);
              const uint64_t b_desc = desc_hi | (uint64_t)(((a_addr + kATileBytes + (uint32_t)g * b_chunk) & 0x3FFFF) >> 4);
#pragma unroll
              for (int k = 0; k < ksteps; ++k)
                umma_f16_2cta(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kit | g | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2cta(empty_bar(st), (uint16_t)3);
          if (kit == p.num_k_iters - 1) umma_commit_2cta(tmem_full_bar(acc), (uint16_t)3);
        }
      }
    }
  } else if (warp < kEpilogueWarps) {
#define VCB_EPI_CASE(K, ACT, RES, F32) \
    case K: \
      if (p.epi_split) conv_epilogue_split<ACT, RES, F32, true>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tmem_full_bar(0), \
                                                                tmem_empty_bar(0), cluster_id, num_clusters, rank, bias_gen + p.cout_pad); \
      else conv_epilogue_fast<ACT, RES, F32, false, true>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tmem_full_bar(0), \
                                                          tmem_empty_bar(0), cluster_id, num_clusters, rank); \
      break;
    switch (p.epi_kind) {
      VCB_EPI_CASE(1, VCB_ACT_SILU, VCB_RES_NONE, false)
      VCB_EPI_CASE(2, VCB_ACT_SILU, VCB_RES_AFTER_ACT, false)
      VCB_EPI_CASE(3, VCB_ACT_SILU_TANH, VCB_RES_NONE, false)
      VCB_EPI_CASE(4, VCB_ACT_SILU_TANH, VCB_RES_AFTER_ACT, false)
      VCB_EPI_CASE(5, VCB_ACT_RELU, VCB_RES_NONE, false)
      VCB_EPI_CASE(6, VCB_ACT_RELU, VCB_RES_BEFORE_ACT, false)
      VCB_EPI_CASE(7, VCB_ACT_NONE, VCB_RES_NONE, true)
      VCB_EPI_CASE(8, VCB_ACT_NONE, VCB_RES_NONE, false)
      default:
        conv_epilogue<true>(p, &tmap_out, tmem_base, out_stage, tmem_full_bar(0), tmem_empty_bar(0), cluster_id, num_clusters, rank);
    }
#undef VCB_EPI_CASE
  }

  tcgen05_fence_before();
  cluster_sync_all();                                   // nobody exits (or frees TMEM) while the peer may still signal it
  if (warp == kMmaWarp) tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// Patch mode (3x3 / stride 1 / pad 1, TMA path): the input of an M tile is loaded ONCE per 64-channel chunk.
// An M tile is R output rows x Xs output columns of one image.  Its input patch, (R+2) x (Xs+2) pixels x 64 channels, arrives
// by one tiled 4-D TMA box (borders and image edges zero-filled by the TMA unit) and sits in shared memory as consecutive
// 128-byte rows, pixel (py, px) at row py*Lp + px, Lp = Xs+2.  Accumulator row m = i*Lp + j is output pixel (y0+i, x0+j); for
// filter tap (r, s) its operand row is m + r*Lp + s, so the tap's A operand is the SAME shared-memory patch read through a
// descriptor whose start address is shifted by (r*Lp + s) rows.  tcgen05.mma applies the 128-byte swizzle to absolute
// shared-memory address bits, so any 128-byte-aligned start is legal with base_offset = 0 (tools/umma_shift_probe.cu,
// profiles/r01_umma_shift_probe.txt).  Lattice columns j >= Xs (and rows i >= R) are junk: they are computed and dropped.
// Versus one im2col box per tap this cuts the L2 -> shared-memory traffic of A by 9 / ((1+2/R)(1+2/Xs)) (3.5-5x); the weight
// tiles stream through their own ring (own producer warp) or stay resident when the packed B fits.
// The lattice of a tile has up to 256 rows: rows 0-127 and 128-255 are two accumulators (two independent MMA chains -- dependent
// tcgen05.mma on one accumulator issue ~190 clk apart) that share every weight tile.
// Warps: 0-7 epilogue (0-3 chain 0, 4-7 chain 1), 8 patch producer, 9 MMA issuer (+ TMEM), 10 weight producer.  One CTA per SM.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kThreadsPatch = 448;      // warps 0-7 epilogue, 8 patch producer, 10 weight producer, 9/11/12/13 MMA issuers
constexpr int kPatchMaxA = 4, kPatchMaxB = 8;

__global__ void __launch_bounds__(kThreadsPatch, 1)
conv_patch_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t a_region = smem_base;
  const uint32_t b_region = a_region + (uint32_t)(p.a_stages * p.patch_stage_bytes);
  const uint32_t b_bytes = p.b_resident ? (uint32_t)p.b_res_bytes : (uint32_t)p.b_stages * b_tile_bytes;
  const uint32_t out_stage = b_region + b_bytes;
  const uint32_t bars = out_stage + (uint32_t)p.out_stage_bytes;
  auto afull = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto aempty = [&](int s) { return bars + 8u * (uint32_t)(kPatchMaxA + s); };
  auto bfull = [&](int s) { return bars + 8u * (uint32_t)(2 * kPatchMaxA + s); };
  auto bempty = [&](int s) { return bars + 8u * (uint32_t)(2 * kPatchMaxA + kPatchMaxB + s); };
  const uint32_t tfull0 = bars + 8u * (2 * kPatchMaxA + 2 * kPatchMaxB);
  const uint32_t tempty0 = tfull0 + 16u;
  const uint32_t tmem_slot = tempty0 + 16u;
  const uint32_t bres_bar = tmem_slot + 8u;
  const uint32_t res_bar0 = tmem_slot + 16u;
  const uint32_t bias_smem = tmem_slot + 32u;
  uint8_t* tail_gen = smem_raw + (tmem_slot - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tail_gen);
  float* bias_gen = reinterpret_cast<float*>(tail_gen + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.cout_pad; i += blockDim.x) bias_gen[i] = __ldg(p.bias + i);
  griddep_launch_dependents();
  const long long t_cta0 = prof_clock(p.prof);
  // One thread issues tcgen05.mma at most every ~107 clk whatever N is; the limit is per issuing thread, not per CTA or per
  // accumulator (tools/umma_rate_probe.cu, profiles/r01_umma_rate_probe.txt: 2 issuers 53 clk, 3-4 issuers ~40-48 clk per MMA per
  // SM).  So each (row chain, K chain) accumulator gets its OWN issuing warp: 2 issuers, 4 for N <= 64.  Every issuer waits on the
  // same full barriers and signals the same empty barriers (arrival count = number of issuers).
  const uint32_t num_issuers = 2u * (uint32_t)p.kchains;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPatchMaxA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), num_issuers); }
    for (int s = 0; s < kPatchMaxB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), num_issuers); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8u * a, num_issuers); mbar_init(tempty0 + 8u * a, kNumEpilogueThreads); }
    mbar_init(bres_bar, 1);
    mbar_init(res_bar0, 1);
    mbar_init(res_bar0 + 8u, 1);
    fence_mbar_init();
  }
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (p.res_mode != VCB_RES_NONE) tma_prefetch_desc(&tmap_res);
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_wait();

  const int chunks = p.chunks_per_tap;
  const int per_img = p.patch_ytiles * p.patch_xsegs;
  const uint32_t bc = (uint32_t)(p.block_n * kBlockK * 2);                     // one (tap, chunk) weight tile

  if (warp == kProducerWarp) {
    if (lane == 0) {         // ---- patch producer: one 4-D box per (tile, channel chunk)
      uint32_t sa = 0, pha = 0;
      for (int tile_seq = blockIdx.x; tile_seq < p.num_tiles; tile_seq += gridDim.x) {
        const int mt = (p.tile_rev ? p.num_tiles - 1 - tile_seq : tile_seq) / p.n_tiles;
        const int ni = mt / per_img, rem = mt - ni * per_img;
        const int yt = rem / p.patch_xsegs, xs = rem - yt * p.patch_xsegs;
        for (int c = 0; c < chunks; ++c) {
          const long long tw = prof_clock(p.prof);
          mbar_wait(aempty(sa), pha ^ 1u, p.fault, FAULT_EMPTY_WAIT, 600 + (int)sa);
          prof_add(p.prof, PROF_PROD_WAIT_EMPTY, prof_clock(p.prof) - tw);
          mbar_arrive_expect_tx(afull(sa), (uint32_t)p.patch_box_bytes);
          tma_load_4d(&tmap_a, afull(sa), a_region + sa * (uint32_t)p.patch_stage_bytes, c * kBlockK, xs * p.patch_Xs - 1,
                      yt * p.patch_R - 1, ni, p.a_policy);
          if (++sa == (uint32_t)p.a_stages) { sa = 0; pha ^= 1u; }
        }
      }
    }
  } else if (warp == kGatherWarp0) {
    if (lane == 0) {         // ---- weight producer
      if (p.b_resident) {
        if ((int)blockIdx.x < p.num_tiles) {
          mbar_arrive_expect_tx(bres_bar, (uint32_t)p.total_chunks * bc);
          for (int kc = 0; kc < p.total_chunks; ++kc) tma_load_2d(&tmap_b, bres_bar, b_region + (uint32_t)kc * bc, kc * kBlockK, 0, p.b_policy);
        }
      } else {
        uint32_t sb = 0, phb = 0;
        for (int tile_seq = blockIdx.x; tile_seq < p.num_tiles; tile_seq += gridDim.x) {
          const int nt = (p.tile_rev ? p.num_tiles - 1 - tile_seq : tile_seq) % p.n_tiles;
          for (int c = 0; c < chunks; ++c)
            for (int t = 0; t < 9; ++t) {
              mbar_wait(bempty(sb), phb ^ 1u, p.fault, FAULT_EMPTY_WAIT, 620 + (int)sb);
              mbar_arrive_expect_tx(bfull(sb), b_tile_bytes);
              tma_load_2d(&tmap_b, bfull(sb), b_region + sb * b_tile_bytes, (t * chunks + c) * kBlockK, nt * p.block_n, p.b_policy);
              if (++sb == (uint32_t)p.b_stages) { sb = 0; phb ^= 1u; }
            }
        }
      }
    }
  } else if (warp == kMmaWarp || warp >= 11) {
    const uint32_t issuer = warp == kMmaWarp ? 0u : (uint32_t)(warp - 10);       // warps 9, 11, 12, 13 -> issuers 0..3
    if (lane == 0 && issuer < num_issuers) {         // ---- MMA issuer of accumulator (K chain kq, row chain h)
      const uint32_t h = issuer & 1u, kq = issuer >> 1;
      const uint32_t idesc = umma_idesc_f16((uint32_t)p.block_n);
      const uint64_t desc_hi = umma_desc_kmajor(0, 1024u, 2u);      // SWIZZLE_128B, 8-row groups 1024 B apart, base_offset 0
      // The issue loop is ONE thread: every integer instruction in it delays the next tcgen05.mma (the first version spent ~8k
      // cycles per tile on ring-index divisions and per-tap address arithmetic).  Everything loop-invariant is hoisted: the nine
      // tap offsets in descriptor units, ring positions as counters, descriptors advanced by additions.
      uint32_t tap_off[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_off[t] = ((uint32_t)((t / 3) * p.patch_Lp + (t % 3)) * 128u) >> 4;
      const uint32_t bc16 = bc >> 4;
      const uint32_t kstep0 = 2u * kq, kstep_inc = 2u * (uint32_t)p.kchains;      // descriptor offset of this issuer's first K step / stride
      const uint32_t ksteps_mine = 4u / (uint32_t)p.kchains;
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tile_iter = 0;
      if (p.b_resident && (int)blockIdx.x < p.num_tiles) mbar_wait(bres_bar, 0u, p.fault, FAULT_FULL_WAIT, 640);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
        const uint32_t acc = (p.acc_stages == 2) ? (tile_iter & 1u) : 0u;
        const uint32_t acc_ph = (p.acc_stages == 2) ? ((tile_iter >> 1) & 1u) : (tile_iter & 1u);
        const long long tw1 = prof_clock(issuer == 0 ? p.prof : nullptr);
        mbar_wait(tempty0 + 8u * acc, acc_ph ^ 1u, p.fault, FAULT_TMEM_EMPTY_WAIT, (int)acc);
        if (issuer == 0) prof_add(p.prof, PROF_MMA_WAIT_TMEM, prof_clock(p.prof) - tw1);
        tcgen05_fence_after();
        // accumulators of a stage: [K chain][row chain: lattice rows 0-127 | 128-255][block_n]
        const uint32_t d_tmem = tmem_base + acc * 2u * (uint32_t)(p.block_n * p.kchains) + (kq * 2u + h) * (uint32_t)p.block_n;
        uint32_t accumulate = 0u;
        uint64_t bres_desc = desc_hi | (uint64_t)((b_region & 0x3FFFF) >> 4);      // resident weights: (tap, chunk) tiles, chunk c at +c*bc16
        for (int c = 0; c < chunks; ++c) {
          const long long tw2 = prof_clock(issuer == 0 ? p.prof : nullptr);
          mbar_wait(afull(sa), pha, p.fault, FAULT_FULL_WAIT, 650 + (int)sa);
          if (issuer == 0) prof_add(p.prof, PROF_MMA_WAIT_FULL, prof_clock(p.prof) - tw2);
          tcgen05_fence_after();
          // row chain 1 = lattice rows 128..255 of the patch
          const uint64_t a_desc0 = (desc_hi | (uint64_t)(((a_region + sa * (uint32_t)p.patch_stage_bytes + h * 16384u) & 0x3FFFF) >> 4)) + kstep0;
          uint64_t b_desc_res = bres_desc + (uint64_t)((uint32_t)c * bc16) + kstep0;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            uint64_t b_desc;
            if (p.b_resident) {
              b_desc = b_desc_res;
              b_desc_res += (uint64_t)((uint32_t)chunks * bc16);
            } else {
              mbar_wait(bfull(sb), phb, p.fault, FAULT_FULL_WAIT, 660 + (int)sb);
              tcgen05_fence_after();
              b_desc = (desc_hi | (uint64_t)(((b_region + sb * b_tile_bytes) & 0x3FFFF) >> 4)) + kstep0;
            }
            const uint64_t a_desc = a_desc0 + tap_off[t];
            for (uint32_t k = 0; k < ksteps_mine; ++k) {      // this issuer's K steps of the 64-wide chunk
              umma_f16(d_tmem, a_desc + (uint64_t)(k * kstep_inc), b_desc + (uint64_t)(k * kstep_inc), idesc, accumulate);
              accumulate = 1u;
            }
            if (!p.b_resident) {
              umma_commit(bempty(sb));
              if (++sb == (uint32_t)p.b_stages) { sb = 0; phb ^= 1u; }
            }
          }
          umma_commit(aempty(sa));
          if (++sa == (uint32_t)p.a_stages) { sa = 0; pha ^= 1u; }
        }
        umma_commit(tfull0 + 8u * acc);
      }
    }
  } else if (warp < kEpilogueWarps) {
#define VCB_EPI_CASE(K, ACT, RES, F32) \
    case K: conv_epilogue_fast<ACT, RES, F32, true, false, true>(p, &tmap_out, &tmap_res, tmem_base, out_stage, bias_smem, res_bar0, tfull0, tempty0); break;
    switch (p.epi_kind) {
      VCB_EPI_CASE(1, VCB_ACT_SILU, VCB_RES_NONE, false)
      VCB_EPI_CASE(2, VCB_ACT_SILU, VCB_RES_AFTER_ACT, false)
      VCB_EPI_CASE(3, VCB_ACT_SILU_TANH, VCB_RES_NONE, false)
      VCB_EPI_CASE(4, VCB_ACT_SILU_TANH, VCB_RES_AFTER_ACT, false)
      VCB_EPI_CASE(5, VCB_ACT_RELU, VCB_RES_NONE, false)
      VCB_EPI_CASE(6, VCB_ACT_RELU, VCB_RES_BEFORE_ACT, false)
      VCB_EPI_CASE(7, VCB_ACT_NONE, VCB_RES_NONE, true)
      VCB_EPI_CASE(8, VCB_ACT_NONE, VCB_RES_NONE, false)
      default: break;
    }
#undef VCB_EPI_CASE
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 0 && p.prof) { prof_add(p.prof, PROF_CTA_TOTAL, clock64() - t_cta0); prof_add(p.prof, PROF_CTAS, 1); }
}

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 (BN folded) -> [cout_pad][K_pad] fp16, K index = (tap * cin_pad + c) for
// the 64-channel-chunk modes and (tap * 4 + c) for A_C4
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias, __half* __restrict__ wp,
                                    float* __restrict__ bp, int cout, int cin, int kh, int kw, int cout_pad, int k_pad,
                                    int cin_pad, int c4, int rowwin) {
  const long long total = (long long)cout_pad * k_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i / k_pad), k = (int)(i - (long long)o * k_pad);
    int tap, c;
    if (c4) { tap = k >> 2; c = k & 3; }
    else if (rowwin) {        // k = r * 64 + s * 16 + c, s = 3 is the zero pixel that completes the 128-byte row
      const int r = k >> 6, sx = (k >> 4) & 3;
      c = k & 15;
      tap = sx < 3 ? r * 3 + sx : kh * kw;
    } else { tap = k / cin_pad; c = k - tap * cin_pad; }
    float v = 0.0f;
    if (o < cout && c < cin && tap < kh * kw) {
      const int r = tap / kw, s = tap - r * kw;
      v = w[(((size_t)o * cin + c) * kh + r) * kw + s];
    }
    wp[i] = __float2half_rn(v);
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < cout_pad; o += gridDim.x * blockDim.x)
    bp[o] = (bias != nullptr && o < cout) ? bias[o] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ConvGeom {
  int a_mode;   // A_TMA / A_GATHER / A_C4
  int P, Q, M;
  int cin_pad, k_pad, num_k_iters, chunks_per_tap;
  int bk, chunks_per_stage, total_chunks;
  int block_n, n_tiles, cout_pad, m_tiles;
  int stages, acc_stages, tmem_cols;
  int out_bufs;         // staged output sub-tiles in flight (1 or 2)
  int ctas_per_sm;      // 1, or 2 co-resident persistent CTAs for narrow-N layers (each <= 110 KiB smem, <= 256 TMEM columns)
  int two_cta;          // CTA-pair kernel (cta_group::2)
  int m256;             // 256-row CTA tiles, two accumulator chains, one CTA per SM
  int tile_m;           // 128 or 256
  int epi_kind;         // specialised epilogue instance (0 = generic)
  int kchains;          // K-split accumulation chains per tile (1, 2, 4)
  int patch;            // patch mode (conv_patch_kernel)
  int patch_R, patch_Xs, patch_Lp, patch_xsegs, patch_ytiles, patch_stage_bytes, a_stages, b_stages;
  int b_resident, b_res_bytes;
  int rowwin;           // VCB_A_ROWWIN (patch_R x patch_Xs pixel blocks, dense lattice)
  size_t smem_bytes;
};

static int conv_geometry(const VcbConvDesc& d_in, ConvGeom& g) {
  VcbConvDesc d = d_in;
  const bool want_stats = (d_in.reserved[1] & 0x200) != 0;     // internal (vcb_conv2d_fwd_stats): keep to the kernels with the split epilogue
  d.reserved[1] &= 0xff;      // bit 8 is a per-launch flag that does not change the geometry (conv2d_fwd)
  if (d.n <= 0 || d.h <= 0 || d.w <= 0 || d.cin <= 0 || d.cout <= 0 || d.kh <= 0 || d.kw <= 0 || d.stride <= 0 ||
      d.pad < 0)
    return set_error(VCB_ERR_INVALID, "conv: non-positive dimension");
  g.P = (d.h + 2 * d.pad - d.kh) / d.stride + 1;
  g.Q = (d.w + 2 * d.pad - d.kw) / d.stride + 1;
  if (g.P <= 0 || g.Q <= 0) return set_error(VCB_ERR_INVALID, "conv: empty output");
  const long long M = (long long)d.n * g.P * g.Q;
  if (M > 0x7fffff00LL || (long long)d.n * d.h * d.w > 0x7fffff00LL) return set_error(VCB_ERR_INVALID, "conv: too many pixels");
  g.M = (int)M;
  int mode = d.a_mode;
  if (mode == VCB_A_AUTO) mode = (d.cin_pitch == 4 && d.cin <= 4) ? VCB_A_C4 : VCB_A_IM2COL_TMA;
  g.rowwin = 0;
  if (mode == VCB_A_ROWWIN) {
    if (d.cin > 16 || d.cin_pitch != 16 || d.kh != 3 || d.kw != 3 || d.stride != 1 || d.pad != 1 || d.cout > 256 ||
        d.res_mode != VCB_RES_NONE || d.out_dtype != VCB_F16 || (d.act != VCB_ACT_SILU && d.act != VCB_ACT_RELU) || d.reserved[0] != 0 ||
        (d.reserved[3] != 0 && d.reserved[3] != 1) || d.reserved[2] != 0)
      return set_error(VCB_ERR_INVALID, "conv: VCB_A_ROWWIN needs a 3x3/s1/p1 layer over cin_pitch = 16 (W-padded input), SiLU or ReLU, fp16 out, no residual");
    g.a_mode = A_TMA;
    g.rowwin = 1;
    g.bk = kBlockK;
    g.chunks_per_tap = 1;
    g.cin_pad = 16;
    g.total_chunks = 3;                                  // filter rows
    // pixel block R x Xs = 128 with the fewest junk pixels
    double best = -1.0;
    for (int xs = 128; xs >= 8; xs >>= 1) {
      const int r = 128 / xs;
      const int nseg = (d.w + xs - 1) / xs, yt = (d.h + r - 1) / r;
      const double util = (double)d.h * d.w / ((double)nseg * yt * 128.0);
      if (util > best + 1e-9) { best = util; g.patch_R = r; g.patch_Xs = xs; g.patch_xsegs = nseg; g.patch_ytiles = yt; }
    }
    g.patch_Lp = g.patch_Xs;
  } else if (mode == VCB_A_C4) {
    if (d.cin_pitch != 4 || d.cin > 4) return set_error(VCB_ERR_INVALID, "conv: A_C4 needs cin<=4 and cin_pitch==4");
    g.a_mode = A_C4;
    g.bk = kBlockK;
    g.cin_pad = 4;
    g.chunks_per_tap = 0;
    g.total_chunks = (d.kh * d.kw + 15) / 16;          // 16 taps x 4 channels per K-step
  } else {
    if (d.cin % 8 != 0 || d.cin_pitch % 8 != 0 || d.cin_pitch < d.cin)
      return set_error(VCB_ERR_INVALID, "conv: cin and cin_pitch must be multiples of 8 (cin_pitch >= cin)");
    g.a_mode = (mode == VCB_A_GATHER) ? A_GATHER : A_TMA;
    // K chunk = one TMA box of bk channels of one tap; the widest swizzle that divides cin avoids zero K padding
    g.bk = kBlockK;
    if (g.a_mode == A_TMA) {
      if (d.reserved[2] == 16 || d.reserved[2] == 32 || d.reserved[2] == 64) g.bk = d.reserved[2];
      else g.bk = (d.cin % 64 == 0) ? 64 : ((d.cin % 32 == 0) ? 32 : (d.cin <= 16 ? 16 : 64));   // measured: 16-wide chunks only pay for stems
    }
    g.chunks_per_tap = (d.cin + g.bk - 1) / g.bk;
    g.cin_pad = g.chunks_per_tap * g.bk;
    g.total_chunks = d.kh * d.kw * g.chunks_per_tap;
    if (g.a_mode == A_TMA && (d.pad > 127 || d.kh > 128 || d.kw > 128 || d.stride > 8))
      return set_error(VCB_ERR_INVALID, "conv: geometry outside the im2col TMA limits");
  }
  g.chunks_per_stage = kBlockK / g.bk;
  g.num_k_iters = (g.total_chunks + g.chunks_per_stage - 1) / g.chunks_per_stage;
  if (d.h > 16000 || d.w > 16000) return set_error(VCB_ERR_INVALID, "conv: image too large");
  g.k_pad = g.total_chunks * g.bk;
  const int cout16 = (d.cout + 15) / 16 * 16;
  if (d.block_n != 0) {
    if (d.block_n % 16 != 0 || d.block_n < 16 || d.block_n > 256) return set_error(VCB_ERR_INVALID, "conv: bad block_n");
    g.block_n = d.block_n;
    g.n_tiles = (cout16 + g.block_n - 1) / g.block_n;
  } else {
    g.n_tiles = (cout16 + 255) / 256;
    g.block_n = ((cout16 + g.n_tiles - 1) / g.n_tiles + 15) / 16 * 16;
    // with several N tiles every staged 64-column sub-tile must belong to one tile only
    if (g.n_tiles > 1) g.block_n = (g.block_n + 63) / 64 * 64;
  }
  if (g.n_tiles > 1 && g.block_n % 64 != 0) return set_error(VCB_ERR_INVALID, "conv: block_n must be a multiple of 64 when cout spans several N tiles");
  g.cout_pad = g.n_tiles * g.block_n;
  g.m_tiles = g.rowwin ? d.n * g.patch_ytiles * g.patch_xsegs : (g.M + kBlockM - 1) / kBlockM;
  const int cout_store = (d.cout + 7) / 8 * 8;
  if (d.cout_pitch < cout_store || d.cout_pitch % 8 != 0)
    return set_error(VCB_ERR_INVALID, "conv: cout_pitch must be a multiple of 8 and >= round_up(cout, 8)");
  if (d.res_mode != VCB_RES_NONE && (d.res_pitch % 8 != 0 || d.res_pitch < cout_store))
    return set_error(VCB_ERR_INVALID, "conv: bad residual pitch");
  // CTA pairs (cta_group::2): measured on B200 at parity with the single-CTA kernel for these layer shapes, so they are
  // opt-in (reserved[3] == 2); reserved[3] == 1 forces the single-CTA kernel.
  // CTA pairs (cta_group::2, each CTA loads half of every weight tile), two co-resident clusters per SM pair: measured 3-9 %
  // faster than two independent 128-row CTAs per SM on 3x3 layers with N >= 128 and at least two waves of pair tiles, at
  // parity elsewhere (profiles/r01_layer_modes.md).  reserved[3]: 1 = never, 2 = pair kernel with one cluster per SM pair,
  // 4 = pair kernel with two.
  const int epi0 = (d.reserved[0] == 0) ? epi_kind_of(d.act, d.res_mode, d.out_dtype == VCB_F32 ? 1 : 0) : 0;
  const bool auto_pair = d.reserved[3] == 0 && g.a_mode == A_TMA && !g.rowwin && epi0 != 0 && d.kh * d.kw > 1 && g.block_n >= 128 &&
                         (long long)((g.m_tiles + 1) / 2) * g.n_tiles >= 2 * 148;
  g.two_cta = (g.a_mode == A_TMA && g.m_tiles >= 2 && (d.reserved[3] == 2 || d.reserved[3] == 4 || auto_pair)) ? 1 : 0;
  if ((d.reserved[3] == 2 || d.reserved[3] == 4) && !g.two_cta) return set_error(VCB_ERR_INVALID, "conv: the CTA-pair kernel needs the TMA path and >= 2 M tiles");
  // specialised epilogue: single-CTA TMA kernel with the staged TMA store (any debug value in reserved[0] forces the generic one)
  g.epi_kind = (g.a_mode == A_TMA && d.reserved[0] == 0) ? epi_kind_of(d.act, d.res_mode, d.out_dtype == VCB_F32 ? 1 : 0) : 0;
  const size_t b_total = (size_t)g.total_chunks * g.block_n * g.bk * 2;
  const size_t tail0 = 1024 /*align slack*/ + 8 * (2 * kMaxStages + 4) + 16 + kBlockM * sizeof(RowInfo) + (size_t)g.cout_pad * 4 + 1024 /*BN-statistics scratch*/ + 64;
  g.m256 = 0; g.tile_m = kBlockM; g.b_resident = 0; g.b_res_bytes = 0; g.kchains = 1; g.patch = 0;
  int chosen = 0;
  // ---- patch mode: 3x3/s1/p1 with 64-channel chunks; reserved[3] == 5 forces it, == 1 (or any other forced mode) forbids it
  {
    const bool can = g.a_mode == A_TMA && !g.rowwin && !want_stats && !g.two_cta && g.epi_kind != 0 && d.kh == 3 && d.kw == 3 && d.stride == 1 && d.pad == 1 && g.bk == 64;
    // automatic for narrow layers (N <= 64) with at least two waves of tiles: measured 4-14 % faster than the 128-row im2col mode
    // (ReID layer 1, YOLOv5m 48->48, YOLOv5s 64->64; profiles/r01_layer_modes.md); wider layers lose (single CTA per SM)
    bool want = can && (d.reserved[3] == 5);
    if (can && d.reserved[3] == 0 && g.block_n <= 64 && g.n_tiles == 1) {
      int best_tiles = 0;
      double best_u = 0.0;
      for (int nseg = 1; nseg <= 16; ++nseg) {
        const int xs = (d.w + nseg - 1) / nseg;
        if (xs + 2 > 128) continue;
        int r = 256 / (xs + 2);
        if (r > d.h) r = d.h;
        if (r < 1) continue;
        const int yt = (d.h + r - 1) / r;
        const double util = (double)d.h * d.w / ((double)yt * nseg * 256.0);
        if (util > best_u + 1e-9) { best_u = util; best_tiles = yt * nseg; }
      }
      if (best_u >= 0.75 && (long long)best_tiles * d.n >= 2 * 148) want = true;
    }
    if (d.reserved[3] == 5 && !can) return set_error(VCB_ERR_INVALID, "conv: patch mode needs a 3x3/s1/p1 layer on the TMA path with 64-channel chunks");
    if (want) {
      // tile shape: R rows x Xs columns with R * (Xs + 2) <= 128, maximising useful rows per 128-row tile over the image
      double best = 0.0;
      for (int nseg = 1; nseg <= 16; ++nseg) {
        const int xs = (d.w + nseg - 1) / nseg;
        if (xs + 2 > 128) continue;
        int r = 256 / (xs + 2);
        if (r > d.h) r = d.h;
        if (r < 1) continue;
        const int yt = (d.h + r - 1) / r;
        const double util = (double)d.h * d.w / ((double)yt * nseg * 256.0);
        if (util > best + 1e-9) { best = util; g.patch_R = r; g.patch_Xs = xs; g.patch_xsegs = nseg; g.patch_ytiles = yt; }
      }
      g.patch_Lp = g.patch_Xs + 2;
      g.patch_stage_bytes = ((256 + 2 * g.patch_Lp + 2) * 128 + 1023) / 1024 * 1024;
      const size_t budget = 227 * 1024;
      const size_t tailp = 1024 + 8 * (2 * 4 + 2 * 8) + 32 + 16 + 16 + (size_t)g.cout_pad * 4 + 64;
      for (int res = 1; res >= 0 && !chosen; --res) {
        if (res && !(g.n_tiles == 1 && b_total <= 112 * 1024)) continue;
        const size_t b_res = res ? (b_total + 1023) / 1024 * 1024 : 0;
        for (int as = 3; as >= 2 && !chosen; --as)          // a third patch stage before a second staging buffer
          for (int bufs = 2; bufs >= 1 && !chosen; --bufs) {
            const size_t fixed = tailp + b_res + (size_t)bufs * 2 * kStageOutBytes + (size_t)as * g.patch_stage_bytes;
            if (fixed > budget) continue;
            int bs = res ? 0 : (int)((budget - fixed) / ((size_t)g.block_n * 128));
            if (bs > 8) bs = 8;
            if (!res && bs < 3) continue;
            g.patch = 1; g.a_stages = as; g.b_stages = bs; g.b_resident = res; g.b_res_bytes = (int)b_res; g.out_bufs = bufs;
            // narrow N: the two row chains do not cover the ~190 clk dependent-issue interval -> split K over two more accumulators
            g.kchains = (g.block_n <= 64 && d.reserved[1] != 3) ? 2 : 1;
            g.ctas_per_sm = 1; g.acc_stages = (4 * g.kchains * g.block_n <= 512) ? 2 : 1;
            g.tile_m = 2 * kBlockM;
            int pow2 = 32;
            while (pow2 < g.acc_stages * 2 * g.kchains * g.block_n) pow2 <<= 1;
            g.tmem_cols = pow2; g.stages = as;
            g.smem_bytes = fixed + (size_t)bs * g.block_n * 128;
            chosen = 1;
          }
      }
      if (!chosen && d.reserved[3] == 5) return set_error(VCB_ERR_INVALID, "conv: patch mode does not fit in shared memory");
    }
  }
  // ---- 256-row tiles (two accumulator chains, one CTA per SM): halves the weight bytes each SM pulls from L2 per output.
  // Needs the specialised epilogue, N <= 192 (three 56 KiB stages must fit) and at least two waves of 256-row tiles
  // (reserved[3] == 3 forces it, == 1 forbids it).
  {
    const bool want = g.a_mode == A_TMA && !g.two_cta && g.epi_kind != 0 && g.block_n <= 192 && d.reserved[3] != 1 &&
                      d.reserved[3] == 3;     // opt-in: measured at or below the two-CTA 128-row mode on every layer shape
    if (d.reserved[3] == 3 && !want) return set_error(VCB_ERR_INVALID, "conv: 256-row tiles need the TMA path, a specialised epilogue and N <= 192");
    if (want && !chosen) {
      const size_t budget = 227 * 1024;
      const size_t stage_out = 2 * (size_t)kStageOutBytes;                  // 256 rows x 128 B
      for (int res = 1; res >= 0 && !chosen; --res) {
        // resident weights (one N tile): the packed B is loaded once per CTA and no weight tile travels with the stages
        if (res && !(g.n_tiles == 1 && b_total <= 112 * 1024)) continue;
        const size_t b_res = res ? (b_total + 1023) / 1024 * 1024 : 0;
        const size_t stage_bytes = 2 * (size_t)kATileBytes + (res ? 0 : (size_t)g.block_n * 128);
        for (int bufs = 2; bufs >= 1 && !chosen; --bufs) {
          const size_t fixed = tail0 + b_res + (size_t)bufs * stage_out;
          if (budget < fixed + 3 * stage_bytes) continue;
          int stages = (int)((budget - fixed) / stage_bytes);
          if (stages > kMaxStages) stages = kMaxStages;
          if (d.stages != 0 && d.stages < stages) stages = d.stages;
          if (stages < 2) continue;
          if (bufs == 2 && stages < 4 && budget >= tail0 + b_res + stage_out + (size_t)(stages + 1) * stage_bytes) continue;   // prefer one more stage
          const int acc = (4 * g.block_n <= 512) ? 2 : 1;
          int pow2 = 32;
          while (pow2 < acc * 2 * g.block_n) pow2 <<= 1;
          g.m256 = 1; g.tile_m = 2 * kBlockM; g.ctas_per_sm = 1; g.acc_stages = acc; g.tmem_cols = pow2; g.stages = stages; g.out_bufs = bufs;
          g.b_resident = res; g.b_res_bytes = (int)b_res;
          g.smem_bytes = fixed + (size_t)stages * stage_bytes;
          chosen = 1;
        }
      }
    }
  }
  if (!chosen) {
  // resident weights: one N tile and a packed B of at most 48 KiB stay in shared memory for the whole kernel
  // (measured: pays off together with two CTAs per SM because it shrinks the per-stage footprint to the A tile)
  g.b_resident = (g.a_mode == A_TMA && !g.two_cta && g.n_tiles == 1 && b_total <= 48 * 1024 && g.m_tiles > 148 && d.reserved[0] != 5) ? 1 : 0;
  g.b_res_bytes = g.b_resident ? (int)((b_total + 1023) / 1024 * 1024) : 0;
  const size_t stage_bytes = (size_t)kATileBytes + (g.b_resident ? 0 : (size_t)g.block_n * (g.two_cta ? 64 : 128));
  const size_t tail = tail0 + g.b_res_bytes;
  const int min_stages = (g.a_mode == A_TMA) ? 2 : kGatherLag + 2;
  // Two co-resident persistent CTAs per SM give the tensor pipe two independent accumulator chains (half the shared
  // memory and half the TMEM each); fall back to 1.
  const bool allow2 = g.a_mode == A_TMA && (!g.two_cta || d.reserved[1] == 5 || d.reserved[3] == 4 || auto_pair) && d.reserved[0] != 7 && (long long)g.m_tiles * g.n_tiles > 2 * 148;
  for (int ctas = allow2 ? 2 : 1; ctas >= 1 && !chosen; --ctas) {
    const size_t budget = (size_t)(227 * 1024) / ctas;
    const int tmem_budget = 512 / ctas;
    if (g.block_n > tmem_budget) continue;
    const int acc = (2 * g.block_n <= tmem_budget) ? 2 : 1;
    int pow2 = 32;
    while (pow2 < acc * g.block_n) pow2 <<= 1;
    for (int bufs = 2; bufs >= 1 && !chosen; --bufs) {
      const size_t fixed = tail + (size_t)bufs * kStageOutBytes;
      if (budget < fixed + (size_t)min_stages * stage_bytes) continue;
      int stages = (int)((budget - fixed) / stage_bytes);
      if (stages > kMaxStages) stages = kMaxStages;
      if (d.stages != 0 && d.stages < stages) stages = d.stages;
      if (stages < min_stages) continue;
      if (ctas == 2 && bufs == 2 && stages < 3) continue;     // prefer a third stage over a second staging buffer
      g.ctas_per_sm = ctas; g.acc_stages = acc; g.tmem_cols = pow2; g.stages = stages; g.out_bufs = bufs;
      g.smem_bytes = fixed + (size_t)stages * stage_bytes;
      // K chains: one tcgen05.mma of this N occupies the tensor pipe for block_n/2 clk but dependent ones issue ~200 clk apart,
      // so the SM wants ~400/block_n independent accumulators; they share this CTA's TMEM budget (single-buffered if need be)
      g.kchains = 1;
      if (g.epi_kind != 0 && !g.two_cta && d.reserved[1] == 4) {     // opt-in (reserved[1] == 4): measured slower than one chain on every shape
        const int total_ksteps = g.total_chunks * (g.bk / 16);
        int want = (400 + g.block_n - 1) / g.block_n;            // chains per SM
        want = (want + ctas - 1) / ctas;                         // per CTA
        int kc = want >= 3 ? 4 : (want >= 2 ? 2 : 1);
        while (kc > 1 && (kc * g.block_n > tmem_budget || total_ksteps < 2 * kc)) kc >>= 1;
        if (kc > 1) {
          g.kchains = kc;
          g.acc_stages = (2 * kc * g.block_n <= tmem_budget) ? 2 : 1;
          int p2 = 32;
          while (p2 < g.acc_stages * kc * g.block_n) p2 <<= 1;
          g.tmem_cols = p2;
        }
      }
      chosen = 1;
    }
  }
  }
  if (!chosen) return set_error(VCB_ERR_INVALID, "conv: not enough shared memory for the pipeline");
  return VCB_OK;
}

int conv_packed_sizes(const VcbConvDesc& d, int64_t* weight_halfs, int64_t* bias_floats) {
  ConvGeom g;
  const int rc = conv_geometry(d, g);
  if (rc != VCB_OK) return rc;
  if (weight_halfs) *weight_halfs = (int64_t)g.cout_pad * g.k_pad;
  if (bias_floats) *bias_floats = g.cout_pad;
  return VCB_OK;
}

int conv_out_hw(const VcbConvDesc& d, int32_t* ho, int32_t* wo) {
  ConvGeom g;
  const int rc = conv_geometry(d, g);
  if (rc != VCB_OK) return rc;
  if (ho) *ho = g.P;
  if (wo) *wo = g.Q;
  return VCB_OK;
}

int conv_pack_weights(const VcbConvDesc& d, const float* w, const float* bias, void* wp, float* bp, cudaStream_t st) {
  ConvGeom g;
  const int rc = conv_geometry(d, g);
  if (rc != VCB_OK) return rc;
  const long long total = (long long)g.cout_pad * g.k_pad;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  pack_weights_kernel<<<blocks, 256, 0, st>>>(w, bias, reinterpret_cast<__half*>(wp), bp, d.cout, d.cin, d.kh, d.kw,
                                              g.cout_pad, g.k_pad, g.cin_pad, g.a_mode == A_C4 ? 1 : 0, g.rowwin);
  return check_cuda(cudaGetLastError(), "pack_weights launch");
}

// Launch configuration shared by the conv kernels; with state().pdl the launch carries the programmatic stream
// serialisation attribute (the kernels call griddepcontrol.wait before they touch global memory).
static void fill_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int block, size_t smem, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (state().pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
  }
}

template <int BK>
static int launch_conv_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr, const ConvParams& p,
                            const ConvGeom& g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(conv_umma_2cta_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv 2cta)");
    attr_set = true;
  }
  const int max_clusters = state().num_sms / 2 * g.ctas_per_sm;
  const int clusters = p.num_pair_tiles < max_clusters ? p.num_pair_tiles : max_clusters;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_launch_config(cfg, attr, 2 * clusters, kThreadsTma, g.smem_bytes, st);
  return check_cuda(cudaLaunchKernelEx(&cfg, conv_umma_2cta_kernel<BK>, ta, tb, to, tr, p), "conv (cta pair) launch");
}

template <int A_MODE, int BK, bool M256 = false>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr, const ConvParams& p,
                       const ConvGeom& g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<A_MODE, BK, M256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv)");
    attr_set = true;
  }
  const int max_ctas = state().num_sms * g.ctas_per_sm;
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  fill_launch_config(cfg, attr, grid, A_MODE == A_TMA ? kThreadsTma : kThreadsGather, g.smem_bytes, st);
  return check_cuda(cudaLaunchKernelEx(&cfg, conv_umma_kernel<A_MODE, BK, M256>, ta, tb, to, tr, p), "conv launch");
}

int conv2d_fwd(const VcbConvDesc& d_in, const void* x, const void* w_packed, const float* bias_packed, const void* residual,
               void* y, cudaStream_t st, const int* stat_seg, double* stat_sums) {
  VcbConvDesc d = d_in;
  const int launch_flags = d.reserved[1] >> 8;      // bit 0: walk the tiles in reverse order
  d.reserved[1] &= 0xff;
  if (stat_sums != nullptr) {
    if (stat_seg == nullptr || d.act != VCB_ACT_NONE || d.res_mode != VCB_RES_NONE || d.out_dtype != VCB_F16 || d.reserved[0] != 0)
      return set_error(VCB_ERR_INVALID, "conv (BN statistics): needs act = none, no residual, fp16 output and a segment table");
    d.reserved[1] |= 0x200;                         // geometry: stay on the kernels that run the split epilogue
  }
  int rc = require_init();
  if (rc != VCB_OK) return rc;
  ConvGeom g;
  rc = conv_geometry(d, g);
  if (rc != VCB_OK) return rc;
  if (x == nullptr || w_packed == nullptr || bias_packed == nullptr || y == nullptr ||
      (d.res_mode != VCB_RES_NONE && residual == nullptr))
    return set_error(VCB_ERR_INVALID, "conv: null pointer");
  if (((uintptr_t)x & 15) || ((uintptr_t)w_packed & 15) || ((uintptr_t)y & 15) || ((uintptr_t)residual & 15) ||
      ((uintptr_t)bias_packed & 15))
    return set_error(VCB_ERR_INVALID, "conv: pointers must be 16-byte aligned");

  ConvParams p{};
  p.N = d.n; p.H = d.h; p.W = d.w; p.Cin = d.cin; p.cin_pitch = d.cin_pitch;
  p.P = g.P; p.Q = g.Q; p.PQ = g.P * g.Q; p.M = g.M;
  p.kh = d.kh; p.kw = d.kw; p.stride = d.stride; p.pad = d.pad;
  p.Cout = d.cout; p.cout_store = (d.cout + 7) / 8 * 8; p.out_pitch = d.cout_pitch;
  p.chunks_per_tap = g.chunks_per_tap; p.num_k_iters = g.num_k_iters;
  p.bk = g.bk; p.chunks_per_stage = g.chunks_per_stage; p.total_chunks = g.total_chunks;
  p.block_n = g.block_n; p.n_tiles = g.n_tiles;
  p.m_tiles = (g.M + g.tile_m - 1) / g.tile_m; p.num_tiles = p.m_tiles * g.n_tiles;
  p.num_pair_tiles = ((g.m_tiles + 1) / 2) * g.n_tiles;
  p.num_stages = g.stages; p.acc_stages = g.acc_stages; p.tmem_cols = g.tmem_cols;
  p.act = d.act; p.res_mode = d.res_mode; p.res_pitch = d.res_pitch; p.out_fp32 = d.out_dtype == VCB_F32 ? 1 : 0;
  p.x = reinterpret_cast<const __half*>(x);
  p.bias = bias_packed;
  p.residual = reinterpret_cast<const __half*>(residual);
  p.out = y;
  p.fault = state().fault_dev;
  p.prof = state().prof_on ? state().prof_dev : nullptr;
  p.epi_direct = d.reserved[0] == 1 ? 1 : 0;
  p.split_b = 0;   // (a second producer thread for the weight tiles measured no gain; code path kept for experiments only)
  p.b_resident = g.b_resident; p.b_res_bytes = g.b_res_bytes;
  p.dbg_skip_epilogue = (d.reserved[0] == 3 || d.reserved[1] == 6 || d.reserved[1] == 7) ? 1 : 0;
  p.dbg_swap = d.reserved[1] == 7 ? 1 : 0;
  if (p.dbg_swap) { p.tmem_cols = 512; p.acc_stages = 2; }
  p.cout_pad = g.cout_pad;
  p.epi_kind = g.epi_kind;
  p.tile_rev = (launch_flags & 1) ? 1 : 0;
  p.ksteps_lim = 4;
  {
    const int sub_cols = d.out_dtype == VCB_F32 ? 32 : 64;
    const int num_sub = (g.block_n + sub_cols - 1) / sub_cols;
    p.epi_split = (state().epi_split && g.epi_kind != 0 && !g.patch && !g.m256 && g.out_bufs == 2 && g.kchains == 1 &&
                   (num_sub >= 2 || g.acc_stages == 2)) ? 1 : 0;
    p.epi_empty_count = (p.epi_split && num_sub == 1 ? kNumEpilogueThreads / 2 : kNumEpilogueThreads) * (g.two_cta ? 2 : 1);
  }
  if (stat_sums != nullptr && (!p.epi_split || g.rowwin))
    return set_error(VCB_ERR_INVALID, "conv (BN statistics): this geometry does not run the split epilogue");
  p.stat_seg = stat_seg;
  p.stat_sums = stat_sums;
  p.a_policy = state().l2_hint ? kL2EvictFirst : kL2EvictNormal;
  p.b_policy = state().l2_hint ? kL2EvictLast : kL2EvictNormal;
  p.kchains = g.kchains;
  if (g.patch) {
    p.patch_R = g.patch_R; p.patch_Xs = g.patch_Xs; p.patch_Lp = g.patch_Lp; p.patch_xsegs = g.patch_xsegs; p.patch_ytiles = g.patch_ytiles;
    p.patch_box_bytes = (g.patch_R + 2) * g.patch_Lp * 128;
    p.patch_stage_bytes = g.patch_stage_bytes;
    p.patch_out_bytes = g.patch_R * g.patch_Xs * 128;
    p.a_stages = g.a_stages; p.b_stages = g.b_stages;
    p.m_tiles = d.n * g.patch_ytiles * g.patch_xsegs;
    p.num_tiles = p.m_tiles * g.n_tiles;
  }
  if (d.reserved[0] == 4) { p.acc_stages = 1; }
  p.out_stage_bufs = g.out_bufs;
  p.out_stage_bytes = g.out_bufs * (g.tile_m * 128);
  p.c4_wide = (g.a_mode == A_C4 && d.kw % 2 == 0 && d.stride % 2 == 0 && d.pad % 2 == 0 && d.w % 2 == 0 && d.reserved[1] != 1) ? 1 : 0;

  const CUtensorMapSwizzle swz = g.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (g.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  alignas(64) CUtensorMap ta, tb, to, tr;
  memset(&ta, 0, sizeof(ta));
  memset(&to, 0, sizeof(to));
  memset(&tr, 0, sizeof(tr));
  if (g.patch) {
    // 4-D views (C, W, H, N): input patch box {64, Xs+2, R+2, 1}; output / residual box {64 | 32, Xs, R, 1}; weights as usual
    const cuuint32_t estr4[4] = {1, 1, 1, 1};
    {
      const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
      const cuuint64_t strides[3] = {(cuuint64_t)d.cin_pitch * 2, (cuuint64_t)d.w * d.cin_pitch * 2, (cuuint64_t)d.h * d.w * d.cin_pitch * 2};
      const cuuint32_t box[4] = {64u, (cuuint32_t)g.patch_Lp, (cuuint32_t)(g.patch_R + 2), 1u};
      const CUresult r = state().encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr4,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(patch input) failed: %d", (int)r);
    }
    {
      const bool f32 = d.out_dtype == VCB_F32;
      const cuuint64_t es = f32 ? 4 : 2;
      const cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
      const cuuint64_t strides[3] = {(cuuint64_t)d.cout_pitch * es, (cuuint64_t)d.w * d.cout_pitch * es, (cuuint64_t)d.h * d.w * d.cout_pitch * es};
      const cuuint32_t box[4] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)g.patch_Xs, (cuuint32_t)g.patch_R, 1u};
      const CUresult r = state().encode_tiled(&to, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, y, dims, strides, box,
                                              estr4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(patch output) failed: %d", (int)r);
    }
    if (d.res_mode != VCB_RES_NONE) {
      const cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
      const cuuint64_t strides[3] = {(cuuint64_t)d.res_pitch * 2, (cuuint64_t)d.w * d.res_pitch * 2, (cuuint64_t)d.h * d.w * d.res_pitch * 2};
      const cuuint32_t box[4] = {64u, (cuuint32_t)g.patch_Xs, (cuuint32_t)g.patch_R, 1u};
      const CUresult r = state().encode_tiled(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(residual), dims, strides, box, estr4,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(patch residual) failed: %d", (int)r);
    }
    {
      const cuuint64_t dims[2] = {(cuuint64_t)g.k_pad, (cuuint64_t)g.cout_pad};
      const cuuint64_t strides[1] = {(cuuint64_t)g.k_pad * 2};
      const cuuint32_t box[2] = {64u, (cuuint32_t)g.block_n};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = state().encode_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    }
    static bool attr_set = false;
    if (!attr_set) {
      const cudaError_t e = cudaFuncSetAttribute(conv_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv patch)");
      attr_set = true;
    }
    const int grid = p.num_tiles < state().num_sms ? p.num_tiles : state().num_sms;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    fill_launch_config(cfg, attr, grid, kThreadsPatch, g.smem_bytes, st);
    return check_cuda(cudaLaunchKernelEx(&cfg, conv_patch_kernel, ta, tb, to, tr, p), "conv (patch) launch");
  }
  if (g.rowwin) {
    // Row-window mode.  Input: W-padded NHWC16, pixel (y, x) at column x + 1.  Tensor-map view (k, x, y, n): element k of window x
    // of row y at byte 2k + 32x + 32(W+2)y -- the pixel stride (32 B) is SMALLER than the 128-byte inner extent, i.e. neighbouring
    // windows overlap; the window of output column x holds input columns x-1 .. x+2 (the last one against zero weights).  Rows
    // y = -1 and y = H are out of range and zero-filled by the TMA unit, the left / right borders are the buffer's zero pad columns.
    p.a_rowwin = 1;
    p.ksteps_lim = 3;
    p.patch_R = g.patch_R; p.patch_Xs = g.patch_Xs; p.patch_Lp = g.patch_Lp; p.patch_xsegs = g.patch_xsegs; p.patch_ytiles = g.patch_ytiles;
    p.patch_out_bytes = kStageOutBytes;
    p.m_tiles = g.m_tiles;
    p.num_tiles = p.m_tiles * g.n_tiles;
    const cuuint32_t estr4[4] = {1, 1, 1, 1};
    {
      const cuuint64_t wp = (cuuint64_t)d.w + 2;
      const cuuint64_t dims[4] = {64u, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
      const cuuint64_t strides[3] = {32u, wp * 32u, (cuuint64_t)d.h * wp * 32u};
      const cuuint32_t box[4] = {64u, (cuuint32_t)g.patch_Xs, (cuuint32_t)g.patch_R, 1u};
      const CUresult r = state().encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr4,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(row windows) failed: %d", (int)r);
    }
    {
      const cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
      const cuuint64_t strides[3] = {(cuuint64_t)d.cout_pitch * 2, (cuuint64_t)d.w * d.cout_pitch * 2, (cuuint64_t)d.h * d.w * d.cout_pitch * 2};
      const cuuint32_t box[4] = {64u, (cuuint32_t)g.patch_Xs, (cuuint32_t)g.patch_R, 1u};
      const CUresult r = state().encode_tiled(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, y, dims, strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(row-window output) failed: %d", (int)r);
    }
    {
      const cuuint64_t dims[2] = {(cuuint64_t)g.k_pad, (cuuint64_t)g.cout_pad};
      const cuuint64_t strides[1] = {(cuuint64_t)g.k_pad * 2};
      const cuuint32_t box[2] = {64u, (cuuint32_t)g.block_n};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = state().encode_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    }
    return launch_conv<A_TMA, 64>(ta, tb, to, tr, p, g, st);
  }
  if (p.epi_kind != 0 && d.res_mode != VCB_RES_NONE) {   // residual: [M][cout] fp16 view with row pitch res_pitch, same boxes as the output
    const cuuint64_t dims[2] = {(cuuint64_t)d.cout, (cuuint64_t)g.M};
    const cuuint64_t strides[1] = {(cuuint64_t)d.res_pitch * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)g.tile_m};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(residual), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(residual) failed: %d", (int)r);
  }
  if (!p.epi_direct) {   // output: [M][cout] slice of an NHWC buffer with row pitch cout_pitch; box = 128 bytes x 128 rows
    const bool f32 = d.out_dtype == VCB_F32;
    const cuuint64_t dims[2] = {(cuuint64_t)d.cout, (cuuint64_t)g.M};
    const cuuint64_t strides[1] = {(cuuint64_t)d.cout_pitch * (f32 ? 4 : 2)};
    const cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)g.tile_m};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&to, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y, dims,
                                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed: %d", (int)r);
  }
  {   // B: [cout_pad][k_pad] fp16, box = 64 (K) x block_n rows, 128-byte swizzle
    const cuuint64_t dims[2] = {(cuuint64_t)g.k_pad, (cuuint64_t)g.cout_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)g.k_pad * 2};
    const cuuint32_t box[2] = {(cuuint32_t)g.bk, (cuuint32_t)(g.two_cta ? g.block_n / 2 : g.block_n)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_packed), dims, strides,
                                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  }
  p.a_tiled = (g.a_mode == A_TMA && d.kh == 1 && d.kw == 1 && d.stride == 1 && d.pad == 0 && d.reserved[1] != 2) ? 1 : 0;
  if (p.a_tiled) {
    // 1x1/s1/p0: A is the [M][cin] matrix itself (row pitch cin_pitch): plain tiled TMA, box = bk channels x 128 rows
    const cuuint64_t dims[2] = {(cuuint64_t)d.cin, (cuuint64_t)g.M};
    const cuuint64_t strides[1] = {(cuuint64_t)d.cin_pitch * 2};
    const cuuint32_t box[2] = {(cuuint32_t)g.bk, (cuuint32_t)g.tile_m};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = state().encode_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
  } else if (g.a_mode == A_TMA) {
    // A: NHWC activations as a (C, W, H, N) tensor in im2col mode.  Bounding box of base pixels:
    // lower corner = -pad, upper corner = pad - (k - 1); traversal stride = conv stride; one load =
    // 128 consecutive output pixels x 64 channels at filter offset (s, r).
    const cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
    const cuuint64_t strides[3] = {(cuuint64_t)d.cin_pitch * 2, (cuuint64_t)d.w * d.cin_pitch * 2,
                                   (cuuint64_t)d.h * d.w * d.cin_pitch * 2};
    const int lower[2] = {-d.pad, -d.pad};
    const int upper[2] = {d.pad - (d.kw - 1), d.pad - (d.kh - 1)};
    const cuuint32_t estr[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
    const CUresult r = state().encode_im2col(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, lower,
                                             upper, (cuuint32_t)g.bk, (cuuint32_t)g.tile_m, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VCB_ERR_CUDA, "cuTensorMapEncodeIm2col failed: %d", (int)r);
    // Driver quirk for small tensors in im2col mode (same adjustment CUTLASS applies, see
    // cute/atom/copy_traits_sm90_im2col.hpp make_im2col_tma_copy_desc): clear bit 21 of the second
    // descriptor word when the tensor spans < 128 KiB on drivers <= 13.1.
    if (state().driver_version <= 13010) {
      const unsigned long long span = (unsigned long long)d.n * d.h * d.w * d.cin_pitch * 2ull;
      if (span < 131072ull) reinterpret_cast<uint64_t*>(&ta)[1] &= ~(1ull << 21);
    }
  }
  switch (g.a_mode) {
    case A_TMA:
      if (g.two_cta) {
        if (g.bk == 64) return launch_conv_2cta<64>(ta, tb, to, tr, p, g, st);
        if (g.bk == 32) return launch_conv_2cta<32>(ta, tb, to, tr, p, g, st);
        return launch_conv_2cta<16>(ta, tb, to, tr, p, g, st);
      }
      if (g.m256) {
        if (g.bk == 64) return launch_conv<A_TMA, 64, true>(ta, tb, to, tr, p, g, st);
        if (g.bk == 32) return launch_conv<A_TMA, 32, true>(ta, tb, to, tr, p, g, st);
        return launch_conv<A_TMA, 16, true>(ta, tb, to, tr, p, g, st);
      }
      if (g.bk == 64) return launch_conv<A_TMA, 64>(ta, tb, to, tr, p, g, st);
      if (g.bk == 32) return launch_conv<A_TMA, 32>(ta, tb, to, tr, p, g, st);
      return launch_conv<A_TMA, 16>(ta, tb, to, tr, p, g, st);
    case A_GATHER: return launch_conv<A_GATHER, 64>(ta, tb, to, tr, p, g, st);
    default: return launch_conv<A_C4, 64>(ta, tb, to, tr, p, g, st);
  }
}

}  // namespace vcb
