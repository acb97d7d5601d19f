// Library-private declarations shared by the .cu translation units of libvcb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/vcb200.h"
#include "vcb_ptx.cuh"

namespace vcb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct State {
  bool initialised = false;
  int device = -1;
  int num_sms = 0;
  int driver_version = 0;
  int pdl = 1;                         // conv launches carry the programmatic-dependent-launch attribute ($VCB_PDL=0 / vcb_set_option turn it
                                       // off): +6-7 % at B = 1 ... 8 (grids smaller than the GPU: the next kernel's prologue overlaps), neutral at B = 64
  EncodeTiledFn encode_tiled = nullptr;
  EncodeIm2colFn encode_im2col = nullptr;
  KernelFault* fault_host = nullptr;   // pinned + mapped: still readable after a trapped kernel
  KernelFault* fault_dev = nullptr;
  int l2_hint = 0;                     // activation TMA loads evict-first, weight loads evict-last ($VCB_L2_HINT / vcb_set_option)
  int epi_split = 1;                   // conv epilogue as two independent four-warp groups ($VCB_EPI_SPLIT / vcb_set_option)
  int prof_on = 0;                     // conv role timers (development aid)
  unsigned long long* prof_dev = nullptr;   // 16 counters in device memory
};
State& state();

int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int require_init();

// conv_umma.cu
int conv_packed_sizes(const VcbConvDesc& d, int64_t* weight_halfs, int64_t* bias_floats);
int conv_out_hw(const VcbConvDesc& d, int32_t* ho, int32_t* wo);
int conv_pack_weights(const VcbConvDesc& d, const float* w, const float* bias, void* wp, float* bp, cudaStream_t st);
int conv2d_fwd(const VcbConvDesc& d, const void* x, const void* w_packed, const float* bias_packed, const void* residual,
               void* y, cudaStream_t st, const int* stat_seg = nullptr, double* stat_sums = nullptr);

}  // namespace vcb
