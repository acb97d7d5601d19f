// Probe: may the A operand of tcgen05.mma (K-major, 128-byte swizzle) start at ANY 128-byte row of a TMA-written tile,
// i.e. not on a 1024-byte swizzle-atom boundary?  A 3x3 convolution could then keep one input patch in shared memory and
// point the nine taps' descriptors at shifted rows of it (9x less A traffic than one im2col box per tap).  Variants of the
// shared-memory descriptor: base_offset field (bits 49-51) = 0, or = (start_address >> 7) & 7 as the PTX ISA describes for
// start addresses inside a swizzle atom.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vehicle_counting_b200/csrc -o gpurun_out/umma_shift_probe tools/umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vcb_ptx.cuh"

using namespace vcb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kRowsA = 160, kK = 64, kN = 64;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* out,
                                            int shift, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + kRowsA * 128, bars = b_smem + kN * 128;
  const uint32_t full = bars, done = bars + 8, slot = bars + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 64u); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(full, (uint32_t)(kRowsA * 128 + kN * 128));
    tma_load_2d(&tm_a, full, a_smem, 0, 0);
    tma_load_2d(&tm_b, full, b_smem, 0, 0);
    mbar_wait_tight(full, 0u, nullptr, 0, 0);
    tcgen05_fence_after();
    const uint32_t a_start = a_smem + (uint32_t)shift * 128u;
    uint64_t a_desc = umma_desc_kmajor(a_start, 1024u, 2u);
    if (use_base_offset) a_desc |= (uint64_t)((a_start >> 7) & 7u) << 49;
    const uint64_t b_desc = umma_desc_kmajor(b_smem, 1024u, 2u);
    const uint32_t idesc = umma_idesc_f16((uint32_t)kN);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait_tight(done, 0u, nullptr, 0, 0);
  tcgen05_fence_after();
  uint32_t v[16];
  for (int c = 0; c < kN; c += 16) {
    tmem_ld_x16(tmem + (uint32_t)c + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(size_t)threadIdx.x * kN + c + i] = __uint_as_float(v[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64u);
}

int main() {
  CK(cudaSetDevice(0));
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  Enc enc = (Enc)fn;
  std::vector<__half> ha(kRowsA * kK), hb(kN * kK);
  std::vector<float> fa(kRowsA * kK), fb(kN * kK);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2half((rand() % 17 - 8) / 8.0f); fa[i] = __half2float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2half((rand() % 13 - 6) / 4.0f); fb[i] = __half2float(hb[i]); }
  __half *da, *db;
  float* dout;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 128 * kN * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  alignas(64) CUtensorMap ta, tb;
  const cuuint32_t es[2] = {1, 1};
  {
    const cuuint64_t dims[2] = {kK, kRowsA}; const cuuint64_t st[1] = {kK * 2}; const cuuint32_t box[2] = {kK, kRowsA};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
  }
  {
    const cuuint64_t dims[2] = {kK, kN}; const cuuint64_t st[1] = {kK * 2}; const cuuint32_t box[2] = {kK, kN};
    if (enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
  }
  const size_t smem = 1024 + kRowsA * 128 + kN * 128 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> hout(128 * kN);
  printf("shift base_offset | max |got - ref|   (ref = A[shift:shift+128] x B^T, exact in fp32)\n");
  for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 27, 31}) {
    for (int bo : {0, 1}) {
      CK(cudaMemset(dout, 0xff, 128 * kN * 4));
      probe<<<1, 128, smem>>>(ta, tb, dout, shift, bo);
      const cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%5d %11d | kernel error: %s\n", shift, bo, cudaGetErrorString(e)); return 0; }
      CK(cudaMemcpy(hout.data(), dout, 128 * kN * 4, cudaMemcpyDeviceToHost));
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += (double)fa[(size_t)(m + shift) * kK + k] * fb[(size_t)n * kK + k];
          const double d = fabs(ref - hout[(size_t)m * kN + n]);
          if (!(d <= worst)) worst = d;
        }
      printf("%5d %11d | %g\n", shift, bo, worst);
    }
  }
  return 0;
}
