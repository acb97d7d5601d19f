// Micro-benchmark: per-SM throughput of tiled-mode TMA loads (L2-resident source) as a function of box rows,
// row bytes and the number of loads kept in flight.  Used to size the conv kernel's pipeline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/tma_probe.cu -lcuda && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// one thread per CTA issues `iters` loads of one box each, keeping `depth` in flight; rows of consecutive loads advance
__global__ void probe(const __grid_constant__ CUtensorMap tm, int box_bytes, int box_rows, int depth, int iters, int total_rows,
                      long long* cycles_out, int hot) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // hot = 1: every CTA walks the same rows (like the weight tiles of a conv layer); 0: disjoint rows per CTA
    const int row0 = hot ? 0 : (int)((long long)blockIdx.x * 9973 % (total_rows - box_rows * (iters + 1)));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      if (i >= depth) mbar_wait(smem_u32(&bars[s]), ((i / depth) - 1) & 1);
      mbar_expect(smem_u32(&bars[s]), box_bytes);
      tma_2d(&tm, smem_u32(&bars[s]), base + s * box_bytes, 0, row0 + (hot ? (i % 27) : i) * box_rows);
    }
    for (int i = iters; i < iters + depth; ++i) {
      const int s = i % depth;
      mbar_wait(smem_u32(&bars[s]), ((i / depth) - 1) & 1);
    }
    cycles_out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  Enc enc = (Enc)fn;
  const int total_rows = 1 << 17;              // x 256 B pitch = 32 MiB: L2 resident after the first pass
  const int pitch = 256;
  void* buf;
  CK(cudaMalloc(&buf, (size_t)total_rows * pitch));
  CK(cudaMemset(buf, 1, (size_t)total_rows * pitch));
  long long* d_cycles;
  CK(cudaMalloc(&d_cycles, sizeof(long long) * 1024));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 400;
  printf("hot row_bytes box_rows depth grid | cycles/op  B/clk/SM  (clock %d kHz)\n", prop.clockRate);
  for (int hot : {0, 1})
  for (int row_bytes : {64, 128}) {
    for (int box_rows : {48, 96, 128, 192, 256}) {
      for (int depth : {2, 4}) {
        const int box_bytes = row_bytes * box_rows;
        if ((size_t)box_bytes * depth > 190 * 1024) continue;
        alignas(64) CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 2), (cuuint64_t)total_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)pitch};
        const cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)box_rows};
        const cuuint32_t es[2] = {1, 1};
        const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        for (int grid : {prop.multiProcessorCount}) {
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, 32, (size_t)box_bytes * depth + 1024>>>(tm, box_bytes, box_rows, depth, iters, total_rows, d_cycles, hot);
            CK(cudaDeviceSynchronize());
          }
          std::vector<long long> h(grid);
          CK(cudaMemcpy(h.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
          double avg = 0;
          for (long long v : h) avg += (double)v;
          avg /= grid;
          printf("%3d %9d %8d %5d %4d | %9.1f %9.2f\n", hot, row_bytes, box_rows, depth, grid, avg / iters, (double)box_bytes * iters / avg);
        }
      }
    }
  }
  return 0;
}
