// Probe: sustained cost of one tcgen05.mma (kind::f16, M = 128, K = 16, both operands in shared memory, 128-byte swizzle)
// as a function of N, of the number of independent accumulators the instructions rotate over, and of the number of
// co-resident CTAs per SM.  No TMA, no epilogue: the operands sit in shared memory for the whole kernel, one thread
// issues `iters` groups of four MMAs (one 64-wide K chunk) and the kernel reports cycles per MMA from the first issue to the
// arrival of the final tcgen05.commit.  Explains why the narrow-N conv layers sit far below the tensor peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vehicle_counting_b200/csrc -o build/umma_rate_probe tools/umma_rate_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vcb_ptx.cuh"

using namespace vcb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

// smem: A tile 128 rows x 128 B (16 KiB), B tile n rows x 128 B; contents are whatever (zeros): timing only
__global__ void __launch_bounds__(128) probe(int n, int chains, int iters, int tmem_cols, int issuers, int shift_rows, int rotate, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 8u * 16384u + 4096u, bars = b_smem + 2u * 256u * 128u;      // 8 A tiles (+ slack for shifted starts), 2 B tiles
  const uint32_t done = bars, slot = bars + 32;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  uint4* z = reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (8 * 16384 + 4096 + 2 * 256 * 128) / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(done, 1); mbar_init(done + 8, 1); mbar_init(done + 16, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(slot, (uint32_t)tmem_cols); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot_ptr;
  // issuer w (w < issuers) is lane 0 of warp 1 + w; each has its own accumulators and its own completion barrier
  if ((threadIdx.x & 31) == 0 && warp >= 1 && warp <= issuers) {
    const uint32_t me = (uint32_t)(warp - 1);
    const uint32_t my_done = done + 8u * me;
    const uint32_t my_tmem = tmem + me * (uint32_t)(n * chains);
    const uint64_t a_desc = umma_desc_kmajor(a_smem + (uint32_t)shift_rows * 128u, 1024u, 2u);   // shift_rows != multiple of 8: start inside a swizzle atom
    const uint64_t b_desc = umma_desc_kmajor(b_smem, 1024u, 2u);
    const uint32_t idesc = umma_idesc_f16((uint32_t)n);
    const long long t0 = clock64();
    uint32_t ks = 0;
    for (int it = 0; it < iters; ++it) {
      // rotate > 1: every group of four MMAs reads a different A tile (16 KiB apart) and B tile, as a real main loop does
      const uint64_t ao = (uint64_t)(((uint32_t)(it + (int)me) % (uint32_t)rotate) * (16384u >> 4));
      const uint64_t bo = (uint64_t)((((uint32_t)it / (uint32_t)rotate) & 1u) * ((256u * 128u) >> 4)) * (rotate > 1 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k, ++ks)
        umma_f16(my_tmem + (ks % (uint32_t)chains) * (uint32_t)n, a_desc + ao + (uint64_t)(2 * k), b_desc + bo + (uint64_t)(2 * k), idesc, ks >= (uint32_t)chains ? 1u : 0u);
    }
    umma_commit(my_done);
    mbar_wait_tight(my_done, 0u, nullptr, 0, 0);
    if (me == 0) out[blockIdx.x] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 4096));
  const size_t smem = 1024 + 8 * 16384 + 4096 + 2 * 256 * 128 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  printf("N chains ctas/SM issuers/CTA | cycles per tcgen05.mma (per issuer) | per SM | ideal N/2 | tensor-pipe utilisation\n");
  for (int n : {32, 48, 64, 96, 128, 192, 256}) {
    for (int ctas : {1, 2, 3, 4}) {
     for (int issuers : {1, 2, 3}) {
      for (int chains : {1, 2}) {
        if (chains == 2 && (issuers > 1 || ctas > 2)) continue;
        int cols = 32;
        while (cols < n * chains * issuers) cols <<= 1;
        if (cols * ctas > 512) continue;
        std::vector<long long> h(sms * ctas);
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<sms * ctas, 128, smem>>>(n, chains, iters, cols, issuers, 0, 1, d_out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * sms * ctas, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (long long v : h) avg += (double)v;
        avg /= h.size();
        const double per = avg / (4.0 * iters), per_sm = per / (ctas * issuers);
        printf("%3d %6d %7d %7d | %8.1f | %8.1f | %5.1f | %5.1f %%\n", n, chains, ctas, issuers, per, per_sm, n / 2.0, 100.0 * (n / 2.0) / per_sm);
      }
     }
    }
  }
  printf("\nA operand start shifted by whole 128-byte rows (patch mode): N issuers shift | cycles per MMA per SM\n");
  for (int n : {64, 192}) {
    for (int issuers : {1, 2, 3}) {
      for (int shift : {0, 8, 1, 3, 29}) {
        int cols = 32;
        while (cols < n * issuers) cols <<= 1;
        if (cols > 512) continue;
        std::vector<long long> h(sms);
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<sms, 128, smem>>>(n, 1, iters, cols, issuers, shift, 1, d_out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (long long v : h) avg += (double)v;
        avg /= h.size();
        printf("%3d %7d %5d | %8.1f\n", n, issuers, shift, avg / (4.0 * iters) / issuers);
      }
    }
  }
  printf("\nfresh operands: every K chunk reads another A tile (8 tiles in rotation) and alternates B tiles: N issuers rotate | cycles per MMA per SM | tensor util\n");
  for (int n : {64, 128, 192, 256}) {
    for (int issuers : {1, 2, 3}) {
      for (int rot : {1, 8}) {
        int cols = 32;
        while (cols < n * issuers) cols <<= 1;
        if (cols > 512) continue;
        std::vector<long long> h(sms);
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<sms, 128, smem>>>(n, 1, iters, cols, issuers, 0, rot, d_out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (long long v : h) avg += (double)v;
        avg /= h.size();
        const double per_sm = avg / (4.0 * iters) / issuers;
        printf("%3d %7d %6d | %8.1f | %5.1f %%\n", n, issuers, rot, per_sm, 100.0 * (n / 2.0) / per_sm);
      }
    }
  }
  return 0;
}