// Issue-rate probe for tcgen05.mma kind::f16 shapes on one SM: how many cycles does one MMA (K=16) take as a function of
// M, N, of where A lives (shared memory vs tensor memory), of accumulator alternation and of the number of issuing
// warps?  Operands are whatever is in shared memory (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -I tair_b200/csrc -o tools/probes/mma_shape_probe.bin tools/probes/mma_shape_probe.cu
#include <cstdio>
#include <cuda.h>
#include "../../include/tair_b200.h"
#include "common.cuh"
using namespace tair;

// mode bit0: A from TMEM; bit1: alternate two accumulators; nissue: 1 or 2 issuing warps (own accumulators)
__global__ void __launch_bounds__(128) probe(int M, int N, int mode, int nissue, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar[2];
  const uint32_t sbase = (smem_u32(smem) + 1023) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (sbase - smem_u32(smem)))[i] = 0;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const int w = threadIdx.x >> 5;
  if ((mode & 4) && w < nissue) {
    // warp-uniform issue path: all 32 lanes run the loop, one elected lane issues (operands stay in uniform registers)
    const int ts = mode & 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_lo = umma_desc_lo(sbase + w * 49152, 16), b_lo = umma_desc_lo(sbase + w * 49152 + 16384, 16);
    const uint32_t d0 = tm + (nissue == 2 ? w * 256 : 0);
    const uint32_t a_t = (nissue == 2) ? d0 + 224 : tm + 256;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ts) umma_ts_lohi(d0, a_t + k * 8, b_lo + 2 * k, hi, idesc, 1);
          else umma_ss_lohi(d0, a_lo + 2 * k, b_lo + 2 * k, hi, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar[w]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[w]), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 2 + w] = t1 - t0;
  } else if ((threadIdx.x & 31) == 0 && w < nissue) {
    const int ts = mode & 1, alt = (mode >> 1) & 1;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_lo = umma_desc_lo(sbase + w * 49152, 16), b_lo = umma_desc_lo(sbase + w * 49152 + 16384, 16);
    // accumulators: nissue==2 -> warp w owns columns [w*256, w*256+N) (N <= 224 when A is in TMEM at +224..)
    const uint32_t d0 = tm + (nissue == 2 ? w * 256 : 0);
    const uint32_t d1 = alt ? tm + 256 : d0;
    const uint32_t a_t = (nissue == 2 || alt) ? d0 + 224 : tm + 256;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = (k & 1) ? d1 : d0;
        if (ts) umma_ts_lohi(d, a_t + k * 8, b_lo + 2 * k, hi, idesc, 1);
        else umma_ss_lohi(d, a_lo + 2 * k, b_lo + 2 * k, hi, idesc, 1);
      }
    }
    umma_commit(smem_u32(&bar[w]));
    mbar_wait(smem_u32(&bar[w]), 0);
    long long t1 = clock64();
    out[blockIdx.x * 2 + w] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 2 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2048;
  for (int nissue = 1; nissue <= 2; ++nissue)
    for (int mode : {0, 1, 4, 5})
      for (int M : {64, 128})
        for (int N : {16, 32, 64, 96, 128, 160, 192, 256}) {
          if ((mode & 1) && (nissue == 2 || (mode & 2)) && N > 192) continue;
          if (nissue == 2 && (mode & 2)) continue;
          cudaMemset(d, 0, 16);
          probe<<<1, 128, 100 * 1024>>>(M, N, mode, nissue, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2];
          cudaMemcpy(h, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = h[0] > h[1] ? h[0] : h[1];
          printf("issuers %d %s %s M=%3d N=%3d: %.1f cycles per MMA per issuer (%s)\n", nissue, (mode & 1) ? "TS" : "SS",
                 (mode & 4) ? "elect " : "lane0 ", M, N, (double)mx / (iters * 4.0), cudaGetErrorString(e));
        }
  return 0;
}
