// MUFU.EX2 issue rate per scheduler: W warps per SMSP, each running a straight-line block of independent ex2.approx
// (optionally with the softmax companions: one FFMA before, one FADD after, one bf16 pack per pair).
// Prints cycles per warp-level MUFU instruction per scheduler.  Result on B200: 8.1 cycles with FFMA + FADD (+FMUL) around
// it — the MUFU peak of 16 lanes per clock and SM — with >= 2 warps per scheduler, 9.4 with a single warp; 10.0 once a
// bf16 pack per pair feeds a serial XOR chain, with F2FP or with integer rounding alike (so the pack is not an XU op).  nvcc -arch=sm_100a -O3 mufu_probe.cu -o mufu_probe.bin
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, float c, float m, int iters) {
  float x[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  float s0 = 0.f, s1 = 0.f;
  unsigned pk = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float a = x[i], b = x[i + 1];
      if (MODE >= 1) { a = fmaf(a, c, m); b = fmaf(b, c, m); }
      float e0, e1;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(b));
      if (MODE >= 2) { s0 += e0; s1 += e1; }
      if (MODE == 3) { __nv_bfloat162 v = __floats2bfloat162_rn(e0, e1); pk ^= *reinterpret_cast<unsigned*>(&v); }
      if (MODE == 4) pk ^= __byte_perm(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632);   // integer rounding
      x[i] = e0 * 0.5f; x[i + 1] = e1 * 0.5f;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + x[3] + __uint_as_float(pk);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(int warps_per_smsp) {
  const int threads = 128 * warps_per_smsp, iters = 200;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, threads>>>(out, cyc, 1.01f, -0.3f, iters);
  k<MODE><<<148, threads>>>(out, cyc, 1.01f, -0.3f, iters);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("mode %d (2 ffma+ex2+fadd, 3 +F2FP pack, 4 +integer-rounded pack)  warps/SMSP %d: %.2f cycles per MUFU warp-instruction per scheduler\n",
         MODE, warps_per_smsp, avg / (double)(iters * 64 * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w = 1; w <= 4; w *= 2) { run<2>(w); run<3>(w); run<4>(w); }
  return 0;
}
