// Micro-benchmark (debug aid): per-SMSP throughput of the instructions the softmax loop is made of.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.01f;
  float c = 0.999f, d = -0.001f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);                                   // MUFU only (dependent per element, 16 independent chains)
      if (MODE == 1) a[i] = fmaf(a[i], c, d);                            // FFMA only
      if (MODE == 2) a[i] = ex2(fmaf(a[i], c, d));                       // FFMA + MUFU
      if (MODE == 3) { float p = ex2(fmaf(a[i], c, d)); a[i] = p + a[i] * 0.5f; }   // FFMA + MUFU + FFMA
      if (MODE == 4) { float p = ex2(fmaf(a[i], c, d)); __nv_bfloat162 h = __floats2bfloat162_rn(p, a[i]); a[i] = p + __uint_as_float(*reinterpret_cast<unsigned*>(&h) & 0x3f800000u); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double per = (double)h[0] / (iters * 16.0);
  printf("%-28s warps/SM %2d (%d per SMSP): %.2f cycles per warp-level element-step; per SMSP: %.2f cycles per warp-instr group\n", name, warps, warps / 4,
         per, per / (warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("MUFU.EX2", 4); run<1>("FFMA", 4); run<2>("FFMA+MUFU", 4); run<3>("FFMA+MUFU+FFMA", 4); run<4>("FFMA+MUFU+F2FP+..", 4); }
    if (w == 8) { run<0>("MUFU.EX2", 8); run<1>("FFMA", 8); run<2>("FFMA+MUFU", 8); run<3>("FFMA+MUFU+FFMA", 8); run<4>("FFMA+MUFU+F2FP+..", 8); }
    if (w == 16) { run<0>("MUFU.EX2", 16); run<1>("FFMA", 16); run<2>("FFMA+MUFU", 16); run<3>("FFMA+MUFU+FFMA", 16); run<4>("FFMA+MUFU+F2FP+..", 16); }
  }
  return 0;
}
