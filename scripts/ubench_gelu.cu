// Micro-benchmark (debug aid): issue cost of GELU epilogue formulations, 8 warps per SM (2 per SMSP) like the GEMM epilogue.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float gelu17(float x) {
  const float xc = fminf(fmaxf(x, -4.0f), 4.0f);
  const float x2 = xc * xc;
  float p = 8.062929977e-11f;
  p = fmaf(p, x2, -7.003156417e-09f);
  p = fmaf(p, x2, 2.716075885e-07f);
  p = fmaf(p, x2, -6.294891059e-06f);
  p = fmaf(p, x2, 9.890726931e-05f);
  p = fmaf(p, x2, -1.133918807e-03f);
  p = fmaf(p, x2, 9.877469438e-03f);
  p = fmaf(p, x2, -6.641058494e-02f);
  p = fmaf(p, x2, 3.989227133e-01f);
  return x * fmaf(p, xc, 0.5f);
}
// scalar, deg 15, clamp via x2 min + saturating fma
__device__ __forceinline__ float gelu15_sat(float x) {
  const float x2 = fminf(x * x, 16.0f);
  float p = -1.301277620e-09f;
  p = fmaf(p, x2, 1.041950039e-07f);
  p = fmaf(p, x2, -3.657106863e-06f);
  p = fmaf(p, x2, 7.485470993e-05f);
  p = fmaf(p, x2, -1.006488016e-03f);
  p = fmaf(p, x2, 9.505389249e-03f);
  p = fmaf(p, x2, -6.588782661e-02f);
  p = fmaf(p, x2, 3.986733839e-01f);
  return x * __saturatef(fmaf(p, x, 0.5f));
}
__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// packed pair, deg 15
__device__ __forceinline__ void gelu15_x2(float& x0, float& x1) {
  const unsigned long long X = pk(x0, x1);
  float s0, s1; upk(mul2(X, X), s0, s1);
  const unsigned long long X2 = pk(fminf(s0, 16.0f), fminf(s1, 16.0f));
  unsigned long long p = pk(-1.301277620e-09f, -1.301277620e-09f);
  p = fma2(p, X2, pk(1.041950039e-07f, 1.041950039e-07f));
  p = fma2(p, X2, pk(-3.657106863e-06f, -3.657106863e-06f));
  p = fma2(p, X2, pk(7.485470993e-05f, 7.485470993e-05f));
  p = fma2(p, X2, pk(-1.006488016e-03f, -1.006488016e-03f));
  p = fma2(p, X2, pk(9.505389249e-03f, 9.505389249e-03f));
  p = fma2(p, X2, pk(-6.588782661e-02f, -6.588782661e-02f));
  p = fma2(p, X2, pk(3.986733839e-01f, 3.986733839e-01f));
  float p0, p1; upk(p, p0, p1);
  x0 = x0 * __saturatef(fmaf(p0, x0, 0.5f));
  x1 = x1 * __saturatef(fmaf(p1, x1, 0.5f));
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float scale) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = (threadIdx.x * 0.01f + i * 0.1f - 2.0f) * scale;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = gelu17(a[i]) + 0.25f;
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = gelu15_sat(a[i]) + 0.25f;
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) { gelu15_x2(a[i], a[i + 1]); a[i] += 0.25f; a[i + 1] += 0.25f; }
    } else if (MODE == 3) {  // pure FFMA2 chain: 8 per pair
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        unsigned long long p = pk(a[i], a[i + 1]); const unsigned long long c = pk(0.999f, 0.998f), d = pk(0.001f, 0.002f);
#pragma unroll
        for (int j = 0; j < 8; ++j) p = fma2(p, c, d);
        upk(p, a[i], a[i + 1]);
      }
    } else if (MODE == 4) {  // pure FFMA (imm) chain: 8 per element
#pragma unroll
      for (int i = 0; i < 32; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[i] = fmaf(a[i], 0.999f, 0.001f);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 500;
  for (int warps : {4, 8}) {
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f); cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    printf("%-22s %d warps/SMSP: %.2f SMSP-cycles per warp-level element\n", name, warps / 4, (double)h[0] / (iters * 32.0) / (warps / 4.0));
  }
  cudaFree(out); cudaFree(cyc);
}
__global__ void acc(float* out) {
  float x = -6.0f + 12.0f * (blockIdx.x * blockDim.x + threadIdx.x) / (float)(gridDim.x * blockDim.x);
  float a = x, b = x + 1e-3f; gelu15_x2(a, b);
  out[3 * (blockIdx.x * blockDim.x + threadIdx.x)] = gelu17(x);
  out[3 * (blockIdx.x * blockDim.x + threadIdx.x) + 1] = gelu15_sat(x);
  out[3 * (blockIdx.x * blockDim.x + threadIdx.x) + 2] = a;
}
int main() {
  run<0>("gelu17 (current)"); run<1>("gelu15 sat scalar"); run<2>("gelu15 sat FFMA2"); run<3>("8x FFMA2 per pair"); run<4>("8x FFMA imm per elem");
  const int n = 1 << 16; float* d; cudaMalloc(&d, 3 * n * 4); acc<<<n / 256, 256>>>(d); float* h = new float[3 * n]; cudaMemcpy(h, d, 3 * n * 4, cudaMemcpyDeviceToHost);
  double e[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) { double x = -6.0 + 12.0 * i / n; double r = 0.5 * x * (1 + erf(x / sqrt(2.0))); for (int j = 0; j < 3; ++j) e[j] = fmax(e[j], fabs(h[3 * i + j] - r)); }
  printf("max abs err vs exact: gelu17 %.3g  gelu15_sat %.3g  gelu15_x2 %.3g\n", e[0], e[1], e[2]);
  return 0;
}
