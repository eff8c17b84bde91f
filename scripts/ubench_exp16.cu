// Micro-benchmark (debug aid): MUFU throughput of ex2.approx.f16x2 (two exponentials per instruction) against ex2.approx.ftz.f32,
// alone and inside the instruction mix of the attention kernel's exp loop (scale, pack, exp, row sum).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/ubench_exp16.cu -o scripts/_bin/ubench_exp16
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned pack_h2(float lo, float hi) { unsigned y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ unsigned hadd2(unsigned a, unsigned b) { unsigned y; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16]; unsigned h[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = -(threadIdx.x * 0.001f + i * 0.01f);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack_h2(a[2 * i], a[2 * i + 1]);
  float c = 0.999f, d = -0.001f;
  unsigned acc0 = 0, acc1 = 0; float s0 = 0.f, s1 = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]);                      // 16 exps: 16 MUFU.f32
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = ex2h2(h[i]);                     // 16 exps: 8 MUFU.f16x2
    } else if (MODE == 2) {                                              // current loop: FFMA, MUFU, FADD per element, one pack per pair
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float p0 = ex2(fmaf(a[i], c, d)), p1 = ex2(fmaf(a[i + 1], c, d));
        s0 += p0; s1 += p1;
        __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
        a[i] = a[i] * 0.5f + __uint_as_float(*reinterpret_cast<unsigned*>(&b) & 0x3f800000u) * 1e-30f; a[i + 1] = a[i + 1] * 0.5f - 1e-3f;
      }
    } else if (MODE == 3) {                                              // f16x2 loop: FFMA x2, pack, MUFU.f16x2, HADD2 per pair
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned x = pack_h2(fmaf(a[i], c, d), fmaf(a[i + 1], c, d));
        unsigned p = ex2h2(x);
        if ((i & 2) == 0) acc0 = hadd2(acc0, p); else acc1 = hadd2(acc1, p);
        a[i] = a[i] * 0.5f + __uint_as_float(p & 0x00010001u) * 1e-30f; a[i + 1] = a[i + 1] * 0.5f - 1e-3f;
      }
    }
  }
  long long t1 = clock64();
  float s = s0 + s1 + __uint_as_float(acc0) + __uint_as_float(acc1);
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float(h[i]);
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
  printf("%-44s warps/SM %2d: %.2f cycles per exponential per warp; %.2f cycles per exponential per SMSP\n", name, warps, per, per / (warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8}) {
    if (w == 4) { run<0>("MUFU.EX2 f32", 4); run<1>("MUFU.EX2 f16x2", 4); run<2>("loop f32: FFMA+MUFU+FADD (+pack/2)", 4); run<3>("loop f16x2: FFMA+pack/2+MUFU/2+HADD2/2", 4); }
    if (w == 8) { run<0>("MUFU.EX2 f32", 8); run<1>("MUFU.EX2 f16x2", 8); run<2>("loop f32: FFMA+MUFU+FADD (+pack/2)", 8); run<3>("loop f16x2: FFMA+pack/2+MUFU/2+HADD2/2", 8); }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
