// Micro-benchmark (debug aid): per-element cost of exp2 on the MUFU pipe, as a Cody-Waite + degree-3 polynomial on the FMA
// pipe (scalar and packed f32x2), and of 3:1 / 2:1 / 1:1 mixes of the two -- 32 independent elements per thread, the
// softmax loop's surrounding work (scale FFMA, row-sum FADD, bf16 pack) included so the numbers transfer to the kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I boxdreamer_b200/csrc -o scripts/_bin/ubench_exp scripts/ubench_exp.cu
#include <cstdio>
#include "common.cuh"
using namespace bd;

__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float magic = 12582912.0f;
  const float t = x + magic;
  const float f = x - (t - magic);
  float p = fmaf(f, 0.0558011f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
// two elements at once on the packed-fp32 pipe
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& r0, float& r1) {
  x0 = fmaxf(x0, -126.0f); x1 = fmaxf(x1, -126.0f);
  const f32x2 magic = pack_f32x2(12582912.0f, 12582912.0f), nmagic = pack_f32x2(-12582912.0f, -12582912.0f);
  const f32x2 x = pack_f32x2(x0, x1);
  const f32x2 t = add_f32x2(x, magic);
  const f32x2 n = add_f32x2(t, nmagic);
  const f32x2 f = fma_f32x2(n, pack_f32x2(-1.0f, -1.0f), x);
  f32x2 p = fma_f32x2(f, pack_f32x2(0.0558011f, 0.0558011f), pack_f32x2(0.2402265f, 0.2402265f));
  p = fma_f32x2(p, f, pack_f32x2(0.6931472f, 0.6931472f));
  p = fma_f32x2(p, f, pack_f32x2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack_f32x2(p, p0, p1); unpack_f32x2(t, t0, t1);
  r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// MODE 0: all MUFU   1: all poly scalar   2: all poly packed   3: k of every 4 on poly (scalar)   4: k of every 4 pairs... packed
template <int MODE, int K>
__global__ void k(float* out, long long* cyc, int iters, float c, float nmc) {
  constexpr int NE = 32;
  float s[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) s[i] = -0.01f * ((threadIdx.x * 7 + i * 13) % 97);
  float sum0 = 0.f, sum1 = 0.f;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NE; i += 8) {   // 8 elements = 4 pairs per group
      float p[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float x0 = fmaf(s[i + j], c, nmc), x1 = fmaf(s[i + j + 1], c, nmc);
        bool poly;
        if (MODE == 0) poly = false;
        else if (MODE == 1 || MODE == 2) poly = true;
        else poly = (j / 2) < K;      // K of the 4 pairs
        if (!poly) { p[j] = ex2_approx(x0); p[j + 1] = ex2_approx(x1); }
        else if (MODE == 1 || MODE == 3) { p[j] = ex2_poly(x0); p[j + 1] = ex2_poly(x1); }
        else ex2_poly2(x0, x1, p[j], p[j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        sum0 += p[j]; sum1 += p[j + 1];
        acc ^= pack_bf16x2(p[j], p[j + 1]);
      }
    }
    nmc += __uint_as_float(acc & 1u);   // keeps the loop body alive; stays ~0
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum0 + sum1 + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int K>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {4, 8}) {
    const int iters = 4000;
    k<MODE, K><<<148, warps * 32>>>(out, cyc, iters, 0.14f, -0.3f); cudaDeviceSynchronize();
    k<MODE, K><<<148, warps * 32>>>(out, cyc, iters, 0.14f, -0.3f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    printf("%-34s %d warp(s)/SMSP: %6.2f cycles per element per warp, %6.2f per element per SMSP\n", name, warps / 4,
           (double)h[0] / (iters * 32.0), (double)h[0] / (iters * 32.0) / (warps / 4));
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 0>("all MUFU");
  run<1, 0>("all polynomial, scalar");
  run<2, 0>("all polynomial, packed f32x2");
  run<3, 1>("1 of 4 polynomial, scalar");
  run<4, 1>("1 of 4 polynomial, packed");
  run<3, 2>("2 of 4 polynomial, scalar");
  run<4, 2>("2 of 4 polynomial, packed");
  return 0;
}
