// Micro-benchmark (debug aid): cost of one softmax tile step of the attention kernel, without MMAs and barriers.
// Each of NW warps (4 = one softmax group, 8 = both groups, i.e. 1 or 2 warps per SM sub-partition) repeats:
//   tcgen05.ld S row (BKV fp32 columns) -> row max (3-input max) -> p = ex2(s*c - m*c), row sum, bf16 pack -> tcgen05.st P
// MODE 0: as in the kernel   MODE 1: no exp (max + pack only)   MODE 2: exp phase only (no TMEM traffic)
// MODE 3: POLY of every 4 exponentials evaluated on the FMA pipe (Cody-Waite + degree-3 polynomial) instead of MUFU
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I boxdreamer_b200/csrc -o scripts/_bin/ubench_softmax scripts/ubench_softmax.cu
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
using namespace bd;

__device__ __forceinline__ float ex2_poly(float x) {
  // 2^x for x <= 0: round-to-nearest split x = n + f, f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial
  // (rel. error ~1e-4, below bf16 resolution); exponent add through integer arithmetic.  x is clamped at -126.
  x = fmaxf(x, -126.0f);
  const float magic = 12582912.0f;  // 1.5 * 2^23
  const float t = x + magic;
  const float n = t - magic;
  const float f = x - n;
  float p = fmaf(f, 0.0558011f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

template <int MODE, int BKV, int POLY>
__global__ void __launch_bounds__(256, 1) k(int iters, float c, long long* cyc, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tptr, 0);
  const int g = warp >> 2, quad = warp & 3;
  const uint32_t t_s = tb + (static_cast<uint32_t>(quad * 32) << 16) + g * 256;
  const uint32_t t_p = t_s + BKV;
  {  // initialise S with finite values
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.01f * ((lane * 7 + i * 13) % 97));
    for (int ch = 0; ch < BKV / 32; ++ch) tmem_st_32x32b_x32(t_s + ch * 32, v);
    tmem_wait_st();
  }
  __syncthreads();
  float m_run = 0.f, l_run = 0.f;
  uint32_t sv[BKV];
  if (MODE == 2) {
#pragma unroll
    for (int ch = 0; ch < BKV / 32; ++ch) tmem_ld_32x32b_x32p(t_s + ch * 32, sv + ch * 32);
    tmem_wait_ld();
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE != 2) {
#pragma unroll
      for (int ch = 0; ch < BKV / 32; ++ch) tmem_ld_32x32b_x32p(t_s + ch * 32, sv + ch * 32);
      tmem_wait_ld();
    }
    float mx = m_run;
    if (MODE != 2) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < BKV; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
      }
      mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    }
    m_run = mx;
    const float nmc = -m_run * c;
    float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < BKV / 2; ++i) sv[i] = pack_bf16x2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1]));
    } else if (MODE == 4 || MODE == 5) {
      // packed-fp32 scale and row sum (FFMA2 / FADD2): half the FMA-pipe instructions of the scalar loop.  MODE 5: no row sum.
      const f32x2 c2 = pack_f32x2(c, c), n2 = pack_f32x2(nmc, nmc);
      f32x2 acc0 = pack_f32x2(0.f, 0.f), acc1 = pack_f32x2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < BKV / 2; i += 2) {
        float x0, x1, x2, x3;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, n2), x0, x1);
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * i + 2]), __uint_as_float(sv[2 * i + 3])), c2, n2), x2, x3);
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
        if (MODE == 4) { acc0 = add_f32x2(acc0, pack_f32x2(p0, p1)); acc1 = add_f32x2(acc1, pack_f32x2(p2, p3)); }
        sv[i] = pack_bf16x2(p0, p1);
        sv[i + 1] = pack_bf16x2(p2, p3);
      }
      float a, b, d, e;
      unpack_f32x2(acc0, a, b); unpack_f32x2(acc1, d, e);
      ps0 = a; ps1 = b; ps2 = d; ps3 = e;
    } else {
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int i = 0; i < BKV / 2; i += 2) {
        const float x0 = fmaf(__uint_as_float(sv[2 * i]), c, nmc), x1 = fmaf(__uint_as_float(sv[2 * i + 1]), c, nmc);
        const float x2 = fmaf(__uint_as_float(sv[2 * i + 2]), c, nmc), x3 = fmaf(__uint_as_float(sv[2 * i + 3]), c, nmc);
        const float p0 = (MODE == 3 && POLY >= 1) ? ex2_poly(x0) : ex2_approx(x0);
        const float p1 = (MODE == 3 && POLY >= 2) ? ex2_poly(x1) : ex2_approx(x1);
        const float p2 = (MODE == 3 && POLY >= 3) ? ex2_poly(x2) : ex2_approx(x2);
        const float p3 = (MODE == 3 && POLY >= 4) ? ex2_poly(x3) : ex2_approx(x3);
        ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
        pk[i] = pack_bf16x2(p0, p1);
        pk[i + 1] = pack_bf16x2(p2, p3);
      }
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < BKV / 2; ++i) sv[i] ^= pk[i] & 0x00010001u;   // feedback keeps the loop alive, values stay finite
      } else {
#pragma unroll
        for (int i = 0; i < BKV / 2; ++i) sv[i] = pk[i];
      }
    }
    l_run += (ps0 + ps1) + (ps2 + ps3);
    if (MODE != 2) {
      tmem_st_32x32b_x32p(t_p, sv);
      if (BKV == 128) tmem_st_32x32b_x32p(t_p + 32, sv + 32);
      else tmem_st_32x32b_x16(t_p + 32, sv + 32);
      tmem_wait_st();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float acc = l_run + m_run;
  if (MODE == 2) for (int i = 0; i < BKV / 2; ++i) acc += __uint_as_float(sv[i]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int MODE, int BKV, int POLY = 0>
static void run(const char* name, int nw) {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 2000;
  for (int r = 0; r < 2; ++r) {
    k<MODE, BKV, POLY><<<148, nw * 32>>>(iters, 0.14f, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  const double per = (double)h[0] / iters;
  printf("%-46s BKV %3d, %d warps/SMSP: %7.1f cycles per tile step = %5.2f per key per SMSP\n", name, BKV, nw / 4, per, per / BKV);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int nw : {4, 8}) {
    run<0, 96>("kernel loop (ld, max, exp, pack, st)", nw);
    run<1, 96>("no exp (ld, max, pack, st)", nw);
    run<2, 96>("exp phase only (fma, ex2, add, pack)", nw);
    run<3, 96, 1>("1 of 4 exps on the FMA pipe", nw);
    run<3, 96, 2>("2 of 4 exps on the FMA pipe", nw);
    run<3, 96, 4>("all exps on the FMA pipe", nw);
    run<4, 96>("packed FFMA2/FADD2 scale + row sum", nw);
    run<5, 96>("packed scale, no row sum (sum on the tensor core)", nw);
    run<0, 128>("kernel loop (ld, max, exp, pack, st)", nw);
    run<4, 128>("packed FFMA2/FADD2 scale + row sum", nw);
    run<2, 128>("exp phase only", nw);
    run<3, 128, 1>("1 of 4 exps on the FMA pipe", nw);
    run<3, 128, 2>("2 of 4 exps on the FMA pipe", nw);
  }
  return 0;
}
