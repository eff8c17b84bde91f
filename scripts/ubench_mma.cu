// Micro-benchmark (debug aid): tensor-pipe cost of the tcgen05.mma shapes the attention kernel issues.
// One CTA per SM; warp 0 issues `groups` x `per_group` MMAs (operands: SW128 K-major smem tiles, or A from TMEM), one
// commit at the end, and reports cycles per MMA.  Data are zeros/garbage: only the issue/execute rate is of interest.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I boxdreamer_b200/csrc -o scripts/_bin/ubench_mma scripts/ubench_mma.cu
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
using namespace bd;

// MODE 0: SS chain (A,B smem), accumulator alternates between two TMEM regions per group
// MODE 1: TS chain (A tmem, B smem)
// MODE 2: attention order: S0 chain, S1 chain (SS, N = NS), PV0 chain, PV1 chain (TS, N = NO, KSTEPS_PV steps)
// MODE 3: SS, two accumulators interleaved step by step      MODE 4: TS, two accumulators interleaved
// MODE 5: SS / TS alternating step by step
// All shapes are compile-time so that the issue loop is straight-line UTCHMMA code (descriptors in uniform registers).
template <int MODE, int N, int KS, int NO, int KPV>
__global__ void __launch_bounds__(128, 1) k(int groups, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i * 2654435761u % 1024;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tptr, 0);
  if (warp == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024, v0 = smem_u32(smem) + 112 * 1024;
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, NO);
    constexpr int BSK = N * 128, VSK = NO * 128;
    long long t0 = clock64();
#pragma unroll 1
    for (int g = 0; g < groups; ++g) {
      if constexpr (MODE == 0) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
          umma_ss_bf16_w(tb + (g & 1) * 256, make_smem_desc_sw128(a0 + (kk / 4) * 16384) + 2 * (kk % 4),
                         make_smem_desc_sw128(b0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
      } else if constexpr (MODE == 1) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
          umma_ts_bf16_w(tb + (g & 1) * 256, tb + 128 + kk * 8, make_smem_desc_sw128(b0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
      } else if constexpr (MODE == 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int kk = 0; kk < KS; ++kk)
            umma_ss_bf16_w(tb + h * 256, make_smem_desc_sw128(a0 + h * 32768 + (kk / 4) * 16384) + 2 * (kk % 4),
                           make_smem_desc_sw128(b0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int kk = 0; kk < KPV; ++kk)
            umma_ts_bf16_w(tb + h * 256 + 256 - NO, tb + h * 256 + N + kk * 8, make_smem_desc_sw128(v0 + (kk / 4) * VSK) + 2 * (kk % 4), idesc_o, 1);
      } else if constexpr (MODE == 3) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            umma_ss_bf16_w(tb + h * 256, make_smem_desc_sw128(a0 + h * 32768 + (kk / 4) * 16384) + 2 * (kk % 4),
                           make_smem_desc_sw128(b0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
      } else if constexpr (MODE == 4) {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            umma_ts_bf16_w(tb + h * 256, tb + h * 256 + 128 + kk * 8, make_smem_desc_sw128(b0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
      } else {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          umma_ss_bf16_w(tb, make_smem_desc_sw128(a0 + (kk / 4) * 16384) + 2 * (kk % 4), make_smem_desc_sw128(b0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
          umma_ts_bf16_w(tb + 256, tb + 256 + 128 + kk * 8, make_smem_desc_sw128(v0 + (kk / 4) * BSK) + 2 * (kk % 4), idesc, kk != 0);
        }
      }
    }
    umma_commit_w(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int MODE, int N, int KS, int NO = 96, int KPV = 6>
static void run(const char* name, int groups = 512) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  auto kern = k<MODE, N, KS, NO, KPV>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int r = 0; r < 2; ++r) {
    kern<<<148, 128, 200 * 1024>>>(groups, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
  const double per_group = (double)h[0] / groups;
  double ideal;
  int n;
  if (MODE == 2) { n = 2 * KS + 2 * KPV; ideal = 2 * KS * (N / 2.0) + 2 * KPV * (NO / 2.0); }
  else if (MODE >= 3) { n = 2 * KS; ideal = n * (N / 2.0); }
  else { n = KS; ideal = n * (N / 2.0); }
  printf("%-52s %8.1f cycles/group of %2d MMAs = %6.1f per MMA (ideal %5.1f) -> %3.0f%% of the dense rate\n", name, per_group, n, per_group / n,
         ideal / n, 100.0 * ideal / per_group);
  cudaFree(cyc);
}

int main() {
  run<0, 256, 4>("SS N256 k4 (GEMM tile)");
  run<0, 192, 6>("SS N192 k6");
  run<0, 128, 6>("SS N128 k6");
  run<0, 96, 6>("SS N96 k6 (S, hd 96)");
  run<0, 128, 4>("SS N128 k4 (S, hd 64)");
  run<0, 64, 4>("SS N64 k4");
  run<1, 96, 6>("TS N96 k6 (PV hd 96, BKV 96)");
  run<1, 96, 8>("TS N96 k8 (PV hd 96, BKV 128)");
  run<1, 128, 8>("TS N128 k8");
  run<1, 64, 8>("TS N64 k8 (PV hd 64, BKV 128)");
  run<1, 256, 4>("TS N256 k4");
  run<3, 96, 6>("SS N96 k6 x2 interleaved accumulators");
  run<3, 128, 4>("SS N128 k4 x2 interleaved accumulators");
  run<4, 96, 6>("TS N96 k6 x2 interleaved accumulators");
  run<4, 64, 8>("TS N64 k8 x2 interleaved accumulators");
  run<5, 96, 6>("SS/TS N96 alternating");
  run<2, 96, 6, 96, 6>("attention hd96 BKV96 : 2x S(N96,k6) + 2x PV(N96,k6)");
  run<2, 128, 6, 96, 8>("attention hd96 BKV128: 2x S(N128,k6) + 2x PV(N96,k8)");
  run<2, 128, 4, 64, 8>("attention hd64 BKV128: 2x S(N128,k4) + 2x PV(N64,k8)");
  run<2, 64, 6, 96, 4>("attention hd96 BKV64 : 2x S(N64,k6) + 2x PV(N96,k4)");
  return 0;
}
