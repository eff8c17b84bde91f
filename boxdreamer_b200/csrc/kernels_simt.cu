// fp32 "exact" compute path (SIMT GEMM + attention, used for the parity gate against the fp32 oracle)
// and the memory-bound glue kernels shared by both precision modes (LayerNorm, im2col, patchify,
// token fusion, query gather, unpatchify + sigmoid).
#include "bd_internal.h"
#include "common.cuh"

namespace bd {

// ---------------------------------------------------------------------------------------------
// fp32 SIMT GEMM: out = epi(A[M,K] . W[N,K]^T + bias)

static constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__device__ __forceinline__ float gelu_erf_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int EPI>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, int M, int N,
                                                       int K, GemmEpi e) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = threadIdx.x >> 2;        // 0..63 tile row
  const int lk = (threadIdx.x & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      const int ra = m0 + lr, rb = n0 + lr;
      As[lk + i][lr] = (ra < M && k < K) ? A[static_cast<long long>(ra) * K + k] : 0.f;
      Bs[lk + i][lr] = (rb < N && k < K) ? W[static_cast<long long>(rb) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = acc[i][j] + e.bias[n];
      if constexpr (EPI == EPI_F32) {
        long long orow = m;
        float add = 0.f;
        if (e.rp_in > 0) {
          const int tr = m % e.rp_in;
          orow = static_cast<long long>(m / e.rp_in) * e.rp_out + e.rp_off + tr;
          if (e.addtab) add = e.addtab[static_cast<long long>(tr) * N + n];
        }
        e.out_f32[orow * e.ldo + n] = v + add;
      } else if constexpr (EPI == EPI_RESID) {
        float* dst = e.out_f32 + static_cast<long long>(m) * e.ldo + n;
        const float g = e.gamma ? e.gamma[n] : 1.f;
        *dst = *dst + g * v;
      } else if constexpr (EPI == EPI_GELU) {
        reinterpret_cast<float*>(e.out_act)[static_cast<long long>(m) * N + n] = gelu_erf_exact(v);
      } else {
        reinterpret_cast<float*>(e.out_act)[static_cast<long long>(m) * N + n] = v;
      }
    }
  }
}

__global__ void qkv_split_f32_kernel(const float* __restrict__ qkv, GemmEpi e, int M) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int per_row = 3 * e.heads;
  if (idx >= static_cast<long long>(M) * per_row) return;
  const int m = static_cast<int>(idx / per_row);
  const int wh = static_cast<int>(idx % per_row);
  const int which = wh / e.heads, head = wh % e.heads;
  const int hd = e.head_dim, d = e.heads * hd;
  const float* src = qkv + static_cast<long long>(m) * 3 * d + which * d + head * hd;
  const int l = m / e.seq, tok = m % e.seq;
  float* dst = reinterpret_cast<float*>(which == 0 ? e.q : (which == 1 ? e.k : e.v)) +
               ((static_cast<long long>(l) * e.heads + head) * e.seq_pad + tok) * hd;
  const float* nw = which == 0 ? e.q_norm_w : (which == 1 ? e.k_norm_w : nullptr);
  if (nw != nullptr) {
    float ss = 0.f;
    for (int j = 0; j < hd; ++j) ss = fmaf(src[j], src[j], ss);
    const float r = rsqrtf(ss / hd + e.rms_eps);
    for (int j = 0; j < hd; ++j) dst[j] = nw[j] * (src[j] * r);
  } else {
    for (int j = 0; j < hd; ++j) dst[j] = src[j];
  }
}

cudaError_t qkv_split_f32(const float* qkv, const GemmEpi& e, int M, cudaStream_t s) {
  const long long n = static_cast<long long>(M) * 3 * e.heads;
  qkv_split_f32_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(qkv, e, M);
  return cudaGetLastError();
}

cudaError_t gemm_f32(const float* A, const float* W, int M, int N, int K, int epi, const GemmEpi& e, cudaStream_t s) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  switch (epi) {
    case EPI_F32: gemm_f32_kernel<EPI_F32><<<grid, 256, 0, s>>>(A, W, M, N, K, e); break;
    case EPI_RESID: gemm_f32_kernel<EPI_RESID><<<grid, 256, 0, s>>>(A, W, M, N, K, e); break;
    case EPI_GELU: gemm_f32_kernel<EPI_GELU><<<grid, 256, 0, s>>>(A, W, M, N, K, e); break;
    case EPI_ACT: gemm_f32_kernel<EPI_ACT><<<grid, 256, 0, s>>>(A, W, M, N, K, e); break;
    case EPI_QKV: {
      // raw qkv into the caller-provided scratch (e.out_act, [M, 3d] fp32), then norm + split
      gemm_f32_kernel<EPI_ACT><<<grid, 256, 0, s>>>(A, W, M, N, K, e);
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess) return err;
      return qkv_split_f32(reinterpret_cast<const float*>(e.out_act), e, M, s);
    }
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// fp32 attention: one CTA per (query row, bh)

__global__ void __launch_bounds__(128) attention_f32_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                            const float* __restrict__ V, float* __restrict__ O, int heads,
                                                            int hd, int seq, int seq_pad, float scale) {
  extern __shared__ float sm[];
  float* sc = sm;            // [seq]
  float* qs = sm + seq;      // [hd]
  __shared__ float red[4];
  const int row = blockIdx.x, bh = blockIdx.y;
  const int tid = threadIdx.x;
  const float* q = Q + (static_cast<long long>(bh) * seq_pad + row) * hd;
  for (int j = tid; j < hd; j += 128) qs[j] = q[j] * scale;
  __syncthreads();
  float mx = -INFINITY;
  for (int key = tid; key < seq; key += 128) {
    const float* kp = K + (static_cast<long long>(bh) * seq_pad + key) * hd;
    float s = 0.f;
    for (int j = 0; j < hd; ++j) s = fmaf(qs[j], kp[j], s);
    sc[key] = s;
    mx = fmaxf(mx, s);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int key = tid; key < seq; key += 128) {
    const float p = expf(sc[key] - mx);
    sc[key] = p;
    sum += p;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  sum = red[0] + red[1] + red[2] + red[3];
  const float inv = 1.0f / sum;
  const int l = bh / heads, head = bh % heads;
  for (int j = tid; j < hd; j += 128) {
    float acc = 0.f;
    const float* vp = V + static_cast<long long>(bh) * seq_pad * hd + j;
    for (int key = 0; key < seq; ++key) acc = fmaf(sc[key], vp[static_cast<long long>(key) * hd], acc);
    O[(static_cast<long long>(l) * seq + row) * (heads * hd) + head * hd + j] = acc * inv;
  }
}

cudaError_t attention_f32(const float* Q, const float* K, const float* V, float* O, int L, int heads, int head_dim, int seq,
                          int seq_pad, float scale, cudaStream_t s) {
  const size_t smem = (static_cast<size_t>(seq) + head_dim) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
  }
  dim3 grid(seq, L * heads);
  attention_f32_kernel<<<grid, 128, smem, s>>>(Q, K, V, O, heads, head_dim, seq, seq_pad, scale);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (one warp per row)

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps, float* __restrict__ out_f32,
                                                        bf16* __restrict__ out_bf16, int rows_out, int d, int rows_out_per,
                                                        int rows_in_per, int row_off) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_out) return;
  long long rin = r;
  if (rows_out_per > 0) rin = static_cast<long long>(r / rows_out_per) * rows_in_per + row_off + (r % rows_out_per);
  const float* xr = x + rin * d;
  float sum = 0.f;
  for (int j = lane; j < d; j += 32) sum += xr[j];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / d;
  float var = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float t = xr[j] - mean;
    var = fmaf(t, t, var);
  }
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / d + eps);
  for (int j = lane; j < d; j += 32) {
    float v = (xr[j] - mean) * rstd;
    if (w) v = v * w[j] + b[j];
    if (out_f32) out_f32[static_cast<long long>(r) * d + j] = v;
    if (out_bf16) out_bf16[static_cast<long long>(r) * d + j] = __float2bfloat16_rn(v);
  }
}

// d == 768 fast path: the row lives in registers (6 x float4 per lane): one HBM read, one write
__global__ void __launch_bounds__(256) layernorm768_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, float eps, float* __restrict__ out_f32,
                                                           bf16* __restrict__ out_bf16, int rows_out, int rows_out_per,
                                                           int rows_in_per, int row_off) {
  constexpr int D = 768;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_out) return;
  long long rin = r;
  if (rows_out_per > 0) rin = static_cast<long long>(r / rows_out_per) * rows_in_per + row_off + (r % rows_out_per);
  const float4* xr = reinterpret_cast<const float4*>(x + rin * D);
  float4 v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = xr[lane + 32 * i];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / D);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    var = fmaf(v[i].x, v[i].x, var); var = fmaf(v[i].y, v[i].y, var);
    var = fmaf(v[i].z, v[i].z, var); var = fmaf(v[i].w, v[i].w, var);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = (lane + 32 * i) * 4;
    float4 o = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);
    if (w) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + c));
      o.x = fmaf(o.x, w4.x, b4.x); o.y = fmaf(o.y, w4.y, b4.y); o.z = fmaf(o.z, w4.z, b4.z); o.w = fmaf(o.w, w4.w, b4.w);
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + static_cast<long long>(r) * D + c) = o;
    if (out_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(r) * D + c) = pk;
    }
  }
}

cudaError_t layernorm(const float* x, const float* w, const float* b, float eps, float* out_f32, bf16* out_bf16, int rows_out,
                      int d, int rows_out_per, int rows_in_per, int row_off, cudaStream_t s) {
  if (d == 768)
    layernorm768_kernel<<<(rows_out + 7) / 8, 256, 0, s>>>(x, w, b, eps, out_f32, out_bf16, rows_out, rows_out_per, rows_in_per,
                                                           row_off);
  else
    layernorm_kernel<<<(rows_out + 7) / 8, 256, 0, s>>>(x, w, b, eps, out_f32, out_bf16, rows_out, d, rows_out_per, rows_in_per,
                                                        row_off);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// im2col for the 14x14/s14 patch-embed conv with the ImageNet normalisation folded in
// (encoder/dinov2.py:45-46, DINOv2 layers/patch_embed.py:65-78) and patchify of the 8-channel corner heat maps
// (betr.py:211-228).  One CTA per (image, patch row): the C x 14 image rows are read coalesced into shared memory,
// then the 16 (or S/14) output token rows -- one contiguous block of the GEMM A operand -- are written coalesced.
//   MODE 0 (im2col)  : column = c*p*p + pr*p + pc   (conv weight order), value = (x - mean_c) / std_c, zero pad to kpad
//   MODE 1 (patchify): column = (pr*p + pc)*C + c   (einsum "nchpwq->nhwpqc")

template <typename TIn, typename TOut, int MODE>
__global__ void __launch_bounds__(256) patch_rows_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int C, int S, int patch,
                                                         int kout) {
  extern __shared__ unsigned char smem_raw[];
  TOut* tile = reinterpret_cast<TOut*>(smem_raw);
  const int g = S / patch;
  const int ph = blockIdx.x;
  const long long l = blockIdx.y;
  const int pitch = S + 2;  // de-conflicts the channel-strided reads of MODE 1
  const int nrows = C * patch;
  for (int idx = threadIdx.x; idx < nrows * S; idx += blockDim.x) {
    const int row = idx / S, x = idx % S;
    const int c = row / patch, pr = row % patch;
    float v = static_cast<float>(in[((l * C + c) * S + ph * patch + pr) * S + x]);
    if (MODE == 0) {
      const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
      const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
      v = (v - mean) / stdv;
    }
    tile[row * pitch + x] = static_cast<TOut>(v);
  }
  __syncthreads();
  TOut* dst = out + (l * g * g + static_cast<long long>(ph) * g) * kout;
  const int kreal = C * patch * patch;
  for (int idx = threadIdx.x; idx < g * kout; idx += blockDim.x) {
    const int pw = idx / kout, col = idx % kout;
    TOut v = static_cast<TOut>(0.f);
    if (col < kreal) {
      int c, pr, pc;
      if (MODE == 0) { c = col / (patch * patch); const int rem = col % (patch * patch); pr = rem / patch; pc = rem % patch; }
      else { c = col % C; const int pp = col / C; pr = pp / patch; pc = pp % patch; }
      v = tile[(c * patch + pr) * pitch + pw * patch + pc];
    }
    dst[idx] = v;
  }
}

// Fast path for 14-pixel patches with bf16 in/out (the tensor path): no run-time divisions, 16-byte global accesses.
// One CTA per (image, patch row).  Loads: warp per image row, lane per 8-pixel chunk.  Stores: MODE 1 writes the 8
// channels of one (token, pr, pc) as one 16-byte vector; MODE 0 writes bf16 pairs (pc, pc+1) -- a (c, pr) run of 14
// columns starts on a 4-byte boundary only.
template <int C, int MODE>
__global__ void __launch_bounds__(256) patch_rows14_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int S, int kout) {
  constexpr int P = 14;
  extern __shared__ unsigned char smem_raw[];
  bf16* tile = reinterpret_cast<bf16*>(smem_raw);
  const int g = S / P;
  const int ph = blockIdx.x;
  const long long l = blockIdx.y;
  const int pitch = S + 8;   // elements; rows stay 16-byte aligned, channel-strided reads spread over the banks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = S / 8;
  for (int row = warp; row < C * P; row += 8) {
    const int c = row / P, pr = row % P;
    const bf16* src = in + ((l * C + c) * S + ph * P + pr) * static_cast<long long>(S);
    for (int ch = lane; ch < chunks; ch += 32) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + ch);
      if (MODE == 0) {
        const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
        const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          h[i] = __floats2bfloat162_rn((f.x - mean) / stdv, (f.y - mean) / stdv);
        }
      }
      *reinterpret_cast<uint4*>(tile + row * pitch + ch * 8) = v;
    }
  }
  __syncthreads();
  bf16* dst = out + (l * g * g + static_cast<long long>(ph) * g) * kout;
  if (MODE == 1) {
    static_assert(MODE == 0 || C == 8, "patchify fast path: 8 channels = one 16-byte vector");
    const int items = g * P * P;   // (token, pr*14 + pc)
    for (int idx = threadIdx.x; idx < items; idx += 256) {
      const int pw = idx / (P * P), pp = idx % (P * P);
      const int pr = pp / P, pc = pp % P;
      const bf16* t = tile + pr * pitch + pw * P + pc;
      uint4 v;
      unsigned short* hv = reinterpret_cast<unsigned short*>(&v);
#pragma unroll
      for (int c = 0; c < 8; ++c) hv[c] = reinterpret_cast<const unsigned short*>(t)[c * P * pitch];
      *reinterpret_cast<uint4*>(dst + static_cast<long long>(pw) * kout + pp * 8) = v;
    }
  } else {
    const int pairs = kout / 2;
    const int real = C * P * (P / 2);   // bf16 pairs that hold pixels; the rest is the zero pad up to kout
    for (int idx = threadIdx.x; idx < g * pairs; idx += 256) {
      const int pw = idx / pairs, j = idx % pairs;
      uint32_t v = 0u;
      if (j < real) {
        const int c = j / (P * (P / 2)), rem = j % (P * (P / 2));
        const int pr = rem / (P / 2), pc = (rem % (P / 2)) * 2;
        v = *reinterpret_cast<const uint32_t*>(tile + (c * P + pr) * pitch + pw * P + pc);
      }
      *reinterpret_cast<uint32_t*>(dst + static_cast<long long>(pw) * kout + 2 * j) = v;
    }
  }
}

template <typename TIn, typename TOut, int MODE>
static cudaError_t launch_patch_rows(const void* in, void* out, int L, int C, int S, int patch, int kout, cudaStream_t s) {
  if constexpr (sizeof(TIn) == 2 && sizeof(TOut) == 2) {
    const bool aligned = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    if (patch == 14 && S % 8 == 0 && aligned && C == (MODE == 0 ? 3 : 8) && kout % 8 == 0) {
      const size_t smem14 = static_cast<size_t>(C) * 14 * (S + 8) * 2;
      auto k14 = patch_rows14_kernel<(MODE == 0 ? 3 : 8), MODE>;
      if (smem14 > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(k14, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem14));
        if (err != cudaSuccess) return err;
      }
      k14<<<dim3(S / 14, L), 256, smem14, s>>>(reinterpret_cast<const bf16*>(in), reinterpret_cast<bf16*>(out), S, kout);
      return cudaGetLastError();
    }
  }
  const size_t smem = static_cast<size_t>(C) * patch * (S + 2) * sizeof(TOut);
  auto kern = patch_rows_kernel<TIn, TOut, MODE>;
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
  }
  dim3 grid(S / patch, L);
  kern<<<grid, 256, smem, s>>>(reinterpret_cast<const TIn*>(in), reinterpret_cast<TOut*>(out), C, S, patch, kout);
  return cudaGetLastError();
}

template <int MODE>
static cudaError_t dispatch_patch_rows(const void* in, int in_bf16, void* out, int out_bf16, int L, int C, int S, int patch, int kout,
                                       cudaStream_t s) {
  if (in_bf16 && out_bf16) return launch_patch_rows<bf16, bf16, MODE>(in, out, L, C, S, patch, kout, s);
  if (in_bf16) return launch_patch_rows<bf16, float, MODE>(in, out, L, C, S, patch, kout, s);
  if (out_bf16) return launch_patch_rows<float, bf16, MODE>(in, out, L, C, S, patch, kout, s);
  return launch_patch_rows<float, float, MODE>(in, out, L, C, S, patch, kout, s);
}

cudaError_t im2col_patches(const void* images, int img_is_bf16, void* out, int out_is_bf16, int L, int S, int patch, int kpad,
                           cudaStream_t s) {
  return dispatch_patch_rows<0>(images, img_is_bf16, out, out_is_bf16, L, 3, S, patch, kpad, s);
}

cudaError_t patchify_heat(const void* feat, int in_is_bf16, void* out, int out_is_bf16, int L, int C, int S, int patch,
                          cudaStream_t s) {
  return dispatch_patch_rows<1>(feat, in_is_bf16, out, out_is_bf16, L, C, S, patch, patch * patch * C, s);
}

// DINOv2 prepare_tokens (vision_transformer.py:213-232): row 0 = cls + pos[0]; rows 1..n_reg = register tokens
__global__ void dino_prefix_kernel(float* __restrict__ X, const float* __restrict__ cls, const float* __restrict__ pos0,
                                   const float* __restrict__ reg, int L, int n_tok, int n_reg, int d) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(L) * (1 + n_reg) * d;
  if (idx >= total) return;
  const int j = static_cast<int>(idx % d);
  const int t = static_cast<int>((idx / d) % (1 + n_reg));
  const long long l = idx / (static_cast<long long>(d) * (1 + n_reg));
  const float v = (t == 0) ? (cls[j] + pos0[j]) : reg[(t - 1) * d + j];
  X[(l * n_tok + t) * d + j] = v;
}

cudaError_t dino_prefix_tokens(float* X, const float* cls, const float* pos0, const float* reg, int L, int n_tok, int n_reg,
                               int d, cudaStream_t s) {
  const long long total = static_cast<long long>(L) * (1 + n_reg) * d;
  dino_prefix_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(X, cls, pos0, reg, L, n_tok, n_reg, d);
  return cudaGetLastError();
}

// BETR token fusion (betr.py:282-295, 351-401): one warp per token
__global__ void __launch_bounds__(256) betr_fuse_kernel(const float* __restrict__ PF, const float* __restrict__ R,
                                                        const float* __restrict__ qtok, const float* __restrict__ pos,
                                                        const int64_t* __restrict__ qidx, float* __restrict__ X, int B, int T,
                                                        int P, int d, float eps) {
  const long long m = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= static_cast<long long>(B) * T * P) return;
  const int p = static_cast<int>(m % P);
  const int t = static_cast<int>((m / P) % T);
  const int b = static_cast<int>(m / (static_cast<long long>(P) * T));
  const bool is_q = (qidx[b] == t);
  const float* rr = R + m * d;
  float sum = 0.f;
  for (int j = lane; j < d; j += 32) sum += rr[j];
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / d;
  float var = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float tt = rr[j] - mean;
    var = fmaf(tt, tt, var);
  }
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / d + eps);
  const float* pf = is_q ? qtok : (PF + m * d);
  for (int j = lane; j < d; j += 32) {
    const float rgb = (rr[j] - mean) * rstd;
    X[m * d + j] = (pf[j] + rgb) + pos[static_cast<long long>(p) * d + j];
  }
}

cudaError_t betr_fuse(const float* PF, const float* R, const float* query_tok, const float* pos, const int64_t* query_idx,
                      float* X, int B, int T, int P, int d, float eps, cudaStream_t s) {
  const long long rows = static_cast<long long>(B) * T * P;
  betr_fuse_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(PF, R, query_tok, pos, query_idx, X, B, T, P, d, eps);
  return cudaGetLastError();
}

__global__ void gather_query_kernel(const float* __restrict__ X, const int64_t* __restrict__ qidx, float* __restrict__ of,
                                    bf16* __restrict__ ob, int B, int T, int P, int d) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * P * d;
  if (idx >= total) return;
  const int j = static_cast<int>(idx % d);
  const int p = static_cast<int>((idx / d) % P);
  const int b = static_cast<int>(idx / (static_cast<long long>(d) * P));
  const float v = X[((static_cast<long long>(b) * T + qidx[b]) * P + p) * d + j];
  if (of) of[idx] = v;
  if (ob) ob[idx] = __float2bfloat16_rn(v);
}

cudaError_t gather_query(const float* X, const int64_t* query_idx, float* out_f32, bf16* out_bf16, int B, int T, int P, int d,
                         cudaStream_t s) {
  const long long total = static_cast<long long>(B) * P * d;
  gather_query_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(X, query_idx, out_f32, out_bf16, B, T, P, d);
  return cudaGetLastError();
}

// unpatchify (betr.py:230-247) + sigmoid + 2x-1 (betr.py:432-435)
__device__ __forceinline__ float heat_from_logit(float l) {
  const float sg = 1.0f / (1.0f + expf(-l));
  return 2.0f * sg - 1.0f;
}
__global__ void unpatchify_sigmoid_kernel(const float* __restrict__ logits, float* __restrict__ heat, int B, int C, int S,
                                          int patch) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * S * S;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % S);
  const int y = static_cast<int>((idx / S) % S);
  const int c = static_cast<int>((idx / (static_cast<long long>(S) * S)) % C);
  const long long b = idx / (static_cast<long long>(S) * S * C);
  const int g = S / patch;
  const int tok = (y / patch) * g + (x / patch);
  const int j = ((y % patch) * patch + (x % patch)) * C + c;
  heat[idx] = heat_from_logit(logits[(b * g * g + tok) * (static_cast<long long>(patch) * patch * C) + j]);
}
// 8 channels: one thread per pixel reads its 8 channel logits as two 16-byte vectors (a warp covers whole 32-byte
// sectors) and writes one float into each of the 8 channel planes (a warp writes 128 contiguous bytes per plane).
__global__ void __launch_bounds__(256) unpatchify_sigmoid8_kernel(const float* __restrict__ logits, float* __restrict__ heat, int B, int S,
                                                                  int patch) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long plane = static_cast<long long>(S) * S;
  if (idx >= B * plane) return;
  const int x = static_cast<int>(idx % S);
  const int y = static_cast<int>((idx / S) % S);
  const long long b = idx / plane;
  const int g = S / patch;
  const int tok = (y / patch) * g + (x / patch);
  const int pp = (y % patch) * patch + (x % patch);
  const float4* src = reinterpret_cast<const float4*>(logits + (b * g * g + tok) * (static_cast<long long>(patch) * patch * 8) + pp * 8);
  const float4 a = __ldg(src), c4 = __ldg(src + 1);
  float* dst = heat + b * 8 * plane + static_cast<long long>(y) * S + x;
  dst[0 * plane] = heat_from_logit(a.x);
  dst[1 * plane] = heat_from_logit(a.y);
  dst[2 * plane] = heat_from_logit(a.z);
  dst[3 * plane] = heat_from_logit(a.w);
  dst[4 * plane] = heat_from_logit(c4.x);
  dst[5 * plane] = heat_from_logit(c4.y);
  dst[6 * plane] = heat_from_logit(c4.z);
  dst[7 * plane] = heat_from_logit(c4.w);
}

cudaError_t unpatchify_sigmoid(const float* logits, float* heat, int B, int C, int S, int patch, cudaStream_t s) {
  if (C == 8 && reinterpret_cast<uintptr_t>(logits) % 16 == 0) {
    const long long pixels = static_cast<long long>(B) * S * S;
    unpatchify_sigmoid8_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, s>>>(logits, heat, B, S, patch);
    return cudaGetLastError();
  }
  const long long total = static_cast<long long>(B) * C * S * S;
  unpatchify_sigmoid_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(logits, heat, B, C, S, patch);
  return cudaGetLastError();
}

// make_bbox_features(type="heatmap") of the dataset (src/datasets/utils/base/bbox_utils.py:263-303) on the device: one
// CTA per (view, corner).  heat = 2 * exp(-d / s) / max - 1 with d the pixel's distance to the projected corner,
// s = (|corner - centre of the 8 corners| / 10)^2 and `max` the maximum of corner i's un-normalised maps over the `group`
// views rasterised by one reference call (`bbox_map[..., i].max()`, :296; the dataset calls it per sample, group = T).
// exp is monotone and the pixel nearest to the corner minimises |dx| and |dy| separately, so a view's maximum is
// exp(-d(nearest pixel) / s): O(1) per view, no reduction pass.  Every operation is a separate IEEE fp32 operation in the
// reference's order (no FMA contraction); only expf may differ from the host's by an ulp.
__device__ __forceinline__ float heat_dist(float bx, float by, int x, int y) {
  const float dx = __fsub_rn(bx, static_cast<float>(x)), dy = __fsub_rn(by, static_cast<float>(y));
  return sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}
__device__ __forceinline__ float heat_scale(const float* c, int i) {   // c: the 8 corners of one view
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sx = __fadd_rn(sx, c[2 * k]); sy = __fadd_rn(sy, c[2 * k + 1]); }
  const float cx = __fdiv_rn(sx, 8.0f), cy = __fdiv_rn(sy, 8.0f);
  const float ex = __fsub_rn(cx, c[2 * i]), ey = __fsub_rn(cy, c[2 * i + 1]);
  const float q = __fdiv_rn(sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey))), 10.0f);
  return __fmul_rn(q, q);
}
__device__ __forceinline__ int heat_nearest(float b, int S) {
  const float r = fminf(fmaxf(nearbyintf(b), 0.0f), static_cast<float>(S - 1));
  return static_cast<int>(r);
}
template <typename TOut>
__global__ void __launch_bounds__(256) bbox_heatmap_kernel(const float* __restrict__ corners, TOut* __restrict__ out, int S, int group) {
  const int i = blockIdx.x;
  const long long l = blockIdx.y;
  const long long g0 = (l / group) * group;
  float vmax = 0.0f;
  for (int v = 0; v < group; ++v) {
    const float* cv = corners + (g0 + v) * 16;
    const float bxv = cv[2 * i], byv = cv[2 * i + 1];
    const float dmin = heat_dist(bxv, byv, heat_nearest(bxv, S), heat_nearest(byv, S));
    vmax = fmaxf(vmax, expf(__fdiv_rn(-dmin, heat_scale(cv, i))));
  }
  const float* c = corners + l * 16;
  const float bx = c[2 * i], by = c[2 * i + 1];
  // per pixel: distance, one multiply, ex2, one FMA.  The fast forms (rsqrt-based distance, __expf, reciprocals hoisted out
  // of the loop) stay within 1e-6 of the reference's IEEE sequence on values in [-1, 1] (the stated tolerance is 4e-6);
  // the normaliser vmax above uses the exact sequence.
  const float neg_inv_s = __fdiv_rn(-1.0f, heat_scale(c, i));
  const float k = __fdiv_rn(2.0f, vmax);
  const int n = S * S;
  TOut* dst = out + (l * 8 + i) * n;
  int x = threadIdx.x % S, y = threadIdx.x / S;
  const int step_y = 256 / S, step_x = 256 % S;
  for (int p = threadIdx.x; p < n; p += 256) {
    const float dx = bx - static_cast<float>(x), dy = by - static_cast<float>(y);
    const float d2 = fmaf(dx, dx, dy * dy);
    const float d = d2 > 0.0f ? d2 * rsqrtf(d2) : 0.0f;
    dst[p] = static_cast<TOut>(fmaf(__expf(d * neg_inv_s), k, -1.0f));
    x += step_x; y += step_y;
    if (x >= S) { x -= S; ++y; }
  }
}

cudaError_t bbox_heatmaps(const float* corners_px, void* out, int out_is_bf16, int L, int S, int group, cudaStream_t s) {
  if (L <= 0) return cudaSuccess;
  if (group <= 0 || L % group != 0) return cudaErrorInvalidValue;
  if (out_is_bf16) bbox_heatmap_kernel<bf16><<<dim3(8, L), 256, 0, s>>>(corners_px, reinterpret_cast<bf16*>(out), S, group);
  else bbox_heatmap_kernel<float><<<dim3(8, L), 256, 0, s>>>(corners_px, reinterpret_cast<float*>(out), S, group);
  return cudaGetLastError();
}

__global__ void cast_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx < n) out[idx] = __float2bfloat16_rn(in[idx]);
}

cudaError_t cast_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t s) {
  cast_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace bd
