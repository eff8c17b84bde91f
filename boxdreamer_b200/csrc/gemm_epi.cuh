// Shared pieces of the tcgen05 GEMM kernels (gemm_tc.cu: one CTA per tile; gemm_tc2.cu: CTA pairs, cta_group::2):
// tile constants, the argument block and the fused epilogue that drains one 128 x BN accumulator tile from TMEM.
#pragma once
#include "bd_internal.h"
#include "common.cuh"

namespace bd {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int GEMM_THREADS = 384;
static constexpr int EPI_WARP0 = 4;
static constexpr int N_EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = 4;
  static constexpr int STAGING_BYTES = N_EPI_WARPS * 32 * 32 * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static constexpr int TMEM_COLS = 512;  // 2 accumulator stages of BN (<= 256) fp32 columns
};

struct GemmArgs {
  int M, N, K;
  GemmEpi e;
  int qkv_col_base = 0;  // EPI_QKV on a column slice of the q|k|v weight: global column of this GEMM's column 0
  int m_fastest = 0;  // tile order: consecutive tiles walk M (swapped-operand V^T GEMM: the token block is shared)
};

__device__ __forceinline__ float gelu_erf_fast(float x) {
  // x * Phi(x) with Phi(x) - 0.5 = 0.5 erf(x / sqrt 2) as a degree-17 odd minimax polynomial on |x| <= 4 (clamped beyond,
  // where Phi - 0.5 = +-0.49997): max |gelu error| 2.2e-5 in fp32 Horner form -- two orders below bf16 resolution --
  // and no MUFU op, so the fc1 epilogue stays off the 16/clk/SM special-function pipe.
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

// Two elements at once on the packed-fp32 pipe (FFMA2): x * Phi(x), Phi(x) - 0.5 = x * q(min(x^2, 16)) with q a degree-7
// minimax polynomial in x^2 (max |gelu error| 3.8e-5 on |x| <= 4); beyond |x| = 4 the saturating fma clamps Phi to
// [0, 1] (error there <= 4 * (1 - Phi(4)) = 1.3e-4 at x = 4, i.e. 3e-5 relative).
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
  const f32x2 X = pack_f32x2(x0, x1);
  float s0, s1;
  unpack_f32x2(mul_f32x2(X, X), s0, s1);
  const f32x2 X2 = pack_f32x2(fminf(s0, 16.0f), fminf(s1, 16.0f));
  f32x2 p = pack_f32x2(-1.301277620e-09f, -1.301277620e-09f);
  p = fma_f32x2(p, X2, pack_f32x2(1.041950039e-07f, 1.041950039e-07f));
  p = fma_f32x2(p, X2, pack_f32x2(-3.657106863e-06f, -3.657106863e-06f));
  p = fma_f32x2(p, X2, pack_f32x2(7.485470993e-05f, 7.485470993e-05f));
  p = fma_f32x2(p, X2, pack_f32x2(-1.006488016e-03f, -1.006488016e-03f));
  p = fma_f32x2(p, X2, pack_f32x2(9.505389249e-03f, 9.505389249e-03f));
  p = fma_f32x2(p, X2, pack_f32x2(-6.588782661e-02f, -6.588782661e-02f));
  p = fma_f32x2(p, X2, pack_f32x2(3.986733839e-01f, 3.986733839e-01f));
  float p0, p1;
  unpack_f32x2(p, p0, p1);
  x0 = x0 * __saturatef(fmaf(p0, x0, 0.5f));
  x1 = x1 * __saturatef(fmaf(p1, x1, 0.5f));
}

// staging tile: 32 rows x 32 words (128 B per row).  Two access patterns share it:
//  (a) word-granular XOR swizzle (stage_write / stage_read): row-owner writes, row-wise 4-byte reads  (EPI_QKV)
//  (b) 16-byte-chunk XOR swizzle (stage_write16 / stage_read16): row-owner writes 8 x 16 B, then each lane reads a
//      16-byte chunk of 8 different rows -> 4 rows x 128 B per warp instruction, conflict-free both ways.
__device__ __forceinline__ void stage_write(uint32_t* tile, int lane, const uint32_t (&w)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) tile[lane * 32 + (j ^ lane)] = w[j];
}
__device__ __forceinline__ uint32_t stage_read(const uint32_t* tile, int rr, int lane) {
  return tile[rr * 32 + (lane ^ rr)];
}
__device__ __forceinline__ void stage_write16(uint32_t* tile, int lane, const uint32_t (&w)[32]) {
  uint4* row = reinterpret_cast<uint4*>(tile + lane * 32);
#pragma unroll
  for (int j = 0; j < 8; ++j) row[j ^ (lane & 7)] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}
__device__ __forceinline__ uint4 stage_read16(const uint32_t* tile, int row, int c4) {
  return reinterpret_cast<const uint4*>(tile + row * 32)[c4 ^ (row & 7)];
}

// Drains this warp's 32 rows x (column share) of one accumulator tile.
//   t_acc : TMEM address of the tile (lane quadrant and accumulator stage already applied)
//   row_w : global row of this warp's first TMEM lane;  n_blk : N-tile index;  grp : column group of this warp
template <int BN, int EPI, int HD>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& args, uint32_t* tile_s, uint32_t t_acc, int row_w, int n_blk,
                                                   int lane, int grp) {
  const GemmEpi& e = args.e;
  const int M = args.M, N = args.N;
  const int my_row = row_w + lane;
  if constexpr (EPI != EPI_QKV) {
    // 32-column chunks.  Row-owner phase: TMEM -> registers -> swizzled smem.  Row-wise phase: lane = (row group
    // rsub, 16-byte column chunk c4); 8 independent 16-byte global accesses per lane are in flight at once.
    long long my_out_row = my_row;
    int my_tab_row = 0;
    if (EPI == EPI_F32 && e.rp_in > 0) {
      my_tab_row = my_row % e.rp_in;
      my_out_row = static_cast<long long>(my_row / e.rp_in) * e.rp_out + e.rp_off + my_tab_row;
    }
    const int c4 = lane & 7, rsub = lane >> 3;
    constexpr int NCH = BN / 64;  // chunks of 32 columns per column-half
    for (int c = 0; c < NCH; ++c) {
      const int col0 = n_blk * BN + grp * (BN / 2) + c * 32;
      if (col0 >= N) continue;  // N tail: nothing to store (warp-uniform)
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_acc + grp * (BN / 2) + c * 32, v);
      tmem_wait_ld();
      stage_write16(tile_s, lane, v);
      __syncwarp();
      const int col = col0 + 4 * c4;
      const bool col_ok = col < N;  // N % 4 == 0: a 16-byte chunk is entirely in or out
      const float4 b4 = col_ok ? __ldg(reinterpret_cast<const float4*>(e.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 a[8];
      long long orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + rsub;
        const uint4 u = stage_read16(tile_s, row, c4);
        a[i] = make_float4(__uint_as_float(u.x) + b4.x, __uint_as_float(u.y) + b4.y, __uint_as_float(u.z) + b4.z,
                           __uint_as_float(u.w) + b4.w);
        orow[i] = __shfl_sync(0xffffffffu, my_out_row, row);
      }
      if constexpr (EPI == EPI_F32) {
        int trow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) trow[i] = __shfl_sync(0xffffffffu, my_tab_row, 4 * i + rsub);
        if (e.addtab != nullptr && col_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (row_w + 4 * i + rsub < M) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(e.addtab + static_cast<long long>(trow[i]) * N + col));
              a[i].x += t4.x; a[i].y += t4.y; a[i].z += t4.z; a[i].w += t4.w;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (row_w + 4 * i + rsub < M && col_ok) *reinterpret_cast<float4*>(e.out_f32 + orow[i] * e.ldo + col) = a[i];
        }
      } else if constexpr (EPI == EPI_RESID) {
        const float4 g4 = (e.gamma != nullptr && col_ok) ? __ldg(reinterpret_cast<const float4*>(e.gamma + col))
                                                          : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // all residual loads first (memory-level parallelism), then the stores
          r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_w + 4 * i + rsub < M && col_ok) r[i] = *reinterpret_cast<const float4*>(e.out_f32 + orow[i] * e.ldo + col);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (row_w + 4 * i + rsub < M && col_ok) {
            r[i].x = fmaf(g4.x, a[i].x, r[i].x); r[i].y = fmaf(g4.y, a[i].y, r[i].y);
            r[i].z = fmaf(g4.z, a[i].z, r[i].z); r[i].w = fmaf(g4.w, a[i].w, r[i].w);
            *reinterpret_cast<float4*>(e.out_f32 + orow[i] * e.ldo + col) = r[i];
          }
        }
      } else {  // EPI_GELU / EPI_ACT: bf16 store, 8 bytes per lane
        bf16* out = reinterpret_cast<bf16*>(e.out_act);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if constexpr (EPI == EPI_GELU) {
            a[i].x = gelu_erf_fast(a[i].x); a[i].y = gelu_erf_fast(a[i].y);
            a[i].z = gelu_erf_fast(a[i].z); a[i].w = gelu_erf_fast(a[i].w);
          }
          if (row_w + 4 * i + rsub < M && col_ok) {
            uint2 pk;
            pk.x = pack_bf16x2(a[i].x, a[i].y);
            pk.y = pack_bf16x2(a[i].z, a[i].w);
            *reinterpret_cast<uint2*>(out + orow[i] * N + col) = pk;
          }
        }
      }
      __syncwarp();
    }
  } else {  // EPI_QKV
    static_assert(EPI != EPI_QKV || (BN % HD == 0), "BN must hold whole heads");
    constexpr int UNITS = BN / HD;
    constexpr int WPR = HD / 2;  // packed words per row
    const int d_model = e.heads * HD;
    const int l = my_row / e.seq, tok = my_row % e.seq;
    const bool row_ok = my_row < M;
    // element offset of (l, head 0, tok, 0) in Q/K; the head term is added per unit
    const long long my_qk_off = (static_cast<long long>(l) * e.heads * e.seq_pad + tok) * HD;
    for (int u = grp; u < UNITS; u += 2) {
      const int col0 = n_blk * BN + u * HD;
      if (col0 >= N) continue;
      const int which = (args.qkv_col_base + col0) / d_model;            // 0 q, 1 k, 2 v
      const int head = ((args.qkv_col_base + col0) % d_model) / HD;
      float f[HD];
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_acc + u * HD + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[c * 32 + j] = __uint_as_float(v[j]) + __ldg(e.bias + col0 + c * 32 + j);
      }
      if (which < 2) {
        const float* nw = (which == 0) ? e.q_norm_w : e.k_norm_w;
        if (nw != nullptr) {
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < HD; ++j) ss = fmaf(f[j], f[j], ss);
          const float r = rsqrtf(ss * (1.0f / HD) + e.rms_eps);
#pragma unroll
          for (int j = 0; j < HD; ++j) f[j] = f[j] * r * __ldg(nw + j);
        }
        bf16* base = reinterpret_cast<bf16*>(which == 0 ? e.q : e.k);
        const long long head_off = static_cast<long long>(head) * e.seq_pad * HD;
#pragma unroll
        for (int piece = 0; piece < (WPR + 31) / 32; ++piece) {
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int wi = piece * 32 + j;
            w[j] = (wi < WPR) ? pack_bf16x2(f[2 * (wi < WPR ? wi : 0)], f[2 * (wi < WPR ? wi : 0) + 1]) : 0u;
          }
          stage_write(tile_s, lane, w);
          __syncwarp();
          const int wi = piece * 32 + lane;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t word = stage_read(tile_s, rr, lane);
            const long long off = __shfl_sync(0xffffffffu, my_qk_off, rr);
            if (row_w + rr < M && wi < WPR) {
              *reinterpret_cast<uint32_t*>(base + head_off + off + 2 * wi) = word;
            }
          }
          __syncwarp();
        }
      } else {
        // V^T [BH, HD, seq_pad]: lanes hold consecutive tokens -> coalesced along the key axis
        bf16* vt = reinterpret_cast<bf16*>(e.v);
        if (row_ok) {
          bf16* dst = vt + (static_cast<long long>(l) * e.heads + head) * HD * e.seq_pad + tok;
#pragma unroll
          for (int j = 0; j < HD; ++j) dst[static_cast<long long>(j) * e.seq_pad] = __float2bfloat16_rn(f[j]);
        }
      }
    }
  }
}

// TMA-out epilogue (EPI_GELU / EPI_ACT: bf16 store, EPI_RESID: fp32 reduce-add performed by the L2) for N % 64 == 0.
// Each lane owns one accumulator row: TMEM -> registers -> bias / GELU / gamma -> 128-byte swizzled staging row ->
// one elected TMA store of the 32-row box.  No per-element global address arithmetic, no residual load in the SM,
// M tail clipped by the tensor map.  sbuf: this warp's two 4 KB staging boxes (1024-byte aligned), buf: toggles.
template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile_tma(const GemmArgs& args, const CUtensorMap* tm_out, uint8_t* sbuf, uint32_t& buf,
                                                       uint32_t t_acc, int row_w, int n_blk, int lane, int grp) {
  static_assert(EPI == EPI_GELU || EPI == EPI_ACT || EPI == EPI_RESID, "TMA-out epilogue kinds");
  const GemmEpi& e = args.e;
  if (row_w >= args.M) return;  // whole 32-row slab beyond M (warp-uniform)
  constexpr int CW = (EPI == EPI_RESID) ? 32 : 64;  // columns per 128-byte staging row
  constexpr int NCH = (BN / 2) / CW;
  const int sw = lane & 7;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int colt = grp * (BN / 2) + c * CW;
    const int col0 = n_blk * BN + colt;
    if (col0 >= args.N) continue;
    uint8_t* sb = sbuf + buf * 4096;
    if (lane == 0) bulk_wait_group_read<1>();  // the store issued two chunks ago has finished reading this box
    __syncwarp();
    uint4* row = reinterpret_cast<uint4*>(sb + lane * 128);
    if constexpr (EPI == EPI_RESID) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_acc + colt, v);
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
      const float4* g4 = reinterpret_cast<const float4*>(e.gamma + col0);
      const bool has_g = e.gamma != nullptr;
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);  // warp-uniform address: one broadcast transaction
        float4 o = make_float4(__uint_as_float(v[4 * j]) + b.x, __uint_as_float(v[4 * j + 1]) + b.y,
                               __uint_as_float(v[4 * j + 2]) + b.z, __uint_as_float(v[4 * j + 3]) + b.w);
        if (has_g) {
          const float4 g = __ldg(g4 + j);
          o.x *= g.x; o.y *= g.y; o.z *= g.z; o.w *= g.w;
        }
        row[j ^ sw] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w));
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_acc + colt + h * 32, v);
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0 + h * 32);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 columns -> one 16-byte chunk of bf16
          const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
          float f0 = __uint_as_float(v[8 * j]) + ba.x, f1 = __uint_as_float(v[8 * j + 1]) + ba.y;
          float f2 = __uint_as_float(v[8 * j + 2]) + ba.z, f3 = __uint_as_float(v[8 * j + 3]) + ba.w;
          float f4 = __uint_as_float(v[8 * j + 4]) + bb.x, f5 = __uint_as_float(v[8 * j + 5]) + bb.y;
          float f6 = __uint_as_float(v[8 * j + 6]) + bb.z, f7 = __uint_as_float(v[8 * j + 7]) + bb.w;
          if constexpr (EPI == EPI_GELU) {
            gelu_erf_fast2(f0, f1); gelu_erf_fast2(f2, f3); gelu_erf_fast2(f4, f5); gelu_erf_fast2(f6, f7);
          }
          row[(h * 4 + j) ^ sw] = make_uint4(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3), pack_bf16x2(f4, f5), pack_bf16x2(f6, f7));
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if constexpr (EPI == EPI_RESID) tma_reduce_add_2d(tm_out, sb, col0, row_w);
      else tma_store_2d(tm_out, sb, col0, row_w);
      bulk_commit_group();
    }
    buf ^= 1;
  }
}

// q|k projection epilogue through TMA (CTA-pair kernel): per head, bias (+ RMSNorm over the head) in registers, then
// HD/32 staging boxes of 32 rows x 32 bf16 (64-byte rows, 64B swizzle) stored into Q or K [L*heads, seq(_pad), HD].
// A 32-row slab that crosses an image boundary (or the end of M) cannot be a TMA box (negative start coordinates fault
// on the device), so those slabs -- 1 in 8 for DINOv2's 261 tokens, none for the decoder -- store their rows directly,
// 4 x 16 bytes per lane and piece.  Pad rows are never written.  sbuf: this warp's four 2 KB boxes.
template <int BN, int HD>
__device__ __forceinline__ void gemm_epilogue_qk_tma(const GemmArgs& args, const CUtensorMap* tm_q, const CUtensorMap* tm_k,
                                                     uint8_t* sbuf, uint32_t& slot, uint32_t t_acc, int row_w, int n_blk, int lane,
                                                     int grp) {
  static_assert(BN % HD == 0 && HD % 32 == 0, "BN must hold whole heads");
  const GemmEpi& e = args.e;
  if (row_w >= args.M) return;
  constexpr int UNITS = BN / HD;
  const int d_model = e.heads * HD;
  const int l0 = row_w / e.seq, tok0 = row_w - l0 * e.seq;
  const bool inside = tok0 + 32 <= e.seq;  // whole slab within one image (then also < M); warp-uniform
  const int my_row = row_w + lane;
  const int my_l = my_row / e.seq, my_tok = my_row - my_l * e.seq;
  const int sw = (lane >> 1) & 3;
#pragma unroll 1
  for (int u = grp; u < UNITS; u += 2) {
    const int col0 = n_blk * BN + u * HD;
    if (col0 >= args.N) continue;
    const int which = col0 / d_model;  // 0 q, 1 k
    const int head = (col0 - which * d_model) / HD;
    float f[HD];
    {
      uint32_t v[HD];
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) tmem_ld_32x32b_x32p(t_acc + u * HD + c * 32, &v[c * 32]);
      tmem_wait_ld();
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
      for (int j = 0; j < HD / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        f[4 * j] = __uint_as_float(v[4 * j]) + b.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
        f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
      }
    }
    const float* nw = (which == 0) ? e.q_norm_w : e.k_norm_w;
    if (nw != nullptr) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < HD; ++j) ss = fmaf(f[j], f[j], ss);
      const float r = rsqrtf(ss * (1.0f / HD) + e.rms_eps);
      const float4* w4 = reinterpret_cast<const float4*>(nw);
#pragma unroll
      for (int j = 0; j < HD / 4; ++j) {
        const float4 w = __ldg(w4 + j);
        f[4 * j] = f[4 * j] * r * w.x; f[4 * j + 1] = f[4 * j + 1] * r * w.y;
        f[4 * j + 2] = f[4 * j + 2] * r * w.z; f[4 * j + 3] = f[4 * j + 3] * r * w.w;
      }
    }
    if (!inside) {
      if (my_row < args.M) {
        bf16* dst = reinterpret_cast<bf16*>(which == 0 ? e.q : e.k) +
                    ((static_cast<long long>(my_l) * e.heads + head) * e.seq_pad + my_tok) * HD;
#pragma unroll
        for (int j = 0; j < HD / 8; ++j) {
          const float* g = &f[8 * j];
          reinterpret_cast<uint4*>(dst)[j] =
              make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
        }
      }
      continue;
    }
    const CUtensorMap* tm = (which == 0) ? tm_q : tm_k;
#pragma unroll
    for (int p = 0; p < HD / 32; ++p) {
      uint8_t* sb = sbuf + slot * 2048;
      if (lane == 0) bulk_wait_group_read<3>();
      __syncwarp();
      uint4* row = reinterpret_cast<uint4*>(sb + lane * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* g = &f[p * 32 + 8 * j];
        row[j ^ sw] = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(tm, sb, p * 32, tok0, l0 * e.heads + head);
        bulk_commit_group();
      }
      slot = (slot + 1) & 3;
    }
  }
}

// V^T epilogue of the swapped-operand GEMM (rows = v features, columns = tokens): lane = one feature row, 64 tokens per
// 128-byte staging row, stored into V^T viewed as [L, d_model, seq(_pad)].  The innermost (token) coordinate of a TMA
// store must be 16-byte aligned and chunks must not cross images, so the dispatcher takes this path only for
// seq % 64 == 0 (the decoder); otherwise V goes through the row-major epilogue's strided V^T stores.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_vt_tma(const GemmArgs& args, const CUtensorMap* tm_v, uint8_t* sbuf, uint32_t& buf,
                                                     uint32_t t_acc, int row_w, int n_blk, int lane, int grp) {
  const GemmEpi& e = args.e;
  if (row_w >= args.M) return;
  const float b = (row_w + lane < args.M) ? __ldg(e.bias + row_w + lane) : 0.f;
  const int sw = lane & 7;
  constexpr int NCH = (BN / 2) / 64;
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int colt = grp * (BN / 2) + c * 64;
    const int n0 = n_blk * BN + colt;  // first token of the chunk (global index l * seq + tok)
    if (n0 >= args.N) continue;
    uint8_t* sb = sbuf + buf * 4096;
    if (lane == 0) bulk_wait_group_read<1>();
    __syncwarp();
    uint4* row = reinterpret_cast<uint4*>(sb + lane * 128);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_acc + colt + h * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        row[(h * 4 + j) ^ sw] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * j]) + b, __uint_as_float(v[8 * j + 1]) + b),
                                           pack_bf16x2(__uint_as_float(v[8 * j + 2]) + b, __uint_as_float(v[8 * j + 3]) + b),
                                           pack_bf16x2(__uint_as_float(v[8 * j + 4]) + b, __uint_as_float(v[8 * j + 5]) + b),
                                           pack_bf16x2(__uint_as_float(v[8 * j + 6]) + b, __uint_as_float(v[8 * j + 7]) + b));
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const int l = n0 / e.seq;
      tma_store_3d(tm_v, sb, n0 - l * e.seq, row_w, l);
      bulk_commit_group();
    }
    buf ^= 1;
  }
}

}  // namespace bd
