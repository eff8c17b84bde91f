// FlashAttention-style non-causal softmax attention on tcgen05 tensor cores (sm_100a).
//
//   O[l, tok, h*HD:(h+1)*HD] = softmax(scale * Q K^T) V          per (image/sample l, head h)
//
// Replaces flash_attn_func / F.scaled_dot_product_attention at blocks.py:259-285 (decoder, 8 x 96, N = T*P)
// and the naive softmax attention of DINOv2 layers/attention.py:56-69 (12 x 64, N = 261).
//
// Operand layouts (written by the QKV GEMM epilogue, gemm_tc.cu EPI_QKV):
//   Q, K : bf16 [BH, seq_pad, HD]   (q, k already RMS-normalised where the model asks for it)
//   V^T  : bf16 [BH, HD, seq_pad]   so that both MMAs take K-major operands
// One CTA = one 128-row query tile of one (l, h); 6 warps:
//   warp 0     TMA producer (Q once; K and V^T tiles, 128B swizzle; zero OOB fill pads HD 96 -> 128)
//   warp 1     TMEM allocator + MMA issuer: S = Q K^T (128 x BKV, fp32 in TMEM), O += P V (128 x HD)
//   warps 2-5  softmax (thread == query row): online max/sum in fp32, exp2 with folded scale, P -> bf16,
//              P handed to the tensor core either through TMEM (TS MMA, variant 1) or swizzled smem (SS, variant 0);
//              O is rescaled in TMEM only when a row maximum moved; final 1/sum and bf16 store.
// Two CTAs are co-resident per SM (256 TMEM columns, <= 113 KB smem each) so one CTA's softmax overlaps the
// other's MMAs.
#include "bd_internal.h"
#include "common.cuh"

namespace bd {

bool get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br);

static constexpr int ATT_THREADS = 192;
static constexpr int BQ = 128;

template <int HD, int BKV, bool PTMEM>
struct AttCfg {
  static constexpr int NQS = (HD + 63) / 64;            // 64-column sub-tiles of Q / K (HD padded to NQS*64 by TMA zero fill)
  static constexpr int NKS = BKV / 64;                  // 64-key sub-tiles of V^T / P
  static constexpr int Q_SUB = BQ * 128;                // bytes per [128 x 64] bf16 sub-tile
  static constexpr int K_SUB = BKV * 128;
  static constexpr int V_SUB = HD * 128;
  static constexpr int Q_BYTES = NQS * Q_SUB;
  static constexpr int K_BYTES = NQS * K_SUB;
  static constexpr int V_BYTES = NKS * V_SUB;
  static constexpr int P_BYTES = PTMEM ? 0 : NKS * Q_SUB;
  static constexpr int BAR_BYTES = 128;
  static constexpr int SMEM_BYTES = Q_BYTES + K_BYTES + V_BYTES + P_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int S_COL = 0;
  static constexpr int O_COL = 128;
  static constexpr int P_COL = 0;  // aliases the first BKV/2 columns of S
};

struct AttArgs {
  bf16* O;
  int heads, seq, seq_pad;
  float scale_log2;  // softmax scale * log2(e)
};

template <int HD, int BKV, bool PTMEM>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttArgs args) {
  using Cfg = AttCfg<HD, BKV, PTMEM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + Cfg::K_BYTES;
  uint8_t* sP = sV + Cfg::V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int bh = blockIdx.y;
  const int seq = args.seq, seq_pad = args.seq_pad;
  const int n_kv = (seq + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast through a shuffle so the compiler keeps the TMEM base (and everything derived from it) in uniform
  // registers: otherwise every tcgen05.mma is wrapped in an R2UR "waterfall" loop that costs ~100 cycles per issue
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int s = 0; s < Cfg::NQS; ++s) tma_load_2d(sQ + s * Cfg::Q_SUB, &tmQ, q_full, s * 64, bh * seq_pad + q0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(k_empty, ph ^ 1);
        mbar_expect_tx(k_full, Cfg::K_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::NQS; ++s) tma_load_2d(sK + s * Cfg::K_SUB, &tmK, k_full, s * 64, bh * seq_pad + j * BKV);
        mbar_wait(v_empty, ph ^ 1);
        mbar_expect_tx(v_full, Cfg::V_BYTES);
#pragma unroll
        for (int s = 0; s < Cfg::NKS; ++s) tma_load_2d(sV + s * Cfg::V_SUB, &tmV, v_full, j * BKV + s * 64, bh * HD);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD);
      const uint32_t d_s = tmem_base + Cfg::S_COL;
      const uint32_t d_o = tmem_base + Cfg::O_COL;
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        // ---- S = Q K_j^T ----
        mbar_wait(k_full, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sQ + (k / 4) * Cfg::Q_SUB)) + 2 * (k % 4);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sK + (k / 4) * Cfg::K_SUB)) + 2 * (k % 4);
          umma_ss_bf16(d_s, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(k_empty);
        umma_commit(s_full);
        // ---- O (+)= P_j V_j ----
        mbar_wait(v_full, ph);
        mbar_wait(p_full, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sV + (k / 4) * Cfg::V_SUB)) + 2 * (k % 4);
          const uint32_t accum = (j != 0 || k != 0) ? 1u : 0u;
          if constexpr (PTMEM) {
            umma_ts_bf16(d_o, tmem_base + Cfg::P_COL + k * 8, bdesc, idesc_o, accum);
          } else {
            const uint64_t adesc = make_smem_desc_sw128(smem_u32(sP + (k / 4) * Cfg::Q_SUB)) + 2 * (k % 4);
            umma_ss_bf16(d_o, adesc, bdesc, idesc_o, accum);
          }
        }
        umma_commit(v_empty);
        if (j == n_kv - 1) umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (thread == query row) =====================
    const int quad = warp & 3;              // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;         // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + Cfg::S_COL;
    const uint32_t t_o = tmem_base + lane_addr + Cfg::O_COL;
    const uint32_t t_p = tmem_base + lane_addr + Cfg::P_COL;
    const float c = args.scale_log2;
    float m_run = -INFINITY;
    float l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int kv0 = j * BKV;
      const bool tail = kv0 + BKV > seq;
      mbar_wait(s_full, ph);
      tc_fence_after();
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int ch = 0; ch < BKV / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_s + ch * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (tail && kv0 + ch * 32 + i >= seq) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * c);  // 0 on the first tile (m_run = -inf)
      const float mc = m_new * c;
      // O rescale (s_full(j) implies PV_{j-1} has completed, so O is stable)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
        for (int ch = 0; ch < HD / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_o + ch * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x32(t_o + ch * 32, v);
        }
      }
      // pass 2: P = exp2(s*c - m*c), row sum, hand P to the tensor core
      float psum = 0.f;
#pragma unroll
      for (int ch = 0; ch < BKV / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_s + ch * 32, v);
        tmem_wait_ld();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
          float p0 = exp2f(fmaf(s0, c, -mc));
          float p1 = exp2f(fmaf(s1, c, -mc));
          if (tail) {
            if (kv0 + ch * 32 + 2 * i >= seq) p0 = 0.f;
            if (kv0 + ch * 32 + 2 * i + 1 >= seq) p1 = 0.f;
          }
          psum += p0 + p1;
          w[i] = pack_bf16x2(p0, p1);
        }
        if constexpr (PTMEM) {
          // A operand in TMEM: row r = lane r, element k -> column k/2 (two bf16 per 32-bit column)
          tmem_st_32x32b_x16(t_p + ch * 16, w);
        } else {
          // K-major SWIZZLE_128B tile: 16-byte chunk index XOR (row % 8); 64 keys per 128-byte row
          const int sub = (ch * 32) / 64;
          const int chunk0 = ((ch * 32) % 64) / 8;  // first 16B chunk of these 32 keys inside the row
          uint8_t* rowp = sP + sub * Cfg::Q_SUB + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (chunk0 + q) ^ (r & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
          }
        }
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;
      if constexpr (PTMEM) {
        tmem_wait_st();
      } else {
        tmem_wait_st();            // O rescale stores
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int row = q0 + r;
    const int l_idx = bh / args.heads, head = bh % args.heads;
    bf16* dst = args.O + (static_cast<long long>(l_idx) * seq + row) * (args.heads * HD) + head * HD;
#pragma unroll
    for (int ch = 0; ch < HD / 32; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_o + ch * 32, v);
      tmem_wait_ld();
      if (row < seq) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
          o.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
          o.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
          o.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + ch * 32 + q * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int HD, int BKV, bool PTMEM>
static cudaError_t launch_att(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int seq, int seq_pad,
                              float scale, cudaStream_t s) {
  using Cfg = AttCfg<HD, BKV, PTMEM>;
  const int BH = L * heads;
  CUtensorMap tq, tk, tv;
  if (!get_tmap_2d_bf16(&tq, Q, static_cast<uint64_t>(BH) * seq_pad, HD, HD, 64, BQ)) return cudaErrorInvalidValue;
  if (!get_tmap_2d_bf16(&tk, K, static_cast<uint64_t>(BH) * seq_pad, HD, HD, 64, BKV)) return cudaErrorInvalidValue;
  if (!get_tmap_2d_bf16(&tv, Vt, static_cast<uint64_t>(BH) * HD, seq_pad, seq_pad, 64, HD)) return cudaErrorInvalidValue;
  auto kern = attn_tc_kernel<HD, BKV, PTMEM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  AttArgs a{O, heads, seq, seq_pad, scale * 1.4426950408889634f};
  dim3 grid((seq + BQ - 1) / BQ, BH);
  kern<<<grid, ATT_THREADS, Cfg::SMEM_BYTES, s>>>(tq, tk, tv, a);
  return cudaGetLastError();
}

cudaError_t attention_tc(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int head_dim, int seq,
                         int seq_pad, float scale, int variant, cudaStream_t s) {
  if (seq_pad % 128 != 0 || seq > seq_pad) return cudaErrorInvalidValue;
  if (variant == 2) return attention_tc2(Q, K, Vt, O, L, heads, head_dim, seq, seq_pad, scale, s);
  if (head_dim == 96) {
    if (variant == 1) return launch_att<96, 128, true>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, s);
    return launch_att<96, 64, false>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, s);
  }
  if (head_dim == 64) {
    if (variant == 1) return launch_att<64, 128, true>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, s);
    return launch_att<64, 128, false>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace bd
