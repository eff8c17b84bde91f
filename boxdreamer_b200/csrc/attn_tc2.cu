// Attention v2: persistent, two query tiles per CTA in ping-pong (sm_100a, tcgen05 + TMEM + TMA).
//
// Same operand layouts and math as attn_tc.cu (Q,K [BH, seq_pad, HD]; V^T [BH, HD, seq_pad]; O token-major), but
//   * one persistent CTA per SM walks (bh, 256-row query pair) work items: TMEM (512 columns), barriers and tensor-map
//     prefetch are set up once, K/V tiles stream through an smem ring across items;
//   * two 128-row query tiles share every K/V tile; the MMA warp interleaves  S0, S1, PV0, S0', PV1, S1', ...  so one
//     tile's softmax (warps 2-5 / 6-9) overlaps the other tile's MMAs -- the tensor pipe only waits for the slower of
//     the MMA stream and the two softmax groups;
//   * P stays in tensor memory (TS-form PV MMA), aliased onto the first 64 columns of its S tile;
//   * the running maximum is only advanced (and O rescaled in TMEM) when it grows by more than 2^8 -- stale maxima are
//     exact because the same maximum scales P and the row sum;
//   * the last K/V tile is trimmed to the next multiple of 16 keys (runtime UMMA N / K), which removes most of the
//     padding waste of short sequences (DINOv2: 261 keys = 2 tiles + 16 keys instead of 3 tiles).
#include "bd_internal.h"
#include "common.cuh"

namespace bd {

bool get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br);

static constexpr int A2_THREADS = 320;
static constexpr int A2_BQ = 128;
static constexpr int A2_BKV = 128;

template <int HD>
struct Att2Cfg {
  static constexpr int NQS = (HD + 63) / 64;
  static constexpr int Q_TILE = NQS * A2_BQ * 128;
  static constexpr int K_TILE = NQS * A2_BKV * 128;
  static constexpr int V_SUB = HD * 128;
  static constexpr int V_TILE = 2 * V_SUB;
  static constexpr int NSTG = (HD > 64) ? 2 : 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 2 * Q_TILE + NSTG * (K_TILE + V_TILE) + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
};

struct Att2Args {
  bf16* O;
  int heads, seq, seq_pad, BH;
  float scale_log2;
};

__device__ __forceinline__ uint32_t s_col(int g) { return static_cast<uint32_t>(g * 128); }
__device__ __forceinline__ uint32_t o_col(int g) { return static_cast<uint32_t>(256 + g * 128); }

template <int HD>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const Att2Args args) {
  using Cfg = Att2Cfg<HD>;
  constexpr int NSTG = Cfg::NSTG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][Q_TILE]
  uint8_t* sK = sQ + 2 * Cfg::Q_TILE;                  // [NSTG][K_TILE]
  uint8_t* sV = sK + NSTG * Cfg::K_TILE;               // [NSTG][V_TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NSTG * Cfg::V_TILE);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* q_empty = bars + 2;             // [1]
  uint64_t* k_full = bars + 3;              // [NSTG]
  uint64_t* k_empty = k_full + NSTG;
  uint64_t* v_full = k_empty + NSTG;
  uint64_t* v_empty = v_full + NSTG;
  uint64_t* s_full = v_empty + NSTG;        // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_full = p_full + 2;            // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = args.seq, seq_pad = args.seq_pad;
  const int n_qt = (seq + A2_BQ - 1) / A2_BQ;            // query tiles per sequence
  const int n_pairs = (n_qt + 1) / 2;
  const int n_items = args.BH * n_pairs;
  const int n_kv = (seq + A2_BKV - 1) / A2_BKV;
  const int tail_keys = seq - (n_kv - 1) * A2_BKV;       // 1..128 valid keys in the last tile
  const int tail_cols = (tail_keys + 15) & ~15;          // MMA N / K extent of the last tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NSTG; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_full[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int bh = item / n_pairs, pair = item % n_pairs;
        const int q0 = pair * 2 * A2_BQ;
        const bool act1 = q0 + A2_BQ < seq;
        mbar_wait(q_empty, (it & 1) ^ 1);  // previous item's S MMAs have consumed Q
        mbar_expect_tx(&q_full[0], Cfg::Q_TILE);
#pragma unroll
        for (int s = 0; s < Cfg::NQS; ++s) tma_load_2d(sQ + s * (A2_BQ * 128), &tmQ, &q_full[0], s * 64, bh * seq_pad + q0);
        if (act1) {
          mbar_expect_tx(&q_full[1], Cfg::Q_TILE);
#pragma unroll
          for (int s = 0; s < Cfg::NQS; ++s)
            tma_load_2d(sQ + Cfg::Q_TILE + s * (A2_BQ * 128), &tmQ, &q_full[1], s * 64, bh * seq_pad + q0 + A2_BQ);
        }
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], Cfg::K_TILE);
#pragma unroll
          for (int s = 0; s < Cfg::NQS; ++s)
            tma_load_2d(sK + st * Cfg::K_TILE + s * (A2_BKV * 128), &tmK, &k_full[st], s * 64, bh * seq_pad + j * A2_BKV);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], Cfg::V_TILE);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            tma_load_2d(sV + st * Cfg::V_TILE + s * Cfg::V_SUB, &tmV, &v_full[st], j * A2_BKV + s * 64, bh * HD);
          if (++st == NSTG) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(A2_BQ, HD);
      int st = 0;
      uint32_t ph = 0;
      uint32_t p_cnt[2] = {0, 0};
      uint32_t q1_cnt = 0;  // q_full[1] / o_full[1] only complete for items whose second tile is active
      int it = 0;

      auto issue_s = [&](int g, int stage, int ncols) {
        const uint32_t idesc_s = make_idesc_bf16(A2_BQ, ncols);
        const uint32_t d = tmem_base + s_col(g);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sQ + g * Cfg::Q_TILE + (k / 4) * (A2_BQ * 128))) + 2 * (k % 4);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sK + stage * Cfg::K_TILE + (k / 4) * (A2_BKV * 128))) + 2 * (k % 4);
          umma_ss_bf16(d, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int g, int stage, int ncols, bool first) {
        const uint32_t d = tmem_base + o_col(g);
        const uint32_t a = tmem_base + s_col(g);  // P aliases the first 64 columns of S
        const int ksteps = ncols / 16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sV + stage * Cfg::V_TILE + (k / 4) * Cfg::V_SUB)) + 2 * (k % 4);
          umma_ts_bf16(d, a + k * 8, bdesc, idesc_o, (first && k == 0) ? 0u : 1u);
        }
      };

      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int pair = item % n_pairs;
        const int q0 = pair * 2 * A2_BQ;
        const bool act1 = q0 + A2_BQ < seq;
        mbar_wait(&q_full[0], it & 1);
        if (act1) {
          mbar_wait(&q_full[1], q1_cnt & 1);
          ++q1_cnt;
        }
        // prologue: S0(0), S1(0)
        {
          const int nc = (n_kv == 1) ? tail_cols : A2_BKV;
          mbar_wait(&k_full[st], ph);
          tc_fence_after();
          issue_s(0, st, nc);
          umma_commit(&s_full[0]);
          if (act1) {
            issue_s(1, st, nc);
            umma_commit(&s_full[1]);
          }
          umma_commit(&k_empty[st]);
          if (n_kv == 1) umma_commit(q_empty);
        }
        for (int j = 0; j < n_kv; ++j) {
          const int nc = (j == n_kv - 1) ? tail_cols : A2_BKV;
          const bool has_next = j + 1 < n_kv;
          const int st_next = (st + 1 == NSTG) ? 0 : st + 1;
          const uint32_t ph_next = (st + 1 == NSTG) ? (ph ^ 1) : ph;
          const int nc_next = (j + 1 == n_kv - 1) ? tail_cols : A2_BKV;
          mbar_wait(&v_full[st], ph);
          // ---- tile 0: O0 += P0 V_j ; then S0 of the next K tile ----
          mbar_wait(&p_full[0], p_cnt[0] & 1);
          ++p_cnt[0];
          tc_fence_after();
          issue_pv(0, st, nc, j == 0);
          if (!has_next) umma_commit(&o_full[0]);
          if (has_next) {
            mbar_wait(&k_full[st_next], ph_next);
            tc_fence_after();
            issue_s(0, st_next, nc_next);
            umma_commit(&s_full[0]);
          }
          // ---- tile 1 ----
          if (act1) {
            mbar_wait(&p_full[1], p_cnt[1] & 1);
            ++p_cnt[1];
            tc_fence_after();
            issue_pv(1, st, nc, j == 0);
            if (!has_next) umma_commit(&o_full[1]);
          }
          umma_commit(&v_empty[st]);
          if (has_next) {
            if (act1) {
              issue_s(1, st_next, nc_next);
              umma_commit(&s_full[1]);
            }
            umma_commit(&k_empty[st_next]);
            if (j + 1 == n_kv - 1) umma_commit(q_empty);  // last S MMAs of this item issued
          }
          st = st_next;
          ph = ph_next;
        }
      }
    }
  } else {
    // ===================== softmax groups (thread == query row of tile g) =====================
    const int g = (warp - 2) >> 2;          // 0: warps 2-5, 1: warps 6-9
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + s_col(g);
    const uint32_t t_o = tmem_base + lane_addr + o_col(g);
    const float c = args.scale_log2;
    uint32_t s_cnt = 0, o_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int bh = item / n_pairs, pair = item % n_pairs;
      const int q0 = pair * 2 * A2_BQ + g * A2_BQ;
      if (q0 >= seq) continue;  // inactive second tile: the whole group skips this item
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        const int kv0 = j * A2_BKV;
        const bool last = (j == n_kv - 1);
        const int nchunks = last ? (tail_cols + 31) / 32 : A2_BKV / 32;
        mbar_wait(&s_full[g], s_cnt & 1);
        ++s_cnt;
        tc_fence_after();
        // pass 1: row maximum over the valid keys
        float mx = -INFINITY;
        for (int ch = 0; ch < nchunks; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_s + ch * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float s = __uint_as_float(v[i]);
            if (last && kv0 + ch * 32 + i >= seq) s = -INFINITY;
            mx = fmaxf(mx, s);
          }
        }
        // lazy maximum update: only move (and rescale O) when the maximum grew by more than 2^8
        float alpha = 1.0f;
        if (j == 0) {
          m_run = mx;
        } else if ((mx - m_run) * c > 8.0f) {
          alpha = exp2f((m_run - mx) * c);
          m_run = mx;
        }
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
          l_run *= alpha;
        }
        const float mc = m_run * c;
        // pass 2: P = exp2(s*c - m*c) -> bf16 -> TMEM (aliasing S), row sum
        float psum = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_s + ch * 32, v);
          tmem_wait_ld();
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = exp2f(fmaf(__uint_as_float(v[2 * i]), c, -mc));
            float p1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
            if (last) {
              if (kv0 + ch * 32 + 2 * i >= seq) p0 = 0.f;
              if (kv0 + ch * 32 + 2 * i + 1 >= seq) p1 = 0.f;
            }
            psum += p0 + p1;
            w[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x16(t_s + ch * 16, w);
        }
        l_run += psum;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // ---- epilogue: O / l -> bf16, token-major ----
      mbar_wait(&o_full[g], o_cnt & 1);
      ++o_cnt;
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
      // O_g is free again once these loads completed; the next item's first PV_g is ordered behind this group's next
      // p_full arrival, which follows in program order.
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int g_att2_sms = 0;

template <int HD>
static cudaError_t launch_att2(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int seq, int seq_pad,
                               float scale, cudaStream_t s) {
  using Cfg = Att2Cfg<HD>;
  const int BH = L * heads;
  CUtensorMap tq, tk, tv;
  if (!get_tmap_2d_bf16(&tq, Q, static_cast<uint64_t>(BH) * seq_pad, HD, HD, 64, A2_BQ)) return cudaErrorInvalidValue;
  if (!get_tmap_2d_bf16(&tk, K, static_cast<uint64_t>(BH) * seq_pad, HD, HD, 64, A2_BKV)) return cudaErrorInvalidValue;
  if (!get_tmap_2d_bf16(&tv, Vt, static_cast<uint64_t>(BH) * HD, seq_pad, seq_pad, 64, HD)) return cudaErrorInvalidValue;
  auto kern = attn_tc2_kernel<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  if (g_att2_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_att2_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_qt = (seq + A2_BQ - 1) / A2_BQ;
  const int n_items = BH * ((n_qt + 1) / 2);
  const int grid = n_items < g_att2_sms ? n_items : g_att2_sms;
  Att2Args a{O, heads, seq, seq_pad, BH, scale * 1.4426950408889634f};
  kern<<<grid, A2_THREADS, Cfg::SMEM_BYTES, s>>>(tq, tk, tv, a);
  return cudaGetLastError();
}

cudaError_t attention_tc2(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int head_dim, int seq,
                          int seq_pad, float scale, cudaStream_t s) {
  if (seq_pad % 128 != 0 || seq > seq_pad || seq <= 0) return cudaErrorInvalidValue;
  if (head_dim == 96) return launch_att2<96>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, s);
  if (head_dim == 64) return launch_att2<64>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, s);
  return cudaErrorInvalidValue;
}

}  // namespace bd
