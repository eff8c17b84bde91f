// Attention: persistent, two query tiles per CTA in ping-pong (sm_100a, tcgen05 + TMEM + TMA).
//
// Replaces flash_attn_func / F.scaled_dot_product_attention at blocks.py:259-285 (decoder, 8 heads x 96) and the DINOv2
// attention (layers/attention.py:56-69, 12 heads x 64).  Operand layouts: Q, K [BH, seq_pad, HD]; V^T [BH, HD, seq_pad]; O
// token-major [L*seq, heads*HD] (written by the QKV GEMM epilogue / read by the proj GEMM).
//   * one persistent CTA per SM walks (bh, 256-row query pair) work items: TMEM (512 columns), barriers and tensor-map
//     prefetch are set up once, K/V tiles stream through an smem ring across items;
//   * two 128-row query tiles share every K/V tile; the MMA warp interleaves  S0, S1, PV0, S0', PV1, S1', ...  so one
//     tile's softmax (warps 2-5 / 6-9) overlaps the other tile's MMAs -- the tensor pipe only waits for the slower of
//     the MMA stream and the two softmax groups;
//   * P stays in tensor memory (TS-form PV MMA) in its own columns, so S(j+1) is issued as soon as a softmax group
//     holds S(j) in registers: in steady state the groups never wait for the tensor pipe (exp/MUFU-paced);
//   * the running maximum is only advanced (and O rescaled in TMEM) when it grows by more than 2^8 -- stale maxima are
//     exact because the same maximum scales P and the row sum;
//   * the last K/V tile is trimmed to the next multiple of 16 keys (runtime UMMA N / K), which removes most of the
//     padding waste of short sequences (DINOv2: 261 keys = 2 tiles + 16 keys instead of 3 tiles).
#include <stdlib.h>

#include "bd_internal.h"
#include "common.cuh"

namespace bd {

bool get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br);

static constexpr int A2_THREADS = 384;  // 3 warpgroups: softmax tile 0, softmax tile 1, {TMA, MMA, -, -}
// The control warps sit in the LAST warpgroup: the SMSP arbiter prefers the highest warp id, and the MMA / TMA issuers
// (few instructions, all on the critical path) must never queue behind the softmax warps sharing their scheduler.
static constexpr int A2_CTRL_WARP = 8;
#ifndef A2_TURNS_MIN_KV
#define A2_TURNS_MIN_KV 0   // sequences with at most this many key tiles run the two softmax groups without turns
#endif
#ifndef A2_PASS_NUM
#define A2_PASS_NUM 3
#define A2_PASS_DEN 4
#endif
static constexpr int A2_BQ = 128;
// One quad of scores in A2_POLY_PERIOD goes through ex2_poly_pair (FMA pipe) instead of the MUFU pipe; 0 = all on MUFU.
// Measured on B200 (decoder shape, isolated / in-step ms per launch): 0 -> 0.4426 / 0.545; 4 -> 0.4386 / 0.510; 3 -> 0.4491; 2 -> 0.4762
// (profiles/r02_attention_variants.txt).  A2_PACK2: scale the scores with packed FFMA2 (needed for the polynomial to pay: the
// issue port, not only the MUFU pipe, limits the exp loop).
#ifndef A2_POLY_PERIOD
#define A2_POLY_PERIOD 4
#define A2_PACK2 1
#endif
#ifndef A2_POLY_PHASE
#define A2_POLY_PHASE (A2_POLY_PERIOD - 1)
#endif

// 2^x for a pair of scores on the FMA pipe (packed fp32): x = s*c + nmc, clamped at -126, split into n = round(x) and
// f = x - n in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, far below the bf16 rounding of P);
// n goes straight into the exponent field.
__device__ __forceinline__ void ex2_poly_pair(float s0, float s1, f32x2 cc, f32x2 nn, float& r0, float& r1) {
  float x0, x1;
  unpack_f32x2(fma_f32x2(pack_f32x2(s0, s1), cc, nn), x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const f32x2 x = pack_f32x2(x0, x1);
  const f32x2 t = add_f32x2(x, pack_f32x2(12582912.0f, 12582912.0f));            // 1.5 * 2^23: the low mantissa bits hold round(x)
  const f32x2 n = add_f32x2(t, pack_f32x2(-12582912.0f, -12582912.0f));
  const f32x2 f = fma_f32x2(n, pack_f32x2(-1.0f, -1.0f), x);
  f32x2 p = fma_f32x2(f, pack_f32x2(0.0551716685f, 0.0551716685f), pack_f32x2(0.2426111251f, 0.2426111251f));
  p = fma_f32x2(p, f, pack_f32x2(0.6932609677f, 0.6932609677f));
  p = fma_f32x2(p, f, pack_f32x2(0.9999280572f, 0.9999280572f));
  float pa, pb, ta, tb;
  unpack_f32x2(p, pa, pb);
  unpack_f32x2(t, ta, tb);
  r0 = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ta) << 23));
  r1 = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(tb) << 23));
}

template <int HD>
struct Att2Cfg {
  static constexpr int BKV = (HD > 64) ? 96 : 128;   // keys per tile: S (BKV) + P (BKV/2) + O (HD) fp32 columns per query tile <= 256
  // Operand tiles are K-major.  Columns 0..63 live in SWIZZLE_128B boxes (128-byte rows); head dim 96 adds a 32-column
  // SWIZZLE_64B box (64-byte rows) instead of a half-empty second 128-byte box, which is what makes room for two Q buffers.
  static constexpr bool SPLIT = HD > 64;
  static constexpr int Q_P0 = A2_BQ * 128, Q_P1 = SPLIT ? A2_BQ * 64 : 0, Q_TILE = Q_P0 + Q_P1;
  static constexpr int K_P0 = BKV * 128, K_P1 = SPLIT ? BKV * 64 : 0, K_TILE = K_P0 + K_P1;
  static constexpr bool VSPLIT = (BKV % 64) != 0;    // keys 64.. of a tile: a 32-key SWIZZLE_64B box (BKV 96) or a second 64-key box
  static constexpr int V_P0 = HD * 128, V_P1 = VSPLIT ? HD * 64 : HD * 128, V_TILE = V_P0 + V_P1;
  static constexpr int NSTG = (HD > 64) ? 3 : 4;
  static constexpr int QBUF = 2;                      // the next item's Q is prefetched; the current one doubles as the O staging tile
  static constexpr int BAR_BYTES = 256 + 2048;       // mbarriers + 2 x 256 hand-over slots (a2_turn_*)
  static constexpr int SMEM_BYTES = QBUF * 2 * Q_TILE + NSTG * (K_TILE + V_TILE) + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int S_OFF = 0, P_OFF = BKV, O_OFF = BKV + BKV / 2;   // column offsets inside a tile's 256-column half
  static_assert(O_OFF + HD <= 256, "tile does not fit its TMEM half");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct Att2Args {
  int heads, seq, seq_pad, BH;
  int q_off;         // first query row handled by this kernel (rows [0, q_off) are done by attention_prefix_rows)
  // Query window (the decoder's LAST layer: only the query view's tokens feed the head, betr.py:419-430): when win_idx != nullptr
  // sequence l attends only from its rows [win_idx[l] * win_rows, +win_rows) -- over ALL keys -- and O is the compact
  // [L, win_rows, heads*HD] tensor.  q_off must be 0 then.
  const long long* win_idx;
  int win_rows;
  float scale_log2;
  long long* trace;  // debug: per-role clock64 stamps of CTA 0's first items (nullptr in production)
};
struct Att2Maps {    // Q, K: [BH*seq_pad, HD]; V^T: [BH*HD, seq_pad]; O: [L][seq][heads*HD] (3-D: rows are clipped at seq)
  CUtensorMap q, q2, k, k2, v, v2, o, o2;   // *2: the 32-column (SWIZZLE_64B) boxes of head dim 96 / of 96-key tiles
};
#define A2_TRACE(role, slot) do { if (args.trace != nullptr && blockIdx.x == 0 && (slot) < 512) args.trace[(role) * 512 + (slot)] = clock64(); } while (0)

// ptxas moves register-only work (the MUFU stream) freely across BAR instructions, so the hand-over is tied into the
// data flow through shared memory: the exp loop's input offset is re-read (volatile) after the bar.sync, and the row sum
// it produces is stored (volatile) before the bar.arrive.  One 4-byte slot per thread and direction.
__device__ __forceinline__ void a2_turn_wait(int g, float& dep, uint32_t slot) {
  asm volatile("st.volatile.shared.f32 [%2], %0;\n\tbar.sync %1, 256;\n\tld.volatile.shared.f32 %0, [%2];"
               : "+f"(dep) : "r"(2 + g), "r"(slot) : "memory");
}
__device__ __forceinline__ void a2_turn_pass(int g, float dep, uint32_t slot) {
  asm volatile("st.volatile.shared.f32 [%2], %0;\n\tbar.arrive %1, 256;" ::"f"(dep), "r"(2 + (g ^ 1)), "r"(slot) : "memory");
}
__device__ __forceinline__ void a2_group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(4 + g) : "memory"); }

// S_g = Q_g K^T : HD/16 UMMAs (128 x ncols x 16), operands in shared memory (low descriptor words)
template <int HD>
__device__ __forceinline__ void a2_issue_s(uint32_t d_tmem, uint32_t q_lo, uint32_t k_lo, uint32_t idesc_s) {
  using Cfg = Att2Cfg<HD>;
#pragma unroll
  for (int k = 0; k < HD / 16; ++k) {
    if (k < 4) umma_ss_lh_w(d_tmem, q_lo + 2 * k, k_lo + 2 * k, SMEM_DESC_HI_SW128, idesc_s, k != 0 ? 1u : 0u);
    else umma_ss_lh_w(d_tmem, q_lo + (Cfg::Q_P0 >> 4) + 2 * (k - 4), k_lo + (Cfg::K_P0 >> 4) + 2 * (k - 4), SMEM_DESC_HI_SW64, idesc_s, 1u);
  }
}
// O_g (+)= P_g V : ncols/16 UMMAs (128 x HD x 16), P from tensor memory, V^T from shared memory
template <int HD, bool CHECK>
__device__ __forceinline__ void a2_issue_pv_k(uint32_t d_tmem, uint32_t p_tmem, uint32_t v_lo, int ncols, uint32_t acc0) {
  using Cfg = Att2Cfg<HD>;
  constexpr uint32_t idesc_o = make_idesc_bf16(A2_BQ, HD);
#pragma unroll
  for (int k = 0; k < Cfg::BKV / 16; ++k) {
    if (CHECK && k * 16 >= ncols) break;
    if (k < 4) umma_ts_lh_w(d_tmem, p_tmem + k * 8, v_lo + 2 * k, SMEM_DESC_HI_SW128, idesc_o, k == 0 ? acc0 : 1u);
    else umma_ts_lh_w(d_tmem, p_tmem + k * 8, v_lo + (Cfg::V_P0 >> 4) + 2 * (k - 4), Cfg::VSPLIT ? SMEM_DESC_HI_SW64 : SMEM_DESC_HI_SW128, idesc_o, 1u);
  }
}
template <int HD>
__device__ __forceinline__ void a2_issue_pv(uint32_t d_tmem, uint32_t p_tmem, uint32_t v_lo, int ncols, uint32_t acc0) {
  if (ncols == Att2Cfg<HD>::BKV) a2_issue_pv_k<HD, false>(d_tmem, p_tmem, v_lo, ncols, acc0);  // full tile: back-to-back UMMAs
  else a2_issue_pv_k<HD, true>(d_tmem, p_tmem, v_lo, ncols, acc0);
}

template <int HD>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ Att2Maps tm, const Att2Args args) {
  using Cfg = Att2Cfg<HD>;
  constexpr int NSTG = Cfg::NSTG;
  constexpr int QBUF = Cfg::QBUF;
  constexpr int BKV = Cfg::BKV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [QBUF][2][Q_TILE]
  uint8_t* sK = sQ + QBUF * 2 * Cfg::Q_TILE;           // [NSTG][K_TILE]
  uint8_t* sV = sK + NSTG * Cfg::K_TILE;               // [NSTG][V_TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NSTG * Cfg::V_TILE);
  uint64_t* q_full = bars;                  // [QBUF][2]
  uint64_t* q_empty = bars + 4;             // [QBUF]  both softmax groups have stored their O tile out of this Q buffer
  uint64_t* k_full = bars + 6;              // [NSTG]
  uint64_t* k_empty = k_full + NSTG;
  uint64_t* v_full = k_empty + NSTG;
  uint64_t* v_empty = v_full + NSTG;
  uint64_t* s_full = v_empty + NSTG;        // [2]  MMA -> softmax: S_g(j) complete
  uint64_t* s_free = s_full + 2;            // [2]  softmax -> MMA: S_g(j) is in registers, S_g may be overwritten
  uint64_t* p_full = s_free + 2;            // [2]  softmax -> MMA: P_g(j) stored (and O_g rescaled)
  uint64_t* pv_done = p_full + 2;           // [2]  MMA -> softmax: O_g += P_g(j) V_j complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 2);
  const uint32_t turn_slot = smem_u32(reinterpret_cast<uint8_t*>(bars) + 256) + threadIdx.x * 4;  // softmax threads: 0..255

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches stay convergent, operands stay in uniform registers
  const int lane = threadIdx.x & 31;
  const int seq = args.seq, seq_pad = args.seq_pad;
  const int q_off = args.q_off;
  const bool windowed = args.win_idx != nullptr;
  const int q_end = windowed ? args.win_rows : seq;      // query rows (relative to the window start) end here
  const int n_qt = (q_end - q_off + A2_BQ - 1) / A2_BQ;  // query tiles per sequence
  const int n_pairs = (n_qt + 1) / 2;
  const int n_items = args.BH * n_pairs;
  const int n_kv = (seq + BKV - 1) / BKV;
  const int tail_keys = seq - (n_kv - 1) * BKV;          // 1..BKV valid keys in the last tile
  const int tail_cols = (tail_keys + 15) & ~15;          // MMA N / K extent of the last tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.q);
    tma_prefetch_desc(&tm.k);
    tma_prefetch_desc(&tm.v);
    tma_prefetch_desc(&tm.o);
    if (Cfg::SPLIT) {
      tma_prefetch_desc(&tm.q2);
      tma_prefetch_desc(&tm.k2);
      tma_prefetch_desc(&tm.o2);
    }
    tma_prefetch_desc(&tm.v2);
    for (int i = 0; i < 4; ++i) mbar_init(&q_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&q_empty[i], 2);
    for (int i = 0; i < NSTG; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 4);
      mbar_init(&p_full[g], 4);
      mbar_init(&pv_done[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == A2_CTRL_WARP + 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast through a shuffle so the compiler keeps the TMEM base (and everything derived from it) in uniform registers
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  // register budget: the two softmax warpgroups keep a whole S row per thread
  if (warp >= A2_CTRL_WARP) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == A2_CTRL_WARP) {
    // ===================== TMA producer (whole warp, elected lane issues) =====================
    int st = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int item_f = blockIdx.x; item_f < n_items; item_f += gridDim.x, ++it) {
      const int item = item_f;
      const int bh = item / n_pairs, pair = item % n_pairs;
      const int qrel = q_off + pair * 2 * A2_BQ;
      const bool act1 = qrel + A2_BQ < q_end;
      const int q0 = qrel + (windowed ? static_cast<int>(args.win_idx[bh / args.heads]) * args.win_rows : 0);
      const int qb = it % QBUF;
      uint8_t* sQi = sQ + qb * 2 * Cfg::Q_TILE;
      if (lane == 0) A2_TRACE(3, it * 8 + 0);
      mbar_wait(&q_empty[qb], ((it / QBUF) & 1) ^ 1);  // the O tiles of the item that used this buffer have been stored
      if (lane == 0) A2_TRACE(3, it * 8 + 1);
      mbar_expect_tx_w(&q_full[qb * 2 + 0], Cfg::Q_TILE);
      tma_load_2d_w(sQi, &tm.q, &q_full[qb * 2 + 0], 0, bh * seq_pad + q0);
      if (Cfg::SPLIT) tma_load_2d_w(sQi + Cfg::Q_P0, &tm.q2, &q_full[qb * 2 + 0], 64, bh * seq_pad + q0);
      if (act1) {
        mbar_expect_tx_w(&q_full[qb * 2 + 1], Cfg::Q_TILE);
        tma_load_2d_w(sQi + Cfg::Q_TILE, &tm.q, &q_full[qb * 2 + 1], 0, bh * seq_pad + q0 + A2_BQ);
        if (Cfg::SPLIT) tma_load_2d_w(sQi + Cfg::Q_TILE + Cfg::Q_P0, &tm.q2, &q_full[qb * 2 + 1], 64, bh * seq_pad + q0 + A2_BQ);
      }
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx_w(&k_full[st], Cfg::K_TILE);
        tma_load_2d_w(sK + st * Cfg::K_TILE, &tm.k, &k_full[st], 0, bh * seq_pad + j * BKV);
        if (Cfg::SPLIT) tma_load_2d_w(sK + st * Cfg::K_TILE + Cfg::K_P0, &tm.k2, &k_full[st], 64, bh * seq_pad + j * BKV);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx_w(&v_full[st], Cfg::V_TILE);
        tma_load_2d_w(sV + st * Cfg::V_TILE, &tm.v, &v_full[st], j * BKV, bh * HD);
        tma_load_2d_w(sV + st * Cfg::V_TILE + Cfg::V_P0, &tm.v2, &v_full[st], j * BKV + 64, bh * HD);
        if (++st == NSTG) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == A2_CTRL_WARP + 1) {
    // ===================== MMA issuer (whole warp, elected lane issues) =====================
    int st = 0;
    uint32_t ph = 0;
    uint32_t pp0 = 0, pp1 = 0, fp0 = 0, fp1 = 0;   // parities of p_full[g] / s_free[g]
    uint32_t q1p = 0;                              // bit qb: parity of q_full[qb][1] (only completes for items with an active second tile)
    // every descriptor is {low word, constant high word}: the loop carries 32-bit low words derived from warp-uniform values
    const uint32_t q_lo0 = make_smem_desc_lo(smem_u32(sQ)), k_lo0 = make_smem_desc_lo(smem_u32(sK)), v_lo0 = make_smem_desc_lo(smem_u32(sV));
    constexpr uint32_t idesc_full = make_idesc_bf16(A2_BQ, BKV);
    const uint32_t idesc_tail = make_idesc_bf16(A2_BQ, tail_cols);
    const uint32_t tS0 = tmem_base + Cfg::S_OFF, tS1 = tmem_base + 256 + Cfg::S_OFF;
    const uint32_t tP0 = tmem_base + Cfg::P_OFF, tP1 = tmem_base + 256 + Cfg::P_OFF;
    const uint32_t tO0 = tmem_base + Cfg::O_OFF, tO1 = tmem_base + 256 + Cfg::O_OFF;
    int it = 0;

    // Issue order per item:  S0(0) S1(0) | for j: [S0(j+1) S1(j+1) as soon as the softmax groups hold S(j) in registers]
    // PV0(j) PV1(j).  S(j+1) is therefore already complete when a group finishes tile j: the groups never wait for the
    // tensor pipe in steady state and the kernel runs at the pace of the exp (MUFU) pipe.
    for (int item_f = blockIdx.x; item_f < n_items; item_f += gridDim.x, ++it) {
      const int item = item_f;
      const int pair = item % n_pairs;
      const bool act1 = q_off + pair * 2 * A2_BQ + A2_BQ < q_end;
      const int qb = it % QBUF;
      const uint32_t q_lo = q_lo0 + qb * ((2 * Cfg::Q_TILE) >> 4);
      const uint32_t q_lo1 = q_lo + (Cfg::Q_TILE >> 4);
      if (lane == 0) A2_TRACE(3, it * 8 + 2);
      mbar_wait(&q_full[qb * 2 + 0], (it / QBUF) & 1);
      if (act1) {
        mbar_wait(&q_full[qb * 2 + 1], (q1p >> qb) & 1);
        q1p ^= 1u << qb;
      }
      if (lane == 0) A2_TRACE(3, it * 8 + 3);
      mbar_wait(&k_full[st], ph);
      if (lane == 0) A2_TRACE(3, it * 8 + 4);
      tc_fence_after();
      {
        const uint32_t id0 = (n_kv == 1) ? idesc_tail : idesc_full;
        const uint32_t k_lo = k_lo0 + st * (Cfg::K_TILE >> 4);
        a2_issue_s<HD>(tS0, q_lo, k_lo, id0);
        umma_commit_w(&s_full[0]);
        if (act1) {
          a2_issue_s<HD>(tS1, q_lo1, k_lo, id0);
          umma_commit_w(&s_full[1]);
        }
        umma_commit_w(&k_empty[st]);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int nc = (j == n_kv - 1) ? tail_cols : BKV;
        const int st_next = (st + 1 == NSTG) ? 0 : st + 1;
        const uint32_t ph_next = (st + 1 == NSTG) ? (ph ^ 1) : ph;
        if (j + 1 < n_kv) {
          const uint32_t idn = (j + 1 == n_kv - 1) ? idesc_tail : idesc_full;
          const uint32_t k_lo = k_lo0 + st_next * (Cfg::K_TILE >> 4);
          mbar_wait(&k_full[st_next], ph_next);
          mbar_wait(&s_free[0], fp0);
          fp0 ^= 1;
          tc_fence_after();
          a2_issue_s<HD>(tS0, q_lo, k_lo, idn);
          umma_commit_w(&s_full[0]);
          if (act1) {
            mbar_wait(&s_free[1], fp1);
            fp1 ^= 1;
            tc_fence_after();
            a2_issue_s<HD>(tS1, q_lo1, k_lo, idn);
            umma_commit_w(&s_full[1]);
          }
          umma_commit_w(&k_empty[st_next]);
        }
        const uint32_t v_lo = v_lo0 + st * (Cfg::V_TILE >> 4);
        const uint32_t acc0 = j == 0 ? 0u : 1u;
        mbar_wait(&v_full[st], ph);
        if (lane == 0) A2_TRACE(0, (it * n_kv + j) * 4 + 0);
        mbar_wait(&p_full[0], pp0);
        if (lane == 0) A2_TRACE(0, (it * n_kv + j) * 4 + 1);
        pp0 ^= 1;
        tc_fence_after();
        a2_issue_pv<HD>(tO0, tP0, v_lo, nc, acc0);
        umma_commit_w(&pv_done[0]);
        if (act1) {
          if (lane == 0) A2_TRACE(0, (it * n_kv + j) * 4 + 2);
          mbar_wait(&p_full[1], pp1);
          if (lane == 0) A2_TRACE(0, (it * n_kv + j) * 4 + 3);
          pp1 ^= 1;
          tc_fence_after();
          a2_issue_pv<HD>(tO1, tP1, v_lo, nc, acc0);
          umma_commit_w(&pv_done[1]);
        }
        umma_commit_w(&v_empty[st]);
        st = st_next;
        ph = ph_next;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ===================== softmax groups (thread == query row of tile g) =====================
    const int g = warp >> 2;                // 0: warps 0-3, 1: warps 4-7
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + g * 256 + Cfg::S_OFF;
    const uint32_t t_p = tmem_base + lane_addr + g * 256 + Cfg::P_OFF;
    const uint32_t t_o = tmem_base + lane_addr + g * 256 + Cfg::O_OFF;
    const float c = args.scale_log2;
    const bool storer = (quad == 0) && (lane == 0);   // issues this group's O stores and releases the Q buffer afterwards
    uint32_t s_cnt = 0, d_cnt = 0;
    int pend_qb = -1, pend_arrivals = 0;    // O store in flight out of Q buffer pend_qb (storer thread only)
    int it = 0;
    for (int item_f = blockIdx.x; item_f < n_items; item_f += gridDim.x, ++it) {
      const int item = item_f;
      const int bh = item / n_pairs, pair = item % n_pairs;
      const int q0 = q_off + pair * 2 * A2_BQ + g * A2_BQ;   // relative to the window start (= absolute without a window): the O row
      if (q0 >= q_end) {  // inactive second tile: the whole group skips this item (but still releases its last staging tile:
                        // the producer may need that Q buffer before this group becomes active again)
        if (storer && pend_qb >= 0) {
          bulk_wait_group_read<0>();
          for (int a = 0; a < pend_arrivals; ++a) mbar_arrive(&q_empty[pend_qb]);
          pend_qb = -1;
        }
        continue;
      }
      const int qb = it % QBUF;
      // Anti-phase: while one group runs its exp loop (MUFU-bound) the other one reads S from TMEM, takes the row
      // maximum, stores P and hands it to the MMA warp.  Strict alternation g0, g1, g0, ... per key tile; group 1 opens
      // group 0's first turn of every item.  Items whose second tile is inactive run group 0 alone, without turns.
      const bool turns = (q_off + pair * 2 * A2_BQ + A2_BQ < q_end) && (n_kv > A2_TURNS_MIN_KV);
      float m_run = -INFINITY, l_run = 0.f;
      if (turns && g == 1) a2_turn_pass(1, l_run, turn_slot + 1024);
      for (int j = 0; j < n_kv; ++j) {
        const int kv0 = j * BKV;
        const bool last = (j == n_kv - 1);
        if (quad == 0 && lane == 0) A2_TRACE(1 + g, (it * n_kv + j) * 8 + 0);
        mbar_wait(&s_full[g], s_cnt & 1);
        if (quad == 0 && lane == 0) A2_TRACE(1 + g, (it * n_kv + j) * 8 + 1);
        ++s_cnt;
        tc_fence_after();
        float alpha = 1.0f;
        if (!last || tail_keys == BKV) {
          // ---------- full tile: the whole S row lives in registers (one TMEM read, no masking) ----------
          uint32_t sv[BKV];
#pragma unroll
          for (int ch = 0; ch < BKV / 32; ++ch) tmem_ld_32x32b_x32p(t_s + ch * 32, sv + ch * 32);
          tmem_wait_ld();
          if (!last) {  // S_g may be overwritten by S_g(j+1) from now on
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
          }
          if (quad == 0 && lane == 0) A2_TRACE(1 + g, (it * n_kv + j) * 8 + 2);
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < BKV; i += 8) {
            mx0 = fmax3(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
            mx2 = fmax3(mx2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
            mx3 = fmax3(mx3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          if (j == 0) {
            m_run = mx;
          } else if ((mx - m_run) * c > 8.0f) {  // lazy maximum: move only when it grew by more than 2^8
            alpha = ex2_approx((m_run - mx) * c);
            m_run = mx;
          }
          l_run *= alpha;
          float nmc = -m_run * c;
          float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
          if (turns) a2_turn_wait(g, nmc, turn_slot);
#ifdef A2_TRACE_FINE   // two more stamps inside the exp phase (scripts/build_variant.sh fine attn_tc2.cu -DA2_TRACE_FINE): they cost ~2 %
          if (quad == 0 && lane == 0) A2_TRACE(1 + g, (it * n_kv + j) * 8 + 4);
#endif
          // The turn is handed over when A2_PASS_NUM/A2_PASS_DEN of the exponentials are done: the other group's first
          // exponentials overlap this group's last ones, which hides the hand-over latency without starving the MUFU pipe.
          constexpr int PASS_I = ((BKV / 2) * A2_PASS_NUM / A2_PASS_DEN) & ~1;
#if A2_POLY_PERIOD > 0 || defined(A2_PACK2)
          const f32x2 cc2 = pack_f32x2(c, c), nn2 = pack_f32x2(nmc, nmc);
#endif
#pragma unroll
          for (int i = 0; i < BKV / 2; i += 2) {
            if (i == PASS_I && turns && !(g == 1 && last))   // group 1's last pass of an item is replaced by the next item's opening pass
              a2_turn_pass(g, (ps0 + ps1) + (ps2 + ps3), turn_slot + 1024);
            float p0, p1, p2, p3;
#if A2_POLY_PERIOD > 0
            // every A2_POLY_PERIOD-th quad of scores takes the FMA-pipe polynomial instead of the MUFU pipe (the MUFU pipe is
            // the kernel's bound: 16 ex2 per clock and SM against 2 x 128 x BKV exponentials per key-tile pair)
            if (((i >> 1) % A2_POLY_PERIOD) == A2_POLY_PHASE) {
              ex2_poly_pair(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1]), cc2, nn2, p0, p1);
              ex2_poly_pair(__uint_as_float(sv[2 * i + 2]), __uint_as_float(sv[2 * i + 3]), cc2, nn2, p2, p3);
            } else
#endif
            {
#ifdef A2_PACK2
              float x0, x1, x2, x3;
              unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), cc2, nn2), x0, x1);
              unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * i + 2]), __uint_as_float(sv[2 * i + 3])), cc2, nn2), x2, x3);
              p0 = ex2_approx(x0); p1 = ex2_approx(x1); p2 = ex2_approx(x2); p3 = ex2_approx(x3);
#else
              p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * i]), c, nmc));
              p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 1]), c, nmc));
              p2 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 2]), c, nmc));
              p3 = ex2_approx(fmaf(__uint_as_float(sv[2 * i + 3]), c, nmc));
#endif
            }
            ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
            sv[i] = pack_bf16x2(p0, p1);       // in place: word i is written after elements 2i, 2i+1 were consumed
            sv[i + 1] = pack_bf16x2(p2, p3);
          }
          l_run += (ps0 + ps1) + (ps2 + ps3);
#ifdef A2_TRACE_FINE
          if (quad == 0 && lane == 0 && args.trace != nullptr && blockIdx.x == 0 && (it * n_kv + j) * 8 + 5 < 512)
            args.trace[(1 + g) * 512 + (it * n_kv + j) * 8 + 5] = clock64() + (__float_as_uint(l_run) & 0);   // after the exp loop (data-dependent)
#endif
          if (j > 0) {  // PV_g(j-1) must have finished reading P_g and writing O_g
            mbar_wait(&pv_done[g], d_cnt & 1);
            ++d_cnt;
            tc_fence_after();
          }
          tmem_st_32x32b_x32p(t_p + 0, sv + 0);
          if constexpr (BKV == 128) tmem_st_32x32b_x32p(t_p + 32, sv + 32);
          else tmem_st_32x32b_x16(t_p + 32, sv + 32);
        } else {
          // ---------- trimmed last tile: 16-column pieces, masked, two passes over TMEM ----------
          const int n16 = tail_cols >> 4;
          float mx = -INFINITY;
          for (int c16 = 0; c16 < n16; ++c16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_s + c16 * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (kv0 + c16 * 16 + i < seq) ? __uint_as_float(v[i]) : -INFINITY);
          }
          if (j == 0) {
            m_run = mx;
          } else if ((mx - m_run) * c > 8.0f) {
            alpha = ex2_approx((m_run - mx) * c);
            m_run = mx;
          }
          l_run *= alpha;
          if (j > 0) {
            mbar_wait(&pv_done[g], d_cnt & 1);
            ++d_cnt;
            tc_fence_after();
          }
          float nmc = -m_run * c;
          float psum = 0.f;
          if (turns) a2_turn_wait(g, nmc, turn_slot);
          for (int c16 = 0; c16 < n16; ++c16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_s + c16 * 16, v);
            tmem_wait_ld();
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, nmc));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, nmc));
              if (kv0 + c16 * 16 + 2 * i >= seq) p0 = 0.f;
              if (kv0 + c16 * 16 + 2 * i + 1 >= seq) p1 = 0.f;
              psum += p0 + p1;
              w[i] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x32b_x8(t_p + c16 * 8, w);
          }
          l_run += psum;
          if (turns && g == 0) a2_turn_pass(g, l_run, turn_slot + 1024);  // the tail tile is always the last one (see the full-tile branch)
        }
        // O rescale (rare: lazy maximum), after PV_g(j-1) completed and with the S registers dead
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
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        if (quad == 0 && lane == 0) A2_TRACE(1 + g, (it * n_kv + j) * 8 + 3);
        // the previous item's O store has long finished reading its staging tile: release that Q buffer to the producer
        if (storer && pend_qb >= 0) {
          bulk_wait_group_read<0>();
          for (int a = 0; a < pend_arrivals; ++a) mbar_arrive(&q_empty[pend_qb]);
          pend_qb = -1;
        }
      }
      // ---- epilogue: O / l -> bf16 into this tile's (dead) Q buffer in the TMA box layout, one bulk tensor store per box ----
      mbar_wait(&pv_done[g], d_cnt & 1);    // also: every S MMA that read the Q tile has completed
      ++d_cnt;
      tc_fence_after();
      if (g == 0 && quad == 0 && lane == 0) A2_TRACE(3, it * 8 + 5);
      const float inv_l = 1.0f / l_run;
      uint8_t* stage = sQ + (qb * 2 + g) * Cfg::Q_TILE;
      const uint32_t row0 = smem_u32(stage) + r * 128, row1 = smem_u32(stage) + Cfg::Q_P0 + r * 64;
#pragma unroll
      for (int ch = 0; ch < HD / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_o + ch * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t ox = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
          const uint32_t oy = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
          const uint32_t oz = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
          const uint32_t ow = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
          const int cch = ch * 4 + q;   // 16-byte chunk (8 columns) of the row
          const uint32_t addr = (cch < 8) ? row0 + ((cch ^ (r & 7)) << 4) : row1 + (((cch - 8) ^ ((r >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ox), "r"(oy), "r"(oz), "r"(ow) : "memory");
        }
      }
      // O_g is free again once these loads completed; the next item's first PV_g is ordered behind this group's next
      // p_full arrival, which follows in program order.
      tc_fence_before();
      fence_proxy_async_smem();
      a2_group_sync(g);
      if (storer) {
        const int l_idx = bh / args.heads, head = bh % args.heads;
        tma_store_3d(&tm.o, stage, head * HD, q0, l_idx);
        if (Cfg::SPLIT) tma_store_3d(&tm.o2, stage + Cfg::Q_P0, head * HD + 64, q0, l_idx);
        bulk_commit_group();
        pend_qb = qb;
        pend_arrivals = (g == 0 && !turns) ? 2 : 1;   // group 0 also arrives for an inactive group 1
      }
    }
    if (storer && pend_qb >= 0) bulk_wait_group<0>();   // the last store must be complete before the CTA's shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A2_CTRL_WARP + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static long long* g_att2_trace = nullptr;
void attention_tc2_set_trace(long long* dev_buf) { g_att2_trace = dev_buf; }

bool get_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0,
                      uint32_t b1);

template <int HD>
static cudaError_t launch_att2(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int seq, int seq_pad,
                               float scale, int q_off, cudaStream_t s, const long long* win_idx = nullptr, int win_rows = 0) {
  using Cfg = Att2Cfg<HD>;
  const int BH = L * heads;
  const uint64_t qk_rows = static_cast<uint64_t>(BH) * seq_pad, v_rows = static_cast<uint64_t>(BH) * HD;
  const uint64_t o_cols = static_cast<uint64_t>(heads) * HD;
  const uint64_t o_rows = win_idx ? win_rows : seq;   // rows per sequence of the output tensor (compact with a query window)
  Att2Maps tm;
  bool ok = get_tmap_2d_bf16(&tm.q, Q, qk_rows, HD, HD, 64, A2_BQ) && get_tmap_2d_bf16(&tm.k, K, qk_rows, HD, HD, 64, Cfg::BKV) &&
            get_tmap_2d_bf16(&tm.v, Vt, v_rows, seq_pad, seq_pad, 64, HD) &&
            get_tmap_3d_bf16(&tm.o, O, o_cols, o_rows, L, o_cols * 2, static_cast<uint64_t>(o_rows) * o_cols * 2, 64, A2_BQ);
  if (ok && Cfg::SPLIT) {
    ok = get_tmap_2d_bf16(&tm.q2, Q, qk_rows, HD, HD, 32, A2_BQ) && get_tmap_2d_bf16(&tm.k2, K, qk_rows, HD, HD, 32, Cfg::BKV) &&
         get_tmap_3d_bf16(&tm.o2, O, o_cols, o_rows, L, o_cols * 2, static_cast<uint64_t>(o_rows) * o_cols * 2, 32, A2_BQ);
  } else if (ok) {
    tm.q2 = tm.q; tm.k2 = tm.k; tm.o2 = tm.o;
  }
  if (ok && Cfg::VSPLIT) ok = get_tmap_2d_bf16(&tm.v2, Vt, v_rows, seq_pad, seq_pad, 32, HD);
  else if (ok) tm.v2 = tm.v;
  if (!ok) return cudaErrorInvalidValue;
  auto kern = attn_tc2_kernel<HD>;
  cudaError_t aerr = tc_ensure_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES);
  if (aerr != cudaSuccess) return aerr;
  const int n_sms = tc_num_sms();
  const int n_qt = ((win_idx ? win_rows : seq) - q_off + A2_BQ - 1) / A2_BQ;
  const int n_items = BH * ((n_qt + 1) / 2);
  const int grid = n_items < n_sms ? n_items : n_sms;
  Att2Args a{heads, seq, seq_pad, BH, q_off, win_idx, win_rows, scale * 1.4426950408889634f, g_att2_trace};
  kern<<<grid, A2_THREADS, Cfg::SMEM_BYTES, s>>>(tm, a);
  return cudaGetLastError();
}

// Rows [0, nrows) (nrows <= 8) of every sequence: one WARP per (image, head), flash-style over 96-key chunks with
// warp-level mma.sync.m16n8k16 (the rows fill the upper half of the 16-row A fragment, the lower half is zero).  Used for
// the few tokens (DINOv2: cls + 4 registers) that would otherwise cost a whole extra 128-row query tile.  K rows and V^T
// rows are read straight from global memory into B fragments with 16- / 8-byte loads; this works because both
// contractions are order-free, so the k slots of a fragment are mapped to whatever elements sit contiguously in memory:
//   S = Q K^T : lane q (= lane % 4) covers dims [q*HD/4, (q+1)*HD/4): k-step s <- dims q*HD/4 + 4s .. +3
//   O = P V   : score tile t (8 keys), column n holds key  16*(t/2) + 4*(n/2) + (n%2) + 2*(t%2),  so the four P values a
//               lane feeds into PV k-step t/2 belong to the four consecutive keys 16*(t/2) + 4q .. +3 of V^T.
// HBM/L2-bound: K and V^T of the (image, head) are streamed exactly once.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

static constexpr int PFX_WARPS = 4;       // warps (= (image, head) pairs) per CTA
static constexpr int PFX_CHUNK_TILES = 12;  // 8-key score tiles per chunk (96 keys)

template <int HD>
__global__ void __launch_bounds__(PFX_WARPS * 32) attention_prefix_rows_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K,
                                                                            const bf16* __restrict__ Vt, bf16* __restrict__ O, int BH,
                                                                            int heads, int seq, int seq_pad, int nrows,
                                                                            float scale_log2) {
  constexpr int KS = HD / 16;          // k-steps of S = Q K^T
  constexpr int DPL = HD / 4;          // dims per lane quarter
  constexpr int NT = PFX_CHUNK_TILES;
  constexpr int OT = HD / 8;           // 8-dim output tiles
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.x * PFX_WARPS + (threadIdx.x >> 5);
  if (bh >= BH) return;
  const int g = lane >> 2, q = lane & 3;
  const bf16* Qb = Q + static_cast<long long>(bh) * seq_pad * HD;
  const bf16* Kb = K + static_cast<long long>(bh) * seq_pad * HD;
  const bf16* Vb = Vt + static_cast<long long>(bh) * HD * seq_pad;
  // A fragments of Q (row g; rows 8..15 of the fragment are zero), pre-multiplied layout: a0 = dims +0,+1; a2 = dims +2,+3
  uint32_t qa0[KS], qa2[KS];
#pragma unroll
  for (int s2 = 0; s2 < KS; ++s2) {
    uint2 v = make_uint2(0u, 0u);
    if (g < nrows) v = *reinterpret_cast<const uint2*>(Qb + static_cast<long long>(g) * HD + q * DPL + 4 * s2);
    qa0[s2] = v.x;
    qa2[s2] = v.y;
  }
  float o[OT][4];
#pragma unroll
  for (int t = 0; t < OT; ++t) { o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }
  float m_run = -INFINITY, l_run = 0.f;
  const int n_chunks = (seq + NT * 8 - 1) / (NT * 8);
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int key0 = ch * NT * 8;
    // ---- scores: NT tiles of 8 keys.  All K loads of the chunk are issued before the first MMA (the kernel is bound by the
    // round trips to L2 / HBM, not by bandwidth: with the loads interleaved into the MMA chain only a few KB per warp were in flight)
    float sc[NT][4];
    uint4 kfrag[NT][KS / 2];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int key = key0 + 16 * (t >> 1) + 4 * (g >> 1) + (g & 1) + 2 * (t & 1);   // the key this lane's B fragment column (n = g) holds
      key = key < seq_pad ? key : seq_pad - 1;
      const bf16* kr = Kb + static_cast<long long>(key) * HD + q * DPL;
#pragma unroll
      for (int s2 = 0; s2 < KS; s2 += 2) kfrag[t][s2 >> 1] = __ldg(reinterpret_cast<const uint4*>(kr + 4 * s2));   // dims of k-steps s2, s2+1
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < KS; s2 += 2) {
        const uint4 kv = kfrag[t][s2 >> 1];
        mma_bf16_16816(sc[t], qa0[s2], 0u, qa2[s2], 0u, kv.x, kv.y);
        mma_bf16_16816(sc[t], qa0[s2 + 1], 0u, qa2[s2 + 1], 0u, kv.z, kv.w);
      }
    }
    // columns 2q, 2q+1 of tile t: keys key0 + 16*(t/2) + 4q + {0,1} + 2*(t%2); mask the tail
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int kk = key0 + 16 * (t >> 1) + 4 * q + 2 * (t & 1);
      if (kk >= seq) sc[t][0] = -INFINITY;
      if (kk + 1 >= seq) sc[t][1] = -INFINITY;
      mx = fmaxf(mx, fmaxf(sc[t][0], sc[t][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);          // chunk 0 always holds valid keys, so m_new is finite
    const float alpha = exp2f((m_run - m_new) * scale_log2);
    m_run = m_new;
    const float nmc = -m_new * scale_log2;
    float psum = 0.f;
    uint32_t pa[NT];   // bf16 pairs (row g): tile t -> keys of k-slot pair
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const float p0 = exp2f(fmaf(sc[t][0], scale_log2, nmc));
      const float p1 = exp2f(fmaf(sc[t][1], scale_log2, nmc));
      psum += p0 + p1;
      pa[t] = pack_bf16x2(p0, p1);
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * alpha + psum;
#pragma unroll
    for (int t = 0; t < OT; ++t) { o[t][0] *= alpha; o[t][1] *= alpha; }
    // ---- O += P V : k-step T = tiles 2T, 2T+1 = keys key0 + 16T + 4q .. +3 (one 8-byte load of V^T per output tile); the loads of
    // half a chunk are issued together, ahead of their MMAs
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint2 vfrag[NT / 4][OT];
#pragma unroll
      for (int T2 = 0; T2 < NT / 4; ++T2) {
        int kk = key0 + 16 * (h * (NT / 4) + T2) + 4 * q;
        kk = kk + 4 <= seq_pad ? kk : seq_pad - 4;    // (P is zero there)
#pragma unroll
        for (int t = 0; t < OT; ++t) vfrag[T2][t] = __ldg(reinterpret_cast<const uint2*>(Vb + static_cast<long long>(8 * t + g) * seq_pad + kk));
      }
#pragma unroll
      for (int T2 = 0; T2 < NT / 4; ++T2) {
        const int TT = h * (NT / 4) + T2;
#pragma unroll
        for (int t = 0; t < OT; ++t) mma_bf16_16816(o[t], pa[2 * TT], 0u, pa[2 * TT + 1], 0u, vfrag[T2][t].x, vfrag[T2][t].y);
      }
    }
  }
  if (g < nrows) {
    const float inv = 1.0f / l_run;
    const int l_idx = bh / heads, head = bh % heads;
    bf16* dst = O + (static_cast<long long>(l_idx) * seq + g) * (heads * HD) + head * HD + 2 * q;
#pragma unroll
    for (int t = 0; t < OT; ++t) *reinterpret_cast<uint32_t*>(dst + 8 * t) = pack_bf16x2(o[t][0] * inv, o[t][1] * inv);
  }
}

template <int HD>
static cudaError_t launch_prefix(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int seq, int seq_pad,
                                 int nrows, float scale, cudaStream_t s) {
  if (nrows > 8 || nrows < 1 || seq_pad < 16) return cudaErrorInvalidValue;
  const float sl = scale * 1.4426950408889634f;
  const int BH = L * heads;
  note_extra_launches(1);
  attention_prefix_rows_kernel<HD><<<(BH + PFX_WARPS - 1) / PFX_WARPS, PFX_WARPS * 32, 0, s>>>(Q, K, Vt, O, BH, heads, seq, seq_pad, nrows, sl);
  return cudaGetLastError();
}

cudaError_t attention_tc(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int head_dim, int seq,
                         int seq_pad, float scale, cudaStream_t s) {
  if (seq_pad % 128 != 0 || seq > seq_pad || seq <= 0) return cudaErrorInvalidValue;
  // a handful of rows beyond a multiple of 128 (DINOv2: 5 + 256) would cost a whole extra query tile: peel them off
  const int rem = seq % A2_BQ;
  static const bool no_peel = getenv("BD_ATT2_NOPEEL") != nullptr;  // debug switch
  const int q_off = (!no_peel && rem > 0 && rem <= 8 && seq > A2_BQ) ? rem : 0;
  cudaError_t err;
  static const char* only = getenv("BD_ATT2_ONLY");  // debug switch: "prefix" or "main"
  if (only && only[0] == 'p') {
    if (!q_off) return cudaSuccess;
    return head_dim == 96 ? launch_prefix<96>(Q, K, Vt, O, L, heads, seq, seq_pad, q_off, scale, s)
                          : launch_prefix<64>(Q, K, Vt, O, L, heads, seq, seq_pad, q_off, scale, s);
  }
  if (only && only[0] == 'm') {
    return head_dim == 96 ? launch_att2<96>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, q_off, s)
                          : launch_att2<64>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, q_off, s);
  }
  if (head_dim == 96) {
    if (q_off && (err = launch_prefix<96>(Q, K, Vt, O, L, heads, seq, seq_pad, q_off, scale, s)) != cudaSuccess) return err;
    return launch_att2<96>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, q_off, s);
  }
  if (head_dim == 64) {
    if (q_off && (err = launch_prefix<64>(Q, K, Vt, O, L, heads, seq, seq_pad, q_off, scale, s)) != cudaSuccess) return err;
    return launch_att2<64>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, q_off, s);
  }
  return cudaErrorInvalidValue;
}

// Attention from a window of query rows per sequence (rows [win_idx[l] * win_rows, +win_rows) over all `seq` keys) into the compact
// O [L, win_rows, heads*head_dim]: the decoder's last layer, whose other rows nothing reads.
cudaError_t attention_tc_window(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int head_dim, int seq,
                                int seq_pad, float scale, const long long* win_idx, int win_rows, cudaStream_t s) {
  if (seq_pad % 128 != 0 || seq > seq_pad || seq <= 0 || !win_idx || win_rows <= 0 || win_rows > seq) return cudaErrorInvalidValue;
  if (head_dim == 96) return launch_att2<96>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, 0, s, win_idx, win_rows);
  if (head_dim == 64) return launch_att2<64>(Q, K, Vt, O, L, heads, seq, seq_pad, scale, 0, s, win_idx, win_rows);
  return cudaErrorInvalidValue;
}

}  // namespace bd
