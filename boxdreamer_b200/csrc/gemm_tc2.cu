// CTA-pair tcgen05 GEMM (cta_group::2): a 256 x BN output tile per pair of SMs.
//
// A single-CTA tcgen05.mma with both operands in shared memory is operand-bandwidth bound (12 KB of smem reads per
// 128-cycle 128x256x16 MMA).  In a CTA pair each SM stages its own 128 rows of A and only HALF of the B tile
// (BN/2 rows); the pair's tensor cores share the B halves, so every SM reads 8 KB per MMA and writes 32 KB instead of
// 48 KB per k-block through TMA -- the configuration cuBLAS uses to reach the B200 GEMM peak.
//
// Roles per CTA (12 warps, same as gemm_tc.cu): warp 0 TMA producer (both CTAs load their own tiles; the bytes are
// accounted on the leader's mbarrier), warp 1 MMA issuer (leader CTA only; completion is multicast to both CTAs'
// barriers), warp 2 TMEM allocator (collective cta_group::2 allocation), warps 4-11 the shared fused epilogue
// (gemm_epi.cuh) on the CTA's own 128 accumulator rows.
#include <stdlib.h>

#include <string>

#include "gemm_epi.cuh"

namespace bd {

bool get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br);
bool get_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br, uint32_t esz);
bool get_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0,
                      uint32_t b1);

// TMAOUT: the epilogue leaves through TMA stores (gemm_epilogue_tile_tma) and double-buffers its 4 KB staging boxes
template <int BN, bool TMAOUT>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = TMAOUT ? 5 : 6;
  static constexpr int STAGING_BYTES = N_EPI_WARPS * 32 * 32 * 4 * (TMAOUT ? 2 : 1);
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
};

template <int BN, int EPI, int HD, bool TMAOUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2, const GemmArgs args) {
  using Cfg = Gemm2Cfg<BN, TMAOUT>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + NSTAGE * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                       // [NSTAGE]  used in the leader: TMA bytes of BOTH CTAs
  uint64_t* empty_bar = bars + NSTAGE;             // [NSTAGE]  each CTA: multicast commit of the leader's MMAs
  uint64_t* tfull_bar = bars + 2 * NSTAGE;         // [2]       each CTA: accumulator complete (multicast commit)
  uint64_t* tempty_bar = bars + 2 * NSTAGE + 2;    // [2]       leader: epilogue warps of both CTAs
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches stay convergent, operands stay in uniform registers
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int M = args.M, N = args.N, K = args.K;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * N_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int tt = tile;
      const int m_blk = args.m_fastest ? tt % tiles_m : tt / tiles_n;
      const int n_blk = args.m_fastest ? tt / tiles_m : tt % tiles_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
        if (leader) mbar_expect_tx_w(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
        tma_load_2d_2sm_w(sa, &tmA, leader_full, kb * BK, m_blk * 2 * BM + static_cast<int>(rank) * BM);
        tma_load_2d_2sm_w(sb, &tmB, leader_full, kb * BK, n_blk * BN + static_cast<int>(rank) * (BN / 2));
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      const uint32_t stage_a = smem_u32(stage_base);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = stage_a + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_ss_bf16_2sm_w(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm_w(&empty_bar[stage], 0x3);   // frees the smem slot in both CTAs
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm_w(&tfull_bar[acc], 0x3);       // accumulator complete -> both epilogues
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = warp - EPI_WARP0;
    const int quad = warp & 3;
    const int grp = ew >> 2;
    uint32_t* tile_s = staging + ew * (TMAOUT ? 2048 : 1024);
    uint32_t sbuf_sel = 0;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int tt = tile;
      const int m_blk = args.m_fastest ? tt % tiles_m : tt / tiles_n;
      const int n_blk = args.m_fastest ? tt / tiles_m : tt % tiles_n;
      const int row_w = m_blk * 2 * BM + static_cast<int>(rank) * BM + quad * 32;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      if constexpr (TMAOUT && EPI == EPI_QKV) {
        gemm_epilogue_qk_tma<BN, HD>(args, &tmOut, &tmOut2, reinterpret_cast<uint8_t*>(tile_s), sbuf_sel, t_acc, row_w, n_blk, lane, grp);
      } else if constexpr (TMAOUT && EPI == EPI_VT) {
        gemm_epilogue_vt_tma<BN>(args, &tmOut, reinterpret_cast<uint8_t*>(tile_s), sbuf_sel, t_acc, row_w, n_blk, lane, grp);
      } else if constexpr (TMAOUT) {
        gemm_epilogue_tile_tma<BN, EPI>(args, &tmOut, reinterpret_cast<uint8_t*>(tile_s), sbuf_sel, t_acc, row_w, n_blk, lane, grp);
      } else {
        gemm_epilogue_tile<BN, EPI, HD>(args, tile_s, t_acc, row_w, n_blk, lane, grp);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      }
    }
    if (TMAOUT && lane == 0) bulk_wait_group_read<0>();  // staging boxes stay valid until the last store has read them
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal into this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}


template <int BN, int EPI, int HD, bool TMAOUT>
static cudaError_t launch2(const bf16* A, const bf16* W, int M, int N, int K, const GemmEpi& e, cudaStream_t s, int col_base = 0) {
  using Cfg = Gemm2Cfg<BN, TMAOUT>;
  CUtensorMap tmA, tmB, tmOut;
  if (!get_tmap_2d_bf16(&tmA, A, M, K, K, BK, BM)) return cudaErrorInvalidValue;
  if (!get_tmap_2d_bf16(&tmB, W, N, K, K, BK, BN / 2)) return cudaErrorInvalidValue;
  tmOut = tmA;
  CUtensorMap tmOut2 = tmA;
  bool m_fastest = false;
  if (TMAOUT) {
    bool ok;
    if (EPI == EPI_QKV) {  // rows = tokens (M = L * seq), q | k columns only
      const uint64_t lh = static_cast<uint64_t>(M / e.seq) * e.heads;
      ok = get_tmap_3d_bf16(&tmOut, e.q, HD, e.seq, lh, HD * 2, static_cast<uint64_t>(e.seq_pad) * HD * 2, 32, 32) &&
           get_tmap_3d_bf16(&tmOut2, e.k, HD, e.seq, lh, HD * 2, static_cast<uint64_t>(e.seq_pad) * HD * 2, 32, 32);
    } else if (EPI == EPI_VT) {  // rows = v features (M = d_model), columns = tokens (N = L * seq)
      ok = get_tmap_3d_bf16(&tmOut, e.v, e.seq, M, N / e.seq, static_cast<uint64_t>(e.seq_pad) * 2,
                            static_cast<uint64_t>(M) * e.seq_pad * 2, 64, 32);
      m_fastest = true;
    } else if (EPI == EPI_RESID) {
      ok = get_tmap_2d(&tmOut, e.out_f32, M, N, e.ldo, 32, 32, 4);
    } else {
      ok = get_tmap_2d(&tmOut, e.out_act, M, N, N, 64, 32, 2);
    }
    if (!ok) return cudaErrorInvalidValue;
  }
  auto kern = gemm_tc2_kernel<BN, EPI, HD, TMAOUT>;
  cudaError_t aerr = tc_ensure_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES);
  if (aerr != cudaSuccess) return aerr;
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int max_pairs = tc_num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  GemmArgs args{M, N, K, e, col_base, m_fastest ? 1 : 0};
  kern<<<2 * pairs, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(tmA, tmB, tmOut, tmOut2, args);
  return cudaGetLastError();
}

cudaError_t gemm_tc(const bf16* A, const bf16* W, int M, int N, int K, int epi, const GemmEpi& e, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || (N % 4) != 0) return cudaErrorInvalidValue;
  const char* tv = getenv("BD_GEMM_TMA_EPI");   // debug switch: 0 = register/LSU epilogue everywhere
  const bool tma_out = (N % 64) == 0 && !(tv && atoi(tv) == 0);
  switch (epi) {
    case EPI_F32: return launch2<256, EPI_F32, 32, false>(A, W, M, N, K, e, s);
    case EPI_RESID:
      return tma_out && (e.ldo % 4) == 0 ? launch2<256, EPI_RESID, 32, true>(A, W, M, N, K, e, s)
                                         : launch2<256, EPI_RESID, 32, false>(A, W, M, N, K, e, s);
    case EPI_GELU:
      return tma_out ? launch2<256, EPI_GELU, 32, true>(A, W, M, N, K, e, s) : launch2<256, EPI_GELU, 32, false>(A, W, M, N, K, e, s);
    case EPI_ACT:
      return tma_out ? launch2<256, EPI_ACT, 32, true>(A, W, M, N, K, e, s) : launch2<256, EPI_ACT, 32, false>(A, W, M, N, K, e, s);
    case EPI_QKV:
      if (N != 3 * e.heads * e.head_dim || (e.heads * e.head_dim) % 192 != 0) return cudaErrorInvalidValue;
      if (!(tv && atoi(tv) == 0) && (M % e.seq) == 0 && (e.seq_pad % 8) == 0 && e.seq >= 32 && (e.head_dim == 96 || e.head_dim == 64)) {
        // q|k: rows = tokens, per-head TMA stores.  v: for seq % 64 == 0 a second GEMM with the operands swapped
        // (V^T = W_v . X^T: rows = features, columns = tokens) whose tiles are V^T boxes; otherwise the row-major
        // epilogue on the v columns.
        const int d_model = e.heads * e.head_dim;
        cudaError_t err = e.head_dim == 96 ? launch2<192, EPI_QKV, 96, true>(A, W, M, 2 * d_model, K, e, s)
                                           : launch2<256, EPI_QKV, 64, true>(A, W, M, 2 * d_model, K, e, s);
        if (err != cudaSuccess) return err;
        note_extra_launches(1);
        GemmEpi ev = e;
        ev.bias = e.bias + 2 * d_model;
        const bf16* Wv = W + static_cast<size_t>(2) * d_model * K;
        if ((e.seq % 64) == 0 && (d_model % 32) == 0) return launch2<256, EPI_VT, 32, true>(Wv, A, d_model, M, K, ev, s);
        return e.head_dim == 96 ? launch2<192, EPI_QKV, 96, false>(A, Wv, M, d_model, K, ev, s, 2 * d_model)
                                : launch2<192, EPI_QKV, 64, false>(A, Wv, M, d_model, K, ev, s, 2 * d_model);
      }
      if (e.head_dim == 96) return launch2<192, EPI_QKV, 96, false>(A, W, M, N, K, e, s);
      if (e.head_dim == 64) return launch2<192, EPI_QKV, 64, false>(A, W, M, N, K, e, s);
      return cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

}  // namespace bd
