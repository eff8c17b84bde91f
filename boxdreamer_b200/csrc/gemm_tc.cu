// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   out = epilogue( A[M,K] (bf16, K-contiguous) . W[N,K]^T (bf16, nn.Linear layout) + bias )
//
// One CTA per SM, 12 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, NSTAGE-deep smem ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, accumulators in TMEM, double-buffered)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       (tcgen05.ld 32x32b -> registers -> fused math -> smem transpose -> coalesced global)
// Replaces the cuBLAS + element-wise sequences of the reference's nn.Linear call sites
// (blocks.py:248,300,859-867; betr.py:151-172; DINOv2 layers/{attention,mlp,patch_embed}.py).
#include <mutex>
#include <string>
#include <unordered_map>

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

template <int BN, int EPI, int HD>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + NSTAGE * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                   // [NSTAGE]
  uint64_t* empty_bar = bars + NSTAGE;         // [NSTAGE]
  uint64_t* tfull_bar = bars + 2 * NSTAGE;     // [2]
  uint64_t* tempty_bar = bars + 2 * NSTAGE + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int M = args.M, N = args.N, K = args.K;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = (M + BM - 1) / BM;
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
      mbar_init(&tempty_bar[i], N_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
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
    // ===================== TMA producer (whole warp, elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx_w(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d_w(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d_w(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, elected lane issues) =====================
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      const uint32_t stage_a = smem_u32(stage_base);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
            // advance along K inside the 128-byte swizzle row: +32 bytes -> +2 in the (addr >> 4) field
            umma_ss_bf16_w(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_w(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int ew = warp - EPI_WARP0;  // 0..7
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int grp = ew >> 2;          // column half
    uint32_t* tile_s = staging + ew * 1024;
    const GemmEpi& e = args.e;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const int row_w = m_blk * BM + quad * 32;  // first row of this warp
      const int my_row = row_w + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;

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
          const int which = col0 / d_model;            // 0 q, 1 k, 2 v
          const int head = (col0 % d_model) / HD;
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
      // all TMEM reads of this accumulator stage are complete (every ld was followed by wait::ld)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local std::string g_tc_err;
const char* tc_last_error() { return g_tc_err.c_str(); }
static int g_num_sms = 0;
void tc_set_num_sms(int n) { g_num_sms = n; }

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// 2D bf16 tensor [rows, cols] (cols contiguous, row pitch `pitch_elems`), box {box_cols, box_rows}, 128B swizzle, zero OOB fill.
bool make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_cols,
                       uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { g_tc_err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_tc_err = "cuTensorMapEncodeTiled failed, code " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

struct TmapKey {
  const void* p; uint64_t rows, cols, pitch; uint32_t bc, br;
  bool operator==(const TmapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && pitch == o.pitch && bc == o.bc && br == o.br;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.p);
    h = h * 1000003u ^ k.rows; h = h * 1000003u ^ k.cols; h = h * 1000003u ^ k.pitch; h = h * 1000003u ^ k.bc; h = h * 1000003u ^ k.br;
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapHash> g_tmaps;
static std::mutex g_tmap_mu;

bool get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  TmapKey k{ptr, rows, cols, pitch, bc, br};
  auto it = g_tmaps.find(k);
  if (it != g_tmaps.end()) { *out = it->second; return true; }
  if (!make_tmap_2d_bf16(out, ptr, rows, cols, pitch, bc, br)) return false;
  g_tmaps.emplace(k, *out);
  return true;
}

template <int BN, int EPI, int HD>
static cudaError_t launch(const bf16* A, const bf16* W, int M, int N, int K, const GemmEpi& e, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  if (!get_tmap_2d_bf16(&tmA, A, M, K, K, BK, BM)) return cudaErrorInvalidValue;
  if (!get_tmap_2d_bf16(&tmB, W, N, K, K, BK, BN)) return cudaErrorInvalidValue;
  auto kern = gemm_tc_kernel<BN, EPI, HD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  GemmArgs args{M, N, K, e};
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(tmA, tmB, args);
  return cudaGetLastError();
}

cudaError_t gemm_tc(const bf16* A, const bf16* W, int M, int N, int K, int epi, const GemmEpi& e, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || (N % 4) != 0) {
    g_tc_err = "gemm_tc: K must be a positive multiple of 8 and N a multiple of 4";
    return cudaErrorInvalidValue;
  }
  switch (epi) {
    case EPI_F32: return launch<256, EPI_F32, 32>(A, W, M, N, K, e, s);
    case EPI_RESID: return launch<256, EPI_RESID, 32>(A, W, M, N, K, e, s);
    case EPI_GELU: return launch<256, EPI_GELU, 32>(A, W, M, N, K, e, s);
    case EPI_ACT: return launch<256, EPI_ACT, 32>(A, W, M, N, K, e, s);
    case EPI_QKV:
      if (N != 3 * e.heads * e.head_dim || (e.heads * e.head_dim) % 192 != 0) { g_tc_err = "gemm_tc: bad QKV shape"; return cudaErrorInvalidValue; }
      if (e.head_dim == 96) return launch<192, EPI_QKV, 96>(A, W, M, N, K, e, s);
      if (e.head_dim == 64) return launch<192, EPI_QKV, 64>(A, W, M, N, K, e, s);
      g_tc_err = "gemm_tc: head_dim must be 64 or 96";
      return cudaErrorInvalidValue;
  }
  g_tc_err = "gemm_tc: unknown epilogue";
  return cudaErrorInvalidValue;
}

}  // namespace bd
