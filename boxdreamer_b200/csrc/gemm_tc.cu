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
#include <stdlib.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "gemm_epi.cuh"

namespace bd {

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

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches stay convergent, operands stay in uniform registers
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
        const int tt = args.reverse ? num_tiles - 1 - tile : tile;
        const int m_blk = tt / tiles_n, n_blk = tt % tiles_n;
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
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int tt = args.reverse ? num_tiles - 1 - tile : tile;
        const int m_blk = tt / tiles_n, n_blk = tt % tiles_n;
      const int row_w = m_blk * BM + quad * 32;  // first row of this warp
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;

      gemm_epilogue_tile<BN, EPI, HD>(args, tile_s, t_acc, row_w, n_blk, lane, grp);
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
static thread_local int g_extra_launches = 0;
void note_extra_launches(int n) { g_extra_launches += n; }
int take_extra_launches() { const int n = g_extra_launches; g_extra_launches = 0; return n; }
static thread_local int g_reverse = 0;
void tc_set_reverse(int r) { g_reverse = r; }
int tc_reverse() { return g_reverse; }
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

// 2D bf16 tensor [rows, cols] (cols contiguous, row pitch `pitch_elems`), box {box_cols, box_rows}, zero OOB fill; the swizzle
// span equals the box row (128 bytes; 64 bytes for the 32-column boxes of the attention kernel's head-dim-96 operands).
// esz = 2: bf16 elements, esz = 4: fp32 elements (epilogue TMA stores / reduce-adds).
static bool make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_cols,
                         uint32_t box_rows, uint32_t esz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { g_tc_err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols * esz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_tc_err = "cuTensorMapEncodeTiled failed, code " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

struct TmapKey {
  const void* p; uint64_t rows, cols, pitch; uint32_t bc, br, esz;
  bool operator==(const TmapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && pitch == o.pitch && bc == o.bc && br == o.br && esz == o.esz;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.p);
    h = h * 1000003u ^ k.rows; h = h * 1000003u ^ k.cols; h = h * 1000003u ^ k.pitch; h = h * 1000003u ^ k.bc; h = h * 1000003u ^ k.br; h = h * 1000003u ^ k.esz;
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapHash> g_tmaps;
static std::mutex g_tmap_mu;

bool get_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br, uint32_t esz) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  TmapKey k{ptr, rows, cols, pitch, bc, br, esz};
  auto it = g_tmaps.find(k);
  if (it != g_tmaps.end()) { *out = it->second; return true; }
  if (!make_tmap_2d(out, ptr, rows, cols, pitch, bc, br, esz)) return false;
  g_tmaps.emplace(k, *out);
  return true;
}
bool get_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t bc, uint32_t br) {
  return get_tmap_2d(out, ptr, rows, cols, pitch, bc, br, 2);
}

// 3D bf16 tensor (d0 contiguous; d1, d2 with byte strides s1, s2), box {b0, b1, 1}; swizzle span = b0 * 2 bytes (64 or 128).
// Used by the q|k|v^T epilogue stores, which rely on the map clipping d1 at the unpadded sequence length.
struct Tmap3Key {
  const void* p; uint64_t d0, d1, d2, s1, s2; uint32_t b0, b1;
  bool operator==(const Tmap3Key& o) const {
    return p == o.p && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && b0 == o.b0 && b1 == o.b1;
  }
};
struct Tmap3Hash {
  size_t operator()(const Tmap3Key& k) const {
    size_t h = reinterpret_cast<size_t>(k.p);
    h = h * 1000003u ^ k.d0; h = h * 1000003u ^ k.d1; h = h * 1000003u ^ k.d2; h = h * 1000003u ^ k.s1; h = h * 1000003u ^ k.s2;
    h = h * 1000003u ^ k.b0; h = h * 1000003u ^ k.b1;
    return h;
  }
};
static std::unordered_map<Tmap3Key, CUtensorMap, Tmap3Hash> g_tmaps3;

bool get_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0,
                      uint32_t b1) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  Tmap3Key k{ptr, d0, d1, d2, s1, s2, b0, b1};
  auto it = g_tmaps3.find(k);
  if (it != g_tmaps3.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { g_tc_err = "cuTensorMapEncodeTiled entry point unavailable"; return false; }
  const CUtensorMapSwizzle swz = (b0 * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : (b0 * 2 == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_tc_err = "cuTensorMapEncodeTiled (3d) failed, code " + std::to_string(static_cast<int>(r));
    return false;
  }
  g_tmaps3.emplace(k, *out);
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
  args.reverse = tc_reverse();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(tmA, tmB, args);
  return cudaGetLastError();
}

cudaError_t gemm_tc(const bf16* A, const bf16* W, int M, int N, int K, int epi, const GemmEpi& e, cudaStream_t s) {
  const char* pv = getenv("BD_GEMM_PAIR");   // read per call so tests can flip it
  const int use_pair = pv ? atoi(pv) : 1;
  if (use_pair && M > 256) return gemm_tc_pair(A, W, M, N, K, epi, e, s);
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
