// Host-side plumbing shared by the tensor-core kernels: tensor-map (TMA descriptor) cache, per-device kernel attributes and
// SM counts, error text, and the counter of extra launches a launcher enqueues.
#include <stdlib.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include <cuda.h>

#include "bd_internal.h"

namespace bd {

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

// Per-device state: SM count and the (kernel, device) pairs whose dynamic-shared-memory opt-in has been set.
// cudaFuncSetAttribute is per device, so one process driving several GPUs must repeat it on each of them.
static std::mutex g_dev_mu;
static int g_sms[64] = {0};
int tc_num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_sms[dev] == 0) cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return g_sms[dev];
}
static std::unordered_map<uint64_t, int> g_smem_attr;   // key: kernel address ^ device
cudaError_t tc_ensure_smem(const void* kernel, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t key = reinterpret_cast<uint64_t>(kernel) * 64u + static_cast<uint64_t>(dev & 63);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  auto it = g_smem_attr.find(key);
  if (it != g_smem_attr.end() && it->second >= bytes) return cudaSuccess;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err == cudaSuccess) g_smem_attr[key] = bytes;
  return err;
}

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


}  // namespace bd
