// Internal launcher interface between the engine (bd_engine.cu) and the kernel translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bd {

typedef __nv_bfloat16 bf16;

// GEMM: out = epilogue(A[M,K] . W[N,K]^T + bias[N])   (nn.Linear layout: W is [out, in], K contiguous)
enum GemmEpiKind {
  EPI_F32 = 0,        // fp32 store, optional row remap + additive table (patch-embed into the token buffer)
  EPI_GELU = 1,       // exact-erf GELU, activation-dtype store (bf16 on the tensor path, fp32 on the SIMT path)
  EPI_RESID = 2,      // resid[M,N] (fp32) += gamma[N] * (acc + bias)   (gamma optional: DINOv2 LayerScale)
  EPI_QKV = 3,        // split q|k|v, per-head RMSNorm on q,k (optional), write Q,K [BH,seq_pad,hd] and V^T [BH,hd,seq_pad]
  EPI_ACT = 4,        // plain activation-dtype store
  EPI_VT = 5,         // internal (gemm_tc2.cu): operands swapped, rows = v features, columns = tokens -> V^T [BH, hd, seq_pad]
};

struct GemmEpi {
  const float* bias = nullptr;
  float* out_f32 = nullptr;   // EPI_F32 / EPI_RESID target
  void* out_act = nullptr;    // EPI_GELU / EPI_ACT target (bf16 or fp32 depending on path), row pitch = N
  const float* gamma = nullptr;
  int ldo = 0;                // row pitch (elements) of out_f32
  // EPI_F32 row remap: out_row = (m / rp_in) * rp_out + rp_off + (m % rp_in); rp_in == 0 -> identity
  int rp_in = 0, rp_out = 0, rp_off = 0;
  const float* addtab = nullptr;  // [rp_in, N], added at row (m % rp_in)
  // EPI_QKV
  void* q = nullptr;   // [BH, seq_pad, hd_pad]   (tensor path: bf16, hd_pad == hd; SIMT: fp32)
  void* k = nullptr;
  void* v = nullptr;   // tensor path: V^T [BH, hd, seq_pad] bf16; SIMT path: V [BH, seq_pad, hd] fp32
  const float* q_norm_w = nullptr;  // [hd] or null (no RMSNorm: DINOv2)
  const float* k_norm_w = nullptr;
  int seq = 0, seq_pad = 0, heads = 0, head_dim = 0;
  float rms_eps = 1e-6f;
};

// --- tensor-core path ---
// CTA-pair tcgen05 GEMM (cta_group::2, 256 x BN tiles) with fused epilogues: gemm_tc2.cu
cudaError_t gemm_tc(const bf16* A, const bf16* W, int M, int N, int K, int epi, const GemmEpi& e, cudaStream_t s);
// Q,K [BH, seq_pad, hd] bf16, Vt [BH, hd, seq_pad] bf16 -> O [L*seq, heads*hd] bf16 (token-major): persistent kernel, two
// query tiles per CTA in ping-pong, P in tensor memory (attn_tc2.cu)
cudaError_t attention_tc(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int head_dim, int seq,
                         int seq_pad, float scale, cudaStream_t s);
// launchers that enqueue more than one kernel report the extra ones here; the engine folds them into bd_launch_count()
void note_extra_launches(int n);
int take_extra_launches();
// per-device plumbing (tc_host.cu): SM count of the current device; dynamic-shared-memory opt-in of a kernel on the current device
int tc_num_sms();
cudaError_t tc_ensure_smem(const void* kernel, int bytes);
const char* tc_last_error();

// --- SIMT fp32 path + memory-bound kernels (kernels_simt.cu) ---
cudaError_t gemm_f32(const float* A, const float* W, int M, int N, int K, int epi, const GemmEpi& e, cudaStream_t s);
cudaError_t attention_tc_window(const bf16* Q, const bf16* K, const bf16* Vt, bf16* O, int L, int heads, int head_dim, int seq,
                                int seq_pad, float scale, const long long* win_idx, int win_rows, cudaStream_t s);
cudaError_t attention_f32(const float* Q, const float* K, const float* V, float* O, int L, int heads, int head_dim,
                          int seq, int seq_pad, float scale, cudaStream_t s);
// LayerNorm over the last dim (d), fp32 in; writes any of: fp32 out, bf16 out.  Row remap for the DINO tail:
// input row = (r / rows_out_per) * rows_in_per + row_off + (r % rows_out_per)   (rows_out_per == 0 -> identity)
cudaError_t layernorm(const float* x, const float* w, const float* b, float eps, float* out_f32, bf16* out_bf16,
                      int rows_out, int d, int rows_out_per, int rows_in_per, int row_off, cudaStream_t s);
// images [L,3,S,S] (fp32 or bf16) -> normalised im2col rows [L*P, kpad] (bf16 or fp32), column order (c, pr, pc)
cudaError_t im2col_patches(const void* images, int img_is_bf16, void* out, int out_is_bf16, int L, int S, int patch,
                           int kpad, cudaStream_t s);
// bbox_feat [L,C,S,S] -> [L*P, patch*patch*C], per-token order (pr, pc, c)   (betr.py:211-228)
cudaError_t patchify_heat(const void* feat, int in_is_bf16, void* out, int out_is_bf16, int L, int C, int S, int patch,
                          cudaStream_t s);
// DINO: rows 0..4 of every sequence: cls + pos[0], 4 register tokens
cudaError_t dino_prefix_tokens(float* X, const float* cls, const float* pos0, const float* reg, int L, int n_tok, int n_reg,
                               int d, cudaStream_t s);
// BETR fusion: X[m] = (view(m) is query ? query_tok : PF[m]) + LN_noaffine(R[m], eps) + pos[m % P]
cudaError_t betr_fuse(const float* PF, const float* R, const float* query_tok, const float* pos, const int64_t* query_idx,
                      float* X, int B, int T, int P, int d, float eps, cudaStream_t s);
// gather the query view's tokens: out[b*P + p] = X[(b*T + qidx[b])*P + p]
cudaError_t gather_query(const float* X, const int64_t* query_idx, float* out_f32, bf16* out_bf16, int B, int T, int P, int d,
                         cudaStream_t s);
// logits [B*P, patch*patch*C] -> heat [B,C,S,S] = 2*sigmoid(l) - 1 (betr.py:230-247, 432-435); optional raw logits image
cudaError_t unpatchify_sigmoid(const float* logits, float* heat, int B, int C, int S, int patch, cudaStream_t s);
cudaError_t cast_f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t s);
cudaError_t pose_metrics(const float* pose_pred, const float* pose_gt, const float* K, const float* pts, long long pts_stride,
                         float* out, int B, int N, cudaStream_t s);
cudaError_t bbox_heatmaps(const float* corners_px, void* out, int out_is_bf16, int L, int S, int group, cudaStream_t s);
// SIMT-path QKV post-processing: qkv [M, 3*d] fp32 -> Q,K (RMSNorm optional), V [BH, seq_pad, hd] fp32
cudaError_t qkv_split_f32(const float* qkv, const GemmEpi& e, int M, cudaStream_t s);

// --- corners + PnP (post.cu) ---
cudaError_t corners_topk(const float* heat, float* corners_px, float* corners_norm, int32_t* idx_out, int B, int C, int S,
                         cudaStream_t s);
struct PnpOpts {
  int mode;          // 0 = reference-parity (DLT on all points -> LM to convergence), 1 = RANSAC hypotheses -> LM on inliers
  int n_hyp;         // mode 1
  float thr_px;      // mode 1 inlier threshold
  uint32_t seed;     // mode 1
  int max_iter;      // LM iterations cap
};
// rec (optional, mode 0, n_pts <= 32): packed [B, 28] record {R|t, 8 normalised corners} written by the kernel's epilogue
// (the payload of the multi-GPU result gather); corners_norm [B, 8, 2] supplies its last 16 floats
cudaError_t pnp_solve(const float* corners_px, const float* bbox3d, const float* K, float* poses, const PnpOpts& o, int B,
                      int n_pts, cudaStream_t s, float* rec = nullptr, const float* corners_norm = nullptr);

}  // namespace bd
