/*
 * boxdreamer_b200 -- C ABI of the B200-native BoxDreamer inference hot path.
 *
 * The reference (zju3dv/BoxDreamer) is pure Python; it has no FFI of its own.  Each entry point
 * below therefore names the reference *Python* interface it stands in for (paths relative to the
 * reference root).  The Python shim `boxdreamer_b200/model.py` binds these with ctypes and mirrors
 * `src/models/BoxDreamerModel.py:BoxDreamer` (same ctor config, state_dict keys, forward(dict)->dict).
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor is a caller-owned, contiguous device buffer
 *     unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); nothing in the
 *     device-pointer entry points synchronises the host;
 *   - every function returns 0 on success or a negative bd_status; `bd_last_error()` returns a
 *     thread-local message for the last failure;
 *   - a handle owns only its workspace, packed weight copies and TMA descriptors; it is bound to
 *     one device and is not thread-safe (one handle per process/GPU, as in Lightning DDP).
 */
#ifndef BOXDREAMER_B200_H
#define BOXDREAMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bd_engine* bd_handle;

typedef enum {
  BD_OK = 0,
  BD_ERR_INVALID = -1,   /* bad argument / shape (reference: AssertionError, e.g. betr.py:269-271) */
  BD_ERR_CUDA = -2,      /* CUDA runtime / launch failure */
  BD_ERR_STATE = -3,     /* weights missing or not finalised */
  BD_ERR_UNSUPPORTED = -4
} bd_status;

typedef enum { BD_F32 = 0, BD_BF16 = 1 } bd_dtype;

/* BD_PRECISION_EXACT: fp32 SIMT kernels end to end (parity gate against the fp32 reference forward).
 * BD_PRECISION_BF16 : bf16 tcgen05 tensor-core GEMMs/attention with fp32 accumulation, fp32 residual
 *                     stream, LayerNorm, RMSNorm and softmax (the reference's autocast(bf16) flow). */
typedef enum { BD_PRECISION_EXACT = 0, BD_PRECISION_BF16 = 1 } bd_precision;

/* configs/model/transformer.yaml:10-71 (decoder) + hub kwargs of dinov2_vitb14_reg (encoder) */
typedef struct {
  int32_t img_size;        /* decoder.img_size (224 / 336)                     */
  int32_t patch_size;      /* 14                                               */
  int32_t d_model;         /* 768                                              */
  int32_t dec_layers;      /* decoder.num_decoder_layers = 12                  */
  int32_t dec_heads;       /* decoder.nhead = 8 (head_dim 96)                  */
  int32_t dino_layers;     /* 12                                               */
  int32_t dino_heads;      /* 12 (head_dim 64)                                 */
  int32_t dino_registers;  /* 4                                                */
  int32_t dino_pretrain_grid; /* 37 (pos_embed is 1 + 37*37 rows)              */
  int32_t precision;       /* bd_precision                                     */
  int32_t attn_variant;    /* reserved (one attention kernel is built: csrc/attn_tc2.cu); pass 2                  */
  int32_t max_batch;       /* B the workspace is sized for                     */
  int32_t max_views;       /* T (references + query)                           */
} bd_config;

const char* bd_last_error(void);
int bd_version(void);

/* BoxDreamer.__init__ (BoxDreamerModel.py:24-110): allocates workspace for max_batch x max_views. */
int bd_create(bd_handle* out, const bd_config* cfg);
int bd_destroy(bd_handle h);

/* load_state_dict (run.py:172-183; keys of SURVEY.md section 8a): one call per tensor.
 * `name` is the reference's key with the leading "BoxDreamer." stripped: "decoder.*" for the 177
 * BETR tensors, "dino.*" for the DINOv2 ViT-B/14-reg tensors (dino.pos_embed is passed already
 * interpolated to 1 + (img_size/14)^2 rows; the bicubic-antialias resample stays in torch, see
 * vision_transformer.py:179-211).  `data` may be a host or device pointer (fp32, contiguous). */
int bd_load_weight(bd_handle h, const char* name, const void* data, const int64_t* shape, int32_t ndim);
/* pack weights for the selected precision (bf16 copies, padded patch-embed matrix); checks completeness */
int bd_finalize_weights(bd_handle h);

/* DinoV2Wrapper.predict (encoder/dinov2.py:48-60): images [L,3,S,S] in [0,1] -> patch tokens [L,P,768] fp32 */
int bd_dino_forward(bd_handle h, const void* images, int32_t images_dtype, float* feats_out, int32_t L, void* stream);

/* BETR.forward (backbone/betr.py:249-308): bbox_feat [B,T,8,S,S], feats [B,T,P,768] fp32, query_idx [B] int64
 * -> heat_out [B,8,S,S] fp32 (= query_ret, 2*sigmoid-1), logits_out [B*P,1568] fp32 (nullable) */
int bd_decoder_forward(bd_handle h, const void* bbox_feat, int32_t bbox_dtype, const float* feats, const int64_t* query_idx,
                       float* heat_out, float* logits_out, int32_t B, int32_t T, void* stream);

/* recover_bb8_corners, heatmap branch (utils/box_utils.py:75-110): heat [B,8,S,S] fp32 ->
 * corners_px [B,8,2], corners_norm [B,8,2], idx_out [B,8,20] int32 (nullable; descending value, ties -> lower index) */
int bd_corners_topk(bd_handle h, const float* heat, float* corners_px, float* corners_norm, int32_t* idx_out, int32_t B,
                    int32_t S, void* stream);

typedef struct {
  int32_t mode;      /* 0: reference parity = cv2.solvePnP(SOLVEPNP_ITERATIVE) semantics (DLT on all points -> LM)
                        1: hypothesis mode = the cv2.solvePnPRansac(..., ITERATIVE) counterpart (subset solves scored on all
                           points by reprojection error, LM refit on the winner's inliers).  n <= 12: the 6/5/4-point subsets
                           are enumerated; pooled proposals of the dense multi-round path (utils/box_utils.py:202-304,
                           n = 8 x sub-batches, up to 64): seeded random 6-point subsets */
  int32_t n_hyp;     /* mode 1: hypotheses per query                   */
  float thr_px;      /* mode 1: inlier threshold in pixels             */
  uint32_t seed;     /* mode 1                                         */
  int32_t max_iter;  /* LM iteration cap (0 -> 30)                     */
} bd_pnp_opts;

/* recover_pose_from_bb8 (utils/box_utils.py:113-199) / recover_pose_from_dense_bb8 (:202-304):
 * corners_px [B,n,2], bbox3d [B,n,3], K [B,3,3] (fp32), 6 <= n <= 64 (mode 0) / 256 (mode 1: pooled proposals of up to 32 sub-batches)
 * -> poses [B,4,4] fp32 world->camera (OpenCV convention); a failed solve leaves the zero matrix. */
int bd_pnp(bd_handle h, const float* corners_px, const float* bbox3d, const float* K, float* poses_out, const bd_pnp_opts* opts,
           int32_t B, int32_t n_pts, void* stream);

/* BoxDreamer.forward (BoxDreamerModel.py:112-191), eval, bb8/heatmap: the whole path on `stream`.
 * images [B,T,3,S,S], bbox_feat [B,T,8,S,S] (same dtype), query_idx [B], bbox3d_q [B,8,3], K_q [B,3,3]
 * (query rows, fp32) -> heat_out [B,8,S,S], corners_px/norm [B,8,2], poses_out [B,4,4].
 * heat_out may be NULL (an internal buffer is used).
 * Shapes of at most BOXDREAMER_B200_GRAPH_MAX_VIEWS (default 24) views are launch-bound: their inputs are staged into the
 * handle's buffers and the ~220 launches are replayed as one CUDA graph (BOXDREAMER_B200_GRAPHS=0: always launch eagerly). */
int bd_forward(bd_handle h, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
               const float* bbox3d_q, const float* K_q, float* heat_out, float* corners_px, float* corners_norm,
               float* poses_out, const bd_pnp_opts* opts, int32_t B, int32_t T, void* stream);

/* The same forward, returning only what a multi-GPU evaluation exchanges: rec_out [B,28] fp32 = {R|t (12, row-major 3x4),
 * 8 normalised corners (16)} per query, written by the PnP kernel's epilogue (no packing pass).  It is the payload of the
 * per-step result all-gather that replaces the pickled all_gather of src/utils/comm.py:179-219 (boxdreamer_b200/dist.py).
 * Iterative PnP (opts->mode 0) only. */
int bd_forward_packed(bd_handle h, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
                      const float* bbox3d_q, const float* K_q, float* rec_out, const bd_pnp_opts* opts, int32_t B, int32_t T,
                      void* stream);

/* Same, with HOST buffers (pinned or pageable): stages H2D, runs, copies corners + poses (and the heat maps
 * when heat_out_host != NULL) back and synchronises.  This is what a ctypes/cgo/JNI caller without device
 * memory of its own would bind. */
int bd_forward_host(bd_handle h, const void* images_host, const void* bbox_feat_host, int32_t in_dtype,
                    const int64_t* query_idx_host, const float* bbox3d_q_host, const float* K_q_host, float* heat_out_host,
                    float* corners_px_host, float* corners_norm_host, float* poses_out_host, const bd_pnp_opts* opts,
                    int32_t B, int32_t T);

/* Pipelined form of the two host-buffer entries (bd_forward_host, and bd_forward_host_px below): `submit` enqueues the H2D
 * copies, the forward and the D2H copies of the results for staging slot 0 or 1 and returns at once; `wait` blocks until that
 * slot's results are in the host buffers passed to submit (which, like the inputs, must stay valid until then).  Alternating
 * the two slots -- submit(k+1) before wait(k) -- hides the input transfer of the next batch behind the current batch's
 * compute; this is the loop a data-loader thread of the reference (run.py / the Lightning predict loop) would drive.
 * Exactly one of bbox_feat_host (maps, as bd_forward_host) and bbox_px_host (projected corners, as bd_forward_host_px) is
 * non-NULL.  bd_forward_host(...) == submit(slot 0) + wait(slot 0). */
int bd_forward_host_submit(bd_handle h, int32_t slot, const void* images_host, const void* bbox_feat_host,
                           const float* bbox_px_host, int32_t in_dtype, const int64_t* query_idx_host, const float* bbox3d_q_host,
                           const float* K_q_host, float* heat_out_host, float* corners_px_host, float* corners_norm_host,
                           float* poses_out_host, const bd_pnp_opts* opts, int32_t B, int32_t T);
int bd_forward_host_wait(bd_handle h, int32_t slot);

/* ---- input synthesis on the device ---- */

/* make_bbox_features(bbox, type="heatmap", shape=(S,S)) of the dataset (src/datasets/utils/base/bbox_utils.py:263-303):
 * bbox_px [L,8,2] fp32 projected box corners in crop pixels (make_proj_bbox, camera_utils.py:62-84)
 * -> out [L,8,S,S] in out_dtype (BD_F32 | BD_BF16; the dataset casts to its `precision`, src/datasets/base.py:715-765).
 * `group`: consecutive views rasterised by ONE reference call -- corner i's maps are divided by their maximum over the call
 * (bbox_utils.py:296); the dataset calls per sample, so group = T (must divide L). */
int bd_make_bbox_features(const float* bbox_px, void* out, int32_t out_dtype, int32_t L, int32_t S, int32_t group, void* stream);

/* bd_forward_host with the reference heat maps rasterised on the device: the caller ships 64 bytes per view
 * (bbox_px_host [B,T,8,2] fp32) instead of 8*S*S elements; everything else as bd_forward_host. */
int bd_forward_host_px(bd_handle h, const void* images_host, const float* bbox_px_host, int32_t in_dtype,
                       const int64_t* query_idx_host, const float* bbox3d_q_host, const float* K_q_host, float* heat_out_host,
                       float* corners_px_host, float* corners_norm_host, float* poses_out_host, const bd_pnp_opts* opts,
                       int32_t B, int32_t T);

/* ---- evaluation metrics on the device ---- */

/* Per-query pose metrics of src/lightning/utils/metrics/metric_utils.py: query_pose_error (:162-210),
 * process_single_bs_2d (:255-306), process_single_bs_add (:331-424).  pose_pred / pose_gt [B,3,4] fp32 (the prediction
 * already scaled and moved to the ground truth's frame, :280-281), K [B,3,3], model_pts [N,3] fp32 shared by all queries
 * (pts_stride = 0) or one cloud per query (pts_stride = 3*N).
 * -> out [B,8] = {rotation error (deg), |t_pred - t_gt| (pose units), in-plane rotation error (deg), mean 2-D projection
 *    error (px), ADD mean distance, ADD-S mean nearest-neighbour distance, model diameter (bounding-box diagonal), 0}. */
int bd_pose_metrics(const float* pose_pred, const float* pose_gt, const float* K, const float* model_pts, int64_t pts_stride,
                    float* out, int32_t B, int32_t N, void* stream);

/* ---- kernel-level entry points (used by the unit tests and bench.py's roofline leg) ---- */

/* out = epilogue(A[M,K] . W[N,K]^T + bias).  precision selects the kernel (bf16: A,W bf16; exact: fp32).
 * epilogue: 0 fp32 store, 1 GELU(erf) -> activation dtype, 2 resid += gamma*(acc+bias), 4 activation-dtype store */
int bd_gemm(const void* A, const void* W, const float* bias, const float* gamma, void* out, int32_t M, int32_t N, int32_t K,
            int32_t epilogue, int32_t precision, void* stream);
/* fused QKV projection: x [L*seq, d] . Wqkv[3d, d]^T + b -> (per-head RMSNorm on q,k when q_norm_w != NULL)
 * -> Q,K [L*heads, seq_pad, hd], V (bf16: V^T [L*heads, hd, seq_pad]; exact: [L*heads, seq_pad, hd]).
 * scratch: exact path only, [L*seq, 3d] fp32. */
int bd_qkv_project(const void* x, const void* W, const float* bias, const float* q_norm_w, const float* k_norm_w, void* Q,
                   void* K, void* V, void* scratch, int32_t L, int32_t seq, int32_t seq_pad, int32_t heads, int32_t head_dim,
                   int32_t precision, void* stream);
/* O [L*seq, heads*hd] = softmax(scale * Q K^T) V on the layouts above; `variant` is reserved (ignored) */
int bd_attention(const void* Q, const void* K, const void* V, void* O, int32_t L, int32_t heads, int32_t head_dim, int32_t seq,
                 int32_t seq_pad, float scale, int32_t precision, int32_t variant, void* stream);
int bd_layernorm(const float* x, const float* w, const float* b, float eps, float* out_f32, void* out_bf16, int32_t rows,
                 int32_t d, void* stream);

/* ---- instrumentation (bench.py: gpu_launches and the in-step roofline timing) ---- */
enum {
  BD_PROF_GEMM_QKV = 0, BD_PROF_ATTENTION = 1 /* decoder (BETR) attention */, BD_PROF_GEMM_PROJ = 2, BD_PROF_GEMM_FC1 = 3, BD_PROF_GEMM_FC2 = 4,
  BD_PROF_GEMM_OTHER = 5, BD_PROF_LAYERNORM = 6, BD_PROF_GLUE = 7, BD_PROF_TOPK = 8, BD_PROF_PNP = 9,
  BD_PROF_ATTENTION_DINO = 10, BD_PROF_ATTENTION_WINDOW = 11 /* the decoder's last block: query-window launch */, BD_PROF_NCAT = 12
};
/* kernels this handle has launched so far */
long long bd_launch_count(bd_handle h);
/* when on, every kernel launch is bracketed by CUDA events on its stream */
int bd_profile_enable(bd_handle h, int32_t on);
/* synchronises, then returns accumulated milliseconds and launch counts per BD_PROF_* category */
int bd_profile_read(bd_handle h, double* ms_out, int64_t* count_out, int32_t reset);

/* debug aid: clock64 stamps of the persistent attention kernel's roles (CTA 0) into dev_buf[3*512] (int64); NULL disables */
int bd_debug_attention_trace(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* BOXDREAMER_B200_H */
