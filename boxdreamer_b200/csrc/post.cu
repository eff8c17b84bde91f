// Corner extraction (top-20 mean) and PnP pose solve on the device: no .cpu() sync, no cv2 call.
//   corners_topk  <- recover_bb8_corners, heatmap branch (src/models/utils/box_utils.py:75-110)
//   pnp_solve     <- cv2.solvePnP(SOLVEPNP_ITERATIVE) + cv2.Rodrigues as used at box_utils.py:171-192
#include <limits.h>

#include "bd_internal.h"

namespace bd {

// ---------------------------------------------------------------------------------------------
// top-20 over one S*S heat map per CTA.  HBM-bound: the map is read once from HBM (pass 1) and once more out of L2.
//   pass 1: every thread takes the maximum of its slice (16-byte coalesced loads, no data-dependent work); the 20th
//           largest of the 256 slice maxima is a lower bound T of the 20th largest pixel (20 distinct pixels are >= T);
//   pass 2: pixels >= T are appended to a shared candidate list (a few dozen for real heat maps);
//   select: each thread keeps a sorted top-20 of its share of the candidates in registers, 20 block-argmax rounds merge.
// Degenerate maps (more than TK_CAP pixels >= T, e.g. a saturated constant map) take the exhaustive path: a sorted
// top-20 of the thread's whole slice.  Both paths select the same set.
// Order: larger (x+1)/2 first, ties -> lower flat index (torch.topk leaves ties unspecified).

static constexpr int TOPK = 20;
static constexpr int TK_THREADS = 256;
static constexpr int TK_CAP = 2048;

__device__ __forceinline__ bool tk_better(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }
__device__ __forceinline__ float tk_value(float h) { return (h + 1.0f) * 0.5f; }  // box_utils.py:79

__device__ __forceinline__ void tk_insert(float (&val)[TOPK], int (&idx)[TOPK], float v, int i) {
  if (tk_better(v, i, val[TOPK - 1], idx[TOPK - 1])) {
    val[TOPK - 1] = v;
    idx[TOPK - 1] = i;
#pragma unroll
    for (int k = TOPK - 1; k > 0; --k) {
      if (tk_better(val[k], idx[k], val[k - 1], idx[k - 1])) {
        const float tv = val[k]; val[k] = val[k - 1]; val[k - 1] = tv;
        const int ti = idx[k]; idx[k] = idx[k - 1]; idx[k - 1] = ti;
      }
    }
  }
}

// block-wide argmax of (bv, bi) under tk_better; every thread receives the winner.  Two barriers.
__device__ __forceinline__ void tk_block_best(float& bv, int& bi, float* s_val, int* s_idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (tk_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
  __syncthreads();
  bv = s_val[0]; bi = s_idx[0];
#pragma unroll
  for (int w = 1; w < TK_THREADS / 32; ++w) {
    if (tk_better(s_val[w], s_idx[w], bv, bi)) { bv = s_val[w]; bi = s_idx[w]; }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TK_THREADS) corners_topk_kernel(const float* __restrict__ heat, float* __restrict__ corners_px,
                                                                  float* __restrict__ corners_norm, int32_t* __restrict__ idx_out,
                                                                  int S) {
  const int map = blockIdx.x;
  const int n = S * S;
  const float* hm = heat + static_cast<long long>(map) * n;
  __shared__ float s_val[TK_THREADS / 32];
  __shared__ int s_idx[TK_THREADS / 32];
  __shared__ float c_val[TK_CAP];
  __shared__ int c_idx[TK_CAP];
  __shared__ int c_cnt;
  if (threadIdx.x == 0) c_cnt = 0;
  const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(hm) & 15) == 0);
  const int n4 = vec ? n / 4 : 0;
  // ---- pass 1: slice maxima ----
  float tmax = -INFINITY;
  for (int i = threadIdx.x; i < n4; i += TK_THREADS) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(hm) + i);
    tmax = fmaxf(fmaxf(tmax, fmaxf(tk_value(h.x), tk_value(h.y))), fmaxf(tk_value(h.z), tk_value(h.w)));
  }
  for (int i = n4 * 4 + threadIdx.x; i < n; i += TK_THREADS) tmax = fmaxf(tmax, tk_value(__ldg(hm + i)));
  // 20th largest slice maximum (NaN-free maps; a thread without pixels holds -inf)
  float thr = -INFINITY;
  {
    float mv = tmax;
    int mi = threadIdx.x;   // "index" = owner thread: unique, so exactly one thread pops per round
    for (int round = 0; round < TOPK; ++round) {
      float bv = mv;
      int bi = mi;
      tk_block_best(bv, bi, s_val, s_idx);
      if (mi == bi) { mv = -INFINITY; mi = INT_MAX; }
      thr = bv;
    }
  }
  // ---- pass 2: candidates >= thr ----
  for (int i = threadIdx.x; i < n4; i += TK_THREADS) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(hm) + i);
    const float v[4] = {tk_value(h.x), tk_value(h.y), tk_value(h.z), tk_value(h.w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (v[k] >= thr) {
        const int pos = atomicAdd(&c_cnt, 1);
        if (pos < TK_CAP) { c_val[pos] = v[k]; c_idx[pos] = 4 * i + k; }
      }
    }
  }
  for (int i = n4 * 4 + threadIdx.x; i < n; i += TK_THREADS) {
    const float v = tk_value(__ldg(hm + i));
    if (v >= thr) {
      const int pos = atomicAdd(&c_cnt, 1);
      if (pos < TK_CAP) { c_val[pos] = v; c_idx[pos] = i; }
    }
  }
  __syncthreads();
  const int cnt = c_cnt;
  // ---- select ----
  float val[TOPK];
  int idx[TOPK];
#pragma unroll
  for (int k = 0; k < TOPK; ++k) { val[k] = -INFINITY; idx[k] = INT_MAX; }
  if (cnt <= TK_CAP && cnt >= TOPK) {
    for (int i = threadIdx.x; i < cnt; i += TK_THREADS) tk_insert(val, idx, c_val[i], c_idx[i]);
  } else {  // exhaustive: degenerate map (or fewer than 20 finite pixels)
    for (int i = threadIdx.x; i < n; i += TK_THREADS) tk_insert(val, idx, tk_value(__ldg(hm + i)), i);
  }
  int sum_x = 0, sum_y = 0;
  for (int round = 0; round < TOPK; ++round) {
    float bv = val[0];
    int bi = idx[0];
    tk_block_best(bv, bi, s_val, s_idx);
    if (idx[0] == bi) {  // the unique owner pops its head
#pragma unroll
      for (int k = 0; k < TOPK - 1; ++k) { val[k] = val[k + 1]; idx[k] = idx[k + 1]; }
      val[TOPK - 1] = -INFINITY; idx[TOPK - 1] = INT_MAX;
    }
    sum_x += bi % S;   // box_utils.py:90-91
    sum_y += bi / S;
    if (threadIdx.x == 0 && idx_out) idx_out[static_cast<long long>(map) * TOPK + round] = bi;
  }
  if (threadIdx.x == 0) {
    const float x = static_cast<float>(sum_x) / static_cast<float>(TOPK);  // mean of exact integers
    const float y = static_cast<float>(sum_y) / static_cast<float>(TOPK);
    corners_px[map * 2 + 0] = x;
    corners_px[map * 2 + 1] = y;
    corners_norm[map * 2 + 0] = (x / static_cast<float>(S)) * 2.0f - 1.0f;  // box_utils.py:104-108
    corners_norm[map * 2 + 1] = (y / static_cast<float>(S)) * 2.0f - 1.0f;
  }
}

cudaError_t corners_topk(const float* heat, float* corners_px, float* corners_norm, int32_t* idx_out, int B, int C, int S,
                         cudaStream_t s) {
  if (B * C <= 0) return cudaSuccess;
  corners_topk_kernel<<<B * C, TK_THREADS, 0, s>>>(heat, corners_px, corners_norm, idx_out, S);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// PnP, reference-parity mode: DLT on all points (normalised image coordinates, smallest eigenvector of
// A^T A by cyclic Jacobi) -> nearest rotation (Newton polar iteration) -> Levenberg-Marquardt on the pixel
// reprojection error over all points, run to convergence, all in fp64.  One thread per query.

#define BD_HD __host__ __device__
static constexpr int PNP_MAXPTS = 64;        // iterative mode (thread-per-query kernel: the problem lives in local memory)
static constexpr int PNP_MAXPTS_POOLED = 256;  // robust mode: pooled proposals of the dense multi-round path, 32 sub-batches x 8 corners
                                               // (the reference hands all N*8 pairs to solvePnPRansac, box_utils.py:202-304)
// point mask: bit i = use point i
struct pmask_t {
  unsigned long long w[PNP_MAXPTS_POOLED / 64];
};
BD_HD inline pmask_t pm_none() { pmask_t m; for (int k = 0; k < PNP_MAXPTS_POOLED / 64; ++k) m.w[k] = 0ull; return m; }
BD_HD inline pmask_t pm_first(int n) {   // points 0 .. n-1
  pmask_t m;
  for (int k = 0; k < PNP_MAXPTS_POOLED / 64; ++k) {
    const int r = n - 64 * k;
    m.w[k] = r >= 64 ? ~0ull : (r <= 0 ? 0ull : ((1ull << r) - 1ull));
  }
  return m;
}
BD_HD inline bool pm_test(const pmask_t& m, int i) { return (m.w[i >> 6] >> (i & 63)) & 1ull; }
BD_HD inline void pm_set(pmask_t& m, int i) { m.w[i >> 6] |= 1ull << (i & 63); }
BD_HD inline bool pm_equal(const pmask_t& a, const pmask_t& b) {
  bool eq = true;
  for (int k = 0; k < PNP_MAXPTS_POOLED / 64; ++k) eq = eq && a.w[k] == b.w[k];
  return eq;
}

// Cyclic Jacobi on a symmetric 12x12 (row-major, flat).  NOTE: written with flat indexing and non-unrolled inner
// loops on purpose -- nvcc 12.9 miscompiles the fully unrolled two-pass update on a `double (&)[12][12]` for sm_100a
// (host and device results diverge; scripts/jacobi_variants.cu reproduces it), the flat form is exact.
BD_HD void jacobi_eig_sym12(double* A, double* V) {
  for (int i = 0; i < 144; ++i) V[i] = 0.0;
  for (int i = 0; i < 12; ++i) V[i * 13] = 1.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < 12; ++i) {
      dg += A[i * 13] * A[i * 13];
      for (int j = i + 1; j < 12; ++j) off += A[i * 12 + j] * A[i * 12 + j];
    }
    if (off <= 1e-36 * dg || off == 0.0) break;  // off-diagonal below fp64 resolution of the diagonal
    for (int p = 0; p < 11; ++p) {
      for (int q = p + 1; q < 12; ++q) {
        const double apq = A[p * 12 + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * 13] - A[p * 13]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll 1
        for (int k = 0; k < 12; ++k) {
          const double akp = A[k * 12 + p], akq = A[k * 12 + q];
          A[k * 12 + p] = c * akp - s * akq;
          A[k * 12 + q] = s * akp + c * akq;
        }
#pragma unroll 1
        for (int k = 0; k < 12; ++k) {
          const double apk = A[p * 12 + k], aqk = A[q * 12 + k];
          A[p * 12 + k] = c * apk - s * aqk;
          A[q * 12 + k] = s * apk + c * aqk;
        }
#pragma unroll 1
        for (int k = 0; k < 12; ++k) {
          const double vkp = V[k * 12 + p], vkq = V[k * 12 + q];
          V[k * 12 + p] = c * vkp - s * vkq;
          V[k * 12 + q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

// Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 12x12 (row-major, flat; only the upper
// triangle and the diagonal are read).  Inverse iteration on a Cholesky factor with a Rayleigh-quotient shift: three
// plain steps on A + eps*tr*I find the neighbourhood, then every step re-factors A - mu*I with mu = rho - 1.5*|A x - rho x|,
// which lies below the eigenvalue nearest to rho (so the factorisation exists once x has settled on the smallest one) and
// converges super-linearly -- a handful of steps where the fixed shift needs up to 60 when noise makes the two smallest
// eigenvalues comparable (that tail dominated the latency of one-thread-per-query PnP).  Returns false when a
// factorisation breaks down or the iteration does not settle (the caller then falls back to Jacobi).
BD_HD bool chol12_shifted(const double* A, double mu, double* L) {   // L L^T = A - mu*I ; L: lower triangle, row-major
  for (int j = 0; j < 12; ++j) {
    double d = A[j * 13] - mu;
#pragma unroll 1
    for (int k = 0; k < j; ++k) d -= L[j * 12 + k] * L[j * 12 + k];
    if (!(d > 1e-300)) return false;
    d = sqrt(d);
    L[j * 13] = d;
#pragma unroll 1
    for (int i = j + 1; i < 12; ++i) {
      double v = A[j * 12 + i];   // upper triangle of the symmetric input
#pragma unroll 1
      for (int k = 0; k < j; ++k) v -= L[i * 12 + k] * L[j * 12 + k];
      L[i * 12 + j] = v / d;
    }
  }
  return true;
}
BD_HD bool chol12_step(const double* L, double* x) {   // x <- normalised (L L^T)^-1 x, sign kept; returns false on breakdown
  double y[12];
#pragma unroll 1
  for (int i = 0; i < 12; ++i) {
    double v = x[i];
#pragma unroll 1
    for (int k = 0; k < i; ++k) v -= L[i * 12 + k] * y[k];
    y[i] = v / L[i * 13];
  }
#pragma unroll 1
  for (int i = 11; i >= 0; --i) {
    double v = y[i];
#pragma unroll 1
    for (int k = i + 1; k < 12; ++k) v -= L[k * 12 + i] * y[k];
    y[i] = v / L[i * 13];
  }
  double n = 0.0, dot = 0.0;
  for (int i = 0; i < 12; ++i) { n += y[i] * y[i]; dot += x[i] * y[i]; }
  n = sqrt(n);
  if (!(n > 0.0) || !isfinite(n)) return false;
  const double sc = (dot < 0.0 ? -1.0 : 1.0) / n;
  for (int i = 0; i < 12; ++i) x[i] = y[i] * sc;
  return true;
}
BD_HD bool smallest_eigvec_sym12(const double* A, double* x) {
  double tr = 0.0;
  for (int i = 0; i < 12; ++i) tr += A[i * 13];
  if (!(tr > 0.0)) return false;
  double L[144];
  if (!chol12_shifted(A, -1e-10 * tr, L)) return false;
  {
    double n0 = 0.0;
    for (int i = 0; i < 12; ++i) { x[i] = 1.0 / (1.0 + i); n0 += x[i] * x[i]; }
    n0 = 1.0 / sqrt(n0);
    for (int i = 0; i < 12; ++i) x[i] *= n0;
  }
  for (int it = 0; it < 3; ++it)
    if (!chol12_step(L, x)) return false;
  double mu_prev = -1e-10 * tr;
  for (int it = 0; it < 40; ++it) {
    // Rayleigh quotient and residual with the symmetric product (upper triangle)
    double y[12], rho = 0.0;
#pragma unroll 1
    for (int i = 0; i < 12; ++i) {
      double v = 0.0;
#pragma unroll 1
      for (int k = 0; k < 12; ++k) v += (k >= i ? A[i * 12 + k] : A[k * 12 + i]) * x[k];
      y[i] = v;
      rho += v * x[i];
    }
    double r2 = 0.0;
    for (int i = 0; i < 12; ++i) r2 += (y[i] - rho * x[i]) * (y[i] - rho * x[i]);
    const double r = sqrt(r2);
    if (r <= 1e-15 * tr) return true;                       // residual at the fp64 resolution of the matrix
    double mu = rho - 1.5 * r;
    bool ok = chol12_shifted(A, mu, L);
    if (!ok) {                                               // x still carries a larger eigenvalue: retreat towards the last good shift
      mu = 0.5 * (mu + mu_prev);
      ok = chol12_shifted(A, mu, L);
      if (!ok) { mu = mu_prev; ok = chol12_shifted(A, mu, L); }
      if (!ok) return false;
    }
    mu_prev = mu < mu_prev ? mu_prev : mu;
    double xo[12];
    for (int i = 0; i < 12; ++i) xo[i] = x[i];
    if (!chol12_step(L, x)) return false;
    double diff = 0.0;
    for (int i = 0; i < 12; ++i) diff += (x[i] - xo[i]) * (x[i] - xo[i]);
    if (diff < 1e-28) return true;
  }
  return false;
}

BD_HD __forceinline__ double det3(const double (&R)[3][3]) {
  return R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
         R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
}

// orthogonal polar factor by Newton iteration R <- (R + R^-T)/2
BD_HD void nearest_rotation(double (&R)[3][3], int iters) {
  for (int it = 0; it < iters; ++it) {
    const double d = det3(R);
    if (fabs(d) < 1e-300) return;
    double C[3][3];  // cofactor matrix = det * R^-T
    C[0][0] = R[1][1] * R[2][2] - R[1][2] * R[2][1];
    C[0][1] = R[1][2] * R[2][0] - R[1][0] * R[2][2];
    C[0][2] = R[1][0] * R[2][1] - R[1][1] * R[2][0];
    C[1][0] = R[0][2] * R[2][1] - R[0][1] * R[2][2];
    C[1][1] = R[0][0] * R[2][2] - R[0][2] * R[2][0];
    C[1][2] = R[0][1] * R[2][0] - R[0][0] * R[2][1];
    C[2][0] = R[0][1] * R[1][2] - R[0][2] * R[1][1];
    C[2][1] = R[0][2] * R[1][0] - R[0][0] * R[1][2];
    C[2][2] = R[0][0] * R[1][1] - R[0][1] * R[1][0];
    double delta = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double nv = 0.5 * (R[i][j] + C[i][j] / d);
        delta += (nv - R[i][j]) * (nv - R[i][j]);
        R[i][j] = nv;
      }
    if (delta < 1e-32) break;
  }
}

BD_HD void rodrigues_exp(const double (&w)[3], double (&E)[3][3]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;  // E = I + a K + b K^2
  if (th < 1e-12) { a = 1.0; b = 0.0; }
  else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  const double K[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0.0;
      for (int k = 0; k < 3; ++k) k2 += K[i][k] * K[k][j];
      E[i][j] = (i == j ? 1.0 : 0.0) + a * K[i][j] + b * k2;
    }
}

// solve (H + lam*diag(H)) x = -g by Gaussian elimination with partial pivoting; false if singular
// H + lam*diag(H) is symmetric positive definite whenever the LM step is well posed: unrolled Cholesky with static
// indexing (everything stays in registers); returns false when a pivot is not positive.
BD_HD __forceinline__ bool solve6_chol(const double (&H)[6][6], const double (&g)[6], double lam, double (&x)[6]) {
  double L[6][6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = H[j][j] * (1.0 + lam);
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 0.0)) return false;
    const double inv = 1.0 / sqrt(d);
    L[j][j] = inv;   // reciprocal of the diagonal entry
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double v = H[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v * inv;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
    y[i] = v * L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= L[k][i] * x[k];
    x[i] = v * L[i][i];
  }
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) ok = ok && isfinite(x[i]);
  return ok;
}

BD_HD bool solve6(const double (&H)[6][6], const double (&g)[6], double lam, double (&x)[6]) {
  if (solve6_chol(H, g, lam, x)) return true;
  double M[6][7];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) M[i][j] = H[i][j] + (i == j ? lam * H[i][i] : 0.0);
    M[i][6] = -g[i];
  }
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    double best = fabs(M[c][c]);
    for (int r = c + 1; r < 6; ++r)
      if (fabs(M[r][c]) > best) { best = fabs(M[r][c]); piv = r; }
    if (!(best > 1e-300)) return false;
    if (piv != c)
      for (int j = 0; j < 7; ++j) { const double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
    for (int r = c + 1; r < 6; ++r) {
      const double f = M[r][c] / M[c][c];
      for (int j = c; j < 7; ++j) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double sacc = M[i][6];
    for (int j = i + 1; j < 6; ++j) sacc -= M[i][j] * x[j];
    x[i] = sacc / M[i][i];
  }
  return true;
}

template <int NP>
struct PnpProblemT {
  double X[NP][3];
  double uv[NP][2];
  double fx, fy, cx, cy;
  int n;
};
typedef PnpProblemT<PNP_MAXPTS> PnpProblem;
typedef PnpProblemT<PNP_MAXPTS_POOLED> PnpProblemPooled;
// The LM / cost routines take an optional point mask (nullptr = all points) so hypothesis refits can share one problem.

// camera-frame point and pixel residual of point i
template <int NP>
BD_HD inline void pnp_point(const PnpProblemT<NP>& pb, int i, const double (&R)[3][3], const double (&t)[3], double (&xc)[3], double& ru,
                            double& rv) {
  for (int a = 0; a < 3; ++a) xc[a] = R[a][0] * pb.X[i][0] + R[a][1] * pb.X[i][1] + R[a][2] * pb.X[i][2] + t[a];
  ru = pb.fx * xc[0] / xc[2] + pb.cx - pb.uv[i][0];
  rv = pb.fy * xc[1] / xc[2] + pb.cy - pb.uv[i][1];
}

template <int NP>
BD_HD double reproj_cost(const PnpProblemT<NP>& pb, const double (&R)[3][3], const double (&t)[3], const pmask_t* mask = nullptr) {
  double cost = 0.0;
  for (int i = 0; i < pb.n; ++i) {
    if (mask && !pm_test(*mask, i)) continue;
    double xc[3], ru, rv;
    pnp_point(pb, i, R, t, xc, ru, rv);
    cost += ru * ru + rv * rv;
  }
  return cost;
}

template <int NP>
BD_HD void pnp_dlt_init(const PnpProblemT<NP>& pb, double (&R)[3][3], double (&t)[3], const pmask_t* mask = nullptr) {
  double A[144], V[144];
  for (int i = 0; i < 144; ++i) A[i] = 0.0;
  for (int i = 0; i < pb.n; ++i) {
    if (mask && !pm_test(*mask, i)) continue;
    const double xn = (pb.uv[i][0] - pb.cx) / pb.fx, yn = (pb.uv[i][1] - pb.cy) / pb.fy;
    double r1[12], r2[12];
    for (int j = 0; j < 12; ++j) { r1[j] = 0.0; r2[j] = 0.0; }
    for (int a = 0; a < 3; ++a) {
      r1[a] = pb.X[i][a]; r2[4 + a] = pb.X[i][a];
      r1[8 + a] = -xn * pb.X[i][a]; r2[8 + a] = -yn * pb.X[i][a];
    }
    r1[3] = 1.0; r2[7] = 1.0; r1[11] = -xn; r2[11] = -yn;
    for (int a = 0; a < 12; ++a)
      for (int b = 0; b < 12; ++b) A[a * 12 + b] += r1[a] * r1[b] + r2[a] * r2[b];
  }
  double Rd[3][3], td[3];
  {
    double ev[12];
    if (smallest_eigvec_sym12(A, ev)) {
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rd[a][b] = ev[a * 4 + b];
        td[a] = ev[a * 4 + 3];
      }
    } else {
      jacobi_eig_sym12(A, V);
      int kmin = 0;
      for (int k = 1; k < 12; ++k)
        if (A[k * 13] < A[kmin * 13]) kmin = k;
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rd[a][b] = V[(a * 4 + b) * 12 + kmin];
        td[a] = V[(a * 4 + 3) * 12 + kmin];
      }
    }
  }
  if (det3(Rd) < 0.0) {
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) Rd[a][b] = -Rd[a][b];
      td[a] = -td[a];
    }
  }
  double nrm = 0.0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) nrm += Rd[a][b] * Rd[a][b];
  nrm = sqrt(nrm);
  const double sc = sqrt(3.0) / fmax(nrm, 1e-300);
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) R[a][b] = Rd[a][b] * sc;  // pre-scaled: close to orthogonal already
    t[a] = td[a] * sc;
  }
  nearest_rotation(R, 60);
}

// Levenberg-Marquardt on the pixel reprojection error.  The per-point camera-frame coordinates and residuals are recomputed
// where they are needed (pnp_point: the same expressions, hence the same values) instead of being kept in per-thread arrays of
// the maximum point count -- which is what capped the pooled path at 64 pairs.
template <int NP>
BD_HD void pnp_lm(const PnpProblemT<NP>& pb, double (&R)[3][3], double (&t)[3], int max_iter, const pmask_t* mask = nullptr) {
  double lam = 1e-3;
  double cost = reproj_cost(pb, R, t, mask);
  for (int it = 0; it < max_iter; ++it) {
    double H[6][6], g[6];
    for (int a = 0; a < 6; ++a) { g[a] = 0.0; for (int b = 0; b < 6; ++b) H[a][b] = 0.0; }
    for (int i = 0; i < pb.n; ++i) {
      if (mask && !pm_test(*mask, i)) continue;
      double xc[3], ru, rv;
      pnp_point(pb, i, R, t, xc, ru, rv);
      const double x = xc[0], y = xc[1], z = xc[2];
      const double du[3] = {pb.fx / z, 0.0, -pb.fx * x / (z * z)};
      const double dv[3] = {0.0, pb.fy / z, -pb.fy * y / (z * z)};
      const double Y[3] = {x - t[0], y - t[1], z - t[2]};
      // d(Xc)/d(omega) = -[Y]_x for the left-multiplicative update R <- exp(omega) R
      const double W[3][3] = {{0, Y[2], -Y[1]}, {-Y[2], 0, Y[0]}, {Y[1], -Y[0], 0}};
      double ju[6], jv[6];
      for (int a = 0; a < 3; ++a) {
        ju[a] = du[0] * W[0][a] + du[1] * W[1][a] + du[2] * W[2][a];
        jv[a] = dv[0] * W[0][a] + dv[1] * W[1][a] + dv[2] * W[2][a];
        ju[3 + a] = du[a];
        jv[3 + a] = dv[a];
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        g[a] += ju[a] * ru + jv[a] * rv;
#pragma unroll
        for (int b = a; b < 6; ++b) H[a][b] += ju[a] * ju[b] + jv[a] * jv[b];
      }
    }
#pragma unroll
    for (int a = 1; a < 6; ++a)
#pragma unroll
      for (int b = 0; b < a; ++b) H[a][b] = H[b][a];
    bool improved = false;
    double step = 0.0, dc = 0.0;
    for (int tr = 0; tr < 12; ++tr) {
      double delta[6];
      if (!solve6(H, g, lam, delta)) { lam *= 10.0; continue; }
      const double w[3] = {delta[0], delta[1], delta[2]};
      double E[3][3], Rn[3][3], tn[3];
      rodrigues_exp(w, E);
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rn[a][b] = E[a][0] * R[0][b] + E[a][1] * R[1][b] + E[a][2] * R[2][b];
        tn[a] = t[a] + delta[3 + a];
      }
      const double cn = reproj_cost(pb, Rn, tn, mask);
      if (isfinite(cn) && cn <= cost) {
        improved = true;
        step = 0.0;
        for (int a = 0; a < 6; ++a) step += delta[a] * delta[a];
        step = sqrt(step);
        dc = cost - cn;
        cost = cn;
        for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) R[a][b] = Rn[a][b]; t[a] = tn[a]; }
        lam = fmax(lam * 0.1, 1e-12);
        break;
      }
      lam *= 10.0;
    }
    if (!improved || step < 1e-13 || dc <= 1e-15 * cost + 1e-300) break;
  }
  nearest_rotation(R, 4);
}

__global__ void __launch_bounds__(64) pnp_iterative_kernel(const float* __restrict__ corners, const float* __restrict__ bbox3d,
                                                           const float* __restrict__ Kmat, float* __restrict__ poses, int B,
                                                           int n_pts, int max_iter) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  PnpProblem pb;
  pb.n = n_pts;
  for (int i = 0; i < n_pts; ++i) {
    for (int a = 0; a < 3; ++a) pb.X[i][a] = static_cast<double>(bbox3d[(static_cast<long long>(q) * n_pts + i) * 3 + a]);
    for (int a = 0; a < 2; ++a) pb.uv[i][a] = static_cast<double>(corners[(static_cast<long long>(q) * n_pts + i) * 2 + a]);
  }
  const float* Kq = Kmat + static_cast<long long>(q) * 9;
  pb.fx = Kq[0]; pb.fy = Kq[4]; pb.cx = Kq[2]; pb.cy = Kq[5];
  double R[3][3], t[3];
  pnp_dlt_init(pb, R, t);
  pnp_lm(pb, R, t, max_iter);
  bool ok = true;
  for (int a = 0; a < 3; ++a) {
    ok = ok && isfinite(t[a]);
    for (int b = 0; b < 3; ++b) ok = ok && isfinite(R[a][b]);
  }
  float* P = poses + static_cast<long long>(q) * 16;
  for (int i = 0; i < 16; ++i) P[i] = 0.f;  // failure => zero pose (box_utils.py:136,194-197)
  if (ok) {
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) P[a * 4 + b] = static_cast<float>(R[a][b]);
      P[a * 4 + 3] = static_cast<float>(t[a]);
    }
    P[15] = 1.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// PnP, reference-parity mode, one WARP per query (n_pts <= 32: the timed path, 8 box corners).  Same algorithm as
// pnp_dlt_init + pnp_lm above -- DLT normal matrix, smallest eigenvector by Rayleigh-shifted inverse iteration on a Cholesky
// factor, Newton polar, Levenberg-Marquardt to convergence, all fp64 -- but cooperative: the 12x12 work is spread over 12
// lanes (column-parallel Cholesky, column-oriented substitutions, matrix and factor in shared memory), every lane owns one
// point in the LM loop (its Jacobian rows and residual), the 27 sums of the normal equations are formed in point order through
// shared memory, and the 6x6 solve / rotation update run redundantly in registers so that every decision is warp-uniform.
// The thread-per-query kernel kept its 12x12 arrays in local memory and ran 64 queries on one SM (1.2 ms per batch of 64 at
// BASELINE config 2); this one spreads the batch over B/4 CTAs and takes tens of microseconds.
static constexpr int PW_WARPS = 4;
static constexpr int PW_MAXPTS = 32;
struct PnpW {
  double X[PW_MAXPTS][3], uv[PW_MAXPTS][2];
  double A[144], L[144];
  double x[12];
  double red[27][PW_MAXPTS];
  double tot[28];
};

__device__ __forceinline__ double w_sum(double v) {   // all-lane sum (butterfly)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// L L^T = A - mu*I, column by column: lane i computes entry (i, j); returns a warp-uniform flag
__device__ bool w_chol12(PnpW& w, double mu, int lane) {
  for (int j = 0; j < 12; ++j) {
    double v = 0.0;
    if (lane >= j && lane < 12) {
      v = w.A[j * 12 + lane] - (lane == j ? mu : 0.0);
      for (int k = 0; k < j; ++k) v -= w.L[lane * 12 + k] * w.L[j * 12 + k];
    }
    const double d = __shfl_sync(0xffffffffu, v, j);
    if (!(d > 1e-300)) return false;
    const double sd = sqrt(d);
    if (lane == j) w.L[j * 13] = sd;
    else if (lane > j && lane < 12) w.L[lane * 12 + j] = v / sd;
    __syncwarp();
  }
  return true;
}

// x <- normalised (L L^T)^-1 x with the sign kept; lane i < 12 holds component i.  Returns a warp-uniform flag.
__device__ bool w_chol12_step(PnpW& w, int lane) {
  const double xi = lane < 12 ? w.x[lane] : 0.0;
  double v = xi, yi = 0.0;
  for (int k = 0; k < 12; ++k) {          // forward substitution, column oriented
    double yk = 0.0;
    if (lane == k) { yk = v / w.L[k * 13]; yi = yk; }
    yk = __shfl_sync(0xffffffffu, yk, k);
    if (lane > k && lane < 12) v -= w.L[lane * 12 + k] * yk;
  }
  v = yi;
  double zi = 0.0;
  for (int k = 11; k >= 0; --k) {         // backward substitution with L^T
    double zk = 0.0;
    if (lane == k) { zk = v / w.L[k * 13]; zi = zk; }
    zk = __shfl_sync(0xffffffffu, zk, k);
    if (lane < k) v -= w.L[k * 12 + lane] * zk;
  }
  const double n = sqrt(w_sum(lane < 12 ? zi * zi : 0.0));
  const double dot = w_sum(lane < 12 ? xi * zi : 0.0);
  if (!(n > 0.0) || !isfinite(n)) return false;
  const double sc = (dot < 0.0 ? -1.0 : 1.0) / n;
  __syncwarp();
  if (lane < 12) w.x[lane] = zi * sc;
  __syncwarp();
  return true;
}

__device__ bool w_smallest_eigvec(PnpW& w, int lane) {
  double tr = 0.0;
  for (int i = 0; i < 12; ++i) tr += w.A[i * 13];
  if (!(tr > 0.0)) return false;
  if (!w_chol12(w, -1e-10 * tr, lane)) return false;
  {
    double n0 = 0.0;
    for (int i = 0; i < 12; ++i) n0 += 1.0 / ((1.0 + i) * (1.0 + i));
    if (lane < 12) w.x[lane] = (1.0 / (1.0 + lane)) / sqrt(n0);
    __syncwarp();
  }
  for (int it = 0; it < 3; ++it)
    if (!w_chol12_step(w, lane)) return false;
  double mu_prev = -1e-10 * tr;
  for (int it = 0; it < 40; ++it) {
    double yv = 0.0, xi = 0.0;
    if (lane < 12) {
      for (int k = 0; k < 12; ++k) yv += w.A[lane * 12 + k] * w.x[k];
      xi = w.x[lane];
    }
    const double rho = w_sum(yv * xi);
    const double dr = yv - rho * xi;
    const double r = sqrt(w_sum(lane < 12 ? dr * dr : 0.0));
    if (r <= 1e-15 * tr) return true;
    double mu = rho - 1.5 * r;
    bool ok = w_chol12(w, mu, lane);
    if (!ok) {
      mu = 0.5 * (mu + mu_prev);
      ok = w_chol12(w, mu, lane);
      if (!ok) { mu = mu_prev; ok = w_chol12(w, mu, lane); }
      if (!ok) return false;
    }
    mu_prev = mu < mu_prev ? mu_prev : mu;
    if (!w_chol12_step(w, lane)) return false;
    const double dx = lane < 12 ? w.x[lane] - xi : 0.0;
    if (w_sum(dx * dx) < 1e-28) return true;
  }
  return false;
}

// element j of the two DLT rows of point i (normalised image coordinates xn, yn)
__device__ __forceinline__ void w_dlt_rows(const PnpW& w, int i, int j, double xn, double yn, double& r1, double& r2) {
  const int c = j & 3, blk = j >> 2;
  const double h = c < 3 ? w.X[i][c] : 1.0;
  r1 = blk == 0 ? h : (blk == 2 ? -xn * h : 0.0);
  r2 = blk == 1 ? h : (blk == 2 ? -yn * h : 0.0);
}

// pixel residual / camera-frame point / squared error of this lane's point
__device__ __forceinline__ double w_point_eval(const PnpW& w, int i, const double (&R)[3][3], const double (&t)[3], double fx, double fy,
                                               double cx, double cy, double (&res)[2], double (&Xc)[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) Xc[a] = R[a][0] * w.X[i][0] + R[a][1] * w.X[i][1] + R[a][2] * w.X[i][2] + t[a];
  res[0] = fx * Xc[0] / Xc[2] + cx - w.uv[i][0];
  res[1] = fy * Xc[1] / Xc[2] + cy - w.uv[i][1];
  return res[0] * res[0] + res[1] * res[1];
}
// sum over the points in point order (every lane returns the same value)
__device__ __forceinline__ double w_ordered_sum(PnpW& w, int n, int lane, double mine) {
  __syncwarp();
  if (lane < n) w.red[0][lane] = mine;
  __syncwarp();
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += w.red[0][i];
  return s;
}

__global__ void __launch_bounds__(PW_WARPS * 32) pnp_iterative_warp_kernel(const float* __restrict__ corners, const float* __restrict__ bbox3d,
                                                                         const float* __restrict__ Kmat, float* __restrict__ poses,
                                                                         float* __restrict__ rec, const float* __restrict__ corners_norm,
                                                                         int B, int n, int max_iter) {
  __shared__ PnpW sw[PW_WARPS];
  const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * PW_WARPS + wq;
  if (q >= B) return;
  PnpW& w = sw[wq];
  if (lane < n) {
#pragma unroll
    for (int a = 0; a < 3; ++a) w.X[lane][a] = static_cast<double>(bbox3d[(static_cast<long long>(q) * n + lane) * 3 + a]);
#pragma unroll
    for (int a = 0; a < 2; ++a) w.uv[lane][a] = static_cast<double>(corners[(static_cast<long long>(q) * n + lane) * 2 + a]);
  }
  const float* Kq = Kmat + static_cast<long long>(q) * 9;
  const double fx = Kq[0], fy = Kq[4], cx = Kq[2], cy = Kq[5];
  __syncwarp();
  // ---- DLT normal matrix: entry e = (a, b), summed over the points in order
  for (int e = lane; e < 144; e += 32) {
    const int a = e / 12, b = e % 12;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      const double xn = (w.uv[i][0] - cx) / fx, yn = (w.uv[i][1] - cy) / fy;
      double a1, a2, b1, b2;
      w_dlt_rows(w, i, a, xn, yn, a1, a2);
      w_dlt_rows(w, i, b, xn, yn, b1, b2);
      acc += a1 * b1 + a2 * b2;
    }
    w.A[e] = acc;
  }
  __syncwarp();
  double R[3][3], t[3];
  {
    double Rd[3][3], td[3];
    const bool have = w_smallest_eigvec(w, lane);
    if (!have) {   // rare: cyclic Jacobi by one lane on local copies (same fallback as the thread-per-query kernel)
      __syncwarp();
      if (lane == 0) {
        double Aj[144], Vj[144];
        for (int i = 0; i < 144; ++i) Aj[i] = w.A[i];
        jacobi_eig_sym12(Aj, Vj);
        int kmin = 0;
        for (int k = 1; k < 12; ++k)
          if (Aj[k * 13] < Aj[kmin * 13]) kmin = k;
        for (int i = 0; i < 12; ++i) w.x[i] = Vj[i * 12 + kmin];
      }
      __syncwarp();
    }
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) Rd[a][b] = w.x[a * 4 + b];
      td[a] = w.x[a * 4 + 3];
    }
    if (det3(Rd) < 0.0) {
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rd[a][b] = -Rd[a][b];
        td[a] = -td[a];
      }
    }
    double nrm = 0.0;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) nrm += Rd[a][b] * Rd[a][b];
    nrm = sqrt(nrm);
    const double sc = sqrt(3.0) / fmax(nrm, 1e-300);
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) R[a][b] = Rd[a][b] * sc;
      t[a] = td[a] * sc;
    }
    nearest_rotation(R, 60);
  }
  // ---- Levenberg-Marquardt (pnp_lm): lane i owns point i
  const int pi = lane < n ? lane : 0;
  double res[2], Xc[3];
  double lam = 1e-3;
  double cost = w_ordered_sum(w, n, lane, w_point_eval(w, pi, R, t, fx, fy, cx, cy, res, Xc));
  for (int it = 0; it < max_iter; ++it) {
    {
      const double x = Xc[0], y = Xc[1], z = Xc[2];
      const double du[3] = {fx / z, 0.0, -fx * x / (z * z)};
      const double dv[3] = {0.0, fy / z, -fy * y / (z * z)};
      const double Y[3] = {x - t[0], y - t[1], z - t[2]};
      const double W[3][3] = {{0, Y[2], -Y[1]}, {-Y[2], 0, Y[0]}, {Y[1], -Y[0], 0}};
      double ju[6], jv[6];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ju[a] = du[0] * W[0][a] + du[1] * W[1][a] + du[2] * W[2][a];
        jv[a] = dv[0] * W[0][a] + dv[1] * W[1][a] + dv[2] * W[2][a];
        ju[3 + a] = du[a];
        jv[3 + a] = dv[a];
      }
      __syncwarp();
      if (lane < n) {
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) w.red[e++][lane] = ju[a] * res[0] + jv[a] * res[1];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = a; b < 6; ++b) w.red[e++][lane] = ju[a] * ju[b] + jv[a] * jv[b];
      }
      __syncwarp();
      if (lane < 27) {
        double sacc = 0.0;
        for (int i = 0; i < n; ++i) sacc += w.red[lane][i];
        w.tot[lane] = sacc;
      }
      __syncwarp();
    }
    double H[6][6], g[6];
    {
      int e = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) g[a] = w.tot[e++];
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) { H[a][b] = w.tot[e++]; H[b][a] = H[a][b]; }
    }
    bool improved = false;
    double step = 0.0, dc = 0.0;
    for (int tr = 0; tr < 12; ++tr) {
      double delta[6];
      if (!solve6(H, g, lam, delta)) { lam *= 10.0; continue; }
      const double wv[3] = {delta[0], delta[1], delta[2]};
      double E[3][3], Rn[3][3], tn[3];
      rodrigues_exp(wv, E);
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) Rn[a][b] = E[a][0] * R[0][b] + E[a][1] * R[1][b] + E[a][2] * R[2][b];
        tn[a] = t[a] + delta[3 + a];
      }
      double resn[2], Xcn[3];
      const double cn = w_ordered_sum(w, n, lane, w_point_eval(w, pi, Rn, tn, fx, fy, cx, cy, resn, Xcn));
      if (isfinite(cn) && cn <= cost) {
        improved = true;
        for (int a = 0; a < 6; ++a) step += delta[a] * delta[a];
        step = sqrt(step);
        dc = cost - cn;
        cost = cn;
        for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) R[a][b] = Rn[a][b]; t[a] = tn[a]; }
        res[0] = resn[0]; res[1] = resn[1];
        Xc[0] = Xcn[0]; Xc[1] = Xcn[1]; Xc[2] = Xcn[2];
        lam = fmax(lam * 0.1, 1e-12);
        break;
      }
      lam *= 10.0;
    }
    if (!improved || step < 1e-13 || dc <= 1e-15 * cost + 1e-300) break;
  }
  nearest_rotation(R, 4);
  if (lane == 0) {
    bool ok = true;
    for (int a = 0; a < 3; ++a) {
      ok = ok && isfinite(t[a]);
      for (int b = 0; b < 3; ++b) ok = ok && isfinite(R[a][b]);
    }
    float* P = poses + static_cast<long long>(q) * 16;
    for (int i = 0; i < 16; ++i) P[i] = 0.f;  // failure => zero pose (box_utils.py:136,194-197)
    if (ok) {
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) P[a * 4 + b] = static_cast<float>(R[a][b]);
        P[a * 4 + 3] = static_cast<float>(t[a]);
      }
      P[15] = 1.0f;
    }
  }
  // packed result record of the multi-GPU gather: [R|t (12, row-major 3x4), 8 normalised corners (16)] (dist.py RECORD)
  if (rec != nullptr) {
    __syncwarp();
    float* rq = rec + static_cast<long long>(q) * 28;
    if (lane < 12) rq[lane] = poses[static_cast<long long>(q) * 16 + lane];
    else if (lane < 28 && corners_norm != nullptr) rq[lane] = corners_norm[static_cast<long long>(q) * 16 + (lane - 12)];
  }
}

// ---------------------------------------------------------------------------------------------
// PnP, hypothesis mode (the robust counterpart of cv2.solvePnPRansac, box_utils.py:158-166 / 266-275): one warp per
// query, one lane per hypothesis.  Hypotheses: every 6-point subset solved from scratch (DLT -> LM), then every 5- and
// 4-point subset (and, beyond those, seeded random 5-subsets) refitted from the all-point solution with a few LM steps; each is scored on ALL points (inlier count at thr_px, then truncated squared error), the warp arg-max wins and
// is polished by LM on its inlier set.  Everything stays on the device; nothing is discarded.

__device__ pmask_t nth_subset_mask(int n, int k, int idx) {  // idx-th k-subset of n points in lexicographic order
  pmask_t mask = pm_none();
  int x = 0;
  for (int i = 0; i < k; ++i) {
    for (;; ++x) {
      // number of subsets that start with x at position i: C(n - x - 1, k - i - 1)
      long long c = 1;
      const int nn = n - x - 1, kk = k - i - 1;
      for (int j = 0; j < kk; ++j) c = c * (nn - j) / (j + 1);
      if (idx < c) break;
      idx -= static_cast<int>(c);
    }
    pm_set(mask, x);
    ++x;
  }
  return mask;
}
__device__ int n_choose_k(int n, int k) {   // fits an int for n <= 64, k <= 6 (C(64, 6) = 74 974 368)
  long long c = 1;
  for (int j = 0; j < k; ++j) c = c * (n - j) / (j + 1);
  return static_cast<int>(c);
}
__device__ __forceinline__ unsigned pnp_hash(unsigned seed, unsigned q, unsigned h) {
  unsigned x = seed ^ (q * 2654435761u) ^ (h * 40503u) ^ 0x9e3779b9u;
  x ^= x << 13; x ^= x >> 17; x ^= x << 5;
  x *= 0x2c1b3c6du; x ^= x >> 15;
  return x;
}

__global__ void __launch_bounds__(128) pnp_hypothesis_kernel(const float* __restrict__ corners, const float* __restrict__ bbox3d,
                                                             const float* __restrict__ Kmat, float* __restrict__ poses, int B,
                                                             int n_pts, int n_hyp, float thr_px, unsigned seed, int max_iter) {
  __shared__ PnpProblemPooled s_pb[4];   // 4 x 10.3 KB: the pooled path's points stay in shared memory
  __shared__ double s_seed[4][12];
  const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + wq;
  if (q >= B) return;
  PnpProblemPooled& pb = s_pb[wq];
  for (int i = lane; i < n_pts; i += 32) {
    for (int a = 0; a < 3; ++a) pb.X[i][a] = static_cast<double>(bbox3d[(static_cast<long long>(q) * n_pts + i) * 3 + a]);
    for (int a = 0; a < 2; ++a) pb.uv[i][a] = static_cast<double>(corners[(static_cast<long long>(q) * n_pts + i) * 2 + a]);
  }
  if (lane == 0) {
    pb.n = n_pts;
    const float* Kq = Kmat + static_cast<long long>(q) * 9;
    pb.fx = Kq[0]; pb.fy = Kq[4]; pb.cx = Kq[2]; pb.cy = Kq[5];
  }
  __syncwarp();
  if (lane == 0) {
    double R[3][3], t[3];
    pnp_dlt_init(pb, R, t);
    pnp_lm(pb, R, t, max_iter);
    for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) s_seed[wq][a * 3 + b] = R[a][b]; s_seed[wq][9 + a] = t[a]; }
  }
  __syncwarp();
  const pmask_t all_mask = pm_first(n_pts);
  const double thr2 = static_cast<double>(thr_px) * thr_px;
  // Few points (the 8 corners of one proposal): the subsets are enumerated.  Pooled proposals (n_pts > 12, dense
  // multi-round path): the subset space is sampled -- seeded random 6-point subsets, each solved from scratch: by lexicographic
  // rank up to 64 points (C(64, 6) fits the 32-bit hash), by six seeded draws without replacement beyond.
  const bool ranked = n_pts <= 64;
  const int c6 = ranked && n_pts >= 6 ? n_choose_k(n_pts, 6) : 0, c5 = ranked && n_pts >= 5 ? n_choose_k(n_pts, 5) : 0,
            c4 = ranked ? n_choose_k(n_pts, 4) : 0;
  const bool enumerate = ranked && static_cast<long long>(c6) + c5 + c4 <= 4096;
  // best-so-far of this lane; hypothesis "-1" is the all-point seed itself
  double bestR[3][3], bestT[3];
  int best_inl = -1;
  double best_err = 1e300;
  pmask_t best_mask = all_mask;
  for (int h = lane - 1; h < n_hyp; h += 32) {
    double R[3][3], t[3];
    for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) R[a][b] = s_seed[wq][a * 3 + b]; t[a] = s_seed[wq][9 + a]; }
    if (h >= 0) {
      pmask_t m;
      bool from_scratch;
      if (enumerate) {
        if (h < c6) m = nth_subset_mask(n_pts, 6, h);
        else if (h < c6 + c5) m = nth_subset_mask(n_pts, 5, h - c6);
        else if (h < c6 + c5 + c4) m = nth_subset_mask(n_pts, 4, h - c6 - c5);
        else m = nth_subset_mask(n_pts, 5, static_cast<int>(pnp_hash(seed, q, h) % static_cast<unsigned>(c5 > 0 ? c5 : 1)));
        from_scratch = h < c6;   // 6-point subsets are solved from scratch (independent of the seed's basin)
      } else if (ranked) {
        m = nth_subset_mask(n_pts, 6, static_cast<int>(pnp_hash(seed, q, h) % static_cast<unsigned>(c6)));
        from_scratch = true;
      } else {
        m = pm_none();
        int picked = 0;
        for (unsigned draw = 0; picked < 6; ++draw) {   // rejection on repeats: at most a handful of extra draws for n_pts > 64
          const int i = static_cast<int>(pnp_hash(seed + 0x51ed2701u * (draw + 1u), q, h) % static_cast<unsigned>(n_pts));
          if (!pm_test(m, i)) { pm_set(m, i); ++picked; }
        }
        from_scratch = true;
      }
      if (from_scratch) pnp_dlt_init(pb, R, t, &m);
      pnp_lm(pb, R, t, 6, &m);
    }
    int inl = 0;
    double err = 0.0;
    pmask_t im = pm_none();
    for (int i = 0; i < n_pts; ++i) {
      double xc[3], ru, rv;
      pnp_point(pb, i, R, t, xc, ru, rv);
      const double e2 = ru * ru + rv * rv;
      if (e2 <= thr2) { ++inl; pm_set(im, i); }   // reprojection error only, as cv2's RANSAC callback
      err += fmin(e2, thr2);
    }
    bool ok = isfinite(err);
    for (int a = 0; a < 3; ++a) ok = ok && isfinite(t[a]);
    if (ok && (inl > best_inl || (inl == best_inl && err < best_err))) {
      best_inl = inl; best_err = err; best_mask = im;
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) bestR[a][b] = R[a][b]; bestT[a] = t[a]; }
    }
  }
  // warp arg-max on (inliers, -err, -lane)
  int win = lane;
  int w_inl = best_inl;
  double w_err = best_err;
  for (int o = 16; o > 0; o >>= 1) {
    const int o_inl = __shfl_xor_sync(0xffffffffu, w_inl, o);
    const double o_err = __shfl_xor_sync(0xffffffffu, w_err, o);
    const int o_win = __shfl_xor_sync(0xffffffffu, win, o);
    if (o_inl > w_inl || (o_inl == w_inl && (o_err < w_err || (o_err == w_err && o_win < win)))) { w_inl = o_inl; w_err = o_err; win = o_win; }
  }
  if (lane == win) {
    float* P = poses + static_cast<long long>(q) * 16;
    for (int i = 0; i < 16; ++i) P[i] = 0.f;
    if (best_inl >= 0) {
      if (best_inl >= 4 && !pm_equal(best_mask, all_mask)) pnp_lm(pb, bestR, bestT, max_iter, &best_mask);  // polish on the inliers
      bool ok = true;
      for (int a = 0; a < 3; ++a) { ok = ok && isfinite(bestT[a]); for (int b = 0; b < 3; ++b) ok = ok && isfinite(bestR[a][b]); }
      if (ok) {
        for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) P[a * 4 + b] = static_cast<float>(bestR[a][b]); P[a * 4 + 3] = static_cast<float>(bestT[a]); }
        P[15] = 1.0f;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Pose metrics of the evaluation step (src/lightning/utils/metrics/metric_utils.py): rotation / translation / in-plane
// error (:162-210), mean 2-D projection error of the model points (:224-306), ADD and ADD-S (:331-424) -- one CTA per
// query, the model points stay on the device (the reference moves the batch to the CPU and walks it with a thread pool
// and a cKDTree per query).  ADD-S is a brute-force nearest neighbour over shared-memory tiles of the predicted points.
// out[q] = {rot_deg, trans_norm, inplane_deg, proj2d_mean, add_mean, adds_mean, diameter, 0}.

static constexpr int PM_THREADS = 256;
static constexpr int PM_TILE = 1024;

__device__ __forceinline__ float pm_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < PM_THREADS / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(PM_THREADS) pose_metrics_kernel(const float* __restrict__ pose_pred, const float* __restrict__ pose_gt,
                                                                  const float* __restrict__ Kmat, const float* __restrict__ pts,
                                                                  long long pts_stride, float* __restrict__ out, int N) {
  __shared__ float sp[12], sg[12], sk[9], red[PM_THREADS / 32];
  __shared__ float tile[PM_TILE][3];
  __shared__ float bmin[3][PM_THREADS / 32], bmax[3][PM_THREADS / 32];
  const int q = blockIdx.x;
  if (threadIdx.x < 12) { sp[threadIdx.x] = pose_pred[q * 12 + threadIdx.x]; sg[threadIdx.x] = pose_gt[q * 12 + threadIdx.x]; }
  if (threadIdx.x < 9) sk[threadIdx.x] = Kmat[q * 9 + threadIdx.x];
  __syncthreads();
  const float* P = pts + static_cast<long long>(q) * pts_stride;
  float proj = 0.f, add = 0.f, adds = 0.f;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  // ADD-S needs, for every ground-truth-posed point, its nearest predicted-posed point: tiles of predicted points in smem
  const int n_round = (N + PM_THREADS - 1) / PM_THREADS;
  for (int r = 0; r < n_round; ++r) {
    const int i = r * PM_THREADS + threadIdx.x;
    const bool live = i < N;
    float x = 0.f, y = 0.f, z = 0.f;
    if (live) { x = P[3 * i]; y = P[3 * i + 1]; z = P[3 * i + 2]; }
    float pp[3], pg[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      pp[a] = sp[a * 4] * x + sp[a * 4 + 1] * y + sp[a * 4 + 2] * z + sp[a * 4 + 3];
      pg[a] = sg[a * 4] * x + sg[a * 4 + 1] * y + sg[a * 4 + 2] * z + sg[a * 4 + 3];
    }
    float best = INFINITY;
    for (int t0 = 0; t0 < N; t0 += PM_TILE) {
      __syncthreads();
      for (int j = threadIdx.x; j < PM_TILE && t0 + j < N; j += PM_THREADS) {
        const float u = P[3 * (t0 + j)], v = P[3 * (t0 + j) + 1], w = P[3 * (t0 + j) + 2];
#pragma unroll
        for (int a = 0; a < 3; ++a) tile[j][a] = sp[a * 4] * u + sp[a * 4 + 1] * v + sp[a * 4 + 2] * w + sp[a * 4 + 3];
      }
      __syncthreads();
      const int nt = min(PM_TILE, N - t0);
      if (live) {
        for (int j = 0; j < nt; ++j) {
          const float dx = tile[j][0] - pg[0], dy = tile[j][1] - pg[1], dz = tile[j][2] - pg[2];
          best = fminf(best, dx * dx + dy * dy + dz * dz);
        }
      }
    }
    if (live) {
      adds += sqrtf(best);
      const float dx = pp[0] - pg[0], dy = pp[1] - pg[1], dz = pp[2] - pg[2];
      add += sqrtf(dx * dx + dy * dy + dz * dz);
      float up[3], ug[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        up[a] = sk[a * 3] * pp[0] + sk[a * 3 + 1] * pp[1] + sk[a * 3 + 2] * pp[2];
        ug[a] = sk[a * 3] * pg[0] + sk[a * 3 + 1] * pg[1] + sk[a * 3 + 2] * pg[2];
      }
      const float ex = up[0] / up[2] - ug[0] / ug[2], ey = up[1] / up[2] - ug[1] / ug[2];
      proj += sqrtf(ex * ex + ey * ey);
      mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
      mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
  }
  const float s_proj = pm_block_sum(proj, red), s_add = pm_block_sum(add, red), s_adds = pm_block_sum(adds, red);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = mn[a], hi = mx[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { bmin[a][threadIdx.x >> 5] = lo; bmax[a][threadIdx.x >> 5] = hi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float d2 = 0.f;
    for (int a = 0; a < 3; ++a) {
      float lo = bmin[a][0], hi = bmax[a][0];
      for (int w = 1; w < PM_THREADS / 32; ++w) { lo = fminf(lo, bmin[a][w]); hi = fmaxf(hi, bmax[a][w]); }
      d2 += (hi - lo) * (hi - lo);
    }
    // rotation_diff = R_pred R_gt^T; angle from its trace, in-plane angle from its first column (metric_utils.py:186-208)
    float rd[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) rd[a][b] = sp[a * 4] * sg[b * 4] + sp[a * 4 + 1] * sg[b * 4 + 1] + sp[a * 4 + 2] * sg[b * 4 + 2];
    const float tr = fminf(fmaxf(rd[0][0] + rd[1][1] + rd[2][2], -1.0f), 3.0f);
    float ang = acosf(fminf(fmaxf((tr - 1.0f) * 0.5f, -1.0f), 1.0f)) * 57.29577951308232f;
    const float tx = sp[3] - sg[3], ty = sp[7] - sg[7], tz = sp[11] - sg[11];
    float te = sqrtf(tx * tx + ty * ty + tz * tz);
    if (!isfinite(ang)) ang = 0.f;
    if (!isfinite(te)) te = 0.f;
    float* o = out + q * 8;
    o[0] = ang;
    o[1] = te;
    o[2] = fabsf(atan2f(rd[1][0], rd[0][0]) * 57.29577951308232f);
    o[3] = s_proj / N;
    o[4] = s_add / N;
    o[5] = s_adds / N;
    o[6] = sqrtf(d2);
    o[7] = 0.f;
  }
}

cudaError_t pose_metrics(const float* pose_pred, const float* pose_gt, const float* K, const float* pts, long long pts_stride,
                         float* out, int B, int N, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if (N <= 0) return cudaErrorInvalidValue;
  pose_metrics_kernel<<<B, PM_THREADS, 0, s>>>(pose_pred, pose_gt, K, pts, pts_stride, out, N);
  return cudaGetLastError();
}

cudaError_t pnp_solve(const float* corners_px, const float* bbox3d, const float* K, float* poses, const PnpOpts& o, int B,
                      int n_pts, cudaStream_t s, float* rec, const float* corners_norm) {
  if (B <= 0) return cudaSuccess;
  if (o.mode != 0 && o.mode != 1) return cudaErrorNotSupported;
  if (n_pts < 6 || n_pts > (o.mode == 1 ? PNP_MAXPTS_POOLED : PNP_MAXPTS)) return cudaErrorInvalidValue;
  const int max_iter = o.max_iter > 0 ? o.max_iter : 30;  // converged within 15 on realistic corners; OpenCV caps its LM at 20
  if (o.mode == 1) {
    const int n_hyp = o.n_hyp > 0 ? o.n_hyp : 154;
    const float thr = o.thr_px > 0.f ? o.thr_px : 2.0f;
    if (rec != nullptr) return cudaErrorNotSupported;
    pnp_hypothesis_kernel<<<(B + 3) / 4, 128, 0, s>>>(corners_px, bbox3d, K, poses, B, n_pts, n_hyp, thr, o.seed, max_iter);
    return cudaGetLastError();
  }
  static const bool thread_kernel = getenv("BD_PNP_THREAD") != nullptr;   // debug switch: the thread-per-query kernel
  if (n_pts <= PW_MAXPTS && (!thread_kernel || rec != nullptr)) {
    pnp_iterative_warp_kernel<<<(B + PW_WARPS - 1) / PW_WARPS, PW_WARPS * 32, 0, s>>>(corners_px, bbox3d, K, poses, rec, corners_norm, B, n_pts,
                                                                                    max_iter);
    return cudaGetLastError();
  }
  if (rec != nullptr) return cudaErrorNotSupported;
  pnp_iterative_kernel<<<(B + 63) / 64, 64, 0, s>>>(corners_px, bbox3d, K, poses, B, n_pts, max_iter);
  return cudaGetLastError();
}

}  // namespace bd
