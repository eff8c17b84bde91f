// TEST INFRASTRUCTURE ONLY: runs the product's PnP arithmetic (the __host__ __device__ functions of
// boxdreamer_b200/csrc/post.cu -- DLT initialisation with the Rayleigh-shifted inverse iteration, Newton polar, LM) on the
// host so that the CPU test suite can compare it with the cv2 fixture and with numpy's eigen-decomposition without a GPU.
// It is built by tests/test_pnp_host_math.py into tests/_build/ and is never part of libboxdreamer_b200.so.
#include "../../boxdreamer_b200/csrc/post.cu"

extern "C" int test_pnp_iterative_host(const float* corners, const float* bbox3d, const float* Kmat, float* poses, int B, int n_pts,
                                       int max_iter) {
  using namespace bd;
  if (n_pts < 6 || n_pts > PNP_MAXPTS) return -1;
  for (int q = 0; q < B; ++q) {
    PnpProblem pb;
    pb.n = n_pts;
    for (int i = 0; i < n_pts; ++i) {
      for (int a = 0; a < 3; ++a) pb.X[i][a] = bbox3d[(static_cast<long long>(q) * n_pts + i) * 3 + a];
      for (int a = 0; a < 2; ++a) pb.uv[i][a] = corners[(static_cast<long long>(q) * n_pts + i) * 2 + a];
    }
    const float* Kq = Kmat + static_cast<long long>(q) * 9;
    pb.fx = Kq[0]; pb.fy = Kq[4]; pb.cx = Kq[2]; pb.cy = Kq[5];
    double R[3][3], t[3];
    pnp_dlt_init(pb, R, t);
    pnp_lm(pb, R, t, max_iter);
    float* P = poses + static_cast<long long>(q) * 16;
    for (int i = 0; i < 16; ++i) P[i] = 0.f;
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) P[a * 4 + b] = static_cast<float>(R[a][b]);
      P[a * 4 + 3] = static_cast<float>(t[a]);
    }
    P[15] = 1.0f;
  }
  return 0;
}

// smallest eigenvector of a symmetric PSD 12x12 (row-major): 1 = converged, 0 = the product would fall back to Jacobi
extern "C" int test_smallest_eigvec12_host(const double* A, double* x) { return bd::smallest_eigvec_sym12(A, x) ? 1 : 0; }

extern "C" void test_jacobi12_host(double* A, double* V) { bd::jacobi_eig_sym12(A, V); }
