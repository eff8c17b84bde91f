// Debug aid: several formulations of the 12x12 symmetric eigen-solve, host vs device.
#include <cstdio>
#include <cmath>
#include <vector>
#include <algorithm>
#define HD __host__ __device__
#define ROT(theta, c, s) { const double t_ = ((theta) >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt((theta) * (theta) + 1.0)); c = 1.0 / sqrt(t_ * t_ + 1.0); s = t_ * c; }

HD void v1_ref(double (&A)[12][12]) {              // original two-pass form on a local array passed by reference
  for (int sweep = 0; sweep < 40; ++sweep)
    for (int p = 0; p < 11; ++p) for (int q = p + 1; q < 12; ++q) {
      const double apq = A[p][q]; if (apq == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * apq); double c, s; ROT(theta, c, s);
      for (int k = 0; k < 12; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
      for (int k = 0; k < 12; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
    }
}
HD void v3_ptr(double* A) {                        // same, flat pointer (works on global memory too)
  for (int sweep = 0; sweep < 40; ++sweep)
    for (int p = 0; p < 11; ++p) for (int q = p + 1; q < 12; ++q) {
      const double apq = A[p * 12 + q]; if (apq == 0.0) continue;
      const double theta = (A[q * 12 + q] - A[p * 12 + p]) / (2.0 * apq); double c, s; ROT(theta, c, s);
      for (int k = 0; k < 12; ++k) { const double a = A[k * 12 + p], b = A[k * 12 + q]; A[k * 12 + p] = c * a - s * b; A[k * 12 + q] = s * a + c * b; }
      for (int k = 0; k < 12; ++k) { const double a = A[p * 12 + k], b = A[q * 12 + k]; A[p * 12 + k] = c * a - s * b; A[q * 12 + k] = s * a + c * b; }
    }
}
HD void v4_nounroll(double (&A)[12][12]) {
  for (int sweep = 0; sweep < 40; ++sweep)
    for (int p = 0; p < 11; ++p) for (int q = p + 1; q < 12; ++q) {
      const double apq = A[p][q]; if (apq == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * apq); double c, s; ROT(theta, c, s);
#pragma unroll 1
      for (int k = 0; k < 12; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
#pragma unroll 1
      for (int k = 0; k < 12; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
    }
}
HD void v5_volatile(double (&A0)[12][12]) {
  volatile double (*A)[12] = A0;
  for (int sweep = 0; sweep < 40; ++sweep)
    for (int p = 0; p < 11; ++p) for (int q = p + 1; q < 12; ++q) {
      const double apq = A[p][q]; if (apq == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * apq); double c, s; ROT(theta, c, s);
      for (int k = 0; k < 12; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
      for (int k = 0; k < 12; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
    }
}
// smallest eigenvector by shifted inverse iteration with an LDL^T-free Gaussian elimination (candidate replacement)
HD void v6_invit(const double (&A)[12][12], double (&x)[12], double* lam_out) {
  double M[12][12]; double tr = 0; for (int i = 0; i < 12; ++i) tr += A[i][i];
  const double shift = -1e-9 * tr;   // (A - shift I) is SPD
  for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) M[i][j] = A[i][j] - (i == j ? shift : 0.0);
  // Cholesky M = L L^T (in place, lower)
  for (int j = 0; j < 12; ++j) {
    double d = M[j][j]; for (int k = 0; k < j; ++k) d -= M[j][k] * M[j][k];
    d = sqrt(fmax(d, 1e-300)); M[j][j] = d;
    for (int i = j + 1; i < 12; ++i) { double v = M[i][j]; for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k]; M[i][j] = v / d; }
  }
  for (int i = 0; i < 12; ++i) x[i] = 1.0 / (1.0 + i);
  double lam = 0;
  for (int it = 0; it < 60; ++it) {
    double y[12];
    for (int i = 0; i < 12; ++i) { double v = x[i]; for (int k = 0; k < i; ++k) v -= M[i][k] * y[k]; y[i] = v / M[i][i]; }
    for (int i = 11; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 12; ++k) v -= M[k][i] * y[k]; y[i] = v / M[i][i]; }
    double n = 0; for (int i = 0; i < 12; ++i) n += y[i] * y[i]; n = sqrt(n);
    lam = 1.0 / n + shift;   // since |x| = 1: |A^-1 x| ~ 1/(lam - shift)
    for (int i = 0; i < 12; ++i) x[i] = y[i] / n;
  }
  *lam_out = lam;
}
struct Out { double ev[5][12]; double lam6; double x6[12]; };
HD void run(const double* A0, Out* o, double* gscratch) {
  double A[12][12];
  for (int i = 0; i < 144; ++i) (&A[0][0])[i] = A0[i]; v1_ref(A); for (int i = 0; i < 12; ++i) o->ev[0][i] = A[i][i];
  for (int i = 0; i < 144; ++i) gscratch[i] = A0[i]; v3_ptr(gscratch); for (int i = 0; i < 12; ++i) o->ev[1][i] = gscratch[i * 13];
  for (int i = 0; i < 144; ++i) (&A[0][0])[i] = A0[i]; v3_ptr(&A[0][0]); for (int i = 0; i < 12; ++i) o->ev[2][i] = A[i][i];
  for (int i = 0; i < 144; ++i) (&A[0][0])[i] = A0[i]; v4_nounroll(A); for (int i = 0; i < 12; ++i) o->ev[3][i] = A[i][i];
  for (int i = 0; i < 144; ++i) (&A[0][0])[i] = A0[i]; v5_volatile(A); for (int i = 0; i < 12; ++i) o->ev[4][i] = A[i][i];
  for (int i = 0; i < 144; ++i) (&A[0][0])[i] = A0[i]; v6_invit(A, o->x6, &o->lam6);
}
__global__ void k(const double* A0, Out* o, double* g) { run(A0, o, g); }
int main() {
  // DLT-like PSD matrix: sum of outer products of 16 rows
  std::vector<double> A(144, 0.0);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
  for (int r = 0; r < 16; ++r) {
    double row[12] = {0}; double X = 0.2 * rnd(), Y = 0.2 * rnd(), Z = 0.2 * rnd(), u = 0.3 * rnd();
    if (r % 2 == 0) { row[0] = X; row[1] = Y; row[2] = Z; row[3] = 1; } else { row[4] = X; row[5] = Y; row[6] = Z; row[7] = 1; }
    row[8] = -u * X; row[9] = -u * Y; row[10] = -u * Z; row[11] = -u;
    for (int a = 0; a < 12; ++a) for (int b = 0; b < 12; ++b) A[a * 12 + b] += row[a] * row[b];
  }
  Out h, d; std::vector<double> hs(144);
  run(A.data(), &h, hs.data());
  double *dA, *dg; Out* dO; cudaMalloc(&dA, 144 * 8); cudaMalloc(&dg, 144 * 8); cudaMalloc(&dO, sizeof(Out));
  cudaMemcpy(dA, A.data(), 144 * 8, cudaMemcpyHostToDevice);
  k<<<1, 1>>>(dA, dO, dg); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  cudaMemcpy(&d, dO, sizeof(Out), cudaMemcpyDeviceToHost);
  const char* names[5] = {"v1 ref-array", "v3 global ptr", "v3 local ptr", "v4 nounroll", "v5 volatile"};
  for (int v = 0; v < 5; ++v) {
    std::vector<double> a(h.ev[v], h.ev[v] + 12), b(d.ev[v], d.ev[v] + 12); std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
    printf("%-14s host:", names[v]); for (double e : a) printf(" %.5e", e); printf("\n%-14s dev :", ""); for (double e : b) printf(" %.5e", e); printf("\n");
  }
  printf("v6 invit lam host %.8e dev %.8e ; x host/dev:\n", h.lam6, d.lam6);
  for (int i = 0; i < 12; ++i) printf(" %.8f/%.8f", h.x6[i], d.x6[i]); printf("\n");
  return 0;
}
