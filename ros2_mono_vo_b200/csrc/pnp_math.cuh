// pnp_math.cuh -- per-hypothesis FP64 arithmetic of solvePnPRansac (EPnP minimal solver, Rodrigues, small solvers).
// Host + device: the kernels of pnp.cu run it on the GPU; tests/cpp/test_pnp_math.cu compiles the same source for the
// host to check it against oracle/pnp_oracle.py without a GPU.
#pragma once
#include <float.h>
#include <math.h>
#include "linalg.cuh"

namespace mvo {

constexpr int kPnpK = 5;
#ifdef MVO_PNP_DEBUG
#define PNP_DBG(...) printf(__VA_ARGS__)
#else
#define PNP_DBG(...)
#endif   // minimal sample of the registrator (SOLVEPNP_EPNP inside solvePnPRansac)

// ------------------------------------------------------------------------------------------------
// cv::SVD::compute of a symmetric 3 x 3 matrix, left singular vectors only (columns of U), with OpenCV's signs:
// one-sided Jacobi on the rows of A^T, pairs (0,1) (0,2) (1,2), eps = 10 DBL_EPSILON, selection sort.
MVO_HD void cv_svd3_sym(const double* A, double* w, double* U /* 3x3 row-major, columns = vectors */) {
  double At[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) At[i][k] = A[k * 3 + i];
  double W[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) W[i] = At[i][0] * At[i][0] + At[i][1] * At[i][1] + At[i][2] * At[i][2];
  const double eps = DBL_EPSILON * 10;
  for (int iter = 0; iter < 30; ++iter) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = i + 1; j < 3; ++j) {
        const double a = W[i], b = W[j];
        double p = At[i][0] * At[j][0] + At[i][1] * At[j][1] + At[i][2] * At[j][2];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        double na = 0, nb = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double t0 = c * At[i][k] + s * At[j][k];
          const double t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0;
          At[j][k] = t1;
          na += t0 * t0;
          nb += t1 * t1;
        }
        W[i] = na;
        W[j] = nb;
        changed = true;
      }
    if (!changed) break;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) W[i] = sqrt(At[i][0] * At[i][0] + At[i][1] * At[i][1] + At[i][2] * At[i][2]);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int j = i;
#pragma unroll
    for (int k = i + 1; k < 3; ++k)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      const double tw = W[i];
      W[i] = W[j];
      W[j] = tw;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double t = At[i][k];
        At[i][k] = At[j][k];
        At[j][k] = t;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    w[i] = W[i];
    const double s = W[i] > DBL_MIN ? 1.0 / W[i] : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) U[k * 3 + i] = At[i][k] * s;
  }
}

// minimum-norm least squares  min |A x - b|  (A is M x N row-major) through the eigen-decomposition of A^T A,
// singular values below 1e-12 of the largest are dropped (the SVD back-substitution OpenCV's solvers use)
template <int M, int N>
MVO_HD void ls_solve(const double* A, const double* b, double* x) {
  double AtA[N * N], V[N * N], Atb[N];
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int k = 0; k < M; ++k) s += A[k * N + i] * A[k * N + j];
      AtA[i * N + j] = s;
    }
    double s = 0;
    for (int k = 0; k < M; ++k) s += A[k * N + i] * b[k];
    Atb[i] = s;
  }
  PNP_DBG("ls%d AtA %.6g %.6g %.6g %.6g Atb %.6g %.6g %.6g\n", N, AtA[0], AtA[1], AtA[N+1], AtA[N*N-1], Atb[0], Atb[1], Atb[N-1]);
  jacobi_eig_reg<N, true>(AtA, V);   // N <= 5: fully unrolled, AtA and V stay in registers; short scalar chain per rotation
  PNP_DBG("ls%d eig %.6g %.6g %.6g V0 %.6g %.6g %.6g\n", N, AtA[0], AtA[N+1], AtA[N*N-1], V[0], V[1], V[2]);
  double lmax = 0;
  for (int i = 0; i < N; ++i) lmax = fmax(lmax, AtA[i * N + i]);
  for (int i = 0; i < N; ++i) x[i] = 0;
  for (int e = 0; e < N; ++e) {
    const double l = AtA[e * N + e];
    if (!(l > lmax * 1e-24)) continue;
    double p = 0;
    for (int i = 0; i < N; ++i) p += V[i * N + e] * Atb[i];
    p /= l;
    for (int i = 0; i < N; ++i) x[i] += p * V[i * N + e];
  }
}

// least squares  min |A x - b|  of a full-rank M x N system by Householder QR -- what OpenCV's EPnP runs for its
// Gauss-Newton steps (modules/calib3d/src/epnp.cpp, qr_solve: column scaling by the largest entry, reflector
// v = a + sign(a_kk) |a| e_k).  A and b are destroyed.  Returns false for a zero column or a diagonal of R that spans more
// than nine orders of magnitude (the caller falls back to the minimum-norm solver).  Fully unrolled: everything stays in registers; 3 N + N slow operations against the ~40 Jacobi
// rotations of ls_solve.
template <int M, int N>
MVO_HD __forceinline__ bool qr_solve(double* A, double* b, double* x) {
  double A1[N], A2[N];
  double dmin = 1e300, dmax = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double eta = 0;
#pragma unroll
    for (int i = k; i < M; ++i) eta = fmax(eta, fabs(A[i * N + k]));
    if (eta == 0) return false;
    const double inv_eta = 1.0 / eta;
    double sum = 0;
#pragma unroll
    for (int i = k; i < M; ++i) {
      A[i * N + k] *= inv_eta;
      sum += A[i * N + k] * A[i * N + k];
    }
    double sigma = sqrt(sum);
    if (A[k * N + k] < 0) sigma = -sigma;
    A[k * N + k] += sigma;
    A1[k] = sigma * A[k * N + k];
    A2[k] = -eta * sigma;
    dmin = fmin(dmin, fabs(A2[k]));
    dmax = fmax(dmax, fabs(A2[k]));
    const double inv_a1 = 1.0 / A1[k];
#pragma unroll
    for (int j = k + 1; j < N; ++j) {
      double d = 0;
#pragma unroll
      for (int i = k; i < M; ++i) d += A[i * N + k] * A[i * N + j];
      const double tau = d * inv_a1;
#pragma unroll
      for (int i = k; i < M; ++i) A[i * N + j] -= tau * A[i * N + k];
    }
    // the same reflector on b
    double d = 0;
#pragma unroll
    for (int i = k; i < M; ++i) d += A[i * N + k] * b[i];
    const double tau = d * inv_a1;
#pragma unroll
    for (int i = k; i < M; ++i) b[i] -= tau * A[i * N + k];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double d = 0;
#pragma unroll
    for (int j = i + 1; j < N; ++j) d += A[i * N + j] * x[j];
    x[i] = (b[i] - d) / A2[i];
  }
  return dmin > 1e-9 * dmax;   // (|r_kk| of an unpivoted QR: a crude rank test -- a doubtful system goes to the minimum-norm solver)
}

// closest rotation to a 3x3 matrix: R = U V^T of its SVD (polar factor), through the eigenvectors of A^T A
template <bool FAST_CS = false>
MVO_HD void polar_rotation(const double* A, double* R) {
  double AtA[9], V[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) AtA[i * 3 + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
  jacobi_eig_reg<3, FAST_CS>(AtA, V);
  // U_e = A v_e / sigma_e for the two largest singular values, the third completes a right-handed pair of frames
  int order[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = i + 1; j < 3; ++j)
      if (AtA[order[j] * 4] > AtA[order[i] * 4]) {
        const int t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
  double v[3][3], u[3][3];
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int k = 0; k < 3; ++k) v[e][k] = V[k * 3 + order[e]];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    double n2 = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      u[e][r] = A[r * 3] * v[e][0] + A[r * 3 + 1] * v[e][1] + A[r * 3 + 2] * v[e][2];
      n2 += u[e][r] * u[e][r];
    }
    const double s = n2 > 0 ? 1.0 / sqrt(n2) : 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) u[e][r] *= s;
  }
  // third singular pair: u3 = +-(u1 x u2) with the sign of det(A) relative to the frames (U V^T keeps det(A)'s sign)
  double vx[3] = {v[0][1] * v[1][2] - v[0][2] * v[1][1], v[0][2] * v[1][0] - v[0][0] * v[1][2], v[0][0] * v[1][1] - v[0][1] * v[1][0]};
  const double sv = vx[0] * v[2][0] + vx[1] * v[2][1] + vx[2] * v[2][2];   // +-1: handedness of (v1, v2, v3)
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
  u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
  u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  const double sd = det3(A) < 0 ? -1.0 : 1.0;
  const double f = (sv < 0 ? -1.0 : 1.0) * sd;
#pragma unroll
  for (int r = 0; r < 3; ++r) u[2][r] *= f;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = u[0][r] * v[0][c] + u[1][r] * v[1][c] + u[2][r] * v[2][c];
}

// cv::Rodrigues, matrix -> vector (the matrix is a rotation already)
MVO_HD void rotation_to_rvec(const double* R, double* r) {
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      r[0] = r[1] = r[2] = 0;
      return;
    }
    double t = (R[0] + 1) * 0.5;
    rx = sqrt(fmax(t, 0.));
    t = (R[4] + 1) * 0.5;
    ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
    t = (R[8] + 1) * 0.5;
    rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
    if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
    theta /= sqrt(rx * rx + ry * ry + rz * rz);
    r[0] = rx * theta;
    r[1] = ry * theta;
    r[2] = rz * theta;
    return;
  }
  const double vth = 1 / (2 * s) * theta;
  r[0] = rx * vth;
  r[1] = ry * vth;
  r[2] = rz * vth;
}

// cv::Rodrigues, vector -> matrix, optionally with dR/dr (3 x 9: row i = d vec(R) / d r_i)
MVO_HD void rvec_to_rotation(const double* r, double* R, double* J) {
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (J) {
      for (int i = 0; i < 27; ++i) J[i] = 0;
      J[5] = -1; J[7] = 1; J[9 + 2] = 1; J[9 + 6] = -1; J[18 + 1] = -1; J[18 + 3] = 1;
    }
    return;
  }
  double s, c;
  sincos(theta, &s, &c);
  const double c1 = 1. - c, it = 1. / theta;
  const double k[3] = {r[0] * it, r[1] * it, r[2] * it};
  const double rrt[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1], k[1] * k[2], k[0] * k[2], k[1] * k[2], k[2] * k[2]};
  const double rx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i] + s * rx[i];
  if (!J) return;
  const double drrt[27] = {k[0] + k[0], k[1], k[2], k[1], 0, 0, k[2], 0, 0,
                           0, k[0], 0, k[0], k[1] + k[1], k[2], 0, k[2], 0,
                           0, 0, k[0], 0, 0, k[1], k[0], k[1], k[2] + k[2]};
  const double drx[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
  for (int i = 0; i < 3; ++i) {
    const double ri = k[i];
    const double a0 = -s * ri, a1 = (s - 2 * c1 * it) * ri, a2 = c1 * it, a3 = (c - s * it) * ri, a4 = s * it;
    for (int q = 0; q < 9; ++q)
      J[i * 9 + q] = a0 * ((q % 4 == 0) ? 1.0 : 0.0) + a1 * rrt[q] + a2 * drrt[i * 9 + q] + a3 * rx[q] + a4 * drx[i * 9 + q];
  }
}

// ------------------------------------------------------------------------------------------------
// EPnP on N points with normalised image coordinates, in three pieces so that the kernel can run the 12 x 12
// eigen-decomposition between the first two warp-cooperatively and the three beta approximations of the third on
// different lanes (pnp.cu); epnp_solve below chains them on one thread (host harness, oracle checks).

// piece 1: control points (PCA with cv::SVD's signs), barycentric coordinates, M^T M (12 x 12)
template <int N>
MVO_HD void epnp_build(const double (*pw)[3], const double (*us)[2], double (*cws)[3], double (*alphas)[4], double* MtM) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = 0;
    for (int i = 0; i < N; ++i) s += pw[i][k];
    cws[0][k] = s / N;
  }
  {
    double C[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double s = 0;
        for (int i = 0; i < N; ++i) s += (pw[i][a] - cws[0][a]) * (pw[i][b] - cws[0][b]);
        C[a * 3 + b] = s;
      }
    double dc[3], U[9];
    cv_svd3_sym(C, dc, U);
#pragma unroll
    for (int i = 1; i < 4; ++i) {
      const double kk = sqrt(dc[i - 1] / N);
#pragma unroll
      for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + kk * U[j * 3 + i - 1];
    }
  }
  PNP_DBG("cws %.10g %.10g %.10g | %.10g %.10g %.10g | %.10g %.10g %.10g | %.10g %.10g %.10g\n", cws[0][0], cws[0][1], cws[0][2], cws[1][0], cws[1][1], cws[1][2], cws[2][0], cws[2][1], cws[2][2], cws[3][0], cws[3][1], cws[3][2]);
  // barycentric coordinates
  {
    double CC[9], ci[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 1; j < 4; ++j) CC[3 * i + j - 1] = cws[j][i] - cws[0][i];
    const double d = det3(CC);
    const double id = 1.0 / d;
    ci[0] = (CC[4] * CC[8] - CC[5] * CC[7]) * id;
    ci[1] = (CC[2] * CC[7] - CC[1] * CC[8]) * id;
    ci[2] = (CC[1] * CC[5] - CC[2] * CC[4]) * id;
    ci[3] = (CC[5] * CC[6] - CC[3] * CC[8]) * id;
    ci[4] = (CC[0] * CC[8] - CC[2] * CC[6]) * id;
    ci[5] = (CC[2] * CC[3] - CC[0] * CC[5]) * id;
    ci[6] = (CC[3] * CC[7] - CC[4] * CC[6]) * id;
    ci[7] = (CC[1] * CC[6] - CC[0] * CC[7]) * id;
    ci[8] = (CC[0] * CC[4] - CC[1] * CC[3]) * id;
    for (int i = 0; i < N; ++i) {
      const double dx = pw[i][0] - cws[0][0], dy = pw[i][1] - cws[0][1], dz = pw[i][2] - cws[0][2];
#pragma unroll
      for (int j = 0; j < 3; ++j) alphas[i][1 + j] = ci[3 * j] * dx + ci[3 * j + 1] * dy + ci[3 * j + 2] * dz;
      alphas[i][0] = 1.0 - alphas[i][1] - alphas[i][2] - alphas[i][3];
    }
  }
  PNP_DBG("alphas0 %.10g %.10g %.10g %.10g alphas4 %.10g %.10g %.10g %.10g\n", alphas[0][0], alphas[0][1], alphas[0][2], alphas[0][3], alphas[N-1][0], alphas[N-1][1], alphas[N-1][2], alphas[N-1][3]);
  // M^T M (12 x 12) from the 2N rows [a_j, 0, -a_j u] and [0, a_j, -a_j v]
  for (int i = 0; i < 144; ++i) MtM[i] = 0;
  for (int i = 0; i < N; ++i) {
    double r1[12], r2[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r1[3 * j] = alphas[i][j];
      r1[3 * j + 1] = 0;
      r1[3 * j + 2] = alphas[i][j] * (0.0 - us[i][0]);
      r2[3 * j] = 0;
      r2[3 * j + 1] = alphas[i][j];
      r2[3 * j + 2] = alphas[i][j] * (0.0 - us[i][1]);
    }
    for (int a = 0; a < 12; ++a)
      for (int b = a; b < 12; ++b) MtM[a * 12 + b] += r1[a] * r1[b] + r2[a] * r2[b];
  }
  for (int a = 0; a < 12; ++a)
    for (int b = 0; b < a; ++b) MtM[a * 12 + b] = MtM[b * 12 + a];
  PNP_DBG("MtM00 %.10g MtM[5][7] %.10g trace-ish %.10g\n", MtM[0], MtM[5*12+7], MtM[0]+MtM[13]+MtM[26]+MtM[143]);
}

// piece 2: after the eigen-decomposition (MtM diagonal = eigenvalues, V columns = eigenvectors): the four eigenvectors of
// the smallest eigenvalues, the 6 x 10 distance-constraint matrix L and rho
MVO_HD void epnp_basis(const double* MtM, const double* V, const double (*cws)[3], double (*v)[12], double* L, double* rho) {
  // the four eigenvectors of the smallest eigenvalues, smallest first
  int sel[4];
  {
    bool used[12];
    for (int i = 0; i < 12; ++i) used[i] = false;
    for (int q = 0; q < 4; ++q) {
      int best = -1;
      for (int i = 0; i < 12; ++i)
        if (!used[i] && (best < 0 || MtM[i * 13] < MtM[best * 13])) best = i;
      used[best] = true;
      sel[q] = best;
    }
  }
  for (int q = 0; q < 4; ++q)
    for (int k = 0; k < 12; ++k) v[q][k] = V[k * 12 + sel[q]];
#ifdef MVO_PNP_ROT
  { const double cr = cos(MVO_PNP_ROT), sr = sin(MVO_PNP_ROT);
    for (int k = 0; k < 12; ++k) { const double a = v[0][k], b = v[1][k]; v[0][k] = cr * a + sr * b; v[1][k] = -sr * a + cr * b; } }
#endif
  // L (6 x 10) and rho
  {
    int a = 0, b = 1;
    for (int j = 0; j < 6; ++j) {
      double d[4][3];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) d[i][k] = v[i][3 * a + k] - v[i][3 * b + k];
      auto dot = [&](int p, int q) { return d[p][0] * d[q][0] + d[p][1] * d[q][1] + d[p][2] * d[q][2]; };
      double* row = L + 10 * j;
      row[0] = dot(0, 0);
      row[1] = 2 * dot(0, 1);
      row[2] = dot(1, 1);
      row[3] = 2 * dot(0, 2);
      row[4] = 2 * dot(1, 2);
      row[5] = dot(2, 2);
      row[6] = 2 * dot(0, 3);
      row[7] = 2 * dot(1, 3);
      row[8] = 2 * dot(2, 3);
      row[9] = dot(3, 3);
      double s = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += (cws[a][k] - cws[b][k]) * (cws[a][k] - cws[b][k]);
      rho[j] = s;
      if (++b > 3) {
        ++a;
        b = a + 1;
      }
    }
  }
  PNP_DBG("sel %d %d %d %d L0 %.10g %.10g %.10g rho %.10g %.10g %.10g %.10g %.10g %.10g\n", sel[0], sel[1], sel[2], sel[3], L[0], L[1], L[2], rho[0], rho[1], rho[2], rho[3], rho[4], rho[5]);
}

// piece 3: one of the three beta approximations (mode 0: N = 4 betas, 1: N = 2, 2: N = 3) + five Gauss-Newton steps +
// absolute orientation; returns the mean reprojection error of the candidate (R, t)
template <int N>
MVO_HD double epnp_mode(int mode, const double* L, const double* rho, const double (*v)[12], const double (*alphas)[4],
                        const double (*pw)[3], const double (*us)[2], double* R, double* t) {
  double betas[4] = {0, 0, 0, 0};
#if defined(MVO_EPNP_CLOCK) && defined(__CUDA_ARCH__)
  const long long m0 = clock64();
#endif
  {
    // The three approximations solve 6 x 4, 6 x 3 and 6 x 5 systems (columns {0, 1, 3, 6}, {0, 1, 2}, {0 .. 4} of L).  The
    // three lanes of a warp run them side by side, so all three are written as one 6 x 5 problem with zero columns at
    // the end: a zero column is an exact zero eigenvalue that no rotation touches (a_pq == 0 is skipped) and the
    // back-substitution drops, so x equals the smaller problem's solution bit for bit and the lanes do not diverge.
    double A[30], x[5];
    const int ncols = mode == 0 ? 4 : (mode == 1 ? 3 : 5);
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int col = mode == 0 ? (j == 2 ? 3 : (j == 3 ? 6 : j)) : j;
        A[i * 5 + j] = j < ncols ? L[i * 10 + col] : 0.0;
      }
    // Full-rank systems (the rule) are solved by Householder QR, each lane its own size; cvSolve(CV_SVD) in OpenCV, i.e.
    // the same least-squares solution.  Rank-deficient ones take the minimum-norm path, uniformly.
    bool solved;
    {
      double Aq[30], bq[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) bq[q] = rho[q];
      if (mode == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) Aq[i * 4 + j] = A[i * 5 + j];
        solved = qr_solve<6, 4>(Aq, bq, x);
      } else if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) Aq[i * 3 + j] = A[i * 5 + j];
        solved = qr_solve<6, 3>(Aq, bq, x);
      } else {
#pragma unroll
        for (int q = 0; q < 30; ++q) Aq[q] = A[q];
        solved = qr_solve<6, 5>(Aq, bq, x);
      }
    }
    if (!solved) ls_solve<6, 5>(A, rho, x);
    if (mode == 0) {
      if (x[0] < 0) {
        betas[0] = sqrt(-x[0]);
        betas[1] = -x[1] / betas[0];
        betas[2] = -x[2] / betas[0];
        betas[3] = -x[3] / betas[0];
      } else {
        betas[0] = sqrt(x[0]);
        betas[1] = x[1] / betas[0];
        betas[2] = x[2] / betas[0];
        betas[3] = x[3] / betas[0];
      }
    } else {
      if (x[0] < 0) {
        betas[0] = sqrt(-x[0]);
        betas[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0;
      } else {
        betas[0] = sqrt(x[0]);
        betas[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0;
      }
      if (x[1] < 0) betas[0] = -betas[0];
      if (mode == 2) betas[2] = x[3] / betas[0];
    }
  }
  PNP_DBG("mode %d betas0 %.10g %.10g %.10g %.10g\n", mode, betas[0], betas[1], betas[2], betas[3]);
#if defined(MVO_EPNP_CLOCK) && defined(__CUDA_ARCH__)
  const long long m1 = clock64();
#endif
  // five Gauss-Newton steps on the six distance constraints
  for (int it = 0; it < 5; ++it) {
    double A[24], bb[6], x[4];
    for (int i = 0; i < 6; ++i) {
      const double* l = L + 10 * i;
      const double b0 = betas[0], b1 = betas[1], b2 = betas[2], b3 = betas[3];
      A[i * 4 + 0] = 2 * l[0] * b0 + l[1] * b1 + l[3] * b2 + l[6] * b3;
      A[i * 4 + 1] = l[1] * b0 + 2 * l[2] * b1 + l[4] * b2 + l[7] * b3;
      A[i * 4 + 2] = l[3] * b0 + l[4] * b1 + 2 * l[5] * b2 + l[8] * b3;
      A[i * 4 + 3] = l[6] * b0 + l[7] * b1 + l[8] * b2 + 2 * l[9] * b3;
      bb[i] = rho[i] - (l[0] * b0 * b0 + l[1] * b0 * b1 + l[2] * b1 * b1 + l[3] * b0 * b2 + l[4] * b1 * b2 +
                        l[5] * b2 * b2 + l[6] * b0 * b3 + l[7] * b1 * b3 + l[8] * b2 * b3 + l[9] * b3 * b3);
    }
    {
      double Aq[24], bq[6];
#pragma unroll
      for (int q = 0; q < 24; ++q) Aq[q] = A[q];
#pragma unroll
      for (int q = 0; q < 6; ++q) bq[q] = bb[q];
      if (!qr_solve<6, 4>(Aq, bq, x)) ls_solve<6, 4>(A, bb, x);   // (OpenCV leaves x untouched when its QR meets a zero column)
    }
    PNP_DBG("  gn %d betas %.6g %.6g %.6g %.6g x %.6g %.6g %.6g %.6g bb %.6g %.6g A0 %.6g %.6g %.6g %.6g Lsum %.10g\n", it, betas[0], betas[1], betas[2], betas[3], x[0], x[1], x[2], x[3], bb[0], bb[5], A[0], A[1], A[2], A[3], L[0]+L[11]+L[22]+L[33]+L[44]+L[55]+L[59]);
#pragma unroll
    for (int q = 0; q < 4; ++q) betas[q] += x[q];
  }
  PNP_DBG("mode %d betas %.10g %.10g %.10g %.10g\n", mode, betas[0], betas[1], betas[2], betas[3]);
#if defined(MVO_EPNP_CLOCK) && defined(__CUDA_ARCH__)
  const long long m2 = clock64();
#endif
  // control points in the camera frame, sign, absolute orientation
  double ccs[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) ccs[j][k] = betas[0] * v[0][3 * j + k] + betas[1] * v[1][3 * j + k] + betas[2] * v[2][3 * j + k] + betas[3] * v[3][3 * j + k];
  double sign = 1.0;
  {
    const double z0 = alphas[0][0] * ccs[0][2] + alphas[0][1] * ccs[1][2] + alphas[0][2] * ccs[2][2] + alphas[0][3] * ccs[3][2];
    if (z0 < 0) sign = -1.0;
  }
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  double pcs[N][3];
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pcs[i][k] = sign * (alphas[i][0] * ccs[0][k] + alphas[i][1] * ccs[1][k] + alphas[i][2] * ccs[2][k] + alphas[i][3] * ccs[3][k]);
      pc0[k] += pcs[i][k];
      pw0[k] += pw[i][k];
    }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    pc0[k] /= N;
    pw0[k] /= N;
  }
  double ABt[9];
  for (int q = 0; q < 9; ++q) ABt[q] = 0;
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) ABt[3 * j + k] += (pcs[i][j] - pc0[j]) * (pw[i][k] - pw0[k]);
#if defined(MVO_EPNP_CLOCK) && defined(__CUDA_ARCH__)
  const long long m3 = clock64();
#endif
  polar_rotation<true>(ABt, R);
#if defined(MVO_EPNP_CLOCK) && defined(__CUDA_ARCH__)
  if (threadIdx.x == 0 && R[0] != 123.0) printf("mode cycles: init %lld gn %lld cc %lld polar %lld\n", m1 - m0, m2 - m1, m3 - m2, clock64() - m3);
#endif
  if (det3(R) < 0) {
    R[6] = -R[6];
    R[7] = -R[7];
    R[8] = -R[8];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = pc0[k] - (R[3 * k] * pw0[0] + R[3 * k + 1] * pw0[1] + R[3 * k + 2] * pw0[2]);
  double err = 0;
  for (int i = 0; i < N; ++i) {
    const double X = R[0] * pw[i][0] + R[1] * pw[i][1] + R[2] * pw[i][2] + t[0];
    const double Y = R[3] * pw[i][0] + R[4] * pw[i][1] + R[5] * pw[i][2] + t[1];
    const double iz = 1.0 / (R[6] * pw[i][0] + R[7] * pw[i][1] + R[8] * pw[i][2] + t[2]);
    const double du = us[i][0] - X * iz, dv = us[i][1] - Y * iz;
    err += sqrt(du * du + dv * dv);
  }
  err /= N;
  PNP_DBG("mode %d err %.10g t %.10g %.10g %.10g R0 %.10g %.10g %.10g\n", mode, err, t[0], t[1], t[2], R[0], R[1], R[2]);
  return err;
}

template <int N>
MVO_HD bool epnp_finite(const double* Rout, const double* tout) {
  bool fin = true;
#pragma unroll
  for (int q = 0; q < 9; ++q) fin = fin && isfinite(Rout[q]);
#pragma unroll
  for (int q = 0; q < 3; ++q) fin = fin && isfinite(tout[q]);
  return fin;
}

// Returns false when the solution is not finite.
template <int N>
MVO_HD bool epnp_solve(const double (*pw)[3], const double (*us)[2], double* Rout, double* tout) {
  double cws[4][3], alphas[N][4], MtM[144], V[144];
  epnp_build<N>(pw, us, cws, alphas, MtM);
  jacobi_eig<12>(MtM, V);
  double v[4][12], L[60], rho[6];
  epnp_basis(MtM, V, cws, v, L, rho);
  double best_err = 0;
  for (int mode = 0; mode < 3; ++mode) {
    double R[9], t[3];
    const double err = epnp_mode<N>(mode, L, rho, v, alphas, pw, us, R, t);
    // OpenCV keeps candidate 1 unless a later one is strictly better (NaN errors never win)
    if (mode == 0 || err < best_err) {
      best_err = err;
#pragma unroll
      for (int q = 0; q < 9; ++q) Rout[q] = R[q];
#pragma unroll
      for (int q = 0; q < 3; ++q) tout[q] = t[q];
    }
  }
  return epnp_finite<N>(Rout, tout);
}

#ifdef __CUDACC__
// The form the kernel runs: one warp per problem.  Every lane builds the (tiny) problem redundantly; the 12 x 12
// eigen-decomposition -- nine tenths of a single thread's time -- runs warp-cooperatively on shared memory
// (jacobi_eig_warp: same rotation order and expressions as jacobi_eig); the three beta approximations run on lanes
// 0, 1, 2 side by side and the winner is picked with OpenCV's rule.  sA / sV: 144 doubles of shared memory each.
// The result is returned in every lane.
// IMPL 0: jacobi_eig_warp (round 1, kept as the cross-check); 1: jacobi_eig_warp3 (one element pair per lane, no
// divergence, short scalar chain for the rotation parameters).
template <int N, int IMPL>
__device__ bool epnp_solve_warp(const double (*pw)[3], const double (*us)[2], double* sA, double* sV, int lane, double* Rout,
                                double* tout) {
#ifdef MVO_EPNP_CLOCK
  const long long c0 = clock64();
#endif
  double cws[4][3], alphas[N][4];
  {
    double MtM[144];
    epnp_build<N>(pw, us, cws, alphas, MtM);
    for (int q = lane; q < 144; q += 32) sA[q] = MtM[q];   // identical in every lane
  }
  __syncwarp();
#ifdef MVO_EPNP_CLOCK
  const long long c1 = clock64();
#endif
  if (IMPL == 0) jacobi_eig_warp<12>(sA, sV, lane);
  else jacobi_eig_warp3<12, true>(sA, sV, lane);
  __syncwarp();
#ifdef MVO_EPNP_CLOCK
  const long long c2 = clock64();
#endif
  double v[4][12], L[60], rho[6];
  epnp_basis(sA, sV, cws, v, L, rho);
#ifdef MVO_EPNP_CLOCK
  const long long c3 = clock64();
#endif
  const int mode = lane < 3 ? lane : 0;
  const double err = epnp_mode<N>(mode, L, rho, v, alphas, pw, us, Rout, tout);
#ifdef MVO_EPNP_CLOCK
  if (lane == 0) printf("cycles: build %lld jacobi %lld basis %lld modes %lld\n", c1 - c0, c2 - c1, c3 - c2, clock64() - c3);
#endif
  // OpenCV keeps candidate 0 unless a later one is strictly better (NaN errors never win)
  const double e0 = __shfl_sync(0xffffffffu, err, 0), e1 = __shfl_sync(0xffffffffu, err, 1), e2 = __shfl_sync(0xffffffffu, err, 2);
  int win = 0;
  double best = e0;
  if (e1 < best) {
    best = e1;
    win = 1;
  }
  if (e2 < best) win = 2;
#pragma unroll
  for (int q = 0; q < 9; ++q) Rout[q] = __shfl_sync(0xffffffffu, Rout[q], win);
#pragma unroll
  for (int q = 0; q < 3; ++q) tout[q] = __shfl_sync(0xffffffffu, tout[q], win);
  __syncwarp();   // sA / sV may be reused by the caller
  return epnp_finite<N>(Rout, tout);
}
#endif

}  // namespace mvo
