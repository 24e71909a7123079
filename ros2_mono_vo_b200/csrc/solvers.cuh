// solvers.cuh -- minimal solvers (4-point H, 7-point F, 5-point E) and error functions, one thread each.
// Restated from the published algorithms OpenCV implements; parity: oracle/ransac_oracle.py.
#pragma once
#include "linalg.cuh"
#include <float.h>

namespace mvo {

// ---- homography from 4 correspondences (per-axis L1-normalised DLT) -------------------------------
// M -> m.  Returns number of models (0 or 1); H row-major, H[8] == 1.
__device__ inline int solve_h4(const float2* M, const float2* m, double* H) {
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
  for (int i = 0; i < 4; ++i) {
    cmx += m[i].x; cmy += m[i].y;
    cMx += M[i].x; cMy += M[i].y;
  }
  cmx /= 4; cmy /= 4; cMx /= 4; cMy /= 4;
  double smx = 0, smy = 0, sMx = 0, sMy = 0;
  for (int i = 0; i < 4; ++i) {
    smx += fabs(m[i].x - cmx); smy += fabs(m[i].y - cmy);
    sMx += fabs(M[i].x - cMx); sMy += fabs(M[i].y - cMy);
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON) return 0;
  smx = 4 / smx; smy = 4 / smy; sMx = 4 / sMx; sMy = 4 / sMy;
  // Four correspondences determine H exactly (up to scale), so the null vector of OpenCV's 8 x 9 DLT system (its 9 x 9
  // Jacobi eigen-decomposition) is the unique projective map P_i -> p_i; it is written down in closed form in the
  // normalised coordinates: with the projective basis (a, b, c | q), P_q = sum lambda_i P_i and p_q = sum mu_i p_i
  // (Cramer: lambda_a = det(P_q, P_b, P_c) ...), H0 ~ sum_i mu_i (prod_{j != i} lambda_j) p_i (P_j x P_k)^T.
  // The basis triple is the best conditioned one (largest |det_src * det_dst|); everything stays in registers
  // (the Gauss-Jordan null space on a local-memory 8 x 9 array was the longest kernel of the H search).
  double X[4], Y[4], x[4], y[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] = (m[i].x - cmx) * smx; y[i] = (m[i].y - cmy) * smy;
    X[i] = (M[i].x - cMx) * sMx; Y[i] = (M[i].y - cMy) * sMy;
  }
  auto det3h = [](double ax, double ay, double bx, double by, double cx, double cy) {
    return ax * (by - cy) - ay * (bx - cx) + (bx * cy - cx * by);
  };
  // determinant of the triple that leaves point t out, source and destination
  double best = -1.0;
  int tq = 3;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int a = t == 0 ? 1 : 0, b = t <= 1 ? 2 : 1, c = t <= 2 ? 3 : 2;
    const double d = fabs(det3h(X[a], Y[a], X[b], Y[b], X[c], Y[c]) * det3h(x[a], y[a], x[b], y[b], x[c], y[c]));
    if (d > best) {
      best = d;
      tq = t;
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t)
    if (tq == t) {   // the left-out point becomes q = 3 (register swap, no dynamic indexing)
      double s;
      s = X[t]; X[t] = X[3]; X[3] = s;
      s = Y[t]; Y[t] = Y[3]; Y[3] = s;
      s = x[t]; x[t] = x[3]; x[3] = s;
      s = y[t]; y[t] = y[3]; y[3] = s;
    }
  const double l0 = det3h(X[3], Y[3], X[1], Y[1], X[2], Y[2]), l1 = det3h(X[0], Y[0], X[3], Y[3], X[2], Y[2]),
               l2 = det3h(X[0], Y[0], X[1], Y[1], X[3], Y[3]);
  const double u0 = det3h(x[3], y[3], x[1], y[1], x[2], y[2]), u1 = det3h(x[0], y[0], x[3], y[3], x[2], y[2]),
               u2 = det3h(x[0], y[0], x[1], y[1], x[3], y[3]);
  const double w[3] = {u0 * (l1 * l2), u1 * (l0 * l2), u2 * (l0 * l1)};
  double h0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const double r[3] = {Y[j] - Y[k], X[k] - X[j], X[j] * Y[k] - X[k] * Y[j]};   // P_j x P_k
    const double pw[3] = {w[i] * x[i], w[i] * y[i], w[i]};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) h0[a * 3 + b] += pw[a] * r[b];
  }
  const double inv_hn[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
  const double hn2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
  double t[9];
  mat3_mul(inv_hn, h0, t);
  mat3_mul(t, hn2, H);
  if (H[8] == 0.0 || !isfinite(H[8])) return 0;
  const double s = 1. / H[8];
  for (int i = 0; i < 9; ++i) H[i] *= s;
  H[8] = 1.0;
  return 1;
}

// ---- fundamental matrix from 7 correspondences ------------------------------------------------------
// x2^T F x1 = 0.  Returns 0..3 models (F[8] == 1 when possible).
__device__ inline int solve_f7(const float2* m1, const float2* m2, double* Fout /* 3 x 9 */) {
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
  for (int i = 0; i < 7; ++i) {
    c1x += m1[i].x; c1y += m1[i].y;
    c2x += m2[i].x; c2y += m2[i].y;
  }
  const double t = 1. / 7;
  c1x *= t; c1y *= t; c2x *= t; c2y *= t;
  double s1 = 0, s2 = 0;
  for (int i = 0; i < 7; ++i) {
    const double dx1 = m1[i].x - c1x, dy1 = m1[i].y - c1y, dx2 = m2[i].x - c2x, dy2 = m2[i].y - c2y;
    s1 += sqrt(dx1 * dx1 + dy1 * dy1);
    s2 += sqrt(dx2 * dx2 + dy2 * dy2);
  }
  s1 *= t; s2 *= t;
  if (s1 < FLT_EPSILON || s2 < FLT_EPSILON) return 0;
  s1 = sqrt(2.) / s1;
  s2 = sqrt(2.) / s2;
  double A[7 * 9];
  for (int i = 0; i < 7; ++i) {
    const double x0 = (m1[i].x - c1x) * s1, y0 = (m1[i].y - c1y) * s1;
    const double x1 = (m2[i].x - c2x) * s2, y1 = (m2[i].y - c2y) * s2;
    double* a = A + i * 9;
    a[0] = x1 * x0; a[1] = x1 * y0; a[2] = x1; a[3] = y1 * x0; a[4] = y1 * y0; a[5] = y1; a[6] = x0; a[7] = y0; a[8] = 1;
  }
  double B[2 * 9];
  null_space<7, 9>(A, B);
  double* f1 = B;
  double* f2 = B + 9;
  for (int i = 0; i < 9; ++i) f1[i] -= f2[i];
  double c[4];
  double t0 = f2[4] * f2[8] - f2[5] * f2[7];
  double t1 = f2[3] * f2[8] - f2[5] * f2[6];
  double t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7];
  t1 = f1[3] * f1[8] - f1[5] * f1[6];
  t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  double r[3];
  const int n = solve_cubic(c[0], c[1], c[2], c[3], r);
  if (n < 1 || n > 3) return 0;
  const double T1[9] = {s1, 0, -s1 * c1x, 0, s1, -s1 * c1y, 0, 0, 1};
  const double T2t[9] = {s2, 0, 0, 0, s2, 0, -s2 * c2x, -s2 * c2y, 1};  // T2^T
  for (int k = 0; k < n; ++k) {
    double lambda = r[k], mu = 1.;
    const double s = f1[8] * r[k] + f2[8];
    double F[9];
    if (fabs(s) > DBL_EPSILON) {
      mu = 1. / s;
      lambda *= mu;
      F[8] = 1.;
    } else {
      F[8] = 0.;
    }
    for (int i = 0; i < 8; ++i) F[i] = f1[i] * lambda + f2[i] * mu;
    double tmp[9];
    mat3_mul(T2t, F, tmp);
    double* out = Fout + k * 9;
    mat3_mul(tmp, T1, out);
    if (fabs(out[8]) > FLT_EPSILON) {
      const double sc = 1. / out[8];
      for (int i = 0; i < 9; ++i) out[i] *= sc;
    }
  }
  return n;
}

// ---- essential matrix from 5 correspondences (Nister) ------------------------------------------------
// linear monomials [x, y, z, 1]; quadratic [xx, xy, xz, x, yy, yz, y, zz, z, 1];
// cubic (Nister's order) [x3, y3, x2y, xy2, x2z, x2, y2z, y2, xyz, xy, xz2, xz, x, yz2, yz, y, z3, z2, z, 1]
__device__ __constant__ unsigned char c_ll2q[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
__device__ __constant__ unsigned char c_ql2c[10][4] = {{0, 2, 4, 5},     {2, 3, 8, 9},     {4, 8, 10, 11},  {5, 9, 11, 12},
                                                      {3, 1, 6, 7},     {8, 6, 13, 14},   {9, 7, 14, 15},  {10, 13, 16, 17},
                                                      {11, 14, 17, 18}, {12, 15, 18, 19}};

__device__ __forceinline__ void pmul_ll(const double* a, const double* b, double* q /* += , 10 */, double s) {
#pragma unroll 1
  for (int i = 0; i < 4; ++i)
#pragma unroll 1
    for (int j = 0; j < 4; ++j) q[c_ll2q[i][j]] += s * a[i] * b[j];
}
__device__ __forceinline__ void pmul_ql(const double* q, const double* l, double* c /* += , 20 */, double s) {
#pragma unroll 1
  for (int i = 0; i < 10; ++i)
#pragma unroll 1
    for (int j = 0; j < 4; ++j) c[c_ql2c[i][j]] += s * q[i] * l[j];
}

// Part 1 of the 5-point solver: null-space basis EE (4 x 9), the z-polynomial matrix Bm (3 x 13, ascending
// coefficients: bx deg 3, by deg 3, b1 deg 4 per row) and det B(z) (degree 10, ascending).  Returns false when
// the 10 x 10 elimination is singular.
__device__ inline bool e5_setup(const double2* q1, const double2* q2, double* EE, double* Bm /* 39 */, double* detp /* 11 */) {
  double A[5 * 9];
  for (int i = 0; i < 5; ++i) {
    const double x1 = q1[i].x, y1 = q1[i].y, x2 = q2[i].x, y2 = q2[i].y;
    double* a = A + i * 9;
    a[0] = x2 * x1; a[1] = x2 * y1; a[2] = x2; a[3] = y2 * x1; a[4] = y2 * y1; a[5] = y2; a[6] = x1; a[7] = y1; a[8] = 1;
  }
  null_space<5, 9>(A, EE);   // E = x*EE0 + y*EE1 + z*EE2 + EE3
  // entries of E as linear polynomials e[rc][4]
  double e[9][4];
  for (int k = 0; k < 9; ++k)
    for (int j = 0; j < 4; ++j) e[k][j] = EE[j * 9 + k];
  // EEt = E E^T (symmetric, quadratic polynomials), idx(r,c) for r<=c
  double eet[6][10];
  const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int r = 0; r < 3; ++r)
    for (int c = r; c < 3; ++c) {
      double* q = eet[sym[r][c]];
      for (int i = 0; i < 10; ++i) q[i] = 0;
      for (int k = 0; k < 3; ++k) pmul_ll(e[r * 3 + k], e[c * 3 + k], q, 1.0);
    }
  double tr[10];
  for (int i = 0; i < 10; ++i) tr[i] = eet[0][i] + eet[3][i] + eet[5][i];
  double M[10 * 20];
  for (int i = 0; i < 200; ++i) M[i] = 0;
  // row 0: det(E)
  {
    double q[10];
    for (int i = 0; i < 10; ++i) q[i] = 0;
    pmul_ll(e[4], e[8], q, 1.0);
    pmul_ll(e[5], e[7], q, -1.0);
    pmul_ql(q, e[0], M, 1.0);
    for (int i = 0; i < 10; ++i) q[i] = 0;
    pmul_ll(e[3], e[8], q, 1.0);
    pmul_ll(e[5], e[6], q, -1.0);
    pmul_ql(q, e[1], M, -1.0);
    for (int i = 0; i < 10; ++i) q[i] = 0;
    pmul_ll(e[3], e[7], q, 1.0);
    pmul_ll(e[4], e[6], q, -1.0);
    pmul_ql(q, e[2], M, 1.0);
  }
  // rows 1..9: 2 E E^T E - tr(E E^T) E
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double* row = M + (1 + r * 3 + c) * 20;
      for (int k = 0; k < 3; ++k) pmul_ql(eet[sym[r][k]], e[k * 3 + c], row, 2.0);
      pmul_ql(tr, e[r * 3 + c], row, -1.0);
    }
  // Gauss-Jordan on the first 10 columns (partial pivoting): M -> [I | G]
  for (int k = 0; k < 10; ++k) {
    int p = k;
    double best = fabs(M[k * 20 + k]);
    for (int i = k + 1; i < 10; ++i) {
      const double v = fabs(M[i * 20 + k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (best == 0.0) return false;
    if (p != k)
      for (int j = k; j < 20; ++j) {
        const double t = M[k * 20 + j];
        M[k * 20 + j] = M[p * 20 + j];
        M[p * 20 + j] = t;
      }
    const double inv = 1.0 / M[k * 20 + k];
    for (int j = k; j < 20; ++j) M[k * 20 + j] *= inv;
    for (int i = 0; i < 10; ++i) {
      if (i == k) continue;
      const double f = M[i * 20 + k];
      if (f != 0.0)
        for (int j = k; j < 20; ++j) M[i * 20 + j] -= f * M[k * 20 + j];
    }
  }
  // B(z): rows <x2z> - z<x2>, <y2z> - z<y2>, <xyz> - z<xy>; tail = [xz2, xz, x, yz2, yz, y, z3, z2, z, 1]
  for (int i = 0; i < 3; ++i) {
    const double* ga = M + (4 + 2 * i) * 20 + 10;
    const double* gb = M + (5 + 2 * i) * 20 + 10;
    double* bx = Bm + i * 13;
    double* by = bx + 4;
    double* b1 = bx + 8;
    bx[0] = ga[2];           bx[1] = ga[1] - gb[2]; bx[2] = ga[0] - gb[1]; bx[3] = -gb[0];
    by[0] = ga[5];           by[1] = ga[4] - gb[5]; by[2] = ga[3] - gb[4]; by[3] = -gb[3];
    b1[0] = ga[9];           b1[1] = ga[8] - gb[9]; b1[2] = ga[7] - gb[8]; b1[3] = ga[6] - gb[7]; b1[4] = -gb[6];
  }
  // det B(z): degree 10
  for (int i = 0; i < 11; ++i) detp[i] = 0;
  auto accum = [&](const double* p0, int d0, const double* p1, int d1, const double* p2, int d2, double s) {
    for (int i = 0; i <= d0; ++i)
      for (int j = 0; j <= d1; ++j) {
        const double v = s * p0[i] * p1[j];
        for (int k = 0; k <= d2; ++k) detp[i + j + k] += v * p2[k];
      }
  };
  const double* B0 = Bm;
  const double* B1 = Bm + 13;
  const double* B2 = Bm + 26;
  accum(B0, 3, B1 + 4, 3, B2 + 8, 4, 1.0);
  accum(B0, 3, B1 + 8, 4, B2 + 4, 3, -1.0);
  accum(B0 + 4, 3, B1, 3, B2 + 8, 4, -1.0);
  accum(B0 + 4, 3, B1 + 8, 4, B2, 3, 1.0);
  accum(B0 + 8, 4, B1, 3, B2 + 4, 3, 1.0);
  accum(B0 + 8, 4, B1 + 4, 3, B2, 3, -1.0);
  return true;
}

// Part 1 as one warp runs it (e5_setup_kernel): the same arithmetic in the same order, with the two expensive pieces
// spread over the lanes -- lane r builds row r of the 10 x 20 constraint matrix, and the Gauss-Jordan elimination works
// on the matrix in shared memory with lane = column.  The small serial pieces (null space, E E^T, B(z), det B) are
// computed redundantly by every lane.  Msh: 200 doubles of shared memory.  The outputs are identical in every lane.
__device__ inline bool e5_setup_warp(const double2* q1, const double2* q2, double* Msh, int lane, double* EE, double* Bm /* 39 */,
                                     double* detp /* 11 */) {
  double A[5 * 9];
  for (int i = 0; i < 5; ++i) {
    const double x1 = q1[i].x, y1 = q1[i].y, x2 = q2[i].x, y2 = q2[i].y;
    double* a = A + i * 9;
    a[0] = x2 * x1; a[1] = x2 * y1; a[2] = x2; a[3] = y2 * x1; a[4] = y2 * y1; a[5] = y2; a[6] = x1; a[7] = y1; a[8] = 1;
  }
  null_space<5, 9>(A, EE);
  double e[9][4];
  for (int k = 0; k < 9; ++k)
    for (int j = 0; j < 4; ++j) e[k][j] = EE[j * 9 + k];
  double eet[6][10];
  const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int r = 0; r < 3; ++r)
    for (int c = r; c < 3; ++c) {
      double* q = eet[sym[r][c]];
      for (int i = 0; i < 10; ++i) q[i] = 0;
      for (int k = 0; k < 3; ++k) pmul_ll(e[r * 3 + k], e[c * 3 + k], q, 1.0);
    }
  double tr[10];
  for (int i = 0; i < 10; ++i) tr[i] = eet[0][i] + eet[3][i] + eet[5][i];
  // lane r builds row r
  if (lane < 10) {
    double row[20];
    for (int i = 0; i < 20; ++i) row[i] = 0;
    if (lane == 0) {
      double q[10];
      for (int i = 0; i < 10; ++i) q[i] = 0;
      pmul_ll(e[4], e[8], q, 1.0);
      pmul_ll(e[5], e[7], q, -1.0);
      pmul_ql(q, e[0], row, 1.0);
      for (int i = 0; i < 10; ++i) q[i] = 0;
      pmul_ll(e[3], e[8], q, 1.0);
      pmul_ll(e[5], e[6], q, -1.0);
      pmul_ql(q, e[1], row, -1.0);
      for (int i = 0; i < 10; ++i) q[i] = 0;
      pmul_ll(e[3], e[7], q, 1.0);
      pmul_ll(e[4], e[6], q, -1.0);
      pmul_ql(q, e[2], row, 1.0);
    } else {
      const int r = (lane - 1) / 3, c = (lane - 1) - r * 3;
      for (int k = 0; k < 3; ++k) pmul_ql(eet[sym[r][k]], e[k * 3 + c], row, 2.0);
      pmul_ql(tr, e[r * 3 + c], row, -1.0);
    }
    for (int j = 0; j < 20; ++j) Msh[lane * 20 + j] = row[j];
  }
  __syncwarp();
  // Gauss-Jordan on the first 10 columns (partial pivoting), lane = column: M -> [I | G]
  bool singular = false;
#pragma unroll 1
  for (int k = 0; k < 10; ++k) {
    int p = k;
    double best = fabs(Msh[k * 20 + k]);
    for (int i = k + 1; i < 10; ++i) {
      const double v = fabs(Msh[i * 20 + k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (best == 0.0) {   // warp-uniform
      singular = true;
      break;
    }
    __syncwarp();
    if (p != k && lane >= k && lane < 20) {
      const double t = Msh[k * 20 + lane];
      Msh[k * 20 + lane] = Msh[p * 20 + lane];
      Msh[p * 20 + lane] = t;
    }
    __syncwarp();
    const double inv = 1.0 / Msh[k * 20 + k];
    __syncwarp();
    if (lane >= k && lane < 20) Msh[k * 20 + lane] *= inv;
    __syncwarp();
    double f[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) f[i] = Msh[i * 20 + k];
    __syncwarp();
    if (lane >= k && lane < 20) {
      const double mk = Msh[k * 20 + lane];
#pragma unroll
      for (int i = 0; i < 10; ++i)
        if (i != k && f[i] != 0.0) Msh[i * 20 + lane] -= f[i] * mk;
    }
    __syncwarp();
  }
  if (singular) return false;
  // B(z): rows <x2z> - z<x2>, <y2z> - z<y2>, <xyz> - z<xy>; tail = [xz2, xz, x, yz2, yz, y, z3, z2, z, 1]
  for (int i = 0; i < 3; ++i) {
    const double* ga = Msh + (4 + 2 * i) * 20 + 10;
    const double* gb = Msh + (5 + 2 * i) * 20 + 10;
    double* bx = Bm + i * 13;
    double* by = bx + 4;
    double* b1 = bx + 8;
    bx[0] = ga[2];           bx[1] = ga[1] - gb[2]; bx[2] = ga[0] - gb[1]; bx[3] = -gb[0];
    by[0] = ga[5];           by[1] = ga[4] - gb[5]; by[2] = ga[3] - gb[4]; by[3] = -gb[3];
    b1[0] = ga[9];           b1[1] = ga[8] - gb[9]; b1[2] = ga[7] - gb[8]; b1[3] = ga[6] - gb[7]; b1[4] = -gb[6];
  }
  // det B(z): degree 10
  for (int i = 0; i < 11; ++i) detp[i] = 0;
  auto accum = [&](const double* p0, int d0, const double* p1, int d1, const double* p2, int d2, double s) {
    for (int i = 0; i <= d0; ++i)
      for (int j = 0; j <= d1; ++j) {
        const double v = s * p0[i] * p1[j];
        for (int k = 0; k <= d2; ++k) detp[i + j + k] += v * p2[k];
      }
  };
  const double* B0 = Bm;
  const double* B1 = Bm + 13;
  const double* B2 = Bm + 26;
  accum(B0, 3, B1 + 4, 3, B2 + 8, 4, 1.0);
  accum(B0, 3, B1 + 8, 4, B2 + 4, 3, -1.0);
  accum(B0 + 4, 3, B1, 3, B2 + 8, 4, -1.0);
  accum(B0 + 4, 3, B1 + 8, 4, B2, 3, 1.0);
  accum(B0 + 8, 4, B1, 3, B2 + 4, 3, 1.0);
  accum(B0 + 8, 4, B1 + 4, 3, B2, 3, -1.0);
  __syncwarp();   // Msh may be reused by the caller
  return true;
}

// Part 2: the essential matrix of one real root z of det B(z).  Returns false for degenerate roots.
__device__ inline bool e5_model_from_root(double z, const double* EE, const double* Bm, double* E /* 9 */) {
  double bz[9];
  for (int i = 0; i < 3; ++i) {
    const double* bx = Bm + i * 13;
    const double* by = bx + 4;
    const double* b1 = bx + 8;
    bz[i * 3 + 0] = ((bx[3] * z + bx[2]) * z + bx[1]) * z + bx[0];
    bz[i * 3 + 1] = ((by[3] * z + by[2]) * z + by[1]) * z + by[0];
    bz[i * 3 + 2] = (((b1[4] * z + b1[3]) * z + b1[2]) * z + b1[1]) * z + b1[0];
  }
  // null vector of the (numerically singular) 3x3: the largest of the three row cross products
  double bestn = -1, xv = 0, yv = 0, wv = 0;
  for (int a = 0; a < 3; ++a) {
    const double* r0 = bz + ((a + 1) % 3) * 3;
    const double* r1 = bz + ((a + 2) % 3) * 3;
    const double cx = r0[1] * r1[2] - r0[2] * r1[1];
    const double cy = r0[2] * r1[0] - r0[0] * r1[2];
    const double cw = r0[0] * r1[1] - r0[1] * r1[0];
    const double nn = cx * cx + cy * cy + cw * cw;
    if (nn > bestn) {
      bestn = nn;
      xv = cx; yv = cy; wv = cw;
    }
  }
  if (!(bestn > 0.0)) return false;
  const double inv = 1.0 / sqrt(bestn);
  if (fabs(wv * inv) < 1e-10) return false;
  const double x = xv / wv, y = yv / wv;
  double nrm = 0;
  for (int i = 0; i < 9; ++i) {
    E[i] = x * EE[i] + y * EE[9 + i] + z * EE[18 + i] + EE[27 + i];
    nrm += E[i] * E[i];
  }
  if (!(nrm > 0.0) || !isfinite(nrm)) return false;
  nrm = 1.0 / sqrt(nrm);
  for (int i = 0; i < 9; ++i) E[i] *= nrm;
  return true;
}

// q1, q2: K-normalised points, constraint q2^T E q1 = 0.  Returns 0..10 models with unit Frobenius norm.
// (sequential driver: used by the host unit test; the CUDA path runs e5_setup per thread and then finds the
// roots with 16 lanes per hypothesis, see ransac.cu)
__device__ inline int solve_e5(const double2* q1, const double2* q2, double* Eout /* 10 x 9 */) {
  double EE[36], Bm[39], detp[11];
  if (!e5_setup(q1, q2, EE, Bm, detp)) return 0;
  double roots[11];
  const int nr = real_roots<10>(detp, roots);
  int count = 0;
  for (int k = 0; k < nr && count < 10; ++k)
    if (e5_model_from_root(roots[k], EE, Bm, Eout + count * 9)) ++count;
  return count;
}

// ---- per-correspondence errors ---------------------------------------------------------------------
// float, non-fused, as HomographyEstimatorCallback::computeError
__device__ __forceinline__ float h_error(const float* Hf, float2 M, float2 m) {
  const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], M.x), __fmul_rn(Hf[7], M.y)), 1.f));
  const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], M.x), __fmul_rn(Hf[1], M.y)), Hf[2]), ww), m.x);
  const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], M.x), __fmul_rn(Hf[4], M.y)), Hf[5]), ww), m.y);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__device__ __forceinline__ float f_error(const double* F, float2 p1, float2 p2) {
  const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
  double a = __dadd_rn(__dadd_rn(__dmul_rn(F[0], x1), __dmul_rn(F[1], y1)), F[2]);
  double b = __dadd_rn(__dadd_rn(__dmul_rn(F[3], x1), __dmul_rn(F[4], y1)), F[5]);
  double c = __dadd_rn(__dadd_rn(__dmul_rn(F[6], x1), __dmul_rn(F[7], y1)), F[8]);
  const double s2 = 1. / __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(x2, a), __dmul_rn(y2, b)), c);
  a = __dadd_rn(__dadd_rn(__dmul_rn(F[0], x2), __dmul_rn(F[3], y2)), F[6]);
  b = __dadd_rn(__dadd_rn(__dmul_rn(F[1], x2), __dmul_rn(F[4], y2)), F[7]);
  c = __dadd_rn(__dadd_rn(__dmul_rn(F[2], x2), __dmul_rn(F[5], y2)), F[8]);
  const double s1 = 1. / __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
  const double d1 = __dadd_rn(__dadd_rn(__dmul_rn(x1, a), __dmul_rn(y1, b)), c);
  return (float)fmax(__dmul_rn(__dmul_rn(d1, d1), s1), __dmul_rn(__dmul_rn(d2, d2), s2));
}

// Sampson error on K-normalised points
__device__ __forceinline__ float e_error(const double* E, double2 x1, double2 x2) {
  const double ex0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x1.x), __dmul_rn(E[1], x1.y)), E[2]);
  const double ex1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], x1.x), __dmul_rn(E[4], x1.y)), E[5]);
  const double ex2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], x1.x), __dmul_rn(E[7], x1.y)), E[8]);
  const double et0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], x2.x), __dmul_rn(E[3], x2.y)), E[6]);
  const double et1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], x2.x), __dmul_rn(E[4], x2.y)), E[7]);
  const double d = __dadd_rn(__dadd_rn(__dmul_rn(x2.x, ex0), __dmul_rn(x2.y, ex1)), ex2);
  const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex0, ex0), __dmul_rn(ex1, ex1)), __dmul_rn(et0, et0)),
                               __dmul_rn(et1, et1));
  return (float)(__dmul_rn(d, d) / den);
}

}  // namespace mvo
