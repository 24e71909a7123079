// linalg.cuh -- small dense FP64 helpers used per thread by the RANSAC / pose kernels.
#pragma once
#ifndef __device__
#include <cuda_runtime.h>
#endif
#include <math.h>

#ifndef MVO_HD
#define MVO_HD __host__ __device__
#endif

namespace mvo {

// Null space of an R x C matrix (R < C, full row rank assumed) by Gauss-Jordan elimination with complete
// pivoting.  A is destroyed.  basis: (C-R) vectors of length C, orthonormalised (modified Gram-Schmidt).
template <int R, int C>
MVO_HD void null_space(double* A /* R*C row-major */, double* basis /* (C-R)*C */) {
  int colperm[C];
#pragma unroll 1
  for (int j = 0; j < C; ++j) colperm[j] = j;
#pragma unroll 1
  for (int k = 0; k < R; ++k) {
    // complete pivot in the trailing block
    int pr = k, pc = k;
    double best = -1.0;
#pragma unroll 1
    for (int i = k; i < R; ++i)
#pragma unroll 1
      for (int j = k; j < C; ++j) {
        const double v = fabs(A[i * C + j]);
        if (v > best) {
          best = v;
          pr = i;
          pc = j;
        }
      }
    if (pr != k)
#pragma unroll 1
      for (int j = 0; j < C; ++j) {
        const double t = A[k * C + j];
        A[k * C + j] = A[pr * C + j];
        A[pr * C + j] = t;
      }
    if (pc != k) {
#pragma unroll 1
      for (int i = 0; i < R; ++i) {
        const double t = A[i * C + k];
        A[i * C + k] = A[i * C + pc];
        A[i * C + pc] = t;
      }
      const int t = colperm[k];
      colperm[k] = colperm[pc];
      colperm[pc] = t;
    }
    const double piv = A[k * C + k];
    const double inv = (piv != 0.0) ? 1.0 / piv : 0.0;
#pragma unroll 1
    for (int j = k; j < C; ++j) A[k * C + j] *= inv;
#pragma unroll 1
    for (int i = 0; i < R; ++i) {
      if (i == k) continue;
      const double f = A[i * C + k];
      if (f != 0.0)
#pragma unroll 1
        for (int j = k; j < C; ++j) A[i * C + j] -= f * A[k * C + j];
    }
  }
  // A = [I | N] in permuted columns; null vectors: x_pivot = -N[:, f], x_free(f) = 1
#pragma unroll 1
  for (int f = 0; f < C - R; ++f) {
    double* v = basis + f * C;
#pragma unroll 1
    for (int j = 0; j < C; ++j) v[j] = 0.0;
#pragma unroll 1
    for (int i = 0; i < R; ++i) v[colperm[i]] = -A[i * C + R + f];
    v[colperm[R + f]] = 1.0;
  }
  // modified Gram-Schmidt
#pragma unroll 1
  for (int f = 0; f < C - R; ++f) {
    double* v = basis + f * C;
#pragma unroll 1
    for (int g = 0; g < f; ++g) {
      const double* u = basis + g * C;
      double d = 0.0;
#pragma unroll 1
      for (int j = 0; j < C; ++j) d += u[j] * v[j];
#pragma unroll 1
      for (int j = 0; j < C; ++j) v[j] -= d * u[j];
    }
    double nrm = 0.0;
#pragma unroll 1
    for (int j = 0; j < C; ++j) nrm += v[j] * v[j];
    nrm = (nrm > 0.0) ? 1.0 / sqrt(nrm) : 0.0;
#pragma unroll 1
    for (int j = 0; j < C; ++j) v[j] *= nrm;
  }
}

// Solve the N x N system A x = b in place (partial pivoting).  Returns false when singular.
template <int N>
MVO_HD bool lu_solve(double* A /* N*N, destroyed */, double* b /* in: rhs, out: x */) {
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(A[k * N + k]);
#pragma unroll 1
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(A[i * N + k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (best == 0.0) return false;
    if (p != k) {
#pragma unroll 1
      for (int j = 0; j < N; ++j) {
        const double t = A[k * N + j];
        A[k * N + j] = A[p * N + j];
        A[p * N + j] = t;
      }
      const double t = b[k];
      b[k] = b[p];
      b[p] = t;
    }
    const double inv = 1.0 / A[k * N + k];
#pragma unroll 1
    for (int i = k + 1; i < N; ++i) {
      const double f = A[i * N + k] * inv;
      if (f != 0.0) {
#pragma unroll 1
        for (int j = k + 1; j < N; ++j) A[i * N + j] -= f * A[k * N + j];
        b[i] -= f * b[k];
      }
    }
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i];
#pragma unroll 1
    for (int j = i + 1; j < N; ++j) s -= A[i * N + j] * b[j];
    b[i] = s / A[i * N + i];
  }
  return true;
}

// Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite N x N matrix by inverse
// iteration on (A + delta I): one LU factorisation with partial pivoting + a few triangular solves.  The
// smallest eigenvalue of the matrices used here (DLT normal matrices) is separated from the next by many
// orders of magnitude, so three iterations reach machine precision.  A is destroyed.
template <int N>
MVO_HD void smallest_eigvec_spd(double* A, double* x) {
  double tr = 0.0;
#pragma unroll 1
  for (int i = 0; i < N; ++i) tr += A[i * N + i];
  const double delta = 1e-15 * tr + 1e-300;
#pragma unroll 1
  for (int i = 0; i < N; ++i) A[i * N + i] += delta;
  int piv[N];
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(A[k * N + k]);
#pragma unroll 1
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(A[i * N + k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    piv[k] = p;
    if (p != k)
#pragma unroll 1
      for (int j = 0; j < N; ++j) {
        const double t = A[k * N + j];
        A[k * N + j] = A[p * N + j];
        A[p * N + j] = t;
      }
    const double d = A[k * N + k];
    const double inv = d != 0.0 ? 1.0 / d : 0.0;
#pragma unroll 1
    for (int i = k + 1; i < N; ++i) {
      const double f = A[i * N + k] * inv;
      A[i * N + k] = f;
#pragma unroll 1
      for (int j = k + 1; j < N; ++j) A[i * N + j] -= f * A[k * N + j];
    }
  }
#pragma unroll 1
  for (int i = 0; i < N; ++i) x[i] = 1.0 / sqrt((double)N) * ((i & 1) ? 0.9 : 1.1);
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {   // apply the row permutation, forward substitution (unit lower)
      const double t = x[k];
      x[k] = x[piv[k]];
      x[piv[k]] = t;
#pragma unroll 1
      for (int j = 0; j < k; ++j) x[k] -= A[k * N + j] * x[j];
    }
#pragma unroll 1
    for (int i = N - 1; i >= 0; --i) {
      double sacc = x[i];
#pragma unroll 1
      for (int j = i + 1; j < N; ++j) sacc -= A[i * N + j] * x[j];
      const double d = A[i * N + i];
      x[i] = d != 0.0 ? sacc / d : sacc;
    }
    double nrm = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; ++i) nrm += x[i] * x[i];
    nrm = nrm > 0.0 ? 1.0 / sqrt(nrm) : 0.0;
#pragma unroll 1
    for (int i = 0; i < N; ++i) x[i] *= nrm;
  }
}

// Rotation parameters of one Jacobi step from (a_pp, a_qq, a_pq).  FAST = false: the expressions of jacobi_eig (theta, t,
// c, s: three divisions and two square roots in a row).  FAST = true: the same rotation without theta and t.  With
// d = a_qq - a_pp, x = d^2 + 4 a_pq^2 and g = 1 / sqrt(x):  c^2 = (1 + |d| g) / 2,  s = sgn(theta) a_pq g / c -- two
// reciprocal square roots and a handful of multiply-adds on the critical path of every rotation; c and s agree with the
// slow form to a few ulp.
template <bool FAST>
MVO_HD __forceinline__ void jacobi_cs(double app, double aqq, double apq, double& c, double& s) {
  if (FAST) {
    const double d = aqq - app;
#ifdef __CUDA_ARCH__
    const double g = rsqrt(fma(d, d, 4.0 * apq * apq));
    const double z = fma(0.5 * fabs(d), g, 0.5);
    const double q = rsqrt(z);                       // 1 / c
#else
    const double g = 1.0 / sqrt(d * d + 4.0 * apq * apq);
    const double z = 0.5 * fabs(d) * g + 0.5;
    const double q = 1.0 / sqrt(z);
#endif
    const double ag = (d >= 0.0 ? apq : -apq) * g;   // sgn(d) a_pq g = sgn(theta) |a_pq| g   (d = 0: theta = +-0 counts as >= 0 ...
    c = z * q;
    s = (d == 0.0 ? fabs(apq) * g : ag) * q;         // ... so t = +1 there whatever the sign of a_pq)
  } else {
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    c = 1.0 / sqrt(t * t + 1.0);
    s = t * c;
  }
}

// Cyclic Jacobi eigen-decomposition of a symmetric N x N matrix.  A is destroyed (diagonal = eigenvalues),
// V (row-major, columns = eigenvectors).
template <int N>
MVO_HD void jacobi_eig(double* A, double* V) {
#pragma unroll 1
  for (int i = 0; i < N; ++i)
#pragma unroll 1
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
#pragma unroll 1
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll 1
    for (int p = 0; p < N - 1; ++p)
#pragma unroll 1
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
}

#ifdef __CUDACC__
// The same cyclic Jacobi run by one warp on matrices in shared memory: identical rotation order and per-element
// expressions (so the result matches jacobi_eig<N> to rounding of the compiler's FMA contraction), but the O(N) row /
// column updates of a rotation are spread over the lanes -- lanes 0..N-1 update A, lanes 16..16+N-1 the eigenvectors.
// The rotation parameters and the convergence sums are computed redundantly by every lane (broadcast reads).
// One thread needs ~1 ms for a 12 x 12 problem (EPnP, DLT); the warp ~0.1 ms.
template <int N>
__device__ void jacobi_eig_warp(double* A, double* V, int lane) {
  static_assert(N <= 16, "lanes 0..15 update A, lanes 16..31 update V");
  for (int i = lane; i < N * N; i += 32) V[i] = (i / N == i % N) ? 1.0 : 0.0;
  __syncwarp();
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
#pragma unroll 1
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll 1
    for (int p = 0; p < N - 1; ++p)
#pragma unroll 1
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;   // warp-uniform
        const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        __syncwarp();   // every lane has read A[p][q], A[p][p], A[q][q]
        if (lane < N) {
          const int k = lane;
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        } else if (lane >= 16 && lane < 16 + N) {
          const int k = lane - 16;
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
        __syncwarp();
        if (lane < N) {
          const int k = lane;
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        __syncwarp();
      }
  }
}
#endif

#ifdef __CUDACC__
// jacobi_eig_warp without divergence and with one barrier less per rotation.  A rotation (p, q) is 32 independent plane
// rotations of element pairs plus the 2 x 2 block: the column update A <- A J touches (A[k][p], A[k][q]), the row update
// A <- J^T A touches (A[p][k], A[q][k]) -- outside the block (k != p, q) they do not see each other's results -- and V
// <- V J touches (V[k][p], V[k][q]).  That is 10 + 10 + 12 pairs: one per lane, every lane running the same
// instructions on its own two addresses (jacobi_eig_warp ran three divergent branches one after the other: 520 of its
// 890 cycles per rotation).  The four entries of the block are computed redundantly by every lane from the broadcast
// values, column update first, and stored by lane 0.  Same rotation order and element-wise expressions as jacobi_eig.
// (A form with the rows in registers and shuffles instead of shared memory needs the (p, q) loops unrolled -- 66 bodies,
// 38 K instructions against a 32 KB instruction cache -- and ran at half the speed of the shared-memory form.)
template <int N, bool FAST_CS, bool SKIP_TINY = FAST_CS>
__device__ void jacobi_eig_warp3(double* A, double* V, int lane) {
  static_assert(3 * N - 4 == 32, "one pair per lane: 2 (N - 2) pairs of A and N of V");
  for (int i = lane; i < N * N; i += 32) V[i] = (i / N == i % N) ? 1.0 : 0.0;
  __syncwarp();
  const int role = lane < N - 2 ? 0 : (lane < 2 * (N - 2) ? 1 : 2);      // column pair, row pair, eigenvector pair
  const int l = lane - (role == 0 ? 0 : (role == 1 ? N - 2 : 2 * (N - 2)));
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
#pragma unroll 1
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll 1
    for (int p = 0; p < N - 1; ++p)
#pragma unroll 1
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;   // warp-uniform
        const double app = A[p * N + p], aqq = A[q * N + q], aqp = A[q * N + p];
        if (SKIP_TINY && sweep >= 3) {
          // an off-diagonal entry below the rounding level of both diagonal entries: the rotation would change neither
          // eigenvalue and turn the eigenvectors by < 1e-15 -- most of the last sweeps (which only confirm convergence)
          const double gq = 128.0 * fabs(apq);
          if (fabs(app) + gq == fabs(app) && fabs(aqq) + gq == fabs(aqq)) continue;   // warp-uniform
        }
        // this lane's pair (loaded ahead of the scalar chain: its latency hides behind the two rsqrt)
        int k = l;
        if (role != 2) {            // the l-th index that is neither p nor q
          k += (k >= p);
          k += (k >= q);
        }
        double* M = role == 2 ? V : A;
        const int ia = role == 1 ? p * N + k : k * N + p, ib = role == 1 ? q * N + k : k * N + q;
        const double x = M[ia], y = M[ib];
        double c, s;
        jacobi_cs<FAST_CS>(app, aqq, apq, c, s);
        __syncwarp();   // every lane has read the block
        M[ia] = c * x - s * y;
        M[ib] = s * x + c * y;
        // the block: rows p and q of A J, then J^T on them
        const double ppc = c * app - s * apq, pqc = s * app + c * apq;
        const double qpc = c * aqp - s * aqq, qqc = s * aqp + c * aqq;
        if (lane == 0) {
          A[p * N + p] = c * ppc - s * qpc;
          A[q * N + p] = s * ppc + c * qpc;
          A[p * N + q] = c * pqc - s * qqc;
          A[q * N + q] = s * pqc + c * qqc;
        }
        __syncwarp();
      }
  }
}
#endif

// The same cyclic Jacobi, fully unrolled for small N: every index is a compile-time constant after unrolling, so A and
// V live in registers instead of local memory (the per-point 4 x 4 triangulation problems run this 256 K times a step).
template <int N, bool FAST_CS = false>
MVO_HD __forceinline__ void jacobi_eig_reg(double* A, double* V) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < N - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq != 0.0) {
          double c, s;
          jacobi_cs<FAST_CS>(A[p * N + p], A[q * N + q], apq, c, s);
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double akp = A[k * N + p], akq = A[k * N + q];
            A[k * N + p] = c * akp - s * akq;
            A[k * N + q] = s * akp + c * akq;
          }
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double apk = A[p * N + k], aqk = A[q * N + k];
            A[p * N + k] = c * apk - s * aqk;
            A[q * N + k] = s * apk + c * aqk;
          }
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double vkp = V[k * N + p], vkq = V[k * N + q];
            V[k * N + p] = c * vkp - s * vkq;
            V[k * N + q] = s * vkp + c * vkq;
          }
        }
      }
  }
}

MVO_HD __forceinline__ double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[6] * m[5]) + m[2] * (m[3] * m[7] - m[6] * m[4]);
}

MVO_HD __forceinline__ void mat3_mul(const double* a, const double* b, double* c) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

// Real roots of a0 x^3 + a1 x^2 + a2 x + a3, following cv::solveCubic's case analysis.  Returns the count.
MVO_HD inline int solve_cubic(double a0, double a1, double a2, double a3, double* x) {
  int n = 0;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) return 0;
      x[0] = -a3 / a2;
      return 1;
    }
    double d = a2 * a2 - 4 * a1 * a3;
    if (d >= 0) {
      d = sqrt(d);
      const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
      if (fabs(q1) > fabs(q2)) {
        x[0] = q1 / a1;
        x[1] = a3 / q1;
      } else {
        x[0] = q2 / a1;
        x[1] = a3 / q2;
      }
      n = d > 0 ? 2 : 1;
    }
    return n;
  }
  a0 = 1. / a0;
  a1 *= a0;
  a2 *= a0;
  a3 *= a0;
  const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
  const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
  const double Qcubed = Q * Q * Q;
  double d = (a1 * a1 * (a2 * a2 - 4 * a1 * a3) + 2 * a2 * (9 * a1 * a3 - 2 * a2 * a2) - 27 * a3 * a3) * (1. / 108);
  const double kPi = 3.14159265358979323846;
  if (d > 0) {
    const double theta = acos(R / sqrt(Qcubed));
    const double sqrtQ = sqrt(Q);
    const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
    x[0] = t0 * cos(t1) - t2;
    x[1] = t0 * cos(t1 + (2. * kPi / 3)) - t2;
    x[2] = t0 * cos(t1 + (4. * kPi / 3)) - t2;
    n = 3;
  } else if (d == 0) {
    if (R >= 0) {
      x[0] = -2 * pow(R, 1. / 3) - a1 / 3;
      x[1] = pow(R, 1. / 3) - a1 / 3;
    } else {
      x[0] = 2 * pow(-R, 1. / 3) - a1 / 3;
      x[1] = -pow(-R, 1. / 3) - a1 / 3;
    }
    n = x[0] == x[1] ? 1 : 2;
  } else {
    d = sqrt(-d);
    double e = pow(d + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    n = 1;
  }
  return n;
}

// All real roots of a polynomial of degree <= DEG (coefficients c[0] + c[1] x + ... + c[DEG] x^DEG) by
// recursive bracketing: the real roots of p lie between consecutive real roots of p'.  Each bracket is
// solved by safeguarded Newton (bisection fallback).  Returns the count (ascending order).  Roots of even
// multiplicity (tangencies) are not reported.
template <int DEG>
MVO_HD int real_roots(const double* c, double* roots) {
  int deg = DEG;
  while (deg > 0 && c[deg] == 0.0) --deg;
  if (deg == 0) return 0;
  double bound = 0.0;  // Cauchy bound on |root|; by Gauss-Lucas it also bounds the roots of all derivatives
#pragma unroll 1
  for (int i = 0; i < deg; ++i) bound = fmax(bound, fabs(c[i] / c[deg]));
  bound = 1.0 + bound;
  double crit[DEG + 1], next[DEG + 1], q[DEG + 1];
  int ncrit = 0;
#pragma unroll 1
  for (int lvl = deg - 1; lvl >= 0; --lvl) {
    const int qd = deg - lvl;  // degree of the lvl-th derivative
#pragma unroll 1
    for (int i = 0; i <= qd; ++i) {
      double w = c[i + lvl];
#pragma unroll 1
      for (int t = 0; t < lvl; ++t) w *= (double)(i + lvl - t);
      q[i] = w;
    }
    auto evalq = [&](double x, double& dv) {
      double v = q[qd];
      dv = 0.0;
#pragma unroll 1
      for (int i = qd - 1; i >= 0; --i) {
        dv = dv * x + v;
        v = v * x + q[i];
      }
      return v;
    };
    int nn = 0;
    double lo = -bound, dummy;
    double flo = evalq(lo, dummy);
#pragma unroll 1
    for (int k = 0; k <= ncrit; ++k) {
      const double hi = (k < ncrit) ? crit[k] : bound;
      const double fhi = evalq(hi, dummy);
      if (flo == 0.0) {
        if (nn == 0 || next[nn - 1] != lo) next[nn++] = lo;
      } else if (fhi != 0.0 && ((flo < 0.0) != (fhi < 0.0))) {
        // safeguarded Newton (rtsafe): bisect whenever Newton leaves the bracket or fails to halve the step
        double xl = (flo < 0.0) ? lo : hi, xh = (flo < 0.0) ? hi : lo;  // f(xl) < 0 < f(xh)
        double x = 0.5 * (lo + hi), dxold = fabs(hi - lo), dx = dxold, dfx;
        double fx = evalq(x, dfx);
#pragma unroll 1
        for (int it = 0; it < 200; ++it) {
          if (((x - xh) * dfx - fx) * ((x - xl) * dfx - fx) > 0.0 || fabs(2.0 * fx) > fabs(dxold * dfx)) {
            dxold = dx;
            dx = 0.5 * (xh - xl);
            const double xn = xl + dx;
            if (xn == x) break;
            x = xn;
          } else {
            dxold = dx;
            dx = fx / dfx;
            const double xn = x - dx;
            if (xn == x) break;
            x = xn;
          }
          if (fabs(dx) <= 2e-16 * fabs(x)) break;
          fx = evalq(x, dfx);
          if (fx == 0.0) break;
          if (fx < 0.0) xl = x; else xh = x;
        }
        next[nn++] = x;
      }
      lo = hi;
      flo = fhi;
    }
    if (flo == 0.0 && (nn == 0 || next[nn - 1] != lo)) next[nn++] = lo;
    ncrit = nn;
#pragma unroll 1
    for (int k = 0; k < nn; ++k) crit[k] = next[k];
  }
#pragma unroll 1
  for (int k = 0; k < ncrit; ++k) roots[k] = crit[k];
  return ncrit;
}

#ifdef __CUDACC__
// ---- warp-cooperative versions (matrix in shared memory) --------------------------------------------------------------
// LU with partial pivoting on an N x N matrix in shared memory by one warp; elementwise the operations of lu_solve<N>.
// STORE_L keeps the multipliers in the lower triangle (smallest_eigvec_spd's factorisation); piv (registers, uniform).
template <int N, bool STORE_L>
__device__ __forceinline__ bool lu_factor_warp(double* A, double* b, int* piv, int lane) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(A[k * N + k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(A[i * N + k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (!STORE_L && best == 0.0) return false;
    piv[k] = p;
    if (p != k) {
      if (lane < N) {
        const double t = A[k * N + lane];
        A[k * N + lane] = A[p * N + lane];
        A[p * N + lane] = t;
      } else if (lane == N && b) {
        const double t = b[k];
        b[k] = b[p];
        b[p] = t;
      }
      __syncwarp();
    }
    const double d = A[k * N + k];
    const double inv = STORE_L ? (d != 0.0 ? 1.0 / d : 0.0) : 1.0 / d;
    constexpr int kCols = N + 1;                   // columns k+1 .. N-1 of the matrix, column N = right-hand side
    const int cols = kCols - (k + 1);
    const int cnt = (N - 1 - k) * cols;
    double fkeep = 0.0;
    for (int e = lane; e < cnt; e += 32) {
      const int i = k + 1 + e / cols, j = k + 1 + e % cols;
      const double f = A[i * N + k] * inv;
      if (j < N) {
        if (STORE_L || f != 0.0) A[i * N + j] -= f * A[k * N + j];
      } else if (b) {
        if (f != 0.0) b[i] -= f * b[k];
      }
      fkeep = f;
    }
    (void)fkeep;
    __syncwarp();
    if (STORE_L) {
      if (lane > k && lane < N) A[lane * N + k] = A[lane * N + k] * inv;
      __syncwarp();
    }
  }
  return true;
}

// solve A x = b (both in shared memory, destroyed; x returned in b): warp LU, then back substitution on lane 0
template <int N>
__device__ __forceinline__ bool lu_solve_warp(double* A, double* b, int lane) {
  int piv[N];
  const bool ok = lu_factor_warp<N, false>(A, b, piv, lane);
  if (!ok) return false;
  if (lane == 0) {
    double x[N];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      double s2 = b[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) s2 -= A[i * N + j] * x[j];
      x[i] = s2 / A[i * N + i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) b[i] = x[i];
  }
  __syncwarp();
  return true;
}

// smallest_eigvec_spd<N> (linalg.cuh) with the factorisation done by the warp; A in shared memory, x (N doubles) too
template <int N, int ITERS = 4>
__device__ __forceinline__ void smallest_eigvec_spd_warp(double* A, double* xs, int lane) {
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) tr += A[i * N + i];
  const double delta = 1e-15 * tr + 1e-300;
  __syncwarp();
  if (lane < N) A[lane * N + lane] += delta;
  __syncwarp();
  int piv[N];
  lu_factor_warp<N, true>(A, nullptr, piv, lane);
  if (lane == 0) {
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 1.0 / sqrt((double)N) * ((i & 1) ? 0.9 : 1.1);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int k = 0; k < N; ++k) {   // row permutation + forward substitution (unit lower triangle)
        // x[k] <-> x[piv[k]] with a run-time piv: select over the unrolled register file
        const int pk = piv[k];
        double xp = x[k];
#pragma unroll
        for (int q = k + 1; q < N; ++q)
          if (q == pk) {
            xp = x[q];
            x[q] = x[k];
          }
        x[k] = xp;
#pragma unroll
        for (int j = 0; j < k; ++j) x[k] -= A[k * N + j] * x[j];
      }
#pragma unroll
      for (int i = N - 1; i >= 0; --i) {
        double sacc = x[i];
#pragma unroll
        for (int j = i + 1; j < N; ++j) sacc -= A[i * N + j] * x[j];
        const double d = A[i * N + i];
        x[i] = d != 0.0 ? sacc / d : sacc;
      }
      double nrm = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) nrm += x[i] * x[i];
      nrm = nrm > 0.0 ? 1.0 / sqrt(nrm) : 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] *= nrm;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) xs[i] = x[i];
  }
  __syncwarp();
}

#endif  // __CUDACC__

}  // namespace mvo
