// pose.cu -- cv::recoverPose and cv::triangulatePoints for B200.
//
// Replaces the calls at /root/reference/src/initializer.cpp:236 (recoverPose) and
// src/initializer.cpp:125 / src/tracker.cpp:149 (triangulatePoints).  Contract: SURVEY.md A.4.
//   pose_decompose_kernel    E -> (R1, R2, +-t): 3x3 SVD through the symmetric eigenproblem of E^T E (FP64).
//   pose_cheirality_kernel   one thread per (correspondence, candidate): linear triangulation = smallest
//                            eigenvector of the 4x4 A^T A (Jacobi, FP64) + the depth / distance tests.
//   pose_select_kernel       first candidate whose count >= all others; writes R, t and the final mask.
//   triangulate_kernel       one thread per correspondence, general 3x4 projections, f32 output.
#include "context.cuh"
#include "linalg.cuh"
#include <algorithm>

#ifndef MVO_TRI_FAST_CS
#define MVO_TRI_FAST_CS true   // rotation parameters of the 4 x 4 Jacobi by the two-rsqrt chain (linalg.cuh: jacobi_cs)
#endif

namespace mvo {

__device__ __forceinline__ void triangulate_one(const double* P0, const double* P1, double x0, double y0, double x1,
                                                double y1, double* X) {
  double A[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    A[k] = x0 * P0[8 + k] - P0[k];
    A[4 + k] = y0 * P0[8 + k] - P0[4 + k];
    A[8 + k] = x1 * P1[8 + k] - P1[k];
    A[12 + k] = y1 * P1[8 + k] - P1[4 + k];
  }
  double S[16], V[16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      const double s = A[i] * A[j] + A[4 + i] * A[4 + j] + A[8 + i] * A[8 + j] + A[12 + i] * A[12 + j];
      S[i * 4 + j] = s;
      S[j * 4 + i] = s;
    }
  jacobi_eig_reg<4, MVO_TRI_FAST_CS>(S, V);
  // eigenvector of the smallest eigenvalue, selected without dynamic indexing (keeps V in registers)
  double lmin = S[0];
#pragma unroll
  for (int i = 0; i < 4; ++i) X[i] = V[i * 4];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    const bool less = S[k * 4 + k] < lmin;
    lmin = less ? S[k * 4 + k] : lmin;
#pragma unroll
    for (int i = 0; i < 4; ++i) X[i] = less ? V[i * 4 + k] : X[i];
  }
}

__global__ void pose_decompose_kernel(const double* __restrict__ Ein, double* __restrict__ cands) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const double* E = Ein + b * 9;
  // V from E^T E, singular values descending
  double S[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i * 3 + j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
  jacobi_eig<3>(S, V);
  int ord[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int c = a + 1; c < 3; ++c)
      if (S[ord[c] * 4] > S[ord[a] * 4]) {
        const int t = ord[a];
        ord[a] = ord[c];
        ord[c] = t;
      }
  double v[3][3], u[3][3];
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) v[k][i] = V[i * 3 + ord[k]];
  // u_k = E v_k / sigma_k for the two non-zero singular values, u_3 = u_1 x u_2
  for (int k = 0; k < 2; ++k) {
    double n = 0;
    for (int i = 0; i < 3; ++i) {
      u[k][i] = E[i * 3] * v[k][0] + E[i * 3 + 1] * v[k][1] + E[i * 3 + 2] * v[k][2];
      n += u[k][i] * u[k][i];
    }
    n = n > 0 ? 1. / sqrt(n) : 0;
    for (int i = 0; i < 3; ++i) u[k][i] *= n;
  }
  {
    // re-orthogonalise u_2 against u_1 (sigma_1 == sigma_2 up to rounding for an essential matrix)
    const double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
    double n = 0;
    for (int i = 0; i < 3; ++i) {
      u[1][i] -= d * u[0][i];
      n += u[1][i] * u[1][i];
    }
    n = n > 0 ? 1. / sqrt(n) : 0;
    for (int i = 0; i < 3; ++i) u[1][i] *= n;
  }
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
  u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
  u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  // U (columns u_k), Vt (rows v_k); enforce det(U) > 0, det(Vt) > 0 as cv::decomposeEssentialMat
  double U[9], Vt[9];
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) {
      U[i * 3 + k] = u[k][i];
      Vt[k * 3 + i] = v[k][i];
    }
  if (det3(U) < 0)
    for (int i = 0; i < 9; ++i) U[i] = -U[i];
  if (det3(Vt) < 0)
    for (int i = 0; i < 9; ++i) Vt[i] = -Vt[i];
  const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
  const double Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double t1[9], R1[9], R2[9];
  mat3_mul(U, W, t1);
  mat3_mul(t1, Vt, R1);
  mat3_mul(U, Wt, t1);
  mat3_mul(t1, Vt, R2);
  const double t[3] = {U[2], U[5], U[8]};
  double* out = cands + b * 48;
  for (int cnd = 0; cnd < 4; ++cnd) {
    const double* R = (cnd & 1) ? R2 : R1;
    const double sg = (cnd & 2) ? -1.0 : 1.0;
    for (int i = 0; i < 9; ++i) out[cnd * 12 + i] = R[i];
    for (int i = 0; i < 3; ++i) out[cnd * 12 + 9 + i] = sg * t[i];
  }
}

__global__ void __launch_bounds__(128)
pose_cheirality_kernel(const double2* __restrict__ q1, const double2* __restrict__ q2, const int32_t* __restrict__ npts,
                       int max_pts, const double* __restrict__ cands, const uint8_t* __restrict__ mask_in, int use_mask,
                       uint8_t* __restrict__ cand_mask, int32_t* __restrict__ cand_good, double dist) {
  const int b = blockIdx.z, cnd = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = npts[b];
  int good = 0;
  if (i < n) {
    const double* Rt = cands + b * 48 + cnd * 12;
    const double P0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const double P1[12] = {Rt[0], Rt[1], Rt[2], Rt[9], Rt[3], Rt[4], Rt[5], Rt[10], Rt[6], Rt[7], Rt[8], Rt[11]};
    const double2 a = q1[(long long)b * max_pts + i], c = q2[(long long)b * max_pts + i];
    double Q[4];
    triangulate_one(P0, P1, a.x, a.y, c.x, c.y, Q);
    bool m = Q[2] * Q[3] > 0;
    const double x = Q[0] / Q[3], y = Q[1] / Q[3], z = Q[2] / Q[3];
    m = m && (z < dist);
    const double z2 = P1[8] * x + P1[9] * y + P1[10] * z + P1[11];
    m = m && (z2 > 0) && (z2 < dist);
    if (use_mask) m = m && (mask_in[(long long)b * max_pts + i] != 0);
    good = m ? 1 : 0;
    cand_mask[((long long)b * 4 + cnd) * max_pts + i] = (uint8_t)good;
  }
  good = warp_sum(good);
  if ((threadIdx.x & 31) == 0 && good) atomicAdd(cand_good + b * 4 + cnd, good);
}

__global__ void __launch_bounds__(256)
pose_select_kernel(const int32_t* __restrict__ npts, int max_pts, const double* __restrict__ cands,
                   const uint8_t* __restrict__ cand_mask, const int32_t* __restrict__ cand_good,
                   double* __restrict__ pose, uint8_t* __restrict__ mask_out, int32_t* __restrict__ result) {
  const int b = blockIdx.x;
  const int g0 = cand_good[b * 4], g1 = cand_good[b * 4 + 1], g2 = cand_good[b * 4 + 2], g3 = cand_good[b * 4 + 3];
  int sel = 3;
  if (g0 >= g1 && g0 >= g2 && g0 >= g3) sel = 0;
  else if (g1 >= g0 && g1 >= g2 && g1 >= g3) sel = 1;
  else if (g2 >= g0 && g2 >= g1 && g2 >= g3) sel = 2;
  const int n = npts[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    mask_out[(long long)b * max_pts + i] = cand_mask[((long long)b * 4 + sel) * max_pts + i] ? 255 : 0;
  if (threadIdx.x < 12) pose[b * 12 + threadIdx.x] = cands[b * 48 + sel * 12 + threadIdx.x];
  if (threadIdx.x == 0) {
    result[b * 8 + 4] = sel == 0 ? g0 : sel == 1 ? g1 : sel == 2 ? g2 : g3;
    result[b * 8 + 5] = sel;
  }
}

__global__ void __launch_bounds__(128)
triangulate_kernel(const double* __restrict__ proj, const float2* __restrict__ p0, const float2* __restrict__ p1,
                   const int32_t* __restrict__ npts, int max_pts, float* __restrict__ X4) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = npts[b];
  if (i >= n) return;
  double P0[12], P1[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    P0[k] = proj[b * 24 + k];
    P1[k] = proj[b * 24 + 12 + k];
  }
  const float2 a = p0[(long long)b * max_pts + i], c = p1[(long long)b * max_pts + i];
  double X[4];
  triangulate_one(P0, P1, (double)a.x, (double)a.y, (double)c.x, (double)c.y, X);
  float* out = X4 + (long long)b * 4 * max_pts;
#pragma unroll
  for (int k = 0; k < 4; ++k) out[(long long)k * max_pts + i] = (float)X[k];
}

int pose_prepare(mvo_ctx* c) {
  RansacBufs& r = c->rs;
  const size_t B = (size_t)c->cfg.batch;
  MVO_CUDA_TRY(c, r.cands.alloc(B * 48));
  MVO_CUDA_TRY(c, r.cand_mask.alloc(B * 4 * (size_t)r.max_pts));
  MVO_CUDA_TRY(c, r.cand_good.alloc(B * 4));
  MVO_CUDA_TRY(c, r.pose.alloc(B * 12));
  MVO_CUDA_TRY(c, r.proj.alloc(B * 24));
  MVO_CUDA_TRY(c, r.X4.alloc(B * 4 * (size_t)r.max_pts));
  return MVO_OK;
}

int pose_recover(mvo_ctx* c, bool use_mask) {
  RansacBufs& r = c->rs;
  const int B = c->cfg.batch;
  MVO_CUDA_TRY(c, cudaMemsetAsync(r.cand_good.p, 0, (size_t)B * 16, c->stream));
  pose_decompose_kernel<<<B, 32, 0, c->stream>>>(r.ln().best_model.p, r.cands.p);
  c->launches++;
  dim3 grid((r.max_pts + 127) / 128, 4, B);
  pose_cheirality_kernel<<<grid, 128, 0, c->stream>>>(r.q1.p, r.q2.p, r.npts.p, r.max_pts, r.cands.p, r.ln().mask.p,
                                                      use_mask ? 1 : 0, r.cand_mask.p, r.cand_good.p, 50.0);
  c->launches++;
  pose_select_kernel<<<B, 256, 0, c->stream>>>(r.npts.p, r.max_pts, r.cands.p, r.cand_mask.p, r.cand_good.p, r.pose.p,
                                               r.ln().mask.p, r.ln().result.p);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

int pose_triangulate(mvo_ctx* c) {
  RansacBufs& r = c->rs;
  dim3 grid((r.max_pts + 127) / 128, c->cfg.batch);
  triangulate_kernel<<<grid, 128, 0, c->stream>>>(r.proj.p, r.p1.p, r.p2.p, r.npts.p, r.max_pts, r.X4.p);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

}  // namespace mvo

using namespace mvo;

static int pose_upload(mvo_ctx* c, const float* p1, const float* p2, int n, const double* K) {
  MVO_REQUIRE_IDLE(c);
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.batch != 1) {
    c->set_error("the single-call geometry API needs a batch==1 context");
    return MVO_ERR_INVALID;
  }
  int rc = ransac_prepare(c, std::max(n, std::max(c->rs.max_pts, 64)), std::max(c->rs.cap_iters, 2000));
  if (rc) return rc;
  rc = pose_prepare(c);
  if (rc) return rc;
  RansacBufs& r = c->rs;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.p1.p, p1, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.p2.p, p2, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.npts.p, &n, 4, cudaMemcpyHostToDevice, c->stream));
  if (K) MVO_CUDA_TRY(c, cudaMemcpyAsync(r.K.p, K, 72, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

extern "C" {

int mvo_recover_pose(mvo_ctx* c, const double* E, const float* p1, const float* p2, int n, const double* K, double* R,
                     double* t, uint8_t* mask_io, int* n_good) {
  if (!c) return MVO_ERR_INVALID;
  if (!E || !p1 || !p2 || !K || !R || !t || n < 1) {
    c->set_error("mvo_recover_pose: bad argument");
    return MVO_ERR_INVALID;
  }
  int rc = pose_upload(c, p1, p2, n, K);
  if (rc) return rc;
  RansacBufs& r = c->rs;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.ln().best_model.p, E, 72, cudaMemcpyHostToDevice, c->stream));
  if (mask_io) MVO_CUDA_TRY(c, cudaMemcpyAsync(r.ln().mask.p, mask_io, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  rc = ransac_normalize(c);
  if (rc) return rc;
  rc = pose_recover(c, mask_io != nullptr);
  if (rc) return rc;
  double pose[12];
  int res[8];
  MVO_CUDA_TRY(c, cudaMemcpyAsync(pose, r.pose.p, 96, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(res, r.ln().result.p, 32, cudaMemcpyDeviceToHost, c->stream));
  if (mask_io) MVO_CUDA_TRY(c, cudaMemcpyAsync(mask_io, r.ln().mask.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 9; ++i) R[i] = pose[i];
  for (int i = 0; i < 3; ++i) t[i] = pose[9 + i];
  if (n_good) *n_good = res[4];
  return MVO_OK;
}

int mvo_triangulate(mvo_ctx* c, const double* P0, const double* P1, const float* p0, const float* p1, int n, float* X4) {
  if (!c) return MVO_ERR_INVALID;
  if (!P0 || !P1 || !p0 || !p1 || !X4 || n < 0) {
    c->set_error("mvo_triangulate: bad argument");
    return MVO_ERR_INVALID;
  }
  if (n == 0) return MVO_OK;
  int rc = pose_upload(c, p0, p1, n, nullptr);
  if (rc) return rc;
  RansacBufs& r = c->rs;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.proj.p, P0, 96, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.proj.p + 12, P1, 96, cudaMemcpyHostToDevice, c->stream));
  rc = pose_triangulate(c);
  if (rc) return rc;
  MVO_CUDA_TRY(c, cudaMemcpy2DAsync(X4, (size_t)n * 4, r.X4.p, (size_t)r.max_pts * 4, (size_t)n * 4, 4,
                                    cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

}  // extern "C"
