// pnp.cu -- cv::solvePnPRansac as the reference's Tracker calls it, for B200.
//
// Replaces /root/reference/src/tracker.cpp:309
//     cv::solvePnPRansac(points_3d, points_2d, K, d, rvec, tvec, false, 100, 8.0, 0.99, inliers)
// (SURVEY.md 8f "next #1").  Contract: oracle/pnp_oracle.py, pinned to cv2 4.13.0.
//
//   pnp_normalize_kernel   image points -> K-normalised coordinates rounded to float (undistortPoints on float input)
//   pnp_sample_kernel      the registrator's 5-point samples from the OpenCV MWC stream (no subset check): every offset
//                          of the stream simulates the sample that would start there, one thread follows the chain
//   pnp_epnp_kernel        warp per hypothesis (FP64): EPnP -- PCA control points with cv::SVD's signs (one-sided
//                          Jacobi, replayed exactly), 12x12 M^T M eigenvectors by a warp-cooperative cyclic Jacobi (one
//                          element pair per lane), the three beta approximations (QR) + 5 Gauss-Newton steps (QR) +
//                          absolute orientation on three lanes, best reprojection error
//   pnp_score_kernel       CTA per hypothesis: float squared reprojection error <= thr^2, warp-reduced counts
//   pnp_bound_kernel       batched calls: hypotheses the adaptive loop can still reach after the first 32 of a stream
//                          (the second epnp / score launch skips the rest)
//   pnp_select_kernel      replays the sequential adaptive loop (strict '>' update, shrinking niters), writes the
//                          inlier mask and the ordered inlier list of the winner
//   pnp_refine_kernel      solvePnP(inliers, ITERATIVE): DLT initialisation + Levenberg-Marquardt as CvLevMarq runs it
//                          (<= 20 iterations, eps FLT_EPSILON); one CTA per stream, deterministic block reductions
#include "context.cuh"
#include "linalg.cuh"
#include "pnp_math.cuh"
#include <float.h>
#include <algorithm>

namespace mvo {

// ------------------------------------------------------------------------------------------------
__global__ void pnp_normalize_kernel(const float2* __restrict__ img, const int32_t* __restrict__ npts, int max_pts,
                                     const double* __restrict__ K, double2* __restrict__ xn) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts[b]) return;
  const double* k = K + b * 9;
  const float2 p = img[(long long)b * max_pts + i];
  // undistortPoints with zero distortion on float points: (u - cx) * (1 / fx), result stored as float
  const float x = (float)(((double)p.x - k[2]) * (1.0 / k[0]));
  const float y = (float)(((double)p.y - k[5]) * (1.0 / k[4]));
  xn[(long long)b * max_pts + i] = make_double2((double)x, (double)y);
}

// the registrator's getSubset without a subset check: distinct indices, a repeated draw is redrawn.  The walk over the
// RNG stream is sequential -- sample k starts where sample k - 1 stopped -- but where a sample that starts at a given
// offset stops does not depend on anything else, so every offset of a 1024-draw window simulates "the sample that would
// start here" in parallel (thread per offset), one thread then follows the chain offset -> next offset from the current
// position (one shared-memory load per sample instead of ~40 dependent instructions per draw: 52 -> 6 us), and the
// threads the chain visited write their samples out.
constexpr int kPnpSampleThreads = 1024;
__global__ void __launch_bounds__(kPnpSampleThreads)
pnp_sample_kernel(const uint32_t* __restrict__ rng, int rng_len, const int32_t* __restrict__ npts, int iters,
                  int32_t* __restrict__ subsets) {
  __shared__ int s_next[kPnpSampleThreads], s_slot[kPnpSampleThreads];
  __shared__ int s_pos, s_it;
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = npts[b];
  int32_t* out = subsets + (long long)b * iters * kPnpK;
  if (n < kPnpK) {   // no sample of distinct points exists (batched callers pass npts = 0 for streams without enough points)
    for (int k = t; k < iters * kPnpK; k += kPnpSampleThreads) out[k] = 0;
    return;
  }
  if (t == 0) {
    s_pos = 0;
    s_it = 0;
  }
  __syncthreads();
  for (;;) {
    const int base = s_pos, it0 = s_it;
    if (it0 >= iters || base >= rng_len) break;
    int idx[kPnpK];
    int o = base + t, i = 0;
    while (i < kPnpK && o < rng_len) {
      const int v = (int)(rng[o++] % (uint32_t)n);
      bool dup = false;
#pragma unroll
      for (int j = 0; j < kPnpK; ++j) dup = dup || (j < i && idx[j] == v);
      if (dup) continue;
#pragma unroll
      for (int j = 0; j < kPnpK; ++j)
        if (j == i) idx[j] = v;
      ++i;
    }
    s_next[t] = i == kPnpK ? o - base : -1;   // (relative; a sample that starts near the end of the window may stop outside)
    s_slot[t] = -1;
    __syncthreads();
    if (t == 0) {
      int cur = 0, it = it0;
      while (it < iters && cur < kPnpSampleThreads && s_next[cur] >= 0) {
        s_slot[cur] = it++;
        cur = s_next[cur];
      }
      // the stream ran out inside a sample (cannot happen with the table sizes in use): stop, the rest is marked below
      s_pos = (cur < kPnpSampleThreads && it < iters) ? rng_len : base + cur;
      s_it = it;
    }
    __syncthreads();
    const int slot = s_slot[t];
    if (slot >= 0) {
#pragma unroll
      for (int j = 0; j < kPnpK; ++j) out[slot * kPnpK + j] = idx[j];
    }
    __syncthreads();
  }
  for (int k = s_it * kPnpK + t; k < iters * kPnpK; k += kPnpSampleThreads) out[k] = 0;   // unusable samples
}

// One warp per hypothesis (epnp_solve_warp, pnp_math.cuh).
constexpr int kEpnpWarps = 2;
constexpr int kPnpFirstRound = 32;   // hypotheses of the first round of a batched call
#ifndef MVO_EPNP_MINB
#define MVO_EPNP_MINB 6   // 168 registers: 12 instead of 8 resident warps per SM (batched tracking step 4.35 -> 3.85 ms; 8 blocks / 128 registers loses again)
#endif
template <int IMPL>
__global__ void __launch_bounds__(kEpnpWarps * 32, MVO_EPNP_MINB)
pnp_epnp_kernel(const float* __restrict__ obj, const double2* __restrict__ xn, int max_pts,
                const int32_t* __restrict__ subsets, int iters, int it_begin, int it_end, const int32_t* __restrict__ bound,
                double* __restrict__ models, int32_t* __restrict__ ok) {
  __shared__ double s_A[kEpnpWarps][144], s_V[kEpnpWarps][144];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = it_begin + blockIdx.x * kEpnpWarps + warp;
  if (it >= it_end) return;
  if (bound && it >= bound[b]) return;   // second round: the sequential loop of this stream stops before this hypothesis
  double pw[kPnpK][3], us[kPnpK][2];
  const int32_t* s = subsets + ((long long)b * iters + it) * kPnpK;
  for (int i = 0; i < kPnpK; ++i) {
    const long long o = (long long)b * max_pts + s[i];
    pw[i][0] = (double)obj[o * 3];
    pw[i][1] = (double)obj[o * 3 + 1];
    pw[i][2] = (double)obj[o * 3 + 2];
    const double2 u = xn[o];
    us[i][0] = u.x;
    us[i][1] = u.y;
  }
  double R[9], t[3];
  const bool good = epnp_solve_warp<kPnpK, IMPL>(pw, us, s_A[warp], s_V[warp], lane, R, t);
  if (lane == 0) {
    double* m = models + ((long long)b * iters + it) * 12;
    for (int q = 0; q < 9; ++q) m[q] = R[q];
    for (int q = 0; q < 3; ++q) m[9 + q] = t[q];
    ok[(long long)b * iters + it] = good ? 1 : 0;
  }
}

// float squared reprojection error of one point (PnPRansacCallback::computeError)
__device__ __forceinline__ float pnp_err(const double* m, const double* K, float X, float Y, float Z, float2 uv) {
  const double xc = m[0] * X + m[1] * Y + m[2] * Z + m[9];
  const double yc = m[3] * X + m[4] * Y + m[5] * Z + m[10];
  const double zc = m[6] * X + m[7] * Y + m[8] * Z + m[11];
  const double iz = zc != 0 ? 1.0 / zc : 1.0;
  const float pu = (float)(xc * iz * K[0] + K[2]), pv = (float)(yc * iz * K[4] + K[5]);
  const float dx = __fsub_rn(uv.x, pu), dy = __fsub_rn(uv.y, pv);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__global__ void __launch_bounds__(256)
pnp_score_kernel(const float* __restrict__ obj, const float2* __restrict__ img, const int32_t* __restrict__ npts,
                 int max_pts, const double* __restrict__ K, const double* __restrict__ models,
                 const int32_t* __restrict__ ok, int iters, int it_begin, const int32_t* __restrict__ bound, float thr2,
                 int32_t* __restrict__ counts) {
  const int b = blockIdx.y, it = it_begin + blockIdx.x;
  if (bound && it >= bound[b]) return;
  const long long h = (long long)b * iters + it;
  if (!ok[h]) {
    if (threadIdx.x == 0) counts[h] = -1;
    return;
  }
  __shared__ double m[12], k[9];
  __shared__ int total;
  if (threadIdx.x < 12) m[threadIdx.x] = models[h * 12 + threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 41) k[threadIdx.x - 32] = K[b * 9 + threadIdx.x - 32];
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const int n = npts[b];
  int c = 0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const long long o = (long long)b * max_pts + i;
    c += pnp_err(m, k, obj[o * 3], obj[o * 3 + 1], obj[o * 3 + 2], img[o]) <= thr2 ? 1 : 0;
  }
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
  __syncthreads();
  if (threadIdx.x == 0) counts[h] = total;
}

// RANSACUpdateNumIters
__device__ int pnp_update_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.);
  p = fmin(p, 1.);
  ep = fmax(ep, 0.);
  ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

// Hypotheses the sequential loop can still reach after the first `first` ones (the loop of pnp_select_kernel, stopped
// there): RANSACUpdateNumIters only ever lowers the bound, so everything at or beyond it is never looked at.
__global__ void pnp_bound_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ npts, int iters, int first,
                                 double conf, int batch, int32_t* __restrict__ bound) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int n = npts[b];
  int niters = iters, best_count = 0;
  for (int it = 0; it < niters && it < first; ++it) {
    const int c = counts[(long long)b * iters + it];
    if (c > max(best_count, kPnpK - 1)) {
      best_count = c;
      niters = pnp_update_iters(conf, (double)(n - c) / n, kPnpK, niters);
    }
  }
  bound[b] = niters;
}

__global__ void __launch_bounds__(1024)
pnp_select_kernel(const float* __restrict__ obj, const float2* __restrict__ img, const int32_t* __restrict__ npts,
                  int max_pts, const double* __restrict__ K, const double* __restrict__ models,
                  const int32_t* __restrict__ counts, int iters, float thr2, double conf, uint8_t* __restrict__ mask,
                  int32_t* __restrict__ inl_idx, double* __restrict__ best_model, int32_t* __restrict__ result) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = npts[b];
  __shared__ int s_best, s_run;
  __shared__ double m[12], k[9];
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (tid == 0) {
    int niters = iters, best = -1, best_count = 0, it = 0;
    for (; it < niters; ++it) {
      const int c = counts[(long long)b * iters + it];
      if (c > max(best_count, kPnpK - 1)) {
        best_count = c;
        best = it;
        niters = pnp_update_iters(conf, (double)(n - c) / n, kPnpK, niters);
      }
    }
    s_best = best;
    s_run = it;
    s_base = 0;
  }
  __syncthreads();
  const int best = s_best;
  if (best < 0) {
    for (int i = tid; i < n; i += 1024) mask[(long long)b * max_pts + i] = 0;
    if (tid == 0) {
      result[b * 8 + 0] = 0;
      result[b * 8 + 1] = s_run;
      result[b * 8 + 2] = -1;
    }
    return;
  }
  if (tid < 12) m[tid] = models[((long long)b * iters + best) * 12 + tid];
  if (tid >= 32 && tid < 41) k[tid - 32] = K[b * 9 + tid - 32];
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const long long o = (long long)b * max_pts + i;
    const int in = (i < n) ? (pnp_err(m, k, obj[o * 3], obj[o * 3 + 1], obj[o * 3 + 2], img[o]) <= thr2 ? 1 : 0) : 0;
    if (i < n) mask[o] = (uint8_t)in;
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (in) inl_idx[(long long)b * max_pts + off + __popc(bal & ((1u << lane) - 1))] = i;
    __syncthreads();
    if (tid == 0)
      for (int w = 0; w < 32; ++w) s_base += s_warp[w];
    __syncthreads();
  }
  if (tid < 12) best_model[b * 12 + tid] = m[tid];
  if (tid == 0) {
    result[b * 8 + 0] = s_base;
    result[b * 8 + 1] = s_run;
    result[b * 8 + 2] = best;
  }
}

// ------------------------------------------------------------------------------------------------
// block-wide deterministic sum of NV doubles per thread -> out[NV] (valid in every thread after the call)
constexpr int kPnpRefThreads = 256;
template <int NV>
__device__ void block_sum(const double* v, double* s_part /* [8][NV] */, double* out /* smem [NV] */, int first = 0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int q = first; q < NV; ++q) {   // entries below `first` are known to be zero (not accumulated by the caller)
    const double s = warp_sum_d(v[q]);
    if (lane == 0) s_part[warp * NV + q] = s;
  }
  __syncthreads();
  if (threadIdx.x >= first && threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < kPnpRefThreads / 32; ++w) s += s_part[w * NV + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// projection residual (and Jacobian rows) of one point for parameters (rvec | tvec)
__device__ __forceinline__ void pnp_point(const double* R, const double* dR /* 27 or null */, const double* t, const double* K,
                                          double X, double Y, double Z, double mu, double mv, double* e, double* J /* 12 */) {
  const double xc = R[0] * X + R[1] * Y + R[2] * Z + t[0];
  const double yc = R[3] * X + R[4] * Y + R[5] * Z + t[1];
  const double zc = R[6] * X + R[7] * Y + R[8] * Z + t[2];
  const double z = zc != 0 ? 1.0 / zc : 1.0;
  const double x = xc * z, y = yc * z;
  e[0] = x * K[0] + K[2] - mu;
  e[1] = y * K[4] + K[5] - mv;
  if (!J) return;
  const double fx = K[0], fy = K[4];
  // d(x, y) / d Xc
  const double dx[3] = {z, 0, -x * z}, dy[3] = {0, z, -y * z};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double* d = dR + 9 * i;
    const double a = d[0] * X + d[1] * Y + d[2] * Z, bq = d[3] * X + d[4] * Y + d[5] * Z, cq = d[6] * X + d[7] * Y + d[8] * Z;
    J[i] = fx * (dx[0] * a + dx[2] * cq);
    J[6 + i] = fy * (dy[1] * bq + dy[2] * cq);
    J[3 + i] = fx * dx[i];
    J[9 + i] = fy * dy[i];
  }
}

// closest rotation to a 3 x 3 matrix with positive determinant by Newton's iteration for the polar factor,
// X <- (X + X^-T) / 2 (quadratic; the input is pre-scaled to singular values near one): registers only, where the
// eigen-decomposition route (polar_rotation) walks local-memory arrays.  false: (near-)singular input.
__device__ __forceinline__ bool polar_rotation_newton(const double* A, double* R) {
  double X[9];
  double fro = 0;
#pragma unroll
  for (int q = 0; q < 9; ++q) fro += A[q] * A[q];
  if (!(fro > 0.0) || !isfinite(fro)) return false;
  const double sc = sqrt(3.0 / fro);
#pragma unroll
  for (int q = 0; q < 9; ++q) X[q] = A[q] * sc;
#pragma unroll 1
  for (int it = 0; it < 24; ++it) {
    double C[9];
    C[0] = X[4] * X[8] - X[5] * X[7]; C[1] = X[5] * X[6] - X[3] * X[8]; C[2] = X[3] * X[7] - X[4] * X[6];
    C[3] = X[2] * X[7] - X[1] * X[8]; C[4] = X[0] * X[8] - X[2] * X[6]; C[5] = X[1] * X[6] - X[0] * X[7];
    C[6] = X[1] * X[5] - X[2] * X[4]; C[7] = X[2] * X[3] - X[0] * X[5]; C[8] = X[0] * X[4] - X[1] * X[3];
    const double det = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
    if (!(fabs(det) > 1e-12)) return false;
    const double id = 0.5 / det;
    double diff = 0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const double xn = 0.5 * X[q] + C[q] * id;      // X^-T = cofactor matrix / det
      diff = fmax(diff, fabs(xn - X[q]));
      X[q] = xn;
    }
    if (diff < 1e-15) break;
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) R[q] = X[q];
  return true;
}

constexpr int kPnpCache = 3072;   // inlier correspondences kept in (dynamic) shared memory: 20 bytes each
// impl 1: the first-generation initial pose (L^T L built by one thread per entry over all inliers gathered from global
// memory, 12 x 12 warp Jacobi, polar factor through a 3 x 3 eigen-decomposition: 650 of the kernel's 786 us at 2000
// landmarks).  impl 2: the inliers are compacted into shared memory once, L^T L is assembled from 40 block sums (its
// blocks are sum a a^T, sum x a a^T, sum y a a^T, sum (x^2 + y^2) a a^T with a = (X, Y, Z, 1)), the null direction comes
// from a warp LU + inverse iteration, the polar factor from Newton's iteration in registers.
__global__ void __launch_bounds__(kPnpRefThreads)
pnp_refine_kernel(const float* __restrict__ obj, const float2* __restrict__ img, int max_pts,
                  const double* __restrict__ K, const int32_t* __restrict__ inl_idx, const double* __restrict__ best_model,
                  const int32_t* __restrict__ result, double* __restrict__ pose_out /* batch * 8: rvec, tvec, flags */,
                  int impl, int cache_cap) {
  extern __shared__ __align__(16) unsigned char pnp_dyn[];
  float4* s_xyzu = reinterpret_cast<float4*>(pnp_dyn);
  float* s_vv = reinterpret_cast<float*>(s_xyzu + cache_cap);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = result[b * 8 + 0];
  if (result[b * 8 + 2] < 0 || n <= 0) return;
  __shared__ double s_part[(kPnpRefThreads / 32) * 43];
  __shared__ double s_out[43];
  __shared__ double s_LtL[144], s_V12[144];
  __shared__ double s_param[6], s_prev[6], s_JtJ[36], s_JtErr[6], s_R[9], s_dR[27], s_k[9];
  __shared__ double s_prev_norm;
  __shared__ int s_state, s_lambda, s_iters, s_planar;
  const int32_t* idx = inl_idx + (long long)b * max_pts;
  const float* O = obj + (long long)b * max_pts * 3;
  const float2* I = img + (long long)b * max_pts;
  if (tid < 9) s_k[tid] = K[b * 9 + tid];
  const int nc = impl != 1 ? min(n, cache_cap) : 0;
  for (int i = tid; i < nc; i += kPnpRefThreads) {
    const int p = idx[i];
    const float2 uv = I[p];
    s_xyzu[i] = make_float4(O[p * 3], O[p * 3 + 1], O[p * 3 + 2], uv.x);
    s_vv[i] = uv.y;
  }
  __syncthreads();
  const double ifx = 1.0 / s_k[0], ify = 1.0 / s_k[4], cx = s_k[2], cy = s_k[5];
  // inlier i: object point and image point (shared-memory copy, or gathered for the ones past the cache)
  auto pt = [&](int i, float& X, float& Y, float& Z, float2& uv) {
    if (i < nc) {
      const float4 q = s_xyzu[i];
      X = q.x; Y = q.y; Z = q.z;
      uv = make_float2(q.w, s_vv[i]);
    } else {
      const int p = idx[i];
      X = O[p * 3]; Y = O[p * 3 + 1]; Z = O[p * 3 + 2];
      uv = I[p];
    }
  };

  // ---- planarity of the inlier object points (cvFindExtrinsicCameraParams2: W[2] / W[1] < 1e-3) ----
  {
    double v[9 + 3];
#pragma unroll
    for (int q = 0; q < 12; ++q) v[q] = 0;
    for (int i = tid; i < n; i += kPnpRefThreads) {
      float Xf, Yf, Zf;
      float2 uvf;
      pt(i, Xf, Yf, Zf, uvf);
      const double X = Xf, Y = Yf, Z = Zf;
      v[0] += X; v[1] += Y; v[2] += Z;
      v[3] += X * X; v[4] += X * Y; v[5] += X * Z; v[6] += Y * Y; v[7] += Y * Z; v[8] += Z * Z;
    }
    block_sum<12>(v, s_part, s_out);
    if (tid == 0) {
      const double mx = s_out[0] / n, my = s_out[1] / n, mz = s_out[2] / n;
      double MM[9] = {s_out[3] - n * mx * mx, s_out[4] - n * mx * my, s_out[5] - n * mx * mz,
                      0, s_out[6] - n * my * my, s_out[7] - n * my * mz, 0, 0, s_out[8] - n * mz * mz};
      MM[3] = MM[1]; MM[6] = MM[2]; MM[7] = MM[5];
      double V[9];
      jacobi_eig<3>(MM, V);
      double w0 = MM[0], w1 = MM[4], w2 = MM[8], tq;
      if (w0 < w1) { tq = w0; w0 = w1; w1 = tq; }
      if (w1 < w2) { tq = w1; w1 = w2; w2 = tq; }
      if (w0 < w1) { tq = w0; w0 = w1; w1 = tq; }
      s_planar = (n < 6 || !(w2 / w1 >= 1e-3)) ? 1 : 0;
    }
    __syncthreads();
  }

  // ---- initial pose ----
  if (!s_planar && impl != 1) {
    // DLT: L^T L (12 x 12) of rows [a 0 -x a], [0 a -y a], a = (X, Y, Z, 1), from the 40 sums of its four distinct blocks
    double v40[40];
#pragma unroll
    for (int q = 0; q < 40; ++q) v40[q] = 0;
    for (int i = tid; i < n; i += kPnpRefThreads) {
      float Xf, Yf, Zf;
      float2 uv;
      pt(i, Xf, Yf, Zf, uv);
      const double a[4] = {(double)Xf, (double)Yf, (double)Zf, 1.0};
      const double x = ((double)uv.x - cx) * ifx, y = ((double)uv.y - cy) * ify, rr = x * x + y * y;
      int o = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c2 = r; c2 < 4; ++c2) {
          const double aa = a[r] * a[c2];
          v40[o] += aa;
          v40[10 + o] += x * aa;
          v40[20 + o] += y * aa;
          v40[30 + o] += rr * aa;
          ++o;
        }
    }
    block_sum<40>(v40, s_part, s_out);
    if (tid < 144) {
      const int r = tid / 12, c2 = tid - r * 12;
      const int rb = r >> 2, cb = c2 >> 2, i = min(r & 3, c2 & 3), j = max(r & 3, c2 & 3);
      const int sidx = i * 4 - (i * (i - 1)) / 2 + (j - i);
      double val = 0.0;
      if (rb == cb) val = rb == 2 ? s_out[30 + sidx] : s_out[sidx];
      else if (rb + cb == 2 && rb != 1) val = -s_out[10 + sidx];     // blocks (0, 2) and (2, 0)
      else if (rb + cb == 3) val = -s_out[20 + sidx];                // blocks (1, 2) and (2, 1)
      s_LtL[tid] = val;
    }
    __syncthreads();
    if (tid < 32) smallest_eigvec_spd_warp<12, 6>(s_LtL, s_V12, tid);   // null direction of L (inverse iteration)
    __syncthreads();
    if (tid == 0) {
      double RR[12];
      for (int q = 0; q < 12; ++q) RR[q] = s_V12[q];
      double M3[9] = {RR[0], RR[1], RR[2], RR[4], RR[5], RR[6], RR[8], RR[9], RR[10]};
      if (det3(M3) < 0) {
        for (int q = 0; q < 12; ++q) RR[q] = -RR[q];
        for (int q = 0; q < 9; ++q) M3[q] = -M3[q];
      }
      double sc = 0;
      for (int q = 0; q < 9; ++q) sc += M3[q] * M3[q];
      sc = sqrt(sc);
      double R[9];
      if (!polar_rotation_newton(M3, R)) polar_rotation(M3, R);
      const double f = sqrt(3.0) / sc;     // norm(R) / norm(RR[:, :3])
      rotation_to_rvec(R, s_param);
      s_param[3] = RR[3] * f;
      s_param[4] = RR[7] * f;
      s_param[5] = RR[11] * f;
    }
  } else if (!s_planar) {
    // DLT: L^T L (12 x 12) of rows [X 0 -xX], [0 X -yX] on normalised points; entry (r, c) per thread
    if (tid < 144) {
      const int r = tid / 12, c = tid - r * 12;
      double s = 0;
      for (int i = 0; i < n; ++i) {
        const int p = idx[i];
        const double Xh[4] = {(double)O[p * 3], (double)O[p * 3 + 1], (double)O[p * 3 + 2], 1.0};
        const float2 uv = I[p];
        const double x = ((double)uv.x - cx) * ifx, y = ((double)uv.y - cy) * ify;
        // row 1: [Xh, 0, -x Xh]; row 2: [0, Xh, -y Xh]
        const double a1 = r < 4 ? Xh[r] : (r < 8 ? 0.0 : -x * Xh[r - 8]);
        const double b1 = c < 4 ? Xh[c] : (c < 8 ? 0.0 : -x * Xh[c - 8]);
        const double a2 = r < 4 ? 0.0 : (r < 8 ? Xh[r - 4] : -y * Xh[r - 8]);
        const double b2 = c < 4 ? 0.0 : (c < 8 ? Xh[c - 4] : -y * Xh[c - 8]);
        s += a1 * b1 + a2 * b2;
      }
      s_LtL[tid] = s;
    }
    __syncthreads();
    if (tid < 32) jacobi_eig_warp<12>(s_LtL, s_V12, tid);
    __syncthreads();
    if (tid == 0) {
      const double* A = s_LtL;
      const double* V = s_V12;
      int best = 0;
      for (int q = 1; q < 12; ++q)
        if (A[q * 13] < A[best * 13]) best = q;
      double RR[12];
      for (int q = 0; q < 12; ++q) RR[q] = V[q * 12 + best];
      double M3[9] = {RR[0], RR[1], RR[2], RR[4], RR[5], RR[6], RR[8], RR[9], RR[10]};
      if (det3(M3) < 0) {
        for (int q = 0; q < 12; ++q) RR[q] = -RR[q];
        for (int q = 0; q < 9; ++q) M3[q] = -M3[q];
      }
      double sc = 0;
      for (int q = 0; q < 9; ++q) sc += M3[q] * M3[q];
      sc = sqrt(sc);
      double R[9];
      polar_rotation(M3, R);
      const double f = sqrt(3.0) / sc;     // norm(R) / norm(RR[:, :3])
      rotation_to_rvec(R, s_param);
      s_param[3] = RR[3] * f;
      s_param[4] = RR[7] * f;
      s_param[5] = RR[11] * f;
    }
  } else if (tid == 0) {
    // planar / tiny inlier sets: start from the winning hypothesis (OpenCV starts from a homography decomposition;
    // the Levenberg-Marquardt minimum is the same)
    const double* m = best_model + b * 12;
    rotation_to_rvec(m, s_param);
    s_param[3] = m[9];
    s_param[4] = m[10];
    s_param[5] = m[11];
  }
  if (tid == 0) {
    s_lambda = -3;
    s_iters = 0;
    s_state = 0;
  }
  __syncthreads();

  // ---- CvLevMarq (updateAlt-free variant used by cvFindExtrinsicCameraParams2) ----
  for (int guard = 0; guard < 600; ++guard) {
    const int state = s_state;          // 0: need J and err at param; 1: need err at param (check)
    if (state == 2) break;
    if (tid == 0) rvec_to_rotation(s_param, s_R, state == 0 ? s_dR : nullptr);
    __syncthreads();
    double v[43];
#pragma unroll 1
    for (int q = 0; q < 43; ++q) v[q] = 0;
    for (int i = tid; i < n; i += kPnpRefThreads) {
      float Xf, Yf, Zf;
      float2 uv;
      pt(i, Xf, Yf, Zf, uv);
      double e[2], J[12];
      pnp_point(s_R, state == 0 ? s_dR : nullptr, s_param + 3, s_k, Xf, Yf, Zf, uv.x, uv.y, e,
                state == 0 ? J : nullptr);
      v[42] += e[0] * e[0] + e[1] * e[1];
      if (state == 0) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          v[36 + a] += J[a] * e[0] + J[6 + a] * e[1];
#pragma unroll
          for (int c = a; c < 6; ++c) v[a * 6 + c] += J[a] * J[c] + J[6 + a] * J[6 + c];
        }
      }
    }
    block_sum<43>(v, s_part, s_out, state == 0 ? 0 : 42);   // the check evaluation only needs the error norm
    if (tid == 0) {
      auto step = [&]() {
        const double lam = exp(s_lambda * log(10.0));
        double A[36], rhs[6];
        for (int a = 0; a < 6; ++a) {
          for (int c = 0; c < 6; ++c) A[a * 6 + c] = s_JtJ[a * 6 + c];
          A[a * 7] *= 1.0 + lam;
          rhs[a] = s_JtErr[a];
        }
        lu_solve<6>(A, rhs);
        for (int a = 0; a < 6; ++a) s_param[a] = s_prev[a] - rhs[a];
      };
      const double err_norm = sqrt(s_out[42]);
      if (state == 0) {
        for (int a = 0; a < 6; ++a) {
          for (int c = a; c < 6; ++c) s_JtJ[a * 6 + c] = s_JtJ[c * 6 + a] = s_out[a * 6 + c];
          s_JtErr[a] = s_out[36 + a];
          s_prev[a] = s_param[a];
        }
        s_prev_norm = err_norm;
        step();
        s_state = 1;
      } else {
        bool accepted = true;
        if (err_norm > s_prev_norm) {
          if (++s_lambda <= 16) {
            step();
            accepted = false;
          }
        }
        if (accepted) {
          s_lambda = max(s_lambda - 1, -16);
          double dn = 0, pn = 0;
          for (int a = 0; a < 6; ++a) {
            dn += (s_param[a] - s_prev[a]) * (s_param[a] - s_prev[a]);
            pn += s_prev[a] * s_prev[a];
          }
          if (++s_iters >= 20 || sqrt(dn) / sqrt(pn) < (double)FLT_EPSILON) s_state = 2;
          else s_state = 0;
        }
      }
    }
    __syncthreads();
  }
  if (tid < 6) pose_out[b * 8 + tid] = s_param[tid];
  if (tid == 0) pose_out[b * 8 + 6] = (double)s_planar;
}

// ================================================================================================
int pnp_prepare(mvo_ctx* c, int max_pts, int iters) {
  PnpBufs& p = c->pnp;
  const size_t B = (size_t)c->cfg.batch;
  if (max_pts > p.max_pts) {
    const size_t n = B * (size_t)max_pts;
    MVO_CUDA_TRY(c, p.obj.alloc(n * 3));
    MVO_CUDA_TRY(c, p.img.alloc(n));
    MVO_CUDA_TRY(c, p.xn.alloc(n));
    MVO_CUDA_TRY(c, p.mask.alloc(n));
    MVO_CUDA_TRY(c, p.inl_idx.alloc(n));
    p.max_pts = max_pts;
  }
  if (iters > p.cap_iters) {
    const size_t n = B * (size_t)iters;
    MVO_CUDA_TRY(c, p.subsets.alloc(n * kPnpK));
    MVO_CUDA_TRY(c, p.models.alloc(n * 12));
    MVO_CUDA_TRY(c, p.ok.alloc(n));
    MVO_CUDA_TRY(c, p.counts.alloc(n));
    MVO_CUDA_TRY(c, p.bound.alloc(B));
    p.cap_iters = iters;
  }
  MVO_CUDA_TRY(c, p.npts.alloc(B));
  MVO_CUDA_TRY(c, p.K.alloc(B * 9));
  MVO_CUDA_TRY(c, p.best_model.alloc(B * 12));
  MVO_CUDA_TRY(c, p.result.alloc(B * 8));
  MVO_CUDA_TRY(c, p.pose_out.alloc(B * 8));
  return MVO_OK;
}

// points in p.obj / p.img, counts in p.npts, K in p.K
// cv::undistortPoints(img, K, dist, R = I, P = K) with its default criteria (5 fixed-point iterations): the pixel the
// point would have in a distortion-free camera.  Coefficients (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]).
struct PnpDist {
  double k[12];
};
__global__ void pnp_undistort_kernel(float2* __restrict__ img, const int32_t* __restrict__ npts, int max_pts,
                                     const double* __restrict__ Kd, PnpDist d) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts[b]) return;
  const double* K = Kd + b * 9;
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  float2 q = img[(long long)b * max_pts + i];
  const double x0 = ((double)q.x - cx) / fx, y0 = ((double)q.y - cy) / fy;
  double x = x0, y = y0;
  const double* k = d.k;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  q.x = (float)(x * fx + cx);
  q.y = (float)(y * fy + cy);
  img[(long long)b * max_pts + i] = q;
}

int pnp_run(mvo_ctx* c, int iters, double reproj_err, double conf) {
  PnpBufs& p = c->pnp;
  RansacBufs& r = c->rs;
  const int B = c->cfg.batch;
  const float thr2 = (float)(reproj_err * reproj_err);
  dim3 gn((p.max_pts + 255) / 256, B);
  pnp_normalize_kernel<<<gn, 256, 0, c->stream>>>(p.img.p, p.npts.p, p.max_pts, p.K.p, p.xn.p);
  pnp_sample_kernel<<<B, kPnpSampleThreads, 0, c->stream>>>(r.rng.p, r.rng_len, p.npts.p, iters, p.subsets.p);
  // Batched calls evaluate the hypotheses in two rounds: the adaptive loop usually stops after a dozen of the (up to)
  // `iters` hypotheses, so the first kPnpFirstRound are solved and scored, a one-thread-per-stream replay of the loop
  // gives each stream's bound, and the second launch exits at once for everything beyond it (32 streams x 100
  // hypotheses: 586 -> ~200 us of EPnP).  A single stream runs all of them side by side anyway: one round.
  const bool two_rounds = B >= 4 && iters > kPnpFirstRound && c->dbg_pnp_rounds != 1;
  const int first = two_rounds ? kPnpFirstRound : iters;
  for (int round = 0; round < (two_rounds ? 2 : 1); ++round) {
    const int it0 = round == 0 ? 0 : first, it1 = round == 0 ? first : iters;
    const int32_t* bound = round == 0 ? nullptr : p.bound.p;
    dim3 ge((it1 - it0 + kEpnpWarps - 1) / kEpnpWarps, B);
    // mvo_debug_set("pnp_epnp_impl", 0): the round-1 form of the 12 x 12 eigen-decomposition (cross-check)
    if (c->dbg_pnp_epnp_impl == 0)
      pnp_epnp_kernel<0><<<ge, kEpnpWarps * 32, 0, c->stream>>>(p.obj.p, p.xn.p, p.max_pts, p.subsets.p, iters, it0, it1, bound,
                                                                p.models.p, p.ok.p);
    else
      pnp_epnp_kernel<1><<<ge, kEpnpWarps * 32, 0, c->stream>>>(p.obj.p, p.xn.p, p.max_pts, p.subsets.p, iters, it0, it1, bound,
                                                                p.models.p, p.ok.p);
    dim3 gs(it1 - it0, B);
    pnp_score_kernel<<<gs, 256, 0, c->stream>>>(p.obj.p, p.img.p, p.npts.p, p.max_pts, p.K.p, p.models.p, p.ok.p, iters, it0,
                                               bound, thr2, p.counts.p);
    c->launches += 2;
    if (two_rounds && round == 0) {
      pnp_bound_kernel<<<(B + 127) / 128, 128, 0, c->stream>>>(p.counts.p, p.npts.p, iters, first, conf, B, p.bound.p);
      c->launches++;
    }
  }
  pnp_select_kernel<<<B, 1024, 0, c->stream>>>(p.obj.p, p.img.p, p.npts.p, p.max_pts, p.K.p, p.models.p, p.counts.p,
                                              iters, thr2, conf, p.mask.p, p.inl_idx.p, p.best_model.p, p.result.p);
  {
    static bool attr_set = false;
    if (!attr_set) {
      MVO_CUDA_TRY(c, cudaFuncSetAttribute(pnp_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPnpCache * 20));
      attr_set = true;
    }
    const int cache = c->dbg_pnp_refine_impl == 1 ? 0 : std::min(p.max_pts, kPnpCache);
    pnp_refine_kernel<<<B, kPnpRefThreads, (size_t)cache * 20, c->stream>>>(p.obj.p, p.img.p, p.max_pts, p.K.p, p.inl_idx.p,
                                                                           p.best_model.p, p.result.p, p.pose_out.p,
                                                                           c->dbg_pnp_refine_impl, cache);
  }
  c->launches += 4;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

}  // namespace mvo

using namespace mvo;

extern "C" int mvo_solve_pnp_ransac(mvo_ctx* c, const float* obj_xyz, const float* img_xy, int n, const double* K,
                                    const double* dist, int n_dist, int iterations, double reproj_err, double confidence,
                                    double* rvec, double* tvec, int32_t* inliers, int* n_inliers) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  if (!obj_xyz || !img_xy || !K || !rvec || !tvec || n < 0 || iterations < 1 || iterations > 4096) {
    c->set_error("mvo_solve_pnp_ransac: bad argument");
    return MVO_ERR_INVALID;
  }
  // distortion coefficients (the reference passes CameraInfo's d, src/tracker.cpp:309): the image points are undistorted
  // on the device (cv::undistortPoints, what OpenCV's minimal solver does) and the whole search -- threshold, inlier
  // list, refinement -- runs in the distortion-free camera.  All-zero coefficients take the untouched fast path.
  PnpDist dk;
  for (double& v : dk.k) v = 0.0;
  bool distorted = false;
  for (int i = 0; dist && i < n_dist; ++i) {
    if (dist[i] == 0.0) continue;
    if (i >= 12) {
      c->set_error("mvo_solve_pnp_ransac: tilted-sensor distortion terms (tau_x, tau_y) are not implemented");
      return MVO_ERR_UNSUPPORTED;
    }
    dk.k[i] = dist[i];
    distorted = true;
  }
  if (n < 6) {
    // OpenCV switches to P3P (n == 4) / one EPnP solve on all points (n == 5); the reference cannot get here
    // (Tracker::update goes LOST below min_tracked_points = 10 observations, src/tracker.cpp:293-297)
    c->set_error("mvo_solve_pnp_ransac: fewer than 6 correspondences");
    return MVO_ERR_DEGENERATE;
  }
  if (c->cfg.batch != 1) {
    c->set_error("the single-call geometry API needs a batch==1 context");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  int rc = ransac_prepare(c, std::max(c->rs.max_pts, 64), std::max(c->rs.cap_iters, 1));   // the RNG table lives there
  if (rc) return rc;
  rc = pnp_prepare(c, std::max(n, c->pnp.max_pts), std::max(iterations, c->pnp.cap_iters));
  if (rc) return rc;
  PnpBufs& p = c->pnp;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(p.obj.p, obj_xyz, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(p.img.p, img_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(p.npts.p, &n, 4, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(p.K.p, K, 72, cudaMemcpyHostToDevice, c->stream));
  if (distorted) {
    pnp_undistort_kernel<<<dim3((n + 255) / 256, 1), 256, 0, c->stream>>>(p.img.p, p.npts.p, p.max_pts, p.K.p, dk);
    c->launches++;
  }
  rc = pnp_run(c, iterations, reproj_err, confidence);
  if (rc) return rc;
  int res[8];
  double pose[8];
  MVO_CUDA_TRY(c, cudaMemcpyAsync(res, p.result.p, 32, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(pose, p.pose_out.p, 64, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (n_inliers) *n_inliers = res[0];
  c->last_ransac_iters = res[1];
  if (res[2] < 0) {
    c->set_error("solvePnPRansac found no model");
    return MVO_ERR_DEGENERATE;
  }
  if (inliers && res[0] > 0) {
    MVO_CUDA_TRY(c, cudaMemcpyAsync(inliers, p.inl_idx.p, (size_t)res[0] * 4, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  for (int i = 0; i < 3; ++i) {
    rvec[i] = pose[i];
    tvec[i] = pose[3 + i];
  }
  return MVO_OK;
}

/* parity hook: the hypotheses of the last mvo_solve_pnp_ransac call (stream 0): 5 sample indices, R|t (12 doubles)
 * and inlier count (-1: solver failed) per iteration */
extern "C" int mvo_pnp_get_hypotheses(mvo_ctx* c, int iterations, int32_t* subsets, double* models, int32_t* counts) {
  if (!c || iterations < 1 || iterations > c->pnp.cap_iters) return MVO_ERR_INVALID;
  PnpBufs& p = c->pnp;
  if (subsets) MVO_CUDA_TRY(c, cudaMemcpyAsync(subsets, p.subsets.p, (size_t)iterations * kPnpK * 4, cudaMemcpyDeviceToHost, c->stream));
  if (models) MVO_CUDA_TRY(c, cudaMemcpyAsync(models, p.models.p, (size_t)iterations * 96, cudaMemcpyDeviceToHost, c->stream));
  if (counts) MVO_CUDA_TRY(c, cudaMemcpyAsync(counts, p.counts.p, (size_t)iterations * 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

/* cv::Rodrigues(rvec, R): src/tracker.cpp:315 -- host-side 3x3, no device work */
extern "C" int mvo_rodrigues(const double* rvec, double* R) {
  if (!rvec || !R) return MVO_ERR_INVALID;
  const double theta = sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return MVO_OK;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
  const double k[3] = {rvec[0] * it, rvec[1] * it, rvec[2] * it};
  const double rrt[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1], k[1] * k[2], k[0] * k[2], k[1] * k[2], k[2] * k[2]};
  const double rx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
  for (int i = 0; i < 9; ++i) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i] + s * rx[i];
  return MVO_OK;
}
