// ransac.cu -- deterministic parallel RANSAC for H / F / E with OpenCV's RNG stream and adaptive stop.
//
// Replaces cv::findHomography / findFundamentalMat / findEssentialMat at
// /root/reference/src/initializer.cpp:82,87,228 and src/tracker.cpp:243,248.  Contract: SURVEY.md A.4.
//
// OpenCV's loop is sequential (draw subset -> solve -> score -> maybe shrink niters).  The RNG stream only
// depends on data-only subset checks, so here (per stream of the group):
//   1. ransac_attempt_kernel  every draw offset of a window simulates one getSubset attempt (grid-wide);
//      ransac_chain_kernel    pointer doubling in shared memory follows the attempt chain; an ordered
//                             compaction yields the exact subsets OpenCV would draw, in iteration order.
//   2. *_solve_kernel         one thread per hypothesis (4-pt H, 7-pt F, 5-pt E; FP64).
//   3. ransac_score_kernel    one block per hypothesis, warp-reduced inlier counts (FP32 for H, FP64 F/E).
//   4. ransac_select_kernel   prefix-max scan replays the strict '>' update + shrinking niters exactly,
//                             then writes the winner's mask.
//   5. h_refine_kernel        DLT on the inliers + 10 Levenberg-Marquardt iterations + final mask (H only).
// The hypotheses are evaluated in two rounds: iterations [0, kRound0) first; the select kernel then knows
// whether OpenCV's loop would already have stopped (it usually has: 10..150 iterations at the inlier ratios
// a VO front-end sees) and sets a per-stream `done` flag that turns round 1 (the remaining iterations up to
// maxIters) into empty launches.  The result is identical to evaluating all maxIters hypotheses.
#include "context.cuh"
#include "solvers.cuh"
#include <algorithm>
#include <vector>

namespace mvo {

constexpr int kDraws = 32768;         // largest RNG window of one sampler pass (uint16 offsets)
constexpr int kMaxAttempts = 16384;   // getSubset attempts followed per pass (large window)
constexpr int kSegIters = 2048;       // subsets produced per sampler pass
constexpr int kRound0 = 128;          // iterations evaluated before the second early-exit check
constexpr int kDraws0 = 4096;         // window / attempts of that sampler pass
constexpr int kAttempts0 = 1024;
#ifndef MVO_RANSAC_ROUND_A
#define MVO_RANSAC_ROUND_A 32
#endif
constexpr int kScoreGrid = 128;       // CTAs per stream of the scoring kernel (it strides over the hypotheses)
constexpr int kRoundA = MVO_RANSAC_ROUND_A;   // iterations evaluated before the first early-exit check (0: no such round)
constexpr int kDrawsA = 1024;
constexpr int kAttemptsA = 256;

template <int MODEL> struct MT;
template <> struct MT<MVO_MODEL_H> { static constexpr int K = 4, MAXIT = 2000, MAXM = 1; };
template <> struct MT<MVO_MODEL_F> { static constexpr int K = 7, MAXIT = 1000, MAXM = 3; };
template <> struct MT<MVO_MODEL_E> { static constexpr int K = 5, MAXIT = 1000, MAXM = 10; };

// ---- subset drawing ----------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ unsigned draw_subset(const uint32_t* __restrict__ rng, unsigned o, unsigned n, int* idx,
                                                unsigned window) {
#pragma unroll 1
  for (int i = 0; i < K; ++i) {
    for (;;) {
      if (o >= window) return window;   // sentinel: ran out of draws
      const int v = (int)(rng[o++] % n);
      bool dup = false;
#pragma unroll 1
      for (int j = 0; j < i; ++j) dup |= (idx[j] == v);
      if (!dup) {
        idx[i] = v;
        break;
      }
    }
  }
  return o;
}

template <int K>
__device__ __forceinline__ bool have_collinear(const float2* p) {
  const int i = K - 1;
#pragma unroll 1
  for (int j = 0; j < i; ++j) {
    const double dx1 = (double)__fsub_rn(p[j].x, p[i].x), dy1 = (double)__fsub_rn(p[j].y, p[i].y);
#pragma unroll 1
    for (int k = 0; k < j; ++k) {
      const double dx2 = (double)__fsub_rn(p[k].x, p[i].x), dy2 = (double)__fsub_rn(p[k].y, p[i].y);
      if (fabs(__dsub_rn(__dmul_rn(dx2, dy1), __dmul_rn(dy2, dx1))) <=
          (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
        return true;
    }
  }
  return false;
}

__device__ __forceinline__ double det3_pts(float2 a, float2 b, float2 c) {
  // determinant(Matx33d(a.x, a.y, 1, b.x, b.y, 1, c.x, c.y, 1))
  const double a00 = a.x, a01 = a.y, a10 = b.x, a11 = b.y, a20 = c.x, a21 = c.y;
  return __dadd_rn(__dsub_rn(__dmul_rn(a00, __dsub_rn(a11, a21)), __dmul_rn(a01, __dsub_rn(a10, a20))),
                   __dsub_rn(__dmul_rn(a10, a21), __dmul_rn(a20, a11)));
}

template <int MODEL>
__device__ __forceinline__ bool check_subset(const float2* s1, const float2* s2) {
  if (MODEL == MVO_MODEL_E) return true;
  constexpr int K = MT<MODEL>::K;
  if (have_collinear<K>(s1) || have_collinear<K>(s2)) return false;
  if (MODEL == MVO_MODEL_H) {
    const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
    int negative = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double da = det3_pts(s1[tt[i][0]], s1[tt[i][1]], s1[tt[i][2]]);
      const double db = det3_pts(s2[tt[i][0]], s2[tt[i][1]], s2[tt[i][2]]);
      negative += (__dmul_rn(da, db) < 0) ? 1 : 0;
    }
    if (negative != 0 && negative != 4) return false;
  }
  return true;
}

// state per stream (8 ints): [0] draw offset into the global RNG table, [1] subsets produced so far, [2] flags
// (2: RNG table exhausted, 4: no further subset can be drawn), [3] done (the adaptive loop has terminated inside
// the evaluated range; later rounds are skipped), [4] hypotheses already solved + scored by earlier rounds, [5] the
// adaptive loop's iteration bound after the last evaluated round (niters only ever shrinks, so no later round needs
// subsets, models or scores beyond it; 0 = not known yet)
template <int MODEL>
__global__ void __launch_bounds__(256)
ransac_attempt_kernel(const uint32_t* __restrict__ rng_table, int rng_len, const float2* __restrict__ p1,
                      const float2* __restrict__ p2, const int32_t* __restrict__ npts, int max_pts,
                      const int32_t* __restrict__ state, int want_total, int window, uint16_t* __restrict__ att_next,
                      uint8_t* __restrict__ att_ok) {
  constexpr int K = MT<MODEL>::K;
  const int b = blockIdx.y;
  const int32_t* st = state + b * 8;
  if (st[5] > 0) want_total = min(want_total, st[5]);
  if (st[3] || st[1] >= want_total) return;
  const int n = npts[b];
  const int draw_base = st[0];
  if (n <= K || draw_base + window > rng_len) return;
  const unsigned o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (unsigned)window) return;
  const uint32_t* rng = rng_table + draw_base;
  int idx[K];
  const unsigned e = draw_subset<K>(rng, o, (unsigned)n, idx, (unsigned)window);
  bool ok = false;
  if (e != (unsigned)window) {
    ok = true;
    if (MODEL != MVO_MODEL_E) {
      const float2* q1 = p1 + (long long)b * max_pts;
      const float2* q2 = p2 + (long long)b * max_pts;
      float2 s1[K], s2[K];
#pragma unroll 1
      for (int i = 0; i < K; ++i) {
        s1[i] = q1[idx[i]];
        s2[i] = q2[idx[i]];
      }
      ok = check_subset<MODEL>(s1, s2);
    }
  }
  att_next[(long long)b * kDraws + o] = (uint16_t)e;
  att_ok[(long long)b * kDraws + o] = ok ? 1 : 0;
}

template <int MODEL>
__global__ void __launch_bounds__(1024)
ransac_chain_kernel(const uint32_t* __restrict__ rng_table, int rng_len, const int32_t* __restrict__ npts,
                    const uint16_t* __restrict__ att_next, const uint8_t* __restrict__ att_ok,
                    int32_t* __restrict__ subsets, int32_t* __restrict__ state, int want_total, int cap_iters,
                    int window, int max_attempts) {
  constexpr int K = MT<MODEL>::K;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* J0 = reinterpret_cast<uint16_t*>(smem_raw);                  // window + 1 (+1 pad)
  uint16_t* J1 = J0 + (window + 2);
  uint16_t* pos = J1 + (window + 2);                                     // max_attempts
  uint16_t* start = pos + max_attempts;                                  // kSegIters
  uint8_t* okb = reinterpret_cast<uint8_t*>(start + kSegIters);          // window
  __shared__ int s_warp[32];
  __shared__ int s_base, s_last_attempt;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = npts[b];
  int32_t* st = state + b * 8;
  if (st[3]) return;
  if (st[5] > 0) want_total = min(want_total, st[5]);
  const int have = st[1];
  const int draw_base = st[0];
  if (have >= want_total) return;
  int32_t* out = subsets + ((long long)b * cap_iters) * K;
  if (n <= K) {
    // count == modelPoints: OpenCV solves the points once; fewer: nothing to do
    if (tid == 0) {
      if (n == K && have == 0) {
        for (int i = 0; i < K; ++i) out[i] = i;
        st[1] = 1;
      }
      st[2] |= 4;  // exhausted
    }
    return;
  }
  if (draw_base + window > rng_len) {
    if (tid == 0) st[2] |= 2;  // RNG table exhausted
    return;
  }
  const uint32_t* rng = rng_table + draw_base;
  const unsigned uw = (unsigned)window;
  for (int o = tid; o <= window; o += 1024) {
    J0[o] = (o < window) ? att_next[(long long)b * kDraws + o] : (uint16_t)window;
    if (o < window) okb[o] = att_ok[(long long)b * kDraws + o];
  }
  if (tid == 0) {
    pos[0] = 0;
    s_base = 0;
    s_last_attempt = -1;
  }
  __syncthreads();
  // attempt chain by pointer doubling: pos[a] = offset where attempt a starts
  uint16_t* cur = J0;
  uint16_t* nxt = J1;
  for (int len = 1; len < max_attempts; len <<= 1) {
    for (int j = tid; j < len; j += 1024) pos[j + len] = cur[pos[j]];
    for (int o = tid; o <= window; o += 1024) nxt[o] = cur[cur[o]];
    __syncthreads();
    uint16_t* t = cur;
    cur = nxt;
    nxt = t;
  }
  // ordered compaction of the successful attempts
  const int want = min(kSegIters, want_total - have);
  for (int a0 = 0; a0 < max_attempts; a0 += 1024) {
    if (s_base >= want) break;
    const int a = a0 + tid;
    const unsigned o = (a < max_attempts) ? pos[a] : uw;
    const int ok = (o < uw) ? (int)okb[o] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= s) v += t;
      }
      s_warp[lane] = v;
    }
    __syncthreads();
    const int rank = s_base + (warp ? s_warp[warp - 1] : 0) + within;
    if (ok && rank < want) {
      start[rank] = (uint16_t)o;
      if (rank == want - 1) s_last_attempt = a;
    }
    __syncthreads();
    if (tid == 0) s_base += s_warp[31];
    __syncthreads();
  }
  const int got = min(s_base, want);
  // materialise the index tuples
  for (int i = tid; i < got; i += 1024) {
    int idx[K];
    draw_subset<K>(rng, start[i], (unsigned)n, idx, uw);
#pragma unroll 1
    for (int k = 0; k < K; ++k) out[(long long)(have + i) * K + k] = idx[k];
  }
  if (tid == 0) {
    st[1] = have + got;
    // where does the next pass continue?  right after the last attempt this pass consumed
    int a_last = (got == want) ? s_last_attempt : max_attempts - 1;
    while (a_last > 0 && pos[a_last] >= uw) --a_last;
    int idx[K];
    const unsigned e = draw_subset<K>(rng, pos[a_last], (unsigned)n, idx, uw);
    if (e != uw) {
      st[0] = draw_base + (int)e;
    } else {
      st[0] = draw_base + pos[a_last];          // unfinished attempt: redo it with a fresh window
      if (pos[a_last] == 0) st[2] |= 4;         // no progress possible
    }
  }
}

// ---- solve ---------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(64)
ransac_solve_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, const double2* __restrict__ q1,
                    const double2* __restrict__ q2, int max_pts, const int32_t* __restrict__ subsets,
                    const int32_t* __restrict__ state, int cap_iters, int h0, int h1, double* __restrict__ models,
                    int32_t* __restrict__ nmodels) {
  constexpr int K = MT<MODEL>::K, MAXM = MT<MODEL>::MAXM;
  const int b = blockIdx.y;
  if (state[b * 8 + 3]) return;   // done: the adaptive loop ended in an earlier round
  const int h = h0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int nsub = min(min(state[b * 8 + 1], cap_iters), h1);
  if (h >= nsub || h < state[b * 8 + 4]) return;
  const int32_t* idx = subsets + ((long long)b * cap_iters + h) * K;
  double* out = models + ((long long)b * cap_iters + h) * MAXM * 9;
  int nm = 0;
  if (MODEL == MVO_MODEL_E) {
    double2 a[K], c[K];
    for (int i = 0; i < K; ++i) {
      a[i] = q1[(long long)b * max_pts + idx[i]];
      c[i] = q2[(long long)b * max_pts + idx[i]];
    }
    double E[10 * 9];
    nm = solve_e5(a, c, E);
    for (int i = 0; i < nm * 9; ++i) out[i] = E[i];
  } else {
    float2 a[K], c[K];
    for (int i = 0; i < K; ++i) {
      a[i] = p1[(long long)b * max_pts + idx[i]];
      c[i] = p2[(long long)b * max_pts + idx[i]];
    }
    if (MODEL == MVO_MODEL_H) {
      double H[9];
      nm = solve_h4(a, c, H);
      if (nm)
        for (int i = 0; i < 9; ++i) out[i] = H[i];
    } else {
      double F[27];
      nm = solve_f7(a, c, F);
      for (int i = 0; i < nm * 9; ++i) out[i] = F[i];
    }
  }
  nmodels[(long long)b * cap_iters + h] = nm;
}

// ---- 5-point solver, split for latency: per-thread algebra, then 16 lanes per hypothesis for the roots ----
constexpr int kE5Scratch = 36 + 39 + 11;   // EE basis, Bm, det polynomial

constexpr int kE5SetupWarps = 4;   // one warp per hypothesis (e5_setup_warp)
__global__ void __launch_bounds__(kE5SetupWarps * 32)
e5_setup_kernel(const double2* __restrict__ q1, const double2* __restrict__ q2, int max_pts,
                const int32_t* __restrict__ subsets, const int32_t* __restrict__ state, int cap_iters, int h0, int h1,
                double* __restrict__ scratch) {
  __shared__ double s_M[kE5SetupWarps][200];
  const int b = blockIdx.y;
  if (state[b * 8 + 3]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = h0 + blockIdx.x * kE5SetupWarps + warp;
  const int nsub = min(min(state[b * 8 + 1], cap_iters), h1);
  if (h >= nsub || h < state[b * 8 + 4]) return;   // warp-uniform
  const int32_t* idx = subsets + ((long long)b * cap_iters + h) * 5;
  double2 a[5], c[5];
  for (int i = 0; i < 5; ++i) {
    a[i] = q1[(long long)b * max_pts + idx[i]];
    c[i] = q2[(long long)b * max_pts + idx[i]];
  }
  double EE[36], Bm[39], detp[11];
  const bool ok = e5_setup_warp(a, c, s_M[warp], lane, EE, Bm, detp);
  double* out = scratch + ((long long)b * cap_iters + h) * kE5Scratch;
  if (lane == 0) {
    for (int i = 0; i < 36; ++i) out[i] = EE[i];
    for (int i = 0; i < 39; ++i) out[36 + i] = ok ? Bm[i] : 0.0;
    for (int i = 0; i < 11; ++i) out[75 + i] = ok ? detp[i] : 0.0;
  }
}

// midpoint of [a, b] in the ordered-integer representation of IEEE doubles: at most ~11 steps to pin the
// binade of a root however wide the bracket is
__device__ __forceinline__ double mid_key(double a, double b) {
  long long ka = __double_as_longlong(a), kb = __double_as_longlong(b);
  ka = ka < 0 ? -(ka & 0x7fffffffffffffffLL) : ka;
  kb = kb < 0 ? -(kb & 0x7fffffffffffffffLL) : kb;
  long long km = (ka >> 1) + (kb >> 1) + (ka & kb & 1);
  if (km < 0) km = (long long)(0x8000000000000000ULL | (unsigned long long)(-km));
  return __longlong_as_double(km);
}

// value and derivative of q (degree qd, ascending coefficients) by Horner:  v = q[qd]; for i = qd-1 .. 0: dv = dv x + v,
// v = v x + q[i].
// Run on a register-resident copy of the coefficients, zero-padded to degree 10 and fully unrolled: the leading zero
// terms leave v = dv = 0 until the true leading coefficient is reached, so value and derivative are bit-identical to the
// variable-degree loop -- without a local-memory load in every step of the dependent chain (the 5-point root finder
// spends its time here: ~25 evaluations per root and derivative level).
__device__ __forceinline__ double poly_eval10(const double (&q)[11], double x, double& dv) {
  double v = q[10];
  dv = 0.0;
#pragma unroll
  for (int i = 9; i >= 0; --i) {
    dv = dv * x + v;
    v = v * x + q[i];
  }
  return v;
}

// root of q inside (lo, hi) given a sign change: Newton safeguarded by bisection in key space
__device__ inline double bracket_root(const double (&q)[11], double lo, double hi, double flo) {
  double xl = (flo < 0.0) ? lo : hi, xh = (flo < 0.0) ? hi : lo;   // f(xl) < 0 < f(xh)
  double x = mid_key(lo, hi), dfx;
  double fx = poly_eval10(q, x, dfx);
  double dxold = fabs(hi - lo);
#pragma unroll 1
  for (int it = 0; it < 100; ++it) {
    if (fx == 0.0) break;
    if (fx < 0.0) xl = x; else xh = x;
    double xn = x - fx / dfx;
    const double a = fmin(xl, xh), bnd = fmax(xl, xh);
    const double dx = fabs(xn - x);
    if (!(xn > a && xn < bnd) || 2.0 * dx > dxold) xn = mid_key(xl, xh);   // bisect
    dxold = fabs(xn - x);
    if (xn == x) break;
    const bool conv = dxold <= 4e-16 * fabs(xn);
    x = xn;
    if (conv) break;
    fx = poly_eval10(q, x, dfx);
  }
  return x;
}

// All ten roots of the degree-10 determinant polynomial at once: Ehrlich-Aberth iteration (simultaneous Newton steps
// with the other roots deflated implicitly: w_k -= N_k / (1 - N_k sum_{j != k} 1 / (w_k - w_j)), N = p / p'), one lane per
// root inside the 16-lane group of a hypothesis, started from Bini's initial guesses (the upper convex hull of
// (k, log |c_k|) gives the root moduli: the roots of these polynomials spread over four orders of magnitude).  About ten
// iterations of ~1000 dependent cycles replace the derivative-level bracketing below (10 levels x up to 100 safeguarded
// Newton / bisection steps: 148 us for the 32 hypotheses of a first round, the longest kernel on a single stream's
// critical path).  Like cv::solvePoly in the reference's findEssentialMat, it finds the complex roots and keeps those with
// |Im z| <= 1e-10; they are polished by Newton steps on the real polynomial and handed out in ascending order, as the
// bracketing path does.  Returns false (uniform inside the 16-lane group) when the iteration did not converge or the
// polynomial is degenerate: the caller then runs the bracketing path.
#ifdef MVO_DK_DEBUG
#define g_dbg_h dbg_h
__device__ __forceinline__ bool aberth_real_roots(const double (&c)[11], int l16, double& myroot, int& ncrit, int dbg_h) {
#else
__device__ __forceinline__ bool aberth_real_roots(const double (&c)[11], int l16, double& myroot, int& ncrit) {
#endif
  const unsigned full = 0xffffffffu;
  // `usable` is uniform inside the 16-lane group but not across the warp's two groups: no early return before the
  // warp-wide shuffles below, an unusable group just idles through the iteration
  bool usable = true;
#pragma unroll
  for (int k = 0; k <= 10; ++k) usable = usable && isfinite(c[k]);
  usable = usable && (c[10] != 0.0) && (c[0] != 0.0);
  // ---- initial guesses: lane k holds log |c_k| (lanes 0 .. 10), the hull walk runs redundantly in every lane ----
  const double mylog = (l16 <= 10 && usable && c[l16 <= 10 ? l16 : 0] != 0.0) ? log(fabs(c[l16 <= 10 ? l16 : 0])) : -1e300;
  double lg[11];
#pragma unroll
  for (int k = 0; k <= 10; ++k) lg[k] = __shfl_sync(full, mylog, k, 16);
  double wr = 0.0, wi = 0.0;
  {
    int k0 = 0, seg = 0;
    double rad = 1.0;
    int seg_lo = 0, seg_hi = 10, seg_id = 0;
#pragma unroll 1
    while (k0 < 10) {
      int best = k0 + 1;
      double bs = -1e300;
#pragma unroll 1
      for (int j = k0 + 1; j <= 10; ++j) {
        const double sl = (lg[j] - lg[k0]) / (double)(j - k0);
        if (sl >= bs) {
          bs = sl;
          best = j;
        }
      }
      if (l16 >= k0 && l16 < best) {
        rad = exp(-bs);                 // |c_k0 / c_best|^(1 / (best - k0))
        seg_lo = k0;
        seg_hi = best;
        seg_id = seg;
      }
      k0 = best;
      ++seg;
    }
    if (!(rad > 1e-150 && rad < 1e150)) usable = false;
    const double ang = 2.0 * (double)(l16 - seg_lo) / (double)(seg_hi - seg_lo) + 0.2 * (double)seg_id + 0.223;
    double sn, cs;
    sincospi(ang, &sn, &cs);
    wr = rad * cs;
    wi = rad * sn;
  }
  {
    // a group is usable only if all of its lanes are
    const unsigned ub = __ballot_sync(full, usable);
    const int half = (threadIdx.x >> 4) & 1;
    usable = ((ub >> (16 * half)) & 0xffffu) == 0xffffu;
  }
  const bool act = l16 < 10 && usable;
  bool conv = !act;
  bool done = false;
  int nres = 0;
  double err2 = 0.0;
#pragma unroll 1
  for (int it = 0; it < 48; ++it) {
    // p(w), p'(w): Horner with real coefficients
    double pr = c[10], pi = 0.0, dr = 0.0, di = 0.0;
    const double aw = sqrt(wr * wr + wi * wi);
    double ab = fabs(c[10]);                    // sum |c_k| |w|^k: the rounding-error scale of the Horner value
#pragma unroll
    for (int k = 9; k >= 0; --k) {
      const double tr = dr * wr - di * wi + pr;
      di = dr * wi + di * wr + pi;
      dr = tr;
      const double t = pr * wr - pi * wi + c[k];
      pi = pr * wi + pi * wr;
      pr = t;
      ab = ab * aw + fabs(c[k]);
    }
    // a root is final when its correction is below 1e-13 |w| (simple roots: the cubic convergence gets there one or two
    // steps after the residual does) or when |p(w)| has been down at the rounding level of its own evaluation for
    // three steps (clustered roots never satisfy a step-size test: their corrections jitter at eps^(1/m))
    nres = (pr * pr + pi * pi <= 1e-29 * ab * ab) ? nres + 1 : 0;
    if (nres >= 3) conv = true;
    // first-order bound on what the evaluation noise lets this root be known to: 4e-15 sum|c_k||w|^k / |p'(w)|
    if (!conv || nres >= 3) err2 = (1.6e-29 * ab * ab) / fmax(dr * dr + di * di, 1e-300);
    // S = sum_{j != k} 1 / (w - w_j)
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const double er = wr - __shfl_sync(full, wr, j, 16), ei = wi - __shfl_sync(full, wi, j, 16);
      const double m = er * er + ei * ei;
      if (j != l16 && m > 0.0) {
        const double im = 1.0 / m;
        sr += er * im;
        si -= ei * im;
      }
    }
    if (!conv) {
      const double dm = dr * dr + di * di;
      double nr = 0.0, ni = 0.0;                 // N = p / p'
      if (dm > 0.0) {
        const double im = 1.0 / dm;
        nr = (pr * dr + pi * di) * im;
        ni = (pi * dr - pr * di) * im;
      }
      // corr = N / (1 - N S)
      const double qr = 1.0 - (nr * sr - ni * si), qi = -(nr * si + ni * sr);
      const double qm = qr * qr + qi * qi;
      double cr = nr, ci = ni;
      if (qm > 0.0) {
        const double im = 1.0 / qm;
        cr = (nr * qr + ni * qi) * im;
        ci = (ni * qr - nr * qi) * im;
      }
      wr -= cr;
      wi -= ci;
      conv = (cr * cr + ci * ci <= 1e-26 * (wr * wr + wi * wi)) || (pr == 0.0 && pi == 0.0);
    }
    const bool bad = act && !(isfinite(wr) && isfinite(wi));
    const unsigned cb = __ballot_sync(full, conv), bb = __ballot_sync(full, bad);
    if (bb) break;                                               // (warp-uniform: both groups fall back)
    if (cb == full) {
      done = true;
      break;
    }
#ifdef MVO_DK_DEBUG
    if (it == 47) printf("noconv lane %d usable %d conv %d w = %g %g cb %08x\n", threadIdx.x & 31, (int)usable, (int)conv, wr, wi, cb);
#endif
  }
#ifdef MVO_DK_DEBUG
  if ((threadIdx.x & 31) == 0) printf("aberth done=%d\n", (int)done);
#endif
  if (!done) return false;
  // A pair of close real roots is a cluster the iteration resolves only to ~sqrt(eps): it ends as m +- i e with a small
  // e > 1e-10 and both roots would be lost.  Such a group (an imaginary part that is neither zero nor clearly non-zero)
  // is left to the bracketing path, which separates the pair by its sign changes.
  // A root inside a tight cluster (|p'| tiny: the roots of some samples all lie within a few per cent of each other) is
  // not known to better than err = noise / |p'|; a group with such a root goes to the bracketing path as well.
  {
    const double w2 = wr * wr + wi * wi;
    const bool vague = act && err2 > 1e-10 * w2;
    const bool suspicious = vague || (act && fabs(wi) > 1e-10 && fabs(wi) <= fmax(1e-6 * fmax(1.0, fabs(wr)), 16.0 * sqrt(err2)));
    const unsigned sb = __ballot_sync(full, suspicious);
    const int half0 = (threadIdx.x >> 4) & 1;
    if ((sb >> (16 * half0)) & 0xffffu) usable = false;
  }
  // classify, polish on the real polynomial, ascending order
  double z = wr;
  bool real = act && fabs(wi) <= 1e-10;
  if (real) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      double dv;
      const double v = poly_eval10(c, z, dv);
      if (dv == 0.0 || !isfinite(v)) break;
      const double zn = z - v / dv;
      if (!isfinite(zn) || zn == z) break;
      z = zn;
    }
  }
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    const double zj = __shfl_sync(full, z, j, 16);
    const bool rj = __shfl_sync(full, (int)real, j, 16) != 0;
    if (rj && (zj < z || (zj == z && j < l16))) ++rank;
  }
  const unsigned rb = __ballot_sync(full, real);
  const int half = (threadIdx.x >> 4) & 1;
  ncrit = __popc((rb >> (16 * half)) & 0xffffu);
  double out = 0.0;
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    const double zj = __shfl_sync(full, z, j, 16);
    const int rkj = __shfl_sync(full, rank, j, 16);
    const bool rj = __shfl_sync(full, (int)real, j, 16) != 0;
    if (rj && rkj == l16) out = zj;
  }
#ifdef MVO_DK_DEBUG
  if (g_dbg_h == 423 && l16 < 11)
    printf("h423 lane %d c=%g w=(%.12g, %.3g) z=%.12g real=%d rank=%d usable=%d ncrit=%d out=%.12g\n", l16, c[l16], wr, wi, z, (int)real, rank,
           (int)usable, ncrit, out);
#endif
  myroot = out;
  return usable;
}

#ifdef MVO_DK_DEBUG
__device__ unsigned g_e5_stats[2];
#endif
constexpr int kRootsThreads = 128;   // 8 hypotheses per block, 16 lanes each
__global__ void __launch_bounds__(kRootsThreads)
e5_roots_kernel(const int32_t* __restrict__ state, int cap_iters, int h0, int h1, const double* __restrict__ scratch,
                double* __restrict__ models, int32_t* __restrict__ nmodels, int impl) {
  const int b = blockIdx.y;
  if (state[b * 8 + 3]) return;
  const int grp = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const int h = h0 + blockIdx.x * (kRootsThreads / 16) + grp;
  const int nsub = min(min(state[b * 8 + 1], cap_iters), h1);
  const bool live = (h < nsub) && (h >= state[b * 8 + 4]);
  const unsigned full = 0xffffffffu;
  const int half = (threadIdx.x >> 4) & 1;          // which 16-lane half of the warp
  const double* sc = scratch + ((long long)b * cap_iters + (live ? h : 0)) * kE5Scratch;
  double c[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) c[i] = live ? sc[75 + i] : 0.0;
  int deg = 10;
  while (deg > 0 && c[deg] == 0.0) --deg;
  // Fujiwara bound on |root|: 2 max_k |c[deg-k] / c[deg]|^(1/k); lane k computes term k
  double bound = 0.0;
  if (deg > 0 && l16 >= 1 && l16 <= deg) bound = pow(fabs(c[deg - l16] / c[deg]), 1.0 / (double)l16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) bound = fmax(bound, __shfl_xor_sync(full, bound, o, 16));
  bound = 2.0 * bound + 1e-300;
  double myroot = 0.0;
  int ncrit = 0;
  // all roots at once (Ehrlich-Aberth iteration); the derivative-level bracketing only where that did not converge
  bool have = false;
  if (impl != 1) {
    double r2 = 0.0;
    int n2 = 0;
#ifdef MVO_DK_DEBUG
    const bool ok = aberth_real_roots(c, l16, r2, n2, live ? h : -1) && deg == 10;
    if (h == 423 && live && l16 == 0) printf("h423 ok=%d deg=%d n2=%d\n", (int)ok, deg, n2);
#else
    const bool ok = aberth_real_roots(c, l16, r2, n2) && deg == 10;
#endif
    // group-uniform result; the bracketing loop below is warp-level code, so it runs when either group needs it
    if (ok) {
      myroot = r2;
      ncrit = n2;
      have = true;
    }
#ifdef MVO_DK_DEBUG
    if (live && l16 == 0) atomicAdd(&g_e5_stats[ok ? 0 : 1], 1u);
#endif
  }
  const bool need_levels = __any_sync(full, !have);
  double root_keep = myroot;
  int ncrit_keep = ncrit;
  if (need_levels) {
    myroot = 0.0;
    ncrit = 0;
  }
#pragma unroll 1
  for (int lvl = need_levels ? 9 : -1; lvl >= 0; --lvl) {
    const bool lvl_on = (deg > 0) && (lvl <= deg - 1);
    const int qd = deg - lvl;
    double q[11];
    if (lvl_on) {
#pragma unroll 1
      for (int i = 0; i <= qd; ++i) {
        double w = c[i + lvl];
#pragma unroll 1
        for (int t = 0; t < lvl; ++t) w *= (double)(i + lvl - t);
        q[i] = w;
      }
    }
    // bracket k = (crit[k-1], crit[k]] with crit[-1] = -bound, crit[ncrit] = +bound
    const double lo_s = __shfl_sync(full, myroot, (l16 > 0 ? l16 - 1 : 0), 16);
    const double hi_s = __shfl_sync(full, myroot, l16, 16);
    bool found = false;
    double root = 0.0;
    if (lvl_on && l16 <= ncrit) {
      double qr[11];   // register copy, zero above the degree
#pragma unroll
      for (int i = 0; i <= 10; ++i) qr[i] = (i <= qd) ? q[i] : 0.0;
      const double lo = (l16 == 0) ? -bound : lo_s;
      const double hi = (l16 < ncrit) ? hi_s : bound;
      double d0;
      const double flo = poly_eval10(qr, lo, d0), fhi = poly_eval10(qr, hi, d0);
      if (fhi == 0.0) {
        found = true;
        root = hi;
      } else if (flo != 0.0 && ((flo < 0.0) != (fhi < 0.0))) {
        found = true;
        root = bracket_root(qr, lo, hi, flo);
      }
    }
    const unsigned bal = (__ballot_sync(full, found) >> (16 * half)) & 0xffffu;
    const int nn = __popc(bal);
    // lane i takes the i-th found root (found roots are ascending in lane order)
    const int src = (l16 < nn) ? (int)__fns(bal, 0, l16 + 1) : 0;
    const double r = __shfl_sync(full, root, src, 16);
    if (lvl_on) {
      myroot = (l16 < nn) ? r : 0.0;
      ncrit = nn;
    }
  }
  if (have) {          // this group's Ehrlich-Aberth result stands (the other group of the warp needed the bracketing)
    myroot = root_keep;
    ncrit = ncrit_keep;
  }
  // lanes 0 .. ncrit-1 hold the real roots of det B(z): one essential matrix each
  bool valid = false;
  double E[9];
  if (live && l16 < ncrit) valid = e5_model_from_root(myroot, sc, sc + 36, E);
  const unsigned vb = (__ballot_sync(full, valid) >> (16 * half)) & 0xffffu;
  if (live) {
    const int pos = __popc(vb & ((1u << l16) - 1));
    if (valid && pos < 10) {
      double* out = models + (((long long)b * cap_iters + h) * 10 + pos) * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) out[i] = E[i];
    }
    if (l16 == 0) nmodels[(long long)b * cap_iters + h] = min(__popc(vb), 10);
  }
}

// ---- score ---------------------------------------------------------------------------------------
template <int MODEL>
__device__ __forceinline__ float model_error(const double* M, const float* Mf, const float2* p1, const float2* p2,
                                             const double2* q1, const double2* q2, long long i) {
  if (MODEL == MVO_MODEL_H) return h_error(Mf, p1[i], p2[i]);
  if (MODEL == MVO_MODEL_F) return f_error(M, p1[i], p2[i]);
  return e_error(M, q1[i], q2[i]);
}

constexpr int kScoreThreads = 128;
template <int MODEL>
__global__ void __launch_bounds__(kScoreThreads)
ransac_score_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, const double2* __restrict__ q1,
                    const double2* __restrict__ q2, const int32_t* __restrict__ npts, int max_pts,
                    const int32_t* __restrict__ state, int cap_iters, int h0, int h1,
                    const double* __restrict__ models, const int32_t* __restrict__ nmodels,
                    const float* __restrict__ thr2, int32_t* __restrict__ counts) {
  constexpr int MAXM = MT<MODEL>::MAXM;
  const int b = blockIdx.y;
  if (state[b * 8 + 3]) return;
  const int nsub = min(min(state[b * 8 + 1], cap_iters), h1);
  const int hbeg = max(h0, state[b * 8 + 4]);
  __shared__ double s_m[MAXM * 9];
  __shared__ float s_mf[MAXM * 9];
  __shared__ int s_cnt[MAXM];
  const int tid = threadIdx.x;
  const int n = npts[b];
  const float t = thr2[b];
  const long long base = (long long)b * max_pts;
  for (int h = hbeg + blockIdx.x; h < nsub; h += gridDim.x) {   // CTA per hypothesis, strided when the grid is capped
    const int nm = nmodels[(long long)b * cap_iters + h];
    if (tid < nm * 9) {
      const double v = models[((long long)b * cap_iters + h) * MAXM * 9 + tid];
      s_m[tid] = v;
      s_mf[tid] = (float)v;
    }
    if (tid < MAXM) s_cnt[tid] = 0;
    __syncthreads();
    for (int m = 0; m < nm; ++m) {
      int c = 0;
      for (int i = tid; i < n; i += kScoreThreads)
        c += (model_error<MODEL>(s_m + m * 9, s_mf + m * 9, p1, p2, q1, q2, base + i) <= t) ? 1 : 0;
      c = warp_sum(c);
      if ((tid & 31) == 0) atomicAdd(&s_cnt[m], c);
    }
    __syncthreads();
    if (tid < MAXM) counts[((long long)b * cap_iters + h) * MAXM + tid] = (tid < nm) ? s_cnt[tid] : -1;
    __syncthreads();
  }
}

// ---- select: replay of the sequential adaptive loop ---------------------------------------------
__host__ __device__ __forceinline__ int update_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.);
  p = fmin(p, 1.);
  ep = fmax(ep, 0.);
  ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// result per stream: [0] inlier count, [1] iterations run, [2] winning iteration, [3] winning model slot
template <int MODEL>
__global__ void __launch_bounds__(1024)
ransac_select_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, const double2* __restrict__ q1,
                     const double2* __restrict__ q2, const int32_t* __restrict__ npts, int max_pts,
                     int32_t* __restrict__ state, int cap_iters, int n_eval, const double* __restrict__ models,
                     const int32_t* __restrict__ counts, const float* __restrict__ thr2, double conf,
                     double* __restrict__ best_model, uint8_t* __restrict__ mask, int32_t* __restrict__ result) {
  constexpr int K = MT<MODEL>::K, MAXM = MT<MODEL>::MAXM, MAXIT = MT<MODEL>::MAXIT;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (state[b * 8 + 3]) return;   // an earlier round already produced the final answer
  const int n = npts[b];
  const int nsub = min(min(min(state[b * 8 + 1], cap_iters), MAXIT), n_eval);
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  __shared__ int s_T, s_bound;
  __shared__ unsigned long long s_win[2048 / 1024 + 1];
  __shared__ double s_m[9];
  __shared__ float s_mf[9];
  __shared__ int s_cnt;
  if (tid == 0) {
    s_carry = 0;
    s_T = 0x7fffffff;   // "the loop did not stop inside the evaluated range"
    s_bound = MAXIT;
    s_cnt = 0;
  }
  __syncthreads();
  // key: count << 24 | (0xFFF - iteration) << 12 ... first iteration / first model wins ties (strict '>')
  // pass A: inclusive prefix max over iterations (chunks of 1024), stop index T
  unsigned long long mykey[2] = {0, 0}, mypre[2] = {0, 0};
  for (int c = 0; c < 2; ++c) {
    const int it = c * 1024 + tid;
    unsigned long long key = 0;
    if (it < nsub) {
      int bestc = -1, bestm = 0;
      for (int m = 0; m < MAXM; ++m) {
        const int v = counts[((long long)b * cap_iters + it) * MAXM + m];
        if (v > bestc) {
          bestc = v;
          bestm = m;
        }
      }
      if (bestc > K - 1) key = ((unsigned long long)bestc << 32) | ((unsigned long long)(0xFFFF - it) << 8) | (unsigned)(0xFF - bestm);
    }
    mykey[c] = key;
    unsigned long long v = key;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, v, s);
      if (lane >= s) v = max(v, o);
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = s_warp[lane];
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, w, s);
        if (lane >= s) w = max(w, o);
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    unsigned long long pre = max(v, s_carry);
    if (warp > 0) pre = max(pre, s_warp[warp - 1]);
    mypre[c] = pre;
    __syncthreads();
    if (tid == 1023) s_carry = pre;
    __syncthreads();
  }
  // niters after iteration `it` = g(best count so far); iteration it+1 runs iff it+1 < that
  for (int c = 0; c < 2; ++c) {
    const int it = c * 1024 + tid;
    if (it < nsub) {
      int niters = MAXIT;
      if (mypre[c]) {
        const int cnt = (int)(mypre[c] >> 32);
        niters = update_iters(conf, (double)(n - cnt) / n, K, MAXIT);
      }
      if (it + 1 >= niters) atomicMin(&s_T, it + 1);
      if (it == nsub - 1) s_bound = niters;   // the bound the sequential loop carries into iteration nsub
    }
  }
  __syncthreads();
  // the loop stops by itself at s_T, or runs out of evaluated / drawable subsets at nsub
  const bool stopped = s_T <= nsub;
  const int T = stopped ? s_T : nsub;   // iterations actually run
  // final answer iff the loop stopped inside the range, all maxIters were evaluated, or no more subsets exist
  const bool final_round = stopped || n_eval >= MAXIT || (state[b * 8 + 2] & 6) != 0;
  __syncthreads();
  if (tid == 0) {
    if (final_round) state[b * 8 + 3] = 1;
    state[b * 8 + 4] = nsub;   // hypotheses [0, nsub) are solved and scored
    state[b * 8 + 5] = s_bound;
  }
  if (!final_round) return;   // the next round extends the range and selects again
  // winner = prefix max at iteration T-1
  for (int c = 0; c < 2; ++c) {
    const int it = c * 1024 + tid;
    if (it == T - 1) s_win[0] = mypre[c];
  }
  if (T == 0 && tid == 0) s_win[0] = 0;
  __syncthreads();
  const unsigned long long win = s_win[0];
  int32_t* res = result + b * 8;
  uint8_t* mk = mask + (long long)b * max_pts;
  if (win == 0) {
    for (int i = tid; i < n; i += 1024) mk[i] = 0;
    if (tid < 9) best_model[b * 9 + tid] = 0.0;
    if (tid == 0) {
      res[0] = 0;
      res[1] = T;
      res[2] = -1;
      res[3] = -1;
    }
    return;
  }
  const int wit = 0xFFFF - (int)((win >> 8) & 0xFFFF), wm = 0xFF - (int)(win & 0xFF);
  if (tid < 9) {
    const double v = models[(((long long)b * cap_iters + wit) * MAXM + wm) * 9 + tid];
    s_m[tid] = v;
    s_mf[tid] = (float)v;
    best_model[b * 9 + tid] = v;
  }
  __syncthreads();
  const float t = thr2[b];
  const long long base = (long long)b * max_pts;
  int c = 0;
  for (int i = tid; i < n; i += 1024) {
    const int in = (model_error<MODEL>(s_m, s_mf, p1, p2, q1, q2, base + i) <= t) ? 1 : 0;
    mk[i] = (uint8_t)in;
    c += in;
  }
  c = warp_sum(c);
  if (lane == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (tid == 0) {
    res[0] = s_cnt;
    res[1] = T;
    res[2] = wit;
    res[3] = wm;
  }
}

// ---- LMedS: what cv::findFundamentalMat(FM_RANSAC) silently runs for 8 <= N < 15 -------------------------
// (LMeDSPointSetRegistrator::run; the Tracker gets there with min_tracked_points = 10: /root/reference/src/tracker.cpp:
// 239-248.)  Same subsets as RANSAC; thread = iteration: median (element N/2 of the sorted errors) of each of its
// models, block-wide minimum with the sequential loop's tie rule (strict '<': first iteration, first model), then
// sigma = 2.5 * 1.4826 * (1 + 5 / (N - 7)) * sqrt(median) >= 0.001 and mask = err <= sigma^2.
__global__ void __launch_bounds__(1024)
f_lmeds_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, const int32_t* __restrict__ npts, int max_pts,
               const int32_t* __restrict__ state, int cap_iters, int niters, const double* __restrict__ models,
               const int32_t* __restrict__ nmodels, double* __restrict__ best_model, uint8_t* __restrict__ mask,
               int32_t* __restrict__ result) {
  constexpr int MAXM = MT<MVO_MODEL_F>::MAXM, NMAX = 16;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(npts[b], NMAX);
  const int nsub = min(min(state[b * 8 + 1], cap_iters), niters);
  const float2* a = p1 + (long long)b * max_pts;
  const float2* c = p2 + (long long)b * max_pts;
  __shared__ unsigned long long s_warp[32];
  __shared__ double s_m[9];
  __shared__ int s_cnt;
  unsigned long long key = ~0ull;
  if (tid < nsub) {
    const int nm = nmodels[(long long)b * cap_iters + tid];
    for (int m = 0; m < nm; ++m) {
      const double* M = models + (((long long)b * cap_iters + tid) * MAXM + m) * 9;
      float e[NMAX];
#pragma unroll
      for (int i = 0; i < NMAX; ++i) e[i] = i < n ? f_error(M, a[i], c[i]) : __int_as_float(0x7f800000);
      // N/2-th smallest: count-based selection keeps e[] in registers (no data-dependent indexing)
      float med = __int_as_float(0x7f800000);
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        int less = 0, leq = 0;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
          less += e[j] < e[i] ? 1 : 0;
          leq += e[j] <= e[i] ? 1 : 0;
        }
        if (i < n && less <= n / 2 && n / 2 < leq) med = e[i];
      }
      const unsigned bits = __float_as_uint(med);
      if (med >= 0.f && bits < 0x7f800000u) {   // finite (a NaN / inf median never beats DBL_MAX in OpenCV either)
        const unsigned long long k2 = ((unsigned long long)bits << 32) | ((unsigned long long)tid << 8) | (unsigned)m;
        key = min(key, k2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
  if (lane == 0) s_warp[warp] = key;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (warp == 0) {
    key = s_warp[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    if (lane == 0) s_warp[0] = key;
  }
  __syncthreads();
  key = s_warp[0];
  int32_t* res = result + b * 8;
  uint8_t* mk = mask + (long long)b * max_pts;
  if (key == ~0ull) {
    if (tid < n) mk[tid] = 0;
    if (tid < 9) best_model[b * 9 + tid] = 0.0;
    if (tid == 0) {
      res[0] = 0;
      res[1] = nsub;
      res[2] = -1;
      res[3] = -1;
    }
    return;
  }
  const int wit = (int)((key >> 8) & 0xFFFFFF), wm = (int)(key & 0xFF);
  if (tid < 9) {
    const double v = models[(((long long)b * cap_iters + wit) * MAXM + wm) * 9 + tid];
    s_m[tid] = v;
    best_model[b * 9 + tid] = v;
  }
  __syncthreads();
  const double median = (double)__uint_as_float((unsigned)(key >> 32));
  double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(median);
  sigma = fmax(sigma, 0.001);
  const float t = (float)(sigma * sigma);
  if (tid < n) {
    const int in = f_error(s_m, a[tid], c[tid]) <= t ? 1 : 0;
    mk[tid] = (uint8_t)in;
    if (in) atomicAdd(&s_cnt, 1);
  }
  __syncthreads();
  if (tid == 0) {
    res[0] = s_cnt;
    res[1] = nsub;
    res[2] = wit;
    res[3] = wm;
  }
}

// ---- homography refinement: DLT on inliers + cv::LMSolver (10 iterations) + mask of the refined H ----
constexpr int kRefThreads = 256;

template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* s_red /* NV * 8 */, int tid) {
  // deterministic: fixed per-thread strides, warp tree, then warps summed in order
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum_d(v[k]);
  __syncthreads();
  if ((tid & 31) == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) s_red[(tid >> 5) * NV + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kRefThreads / 32; ++w) s += s_red[w * NV + k];
    v[k] = s;
  }
}

__device__ __forceinline__ void h_residual(const double* h, double Mx, double My, double mx, double my, double& ww,
                                           double& xi, double& yi, double& ex, double& ey) {
  ww = h[6] * Mx + h[7] * My + 1.;
  ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
  xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
  yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
  ex = xi - mx;
  ey = yi - my;
}

__global__ void __launch_bounds__(kRefThreads)
h_refine_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, const int32_t* __restrict__ npts,
                int max_pts, const float* __restrict__ thr2, double* __restrict__ best_model,
                uint8_t* __restrict__ mask, int32_t* __restrict__ result, int32_t* __restrict__ inl_idx) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = npts[b];
  int32_t* res = result + b * 8;
  if (res[0] <= 0) return;   // no model
  const float2* P1 = p1 + (long long)b * max_pts;
  const float2* P2 = p2 + (long long)b * max_pts;
  uint8_t* mk = mask + (long long)b * max_pts;
  int32_t* idx = inl_idx + (long long)b * max_pts;
  __shared__ double s_red[(kRefThreads / 32) * 44];
  __shared__ int s_warp[kRefThreads / 32];
  __shared__ int s_base;
  __shared__ double s_x[8], s_xd[8], s_A[64], s_v[8], s_ctl[8];
  __shared__ float s_hf[9];
  __shared__ int s_cnt;
  // ordered compaction of the inlier indices
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += kRefThreads) {
    const int i = i0 + tid;
    const int ok = (i < n) ? (mk[i] != 0) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (ok) idx[off + __popc(bal & ((1u << lane) - 1))] = i;
    __syncthreads();
    if (tid == 0)
      for (int w = 0; w < kRefThreads / 32; ++w) s_base += s_warp[w];
    __syncthreads();
  }
  const int ni = s_base;
  if (ni < 4) return;
  // ---- DLT (HomographyEstimatorCallback::runKernel on the inliers) ----
  double v4[4] = {0, 0, 0, 0};
  for (int k = tid; k < ni; k += kRefThreads) {
    const float2 M = P1[idx[k]], m = P2[idx[k]];
    v4[0] += m.x; v4[1] += m.y; v4[2] += M.x; v4[3] += M.y;
  }
  block_sum<4>(v4, s_red, tid);
  const double cmx = v4[0] / ni, cmy = v4[1] / ni, cMx = v4[2] / ni, cMy = v4[3] / ni;
  double d4[4] = {0, 0, 0, 0};
  for (int k = tid; k < ni; k += kRefThreads) {
    const float2 M = P1[idx[k]], m = P2[idx[k]];
    d4[0] += fabs(m.x - cmx); d4[1] += fabs(m.y - cmy); d4[2] += fabs(M.x - cMx); d4[3] += fabs(M.y - cMy);
  }
  block_sum<4>(d4, s_red, tid);
  if (fabs(d4[0]) < DBL_EPSILON || fabs(d4[1]) < DBL_EPSILON || fabs(d4[2]) < DBL_EPSILON || fabs(d4[3]) < DBL_EPSILON) return;
  const double smx = ni / d4[0], smy = ni / d4[1], sMx = ni / d4[2], sMy = ni / d4[3];
  // LtL blocks: S = sum a a^T, Sx = sum x a a^T, Sy = sum y a a^T, Sr = sum (x^2+y^2) a a^T with a = (X, Y, 1)
  double acc[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) acc[k] = 0;
  for (int k = tid; k < ni; k += kRefThreads) {
    const float2 Mp = P1[idx[k]], mp = P2[idx[k]];
    const double x = (mp.x - cmx) * smx, y = (mp.y - cmy) * smy;
    const double X = (Mp.x - cMx) * sMx, Y = (Mp.y - cMy) * sMy;
    const double aa[6] = {X * X, X * Y, X, Y * Y, Y, 1.0};
    const double r = x * x + y * y;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      acc[j] += aa[j];
      acc[6 + j] += x * aa[j];
      acc[12 + j] += y * aa[j];
      acc[18 + j] += r * aa[j];
    }
  }
  block_sum<24>(acc, s_red, tid);
  if (tid == 0) {
    double L[81];
    for (int i = 0; i < 81; ++i) L[i] = 0;
    const int sidx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const int s = sidx[i][j];
        L[i * 9 + j] = acc[s];
        L[(3 + i) * 9 + 3 + j] = acc[s];
        L[i * 9 + 6 + j] = -acc[6 + s];
        L[(6 + j) * 9 + i] = -acc[6 + s];
        L[(3 + i) * 9 + 6 + j] = -acc[12 + s];
        L[(6 + j) * 9 + 3 + i] = -acc[12 + s];
        L[(6 + i) * 9 + 6 + j] = acc[18 + s];
      }
    // eigenvector of the smallest eigenvalue (inverse iteration; the DLT null direction)
    double h0[9];
    smallest_eigvec_spd<9>(L, h0);
    const double inv_hn[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
    const double hn2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double t[9], H[9];
    mat3_mul(inv_hn, h0, t);
    mat3_mul(t, hn2, H);
    const double s = 1. / H[8];
    for (int i = 0; i < 8; ++i) s_x[i] = H[i] * s;
  }
  __syncthreads();
  // ---- LM (cv::LMSolverImpl::run, maxIters 10, eps FLT_EPSILON) ----
  // initial residual / normal equations
  double lam = 1.0, lc = 0.75, S = 0;
  double A[36], vv[8];   // thread 0 keeps the reduced normal matrix (upper triangle) and J^T r
  auto normal_eq = [&](const double* h) {
    double a44[45];
#pragma unroll
    for (int k = 0; k < 45; ++k) a44[k] = 0;
    for (int k = tid; k < ni; k += kRefThreads) {
      const float2 Mp = P1[idx[k]], mp = P2[idx[k]];
      const double Mx = Mp.x, My = Mp.y;
      double ww, xi, yi, ex, ey;
      h_residual(h, Mx, My, (double)mp.x, (double)mp.y, ww, xi, yi, ex, ey);
      const double jx[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
      const double jy[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
      int o = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = i; j < 8; ++j) a44[o++] += jx[i] * jx[j] + jy[i] * jy[j];
#pragma unroll
      for (int i = 0; i < 8; ++i) a44[36 + i] += jx[i] * ex + jy[i] * ey;
      a44[44] += ex * ex + ey * ey;
    }
    // 45 values: reduce in two halves to bound registers
    block_sum<24>(a44, s_red, tid);
    block_sum<21>(a44 + 24, s_red, tid);
    if (tid == 0) {
      for (int k = 0; k < 36; ++k) A[k] = a44[k];
      for (int k = 0; k < 8; ++k) vv[k] = a44[36 + k];
      S = a44[44];
    }
  };
  auto residual_only = [&](const double* h, double& Sd, double& rinf) {
    double two[2] = {0, 0};
    for (int k = tid; k < ni; k += kRefThreads) {
      const float2 Mp = P1[idx[k]], mp = P2[idx[k]];
      double ww, xi, yi, ex, ey;
      h_residual(h, (double)Mp.x, (double)Mp.y, (double)mp.x, (double)mp.y, ww, xi, yi, ex, ey);
      two[0] += ex * ex + ey * ey;
      two[1] = fmax(two[1], fmax(fabs(ex), fabs(ey)));
    }
    // sum for [0], max for [1]
    double mx = two[1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double one[1] = {two[0]};
    block_sum<1>(one, s_red, tid);
    __syncthreads();
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    double m = 0;
    for (int w = 0; w < kRefThreads / 32; ++w) m = fmax(m, s_red[w]);
    __syncthreads();
    Sd = one[0];
    rinf = m;
  };
  normal_eq(s_x);
  double Dg[8];
  if (tid == 0) {
    int o = 0;
    for (int i = 0; i < 8; ++i) {
      Dg[i] = A[o];
      o += 8 - i;
    }
  }
  double rinf_cur;
  {
    double tmpS;
    residual_only(s_x, tmpS, rinf_cur);
  }
  for (int iter = 0;;) {
    // thread 0: solve (A + lam*D) d = v
    if (tid == 0) {
      double Ap[64], d[8];
      int o = 0;
      for (int i = 0; i < 8; ++i)
        for (int j = i; j < 8; ++j) {
          Ap[i * 8 + j] = Ap[j * 8 + i] = A[o++];
        }
      for (int i = 0; i < 64; ++i) s_A[i] = Ap[i];
      for (int i = 0; i < 8; ++i) Ap[i * 9] += lam * Dg[i];
      for (int i = 0; i < 8; ++i) d[i] = vv[i];
      if (!lu_solve<8>(Ap, d))
        for (int i = 0; i < 8; ++i) d[i] = 0;
      for (int i = 0; i < 8; ++i) {
        s_v[i] = d[i];
        s_xd[i] = s_x[i] - d[i];
      }
    }
    __syncthreads();
    double Sd, rinf_d;
    residual_only(s_xd, Sd, rinf_d);
    if (tid == 0) {
      // dS = d . (2 v - A d);  R = (S - Sd) / dS
      double dS = 0, tdot = 0, dinf = 0;
      for (int i = 0; i < 8; ++i) {
        double ad = 0;
        for (int j = 0; j < 8; ++j) ad += s_A[i * 8 + j] * s_v[j];
        dS += s_v[i] * (2 * vv[i] - ad);
        tdot += s_v[i] * vv[i];
        dinf = fmax(dinf, fabs(s_v[i]));
      }
      const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
      if (R > 0.75) {
        lam *= 0.5;
        if (lam < lc) lam = 0;
      } else if (R < 0.25) {
        double nu = (Sd - S) / (fabs(tdot) > DBL_EPSILON ? tdot : 1) + 2;
        nu = fmin(fmax(nu, 2.), 10.);
        if (lam == 0) {
          // lc = 1 / max |diag(A^-1)|
          double maxval = DBL_EPSILON;
          for (int c = 0; c < 8; ++c) {
            double Ap[64], e[8];
            for (int i = 0; i < 64; ++i) Ap[i] = s_A[i];
            for (int i = 0; i < 8; ++i) e[i] = (i == c) ? 1.0 : 0.0;
            if (lu_solve<8>(Ap, e)) maxval = fmax(maxval, fabs(e[c]));
          }
          lam = lc = 1. / maxval;
          nu *= 0.5;
        }
        lam *= nu;
      }
      s_ctl[0] = (Sd < S) ? 1.0 : 0.0;
      s_ctl[1] = dinf;
    }
    __syncthreads();
    const bool accept = s_ctl[0] != 0.0;
    const double dinf = s_ctl[1];
    __syncthreads();
    if (accept) {
      if (tid < 8) s_x[tid] = s_xd[tid];
      __syncthreads();
      normal_eq(s_x);
      rinf_cur = rinf_d;
    }
    ++iter;
    const bool proceed = iter < 10 && dinf >= (double)FLT_EPSILON && rinf_cur >= (double)FLT_EPSILON;
    if (!proceed) break;
  }
  __syncthreads();
  // ---- final model + mask of the refined H ----
  if (tid < 8) {
    best_model[b * 9 + tid] = s_x[tid];
    s_hf[tid] = (float)s_x[tid];
  }
  if (tid == 8) {
    best_model[b * 9 + 8] = 1.0;
    s_hf[8] = 1.f;
    s_cnt = 0;
  }
  __syncthreads();
  const float t = thr2[b];
  int c = 0;
  for (int i = tid; i < n; i += kRefThreads) {
    const int in = (h_error(s_hf, P1[i], P2[i]) <= t) ? 1 : 0;
    mk[i] = (uint8_t)in;
    c += in;
  }
  c = warp_sum(c);
  if (lane == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (tid == 0) res[0] = s_cnt;
}

// ---- homography refinement, second generation ---------------------------------------------------------
// Same algorithm and (element for element) the same arithmetic as h_refine_kernel above, re-organised for latency -- the
// kernel is one CTA per stream and sits on the critical path of a single-stream frame:
//   * the ordered inlier correspondences are compacted into shared memory once (every pass of the old kernel re-read
//     idx[k] -> P1[idx[k]], P2[idx[k]] from global memory: two dependent L2 round trips per pass, 31 % of its time);
//   * the serial pieces that ran on thread 0 with local-memory arrays (9 x 9 LU + inverse iteration for the DLT null
//     vector, the 8 x 8 LU of every Levenberg-Marquardt step: 35 % of the time with 255 threads waiting) run on warp 0 with
//     the matrix in shared memory: pivot search redundantly in every lane, row swap and elimination entries over lanes;
//   * one pass per LM iteration: the residual at the trial point and -- speculatively -- the normal equations there are
//     accumulated together (the old kernel made a second pass after accepting the step); the cross-warp sums are formed
//     by warp 0 only, in the same warp order.
constexpr int kRefCache = 2048;   // inlier correspondences kept in shared memory (32 KB); beyond that: global gather

// warp tree (xor butterfly) per value, lane 0 of every warp -> s_part[warp][NV]; the caller synchronises
template <int NV>
__device__ __forceinline__ void warp_partials(double* v, double* s_part, int lane, int warp) {
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum_d(v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) s_part[warp * NV + k] = v[k];
}
// every thread forms the block totals from the per-warp partials, warps in order (as block_sum)
template <int NV>
__device__ __forceinline__ void block_totals(double* v, const double* s_part) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kRefThreads / 32; ++w) t += s_part[w * NV + k];
    v[k] = t;
  }
}

__global__ void __launch_bounds__(kRefThreads)
h_refine2_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, const int32_t* __restrict__ npts,
                 int max_pts, const float* __restrict__ thr2, double* __restrict__ best_model,
                 uint8_t* __restrict__ mask, int32_t* __restrict__ result, int32_t* __restrict__ inl_idx) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kRefThreads / 32;
  const int n = npts[b];
  int32_t* res = result + b * 8;
  if (res[0] <= 0) return;   // no model
  const float2* P1 = p1 + (long long)b * max_pts;
  const float2* P2 = p2 + (long long)b * max_pts;
  uint8_t* mk = mask + (long long)b * max_pts;
  int32_t* idx = inl_idx + (long long)b * max_pts;
  __shared__ float4 s_pts[kRefCache];
  __shared__ double s_part[kWarps * 47];
  __shared__ double s_L[81];
  __shared__ double s_A[64], s_Ap[64], s_d[8], s_vv[8], s_Dg[8], s_x[8], s_xd[8], s_v[8], s_ad[8], s_h0[9];
  __shared__ double s_S, s_lam, s_lc, s_rinf;
  __shared__ int s_warp[kWarps];
  __shared__ int s_flag[2];
  __shared__ float s_hf[9];
  __shared__ int s_cnt;
  // ---- ordered compaction of the inliers: indices to global memory, correspondences to shared memory ----
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += kRefThreads) {
    const int i = i0 + tid;
    const int ok = (i < n) ? (mk[i] != 0) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int cw = s_warp[w];
      if (w < warp) off += cw;
      tot += cw;
    }
    if (ok) {
      const int k = off + __popc(bal & ((1u << lane) - 1));
      idx[k] = i;
      if (k < kRefCache) {
        const float2 M = P1[i], m = P2[i];
        s_pts[k] = make_float4(M.x, M.y, m.x, m.y);
      }
    }
    base += tot;
    __syncthreads();
  }
  const int ni = base;
  if (ni < 4) return;
  auto pt = [&](int k) -> float4 {
    if (k < kRefCache) return s_pts[k];
    const float2 M = P1[idx[k]], m = P2[idx[k]];
    return make_float4(M.x, M.y, m.x, m.y);
  };
  // ---- DLT (HomographyEstimatorCallback::runKernel on the inliers) ----
  double v4[4] = {0, 0, 0, 0};
  for (int k = tid; k < ni; k += kRefThreads) {
    const float4 q = pt(k);
    v4[0] += q.z; v4[1] += q.w; v4[2] += q.x; v4[3] += q.y;
  }
  warp_partials<4>(v4, s_part, lane, warp);
  __syncthreads();
  block_totals<4>(v4, s_part);
  __syncthreads();
  const double cmx = v4[0] / ni, cmy = v4[1] / ni, cMx = v4[2] / ni, cMy = v4[3] / ni;
  double d4[4] = {0, 0, 0, 0};
  for (int k = tid; k < ni; k += kRefThreads) {
    const float4 q = pt(k);
    d4[0] += fabs(q.z - cmx); d4[1] += fabs(q.w - cmy); d4[2] += fabs(q.x - cMx); d4[3] += fabs(q.y - cMy);
  }
  warp_partials<4>(d4, s_part, lane, warp);
  __syncthreads();
  block_totals<4>(d4, s_part);
  __syncthreads();
  if (fabs(d4[0]) < DBL_EPSILON || fabs(d4[1]) < DBL_EPSILON || fabs(d4[2]) < DBL_EPSILON || fabs(d4[3]) < DBL_EPSILON) return;
  const double smx = ni / d4[0], smy = ni / d4[1], sMx = ni / d4[2], sMy = ni / d4[3];
  {
    // LtL blocks: S = sum a a^T, Sx = sum x a a^T, Sy = sum y a a^T, Sr = sum (x^2+y^2) a a^T with a = (X, Y, 1)
    double acc[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) acc[k] = 0;
    for (int k = tid; k < ni; k += kRefThreads) {
      const float4 q = pt(k);
      const double x = (q.z - cmx) * smx, y = (q.w - cmy) * smy;
      const double X = (q.x - cMx) * sMx, Y = (q.y - cMy) * sMy;
      const double aa[6] = {X * X, X * Y, X, Y * Y, Y, 1.0};
      const double r = x * x + y * y;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        acc[j] += aa[j];
        acc[6 + j] += x * aa[j];
        acc[12 + j] += y * aa[j];
        acc[18 + j] += r * aa[j];
      }
    }
    warp_partials<24>(acc, s_part, lane, warp);
    __syncthreads();
    if (warp == 0) {
      // lane k < 24 owns total k; the 9 x 9 normal matrix is scattered from them
      double t = 0;
      if (lane < 24)
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_part[w * 24 + lane];
      for (int e = lane; e < 81; e += 32) s_L[e] = 0;
      __syncwarp();
      if (lane < 24) {
        const int blk = lane / 6, s6 = lane % 6;
        // symmetric 3 x 3 index pairs of packed entry s6: {0:(0,0), 1:(0,1), 2:(0,2), 3:(1,1), 4:(1,2), 5:(2,2)}
        const int pi = s6 < 3 ? 0 : (s6 < 5 ? 1 : 2);
        const int pj = s6 < 3 ? s6 : (s6 < 5 ? s6 - 2 : 2);
#pragma unroll
        for (int sw = 0; sw < 2; ++sw) {
          const int i = sw ? pj : pi, j = sw ? pi : pj;
          if (sw && pi == pj) break;
          if (blk == 0) {
            s_L[i * 9 + j] = t;
            s_L[(3 + i) * 9 + 3 + j] = t;
          } else if (blk == 1) {
            s_L[i * 9 + 6 + j] = -t;
            s_L[(6 + j) * 9 + i] = -t;
          } else if (blk == 2) {
            s_L[(3 + i) * 9 + 6 + j] = -t;
            s_L[(6 + j) * 9 + 3 + i] = -t;
          } else {
            s_L[(6 + i) * 9 + 6 + j] = t;
          }
        }
      }
      __syncwarp();
      smallest_eigvec_spd_warp<9>(s_L, s_h0, lane);   // the DLT null direction (inverse iteration)
      if (lane == 0) {
        double h0[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) h0[i] = s_h0[i];
        const double inv_hn[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
        const double hn2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
        double t9[9], H[9];
        mat3_mul(inv_hn, h0, t9);
        mat3_mul(t9, hn2, H);
        const double sc = 1. / H[8];
        for (int i = 0; i < 8; ++i) s_x[i] = H[i] * sc;
        s_lam = 1.0;
        s_lc = 0.75;
      }
    }
    __syncthreads();
  }
  // ---- LM (cv::LMSolverImpl::run, maxIters 10, eps FLT_EPSILON) ----
  // one pass: residual sum / max and the normal equations J^T J (upper triangle), J^T r at h
  auto pass = [&](const double* h) {
    double a47[47];
#pragma unroll
    for (int k = 0; k < 47; ++k) a47[k] = 0;
    const double h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3], h4 = h[4], h5 = h[5], h6 = h[6], h7 = h[7];
    for (int k = tid; k < ni; k += kRefThreads) {
      const float4 q = pt(k);
      const double Mx = q.x, My = q.y;
      double ww = h6 * Mx + h7 * My + 1.;
      ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
      const double xi = (h0 * Mx + h1 * My + h2) * ww;
      const double yi = (h3 * Mx + h4 * My + h5) * ww;
      const double ex = xi - (double)q.z, ey = yi - (double)q.w;
      const double jx[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
      const double jy[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
      int o = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = i; j < 8; ++j) a47[o++] += jx[i] * jx[j] + jy[i] * jy[j];
#pragma unroll
      for (int i = 0; i < 8; ++i) a47[36 + i] += jx[i] * ex + jy[i] * ey;
      a47[44] += ex * ex + ey * ey;
      a47[45] = fmax(a47[45], fmax(fabs(ex), fabs(ey)));
    }
#pragma unroll
    for (int k = 0; k < 45; ++k) a47[k] = warp_sum_d(a47[k]);
    double mx = a47[45];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 45; ++k) s_part[warp * 47 + k] = a47[k];
      s_part[warp * 47 + 45] = mx;
    }
  };
  // warp 0: block totals of a pass; lane k owns totals k and k + 32 (sum over the warps in order; entry 45 is a max)
  auto totals = [&](double& t0, double& t1) {
    t0 = 0;
    t1 = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      t0 += s_part[w * 47 + lane];
      if (lane + 32 < 45) t1 += s_part[w * 47 + lane + 32];
      if (lane + 32 == 45) t1 = fmax(t1, s_part[w * 47 + 45]);
    }
  };
  // scatter the totals of a pass into the normal matrix (full symmetric), J^T r, S, |r|_inf
  auto commit = [&](double t0, double t1) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int e = lane + 32 * half;
      const double t = half ? t1 : t0;
      if (e < 36) {
        int i = 0, o = e;
        while (o >= 8 - i) {
          o -= 8 - i;
          ++i;
        }
        const int j = i + o;
        s_A[i * 8 + j] = t;
        s_A[j * 8 + i] = t;
      } else if (e < 44) {
        s_vv[e - 36] = t;
      } else if (e == 44) {
        s_S = t;
      } else if (e == 45) {
        s_rinf = t;
      }
    }
  };
  pass(s_x);
  __syncthreads();
  if (warp == 0) {
    double t0, t1;
    totals(t0, t1);
    commit(t0, t1);
    __syncwarp();
    if (lane < 8) s_Dg[lane] = s_A[lane * 9];
    __syncwarp();
  }
  for (int iter = 0;;) {
    if (warp == 0) {
      // solve (A + lam * D) d = v
      const double lam = s_lam;
      for (int e = lane; e < 64; e += 32) s_Ap[e] = s_A[e];
      __syncwarp();
      if (lane < 8) {
        s_Ap[lane * 9] += lam * s_Dg[lane];      // one fused multiply-add, as in the first-generation kernel
        s_d[lane] = s_vv[lane];
      }
      __syncwarp();
      const bool ok = lu_solve_warp<8>(s_Ap, s_d, lane);
      if (lane < 8) {
        const double d = ok ? s_d[lane] : 0.0;
        s_v[lane] = d;
        s_xd[lane] = s_x[lane] - d;
      }
    }
    __syncthreads();
    pass(s_xd);
    __syncthreads();
    if (warp == 0) {
      double t0, t1;
      totals(t0, t1);
      const double Sd = __shfl_sync(0xffffffffu, t1, 44 - 32), rinf_d = __shfl_sync(0xffffffffu, t1, 45 - 32);
      if (lane < 8) {
        double ad = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) ad += s_A[lane * 8 + j] * s_v[j];
        s_ad[lane] = ad;
      }
      __syncwarp();
      // dS = d . (2 v - A d);  R = (S - Sd) / dS   (every lane, redundantly: uniform control flow for the warp solves)
      const double S = s_S;
      double dS = 0, tdot = 0, dinf = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dS += s_v[i] * (2 * s_vv[i] - s_ad[i]);
        tdot += s_v[i] * s_vv[i];
        dinf = fmax(dinf, fabs(s_v[i]));
      }
      double lam = s_lam, lc = s_lc;
      const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
      if (R > 0.75) {
        lam *= 0.5;
        if (lam < lc) lam = 0;
      } else if (R < 0.25) {
        double nu = (Sd - S) / (fabs(tdot) > DBL_EPSILON ? tdot : 1) + 2;
        nu = fmin(fmax(nu, 2.), 10.);
        if (lam == 0) {
          // lc = 1 / max |diag(A^-1)|
          double maxval = DBL_EPSILON;
          for (int c = 0; c < 8; ++c) {
            __syncwarp();
            for (int e = lane; e < 64; e += 32) s_Ap[e] = s_A[e];
            if (lane < 8) s_d[lane] = (lane == c) ? 1.0 : 0.0;
            __syncwarp();
            if (lu_solve_warp<8>(s_Ap, s_d, lane)) maxval = fmax(maxval, fabs(s_d[c]));
          }
          lam = lc = 1. / maxval;
          nu *= 0.5;
        }
        lam *= nu;
      }
      __syncwarp();
      const bool accept = Sd < S;
      if (accept) {
        if (lane < 8) s_x[lane] = s_xd[lane];
        commit(t0, t1);                      // the trial point's normal equations, S = Sd, |r|_inf
      }
      if (lane == 0) {
        s_lam = lam;
        s_lc = lc;
        const double rinf_cur = accept ? rinf_d : s_rinf;
        s_flag[0] = (iter + 1 < 10 && dinf >= (double)FLT_EPSILON && rinf_cur >= (double)FLT_EPSILON) ? 1 : 0;
      }
    }
    __syncthreads();
    ++iter;
    if (!s_flag[0]) break;
  }
  // ---- final model + mask of the refined H ----
  if (tid < 8) {
    best_model[b * 9 + tid] = s_x[tid];
    s_hf[tid] = (float)s_x[tid];
  }
  if (tid == 8) {
    best_model[b * 9 + 8] = 1.0;
    s_hf[8] = 1.f;
    s_cnt = 0;
  }
  __syncthreads();
  const float t = thr2[b];
  int c = 0;
  for (int i = tid; i < n; i += kRefThreads) {
    const int in = (h_error(s_hf, P1[i], P2[i]) <= t) ? 1 : 0;
    mk[i] = (uint8_t)in;
    c += in;
  }
  c = warp_sum(c);
  if (lane == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (tid == 0) res[0] = s_cnt;
}

// ---- helpers ---------------------------------------------------------------------------------------
__global__ void normalize_points_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2,
                                        const int32_t* __restrict__ npts, int max_pts, const double* __restrict__ K,
                                        double2* __restrict__ q1, double2* __restrict__ q2) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts[b]) return;
  const double fx = K[b * 9 + 0], fy = K[b * 9 + 4], cx = K[b * 9 + 2], cy = K[b * 9 + 5];
  const long long o = (long long)b * max_pts + i;
  q1[o] = make_double2(((double)p1[o].x - cx) / fx, ((double)p1[o].y - cy) / fy);
  q2[o] = make_double2(((double)p2[o].x - cx) / fx, ((double)p2[o].y - cy) / fy);
}

// ====================================================================================================
// host side
static void fill_rng_table(uint64_t seed, std::vector<uint32_t>& t, size_t count) {
  t.resize(count);
  uint64_t s = seed;
  for (size_t i = 0; i < count; ++i) {
    s = (uint64_t)(uint32_t)s * 4164903690ull + (s >> 32);
    t[i] = (uint32_t)s;
  }
}

static size_t chain_smem(int window, int attempts) {
  return (size_t)(window + 2) * 2 * 2 + (size_t)attempts * 2 + (size_t)kSegIters * 2 + (size_t)window + 16;
}

int ransac_prepare(mvo_ctx* c, int max_pts, int cap_iters, int lanes) {
  RansacBufs& r = c->rs;
  const size_t B = (size_t)c->cfg.batch;
  if (!r.rng_ready) {
    std::vector<uint32_t> t;
    r.rng_len = 1 << 19;   // 524288 draws: enough for 16384-hypothesis sweeps
    fill_rng_table(c->cfg.ransac_seed, t, (size_t)r.rng_len);
    MVO_CUDA_TRY(c, r.rng.alloc(t.size()));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(r.rng.p, t.data(), t.size() * 4, cudaMemcpyHostToDevice, c->stream));
    MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const int smem = (int)chain_smem(kDraws, kMaxAttempts);
    MVO_CUDA_TRY(c, cudaFuncSetAttribute(ransac_chain_kernel<MVO_MODEL_H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MVO_CUDA_TRY(c, cudaFuncSetAttribute(ransac_chain_kernel<MVO_MODEL_F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MVO_CUDA_TRY(c, cudaFuncSetAttribute(ransac_chain_kernel<MVO_MODEL_E>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    r.rng_ready = true;
  }
  if (max_pts > r.max_pts) {
    const size_t n = B * (size_t)max_pts;
    MVO_CUDA_TRY(c, r.p1.alloc(n));
    MVO_CUDA_TRY(c, r.p2.alloc(n));
    MVO_CUDA_TRY(c, r.q1.alloc(n));
    MVO_CUDA_TRY(c, r.q2.alloc(n));
    r.max_pts = max_pts;
  }
  r.cap_iters = std::max(r.cap_iters, cap_iters);
  MVO_CUDA_TRY(c, r.npts.alloc(B));
  MVO_CUDA_TRY(c, r.K.alloc(B * 9));
  for (int l = 0; l < std::min(std::max(lanes, 1), RansacBufs::kLanes); ++l) {
    RansacLane& ln = r.lane[l];
    if (!ln.ready) {
      MVO_CUDA_TRY(c, ln.att_next.alloc(B * kDraws));
      MVO_CUDA_TRY(c, ln.att_ok.alloc(B * kDraws));
      MVO_CUDA_TRY(c, ln.state.alloc(B * 8));
      MVO_CUDA_TRY(c, ln.thr2.alloc(B));
      MVO_CUDA_TRY(c, ln.best_model.alloc(B * 9));
      MVO_CUDA_TRY(c, ln.result.alloc(B * 8));
      ln.ready = true;
    }
    if (r.max_pts > ln.max_pts) {
      const size_t n = B * (size_t)r.max_pts;
      MVO_CUDA_TRY(c, ln.mask.alloc(n));
      MVO_CUDA_TRY(c, ln.inl_idx.alloc(n));
      ln.max_pts = r.max_pts;
    }
    if (r.cap_iters > ln.cap_iters) {
      const size_t n = B * (size_t)r.cap_iters;
      MVO_CUDA_TRY(c, ln.subsets.alloc(n * 8));
      MVO_CUDA_TRY(c, ln.models.alloc(n * kMaxHypModels * 9));
      MVO_CUDA_TRY(c, ln.nmodels.alloc(n));
      MVO_CUDA_TRY(c, ln.counts.alloc(n * kMaxHypModels));
      MVO_CUDA_TRY(c, ln.e5_scratch.alloc(n * kE5Scratch));
      ln.cap_iters = r.cap_iters;
    }
  }
  return MVO_OK;
}

// one sampler pass: up to kSegIters more subsets (until `want_total`), from an RNG window of `window` draws
template <int MODEL>
static void sample_pass(mvo_ctx* c, int want_total, int window, int attempts) {
  RansacBufs& r = c->rs;
  const int B = c->cfg.batch;
  dim3 ga((window + 255) / 256, B);
  ransac_attempt_kernel<MODEL><<<ga, 256, 0, c->stream>>>(r.rng.p, r.rng_len, r.p1.p, r.p2.p, r.npts.p, r.max_pts,
                                                          r.ln().state.p, want_total, window, r.ln().att_next.p, r.ln().att_ok.p);
  ransac_chain_kernel<MODEL><<<B, 1024, chain_smem(window, attempts), c->stream>>>(
      r.rng.p, r.rng_len, r.npts.p, r.ln().att_next.p, r.ln().att_ok.p, r.ln().subsets.p, r.ln().state.p, want_total, r.cap_iters, window,
      attempts);
  c->launches += 2;
}

template <int MODEL>
static void solve_score(mvo_ctx* c, int h0, int h1) {
  RansacBufs& r = c->rs;
  const int B = c->cfg.batch;
  const int count = h1 - h0;
  // the solvers are latency bound (one FP64 thread per hypothesis): few threads per block spreads them over SMs
  int tpb = 8;
  while (tpb < 64 && (long long)B * count > 148LL * 4 * tpb) tpb <<= 1;
  dim3 gs((count + tpb - 1) / tpb, B);
  if (MODEL == MVO_MODEL_E) {
    dim3 gw((count + kE5SetupWarps - 1) / kE5SetupWarps, B);
    e5_setup_kernel<<<gw, kE5SetupWarps * 32, 0, c->stream>>>(r.q1.p, r.q2.p, r.max_pts, r.ln().subsets.p, r.ln().state.p, r.cap_iters, h0, h1,
                                               r.ln().e5_scratch.p);
    dim3 gr((count + kRootsThreads / 16 - 1) / (kRootsThreads / 16), B);
    e5_roots_kernel<<<gr, kRootsThreads, 0, c->stream>>>(r.ln().state.p, r.cap_iters, h0, h1, r.ln().e5_scratch.p, r.ln().models.p,
                                                         r.ln().nmodels.p, c->dbg_e5_roots_impl);
    c->launches++;
  } else {
    ransac_solve_kernel<MODEL><<<gs, tpb, 0, c->stream>>>(r.p1.p, r.p2.p, r.q1.p, r.q2.p, r.max_pts, r.ln().subsets.p,
                                                          r.ln().state.p, r.cap_iters, h0, h1, r.ln().models.p, r.ln().nmodels.p);
  }
  dim3 gc(std::min(count, kScoreGrid), B);
  ransac_score_kernel<MODEL><<<gc, kScoreThreads, 0, c->stream>>>(r.p1.p, r.p2.p, r.q1.p, r.q2.p, r.npts.p, r.max_pts,
                                                                 r.ln().state.p, r.cap_iters, h0, h1, r.ln().models.p,
                                                                 r.ln().nmodels.p, r.ln().thr2.p, r.ln().counts.p);
  c->launches += 2;
}

template <int MODEL>
static void select_pass(mvo_ctx* c, int n_eval, double conf) {
  RansacBufs& r = c->rs;
  ransac_select_kernel<MODEL><<<c->cfg.batch, 1024, 0, c->stream>>>(r.p1.p, r.p2.p, r.q1.p, r.q2.p, r.npts.p, r.max_pts,
                                                                  r.ln().state.p, r.cap_iters, n_eval, r.ln().models.p, r.ln().counts.p,
                                                                  r.ln().thr2.p, conf, r.ln().best_model.p, r.ln().mask.p, r.ln().result.p);
  c->launches++;
}

// full find*: points must already be in r.p1/r.p2 (and r.q1/r.q2 for E), counts in r.npts, thresholds in r.ln().thr2
template <int MODEL>
static int find_model(mvo_ctx* c, double conf) {
  RansacBufs& r = c->rs;
  constexpr int MAXIT = MT<MODEL>::MAXIT;
  MVO_CUDA_TRY(c, cudaMemsetAsync(r.ln().state.p, 0, (size_t)c->cfg.batch * 32, c->stream));
  // OpenCV's loop stops after a handful of iterations when most correspondences are inliers (1 - conf is reached
  // quickly): the first round is small so that easy frames do not pay for 128 solved and scored hypotheses
  if (kRoundA > 0) {
    sample_pass<MODEL>(c, kRoundA, kDrawsA, kAttemptsA);
    solve_score<MODEL>(c, 0, kRoundA);
    select_pass<MODEL>(c, kRoundA, conf);
  }
  // the first kRound0 iterations; usually the adaptive loop has stopped inside them
  sample_pass<MODEL>(c, kRound0, kDraws0, kAttempts0);
  solve_score<MODEL>(c, 0, kRound0);
  select_pass<MODEL>(c, kRound0, conf);
  // last round: everything up to maxIters (empty launches for streams that are done; the scoring grid is capped so that
  // those empty launches do not flood the CTA dispatcher while the next step's kernels run beside them)
  sample_pass<MODEL>(c, MAXIT, kDraws, kMaxAttempts);
  solve_score<MODEL>(c, 0, MAXIT);
  select_pass<MODEL>(c, MAXIT, conf);
  MVO_CUDA_TRY(c, cudaGetLastError());
#ifdef MVO_DK_DEBUG
  if (MODEL == MVO_MODEL_E) {
    unsigned hs[2];
    cudaStreamSynchronize(c->stream);
    cudaMemcpyFromSymbol(hs, g_e5_stats, 8);
    fprintf(stderr, "e5 roots: aberth %u, bracketing %u\n", hs[0], hs[1]);
  }
#endif
  return MVO_OK;
}

int ransac_find(mvo_ctx* c, int model, double conf) {
  RansacBufs& r = c->rs;
  int rc;
  if (model == MVO_MODEL_H) {
    rc = find_model<MVO_MODEL_H>(c, conf);
    if (rc) return rc;
    if (c->dbg_h_refine_impl == 1)   // the first-generation kernel, kept as the in-tree cross-check
      h_refine_kernel<<<c->cfg.batch, kRefThreads, 0, c->stream>>>(r.p1.p, r.p2.p, r.npts.p, r.max_pts, r.ln().thr2.p,
                                                                  r.ln().best_model.p, r.ln().mask.p, r.ln().result.p, r.ln().inl_idx.p);
    else
      h_refine2_kernel<<<c->cfg.batch, kRefThreads, 0, c->stream>>>(r.p1.p, r.p2.p, r.npts.p, r.max_pts, r.ln().thr2.p,
                                                                   r.ln().best_model.p, r.ln().mask.p, r.ln().result.p, r.ln().inl_idx.p);
    c->launches++;
    MVO_CUDA_TRY(c, cudaGetLastError());
    return MVO_OK;
  }
  if (model == MVO_MODEL_F) return find_model<MVO_MODEL_F>(c, conf);
  return find_model<MVO_MODEL_E>(c, conf);
}

// C4 sweep: m hypotheses, no early exit
template <int MODEL>
static int sweep_model(mvo_ctx* c, int m) {
  RansacBufs& r = c->rs;
  MVO_CUDA_TRY(c, cudaMemsetAsync(r.ln().state.p, 0, (size_t)c->cfg.batch * 32, c->stream));
  const int passes = (m + kSegIters - 1) / kSegIters + 2;
  for (int l = 0; l < passes; ++l) sample_pass<MODEL>(c, m, kDraws, kMaxAttempts);
  solve_score<MODEL>(c, 0, m);
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

int ransac_normalize(mvo_ctx* c) {
  RansacBufs& r = c->rs;
  dim3 grid((r.max_pts + 255) / 256, c->cfg.batch);
  normalize_points_kernel<<<grid, 256, 0, c->stream>>>(r.p1.p, r.p2.p, r.npts.p, r.max_pts, r.K.p, r.q1.p, r.q2.p);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

int ransac_sweep(mvo_ctx* c, int model, int m) {
  if (model == MVO_MODEL_H) return sweep_model<MVO_MODEL_H>(c, m);
  if (model == MVO_MODEL_F) return sweep_model<MVO_MODEL_F>(c, m);
  return sweep_model<MVO_MODEL_E>(c, m);
}

}  // namespace mvo

// ====================================================================================================
using namespace mvo;

static int upload_points(mvo_ctx* c, const float* p1, const float* p2, int n, double thr2, const double* K, int cap_iters) {
  MVO_REQUIRE_IDLE(c);
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if (c->cfg.batch != 1) {
    c->set_error("the single-call geometry API needs a batch==1 context");
    return MVO_ERR_INVALID;
  }
  int rc = ransac_prepare(c, std::max(n, std::max(c->rs.max_pts, 64)), std::max(cap_iters, std::max(c->rs.cap_iters, 2000)));
  if (rc) return rc;
  RansacBufs& r = c->rs;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.p1.p, p1, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.p2.p, p2, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  struct { int n; float t; } h = {n, (float)thr2};
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.npts.p, &h.n, 4, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(r.ln().thr2.p, &h.t, 4, cudaMemcpyHostToDevice, c->stream));
  if (K) MVO_CUDA_TRY(c, cudaMemcpyAsync(r.K.p, K, 72, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // h lives on this stack frame
  return MVO_OK;
}

static int download_result(mvo_ctx* c, int n, double* model, uint8_t* mask, int* n_inl) {
  RansacBufs& r = c->rs;
  int res[8];
  MVO_CUDA_TRY(c, cudaMemcpyAsync(res, r.ln().result.p, 32, cudaMemcpyDeviceToHost, c->stream));
  if (model) MVO_CUDA_TRY(c, cudaMemcpyAsync(model, r.ln().best_model.p, 72, cudaMemcpyDeviceToHost, c->stream));
  if (mask) MVO_CUDA_TRY(c, cudaMemcpyAsync(mask, r.ln().mask.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (n_inl) *n_inl = res[0];
  c->last_ransac_iters = res[1];
  if (res[2] < 0) {
    c->set_error("RANSAC found no model");
    return MVO_ERR_DEGENERATE;
  }
  return MVO_OK;
}

// cv::findFundamentalMat below 15 correspondences: N == 7 returns the 7-point solution of the points themselves with an
// all-ones mask; 8 <= N < 15 runs LMedS (niters from confidence and a fixed 45 % outlier ratio, at least 3).
static int ransac_f_small(mvo_ctx* c, const float* p1, const float* p2, int n, double conf, double* F, uint8_t* mask,
                          int* n_inliers) {
  constexpr int MAXIT = MT<MVO_MODEL_F>::MAXIT;
  int rc = upload_points(c, p1, p2, n, 1.0, nullptr, MAXIT);
  if (rc) return rc;
  RansacBufs& r = c->rs;
  MVO_CUDA_TRY(c, cudaMemsetAsync(r.ln().state.p, 0, 32, c->stream));
  int niters = 1;
  if (n == 7) {
    const int32_t st[8] = {0, 1, 0, 0, 0, 0, 0, 0}, idx[7] = {0, 1, 2, 3, 4, 5, 6};
    MVO_CUDA_TRY(c, cudaMemcpyAsync(r.ln().state.p, st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(r.ln().subsets.p, idx, sizeof(idx), cudaMemcpyHostToDevice, c->stream));
    MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // st / idx live on this stack frame
  } else {
    niters = std::max(update_iters(conf, 0.45, 7, MAXIT), 3);
    sample_pass<MVO_MODEL_F>(c, niters, kDraws, kMaxAttempts);
  }
  ransac_solve_kernel<MVO_MODEL_F><<<dim3((niters + 63) / 64, 1), 64, 0, c->stream>>>(
      r.p1.p, r.p2.p, r.q1.p, r.q2.p, r.max_pts, r.ln().subsets.p, r.ln().state.p, r.cap_iters, 0, niters, r.ln().models.p,
      r.ln().nmodels.p);
  c->launches++;
  if (n == 7) {
    int nm = 0;
    MVO_CUDA_TRY(c, cudaMemcpyAsync(&nm, r.ln().nmodels.p, 4, cudaMemcpyDeviceToHost, c->stream));
    if (F) MVO_CUDA_TRY(c, cudaMemcpyAsync(F, r.ln().models.p, 72, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->last_ransac_iters = 1;
    if (nm <= 0) {
      c->set_error("mvo_find_fundamental: the 7-point solver found no model");
      return MVO_ERR_DEGENERATE;
    }
    if (mask) std::fill(mask, mask + n, (uint8_t)1);
    if (n_inliers) *n_inliers = n;
    return MVO_OK;
  }
  f_lmeds_kernel<<<1, 1024, 0, c->stream>>>(r.p1.p, r.p2.p, r.npts.p, r.max_pts, r.ln().state.p, r.cap_iters, niters,
                                            r.ln().models.p, r.ln().nmodels.p, r.ln().best_model.p, r.ln().mask.p, r.ln().result.p);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  int n_in = 0;
  rc = download_result(c, n, F, mask, &n_in);
  if (n_inliers) *n_inliers = n_in;
  if (rc) return rc;
  if (n_in < 7) {   // OpenCV: run() fails (empty matrix) although the mask has been written
    c->set_error("mvo_find_fundamental: LMedS kept fewer than 7 inliers");
    return MVO_ERR_DEGENERATE;
  }
  return MVO_OK;
}

extern "C" {

int mvo_find_homography(mvo_ctx* c, const float* p1, const float* p2, int n, double thr, double* H, uint8_t* mask,
                        int* n_inliers) {
  if (!c) return MVO_ERR_INVALID;
  if (!p1 || !p2 || n < 0) {
    c->set_error("mvo_find_homography: null argument");
    return MVO_ERR_INVALID;
  }
  if (n < 4) {
    c->set_error("mvo_find_homography: fewer than 4 correspondences");
    return MVO_ERR_DEGENERATE;
  }
  int rc = upload_points(c, p1, p2, n, thr * thr, nullptr, 2000);
  if (rc) return rc;
  rc = ransac_find(c, MVO_MODEL_H, 0.995);
  if (rc) return rc;
  return download_result(c, n, H, mask, n_inliers);
}

int mvo_find_fundamental(mvo_ctx* c, const float* p1, const float* p2, int n, double thr, double conf, double* F,
                         uint8_t* mask, int* n_inliers) {
  if (!c) return MVO_ERR_INVALID;
  if (!p1 || !p2 || n < 0) {
    c->set_error("mvo_find_fundamental: null argument");
    return MVO_ERR_INVALID;
  }
  if (n < 7) {
    c->set_error("mvo_find_fundamental: fewer than 7 correspondences");
    return MVO_ERR_DEGENERATE;
  }
  if (thr <= 0) thr = 3;
  if (conf < DBL_EPSILON || conf > 1 - DBL_EPSILON) conf = 0.99;
  if (n < 15) return ransac_f_small(c, p1, p2, n, conf, F, mask, n_inliers);   // OpenCV: direct (N == 7) or LMedS
  int rc = upload_points(c, p1, p2, n, thr * thr, nullptr, 1000);
  if (rc) return rc;
  rc = ransac_find(c, MVO_MODEL_F, conf);
  if (rc) return rc;
  return download_result(c, n, F, mask, n_inliers);
}

int mvo_find_essential(mvo_ctx* c, const float* p1, const float* p2, int n, const double* K, double conf, double thr,
                       double* E, uint8_t* mask, int* n_inliers) {
  if (!c) return MVO_ERR_INVALID;
  if (!p1 || !p2 || !K || n < 0) {
    c->set_error("mvo_find_essential: null argument");
    return MVO_ERR_INVALID;
  }
  if (n < 5) {
    c->set_error("mvo_find_essential: fewer than 5 correspondences");
    return MVO_ERR_DEGENERATE;
  }
  const double t = thr / ((K[0] + K[4]) / 2);
  int rc = upload_points(c, p1, p2, n, t * t, K, 1000);
  if (rc) return rc;
  rc = ransac_normalize(c);
  if (rc) return rc;
  rc = ransac_find(c, MVO_MODEL_E, conf);
  if (rc) return rc;
  return download_result(c, n, E, mask, n_inliers);
}

int mvo_score_hypotheses(mvo_ctx* c, int model, const float* p1, const float* p2, int n, const double* K, double thr,
                         int m, int32_t* sample_idx, int32_t* counts, double* models) {
  if (!c) return MVO_ERR_INVALID;
  if (!p1 || !p2 || !counts || n < 8 || m < 1 || m > 16384 || model < 0 || model > 2 || (model == MVO_MODEL_E && !K)) {
    c->set_error("mvo_score_hypotheses: bad argument");
    return MVO_ERR_INVALID;
  }
  double t2 = thr * thr;
  if (model == MVO_MODEL_E) {
    const double t = thr / ((K[0] + K[4]) / 2);
    t2 = t * t;
  }
  int rc = upload_points(c, p1, p2, n, t2, K, m);
  if (rc) return rc;
  if (model == MVO_MODEL_E) {
    rc = ransac_normalize(c);
    if (rc) return rc;
  }
  rc = ransac_sweep(c, model, m);
  if (rc) return rc;
  RansacBufs& r = c->rs;
  const int kk = model == MVO_MODEL_H ? 4 : model == MVO_MODEL_F ? 7 : 5;
  const int mm = model == MVO_MODEL_H ? 1 : model == MVO_MODEL_F ? 3 : 10;
  int st[8];
  MVO_CUDA_TRY(c, cudaMemcpyAsync(st, r.ln().state.p, 32, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(counts, r.ln().counts.p, (size_t)m * mm * 4, cudaMemcpyDeviceToHost, c->stream));
  if (sample_idx) MVO_CUDA_TRY(c, cudaMemcpyAsync(sample_idx, r.ln().subsets.p, (size_t)m * kk * 4, cudaMemcpyDeviceToHost, c->stream));
  if (models) MVO_CUDA_TRY(c, cudaMemcpyAsync(models, r.ln().models.p, (size_t)m * mm * 72, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (st[1] < m) {
    c->set_error("could not draw the requested number of valid minimal samples");
    return MVO_ERR_DEGENERATE;
  }
  return MVO_OK;
}

}  // extern "C"
