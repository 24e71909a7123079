// orb.cu -- cv::ORB::detectAndCompute re-designed for B200 (sm_100a).
//
// Replaces the call at /root/reference/src/feature_processor.cpp:22 (detector_->detectAndCompute).
// Arithmetic contract: SURVEY.md Appendix A.1 (bit-exact keypoint set / response / angle / descriptors
// against cv2 4.13.0).  Structure (per stream group, batch dimension = blockIdx.z):
//
//   K1  orb_level_kernel<RESIZE>   one launch per pyramid level: INTER_LINEAR_EXACT resize of the 64x32
//                                  tile (+4 halo) from level l-1 staged in smem with 128-bit loads ->
//                                  level l store + FAST-9/16 (quick-reject + compacted full score) +
//                                  3x3 NMS + edge filter + score histogram + 7x7 float blur store.
//   K2a orb_select_kernel          CTA per (level, stream): retainBest(2 n_l) cut from the score histogram (parallel
//                                  suffix scan) + compaction of the surviving candidates.
//   K2  orb_harris_angle_kernel    warp per surviving candidate: 7x7 Harris
//                                  (int32 sums, non-fused FP32) + intensity-centroid angle.
//   K3  orb_rank_kernel            per level all-pairs rank on a unique 64-bit key -> sorted scatter.
//   K4  orb_finalize_kernel        retainBest(n_l) cut with ties, cross-level prefix, mvo_keypoint write.
//   K5  orb_brief_kernel           warp per keypoint, lane per descriptor byte (16 rotated samples).
#include "context.cuh"
#include <math.h>
#include <algorithm>

namespace mvo {

// ------------------------------------------------------------------------------------------------
// constants (bit patterns verified against cv2 4.13.0, see oracle/orb_oracle.py)
__constant__ float c_gauss[7] = {0.07015932351350784f, 0.13107487559318542f, 0.1907128244638443f,
                                 0.21610593795776367f, 0.1907128244638443f,  0.13107487559318542f,
                                 0.07015932351350784f};
// rBRIEF sampling pattern.  In global memory, not __constant__: every lane reads its own 16 points (a
// lane-divergent index serialises 32-way on the constant cache); two 128-bit __ldg per lane instead.
__device__ __align__(16) signed char d_pattern[512][2] = {
#include "brief_pattern_31.inc"
};

constexpr int TW = 64, TH = 32, HALO = 4;
constexpr int TS = 96;    // smem tile row stride in bytes
constexpr int TX0 = 16;   // smem column of tile-local x = 0 (keeps the interior 16-byte aligned)
constexpr int SROWS = TH + 2 * HALO;  // 40
constexpr int SCOLS = TW + 2 * HALO;  // 72
constexpr int SRC_ROWS_MAX = 56;
constexpr int SRC_COLS_MAX = 128;
constexpr int FW = TW + 2, FH = TH + 2;  // FAST score region (1 halo for NMS)
constexpr int FS = 68;                   // score row stride
#ifndef MVO_ORB_MINB
#define MVO_ORB_MINB 6   // 40 registers, no spills, 6 CTAs / SM (5: 48 registers, 3 % slower; 7: spills)
#endif
constexpr int kLevelThreads = 256;
constexpr int kGridMaxCells = 8192;      // occupancy bitmap of orb_finalize_kernel (1 KB of shared memory)
constexpr int kFastCols = (TW + 8) / 4;  // 18 aligned 4-pixel groups cover x = -4 .. TW+3
constexpr int kFastLanes = kLevelThreads / kFastCols;  // 14 row lanes
constexpr int kHSegs = 3;                               // row segments of the horizontal resize pass
constexpr int kHRows = (SROWS + kHSegs - 1) / kHSegs;   // 14

struct LevelArgs {
  const uint8_t* src;
  uint8_t* dst;
  uint8_t* blur;
  long long frame_stride;
  int sw, sh, spitch;
  int w, h, pitch;
  const uint32_t* xtab;
  const uint32_t* ytab;
  uint32_t* cand_xy;
  int32_t* cand_score;
  int32_t* cand_count;
  uint32_t* hist;
  int32_t* flags;
  int cand_cap, cand_total, cand_off;
  int level;
  int do_fast;   // 0: pyramid + blur only (orb_compute hook)
};

// FAST-9/16 score of one pixel: max over the 16 arcs of 9 contiguous ring pixels of min(d) and of min(-d),
// d = centre - ring.  Both polarities travel in one register as biased 16-bit lanes (low: d + 256, high:
// -d + 256, both in [1, 511]) so that every min / max is one VIMNMX3.U16x2.  No min/max result is ever negated
// (ptxas 12.9 for sm_100a drops the negation when it fuses max(a, -max(b, c)) into VIMNMX3).
__device__ __forceinline__ int fast_full_score(const uint8_t* t, int idx) {
  // OpenCV ring order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
  const int off[16] = {3 * TS,      3 * TS + 1,  2 * TS + 2,  TS + 3,  3,        -TS + 3,  -2 * TS + 2, -3 * TS + 1,
                       -3 * TS,     -3 * TS - 1, -2 * TS - 2, -TS - 3, -3,       TS - 3,   2 * TS - 2,  3 * TS - 1};
  const uint32_t c = t[idx];
  const uint32_t kc = 0x01000100u - 65535u * c;   // lanes (256 + c, 256 - c) before the ring pixel is applied
  uint32_t v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = kc + 65535u * (uint32_t)t[idx + off[k]];   // (256 + c - r, 256 + r - c)
  uint32_t m3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) m3[k] = __vimin3_u16x2(v[k], v[(k + 1) & 15], v[(k + 2) & 15]);
  uint32_t best = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const uint32_t a = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
    const uint32_t b = __vimin3_u16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
    best = __vimax3_u16x2(best, a, b);
  }
  return (int)max(best & 0xffffu, best >> 16) - 256;  // corner iff > threshold; score = value - 1
}

// per-byte "d > kFastThr" for four packed absolute differences: bit 7 of every byte
__device__ __forceinline__ uint32_t bytes_gt_thr(uint32_t d) {
  return ((d & 0x7f7f7f7fu) + 0x01010101u * (uint32_t)(127 - kFastThr)) | d;
}

template <bool RESIZE>
__global__ void __launch_bounds__(kLevelThreads, MVO_ORB_MINB) orb_level_kernel(const LevelArgs a) {
  __shared__ __align__(16) uint8_t tile[SROWS * TS];
  // scratch: resize staging (source rows + vertical pass, 8.8 fixed point) is dead before the blur row buffer is live
  __shared__ __align__(16) uint8_t scratch[SRC_ROWS_MAX * SRC_COLS_MAX + SROWS * SRC_COLS_MAX * 2];
  __shared__ __align__(16) uint8_t score[(FH * FS + 15) / 16 * 16];
  __shared__ uint16_t work[FW * FH];
  __shared__ uint32_t hist_s[256];
  __shared__ uint32_t emit_xy[TW * TH / 4];
  __shared__ uint8_t emit_sc[TW * TH / 4];
  __shared__ uint32_t tab_x[SCOLS], tab_y[SROWS];
  __shared__ int n_work, n_emit, emit_base;

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
  const int w = a.w, h = a.h;
  uint8_t* dst = a.dst + (long long)b * a.frame_stride;
  uint8_t* blur = a.blur + (long long)b * a.frame_stride;

  if (tid == 0) {
    n_work = 0;
    n_emit = 0;
  }
  hist_s[tid] = 0;  // kLevelThreads == 256
  if (tid < (int)sizeof(score) / 16) reinterpret_cast<uint4*>(score)[tid] = make_uint4(0, 0, 0, 0);

  if (RESIZE) {
    // INTER_LINEAR_EXACT is exact integer arithmetic up to the final shift: out = (sum wx*wy*src + 2^15) >> 16,
    // so the two passes commute.  Vertical first: it runs on whole staged words (two pixels per 16-bit lane pair),
    // the horizontal gather then walks columns with its coefficients in registers.
    const uint8_t* src = a.src + (long long)b * a.frame_stride;
    uint32_t* src_s = reinterpret_cast<uint32_t*>(scratch);                                     // [row][32 words]
    uint16_t* vbuf = reinterpret_cast<uint16_t*>(scratch + SRC_ROWS_MAX * SRC_COLS_MAX);        // [SROWS][128]
    if (tid < SCOLS) tab_x[tid] = __ldg(a.xtab + min(max(tx0 - HALO + tid, 0), w - 1));
    if (tid >= 128 && tid < 128 + SROWS) tab_y[tid - 128] = __ldg(a.ytab + min(max(ty0 - HALO + tid - 128, 0), h - 1));
    __syncthreads();
    const int r0 = tab_y[0] & 0xffff;
    const int r1 = min((int)(tab_y[SROWS - 1] & 0xffff) + 1, a.sh - 1);
    const int nrows = min(r1 - r0 + 1, SRC_ROWS_MAX);
    const int c0a = (int)(tab_x[0] & 0xffff) & ~15;
    const int c1e = min((int)(tab_x[SCOLS - 1] & 0xffff) + 1, a.sw - 1);
    const int nch = min((c1e - c0a) / 16 + 1, SRC_COLS_MAX / 16);
    // stage A: source rows, 128-bit loads
    {
      // thread = (row within a pass of 32 rows, 16-byte chunk): no division by the run-time chunk count
      static_assert(SRC_COLS_MAX / 16 == 8 && kLevelThreads == 256, "stage A maps 8 chunks x 32 rows per pass");
      const int k = tid & 7;
      if (k < nch) {
        for (int r = tid >> 3; r < nrows; r += kLevelThreads / 8) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (long long)(r0 + r) * a.spitch + c0a + k * 16));
          *reinterpret_cast<uint4*>(src_s + r * (SRC_COLS_MAX / 4) + k * 4) = v;
        }
      }
    }
    __syncthreads();
    // stage V: vertical pass, one staged word (4 source pixels) per lane, one output row per warp and round
    {
      const int wd = tid & 31;
      if (wd < nch * 4) {
        for (int yl = tid >> 5; yl < SROWS; yl += kLevelThreads / 32) {
          const uint32_t e = tab_y[yl];
          const int o = e & 0xffff;
          const uint32_t c1 = e >> 16, c0 = 256u - c1;
          const int ra = min(o - r0, nrows - 1), rb = min(min(o + 1, a.sh - 1) - r0, nrows - 1);
          const uint32_t pa = src_s[ra * (SRC_COLS_MAX / 4) + wd], pb = src_s[rb * (SRC_COLS_MAX / 4) + wd];
          const uint32_t te = (pa & 0x00ff00ffu) * c0 + (pb & 0x00ff00ffu) * c1;                // pixels 0, 2
          const uint32_t to = ((pa >> 8) & 0x00ff00ffu) * c0 + ((pb >> 8) & 0x00ff00ffu) * c1;  // pixels 1, 3
          uint2 o2;
          o2.x = __byte_perm(te, to, 0x5410);   // (pixel 0, pixel 1) as u16 pair
          o2.y = __byte_perm(te, to, 0x7632);   // (pixel 2, pixel 3)
          *reinterpret_cast<uint2*>(vbuf + yl * SRC_COLS_MAX + wd * 4) = o2;
        }
      }
    }
    __syncthreads();
    // stage H: horizontal pass, thread = output column (coefficients in registers) x row segment
    if (tid < SCOLS * kHSegs) {
      const int seg = tid / SCOLS, xl = tid - seg * SCOLS;
      const uint32_t e = tab_x[xl];
      const int o = e & 0xffff;
      const uint32_t c1 = e >> 16, c0 = 256u - c1;
      const int oa = o - c0a, ob = min(o + 1, a.sw - 1) - c0a;
      const uint16_t* va = vbuf + seg * kHRows * SRC_COLS_MAX;
      uint8_t* tp = tile + seg * kHRows * TS + TX0 - HALO + xl;
#pragma unroll
      for (int j = 0; j < kHRows; ++j) {
        if (seg * kHRows + j < SROWS) {
          const uint32_t v = ((uint32_t)va[j * SRC_COLS_MAX + oa] * c0 + (uint32_t)va[j * SRC_COLS_MAX + ob] * c1 + 32768u) >> 16;
          tp[j * TS] = (uint8_t)v;
        }
      }
    }
    __syncthreads();
    // level store, 128-bit
    for (int i = tid; i < TH * (TW / 16); i += kLevelThreads) {
      const int yl = i / (TW / 16), k = i - yl * (TW / 16);
      const int gy = ty0 + yl;
      if (gy < h)
        *reinterpret_cast<uint4*>(dst + (long long)gy * a.pitch + tx0 + k * 16) =
            *reinterpret_cast<const uint4*>(tile + (yl + HALO) * TS + TX0 + k * 16);
    }
  } else {
    // level 0: the image itself; load tile + halo with 128-bit loads
    for (int i = tid; i < SROWS * (TS / 16); i += kLevelThreads) {
      const int yl = i / (TS / 16), k = i - yl * (TS / 16);
      const int gy = min(max(ty0 - HALO + yl, 0), h - 1);
      const int gx0 = tx0 - TX0 + k * 16;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gx0 >= 0 && gx0 < a.pitch) v = __ldg(reinterpret_cast<const uint4*>(dst + (long long)gy * a.pitch + gx0));
      *reinterpret_cast<uint4*>(tile + yl * TS + k * 16) = v;
    }
  }
  __syncthreads();

  // ---- tiles on the image border: REFLECT_101 halo (only the 7x7 blur looks outside the image) --------
  if (tx0 == 0 || tx0 + TW + HALO > w || ty0 == 0 || ty0 + TH + HALO > h) {
    const int xlo = tx0 - HALO, xhi = tx0 + TW + HALO - 1, ylo = ty0 - HALO, yhi = ty0 + TH + HALO - 1;
    for (int i = tid; i < SROWS * 2 * HALO; i += kLevelThreads) {
      const int yl = i / (2 * HALO), j = i - yl * (2 * HALO);
      const int gx = j < HALO ? j - HALO : w + j - HALO;
      if (gx >= xlo && gx <= xhi) {
        const int sx = min(max(min(max(reflect101(gx, w), 0), w - 1), xlo), xhi);
        tile[yl * TS + TX0 + gx - tx0] = tile[yl * TS + TX0 + sx - tx0];
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * HALO * SCOLS; i += kLevelThreads) {
      const int j = i / SCOLS, xc = i - j * SCOLS;
      const int gy = j < HALO ? j - HALO : h + j - HALO;
      if (gy >= ylo && gy <= yhi) {
        const int sy = min(max(min(max(reflect101(gy, h), 0), h - 1), ylo), yhi);
        tile[(gy - ylo) * TS + TX0 - HALO + xc] = tile[(sy - ylo) * TS + TX0 - HALO + xc];
      }
    }
    __syncthreads();
  }

  // ---- FAST-9/16 and the ORB-internal blur, interleaved -------------------------------------------------------
  // Both only read the finished tile, so their phases share barriers (three instead of seven) and threads that run
  // out of FAST work pick up blur work:
  //   A  FAST quick reject (packed, 4 px per thread)        +  blur row pass      -> barrier
  //   B  full FAST score of the survivor list               +  blur column pass   -> barrier
  //   C  3x3 NMS + edge filter over the survivor list -> barrier -> D  append to the level's candidate list
  float* rowbuf = reinterpret_cast<float*>(scratch);  // (TH + 6) x TW floats = 9728 B; the resize staging is dead
  const float g0 = c_gauss[0], g1 = c_gauss[1], g2 = c_gauss[2], g3 = c_gauss[3];
  if (a.do_fast) {
    // phase 1, four pixels per thread: an arc of 9 contains one pixel of every antipodal pair, so a corner needs
    // |centre - ring| > thr on one of (0, 8) and on one of (4, 12); packed bytes, VABSDIFF4.  A thread keeps the
    // survivors of its three rows as a 12-bit mask; one warp scan + one shared atomic per warp places them in the
    // survivor list (per-pixel atomics cost more than the test itself).
    {
      const bool active = tid < kFastCols * kFastLanes;
      const int rl = tid / kFastCols, wi = tid - rl * kFastCols;
      const int xl0 = 4 * wi - 4, gx0 = tx0 + xl0;
      const int jlo = max(max(-1 - xl0, kEdge - 1 - gx0), 0), jhi = min(min(TW - xl0, w - kEdge - gx0), 3);
      uint32_t xmask = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j >= jlo && j <= jhi) xmask |= 0x80u << (8 * j);
      if (!active) xmask = 0;
      const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
      uint32_t found = 0;
      if (xmask) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int yl = rl - 1 + k * kFastLanes;
          const int gy = ty0 + yl;
          if (yl <= TH && gy >= kEdge - 1 && gy < h - (kEdge - 1)) {
            const int wq = ((yl + HALO) * TS + TX0 + xl0) >> 2;
            const uint32_t c = t32[wq];
            uint32_t m = bytes_gt_thr(__vabsdiffu4(c, t32[wq + 3 * (TS / 4)])) |
                         bytes_gt_thr(__vabsdiffu4(c, t32[wq - 3 * (TS / 4)]));
            m &= xmask;
            if (m) {
              const uint32_t l = t32[wq - 1], r = t32[wq + 1];
              const uint32_t p12 = __byte_perm(l, c, 0x4321);   // pixels x-3 .. x
              const uint32_t p4 = __byte_perm(c, r, 0x6543);    // pixels x+3 .. x+6
              m &= bytes_gt_thr(__vabsdiffu4(c, p4)) | bytes_gt_thr(__vabsdiffu4(c, p12));
              // bit 7 of byte j -> bit j
              const uint32_t nib = ((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u);
              found |= nib << (4 * k);
            }
          }
        }
      }
      const int cnt = __popc(found);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
      }
      int base = 0;
      if ((tid & 31) == 31 && incl) base = atomicAdd(&n_work, incl);
      base = __shfl_sync(0xffffffffu, base, 31);
      int pos = base + incl - cnt;
      while (found) {
        const int bit = __ffs(found) - 1;
        found &= found - 1;
        const int yl = rl - 1 + (bit >> 2) * kFastLanes;
        work[pos++] = (uint16_t)((yl + 1) * FW + xl0 + (bit & 3) + 1);
      }
    }
  }
  {
    // blur row pass: float sepFilter, FMA order of OpenCV's AVX2 path; 4 pixels per thread from three aligned words
    // (the halo already holds REFLECT_101 pixels)
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
    for (int i = tid; i < (TH + 6) * (TW / 4); i += kLevelThreads) {
      const int yr = i / (TW / 4), xg = i - yr * (TW / 4);
      const int wq = ((yr + 1) * TS + TX0 + 4 * xg) >> 2;
      const uint32_t w0 = t32[wq - 1], w1 = t32[wq], w2 = t32[wq + 1];
      float f[10];
      f[0] = (float)((w0 >> 8) & 0xff);
      f[1] = (float)((w0 >> 16) & 0xff);
      f[2] = (float)(w0 >> 24);
      f[3] = (float)(w1 & 0xff);
      f[4] = (float)((w1 >> 8) & 0xff);
      f[5] = (float)((w1 >> 16) & 0xff);
      f[6] = (float)(w1 >> 24);
      f[7] = (float)(w2 & 0xff);
      f[8] = (float)((w2 >> 8) & 0xff);
      f[9] = (float)((w2 >> 16) & 0xff);
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sacc = __fmul_rn(g0, f[j]);
        sacc = __fmaf_rn(g1, f[j + 1], sacc);
        sacc = __fmaf_rn(g2, f[j + 2], sacc);
        sacc = __fmaf_rn(g3, f[j + 3], sacc);
        sacc = __fmaf_rn(g2, f[j + 4], sacc);
        sacc = __fmaf_rn(g1, f[j + 5], sacc);
        sacc = __fmaf_rn(g0, f[j + 6], sacc);
        o[j] = sacc;
      }
      reinterpret_cast<float4*>(rowbuf)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();   // ---- barrier 1: survivor list and row buffer complete

  const int nw = a.do_fast ? n_work : 0;
  // phase B: full score of the compacted survivors
  for (int i = tid; i < nw; i += kLevelThreads) {
    const int p = work[i];
    const int yl = p / FW - 1, xl = p - (yl + 1) * FW - 1;
    const int best = fast_full_score(tile, (yl + HALO) * TS + TX0 + xl);
    if (best > kFastThr) score[(yl + 1) * FS + xl + 1] = (uint8_t)(best - 1);
  }
  {
    // blur column pass: 4 pixels x 2 rows per thread, symmetric taps added first (OpenCV's symmetric column filter)
    const int xg = tid & (TW / 4 - 1), yl0 = (tid / (TW / 4)) * 2;
    const float4* rb = reinterpret_cast<const float4*>(rowbuf) + yl0 * (TW / 4) + xg;
    float4 r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = rb[k * (TW / 4)];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gy = ty0 + yl0 + j;
      uint32_t packed = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float m3 = reinterpret_cast<const float*>(&r[j + 3])[q];
        const float m2 = __fadd_rn(reinterpret_cast<const float*>(&r[j + 2])[q], reinterpret_cast<const float*>(&r[j + 4])[q]);
        const float m1 = __fadd_rn(reinterpret_cast<const float*>(&r[j + 1])[q], reinterpret_cast<const float*>(&r[j + 5])[q]);
        const float m0 = __fadd_rn(reinterpret_cast<const float*>(&r[j])[q], reinterpret_cast<const float*>(&r[j + 6])[q]);
        float sacc = __fmul_rn(g3, m3);
        sacc = __fmaf_rn(g2, m2, sacc);
        sacc = __fmaf_rn(g1, m1, sacc);
        sacc = __fmaf_rn(g0, m0, sacc);
        uint32_t v;
        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(v) : "f"(sacc));   // saturate_cast<uchar>(cvRound(s))
        packed |= v << (8 * q);
      }
      if (gy < h) *reinterpret_cast<uint32_t*>(blur + (long long)gy * a.pitch + tx0 + 4 * xg) = packed;
    }
  }
  if (!a.do_fast) return;
  __syncthreads();   // ---- barrier 2: scores complete

  // phase C: 3x3 non-max suppression + 31 px edge filter, again over the survivor list only
  for (int i = tid; i < nw; i += kLevelThreads) {
    const int p = work[i];
    const int yl = p / FW - 1, xl = p - (yl + 1) * FW - 1;
    if (xl < 0 || xl >= TW || yl < 0 || yl >= TH) continue;
    const int gx = tx0 + xl, gy = ty0 + yl;
    const uint8_t* sc = score + (yl + 1) * FS + xl + 1;
    const int v = sc[0];
    if (v > 0 && gx >= kEdge && gx < w - kEdge && gy >= kEdge && gy < h - kEdge) {
      if (v > sc[-1] && v > sc[1] && v > sc[-FS - 1] && v > sc[-FS] && v > sc[-FS + 1] && v > sc[FS - 1] &&
          v > sc[FS] && v > sc[FS + 1]) {
        const int e = atomicAdd(&n_emit, 1);
        emit_xy[e] = (uint32_t)gx | ((uint32_t)gy << 16);
        emit_sc[e] = (uint8_t)v;
        atomicAdd(&hist_s[v], 1u);
      }
    }
  }
  __syncthreads();   // ---- barrier 3
  const int ne = n_emit;
  if (ne > 0) {
    if (tid == 0) emit_base = atomicAdd(a.cand_count + b * kLevels + a.level, ne);
    __syncthreads();
    const int base = emit_base;
    for (int i = tid; i < ne; i += kLevelThreads) {
      const int pos = base + i;
      if (pos < a.cand_cap) {
        const long long o = (long long)b * a.cand_total + a.cand_off + pos;
        a.cand_xy[o] = emit_xy[i];
        a.cand_score[o] = emit_sc[i];
      } else {
        atomicOr(a.flags + b, 1);
      }
    }
    if (hist_s[tid]) atomicAdd(a.hist + ((long long)b * kLevels + a.level) * 256 + tid, hist_s[tid]);
  }
}

// ------------------------------------------------------------------------------------------------
// BGR -> gray, (B*3735 + G*19235 + R*9798 + 16384) >> 15   (cvtColor BGR2GRAY)
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, int stride, long long in_frame_stride,
                                   uint8_t* __restrict__ gray, int pitch, long long frame_stride, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const uint8_t* p = bgr + (long long)b * in_frame_stride + (long long)y * stride + 3 * x;
  gray[(long long)b * frame_stride + (long long)y * pitch + x] =
      (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

// ------------------------------------------------------------------------------------------------
// K2: Harris response, one warp per candidate that survives retainBest(2 n_l) by FAST score (the IC angle is computed
// later, for the n_l candidates that survive the Harris ranking: orb_brief_kernel)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = __uint_as_float(0x4265226fu), p3 = __uint_as_float(0xc19556eeu);
  const float p5 = __uint_as_float(0x410e9fbfu), p7 = __uint_as_float(0xc0228ad9u);
  const float eps = __uint_as_float(0x25800000u);  // (float)DBL_EPSILON
  const float ax = fabsf(x), ay = fabsf(y);
  const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
  const float c = __fdiv_rn(mn, __fadd_rn(mx, eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fadd_rn(__fmul_rn(p7, c2), p5);
  a = __fadd_rn(__fmul_rn(a, c2), p3);
  a = __fadd_rn(__fmul_rn(a, c2), p1);
  a = __fmul_rn(a, c);
  if (ax < ay) a = __fsub_rn(90.f, a);
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

// K2a: retainBest(2 n_l) by FAST score, one CTA per (level, stream): the cut r = largest score with
// #(score >= r) >= 2 n_l from the level's score histogram (parallel suffix scan), then the candidates with
// score >= r (ties kept, as cv::KeyPointsFilter::retainBest) are compacted into a dense list for the Harris kernel.
constexpr int kSelThreads = 256;
__global__ void __launch_bounds__(kSelThreads)
orb_select_kernel(const __grid_constant__ OrbGeom g, const uint32_t* __restrict__ cand_xy,
                  const int32_t* __restrict__ cand_score, const int32_t* __restrict__ cand_count,
                  const uint32_t* __restrict__ hist, uint32_t* __restrict__ sel_xy, int32_t* __restrict__ sel_count) {
  const int level = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cap = g.lv[level].cand_cap, quota = g.lv[level].quota, off = g.lv[level].cand_off;
  const int count = min(cand_count[b * kLevels + level], cap);
  __shared__ uint32_t s_warp[kSelThreads / 32];
  __shared__ int s_thr, s_n;
  // suffix sums of the histogram: thread t owns score 255 - t, so an inclusive prefix scan over t is the suffix count
  const uint32_t hv = hist[((long long)b * kLevels + level) * 256 + (255 - tid)];
  uint32_t incl = hv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  if (tid == 0) {
    s_thr = 0;
    s_n = 0;
  }
  __syncthreads();
  uint32_t base = 0;
  for (int w2 = 0; w2 < warp; ++w2) base += s_warp[w2];
  incl += base;
  const uint32_t k = 2u * (uint32_t)quota;
  // the first t (largest score) whose suffix count reaches k defines the cut; exactly one thread satisfies this
  if (count > (int)k && incl >= k && incl - hv < k) s_thr = 255 - tid;
  __syncthreads();
  const int thr = s_thr;
  const long long cbase = (long long)b * g.cand_total + off;
  for (int i0 = 0; i0 < count; i0 += kSelThreads) {
    const int i = i0 + tid;
    const bool keep = i < count && cand_score[cbase + i] >= thr;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    int wbase = 0;
    if (lane == 0 && bal) wbase = atomicAdd(&s_n, __popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (keep) sel_xy[cbase + wbase + __popc(bal & ((1u << lane) - 1))] = cand_xy[cbase + i];
  }
  __syncthreads();
  if (tid == 0) sel_count[b * kLevels + level] = s_n;
}

constexpr int kHarrisWarps = 8;
#ifndef MVO_HARRIS_PER_WARP
#define MVO_HARRIS_PER_WARP 8
#endif
constexpr int kHarrisPerWarp = MVO_HARRIS_PER_WARP;   // candidates of level 0 per warp (software-pipelined loop)

__global__ void __launch_bounds__(kHarrisWarps * 32)
orb_harris_angle_kernel(const __grid_constant__ OrbGeom g, const uint8_t* __restrict__ pyr, const uint32_t* __restrict__ sel_xy,
                        const int32_t* __restrict__ sel_count, unsigned long long* __restrict__ c2_key,
                        float2* __restrict__ c2_ra, int32_t* __restrict__ c2_count) {
  const int level = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int count = sel_count[b * kLevels + level];
  const int warp = tid >> 5, lane = tid & 31;
  if (blockIdx.x == 0 && tid == 0) c2_count[b * kLevels + level] = count;
  if (blockIdx.x * kHarrisWarps >= count) return;
  const int pitch = g.lv[level].pitch, cand_off = g.lv[level].cand_off;
  const uint8_t* img = pyr + (long long)b * g.frame_stride + g.lv[level].off;
  // Grid-stride over this level's selected candidates, about four per warp, software-pipelined: the 16 neighbourhood
  // bytes of the next candidate are requested before the current one is reduced (one candidate per warp and launch
  // wave left the kernel waiting on its gathers: 14 waves of one L2 round trip each).
  // Harris 7x7 block of Sobel-3 gradients: 49 positions over the 32 lanes (2 rounds); per position the 8 neighbours.
  const int stride_c = gridDim.x * kHarrisWarps;
  const long long sel_base = (long long)b * g.cand_total + cand_off;
  auto fetch = [&](int ci, uint32_t& xy, int (&px)[16]) {
    xy = sel_xy[sel_base + ci];
    const int x = xy & 0xffff, y = xy >> 16;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int p = lane + 32 * r;
      if (p < 49) {
        const int dy = p / 7 - 3, dx = p - (p / 7) * 7 - 3;
        const uint8_t* q = img + (long long)(y + dy) * pitch + x + dx;
        px[8 * r + 0] = q[-pitch - 1]; px[8 * r + 1] = q[-pitch]; px[8 * r + 2] = q[-pitch + 1];
        px[8 * r + 3] = q[-1];         px[8 * r + 4] = q[1];
        px[8 * r + 5] = q[pitch - 1];  px[8 * r + 6] = q[pitch];  px[8 * r + 7] = q[pitch + 1];
      }
    }
  };
  int ci = blockIdx.x * kHarrisWarps + warp;
  uint32_t xy = 0, xy_n = 0;
  int px[16], pn[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) px[q] = pn[q] = 0;
  if (ci < count) fetch(ci, xy, px);
  for (; ci < count; ci += stride_c) {
  const int ci_n = ci + stride_c;
  if (ci_n < count) fetch(ci_n, xy_n, pn);
  const int x = xy & 0xffff, y = xy >> 16;
  int sa = 0, sb = 0, sc = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int p = lane + 32 * r;
    if (p < 49) {
      const int a00 = px[8 * r + 0], a01 = px[8 * r + 1], a02 = px[8 * r + 2];
      const int a10 = px[8 * r + 3], a12 = px[8 * r + 4];
      const int a20 = px[8 * r + 5], a21 = px[8 * r + 6], a22 = px[8 * r + 7];
      const int ix = (a12 - a10) * 2 + (a02 - a00) + (a22 - a20);
      const int iy = (a21 - a01) * 2 + (a20 - a00) + (a22 - a02);
      sa += ix * ix;
      sb += iy * iy;
      sc += ix * iy;
    }
  }
  // |sums| < 2^31 (49 x 1020^2; 31 x 15 x 31 x 255): one REDUX.SUM each instead of a five-step shuffle tree
  sa = __reduce_add_sync(0xffffffffu, sa);
  sb = __reduce_add_sync(0xffffffffu, sb);
  sc = __reduce_add_sync(0xffffffffu, sc);

  if (lane == 0) {
    const float s4 = __uint_as_float(0x25ddced1u);  // ((1/(4*7*255))^4 accumulated in float
    const float k = __uint_as_float(0x3d23d70au);   // 0.04f
    const float fa = (float)sa, fb = (float)sb, fc = (float)sc;
    const float t1 = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
    const float apb = __fadd_rn(fa, fb);
    const float t2 = __fmul_rn(__fmul_rn(k, apb), apb);
    const float resp = __fmul_rn(__fsub_rn(t1, t2), s4);
    // (the candidate's own index is dense and unique already: no counter to contend for)
    const long long o = (long long)b * g.cand_total + cand_off + ci;
    // ascending key == (response desc, y asc, x asc); -0.f canonicalised so that ties compare equal
    const uint32_t ro = ~float_orderable(__fadd_rn(resp, 0.f));
    c2_key[o] = ((unsigned long long)ro << 32) | ((unsigned long long)y << 16) | (unsigned long long)x;
    c2_ra[o] = make_float2(resp, 0.f);   // the angle is computed for the survivors only (orb_brief_kernel)
  }
  xy = xy_n;
#pragma unroll
  for (int q = 0; q < 16; ++q) px[q] = pn[q];
  }  // candidate loop
}

// ------------------------------------------------------------------------------------------------
// K3: rank each Harris candidate inside its level by all-pairs comparison of the unique key
constexpr int kRankThreads = 256;
__global__ void __launch_bounds__(kRankThreads)
orb_rank_kernel(const __grid_constant__ OrbGeom g, const unsigned long long* __restrict__ key, const float2* __restrict__ ra,
                const int32_t* __restrict__ c2_count, unsigned long long* __restrict__ key_sorted,
                float2* __restrict__ ra_sorted) {
  const int level = blockIdx.y, b = blockIdx.z;
  const LevelGeom lv = g.lv[level];
  const int m = c2_count[b * kLevels + level];
  const long long base = (long long)b * g.cand_total + lv.cand_off;
  __shared__ unsigned long long s_key[kRankThreads];
  // the grid covers 2 n_0 candidates plus slack (a level keeps ~2 n_l, ties can add a few): block-stride beyond that.
  // (It used to cover the whole candidate capacity: 114 CTAs per level and stream of which 4 had work.)
  for (int blk = blockIdx.x; blk * kRankThreads < m; blk += gridDim.x) {
    const int i = blk * kRankThreads + threadIdx.x;
    const unsigned long long mine = (i < m) ? key[base + i] : ~0ull;
    int rank = 0;
    for (int j0 = 0; j0 < m; j0 += kRankThreads) {
      const int j = j0 + threadIdx.x;
      __syncthreads();
      s_key[threadIdx.x] = (j < m) ? key[base + j] : ~0ull;
      __syncthreads();
      const int lim = min(kRankThreads, m - j0);
#pragma unroll 8
      for (int t = 0; t < lim; ++t) rank += (s_key[t] < mine) ? 1 : 0;
    }
    if (i < m) {
      key_sorted[base + rank] = mine;
      ra_sorted[base + rank] = ra[base + i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: retainBest(n_l) with ties + cross-level compaction into mvo_keypoint records
__global__ void __launch_bounds__(1024)
orb_finalize_kernel(const __grid_constant__ OrbGeom g, const unsigned long long* __restrict__ key_sorted,
                    const float2* __restrict__ ra_sorted, const int32_t* __restrict__ c2_count,
                    mvo_keypoint* __restrict__ kps, float2* __restrict__ kp_xy, int32_t* __restrict__ kp_count,
                    int32_t* __restrict__ flags, int32_t* __restrict__ occ) {
  const int b = blockIdx.x;
  __shared__ int s_keep[kLevels], s_off[kLevels + 1];
  __shared__ uint32_t s_grid[kGridMaxCells / 32];   // one bit per cell of the keypoint-distribution grid
  __shared__ int s_occ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grid_cells = g.grid_rows * g.grid_cols;
  for (int i = tid; i < kGridMaxCells / 32; i += blockDim.x) s_grid[i] = 0;
  if (tid == 0) s_occ = 0;
  if (warp < kLevels) {
    const LevelGeom lv = g.lv[warp];
    const int m = c2_count[b * kLevels + warp];
    const long long base = (long long)b * g.cand_total + lv.cand_off;
    int keep = m;
    if (lv.quota <= 0) keep = 0;
    else if (m > lv.quota) {
      // keys ascending == response descending; cut = response of element quota-1, ties stay
      const uint32_t cut = (uint32_t)(key_sorted[base + lv.quota - 1] >> 32);
      int cnt = 0;
      for (int i = lv.quota + lane; i < m; i += 32) cnt += ((uint32_t)(key_sorted[base + i] >> 32) == cut) ? 1 : 0;
      keep = lv.quota + warp_sum(cnt);
    }
    if (lane == 0) s_keep[warp] = keep;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int l = 0; l < kLevels; ++l) {
      s_off[l] = acc;
      acc += s_keep[l];
    }
    s_off[kLevels] = acc;
    if (acc > g.kp_cap) atomicOr(flags + b, 2);
    kp_count[b] = min(acc, g.kp_cap);
  }
  __syncthreads();
  for (int l = 0; l < kLevels; ++l) {
    const LevelGeom lv = g.lv[l];
    const long long base = (long long)b * g.cand_total + lv.cand_off;
    for (int i = tid; i < s_keep[l]; i += blockDim.x) {
      const int o = s_off[l] + i;
      if (o >= g.kp_cap) break;
      const unsigned long long k = key_sorted[base + i];
      const float2 r = ra_sorted[base + i];
      mvo_keypoint kp;
      kp.x = __fmul_rn((float)(int)(k & 0xffff), lv.scale);
      kp.y = __fmul_rn((float)(int)((k >> 16) & 0xffff), lv.scale);
      kp.size = __fmul_rn(31.f, lv.scale);
      kp.angle = r.y;
      kp.response = r.x;
      kp.octave = l;
      kp.class_id = -1;
      kps[(long long)b * g.kp_cap + o] = kp;
      kp_xy[(long long)b * g.kp_cap + o] = make_float2(kp.x, kp.y);
      if (g.grid_div > 0) {
        // int r = obs.keypoint.pt.y / occupancy_grid_div_, c = obs.keypoint.pt.x / occupancy_grid_div_ (float / int)
        const int r = (int)__fdiv_rn(kp.y, (float)g.grid_div), cc = (int)__fdiv_rn(kp.x, (float)g.grid_div);
        const int cell = r * g.grid_cols + cc;     // grid.at<uchar>(r, c) of a continuous rows x cols Mat
        if (cell >= 0 && cell < grid_cells) atomicOr(&s_grid[cell >> 5], 1u << (cell & 31));
      }
    }
  }
  if (occ) {
    __syncthreads();
    int cnt = 0;
    for (int i = tid; i < kGridMaxCells / 32; i += blockDim.x) cnt += __popc(s_grid[i]);
    cnt = warp_sum(cnt);
    if (lane == 0 && cnt) atomicAdd(&s_occ, cnt);
    __syncthreads();
    if (tid == 0) {
      occ[b * 2] = g.grid_div > 0 ? s_occ : -1;
      occ[b * 2 + 1] = grid_cells;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5: rotated BRIEF, warp per keypoint, lane = descriptor byte (16 samples)
constexpr int kBriefWarps = 8;
__global__ void __launch_bounds__(kBriefWarps * 32)
orb_brief_kernel(const __grid_constant__ OrbGeom g, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                 mvo_keypoint* __restrict__ kps, const int32_t* __restrict__ kp_count, uint8_t* __restrict__ desc,
                 uint8_t* __restrict__ valid, int compute_angle) {
  const int b = blockIdx.y;
  const int n = kp_count[b];
  const int i = blockIdx.x * kBriefWarps + (threadIdx.x >> 5);
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  mvo_keypoint kp = kps[(long long)b * g.kp_cap + i];
  const int l = min(max(kp.octave, 0), kLevels - 1);
  const LevelGeom lv = g.lv[l];
  const int cx = __float2int_rn(__fmul_rn(kp.x, lv.inv_scale));
  const int cy = __float2int_rn(__fmul_rn(kp.y, lv.inv_scale));
  if (compute_angle) {
    // IC angle of a detected keypoint (it lies >= 31 px inside its level) on the UNBLURRED level
    const uint8_t* img = pyr + (long long)b * g.frame_stride + lv.off;
    const int pitch = lv.pitch, x = cx, y = cy;
    // intensity centroid over the radius-15 disc: lane = column u + 15; fully unrolled so that the 31 row loads
    // of a lane are independent and in flight together (disc half-widths fold to constants)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
      constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
      const int u = lane - 15;
      const int au = abs(u);
      // (the disc is symmetric: column u spans the rows |v| <= kUmax[|u|]; one running row pointer instead of a 64-bit
      // multiply-add per row -- the address arithmetic was 30 % of the kernel's instructions)
      int vmax = 3;
#pragma unroll
      for (int k = 0; k < 15; ++k) vmax = (au == k) ? kUmax[k] : vmax;
      const uint8_t* q = img + (long long)(y - 15) * pitch + x + u;
      int colsum = 0;
#pragma unroll
      for (int v = -15; v <= 15; ++v) {
        if ((v < 0 ? -v : v) <= vmax) {
          const int p = *q;
          colsum += p;
          m01 += v * p;
        }
        q += pitch;
      }
      m10 = u * colsum;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);

    kp.angle = fast_atan2_deg((float)m01, (float)m10);
    if (lane == 0) kps[(long long)b * g.kp_cap + i].angle = kp.angle;
  }
  if (!desc) return;
  uint8_t* out = desc + ((long long)b * g.kp_cap + i) * 32;
  const bool ok = (kp.octave >= 0 && kp.octave < kLevels && cx >= kEdge && cx < lv.w - kEdge && cy >= kEdge &&
                   cy < lv.h - kEdge);
  if (valid && lane == 0) valid[(long long)b * g.kp_cap + i] = ok ? 1 : 0;
  if (!ok) {
    out[lane] = 0;
    return;
  }
  const float ang = __fmul_rn(kp.angle, __uint_as_float(0x3c8efa35u));  // (float)(CV_PI/180)
  double dsa, dca;
  sincos((double)ang, &dsa, &dca);
  const float ca = (float)dca, sa = (float)dsa;
  const uint8_t* img = blur + (long long)b * g.frame_stride + lv.off + (long long)cy * lv.pitch + cx;
  // this lane's 16 sample points (8 comparisons) = 32 bytes of the pattern
  const uint4* pp = reinterpret_cast<const uint4*>(&d_pattern[0][0]) + lane * 2;
  const uint4 pa = __ldg(pp), pb = __ldg(pp + 1);
  const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
  const int pitch = lv.pitch;
  int val[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const float px = (float)(signed char)(pw[j] >> (16 * s)), py = (float)(signed char)(pw[j] >> (16 * s + 8));
      const float rx = __fsub_rn(__fmul_rn(px, ca), __fmul_rn(py, sa));
      const float ry = __fadd_rn(__fmul_rn(px, sa), __fmul_rn(py, ca));
      val[2 * j + s] = img[__float2int_rn(ry) * pitch + __float2int_rn(rx)];
    }
  }
  unsigned byte = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) byte |= (val[2 * j] < val[2 * j + 1] ? 1u : 0u) << j;
  out[lane] = (uint8_t)byte;
}

// ================================================================================================
// host side
// keypoint-distribution grid of Initializer::good_keypoint_distribution (src/initializer.cpp:57-58)
void orb_set_grid(mvo_ctx* c, int w, int h) {
  OrbGeom& g = c->geom;
  const int div = c->occupancy_div;
  g.grid_div = 0;
  g.grid_rows = g.grid_cols = 0;
  if (div > 0 && w > 0 && h > 0) {
    const int rows = h / div, cols = w / div;
    if (rows > 0 && cols > 0 && rows * cols <= kGridMaxCells) {
      g.grid_div = div;
      g.grid_rows = rows;
      g.grid_cols = cols;
    }
  }
}

static int build_geometry(mvo_ctx* c, int w, int h, std::vector<uint32_t>& xt, std::vector<uint32_t>& yt) {
  OrbGeom& g = c->geom;
  const int n = c->cfg.nfeatures;
  // per-level quotas, all-float arithmetic as in the OpenCV source
  const double scale_factor = (double)1.2f;
  const float factor = (float)(1.0 / scale_factor);
  float nd = n * (1 - factor) / (1 - (float)pow((double)factor, (double)kLevels));
  int sum = 0;
  int quotas[kLevels];
  for (int l = 0; l < kLevels - 1; ++l) {
    quotas[l] = (int)lrintf(nd);
    sum += quotas[l];
    nd *= factor;
  }
  quotas[kLevels - 1] = std::max(n - sum, 0);

  long long off = 0;
  int cand_off = 0, xoff = 0, yoff = 0;
  xt.clear();
  yt.clear();
  for (int l = 0; l < kLevels; ++l) {
    LevelGeom& lv = g.lv[l];
    lv.scale = (float)pow(scale_factor, (double)l);
    lv.inv_scale = 1.f / lv.scale;
    lv.w = (int)lrintf((float)w * lv.inv_scale);
    lv.h = (int)lrintf((float)h * lv.inv_scale);
    if (lv.w < 1 || lv.h < 1 || lv.w > 65535 || lv.h > 65535) return MVO_ERR_INVALID;
    lv.pitch = (int)align_up((size_t)lv.w, 128);
    lv.quota = quotas[l];
    lv.off = off;
    off += (long long)lv.pitch * align_up((size_t)lv.h, 8);
    // FAST + NMS candidates per level and frame: 1 / 16 of the pixels covers natural images many times over; a frame of
    // pure sensor noise has ~10 % corners -- the detect call then doubles cand_scale and runs again (capi.cu)
    lv.cand_cap = std::max(4096, (int)std::min<long long>((long long)(lv.w * lv.h) / 16 * c->cand_scale, (long long)lv.w * lv.h / 2 + 4096));
    lv.cand_off = cand_off;
    cand_off += lv.cand_cap;
    lv.xtab_off = xoff;
    lv.ytab_off = yoff;
    if (l > 0) {
      const LevelGeom& pv = g.lv[l - 1];
      auto fill = [](std::vector<uint32_t>& t, int src, int dst) {
        const double scale = (double)src / (double)dst;
        for (int v = 0; v < dst; ++v) {
          const double f = scale * (v + 0.5) - 0.5;
          int i = (int)floor(f);
          int c1 = (int)lrint((f - i) * 256.0);
          if (i < 0) {
            i = 0;
            c1 = 0;
          }
          if (i >= src - 1) {
            i = src - 1;
            c1 = 0;
          }
          t.push_back((uint32_t)i | ((uint32_t)c1 << 16));
        }
      };
      fill(xt, pv.w, lv.w);
      fill(yt, pv.h, lv.h);
      // staging bounds of orb_level_kernel
      for (int t0 = 0; t0 < lv.w; t0 += TW) {
        const int lo = std::max(t0 - HALO, 0), hi = std::min(t0 + TW + HALO - 1, lv.w - 1);
        const int c0 = (int)(xt[xoff + lo] & 0xffff) & ~15, c1 = std::min((int)(xt[xoff + hi] & 0xffff) + 1, pv.w - 1);
        if ((c1 - c0) / 16 + 1 > SRC_COLS_MAX / 16) return MVO_ERR_UNSUPPORTED;
      }
      for (int t0 = 0; t0 < lv.h; t0 += TH) {
        const int lo = std::max(t0 - HALO, 0), hi = std::min(t0 + TH + HALO - 1, lv.h - 1);
        const int r0 = (int)(yt[yoff + lo] & 0xffff), r1 = std::min((int)(yt[yoff + hi] & 0xffff) + 1, pv.h - 1);
        if (r1 - r0 + 1 > SRC_ROWS_MAX) return MVO_ERR_UNSUPPORTED;
      }
      xoff += lv.w;
      yoff += lv.h;
    }
  }
  g.frame_stride = (long long)align_up((size_t)off + 256, 256);
  g.cand_total = cand_off;
  g.nfeatures = n;
  g.batch = c->cfg.batch;
  g.kp_cap = n + n / 4 + 64;
  orb_set_grid(c, w, h);
  return MVO_OK;
}

int orb_prepare(mvo_ctx* c, int w, int h) {
  if (w == c->geom_w && h == c->geom_h) return MVO_OK;
  if (w < 16 || h < 16 || w > c->cfg.max_width || h > c->cfg.max_height) {
    c->set_error("image size out of the range given to mvo_create");
    return MVO_ERR_INVALID;
  }
  std::vector<uint32_t> xt, yt;
  int rc = build_geometry(c, w, h, xt, yt);
  if (rc != MVO_OK) {
    c->set_error("unsupported pyramid geometry");
    return rc;
  }
  const OrbGeom& g = c->geom;
  const size_t B = (size_t)g.batch;
  MVO_CUDA_TRY(c, c->pyr.alloc(B * g.frame_stride));
  MVO_CUDA_TRY(c, c->blur.alloc(B * g.frame_stride));
  MVO_CUDA_TRY(c, c->xtab.alloc(xt.size()));
  MVO_CUDA_TRY(c, c->ytab.alloc(yt.size()));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->xtab.p, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->ytab.p, yt.data(), yt.size() * 4, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // xt / yt are stack vectors
  MVO_CUDA_TRY(c, c->cand_xy.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->cand_score.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->cand_sel.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->sel_count.alloc(B * kLevels));
  MVO_CUDA_TRY(c, c->cand_count.alloc(B * kLevels));
  MVO_CUDA_TRY(c, c->hist.alloc(B * kLevels * 256));
  MVO_CUDA_TRY(c, c->c2_key.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->c2_key_sorted.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->c2_ra.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->c2_ra_sorted.alloc(B * g.cand_total));
  MVO_CUDA_TRY(c, c->c2_count.alloc(B * kLevels));
  MVO_CUDA_TRY(c, c->kps.alloc(B * g.kp_cap));
  MVO_CUDA_TRY(c, c->kp_xy.alloc(B * g.kp_cap));
  MVO_CUDA_TRY(c, c->desc.alloc(B * g.kp_cap * 32));
  MVO_CUDA_TRY(c, c->kp_valid.alloc(B * g.kp_cap));
  MVO_CUDA_TRY(c, c->kp_count.alloc(B));
  MVO_CUDA_TRY(c, c->flags.alloc(B));
  MVO_CUDA_TRY(c, c->occ.alloc(B * 2));
  MVO_CUDA_TRY(c, cudaMemsetAsync(c->pyr.p, 0, B * g.frame_stride, c->stream));
  MVO_CUDA_TRY(c, cudaMemsetAsync(c->blur.p, 0, B * g.frame_stride, c->stream));
  c->geom_w = w;
  c->geom_h = h;
  return MVO_OK;
}

// rows of `w` gray bytes: (src, spitch, frame distance) -> (dst, dpitch, frame distance); 16 bytes per thread
__global__ void __launch_bounds__(256)
gray_rows_kernel(const uint8_t* __restrict__ src, int spitch, long long src_frame_stride, uint8_t* __restrict__ dst,
                 int dpitch, long long dst_frame_stride, int w, int h) {
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (x >= w) return;
  const uint8_t* s = src + (long long)b * src_frame_stride + (long long)y * spitch + x;
  uint8_t v[16];
  const int nb = min(16, w - x);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = i < nb ? s[i] : 0;
  *reinterpret_cast<uint4*>(dst + (long long)b * dst_frame_stride + (long long)y * dpitch + x) = *reinterpret_cast<const uint4*>(v);
}

// Host gray frames (batch x h rows of `stride` bytes) -> pitched device rows.  One flat DMA into a staging buffer plus
// an unpack kernel: a 2-D copy of 1241-byte rows from pageable memory runs far below PCIe speed (the synchronous ORB /
// LK calls spent a quarter of their time in it).
int upload_gray_rows(mvo_ctx* c, const uint8_t* host, int w, int h, int stride, int batch, uint8_t* dst, int dpitch,
                     long long dst_frame_stride) {
  const size_t fbytes = (size_t)h * stride;
  // the last row of the last frame ends after w bytes: a strided view (cv::Mat ROI) owns nothing behind it
  const size_t nbytes = fbytes * (batch - 1) + (size_t)(h - 1) * stride + (size_t)w;
  MVO_CUDA_TRY(c, c->img_in.alloc(fbytes * batch + 16));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->img_in.p, host, nbytes, cudaMemcpyHostToDevice, c->stream));
  dim3 grid((w + 16 * 256 - 1) / (16 * 256), h, batch);
  gray_rows_kernel<<<grid, 256, 0, c->stream>>>(c->img_in.p, stride, (long long)fbytes, dst, dpitch, dst_frame_stride, w, h);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

// images: batch frames, each h x stride bytes (host or device)
int orb_upload(mvo_ctx* c, const uint8_t* img, int w, int h, int stride, int channels, int on_device) {
  const OrbGeom& g = c->geom;
  const LevelGeom& l0 = g.lv[0];
  if (channels == 1) {
    if (!on_device) return upload_gray_rows(c, img, w, h, stride, g.batch, c->pyr.p + l0.off, l0.pitch, g.frame_stride);
    for (int b = 0; b < g.batch; ++b)
      MVO_CUDA_TRY(c, cudaMemcpy2DAsync(c->pyr.p + (size_t)b * g.frame_stride + l0.off, l0.pitch,
                                        img + (size_t)b * h * stride, stride, w, h, cudaMemcpyDeviceToDevice, c->stream));
    return MVO_OK;
  }
  if (channels != 3) return MVO_ERR_INVALID;
  const size_t fbytes = (size_t)h * stride;
  const uint8_t* src = img;
  if (!on_device) {
    const size_t nbytes = fbytes * (g.batch - 1) + (size_t)(h - 1) * stride + (size_t)w * 3;   // see upload_gray_rows
    MVO_CUDA_TRY(c, c->img_in.alloc(fbytes * g.batch));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(c->img_in.p, img, nbytes, cudaMemcpyHostToDevice, c->stream));
    src = c->img_in.p;
  }
  dim3 grid((w + 255) / 256, h, g.batch);
  bgr_to_gray_kernel<<<grid, 256, 0, c->stream>>>(src, stride, (long long)fbytes, c->pyr.p + l0.off, l0.pitch,
                                                  g.frame_stride, w, h);
  c->launches++;
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

static int launch_levels(mvo_ctx* c, int do_fast) {
  const OrbGeom& g = c->geom;
  for (int l = 0; l < kLevels; ++l) {
    const LevelGeom& lv = g.lv[l];
    LevelArgs a{};
    a.dst = c->pyr.p + lv.off;
    a.blur = c->blur.p + lv.off;
    a.frame_stride = g.frame_stride;
    a.w = lv.w;
    a.h = lv.h;
    a.pitch = lv.pitch;
    a.cand_xy = c->cand_xy.p;
    a.cand_score = c->cand_score.p;
    a.cand_count = c->cand_count.p;
    a.hist = c->hist.p;
    a.flags = c->flags.p;
    a.cand_cap = lv.cand_cap;
    a.cand_total = g.cand_total;
    a.cand_off = lv.cand_off;
    a.level = l;
    a.do_fast = do_fast;
    dim3 grid((lv.w + TW - 1) / TW, (lv.h + TH - 1) / TH, g.batch);
    if (l == 0) {
      orb_level_kernel<false><<<grid, kLevelThreads, 0, c->stream>>>(a);
    } else {
      const LevelGeom& pv = g.lv[l - 1];
      a.src = c->pyr.p + pv.off;
      a.sw = pv.w;
      a.sh = pv.h;
      a.spitch = pv.pitch;
      a.xtab = c->xtab.p + lv.xtab_off;
      a.ytab = c->ytab.p + lv.ytab_off;
      orb_level_kernel<true><<<grid, kLevelThreads, 0, c->stream>>>(a);
    }
    c->launches++;
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

static int clear_counters(mvo_ctx* c) {
  const size_t B = (size_t)c->geom.batch;
  MVO_CUDA_TRY(c, cudaMemsetAsync(c->cand_count.p, 0, B * kLevels * 4, c->stream));
  MVO_CUDA_TRY(c, cudaMemsetAsync(c->c2_count.p, 0, B * kLevels * 4, c->stream));
  MVO_CUDA_TRY(c, cudaMemsetAsync(c->hist.p, 0, B * kLevels * 256 * 4, c->stream));
  MVO_CUDA_TRY(c, cudaMemsetAsync(c->flags.p, 0, B * 4, c->stream));
  return MVO_OK;
}

int orb_run_levels_only(mvo_ctx* c) {
  int rc = clear_counters(c);
  if (rc) return rc;
  return launch_levels(c, 0);
}

// the eight fused level kernels exactly as orb_run_detect launches them (pyramid + FAST + NMS + blur), nothing after
int orb_run_levels_fast(mvo_ctx* c) {
  int rc = clear_counters(c);
  if (rc) return rc;
  return launch_levels(c, 1);
}

int orb_run_detect(mvo_ctx* c, bool want_desc) {
  const OrbGeom& g = c->geom;
  int rc = clear_counters(c);
  if (rc) return rc;
  if (!c->capturing) cudaEventRecord(c->timers[9].beg, c->stream);   // "orb_dense": the eight fused level kernels
  rc = launch_levels(c, 1);
  if (rc) return rc;
  if (!c->capturing) {
    cudaEventRecord(c->timers[9].end, c->stream);
    c->timers[9].used = true;
  }
  int max_cap = 0;
  for (int l = 0; l < kLevels; ++l) max_cap = std::max(max_cap, g.lv[l].cand_cap);
  {
    dim3 gsel(kLevels, g.batch);
    orb_select_kernel<<<gsel, kSelThreads, 0, c->stream>>>(g, c->cand_xy.p, c->cand_score.p, c->cand_count.p, c->hist.p,
                                                          c->cand_sel.p, c->sel_count.p);
    c->launches++;
    // warp per selected candidate; level 0 keeps ~2 n_0 (+ ties): size the grid for that, grid-stride beyond
    const int want = (2 * g.lv[0].quota + g.lv[0].quota / 4 + 64 + kHarrisPerWarp - 1) / kHarrisPerWarp;
    dim3 grid(std::max(1, std::min((want + kHarrisWarps - 1) / kHarrisWarps, (max_cap + kHarrisWarps - 1) / kHarrisWarps)), kLevels,
              g.batch);
    orb_harris_angle_kernel<<<grid, kHarrisWarps * 32, 0, c->stream>>>(g, c->pyr.p, c->cand_sel.p, c->sel_count.p,
                                                                      c->c2_key.p, c->c2_ra.p, c->c2_count.p);
    c->launches++;
  }
  {
    const int want = 2 * g.lv[0].quota + g.lv[0].quota / 4 + 64;
    dim3 grid(std::min((want + kRankThreads - 1) / kRankThreads, (max_cap + kRankThreads - 1) / kRankThreads), kLevels, g.batch);
    orb_rank_kernel<<<grid, kRankThreads, 0, c->stream>>>(g, c->c2_key.p, c->c2_ra.p, c->c2_count.p,
                                                         c->c2_key_sorted.p, c->c2_ra_sorted.p);
    c->launches++;
  }
  orb_finalize_kernel<<<g.batch, 1024, 0, c->stream>>>(g, c->c2_key_sorted.p, c->c2_ra_sorted.p, c->c2_count.p,
                                                      c->kps.p, c->kp_xy.p, c->kp_count.p, c->flags.p, c->occ.p);
  c->launches++;
  {
    // IC angle of the surviving keypoints, then (want_desc) their rBRIEF descriptors
    dim3 grid((g.kp_cap + kBriefWarps - 1) / kBriefWarps, g.batch);
    orb_brief_kernel<<<grid, kBriefWarps * 32, 0, c->stream>>>(g, c->pyr.p, c->blur.p, c->kps.p, c->kp_count.p,
                                                              want_desc ? c->desc.p : nullptr, nullptr, 1);
    c->launches++;
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

// descriptors for keypoints already uploaded into c->kps (stream 0), count n in c->kp_count
int orb_run_brief_given(mvo_ctx* c, int n) {
  const OrbGeom& g = c->geom;
  dim3 grid((n + kBriefWarps - 1) / kBriefWarps, 1);
  if (n > 0) {
    orb_brief_kernel<<<grid, kBriefWarps * 32, 0, c->stream>>>(g, c->pyr.p, c->blur.p, c->kps.p, c->kp_count.p, c->desc.p,
                                                              c->kp_valid.p, 0);
    c->launches++;
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

}  // namespace mvo
