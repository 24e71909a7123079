// lk.cu -- pyramidal Lucas-Kanade tracking (cv::calcOpticalFlowPyrLK, all defaults) for B200.
//
// Replaces the call at /root/reference/src/tracker.cpp:68-69.  Contract: SURVEY.md A.3 / oracle/lk_oracle.py
// (status identical, positions within 0.01 px of OpenCV; here the integer sums are exact, which is tighter
// than OpenCV's own float accumulation).
//
//   lk_pyrdown_tile_kernel  integer 5x5 [1 4 6 4 1]^2 pyrDown: word loads, packed 16-bit vertical pass (lk_pyrdown_kernel:
//                      the byte-per-thread version, for images narrower than a tile).
//   lk_track_kernel    one warp per point, all pyramid levels inside the kernel (no inter-point dependency):
//                      24x24 raw patch in smem -> Scharr derivatives computed on the fly (no derivative
//                      image is ever written: saves 5.3 P0 bytes of dense traffic per frame) -> Q14 bilinear
//                      patch / gradient values in shared memory -> 2x2 normal matrix by warp reduction ->
//                      <= 30 Gauss-Newton iterations sampling J from a 32x32 smem region (texture-free
//                      bilinear), re-staged only when the window leaves it.
#include "context.cuh"
#include "host_hash.hpp"
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace mvo {

constexpr int LKW = 21;
constexpr int kLkWarps = 4;
#ifndef MVO_LK_MINB
#define MVO_LK_MINB 7
#endif
constexpr int kRaw = 24;      // raw patch side: 22 interp rows + 1 Scharr ring
constexpr int kJReg = 32;     // staged J region side
constexpr int kJMargin = 5;
constexpr int kWin = LKW * LKW;          // 441 window pixels
constexpr int kSlots = (kWin + 31) / 32; // 14 window pixels per lane

struct LkLevel {
  int w, h, pitch;
  long long off;
};
struct LkGeom {
  LkLevel lv[kLkLevels];
  int nlevels;
  long long frame_stride;
};

// ---- pyrDown --------------------------------------------------------------------------------
constexpr int PDW = 32, PDH = 8;   // output tile
__global__ void __launch_bounds__(PDW * PDH)
lk_pyrdown_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw,
                  int dh, int dpitch, long long frame_stride_src, long long frame_stride_dst) {
  __shared__ uint8_t tile[(2 * PDH + 3)][2 * PDW + 4];
  __shared__ uint16_t hsum[(2 * PDH + 3)][PDW];
  const int b = blockIdx.z;
  src += (long long)b * frame_stride_src;
  dst += (long long)b * frame_stride_dst;
  const int ox0 = blockIdx.x * PDW, oy0 = blockIdx.y * PDH;
  const int tid = threadIdx.y * PDW + threadIdx.x;
  const int sx0 = 2 * ox0 - 2, sy0 = 2 * oy0 - 2;
  for (int i = tid; i < (2 * PDH + 3) * (2 * PDW + 3); i += PDW * PDH) {
    const int r = i / (2 * PDW + 3), c = i - r * (2 * PDW + 3);
    const int y = min(reflect101(sy0 + r, sh), sh - 1), x = min(reflect101(sx0 + c, sw), sw - 1);
    tile[r][c] = src[(long long)max(y, 0) * spitch + max(x, 0)];
  }
  __syncthreads();
  for (int i = tid; i < (2 * PDH + 3) * PDW; i += PDW * PDH) {
    const int r = i / PDW, c = i - r * PDW;
    const uint8_t* t = &tile[r][2 * c];
    hsum[r][c] = (uint16_t)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
  if (ox < dw && oy < dh) {
    const int r = 2 * threadIdx.y, c = threadIdx.x;
    const int v = hsum[r][c] + 4 * hsum[r + 1][c] + 6 * hsum[r + 2][c] + 4 * hsum[r + 3][c] + hsum[r + 4][c];
    dst[(long long)oy * dpitch + ox] = (uint8_t)((v + 128) >> 8);
  }
}

// The same filter with word-sized traffic (the byte-per-thread kernel above spends ~150 instructions per output pixel and
// is kept for images narrower than one tile): 64 x 16 output tile, source rows staged with aligned 32-bit loads (row
// indices reflected at load time, the at most two reflected columns on either image edge patched in shared memory),
// horizontal pass two outputs per item from three words, vertical pass on packed 16-bit pairs (5-tap sums of 5-tap sums
// stay below 2^16), four outputs per 32-bit store.
constexpr int PD2W = 64, PD2H = 16;
constexpr int PD2_ROWS = 2 * PD2H + 3;            // 35 source rows
constexpr int PD2_WORDS = (2 * PD2W + 8) / 4;     // 34 words: source columns 2 ox0 - 4 .. 2 ox0 + 131
constexpr int PD2_STRIDE = PD2_WORDS + 1;         // smem row stride in words
__global__ void __launch_bounds__(256)
lk_pyrdown_tile_kernel(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dw,
                       int dh, int dpitch, long long frame_stride_src, long long frame_stride_dst) {
  __shared__ uint32_t tile[PD2_ROWS * PD2_STRIDE];
  __shared__ __align__(8) uint16_t hs[PD2_ROWS * PD2W];
  const int b = blockIdx.z, tid = threadIdx.x;
  src += (long long)b * frame_stride_src;
  dst += (long long)b * frame_stride_dst;
  const int ox0 = blockIdx.x * PD2W, oy0 = blockIdx.y * PD2H;
  const int sx0 = 2 * ox0 - 4, sy0 = 2 * oy0 - 2;   // sx0 is a multiple of 4
  for (int i = tid; i < PD2_ROWS * PD2_WORDS; i += 256) {
    const int r = i / PD2_WORDS, wq = i - r * PD2_WORDS;
    const int y = min(max(reflect101(sy0 + r, sh), 0), sh - 1);
    const int x = sx0 + 4 * wq;
    uint32_t v = 0;
    if (x >= 0 && x < spitch) v = __ldg(reinterpret_cast<const uint32_t*>(src + (long long)y * spitch + x));
    tile[r * PD2_STRIDE + wq] = v;
  }
  __syncthreads();
  // REFLECT_101 columns: -2, -1 on the left edge, sw, sw + 1 on the right edge (nothing further out feeds a stored output)
  uint8_t* tb = reinterpret_cast<uint8_t*>(tile);
  if (sx0 < 0 || sx0 + 4 * PD2_WORDS > sw) {
    for (int i = tid; i < PD2_ROWS * 4; i += 256) {
      const int r = i >> 2, k = i & 3;
      const int col = (k < 2) ? (k - 2) : (sw + k - 2);          // -2, -1, sw, sw + 1
      const int from = min(max(reflect101(col, sw), 0), sw - 1);
      if (col >= sx0 && col < sx0 + 4 * PD2_WORDS && from >= sx0 && from < sx0 + 4 * PD2_WORDS)
        tb[r * PD2_STRIDE * 4 + col - sx0] = tb[r * PD2_STRIDE * 4 + from - sx0];
    }
    __syncthreads();
  }
  // horizontal [1 4 6 4 1]: outputs 2 cp and 2 cp + 1 of row r from the words cp, cp + 1, cp + 2
  for (int i = tid; i < PD2_ROWS * (PD2W / 2); i += 256) {
    const int r = i / (PD2W / 2), cp = i - r * (PD2W / 2);
    const uint32_t* t = tile + r * PD2_STRIDE + cp;
    const uint32_t w0 = t[0], w1 = t[1], w2 = t[2];
    const uint32_t b2 = (w0 >> 16) & 0xff, b3 = w0 >> 24, b4 = w1 & 0xff, b5 = (w1 >> 8) & 0xff, b6 = (w1 >> 16) & 0xff,
                   b7 = w1 >> 24, b8 = w2 & 0xff;
    const uint32_t h0 = b2 + b6 + 4 * (b3 + b5) + 6 * b4, h1 = b4 + b8 + 4 * (b5 + b7) + 6 * b6;
    *reinterpret_cast<uint32_t*>(hs + r * PD2W + 2 * cp) = h0 | (h1 << 16);
  }
  __syncthreads();
  // vertical [1 4 6 4 1] on packed pairs, + 128 >> 8: four outputs per thread
  {
    const int oyl = tid >> 4, g = tid & 15;
    const int oy = oy0 + oyl, ox = ox0 + 4 * g;
    if (oy < dh && ox < dw) {
      const uint2* hp = reinterpret_cast<const uint2*>(hs + (2 * oyl) * PD2W + 4 * g);
      const uint2 a = hp[0], bq = hp[PD2W / 4], c = hp[2 * (PD2W / 4)], d = hp[3 * (PD2W / 4)], e = hp[4 * (PD2W / 4)];
      const uint32_t lo = a.x + e.x + 4 * (bq.x + d.x) + 6 * c.x + 0x00800080u;   // per 16-bit lane <= 65280 + 128
      const uint32_t hi = a.y + e.y + 4 * (bq.y + d.y) + 6 * c.y + 0x00800080u;
      const uint32_t px = __byte_perm(lo, hi, 0x7531);                             // the high byte of every lane == >> 8
      uint8_t* o = dst + (long long)oy * dpitch + ox;
      if (ox + 3 < dw) {
        *reinterpret_cast<uint32_t*>(o) = px;
      } else {
        for (int k = 0; k < dw - ox; ++k) o[k] = (uint8_t)(px >> (8 * k));
      }
    }
  }
}

// ---- tracking -------------------------------------------------------------------------------
// Per-warp scratch.  The two uses never overlap in time: the template patch (I, Ix|Iy) is written by the fused
// setup (one 64-bit store per pixel) and read back into registers before the first J region of the level is staged.
struct LkWarpSmem {
  uint32_t jq[kJReg * kJReg];        // J "quads": (J[y][x], J[y][x+1], J[y+1][x], J[y+1][x+1]) per position
  uint2 patch[kSlots * 32];          // template window: x = Ix (low 16, signed) | Iy (high 16, signed), y = I (Q5)
};

__device__ __forceinline__ int safe_reflect(int i, int n) { return min(max(reflect101(i, n), 0), n - 1); }

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
  const float s = 16384.f;
  const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
  w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), s));
  w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), s));
  w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), s));
  w11 = 16384 - w00 - w01 - w10;
}

// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1   (signed 16-bit weights, unsigned pixels): IDP.2A.LO.S16.U8
__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// d = c + a.lo16 * b.byte2 + a.hi16 * b.byte3
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_w(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

// Stage the 32x32 J region with top-left (jx0, jy0) as quads.  Kept out of line: the kernel is instruction-cache bound
// otherwise (ncu: 26 % no-instruction stalls with everything unrolled and inlined at three call sites).
// Region inside the image (the common case, warp-uniform branch): lane = (row group of eight quad rows, aligned
// 4-pixel word column).  A lane walks its nine pixel rows with two aligned 32-bit loads per row, cuts the five
// pixels it needs out of the eight loaded bytes with two PRMT whose selectors carry the misalignment of jx0, and
// stores four quads per row with one 128-bit STS: ~11 instructions per 4 quads instead of ~12 per quad for the
// byte-per-lane walk (which remains for regions that touch the border and reflect).
__device__ __noinline__ void lk_stage_j(uint32_t* jq, const uint8_t* __restrict__ J, int jx0, int jy0, int w, int h,
                                        int pitch, int lane) {
  if (jx0 >= 0 && jx0 + kJReg <= w && jy0 >= 0 && jy0 + kJReg <= h) {
    const int g = lane >> 3, wi = lane & 7;
    const uint32_t sh = (uint32_t)(jx0 & 3);
    const uint32_t sel01 = 0x2110u + 0x1111u * sh, sel23 = 0x4332u + 0x1111u * sh;
    const uint8_t* p = J + (jy0 + 8 * g) * pitch + (jx0 & ~3) + 4 * wi;
    // pixel row jy0 + 32 feeds only quad row 31, which no window reads; it may lie below the image: clamp it
    const int rmax = h - 1 - (jy0 + 8 * g);
    uint4* out = reinterpret_cast<uint4*>(jq + 8 * g * kJReg + 4 * wi);
    uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(p)), w1 = __ldg(reinterpret_cast<const uint32_t*>(p + 4));
    uint32_t t01 = __byte_perm(w0, w1, sel01), t23 = __byte_perm(w0, w1, sel23);   // pixel pairs (0,1)(1,2) / (2,3)(3,4)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint8_t* pr = p + min(k + 1, rmax) * pitch;
      w0 = __ldg(reinterpret_cast<const uint32_t*>(pr));
      w1 = __ldg(reinterpret_cast<const uint32_t*>(pr + 4));
      const uint32_t b01 = __byte_perm(w0, w1, sel01), b23 = __byte_perm(w0, w1, sel23);
      out[k * (kJReg / 4)] = make_uint4(__byte_perm(t01, b01, 0x5410), __byte_perm(t01, b01, 0x7632),
                                        __byte_perm(t23, b23, 0x5410), __byte_perm(t23, b23, 0x7632));
      t01 = b01;
      t23 = b23;
    }
    return;
  }
  // lane = column of the region, reflecting row / column walk
  const uint8_t* p = J + safe_reflect(jx0 + lane, w);
  uint32_t pprev = 0;
#pragma unroll 1
  for (int r0 = 0; r0 < kJReg; r0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[safe_reflect(jy0 + r0 + k, h) * pitch];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t pr = v[k] | (__shfl_down_sync(0xffffffffu, v[k], 1) << 8);
      if (r0 + k > 0) jq[(r0 + k - 1) * kJReg + lane] = pprev | (pr << 16);
      pprev = pr;
    }
  }
}

// Fused template setup of one window on one image plane: lane = raw column (x = ix - 1 + lane, 24 columns), the 24
// raw rows are walked in registers: Scharr (3,10,3) derivatives from shuffled neighbours, Q14 bilinear interpolation
// of I / Ix / Iy.  Results go to pt (x = Ix | Iy << 16, y = I) in pixel order k = r * 21 + c; the normal-matrix sums are
// taken by the caller when it reads the window back (14 samples per lane instead of 24 rows per lane).
template <bool kFastPath = true>
__device__ __forceinline__ void lk_setup_patch(const uint8_t* __restrict__ I, int ix, int iy, int w, int h, int pitch,
                                               int lane, int w00, int w01, int w10, int w11, uint2* pt_out) {
  const uint32_t Wt = pack_w(w00, w01), Wb = pack_w(w10, w11);
  const int colx = ix - 1 + lane;
  const int xr = safe_reflect(colx, w);
  const bool der_col = lane >= 1 && lane <= kRaw - 2 && colx >= 0 && colx < w;  // derivative column inside the image
  const bool win_col = lane >= 1 && lane <= LKW;                                 // window column c = lane - 1
  int dx0 = 0, dx1 = 0, sm0 = 0, sm1 = 0;
  uint32_t pair0 = 0, pair1 = 0, pd_prev = 0, pdn_prev = 0;
  // rows in groups of four independent loads; the walk itself is rolled (code size: see lk_stage_j).  row_off / row_ok
  // carry the (warp-uniform) border handling: plain strides when the 32 x 24 footprint lies inside the image.
  auto walk = [&](const uint8_t* p, auto row_off, auto row_ok) {
#pragma unroll 1
    for (int r0 = 0; r0 < kRaw; r0 += 4) {
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = p[row_off(r0 + k)];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + k;
        const int c = (int)v[k];
        const int lv_ = __shfl_up_sync(0xffffffffu, c, 1), rv = __shfl_down_sync(0xffffffffu, c, 1);
        const int dx2 = rv - lv_, sm2 = 3 * (lv_ + rv) + 10 * c;
        const uint32_t pair2 = (uint32_t)c | ((uint32_t)rv << 8);
        // derivative row rr = r - 2 (y = iy + rr); meaningful from r = 2 on, harmless before (the results are not stored)
        const int rr = r - 2;
        const int gx = 3 * (dx0 + dx2) + 10 * dx1, gy = sm2 - sm0;
        const uint32_t pd = (der_col && row_ok(rr)) ? pack_w(gx, gy) : 0u;
        const uint32_t pdn = __shfl_down_sync(0xffffffffu, pd, 1);
        // window row pr = rr - 1
        const int pr = rr - 1;
        const int iv = dp2a_lo_su(Wb, pair1, dp2a_lo_su(Wt, pair0, 1 << 8)) >> 9;
        const int x00 = (int)(short)(pd_prev & 0xffffu), x01 = (int)(short)(pdn_prev & 0xffffu);
        const int x10 = (int)(short)(pd & 0xffffu), x11 = (int)(short)(pdn & 0xffffu);
        const int y00 = (int)pd_prev >> 16, y01 = (int)pdn_prev >> 16, y10 = (int)pd >> 16, y11 = (int)pdn >> 16;
        const int vx = (x00 * w00 + x01 * w01 + x10 * w10 + x11 * w11 + (1 << 13)) >> 14;
        const int vy = (y00 * w00 + y01 * w01 + y10 * w10 + y11 * w11 + (1 << 13)) >> 14;
        if (win_col && pr >= 0) {
          pt_out[pr * LKW + lane - 1] = make_uint2(pack_w(vx, vy), (uint32_t)iv);
        }
        pd_prev = pd;
        pdn_prev = pdn;
        dx0 = dx1;
        dx1 = dx2;
        sm0 = sm1;
        sm1 = sm2;
        pair0 = pair1;
        pair1 = pair2;
      }
    }
  };
  if (kFastPath && ix >= 1 && ix - 1 + 32 <= w && iy >= 1 && iy - 1 + kRaw <= h) {
    // Footprint inside the image (the common case): no derivative is masked, so interpolation and Scharr commute --
    // everything is exact integer arithmetic up to the final shifts.  Interpolate the raw patch first
    // (T = sum w * raw, two IDP.2A per value), then take the Scharr of T: sum w * der == Scharr(T).  One shuffled
    // value per row instead of packed derivative pairs, no 8-multiply interpolation of Ix / Iy.
    const uint8_t* p = I + (iy - 1) * pitch + colx;
    int t1 = 0;   // T of the previous row
#pragma unroll 1
    for (int r0 = 0; r0 < kRaw; r0 += 4) {
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = p[(r0 + k) * pitch];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + k;                                   // raw row; T row r - 2; window row pr = r - 3
        const uint32_t c = v[k];
        const uint32_t rv = __shfl_down_sync(0xffffffffu, c, 1);
        const uint32_t pair2 = c | (rv << 8);
        const int tn = dp2a_lo_su(Wb, pair2, dp2a_lo_su(Wt, pair1, 0));
        const int tl = __shfl_up_sync(0xffffffffu, tn, 1), tr = __shfl_down_sync(0xffffffffu, tn, 1);
        const int dx2 = tr - tl, sm2 = 3 * (tl + tr) + 10 * tn;
        const int pr = r - 3;
        const int vx = (3 * (dx0 + dx2) + 10 * dx1 + (1 << 13)) >> 14;
        const int vy = (sm2 - sm0 + (1 << 13)) >> 14;
        const int iv = (t1 + (1 << 8)) >> 9;
        if (win_col && pr >= 0) {
          pt_out[pr * LKW + lane - 1] = make_uint2(pack_w(vx, vy), (uint32_t)iv);
        }
        dx0 = dx1;
        dx1 = dx2;
        sm0 = sm1;
        sm1 = sm2;
        t1 = tn;
        pair1 = pair2;
      }
    }
  } else {
    walk(I + xr, [&](int r) { return safe_reflect(iy - 1 + r, h) * pitch; },
         [&](int rr) { return iy + rr >= 0 && iy + rr < h; });
  }
}

__global__ void __launch_bounds__(kLkWarps * 32, MVO_LK_MINB)
lk_track_kernel(const __grid_constant__ LkGeom g, const uint8_t* __restrict__ pyrI, const uint8_t* __restrict__ pyrJ,
                const float2* __restrict__ pts, const int32_t* __restrict__ npts_dev, int max_pts,
                float2* __restrict__ out_pts, uint8_t* __restrict__ status, float* __restrict__ err) {
  __shared__ LkWarpSmem sm_all[kLkWarps];
  // window pixel k = lane + 32 j  ->  byte offset of its quad inside the staged J region.  A shared table read with one
  // LDS per sample: fourteen offsets per lane do not fit beside the template in 96 registers, and rebuilding them
  // (k / 21, k % 21) costs five integer instructions per sample in the iteration loop.
  __shared__ uint16_t qtab[kSlots * 32];
  for (int k = threadIdx.x; k < kSlots * 32; k += kLkWarps * 32) {
    const int r = k / LKW, c = k - r * LKW;
    qtab[k] = (uint16_t)((r * kJReg + c) * 4);
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int n = min(npts_dev[b], max_pts);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kLkWarps + warp;
  if (i >= n) return;
  LkWarpSmem& sm = sm_all[warp];
  const float2 p0 = pts[(long long)b * max_pts + i];
  float nx = 0.f, ny = 0.f, e = 0.f;
  int st = 1;
  const float flt_scale = 1.f / (1 << 20);
  const uint16_t* qt = qtab + lane;

  for (int L = g.nlevels - 1; L >= 0; --L) {
    const LkLevel lv = g.lv[L];
    const uint8_t* I = pyrI + (long long)b * g.frame_stride + lv.off;
    const uint8_t* J = pyrJ + (long long)b * g.frame_stride + lv.off;
    const int w = lv.w, h = lv.h, pitch = lv.pitch;
    const float s = 1.f / (float)(1 << L);
    const float ppx = __fmul_rn(p0.x, s), ppy = __fmul_rn(p0.y, s);
    if (L == g.nlevels - 1) {
      nx = ppx;
      ny = ppy;
    } else {
      nx = __fmul_rn(nx, 2.f);
      ny = __fmul_rn(ny, 2.f);
    }
    const float qx = __fsub_rn(ppx, 10.f), qy = __fsub_rn(ppy, 10.f);
    const int ix = (int)floorf(qx), iy = (int)floorf(qy);
    // (a NaN coordinate floors to INT_MIN in OpenCV: out of range on every level)
    if (ix < -LKW || ix >= w || iy < -LKW || iy >= h || qx != qx || qy != qy) {
      if (L == 0) {
        st = 0;
        e = 0.f;
      }
      continue;
    }
    int w00, w01, w10, w11;
    lk_weights(__fsub_rn(qx, (float)ix), __fsub_rn(qy, (float)iy), w00, w01, w10, w11);

    // ---- fused template setup (lk_setup_patch): the window lands in smem in pixel order k = r * 21 + c
    int sA11 = 0, sA12 = 0, sA22 = 0;
    __syncwarp();   // every lane is done with the previous level's template
    lk_setup_patch(I, ix, iy, w, h, pitch, lane, w00, w01, w10, w11, sm.patch);
    __syncwarp();
    // normal-matrix sums (exact integers) over the window pixels owned by this lane (pixel k = lane + 32 j).  The
    // template itself stays in shared memory: 42 registers less per thread buy two more CTAs per SM.
    const uint2* tp = sm.patch + lane;
#pragma unroll
    for (int j = 0; j < kSlots; ++j) {
      if (lane + 32 * j < kWin) {
        const uint32_t d = tp[32 * j].x;
        const int gx = (int)(short)(d & 0xffffu), gy = (int)d >> 16;
        sA11 += gx * gx;
        sA12 += gx * gy;
        sA22 += gy * gy;
      }
    }
    const float A11 = __fmul_rn((float)warp_sum_wide(sA11), flt_scale);
    const float A12 = __fmul_rn((float)warp_sum_wide(sA12), flt_scale);
    const float A22 = __fmul_rn((float)warp_sum_wide(sA22), flt_scale);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * LKW * LKW));
    if ((double)min_eig < 1e-4 || D < 1.1920929e-07f) {
      if (L == 0) st = 0;
      continue;
    }
    D = __fdiv_rn(1.f, D);
    float cx = __fsub_rn(nx, 10.f), cy = __fsub_rn(ny, 10.f);
    float pdx = 0.f, pdy = 0.f;
    int jx0 = -100000, jy0 = -100000;
    for (int it = 0; it < 30; ++it) {
      const int inx = (int)floorf(cx), iny = (int)floorf(cy);
      if (inx < -LKW || inx >= w || iny < -LKW || iny >= h) {
        if (L == 0) st = 0;
        break;
      }
      if (inx < jx0 || inx - jx0 > kJReg - LKW - 1 || iny < jy0 || iny - jy0 > kJReg - LKW - 1) {
        jx0 = inx - kJMargin;
        jy0 = iny - kJMargin;
        __syncwarp();
        lk_stage_j(sm.jq, J, jx0, jy0, w, h, pitch, lane);
        __syncwarp();
      }
      int v00, v01, v10, v11;
      lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), v00, v01, v10, v11);
      const uint32_t Vt = pack_w(v00, v01), Vb = pack_w(v10, v11);
      const uint32_t* jb = sm.jq + (iny - jy0) * kJReg + (inx - jx0);
      int sb1 = 0, sb2 = 0;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        if (lane + 32 * j < kWin) {
          const uint32_t q = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(jb) + qt[32 * j]);
          const uint2 tv = tp[32 * j];
          const int diff = (dp2a_hi_su(Vb, q, dp2a_lo_su(Vt, q, 1 << 8)) >> 9) - (int)tv.y;
          sb1 += diff * (int)(short)(tv.x & 0xffffu);
          sb2 += diff * ((int)tv.x >> 16);
        }
      }
      const float b1 = __fmul_rn((float)warp_sum_wide(sb1), flt_scale);
      const float b2 = __fmul_rn((float)warp_sum_wide(sb2), flt_scale);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      cx = __fadd_rn(cx, dx);
      cy = __fadd_rn(cy, dy);
      nx = __fadd_rn(cx, 10.f);
      ny = __fadd_rn(cy, 10.f);
      if ((double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01) break;
      if (it > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
        nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f));
        ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
        break;
      }
      pdx = dx;
      pdy = dy;
    }
    if (L == 0 && st) {
      const float fx = __fsub_rn(nx, 10.f), fy = __fsub_rn(ny, 10.f);
      const int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -LKW || inx >= w || iny < -LKW || iny >= h) {
        st = 0;
      } else {
        if (inx < jx0 || inx - jx0 > kJReg - LKW - 1 || iny < jy0 || iny - jy0 > kJReg - LKW - 1) {
          jx0 = inx - kJMargin;
          jy0 = iny - kJMargin;
          __syncwarp();
          lk_stage_j(sm.jq, J, jx0, jy0, w, h, pitch, lane);
          __syncwarp();
        }
        int v00, v01, v10, v11;
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), v00, v01, v10, v11);
        const uint32_t Vt = pack_w(v00, v01), Vb = pack_w(v10, v11);
        const uint32_t* jb = sm.jq + (iny - jy0) * kJReg + (inx - jx0);
        int se = 0;
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
          if (lane + 32 * j < kWin) {
            const uint32_t q = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(jb) + qt[32 * j]);
            se += abs((dp2a_hi_su(Vb, q, dp2a_lo_su(Vt, q, 1 << 8)) >> 9) - (int)tp[32 * j].y);
          }
        }
        se = __reduce_add_sync(0xffffffffu, se);
        e = __fdiv_rn((float)se, (float)(32 * LKW * LKW));
      }
    }
  }
  if (lane == 0) {
    const long long o = (long long)b * max_pts + i;
    out_pts[o] = make_float2(nx, ny);
    status[o] = (uint8_t)st;
    err[o] = e;
  }
}


// ---- lk_track2_kernel: TMA-staged tiles, vertical sample runs, shuffle-free setup --------------------------------
// Second generation of the gray tracker (the first one, lk_track_kernel above, stays in the library as the cross-check:
// mvo_debug_set(ctx, "lk_impl", 1) selects it at run time; both produce bit-identical results).
//
//   * Tiles by TMA.  The two tiles a level needs -- the raw 24 x 24 template patch around the point in the previous image
//     and the 30 x 29 pixel region of the next image the window samples -- are 2-D byte windows at arbitrary alignment:
//     one cp.async.bulk.tensor (a 32 x 29 box of a per-level 3-D tensor map {x, y, stream}) per tile, issued by one lane,
//     completion on a per-warp mbarrier.  The tile lands with its top-left pixel at byte 0, so no alignment arithmetic is
//     left in the kernel, no registers are held by loads in flight, and both loads are issued ahead of use: the template
//     positions depend only on the input point, so the patch of level L - 1 is prefetched while level L iterates; the J
//     region of a level is requested before its template is computed (ncu r02b: a third of the first version's stall
//     samples were the staging loads).
//   * The 441 window samples are owned as 63 vertical RUNS of seven pixels (run = 3 * column + row third); lane L owns
//     runs L and L + 32.  Template setup needs no shuffle: every lane walks its own 10 raw rows x 4 raw columns per run,
//     interpolates the 3 x 9 T values it needs itself (6 IDP.2A per row) and takes the Scharr of T in registers.
//   * The J quads of a run sit at base + i * pitch: immediate offsets, no offset table, and with a quad pitch of 29 words
//     (bank = x - 3 y mod 32) the 32 lanes of a sample slot hit 32 different banks (the 32-word pitch of the first kernel
//     is 1.7-way conflicted: ncu r01t, 40 % of its shared wavefronts).
//   * Code size is part of the design: the first unrolled version of this kernel (47 KB) thrashed the 32 KB instruction
//     cache (ncu r02b: 95.7 % hit rate, GPC instruction fetch at 72 % of peak); the two runs of the setup share one body.
constexpr int kJW = 29;                 // quad columns == pitch in words
// The border / restaging helpers of the second-generation tracker are inlined.  They used to be out of line (code size);
// with thirteen arguments (some passed on the stack) and 50 - 70 bytes of registers saved around the calls, two builds
// of the kernel -- the BGR8 team form at 96 registers, and the gray form once two more 64-bit sums were live across the
// calls -- returned wrong positions for points with border levels while every build without those saves was correct
// (same source at 128 registers; this inlined form).  Inlined, the kernel needs no spills at 96 registers and is 6 %
// faster (0.647 -> 0.607 ms per 32 x 2000 points).
#ifndef MVO_LK_OUTLINE
#define MVO_LK_OUTLINE __forceinline__
#endif
constexpr int kJH = 28;                 // quad rows
constexpr int kJSlackX = kJW - LKW;     // 8: inx - jx0 in [0, 8]
constexpr int kJSlackY = kJH - LKW;     // 7
constexpr int kJMarginX = 4, kJMarginY = 3;
constexpr int kRuns = 63, kRunLen = 7;
// TMA boxes: the innermost coordinate of a tiled copy must be 16-byte aligned (anything else faults with "illegal
// instruction" on sm_100a: scratch/tma_probe.cu), so a tile starts at x & ~15 and is 48 bytes wide: 15 + 30 columns.
constexpr int kTileW = 48, kTileWords = kTileW / 4;
constexpr int kTileHJ = kJH + 1;        // next-image tile: 29 rows
constexpr int kTileHI = kRaw;           // template tile: 24 rows

#ifndef MVO_LK2_TI_SMEM
#define MVO_LK2_TI_SMEM 1               // 1: template intensities stay in shared memory (14 registers less), 0: registers
#endif

struct __align__(128) LkWarpSmem2 {
  // TMA destinations are 128-byte aligned.  The quad builder reads one word past the end of its tile (into jq: harmless).
  uint32_t rawI[kTileHI * kTileWords];       // template tile: 1152 B
  union {
    struct {
      uint32_t rawJ[352];                    // next-image tile: 29 rows x 48 bytes = 1392 B, padded to 1408
      uint32_t jq[kJH * kJW];                // quads of the staged J region: 3248 B
    };
    uint2 patch[kWin];                       // border setup: template in pixel order (3528 B), gathered into tx / ti; no J
  };                                         // tile is in flight then
  uint32_t tx[2 * kRunLen * 32];             // template Ix | Iy << 16, slot-major: [k][lane], k = 7 * run_of_lane + i
  int32_t ti[2 * kRunLen * 32];              // template I (Q5)
  unsigned long long bar[2];                 // mbarriers: [0] template tile, [1] J tile
};
static_assert(sizeof(LkWarpSmem2) == 9472, "six 4-warp CTAs per SM");
static_assert(kTileHJ * kTileWords <= 352 && kTileHI * kTileWords * 4 % 128 == 0, "tile buffers");
struct LkTmaps {
  CUtensorMap m[2][kLkLevels];               // [0]: previous image (template), [1]: next image; one map per level
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// (bounded: a tile that never arrives -- a broken tensor map -- traps instead of hanging the GPU)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  for (int spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1 << 24)) __trap();
}
// one box (48 bytes x `rows`) of stream b at (x, y), x a multiple of 16: global -> shared, completion (box bytes) on
// `bar`.  Issued by one lane.
__device__ __forceinline__ void tma_tile(void* dst, const CUtensorMap* map, int x, int y, int b, unsigned long long* bar,
                                         int rows) {
  // generic-proxy accesses to the destination (this warp's earlier reads) are ordered before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(bar, kTileW * rows);
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(b), "r"(smem_u32(bar))
      : "memory");
}

// quads of the staged tile: lane = (row phase g = lane >> 3, word column wi = lane & 7), quad rows g, g + 4, ...: with the
// 29-word pitch the 32 stores of a step go to 32 different banks, and the four word loads are conflict-free too.
__device__ __forceinline__ void lk_build_quads(uint32_t* jq, const uint32_t* rawJ, int lane, int sub) {
  const int g = lane >> 3, wi = lane & 7;
  const uint32_t* p = rawJ + g * kTileWords + (sub >> 2) + wi;   // sub = jx0 & 15: column of the region inside the tile
  const uint32_t sh = (uint32_t)(sub & 3);
  const uint32_t sel01 = 0x2110u + 0x1111u * sh, sel23 = 0x4332u + 0x1111u * sh;
  uint32_t* out = jq + g * kJW + 4 * wi;
#pragma unroll
  for (int k = 0; k < kJH / 4; ++k) {
    const uint32_t w0 = p[4 * k * kTileWords], w1 = p[4 * k * kTileWords + 1];
    const uint32_t w2 = p[(4 * k + 1) * kTileWords], w3 = p[(4 * k + 1) * kTileWords + 1];
    const uint32_t t01 = __byte_perm(w0, w1, sel01), t23 = __byte_perm(w0, w1, sel23);
    const uint32_t b01 = __byte_perm(w2, w3, sel01), b23 = __byte_perm(w2, w3, sel23);
    uint32_t* o = out + 4 * k * kJW;
    o[0] = __byte_perm(t01, b01, 0x5410);
    if (wi < 7) {   // quads 29 .. 31 do not exist
      o[1] = __byte_perm(t01, b01, 0x7632);
      o[2] = __byte_perm(t23, b23, 0x5410);
      o[3] = __byte_perm(t23, b23, 0x7632);
    }
  }
}

// J region that touches the image border: lane = pixel column, reflecting row / column walk straight from global memory
__device__ MVO_LK_OUTLINE void lk_stage_j2_border(uint32_t* jq, const uint8_t* __restrict__ J, int jx0, int jy0, int w, int h,
                                                int pitch, int lane) {
  const uint8_t* p = J + safe_reflect(jx0 + lane, w);
  uint32_t pprev = 0;
#pragma unroll 1
  for (int r0 = 0; r0 < 32; r0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[safe_reflect(jy0 + r0 + k, h) * pitch];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t pr = v[k] | (__shfl_down_sync(0xffffffffu, v[k], 1) << 8);
      const int qr = r0 + k - 1;
      if (qr >= 0 && qr < kJH && lane < kJW) jq[qr * kJW + lane] = pprev | (pr << 16);
      pprev = pr;
    }
  }
}

// Border windows: the first kernel's row walk writes the template in pixel order (it handles reflected image rows /
// columns and zeroed derivatives outside the image); out of line, rare.
__device__ MVO_LK_OUTLINE void lk_setup_border(const uint8_t* __restrict__ I, int ix, int iy, int w, int h, int pitch, int lane,
                                             int w00, int w01, int w10, int w11, uint2* pt_out) {
  lk_setup_patch<false>(I, ix, iy, w, h, pitch, lane, w00, w01, w10, w11, pt_out);
}

// One run of the interior setup: raw rows 7 rseg .. 7 rseg + 9, raw columns c .. c + 3 of the staged tile (its top-left
// pixel is (ix - 1, iy - 1)) -> seven template samples (window rows 7 rseg + i, column c) into this lane's slots.
__device__ __forceinline__ void lk_setup_run(const uint32_t* __restrict__ raw, int c, int rseg, uint32_t Wt, uint32_t Wb,
                                             uint32_t* __restrict__ tx, int32_t* __restrict__ ti) {
  // c: byte column inside the tile (window column + (ix - 1) & 15)
  const uint32_t* rp = raw + (kRunLen * rseg) * kTileWords + (c >> 2);
  const uint32_t fs = (uint32_t)(c & 3) * 8u;
  uint32_t Wp = 0, Mp = 0;
  int tcp = 0, dx0 = 0, dx1 = 0, sm0 = 0, sm1 = 0;
#pragma unroll
  for (int j = 0; j < kRunLen + 3; ++j) {
    const uint32_t W = __funnelshift_r(rp[j * kTileWords], rp[j * kTileWords + 1], fs);   // raw bytes c .. c + 3 of row j
    const uint32_t M = __byte_perm(W, 0u, 0x4421);                                          // (b1, b2) in the low half
    if (j >= 1) {
      const int tl = dp2a_lo_su(Wb, W, dp2a_lo_su(Wt, Wp, 0));
      const int tc = dp2a_lo_su(Wb, M, dp2a_lo_su(Wt, Mp, 0));
      const int tr = dp2a_hi_su(Wb, W, dp2a_hi_su(Wt, Wp, 0));
      const int dx2 = tr - tl, sm2 = 3 * (tl + tr) + 10 * tc;
      if (j >= 3) {
        const int vx = (3 * (dx0 + dx2) + 10 * dx1 + (1 << 13)) >> 14;
        const int vy = (sm2 - sm0 + (1 << 13)) >> 14;
        tx[(j - 3) * 32] = __byte_perm((uint32_t)vx, (uint32_t)vy, 0x5410);
        ti[(j - 3) * 32] = (tcp + (1 << 8)) >> 9;
      }
      dx0 = dx1;
      dx1 = dx2;
      sm0 = sm1;
      sm1 = sm2;
      tcp = tc;
    }
    Wp = W;
    Mp = M;
  }
}

// Out of line (code size): wait for the requested tile and / or fetch a new one so that the quads cover the window at
// (inx, iny).  Returns the region origin and the parity of the J barrier.
__device__ MVO_LK_OUTLINE int3 lk_restage(LkWarpSmem2& sm, const CUtensorMap* map, const uint8_t* __restrict__ J, int inx,
                                        int iny, int w, int h, int pitch, int lane, int b, bool jpend, int jx0, int jy0,
                                        uint32_t parJ) {
  bool have_tile = jpend;   // a tile is in flight (it covers the start position of the level by construction)
  if (!(jpend) || inx < jx0 || inx - jx0 > kJSlackX || iny < jy0 || iny - jy0 > kJSlackY) {
    if (jpend) {            // in flight but useless: drain it before the buffer is reused
      mbar_wait(&sm.bar[1], parJ);
      parJ ^= 1;
    }
    jx0 = inx - kJMarginX;
    jy0 = iny - kJMarginY;
    __syncwarp();           // every lane is done with the quads / the tile
    have_tile = jx0 >= 0 && jx0 + kJW + 1 <= w && jy0 >= 0 && jy0 + kTileHJ <= h;
    if (have_tile && lane == 0) tma_tile(sm.rawJ, map, jx0 & ~15, jy0, b, &sm.bar[1], kTileHJ);
  }
  if (have_tile) {
    mbar_wait(&sm.bar[1], parJ);
    parJ ^= 1;
    lk_build_quads(sm.jq, sm.rawJ, lane, jx0 & 15);
  } else {
    lk_stage_j2_border(sm.jq, J, jx0, jy0, w, h, pitch, lane);
  }
  __syncwarp();
  return make_int3(jx0, jy0, (int)parJ);
}

#ifndef MVO_LK2_MINB
#define MVO_LK2_MINB 5
#endif
#ifndef MVO_LK2_WARPS
#define MVO_LK2_WARPS 4
#endif
constexpr int kLk2Warps = MVO_LK2_WARPS;

// PERSIST: the grid is a fixed number of resident CTAs and every warp draws points (work items: the points of all
// streams, stream-major) from a global counter until none is left.  Points differ in their iteration counts and a CTA
// stays resident until its slowest warp is done, so with one point per warp the achieved occupancy was 23 % of a
// theoretical 31 %; drawing work keeps every warp slot busy to the end of the kernel.
constexpr int kLk2MaxBatch = 1024;   // streams of a group the persistent form indexes (larger groups: one point per warp)
// CN = 3 (BGR8, what the node feeds: /root/reference/src/mono_vo.cpp:94): OpenCV's window then spans the three colour
// planes -- every sum of the gray algorithm (the 2 x 2 normal matrix, the mismatch vector of each iteration, the L1 error)
// runs over 21 x 21 x 3 values and nothing else changes.  A TEAM of three warps tracks one point, one warp per plane with
// exactly the gray code path (its own TMA tiles, template, quads); at the three reduction points the warps exchange their
// exact integer partial sums through shared memory behind a named barrier, so all three take the same decisions and
// iterate in lock step.  (The first-generation lk_track_cn_kernel keeps all three planes in one warp: 23 KB of shared
// memory per warp, 8 resident warps per SM, 3.5 ms per 32 x 2000 points; it stays as the cross-check, lk_impl = 1.)
constexpr int kLk2WarpsCn = 3;   // CN = 3: one team per CTA (static shared memory stays below 48 KB)
// (five CTAs per SM = 128 registers.  Six CTAs at 112 registers (__maxnreg__) are no faster: 2.43 against 2.38 ms -- the
// kernel is issue-bound like the gray one.  A 96-register build with the helpers out of line returned wrong positions at
// border points, see MVO_LK_OUTLINE; tests/test_gpu_lk.py compares this kernel with the first-generation one on
// border-heavy point sets so that a regression of that kind is seen.)
#ifndef MVO_LK_T_MINB
#define MVO_LK_T_MINB 5
#endif
// MUL = 3 (with CN = 1): a BGR8 stream whose three planes are identical in both frames -- a gray camera behind the node's
// BGR8 conversion (/root/reference/src/mono_vo.cpp:94), which is what a KITTI gray sequence becomes.  Every sum over the
// three planes is then exactly three times the sum over one: the gray code path runs on plane 0 (plane index 3 b) and
// multiplies its exact integer sums by three where the team kernel would have added three equal partial sums -- the same
// integers, so the same floats afterwards.  colour_a / colour_b (per stream, nonzero = some pixel of that frame has
// differing channels; written by the kernels that split BGR8 frames into planes) select the streams of a launch:
// want_colour = 1 the team kernel's, 0 the MUL = 3 kernel's; null = every stream.
template <bool PERSIST, int CN, int MUL = 1>
__global__ void __launch_bounds__((CN == 1 ? kLk2Warps : kLk2WarpsCn) * 32, CN == 1 ? MVO_LK2_MINB : MVO_LK_T_MINB)
lk_track2_kernel(const __grid_constant__ LkGeom g, const __grid_constant__ LkTmaps tm, const uint8_t* __restrict__ pyrI,
                 const uint8_t* __restrict__ pyrJ, const float2* __restrict__ pts, const int32_t* __restrict__ npts_dev,
                 int max_pts, float2* __restrict__ out_pts, uint8_t* __restrict__ status, float* __restrict__ err,
                 int* __restrict__ work_counter, int batch, const int32_t* __restrict__ colour_a,
                 const int32_t* __restrict__ colour_b, int want_colour) {
  static_assert(MUL == 1 || (MUL == 3 && CN == 1), "MUL = 3: the gray path on identical BGR planes");
  constexpr int PLANES = CN * MUL;   // planes per stream in the pyramid buffers
  auto selected = [&](int q) { return !colour_a || (((colour_a[q] | colour_b[q]) != 0) == (want_colour != 0)); };
  constexpr int NW = CN == 1 ? kLk2Warps : kLk2WarpsCn;
  static_assert(CN == 1 || (CN == 3 && NW % CN == 0 && PERSIST), "teams of CN warps, persistent form only");
  __shared__ LkWarpSmem2 sm_all[NW];
  __shared__ int s_cum[PERSIST ? kLk2MaxBatch + 1 : 1];   // s_cum[b] = points of the streams before b
  __shared__ long long s_team[CN == 1 ? 1 : NW / CN][2][CN][3];   // team exchange: [team][slot parity][plane][value]
  __shared__ int s_item[CN == 1 ? 1 : NW / CN][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = CN == 1 ? 0 : warp / CN, ch = CN == 1 ? 0 : warp - team * CN;
  uint32_t tpar = 0, fpar = 0;
  auto team_barrier = [&]() {
    if (CN > 1) {
      __syncwarp();   // (bar.sync is an aligned barrier: the warp must arrive converged)
      asm volatile("barrier.sync %0, %1;" ::"r"(1 + team), "r"(32 * CN) : "memory");
    }
  };
  // exact integer sums over the planes of the team (identity for gray); every warp of the team gets the same totals
  auto team_sum3 = [&](long long& v0, long long& v1, long long& v2) {
    if (CN == 1) {
      if (MUL != 1) {
        v0 *= MUL;
        v1 *= MUL;
        v2 *= MUL;
      }
      return;
    }
    if (lane == 0) {
      s_team[team][tpar][ch][0] = v0;
      s_team[team][tpar][ch][1] = v1;
      s_team[team][tpar][ch][2] = v2;
    }
    team_barrier();
    v0 = v1 = v2 = 0;
#pragma unroll
    for (int q = 0; q < CN; ++q) {
      v0 += s_team[team][tpar][q][0];
      v1 += s_team[team][tpar][q][1];
      v2 += s_team[team][tpar][q][2];
    }
    tpar ^= 1;   // two slots: a warp can only reach its next-but-one exchange after everybody has read this one
  };
  LkWarpSmem2& sm = sm_all[warp];
  if (lane == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (PERSIST) {
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int q = 0; q < batch; ++q) {
        s_cum[q] = acc;
        acc += selected(q) ? min(npts_dev[q], max_pts) : 0;
      }
      s_cum[batch] = acc;
    }
    __syncthreads();
  }
  const int total_items = PERSIST ? s_cum[batch] : 0;
  uint32_t parI = 0, parJ = 0;
  const float flt_scale = 1.f / (1 << 20);
  // the two runs of this lane: run -> (column c, row third rseg); lane 31 has no second run
  const int c0 = lane / 3, r0s = lane - 3 * c0;
  const bool has1 = lane + 32 < kRuns;
  const int run1 = has1 ? lane + 32 : kRuns - 1;
  const int c1 = run1 / 3, r1s = run1 - 3 * c1;
  const int qoff0 = (kRunLen * r0s) * kJW + c0, qoff1 = (kRunLen * r1s) * kJW + c1;
  uint32_t* const txl = sm.tx + lane;
  int32_t* const til = sm.ti + lane;
  for (;;) {
  int b, i;
  if (PERSIST) {
    int item = 0;
    if (CN == 1) {
      if (lane == 0) item = atomicAdd(work_counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
    } else {
      // two slots: a point can end without a single exchange (every level out of range), so the drawing warp may be
      // back here before a team mate has read the previous item
      if (ch == 0 && lane == 0) s_item[team][fpar] = atomicAdd(work_counter, 1);
      team_barrier();
      item = s_item[team][fpar];
      fpar ^= 1;
    }
    if (item >= total_items) break;
    int lo = 0, hi = batch;                 // largest b with s_cum[b] <= item
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_cum[mid] <= item) lo = mid;
      else hi = mid;
    }
    b = lo;
    i = item - s_cum[b];
  } else {
    b = blockIdx.y;
    i = blockIdx.x * kLk2Warps + warp;
    if (i >= min(npts_dev[b], max_pts) || !selected(b)) return;
  }
  const float2 p0 = pts[(long long)b * max_pts + i];
  const int plane = b * PLANES + ch;     // planar pyramids: plane (b * PLANES + ch); also the z coordinate of the tensor maps
  float nx = 0.f, ny = 0.f, e = 0.f;
  int st = 1;

  // template tile of a level: position and fast-path test depend on the input point only
  auto tmpl_pos = [&](int L, int& ix, int& iy, float& qx, float& qy) {
    const float s = 1.f / (float)(1 << L);
    qx = __fsub_rn(__fmul_rn(p0.x, s), 10.f);
    qy = __fsub_rn(__fmul_rn(p0.y, s), 10.f);
    ix = (int)floorf(qx);
    iy = (int)floorf(qy);
    const int w = g.lv[L].w, h = g.lv[L].h;
    // (a NaN coordinate floors to INT_MIN in OpenCV: out of range on every level)
    const bool in_range = !(ix < -LKW || ix >= w || iy < -LKW || iy >= h || qx != qx || qy != qy);
    const bool interior = in_range && ix >= 1 && ix + 23 <= w && iy >= 1 && iy + 23 <= h;
    return in_range ? (interior ? 2 : 1) : 0;
  };
  bool ireq = false;   // the template tile of the coming level has been requested
  {
    int ix, iy;
    float qx, qy;
    ireq = tmpl_pos(g.nlevels - 1, ix, iy, qx, qy) == 2;
    if (ireq && lane == 0) tma_tile(sm.rawI, &tm.m[0][g.nlevels - 1], (ix - 1) & ~15, iy - 1, plane, &sm.bar[0], kTileHI);
  }

  for (int L = g.nlevels - 1; L >= 0; --L) {
    const LkLevel lv = g.lv[L];
    const uint8_t* I = pyrI + (long long)plane * g.frame_stride + lv.off;
    const uint8_t* J = pyrJ + (long long)plane * g.frame_stride + lv.off;
    const int w = lv.w, h = lv.h, pitch = lv.pitch;
    if (L == g.nlevels - 1) {
      const float s = 1.f / (float)(1 << L);
      nx = __fmul_rn(p0.x, s);
      ny = __fmul_rn(p0.y, s);
    } else {
      nx = __fmul_rn(nx, 2.f);
      ny = __fmul_rn(ny, 2.f);
    }
    int ix, iy;
    float qx, qy;
    const int kind = tmpl_pos(L, ix, iy, qx, qy);
    if (kind == 0) {
      if (L == 0) {
        st = 0;
        e = 0.f;
      }
      ireq = false;
      continue;    // (tmpl_pos is monotone over the levels: no tile was requested for this one)
    }
    // ---- request the J region around the start position of this level (it is needed after the template setup) ----
    float cx = __fsub_rn(nx, 10.f), cy = __fsub_rn(ny, 10.f);
    int jx0 = -100000, jy0 = -100000;
    bool jpend = false;
    {
      const int inx = (int)floorf(cx), iny = (int)floorf(cy);
      const int tx0 = inx - kJMarginX, ty0 = iny - kJMarginY;
      // (a border template builds its window in scratch that aliases the J tile: no request ahead of it then)
      if (kind == 2 && tx0 >= 0 && tx0 + kJW + 1 <= w && ty0 >= 0 && ty0 + kTileHJ <= h) {
        jx0 = tx0;
        jy0 = ty0;
        jpend = true;
        __syncwarp();   // every lane is done with the previous level's J tile and quads
        if (lane == 0) tma_tile(sm.rawJ, &tm.m[1][L], jx0 & ~15, jy0, plane, &sm.bar[1], kTileHJ);
      }
    }
    int w00, w01, w10, w11;
    lk_weights(__fsub_rn(qx, (float)ix), __fsub_rn(qy, (float)iy), w00, w01, w10, w11);

    // ---- template of this level: tx / ti slots of this lane ----
    if (kind == 2) {
      if (!ireq) {   // (a level above was out of range: nothing was prefetched)
        __syncwarp();
        if (lane == 0) tma_tile(sm.rawI, &tm.m[0][L], (ix - 1) & ~15, iy - 1, plane, &sm.bar[0], kTileHI);
      }
      mbar_wait(&sm.bar[0], parI);
      parI ^= 1;
      const uint32_t Wt = pack_w(w00, w01), Wb = pack_w(w10, w11);
      const int sub = (ix - 1) & 15;
#pragma unroll 1
      for (int t = 0; t < 2; ++t)
        lk_setup_run(sm.rawI, sub + (t ? c1 : c0), t ? r1s : r0s, Wt, Wb, txl + t * kRunLen * 32, til + t * kRunLen * 32);
    } else {
      __syncwarp();   // patch aliases the quads of the previous level
      lk_setup_border(I, ix, iy, w, h, pitch, lane, w00, w01, w10, w11, sm.patch);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < kRunLen; ++k) {
        const uint2 a = sm.patch[(kRunLen * r0s + k) * LKW + c0], bq = sm.patch[(kRunLen * r1s + k) * LKW + c1];
        txl[k * 32] = a.x;
        til[k * 32] = (int)a.y;
        txl[(kRunLen + k) * 32] = bq.x;
        til[(kRunLen + k) * 32] = (int)bq.y;
      }
    }
    // prefetch the template tile of the next level (the tile buffer has been consumed)
    ireq = false;
    if (L > 0) {
      int nix, niy;
      float nqx, nqy;
      ireq = tmpl_pos(L - 1, nix, niy, nqx, nqy) == 2;
      __syncwarp();
      if (ireq && lane == 0) tma_tile(sm.rawI, &tm.m[0][L - 1], (nix - 1) & ~15, niy - 1, plane, &sm.bar[0], kTileHI);
    }
    uint32_t txy[2 * kRunLen];
#pragma unroll
    for (int k = 0; k < 2 * kRunLen; ++k) txy[k] = (k >= kRunLen && !has1) ? 0u : txl[k * 32];   // lane 31: no 2nd run
#if !MVO_LK2_TI_SMEM
    int tir[2 * kRunLen];
#pragma unroll
    for (int k = 0; k < 2 * kRunLen; ++k) tir[k] = til[k * 32];
#define LK2_TI(k) tir[k]
#else
#define LK2_TI(k) til[(k) * 32]
#endif
    // The mismatch vector of an iteration is sum (J - I) g = sum J g - sum I g over the window, in exact integers: the
    // second sum does not change while the level iterates, so it is taken once here and the iteration loop neither
    // loads the template intensities nor subtracts them (two of its ten instructions per sample).
    int sA11 = 0, sA12 = 0, sA22 = 0, sI1 = 0, sI2 = 0;
#pragma unroll
    for (int k = 0; k < 2 * kRunLen; ++k) {
      const int gx = (int)(short)(txy[k] & 0xffffu), gy = (int)txy[k] >> 16;
      sA11 += gx * gx;
      sA12 += gx * gy;
      sA22 += gy * gy;
      const int ti = LK2_TI(k);   // |I|, |J| <= 255 * 32 and |g| <= 16 * 255 (Scharr of u8): a lane's 14 products stay below 2^29
      sI1 += ti * gx;
      sI2 += ti * gy;
    }
    long long tA11 = warp_sum_wide(sA11), tA12 = warp_sum_wide(sA12), tA22 = warp_sum_wide(sA22);
    const long long tI1 = warp_sum_wide(sI1), tI2 = warp_sum_wide(sI2);
    team_sum3(tA11, tA12, tA22);

    const float A11 = __fmul_rn((float)tA11, flt_scale);
    const float A12 = __fmul_rn((float)tA12, flt_scale);
    const float A22 = __fmul_rn((float)tA22, flt_scale);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * LKW * LKW));
    const bool degenerate = (double)min_eig < 1e-4 || D < 1.1920929e-07f;
    if (degenerate && L == 0) st = 0;
    D = __fdiv_rn(1.f, D);
    float pdx = 0.f, pdy = 0.f;

    // (re)stage the J region so that it covers the window at integer position (inx, iny)
    auto cover = [&](int inx, int iny) {
      if (jpend || inx < jx0 || inx - jx0 > kJSlackX || iny < jy0 || iny - jy0 > kJSlackY) {
        const int3 r = lk_restage(sm, &tm.m[1][L], J, inx, iny, w, h, pitch, lane, plane, jpend, jx0, jy0, parJ);
        jx0 = r.x;
        jy0 = r.y;
        parJ = (uint32_t)r.z;
        jpend = false;
      }
    };

    for (int it = 0; it < 30 && !degenerate; ++it) {
      const int inx = (int)floorf(cx), iny = (int)floorf(cy);
      if (inx < -LKW || inx >= w || iny < -LKW || iny >= h) {
        if (L == 0) st = 0;
        break;
      }
      cover(inx, iny);
      int v00, v01, v10, v11;
      lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), v00, v01, v10, v11);
      const uint32_t Vt = pack_w(v00, v01), Vb = pack_w(v10, v11);
      const uint32_t* jb = sm.jq + (iny - jy0) * kJW + (inx - jx0);
      const uint32_t* q0 = jb + qoff0;
      const uint32_t* q1 = jb + qoff1;
      int sb1 = 0, sb2 = 0;
#pragma unroll
      for (int k = 0; k < kRunLen; ++k) {
        const uint32_t qa = q0[k * kJW], qb = q1[k * kJW];
        const int ja = dp2a_hi_su(Vb, qa, dp2a_lo_su(Vt, qa, 1 << 8)) >> 9;
        const int jb = dp2a_hi_su(Vb, qb, dp2a_lo_su(Vt, qb, 1 << 8)) >> 9;
        sb1 += ja * (int)(short)(txy[k] & 0xffffu) + jb * (int)(short)(txy[kRunLen + k] & 0xffffu);
        sb2 += ja * ((int)txy[k] >> 16) + jb * ((int)txy[kRunLen + k] >> 16);
      }
      long long tb1 = warp_sum_wide(sb1) - tI1, tb2 = warp_sum_wide(sb2) - tI2, tb3 = 0;
      team_sum3(tb1, tb2, tb3);
      const float b1 = __fmul_rn((float)tb1, flt_scale);
      const float b2 = __fmul_rn((float)tb2, flt_scale);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      cx = __fadd_rn(cx, dx);
      cy = __fadd_rn(cy, dy);
      nx = __fadd_rn(cx, 10.f);
      ny = __fadd_rn(cy, 10.f);
      if ((double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01) break;
      if (it > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
        nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f));
        ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
        break;
      }
      pdx = dx;
      pdy = dy;
    }
    if (L == 0 && st) {
      const float fx = __fsub_rn(nx, 10.f), fy = __fsub_rn(ny, 10.f);
      const int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -LKW || inx >= w || iny < -LKW || iny >= h) {
        st = 0;
      } else {
        cover(inx, iny);
        int v00, v01, v10, v11;
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), v00, v01, v10, v11);
        const uint32_t Vt = pack_w(v00, v01), Vb = pack_w(v10, v11);
        const uint32_t* jb = sm.jq + (iny - jy0) * kJW + (inx - jx0);
        const uint32_t* q0 = jb + qoff0;
        const uint32_t* q1 = jb + qoff1;
        int se = 0;
#pragma unroll
        for (int k = 0; k < kRunLen; ++k) {
          const uint32_t qa = q0[k * kJW], qb = q1[k * kJW];
          se += abs((dp2a_hi_su(Vb, qa, dp2a_lo_su(Vt, qa, 1 << 8)) >> 9) - LK2_TI(k));
          const int eb = abs((dp2a_hi_su(Vb, qb, dp2a_lo_su(Vt, qb, 1 << 8)) >> 9) - LK2_TI(kRunLen + k));
          se += has1 ? eb : 0;
        }
        long long te = __reduce_add_sync(0xffffffffu, se), te1 = 0, te2 = 0;
        team_sum3(te, te1, te2);
        e = __fdiv_rn((float)(int)te, (float)(32 * LKW * PLANES * LKW));
      }
    }
    if (jpend) {   // a requested tile nobody consumed (degenerate level, window left the image): drain it
      mbar_wait(&sm.bar[1], parJ);
      parJ ^= 1;
    }
#undef LK2_TI
  }
  if (lane == 0 && ch == 0) {
    const long long o = (long long)b * max_pts + i;
    out_pts[o] = make_float2(nx, ny);
    status[o] = (uint8_t)st;
    err[o] = e;
  }
  if (!PERSIST) break;
  __syncwarp();
  }   // work loop
}

// ---- multi-channel windows (the node feeds BGR8: /root/reference/src/mono_vo.cpp:94, src/tracker.cpp:68) -------
// OpenCV's window then spans all channels: every sum runs over 21 x 21 x cn values, only err is normalised by cn.
// Planar layout: plane (b * CN + ch) of the pyramid buffer.  The template stays in smem (3 x 441 values do not fit
// in registers); otherwise the same arithmetic and control flow as lk_track_kernel.
constexpr int kLkWarpsCn = 2;
template <int CN>
struct LkWarpSmemCn {
  uint32_t jq[CN][kJReg * kJReg];
  uint2 t[CN][kSlots * 32];            // x = Ix | Iy << 16, y = I
};

template <int CN>
__global__ void __launch_bounds__(kLkWarpsCn * 32)
lk_track_cn_kernel(const __grid_constant__ LkGeom g, const uint8_t* __restrict__ pyrI, const uint8_t* __restrict__ pyrJ,
                   const float2* __restrict__ pts, const int32_t* __restrict__ npts_dev, int max_pts,
                   float2* __restrict__ out_pts, uint8_t* __restrict__ status, float* __restrict__ err) {
  static_assert(CN <= 3, "the 32-bit per-lane accumulators below are sized for at most three channels");
  __shared__ LkWarpSmemCn<CN> sm_all[kLkWarpsCn];
  const int b = blockIdx.y;
  const int n = min(npts_dev[b], max_pts);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kLkWarpsCn + warp;
  if (i >= n) return;
  LkWarpSmemCn<CN>& sm = sm_all[warp];
  const float2 p0 = pts[(long long)b * max_pts + i];
  float nx = 0.f, ny = 0.f, e = 0.f;
  int st = 1;
  const float flt_scale = 1.f / (1 << 20);
  int qoff[kSlots];
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int k = lane + 32 * j;
    const int r = k / LKW, c = k - r * LKW;
    qoff[j] = r * kJReg + c;
  }
  for (int L = g.nlevels - 1; L >= 0; --L) {
    const LkLevel lv = g.lv[L];
    const uint8_t* I = pyrI + (long long)b * CN * g.frame_stride + lv.off;
    const uint8_t* J = pyrJ + (long long)b * CN * g.frame_stride + lv.off;
    const int w = lv.w, h = lv.h, pitch = lv.pitch;
    const float s = 1.f / (float)(1 << L);
    const float ppx = __fmul_rn(p0.x, s), ppy = __fmul_rn(p0.y, s);
    if (L == g.nlevels - 1) {
      nx = ppx;
      ny = ppy;
    } else {
      nx = __fmul_rn(nx, 2.f);
      ny = __fmul_rn(ny, 2.f);
    }
    const float qx = __fsub_rn(ppx, 10.f), qy = __fsub_rn(ppy, 10.f);
    const int ix = (int)floorf(qx), iy = (int)floorf(qy);
    // (a NaN coordinate floors to INT_MIN in OpenCV: out of range on every level)
    if (ix < -LKW || ix >= w || iy < -LKW || iy >= h || qx != qx || qy != qy) {
      if (L == 0) {
        st = 0;
        e = 0.f;
      }
      continue;
    }
    int w00, w01, w10, w11;
    lk_weights(__fsub_rn(qx, (float)ix), __fsub_rn(qy, (float)iy), w00, w01, w10, w11);
    int tA11 = 0, tA12 = 0, tA22 = 0;   // per-lane sums over CN <= 3 channels stay below 2^31 (<= 21 * 4080^2 per channel)
    __syncwarp();
#pragma unroll 1
    for (int ch = 0; ch < CN; ++ch)
      lk_setup_patch(I + (long long)ch * g.frame_stride, ix, iy, w, h, pitch, lane, w00, w01, w10, w11, sm.t[ch]);
    __syncwarp();
#pragma unroll 1
    for (int ch = 0; ch < CN; ++ch)
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        const int k = lane + 32 * j;
        if (k < kWin) {
          const uint32_t d = sm.t[ch][k].x;
          const int gx = (int)(short)(d & 0xffffu), gy = (int)d >> 16;
          tA11 += gx * gx;
          tA12 += gx * gy;
          tA22 += gy * gy;
        }
      }
    const float A11 = __fmul_rn((float)warp_sum_wide(tA11), flt_scale);
    const float A12 = __fmul_rn((float)warp_sum_wide(tA12), flt_scale);
    const float A22 = __fmul_rn((float)warp_sum_wide(tA22), flt_scale);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * LKW * LKW));
    if ((double)min_eig < 1e-4 || D < 1.1920929e-07f) {
      if (L == 0) st = 0;
      continue;
    }
    D = __fdiv_rn(1.f, D);
    float cx = __fsub_rn(nx, 10.f), cy = __fsub_rn(ny, 10.f);
    float pdx = 0.f, pdy = 0.f;
    int jx0 = -100000, jy0 = -100000;
    for (int it = 0; it < 30; ++it) {
      const int inx = (int)floorf(cx), iny = (int)floorf(cy);
      if (inx < -LKW || inx >= w || iny < -LKW || iny >= h) {
        if (L == 0) st = 0;
        break;
      }
      if (inx < jx0 || inx - jx0 > kJReg - LKW - 1 || iny < jy0 || iny - jy0 > kJReg - LKW - 1) {
        jx0 = inx - kJMargin;
        jy0 = iny - kJMargin;
        __syncwarp();
#pragma unroll 1
        for (int ch = 0; ch < CN; ++ch) lk_stage_j(sm.jq[ch], J + (long long)ch * g.frame_stride, jx0, jy0, w, h, pitch, lane);
        __syncwarp();
      }
      int v00, v01, v10, v11;
      lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), v00, v01, v10, v11);
      const uint32_t Vt = pack_w(v00, v01), Vb = pack_w(v10, v11);
      const int joff = (iny - jy0) * kJReg + (inx - jx0);
      int tb1 = 0, tb2 = 0;   // <= CN * 14 * 8160 * 4080 < 2^31
#pragma unroll 1
      for (int ch = 0; ch < CN; ++ch) {
        const uint32_t* jb = sm.jq[ch] + joff;
        int sb1 = 0, sb2 = 0;
#pragma unroll
        for (int j = 0; j < kSlots; ++j) {
          const int k = lane + 32 * j;
          if (k < kWin) {
            const uint32_t q = jb[qoff[j]];
            const uint2 tv = sm.t[ch][k];
            const uint32_t d = tv.x;
            const int diff = (dp2a_hi_su(Vb, q, dp2a_lo_su(Vt, q, 1 << 8)) >> 9) - (int)tv.y;
            sb1 += diff * (int)(short)(d & 0xffffu);
            sb2 += diff * ((int)d >> 16);
          }
        }
        tb1 += sb1;
        tb2 += sb2;
      }
      const float b1 = __fmul_rn((float)warp_sum_wide(tb1), flt_scale);
      const float b2 = __fmul_rn((float)warp_sum_wide(tb2), flt_scale);
      const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
      const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
      cx = __fadd_rn(cx, dx);
      cy = __fadd_rn(cy, dy);
      nx = __fadd_rn(cx, 10.f);
      ny = __fadd_rn(cy, 10.f);
      if ((double)dx * (double)dx + (double)dy * (double)dy <= 0.01 * 0.01) break;
      if (it > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
        nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f));
        ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
        break;
      }
      pdx = dx;
      pdy = dy;
    }
    if (L == 0 && st) {
      const float fx = __fsub_rn(nx, 10.f), fy = __fsub_rn(ny, 10.f);
      const int inx = (int)floorf(fx), iny = (int)floorf(fy);
      if (inx < -LKW || inx >= w || iny < -LKW || iny >= h) {
        st = 0;
      } else {
        if (inx < jx0 || inx - jx0 > kJReg - LKW - 1 || iny < jy0 || iny - jy0 > kJReg - LKW - 1) {
          jx0 = inx - kJMargin;
          jy0 = iny - kJMargin;
          __syncwarp();
#pragma unroll 1
          for (int ch = 0; ch < CN; ++ch) lk_stage_j(sm.jq[ch], J + (long long)ch * g.frame_stride, jx0, jy0, w, h, pitch, lane);
          __syncwarp();
        }
        int v00, v01, v10, v11;
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), v00, v01, v10, v11);
        const uint32_t Vt = pack_w(v00, v01), Vb = pack_w(v10, v11);
        const int joff = (iny - jy0) * kJReg + (inx - jx0);
        int se = 0;
#pragma unroll 1
        for (int ch = 0; ch < CN; ++ch) {
          const uint32_t* jb = sm.jq[ch] + joff;
#pragma unroll
          for (int j = 0; j < kSlots; ++j) {
            const int k = lane + 32 * j;
            if (k < kWin) {
              const uint32_t q = jb[qoff[j]];
              se += abs((dp2a_hi_su(Vb, q, dp2a_lo_su(Vt, q, 1 << 8)) >> 9) - (int)sm.t[ch][k].y);
            }
          }
        }
        se = __reduce_add_sync(0xffffffffu, se);
        e = __fdiv_rn((float)se, (float)(32 * LKW * CN * LKW));
      }
    }
  }
  if (lane == 0) {
    const long long o = (long long)b * max_pts + i;
    out_pts[o] = make_float2(nx, ny);
    status[o] = (uint8_t)st;
    err[o] = e;
  }
}

// interleaved BGR rows -> three planes (level 0 of planes b*3 + ch)
// colour[b] becomes nonzero when some pixel of stream b has differing channels (lk_track2_kernel's MUL = 3 path)
__global__ void lk_split3_kernel(const uint8_t* __restrict__ src, int stride, long long in_frame_stride,
                                 uint8_t* __restrict__ dst, int pitch, long long frame_stride, int w, int h,
                                 int32_t* __restrict__ colour) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const uint8_t* p = src + (long long)b * in_frame_stride + (long long)y * stride + 3 * x;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) dst[((long long)b * 3 + ch) * frame_stride + (long long)y * pitch + x] = p[ch];
  if (((p[0] ^ p[1]) | (p[1] ^ p[2])) && __ldcg(colour + b) == 0) atomicOr(colour + b, 1);
}

// ================================================================================================
static void lk_geometry(int w, int h, LkGeom& g) {
  long long off = 0;
  int lw = w, lh = h;
  g.nlevels = 0;
  for (int l = 0; l < kLkLevels; ++l) {
    LkLevel& lv = g.lv[l];
    lv.w = lw;
    lv.h = lh;
    lv.pitch = (int)align_up((size_t)lw, 128);
    lv.off = off;
    off += (long long)lv.pitch * align_up((size_t)lh, 8);
    g.nlevels = l + 1;
    const int nw = (lw + 1) / 2, nh = (lh + 1) / 2;
    if (nw <= LKW || nh <= LKW) break;
    lw = nw;
    lh = nh;
  }
  g.frame_stride = (long long)align_up((size_t)off + 256, 256);
}

// tensor maps of the pyramid levels (driver entry point through the runtime: the library does not link libcuda)
static int lk_make_tmaps(mvo_ctx* c, const LkGeom& g, int planes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MVO_CUDA_TRY(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      c->set_error("cuTensorMapEncodeTiled is not available in this driver");
      return MVO_ERR_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  for (int k = 0; k < 2; ++k)
    for (int kind = 0; kind < 2; ++kind)     // 0: template tiles (24 rows), 1: next-image tiles (29 rows)
      for (int l = 0; l < g.nlevels; ++l) {
        const LkLevel& lv = g.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)lv.pitch, (cuuint64_t)align_up((size_t)lv.h, 8), (cuuint64_t)planes};
        const cuuint64_t strides[2] = {(cuuint64_t)lv.pitch, (cuuint64_t)g.frame_stride};
        const cuuint32_t box[3] = {(cuuint32_t)kTileW, (cuuint32_t)(kind ? kTileHJ : kTileHI), 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode(reinterpret_cast<CUtensorMap*>(c->lk_tmaps[k][kind][l]), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                                  c->lk_pyr[k].p + lv.off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          c->set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
          return MVO_ERR_CUDA;
        }
      }
  c->lk_tmaps_ok = true;
  return MVO_OK;
}

int lk_prepare(mvo_ctx* c, int w, int h, int max_pts, int cn) {
  LkGeom g;
  lk_geometry(w, h, g);
  const size_t B = (size_t)c->cfg.batch;
  const uint8_t* old_base[2] = {c->lk_pyr[0].p, c->lk_pyr[1].p};
  for (int k = 0; k < 2; ++k) MVO_CUDA_TRY(c, c->lk_pyr[k].alloc(B * cn * g.frame_stride));
  if (!c->lk_tmaps_ok || c->lk_w != w || c->lk_h != h || c->lk_cn != cn || old_base[0] != c->lk_pyr[0].p ||
      old_base[1] != c->lk_pyr[1].p) {
    c->lk_tmaps_ok = false;
    const int rc = lk_make_tmaps(c, g, (int)(B * cn));
    if (rc) return rc;
  }
  MVO_CUDA_TRY(c, c->lk_pts_in.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_pts_out.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_status.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_err.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_npts.alloc(B));
  for (int k = 0; k < 2; ++k) MVO_CUDA_TRY(c, c->lk_colour[k].alloc(B));
  MVO_CUDA_TRY(c, c->lk_work.alloc(2));
  c->lk_w = w;
  c->lk_h = h;
  c->lk_max_pts = max_pts;
  c->lk_cn = cn;
  return MVO_OK;
}

uint8_t* lk_level0(mvo_ctx* c, int which, int* pitch, long long* frame_stride) {
  LkGeom g;
  lk_geometry(c->lk_w, c->lk_h, g);
  *pitch = g.lv[0].pitch;
  *frame_stride = g.frame_stride;
  return c->lk_pyr[which].p + g.lv[0].off;
}

// copy level 0 (batch frames, h x stride each) into LK pyramid buffer `which`, then build levels 1..
// on_device: 0 host images, 1 device images, 2 the ORB pyramid's level 0 (gray only), 3 level 0 is already in place
// (the group step's unpack kernel writes it together with the ORB level 0)
int lk_build_pyramid(mvo_ctx* c, int which, const uint8_t* img, int stride, int on_device) {
  LkGeom g;
  lk_geometry(c->lk_w, c->lk_h, g);
  const int B = c->cfg.batch, cn = c->lk_cn;
  uint8_t* base = c->lk_pyr[which].p;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (on_device == 3) {
    // level 0 is in place already (all planes)
  } else if (cn == 3) {
    // interleaved BGR -> staging -> three planes per stream
    const size_t fbytes = (size_t)c->lk_h * stride;
    const uint8_t* src = img;
    if (!on_device) {
      const size_t nbytes = fbytes * (B - 1) + (size_t)(c->lk_h - 1) * stride + (size_t)c->lk_w * 3;   // not past the last row
      MVO_CUDA_TRY(c, c->img_in.alloc(fbytes * B));
      MVO_CUDA_TRY(c, cudaMemcpyAsync(c->img_in.p, img, nbytes, cudaMemcpyHostToDevice, c->stream));
      src = c->img_in.p;
    }
    dim3 grid((c->lk_w + 255) / 256, c->lk_h, B);
    MVO_CUDA_TRY(c, cudaMemsetAsync(c->lk_colour[which].p, 0, sizeof(int32_t) * B, c->stream));
    lk_split3_kernel<<<grid, 256, 0, c->stream>>>(src, stride, (long long)fbytes, base, g.lv[0].pitch, g.frame_stride,
                                                  c->lk_w, c->lk_h, c->lk_colour[which].p);
    c->launches++;
  } else if (on_device == 2) {
    // source is a batch of pitched device images with frame stride given by `stride` == pitch and
    // frame distance c->geom.frame_stride (ORB pyramid level 0)
    for (int b = 0; b < B; ++b)
      MVO_CUDA_TRY(c, cudaMemcpy2DAsync(base + (size_t)b * g.frame_stride, g.lv[0].pitch,
                                        img + (size_t)b * c->geom.frame_stride, stride, c->lk_w, c->lk_h,
                                        cudaMemcpyDeviceToDevice, c->stream));
  } else if (!on_device) {
    const int rc = upload_gray_rows(c, img, c->lk_w, c->lk_h, stride, B, base, g.lv[0].pitch, g.frame_stride);
    if (rc) return rc;
  } else {
    for (int b = 0; b < B; ++b)
      MVO_CUDA_TRY(c, cudaMemcpy2DAsync(base + (size_t)b * g.frame_stride, g.lv[0].pitch,
                                        img + (size_t)b * c->lk_h * stride, stride, c->lk_w, c->lk_h, kind,
                                        c->stream));
  }
  for (int l = 1; l < g.nlevels; ++l) {
    const LkLevel& s = g.lv[l - 1];
    const LkLevel& d = g.lv[l];
    if (s.w >= 8 && s.h >= 4) {
      dim3 grid((d.w + PD2W - 1) / PD2W, (d.h + PD2H - 1) / PD2H, B * cn);
      lk_pyrdown_tile_kernel<<<grid, 256, 0, c->stream>>>(base + s.off, s.w, s.h, s.pitch, base + d.off, d.w, d.h, d.pitch,
                                                          g.frame_stride, g.frame_stride);
    } else {
      dim3 grid((d.w + PDW - 1) / PDW, (d.h + PDH - 1) / PDH, B * cn), block(PDW, PDH);
      lk_pyrdown_kernel<<<grid, block, 0, c->stream>>>(base + s.off, s.w, s.h, s.pitch, base + d.off, d.w, d.h, d.pitch,
                                                       g.frame_stride, g.frame_stride);
    }
    c->launches++;
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

// track points (device arrays) from pyramid `prev_which` to pyramid `next_which`
int lk_run(mvo_ctx* c, int prev_which, int next_which, const float2* pts_dev, const int32_t* npts_dev, int max_pts,
           float2* out_dev, uint8_t* status_dev, float* err_dev) {
  LkGeom g;
  lk_geometry(c->lk_w, c->lk_h, g);
  if (max_pts > 0) {
    if (c->lk_cn == 3 && (c->dbg_lk_impl == 1 || c->cfg.batch > kLk2MaxBatch)) {
      dim3 grid((max_pts + kLkWarpsCn - 1) / kLkWarpsCn, c->cfg.batch);
      lk_track_cn_kernel<3><<<grid, kLkWarpsCn * 32, 0, c->stream>>>(g, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p,
                                                                    pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev);
    } else if (c->lk_cn == 3) {
      // BGR8: teams of three warps (one per colour plane) on the second-generation kernel, persistent form, for the
      // streams with real colour; the streams whose planes are identical in both frames (a gray camera behind the
      // node's BGR8 conversion) take the gray path with tripled sums (mvo_debug_set("lk_bgr_gray", 0): all to the teams)
      LkTmaps tm;
      memcpy(tm.m[0], c->lk_tmaps[prev_which][0], sizeof(tm.m[0]));
      memcpy(tm.m[1], c->lk_tmaps[next_which][1], sizeof(tm.m[1]));
      const long long items = (long long)max_pts * c->cfg.batch;
      const bool split = c->dbg_lk_bgr_gray != 0;
      const int32_t* ca = split ? c->lk_colour[prev_which].p : nullptr;
      const int32_t* cb = split ? c->lk_colour[next_which].p : nullptr;
      MVO_CUDA_TRY(c, cudaMemsetAsync(c->lk_work.p, 0, 8, c->stream));
      const int ctas = (int)std::min<long long>(148LL * MVO_LK_T_MINB, std::max<long long>(items, 1));
      lk_track2_kernel<true, 3><<<dim3(ctas, 1), kLk2WarpsCn * 32, 0, c->stream>>>(
          g, tm, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev,
          c->lk_work.p, c->cfg.batch, ca, cb, 1);
      if (split) {
        if (items < 148LL * MVO_LK2_MINB * kLk2Warps * 4)
          lk_track2_kernel<false, 1, 3><<<dim3((max_pts + kLk2Warps - 1) / kLk2Warps, c->cfg.batch), kLk2Warps * 32, 0, c->stream>>>(
              g, tm, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev,
              nullptr, c->cfg.batch, ca, cb, 0);
        else
          lk_track2_kernel<true, 1, 3><<<dim3(148 * MVO_LK2_MINB, 1), kLk2Warps * 32, 0, c->stream>>>(
              g, tm, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev,
              c->lk_work.p + 1, c->cfg.batch, ca, cb, 0);
        c->launches++;
      }
    } else {
      dim3 grid((max_pts + kLkWarps - 1) / kLkWarps, c->cfg.batch);
      // mvo_debug_set("lk_impl", 1): the first-generation kernel (kept as the in-tree cross-check; identical results)
      if (c->dbg_lk_impl == 1)
        lk_track_kernel<<<grid, kLkWarps * 32, 0, c->stream>>>(g, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev,
                                                              npts_dev, max_pts, out_dev, status_dev, err_dev);
      else {
        LkTmaps tm;
        memcpy(tm.m[0], c->lk_tmaps[prev_which][0], sizeof(tm.m[0]));
        memcpy(tm.m[1], c->lk_tmaps[next_which][1], sizeof(tm.m[1]));
        const long long items = (long long)max_pts * c->cfg.batch;
        if (c->dbg_lk_impl == 3 || items < 148LL * MVO_LK2_MINB * kLk2Warps * 4 || c->cfg.batch > kLk2MaxBatch) {
          // few points (single stream): one point per warp fills the GPU better than a persistent grid
          lk_track2_kernel<false, 1><<<dim3((max_pts + kLk2Warps - 1) / kLk2Warps, c->cfg.batch), kLk2Warps * 32, 0, c->stream>>>(
              g, tm, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev,
              nullptr, c->cfg.batch, nullptr, nullptr, 0);
        } else {
          MVO_CUDA_TRY(c, cudaMemsetAsync(c->lk_work.p, 0, 4, c->stream));
          const int per_sm = std::min(std::max(c->dbg_lk_ctas_per_sm, 1), MVO_LK2_MINB);
          lk_track2_kernel<true, 1><<<dim3(148 * per_sm, 1), kLk2Warps * 32, 0, c->stream>>>(
              g, tm, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev,
              c->lk_work.p, c->cfg.batch, nullptr, nullptr, 0);
        }
      }
    }
    c->launches++;
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

}  // namespace mvo

using namespace mvo;

// parity tap: pyramid level `level` of the previous (which = 0) / next (which = 1) image of the last mvo_lk_track call
// (channel plane `plane` for BGR input)
extern "C" int mvo_lk_get_level(mvo_ctx* c, int which, int level, int plane, uint8_t* out, int out_stride, int* w, int* h) {
  if (!c || which < 0 || which > 1 || level < 0 || plane < 0 || c->lk_w <= 0 || plane >= c->lk_cn) return MVO_ERR_INVALID;
  LkGeom g;
  lk_geometry(c->lk_w, c->lk_h, g);
  if (level >= g.nlevels) return MVO_ERR_INVALID;
  const LkLevel& lv = g.lv[level];
  if (w) *w = lv.w;
  if (h) *h = lv.h;
  if (!out) return MVO_OK;
  if (out_stride < lv.w) return MVO_ERR_INVALID;
  const int slot = which == 0 ? c->lk_slot_prev : c->lk_slot_next;   // the pyramid cache may have swapped the slots
  MVO_CUDA_TRY(c, cudaMemcpy2DAsync(out, out_stride, c->lk_pyr[slot].p + (size_t)plane * g.frame_stride + lv.off, lv.pitch, lv.w,
                                    lv.h, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}

extern "C" int mvo_lk_track(mvo_ctx* c, const uint8_t* prev, const uint8_t* next, int w, int h, int stride,
                            int channels, const float* prev_xy, int n, float* next_xy, uint8_t* status, float* err) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  if (!prev || !next || n < 0 || (n > 0 && (!prev_xy || !next_xy || !status || !err))) {
    c->set_error("mvo_lk_track: null argument");
    return MVO_ERR_INVALID;
  }
  if (channels != 1 && channels != 3) {
    c->set_error("mvo_lk_track: images must have 1 (gray) or 3 (BGR) channels");
    return MVO_ERR_UNSUPPORTED;
  }
  if (w <= LKW || h <= LKW || w > c->cfg.max_width || h > c->cfg.max_height || stride < w * channels) {
    c->set_error("mvo_lk_track: image size out of range");
    return MVO_ERR_INVALID;
  }
  if (c->cfg.batch != 1) {
    c->set_error("mvo_lk_track needs a batch==1 context");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if (n == 0) return MVO_OK;
  // the ping-pong pyramids are (re)built below: the group entry points must not take them for their previous frame
  c->have_prev = false;
  c->trk_have_frame = false;
  if (c->lk_w != w || c->lk_h != h || c->lk_cn != channels || c->lk_max_pts < n) c->lk_hash[0] = c->lk_hash[1] = 0;
  int rc = lk_prepare(c, w, h, std::max(n, c->lk_max_pts), channels);
  if (rc) return rc;
  // Pyramid cache: the tracker's "prev" image is the "next" image of its previous call
  // (/root/reference/src/tracker.cpp:68-69, :331) in a fresh buffer, so a pyramid slot is recognised by the content
  // hash of the image it was built from: a resident image is neither uploaded nor reduced again.
  int sp = 0, sn = 1;
  uint64_t hp = 0, hn = 0;
  if (c->cache_enabled) {
    hp = content_hash_rows(prev, h, (size_t)w * channels, (size_t)stride);
    hn = content_hash_rows(next, h, (size_t)w * channels, (size_t)stride);
    if (hp == c->lk_hash[1]) sp = 1, sn = 0;
    else if (hp != c->lk_hash[0] && hn == c->lk_hash[0]) sp = 1, sn = 0;
  }
  const bool have_p = c->cache_enabled && c->lk_hash[sp] == hp, have_n = c->cache_enabled && c->lk_hash[sn] == hn && hn != hp;
  c->cache_stats[2] += (have_p ? 1 : 0) + (have_n ? 1 : 0);
  c->cache_stats[3] += (have_p ? 0 : 1) + (have_n ? 0 : 1);
  if (!have_p) {
    c->lk_hash[sp] = 0;
    rc = lk_build_pyramid(c, sp, prev, stride, 0);
    if (rc) return rc;
    c->lk_hash[sp] = hp;
  }
  if (!have_n) {
    c->lk_hash[sn] = 0;
    rc = lk_build_pyramid(c, sn, next, stride, 0);
    if (rc) return rc;
    c->lk_hash[sn] = hn;
  }
  c->lk_slot_prev = sp;
  c->lk_slot_next = sn;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->lk_pts_in.p, prev_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->lk_npts.p, &n, 4, cudaMemcpyHostToDevice, c->stream));
  rc = lk_run(c, sp, sn, c->lk_pts_in.p, c->lk_npts.p, c->lk_max_pts, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p);
  if (rc) return rc;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(next_xy, c->lk_pts_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(status, c->lk_status.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(err, c->lk_err.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}
