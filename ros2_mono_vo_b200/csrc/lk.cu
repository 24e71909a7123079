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
#include <algorithm>

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
  if (ix >= 1 && ix - 1 + 32 <= w && iy >= 1 && iy - 1 + kRaw <= h) {
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
__global__ void lk_split3_kernel(const uint8_t* __restrict__ src, int stride, long long in_frame_stride,
                                 uint8_t* __restrict__ dst, int pitch, long long frame_stride, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const uint8_t* p = src + (long long)b * in_frame_stride + (long long)y * stride + 3 * x;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) dst[((long long)b * 3 + ch) * frame_stride + (long long)y * pitch + x] = p[ch];
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

int lk_prepare(mvo_ctx* c, int w, int h, int max_pts, int cn) {
  LkGeom g;
  lk_geometry(w, h, g);
  const size_t B = (size_t)c->cfg.batch;
  for (int k = 0; k < 2; ++k) MVO_CUDA_TRY(c, c->lk_pyr[k].alloc(B * cn * g.frame_stride));
  MVO_CUDA_TRY(c, c->lk_pts_in.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_pts_out.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_status.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_err.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_npts.alloc(B));
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
  if (cn == 3) {
    // interleaved BGR -> staging -> three planes per stream
    const size_t fbytes = (size_t)c->lk_h * stride;
    const uint8_t* src = img;
    if (!on_device) {
      MVO_CUDA_TRY(c, c->img_in.alloc(fbytes * B));
      MVO_CUDA_TRY(c, cudaMemcpyAsync(c->img_in.p, img, fbytes * B, cudaMemcpyHostToDevice, c->stream));
      src = c->img_in.p;
    }
    dim3 grid((c->lk_w + 255) / 256, c->lk_h, B);
    lk_split3_kernel<<<grid, 256, 0, c->stream>>>(src, stride, (long long)fbytes, base, g.lv[0].pitch, g.frame_stride,
                                                  c->lk_w, c->lk_h);
    c->launches++;
  } else if (on_device == 3) {
    // nothing to copy
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
    if (c->lk_cn == 3) {
      dim3 grid((max_pts + kLkWarpsCn - 1) / kLkWarpsCn, c->cfg.batch);
      lk_track_cn_kernel<3><<<grid, kLkWarpsCn * 32, 0, c->stream>>>(g, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p,
                                                                    pts_dev, npts_dev, max_pts, out_dev, status_dev, err_dev);
    } else {
      dim3 grid((max_pts + kLkWarps - 1) / kLkWarps, c->cfg.batch);
      lk_track_kernel<<<grid, kLkWarps * 32, 0, c->stream>>>(g, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev,
                                                            npts_dev, max_pts, out_dev, status_dev, err_dev);
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
  MVO_CUDA_TRY(c, cudaMemcpy2DAsync(out, out_stride, c->lk_pyr[which].p + (size_t)plane * g.frame_stride + lv.off, lv.pitch, lv.w,
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
  int rc = lk_prepare(c, w, h, std::max(n, c->lk_max_pts), channels);
  if (rc) return rc;
  rc = lk_build_pyramid(c, 0, prev, stride, 0);
  if (rc) return rc;
  rc = lk_build_pyramid(c, 1, next, stride, 0);
  if (rc) return rc;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->lk_pts_in.p, prev_xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->lk_npts.p, &n, 4, cudaMemcpyHostToDevice, c->stream));
  rc = lk_run(c, 0, 1, c->lk_pts_in.p, c->lk_npts.p, c->lk_max_pts, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p);
  if (rc) return rc;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(next_xy, c->lk_pts_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(status, c->lk_status.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(err, c->lk_err.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return MVO_OK;
}
