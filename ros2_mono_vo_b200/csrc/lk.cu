// lk.cu -- pyramidal Lucas-Kanade tracking (cv::calcOpticalFlowPyrLK, all defaults) for B200.
//
// Replaces the call at /root/reference/src/tracker.cpp:68-69.  Contract: SURVEY.md A.3 / oracle/lk_oracle.py
// (status identical, positions within 0.01 px of OpenCV; here the integer sums are exact, which is tighter
// than OpenCV's own float accumulation).
//
//   lk_pyrdown_kernel  integer 5x5 [1 4 6 4 1]^2 pyrDown, smem tile with REFLECT_101 halo.
//   lk_track_kernel    one warp per point, all pyramid levels inside the kernel (no inter-point dependency):
//                      24x24 raw patch in smem -> Scharr derivatives computed on the fly (no derivative
//                      image is ever written: saves 5.3 P0 bytes of dense traffic per frame) -> Q14 bilinear
//                      patch / gradient values in registers -> 2x2 normal matrix by warp reduction ->
//                      <= 30 Gauss-Newton iterations sampling J from a 32x32 smem region (texture-free
//                      bilinear), re-staged only when the window leaves it.
#include "context.cuh"
#include <algorithm>

namespace mvo {

constexpr int LKW = 21;
constexpr int kLkWarps = 8;
constexpr int kRaw = 24;      // raw patch side: 22 interp rows + 1 Scharr ring
constexpr int kDer = 22;      // derivative patch side
constexpr int kJReg = 32;     // staged J region side
constexpr int kJMargin = 5;

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

// ---- tracking -------------------------------------------------------------------------------
struct LkWarpSmem {
  uint8_t raw[kRaw * kRaw];
  short2 der[kDer * kDer];
  uint8_t jreg[kJReg * kJReg];
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

__global__ void __launch_bounds__(kLkWarps * 32)
lk_track_kernel(const LkGeom g, const uint8_t* __restrict__ pyrI, const uint8_t* __restrict__ pyrJ,
                const float2* __restrict__ pts, const int32_t* __restrict__ npts_dev, int max_pts,
                float2* __restrict__ out_pts, uint8_t* __restrict__ status, float* __restrict__ err) {
  __shared__ LkWarpSmem sm_all[kLkWarps];
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
    if (ix < -LKW || ix >= w || iy < -LKW || iy >= h) {
      if (L == 0) {
        st = 0;
        e = 0.f;
      }
      continue;
    }
    int w00, w01, w10, w11;
    lk_weights(__fsub_rn(qx, (float)ix), __fsub_rn(qy, (float)iy), w00, w01, w10, w11);

    __syncwarp();
    for (int k = lane; k < kRaw * kRaw; k += 32) {
      const int r = k / kRaw, c = k - r * kRaw;
      sm.raw[k] = I[(long long)safe_reflect(iy - 1 + r, h) * pitch + safe_reflect(ix - 1 + c, w)];
    }
    __syncwarp();
    for (int k = lane; k < kDer * kDer; k += 32) {
      const int r = k / kDer, c = k - r * kDer;
      const int y = iy + r, x = ix + c;
      short2 d = make_short2(0, 0);
      if (x >= 0 && x < w && y >= 0 && y < h) {
        const uint8_t* q = sm.raw + (r + 1) * kRaw + c + 1;
        const int a0 = q[-kRaw - 1], a1 = q[-kRaw], a2 = q[-kRaw + 1];
        const int b0 = q[-1], b1 = q[0], b2 = q[1];
        const int c0 = q[kRaw - 1], c1 = q[kRaw], c2 = q[kRaw + 1];
        (void)b1;
        const int t0l = (a0 + c0) * 3 + b0 * 10, t0r = (a2 + c2) * 3 + b2 * 10;
        const int t1l = c0 - a0, t1m = c1 - a1, t1r = c2 - a2;
        d.x = (short)(t0r - t0l);
        d.y = (short)((t1r + t1l) * 3 + t1m * 10);
      }
      sm.der[k] = d;
    }
    __syncwarp();
    // window values owned by this lane (pixel k = lane + 32 j)
    int Iv[14], Ixv[14], Iyv[14];
    int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
    for (int j = 0; j < 14; ++j) {
      const int k = lane + 32 * j;
      Iv[j] = Ixv[j] = Iyv[j] = 0;
      if (k < LKW * LKW) {
        const int r = k / LKW, c = k - r * LKW;
        const uint8_t* q = sm.raw + (r + 1) * kRaw + c + 1;
        Iv[j] = (q[0] * w00 + q[1] * w01 + q[kRaw] * w10 + q[kRaw + 1] * w11 + (1 << 8)) >> 9;
        const short2 d00 = sm.der[r * kDer + c], d01 = sm.der[r * kDer + c + 1];
        const short2 d10 = sm.der[(r + 1) * kDer + c], d11 = sm.der[(r + 1) * kDer + c + 1];
        const int vx = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << 13)) >> 14;
        const int vy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << 13)) >> 14;
        Ixv[j] = vx;
        Iyv[j] = vy;
        sA11 += vx * vx;
        sA12 += vx * vy;
        sA22 += vy * vy;
      }
    }
    const float A11 = __fmul_rn((float)warp_sum_ll(sA11), flt_scale);
    const float A12 = __fmul_rn((float)warp_sum_ll(sA12), flt_scale);
    const float A22 = __fmul_rn((float)warp_sum_ll(sA22), flt_scale);
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
        for (int r = 0; r < kJReg; ++r)
          sm.jreg[r * kJReg + lane] = J[(long long)safe_reflect(jy0 + r, h) * pitch + safe_reflect(jx0 + lane, w)];
        __syncwarp();
      }
      int v00, v01, v10, v11;
      lk_weights(__fsub_rn(cx, (float)inx), __fsub_rn(cy, (float)iny), v00, v01, v10, v11);
      const uint8_t* jb = sm.jreg + (iny - jy0) * kJReg + (inx - jx0);
      int sb1 = 0, sb2 = 0;
#pragma unroll
      for (int j = 0; j < 14; ++j) {
        const int k = lane + 32 * j;
        if (k < LKW * LKW) {
          const int r = k / LKW, c = k - r * LKW;
          const uint8_t* q = jb + r * kJReg + c;
          const int diff = ((q[0] * v00 + q[1] * v01 + q[kJReg] * v10 + q[kJReg + 1] * v11 + (1 << 8)) >> 9) - Iv[j];
          sb1 += diff * Ixv[j];
          sb2 += diff * Iyv[j];
        }
      }
      const float b1 = __fmul_rn((float)warp_sum_ll(sb1), flt_scale);
      const float b2 = __fmul_rn((float)warp_sum_ll(sb2), flt_scale);
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
          for (int r = 0; r < kJReg; ++r)
            sm.jreg[r * kJReg + lane] = J[(long long)safe_reflect(jy0 + r, h) * pitch + safe_reflect(jx0 + lane, w)];
          __syncwarp();
        }
        int v00, v01, v10, v11;
        lk_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), v00, v01, v10, v11);
        const uint8_t* jb = sm.jreg + (iny - jy0) * kJReg + (inx - jx0);
        int se = 0;
#pragma unroll
        for (int j = 0; j < 14; ++j) {
          const int k = lane + 32 * j;
          if (k < LKW * LKW) {
            const int r = k / LKW, c = k - r * LKW;
            const uint8_t* q = jb + r * kJReg + c;
            se += abs(((q[0] * v00 + q[1] * v01 + q[kJReg] * v10 + q[kJReg + 1] * v11 + (1 << 8)) >> 9) - Iv[j]);
          }
        }
        se = warp_sum(se);
        e = __fdiv_rn((float)se, (float)(32 * LKW * LKW));
      }
    }
  }
  if (lane == 0) {
    const long long o = (long long)b * max_pts + i;
    out_pts[o] = make_float2(nx, ny);
    status[o] = (uint8_t)st;
    err[o] = st ? e : e;
  }
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

int lk_prepare(mvo_ctx* c, int w, int h, int max_pts) {
  LkGeom g;
  lk_geometry(w, h, g);
  const size_t B = (size_t)c->cfg.batch;
  for (int k = 0; k < 2; ++k) MVO_CUDA_TRY(c, c->lk_pyr[k].alloc(B * g.frame_stride));
  MVO_CUDA_TRY(c, c->lk_pts_in.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_pts_out.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_status.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_err.alloc(B * (size_t)max_pts));
  MVO_CUDA_TRY(c, c->lk_npts.alloc(B));
  c->lk_w = w;
  c->lk_h = h;
  c->lk_max_pts = max_pts;
  return MVO_OK;
}

// copy level 0 (batch frames, h x stride each) into LK pyramid buffer `which`, then build levels 1..
int lk_build_pyramid(mvo_ctx* c, int which, const uint8_t* img, int stride, int on_device) {
  LkGeom g;
  lk_geometry(c->lk_w, c->lk_h, g);
  const int B = c->cfg.batch;
  uint8_t* base = c->lk_pyr[which].p;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (on_device == 2) {
    // source is a batch of pitched device images with frame stride given by `stride` == pitch and
    // frame distance c->geom.frame_stride (ORB pyramid level 0)
    for (int b = 0; b < B; ++b)
      MVO_CUDA_TRY(c, cudaMemcpy2DAsync(base + (size_t)b * g.frame_stride, g.lv[0].pitch,
                                        img + (size_t)b * c->geom.frame_stride, stride, c->lk_w, c->lk_h,
                                        cudaMemcpyDeviceToDevice, c->stream));
  } else {
    for (int b = 0; b < B; ++b)
      MVO_CUDA_TRY(c, cudaMemcpy2DAsync(base + (size_t)b * g.frame_stride, g.lv[0].pitch,
                                        img + (size_t)b * c->lk_h * stride, stride, c->lk_w, c->lk_h, kind,
                                        c->stream));
  }
  for (int l = 1; l < g.nlevels; ++l) {
    const LkLevel& s = g.lv[l - 1];
    const LkLevel& d = g.lv[l];
    dim3 grid((d.w + PDW - 1) / PDW, (d.h + PDH - 1) / PDH, B), block(PDW, PDH);
    lk_pyrdown_kernel<<<grid, block, 0, c->stream>>>(base + s.off, s.w, s.h, s.pitch, base + d.off, d.w, d.h, d.pitch,
                                                     g.frame_stride, g.frame_stride);
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
  dim3 grid((max_pts + kLkWarps - 1) / kLkWarps, c->cfg.batch);
  if (max_pts > 0) {
    lk_track_kernel<<<grid, kLkWarps * 32, 0, c->stream>>>(g, c->lk_pyr[prev_which].p, c->lk_pyr[next_which].p, pts_dev,
                                                          npts_dev, max_pts, out_dev, status_dev, err_dev);
    c->launches++;
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  return MVO_OK;
}

}  // namespace mvo

using namespace mvo;

extern "C" int mvo_lk_track(mvo_ctx* c, const uint8_t* prev, const uint8_t* next, int w, int h, int stride,
                            int channels, const float* prev_xy, int n, float* next_xy, uint8_t* status, float* err) {
  if (!c) return MVO_ERR_INVALID;
  if (!prev || !next || n < 0 || (n > 0 && (!prev_xy || !next_xy || !status || !err))) {
    c->set_error("mvo_lk_track: null argument");
    return MVO_ERR_INVALID;
  }
  if (channels != 1) {
    c->set_error("mvo_lk_track: only single-channel images are implemented (cn=3 windows are a listed next step)");
    return MVO_ERR_UNSUPPORTED;
  }
  if (w <= LKW || h <= LKW || w > c->cfg.max_width || h > c->cfg.max_height || stride < w) {
    c->set_error("mvo_lk_track: image size out of range");
    return MVO_ERR_INVALID;
  }
  if (c->cfg.batch != 1) {
    c->set_error("mvo_lk_track needs a batch==1 context");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if (n == 0) return MVO_OK;
  int rc = lk_prepare(c, w, h, std::max(n, c->lk_max_pts));
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
