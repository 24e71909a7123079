// common.cuh -- context, geometry and small device helpers shared by all kernels of libmonovo_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/monovo_b200.h"

namespace mvo {

constexpr int kLevels = 8;        // cv::ORB default nlevels
constexpr int kEdge = 31;         // edgeThreshold
constexpr int kFastThr = 20;      // fastThreshold
constexpr int kLkLevels = 4;      // calcOpticalFlowPyrLK default maxLevel 3 -> levels 0..3
constexpr int kMaxHypModels = 10; // 5-point solver returns <= 10 models

struct LevelGeom {
  int w, h, pitch;        // pitch in bytes, multiple of 128
  int quota;              // n_l (features kept on this level)
  int cand_cap;           // capacity of the FAST candidate list of this level
  int cand_off;           // offset of this level inside the per-frame candidate arrays
  int xtab_off, ytab_off; // offsets into the resize coefficient tables
  long long off;          // byte offset of the level inside one frame's pyramid buffer
  float scale;            // (float)pow((double)1.2f, l)
  float inv_scale;        // 1.f / scale
};

struct OrbGeom {
  LevelGeom lv[kLevels];
  long long frame_stride;   // bytes between consecutive frames' pyramids
  int cand_total;           // sum of cand_cap over levels (per frame)
  int kp_cap;               // final keypoint capacity per frame
  int nfeatures;
  int batch;
  int grid_div, grid_rows, grid_cols;   // keypoint-distribution grid (src/initializer.cpp:57-61); grid_div 0 = off
};

// Candidate after FAST+NMS: packed position and score.
// Candidate after Harris: 64-bit sort key + response + angle.

#define MVO_CUDA_TRY(ctx, expr)                                                   \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));       \
      return MVO_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

// bumped by every real (re)allocation: a captured CUDA graph holds raw pointers and is re-captured when it changes
inline unsigned long long& alloc_epoch() {
  static unsigned long long e = 0;
  return e;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    ++alloc_epoch();
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t n = 0;
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  PinBuf(PinBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  PinBuf& operator=(PinBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~PinBuf() { release(); }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMallocHost((void**)&p, (count ? count : 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    ++alloc_epoch();
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
  // BORDER_REFLECT_101 for |overshoot| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Exact warp sum of 32-bit addends whose total needs more than 32 bits: the low and the high 16 bits are reduced
// separately by the warp-reduction unit (REDUX.SUM, one instruction each instead of ten shuffles + ten adds).
__device__ __forceinline__ long long warp_sum_wide(int v) {
  const int lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
  const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
  return ((long long)hi << 16) + lo;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// float -> uint32 whose unsigned order equals the float order (no NaNs expected)
__device__ __forceinline__ uint32_t float_orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace mvo
