// group.cu -- mvo_group_step: the whole front-end frame for a group of independent camera streams.
//
// One "front-end frame"  per stream: ORB(frame t) -> kNN+ratio(desc t-1, desc t) ->
// LK(keypoints t-1 -> frame t) -> H + F RANSAC -> E RANSAC -> recoverPose -> triangulate.  It chains the same
// kernels the single-call ABI uses (every kernel carries the stream index in its grid), keeps all
// intermediate data on the device and returns one small result record per stream.
// Call sites it stands for: Tracker::update (/root/reference/src/tracker.cpp:274-333) and
// Initializer::try_initializing (src/initializer.cpp:165-313).
#include "context.cuh"
#include <nvtx3/nvToolsExt.h>
#include <string.h>
#include <algorithm>

namespace mvo {

enum Stage { ST_ORB = 0, ST_KNN, ST_LK, ST_H, ST_F, ST_E, ST_POSE, ST_TRI, ST_TOTAL, ST_ORB_DENSE };
static const char* kStageNames[mvo_ctx::kNumStages] = {"orb", "knn", "lk", "ransac_h", "ransac_f", "ransac_e",
                                                       "pose", "triangulate", "total", "orb_dense"};

// keep LK tracks with status && err < err_thr (src/tracker.cpp:70-77), ordered compaction into (p1, p2)
__global__ void __launch_bounds__(1024)
lk_collect_kernel(const float2* __restrict__ prev_xy, const float2* __restrict__ next_xy,
                  const uint8_t* __restrict__ status, const float* __restrict__ err, const int32_t* __restrict__ nprev,
                  int max_pts, float err_thr, float2* __restrict__ p1, float2* __restrict__ p2,
                  int32_t* __restrict__ npts) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(nprev[b], max_pts);
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const long long o = (long long)b * max_pts + i;
    const int ok = (i < n) ? (status[o] != 0 && err[o] < err_thr) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (ok) {
      const long long d = (long long)b * max_pts + off + __popc(bal & ((1u << lane) - 1));
      p1[d] = prev_xy[o];
      p2[d] = next_xy[o];
    }
    __syncthreads();
    if (tid == 0)
      for (int w = 0; w < 32; ++w) s_base += s_warp[w];
    __syncthreads();
  }
  if (tid == 0) npts[b] = s_base;
}

// per-stream squared threshold; optionally replicate the camera matrix of stream 0 to every stream
__global__ void fill_params_kernel(float* thr2, float v, double* K, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  thr2[b] = v;
  if (K && b > 0)
    for (int i = 0; i < 9; ++i) K[b * 9 + i] = K[i];
}

// P0 = K [I | 0], P1 = K [R | t]   (src/initializer.cpp:119-123)
__global__ void make_proj_kernel(const double* __restrict__ K, const double* __restrict__ pose, double* __restrict__ proj,
                                 int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* k = K + b * 9;
  const double* R = pose + b * 12;
  const double* t = R + 9;
  double* P0 = proj + b * 24;
  double* P1 = P0 + 12;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      P0[i * 4 + j] = k[i * 3 + j];
      P1[i * 4 + j] = k[i * 3] * R[j] + k[i * 3 + 1] * R[3 + j] + k[i * 3 + 2] * R[6 + j];
    }
    P0[i * 4 + 3] = 0;
    P1[i * 4 + 3] = k[i * 3] * t[0] + k[i * 3 + 1] * t[1] + k[i * 3 + 2] * t[2];
  }
}

// chirality count of the triangulated points (src/initializer.cpp:134-157): z > 0 in both cameras.  With `cloud` the
// valid points are also compacted in track order and written as points3d_to_pointcloud_msg would pack them
// (src/utils.cpp:225-241: ROS x = z, y = -x, z = -y; 12 bytes per point) -- MVO_OUT_CLOUD.
constexpr int kTriCountThreads = 1024;   // one CTA per stream walks its ~1700 tracks in chunks: two chunks instead of seven
__global__ void __launch_bounds__(kTriCountThreads)
tri_count_kernel(const float* __restrict__ X4, const uint8_t* __restrict__ mask, const double* __restrict__ pose,
                 const int32_t* __restrict__ npts, int max_pts, int32_t* __restrict__ out, float* __restrict__ cloud,
                 int cloud_cap) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = npts[b];
  const double* R = pose + b * 12;
  const double* t = R + 9;
  const float* X = X4 + (long long)b * 4 * max_pts;
  __shared__ int s_warp[kTriCountThreads / 32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += kTriCountThreads) {
    const int i = i0 + tid;
    bool ok = false;
    float x = 0, y = 0, z = 0;
    if (i < n && mask[(long long)b * max_pts + i]) {
      const float w = X[3LL * max_pts + i];
      const float sc = w != 0.f ? 1.f / w : 1.f;        // convertPointsFromHomogeneous
      x = X[i] * sc;
      y = X[1LL * max_pts + i] * sc;
      z = X[2LL * max_pts + i] * sc;
      if (z > 0) {
        const double z2 = R[6] * (double)x + R[7] * (double)y + R[8] * (double)z + t[2];
        ok = z2 > 0;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base, tot = 0;
#pragma unroll
    for (int w2 = 0; w2 < kTriCountThreads / 32; ++w2) {
      if (w2 < warp) off += s_warp[w2];
      tot += s_warp[w2];
    }
    if (ok && cloud) {
      const int k = off + __popc(bal & ((1u << lane) - 1));
      if (k < cloud_cap) {
        float* o = cloud + ((long long)b * cloud_cap + k) * 3;
        o[0] = z;
        o[1] = -x;
        o[2] = -y;
      }
    }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  if (tid == 0) out[b] = s_base;
}

__global__ void gather_results_kernel(const int32_t* kp_count, const int32_t* nmatch, const int32_t* ntracked,
                                      const int32_t* res_h, const int32_t* res_f, const int32_t* res_e,
                                      const int32_t* ntri, const double* pose, int have_prev,
                                      mvo_frame_result* out, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  mvo_frame_result r;
  memset(&r, 0, sizeof(r));
  r.n_keypoints = kp_count[b];
  if (have_prev) {
    r.n_matches = nmatch[b];
    r.n_tracked = ntracked[b];
    r.score_h = res_h[b];
    r.score_f = res_f[b];
    r.n_inliers_e = res_e[b * 8 + 0];
    r.n_pose_good = res_e[b * 8 + 4];
    r.n_triangulated = ntri[b];
    for (int i = 0; i < 9; ++i) r.R[i] = pose[b * 12 + i];
    for (int i = 0; i < 3; ++i) r.t[i] = pose[b * 12 + 9 + i];
  }
  out[b] = r;
}

// staged frames (batch x h rows of `spitch` bytes) -> level 0 of every stream's ORB pyramid and of its LK pyramid, 16 bytes
// per thread (one launch instead of a 2-D device copy per stream)
__global__ void __launch_bounds__(256)
unpack_frames_kernel(const uint8_t* __restrict__ src, int spitch, long long src_frame_stride, uint8_t* __restrict__ dst,
                     int dpitch, long long dst_frame_stride, uint8_t* __restrict__ dst2, int dpitch2,
                     long long dst2_frame_stride, int w, int h) {
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (x >= w) return;
  const uint8_t* s = src + (long long)b * src_frame_stride + (long long)y * spitch + x;
  uint8_t* d = dst + (long long)b * dst_frame_stride + (long long)y * dpitch + x;   // 16-byte aligned (pitch % 128 == 0)
  uint8_t v[16];
  const int nb = min(16, w - x);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = i < nb ? s[i] : 0;
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(v);
  // second copy: level 0 of the LK pyramid (it outlives the ORB pyramid by one step)
  if (dst2) *reinterpret_cast<uint4*>(dst2 + (long long)b * dst2_frame_stride + (long long)y * dpitch2 + x) = *reinterpret_cast<const uint4*>(v);
}

__global__ void copy_i32_strided_kernel(const int32_t* src, int stride, int32_t* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i * stride];
}

// staged BGR8 frames (batch x h rows of `spitch` bytes, 3 bytes per pixel) -> gray level 0 of every stream's ORB pyramid
// ((B*3735 + G*19235 + R*9798 + 16384) >> 15 == cvtColor BGR2GRAY) and the three colour planes of level 0 of
// its LK pyramid (cv::calcOpticalFlowPyrLK tracks on all channels of a 3-channel Mat); 4 pixels per thread, one aligned
// 32-bit store per destination.  gray may be NULL (tracking frame: LK only).
__global__ void __launch_bounds__(256)
unpack_bgr_frames_kernel(const uint8_t* __restrict__ src, int spitch, long long src_frame_stride, uint8_t* __restrict__ gray,
                         int gpitch, long long gray_frame_stride, uint8_t* __restrict__ planes, int ppitch,
                         long long plane_stride, int w, int h, int32_t* __restrict__ colour) {
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x >= w) return;
  const uint8_t* s = src + (long long)b * src_frame_stride + (long long)y * spitch + 3 * x;
  const int nb = min(4, w - x);
  uint32_t g = 0, p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < nb) {
      const uint32_t cb = s[3 * i], cg = s[3 * i + 1], cr = s[3 * i + 2];
      g |= ((cb * 3735u + cg * 19235u + cr * 9798u + 16384u) >> 15) << (8 * i);
      p0 |= cb << (8 * i);
      p1 |= cg << (8 * i);
      p2 |= cr << (8 * i);
    }
  }
  if (gray) *reinterpret_cast<uint32_t*>(gray + (long long)b * gray_frame_stride + (long long)y * gpitch + x) = g;
  uint8_t* d = planes + (long long)b * 3 * plane_stride + (long long)y * ppitch + x;
  *reinterpret_cast<uint32_t*>(d) = p0;
  *reinterpret_cast<uint32_t*>(d + plane_stride) = p1;
  *reinterpret_cast<uint32_t*>(d + 2 * plane_stride) = p2;
  // some pixel with differing channels: the stream is not a gray camera behind a BGR8 conversion (lk_track2_kernel, MUL = 3)
  if (((p0 ^ p1) | (p1 ^ p2)) && __ldcg(colour + b) == 0) atomicOr(colour + b, 1);
}

__global__ void replicate_k_kernel(double* K, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > 0 && b < batch)
    for (int i = 0; i < 9; ++i) K[b * 9 + i] = K[i];
}

// Tracker::track_frame_with_optical_flow's filter (src/tracker.cpp:70-77) for the tracking frame: observations with
// status && err < err_thr are kept in order; their new positions and landmark coordinates become the stream's next
// observation list and the input of solvePnPRansac.  Fewer than 6 kept points: PnP is skipped (npts 0).
__global__ void __launch_bounds__(1024)
track_collect_kernel(const float2* __restrict__ next_xy, const uint8_t* __restrict__ status, const float* __restrict__ err,
                     const float* __restrict__ obj, const int32_t* __restrict__ nprev, int cap, float err_thr,
                     float2* __restrict__ out_xy, float* __restrict__ out_obj, int32_t* __restrict__ out_n,
                     int32_t* __restrict__ src_idx, float2* __restrict__ pnp_img, float* __restrict__ pnp_obj,
                     int32_t* __restrict__ pnp_n, int pnp_cap) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(nprev[b], cap);
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const long long o = (long long)b * cap + i;
    const int ok = (i < n) ? (status[o] != 0 && err[o] < err_thr) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (ok) {
      const int k = off + __popc(bal & ((1u << lane) - 1));
      const long long d = (long long)b * cap + k, dp = (long long)b * pnp_cap + k;
      const float2 p = next_xy[o];
      const float X = obj[o * 3], Y = obj[o * 3 + 1], Z = obj[o * 3 + 2];
      out_xy[d] = p;
      out_obj[d * 3] = X; out_obj[d * 3 + 1] = Y; out_obj[d * 3 + 2] = Z;
      src_idx[d] = i;
      pnp_img[dp] = p;
      pnp_obj[dp * 3] = X; pnp_obj[dp * 3 + 1] = Y; pnp_obj[dp * 3 + 2] = Z;
    }
    __syncthreads();
    if (tid == 0)
      for (int w = 0; w < 32; ++w) s_base += s_warp[w];
    __syncthreads();
  }
  if (tid == 0) {
    out_n[b] = s_base;
    pnp_n[b] = s_base >= 6 ? s_base : 0;
  }
}

__global__ void gather_track_results_kernel(const int32_t* nprev, const int32_t* ntracked, const int32_t* pnp_result,
                                            const double* pose_out, int have_prev, mvo_track_result* out, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  mvo_track_result r;
  memset(&r, 0, sizeof(r));
  if (have_prev) {
    r.n_prev = nprev[b];
    r.n_tracked = ntracked[b];
    const bool ok = r.n_tracked >= 6 && pnp_result[b * 8 + 2] >= 0;
    r.n_pnp_inliers = ok ? pnp_result[b * 8 + 0] : 0;
    r.pnp_ok = ok ? 1 : 0;
    if (ok)
      for (int i = 0; i < 3; ++i) {
        r.rvec[i] = pose_out[b * 8 + i];
        r.tvec[i] = pose_out[b * 8 + 3 + i];
      }
  }
  out[b] = r;
}

}  // namespace mvo

using namespace mvo;

// every stage is also an NVTX range (domain-less, named like mvo_stage_ms' stages) around its enqueue calls, so a
// timeline tool attributes the kernels of a step to ORB / kNN / LK / the model searches (tracing row)
#define STAGE_BEG(c, s)                                                          \
  do {                                                                           \
    nvtxRangePushA(kStageNames[s]);                                              \
    if (!(c)->capturing) cudaEventRecord((c)->timers[s].beg, (c)->stream);       \
  } while (0)
#define STAGE_END(c, s)                                                          \
  do {                                                                           \
    if (!(c)->capturing) {                                                       \
      cudaEventRecord((c)->timers[s].end, (c)->stream);                          \
      (c)->timers[s].used = true;                                                \
    }                                                                            \
    nvtxRangePop();                                                              \
  } while (0)
// wait on an event recorded by an EARLIER step: meaningless (and not allowed) while a step is captured into a graph --
// the graph form only runs synchronous steps, every earlier step is complete
#define WAIT_PREV_STEP(c, st, ev)                          \
  do {                                                     \
    if (!(c)->capturing) cudaStreamWaitEvent(st, ev, 0);   \
  } while (0)

// layout of a slot's pinned output block for the configured outputs
static int out_prepare(mvo_ctx* c, int cap) {
  OutLayout& L = c->out_layout;
  const size_t B = (size_t)c->cfg.batch;
  if (L.cap != cap || L.batch != (int)B || L.mask_bits != c->out_mask) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
      const size_t o = off;
      off += align_up(bytes, 256);
      return o;
    };
    L = OutLayout();
    L.cap = cap;
    L.batch = (int)B;
    L.mask_bits = c->out_mask;
    L.prev_count = take(B * 4);
    if (c->out_mask & MVO_OUT_KEYPOINTS) {
      L.kps = take(B * cap * sizeof(mvo_keypoint));
      L.desc = take(B * cap * 32);
    }
    if (c->out_mask & MVO_OUT_MATCHES) L.matches = take(B * cap * sizeof(mvo_dmatch));
    if (c->out_mask & MVO_OUT_TRACKS) {
      L.lk_xy = take(B * cap * 8);
      L.lk_status = take(B * cap);
      L.lk_err = take(B * cap * 4);
    }
    if (c->out_mask & MVO_OUT_MODELS) {
      for (int k = 0; k < 4; ++k) L.mask[k] = take(B * cap);
      L.models = take(B * 27 * 8);
    }
    if (c->out_mask & MVO_OUT_POINTS3D) L.x4 = take(B * 4 * cap * 4);
    if (c->out_mask & MVO_OUT_CLOUD) L.cloud = take(B * cap * 12);
    L.total = off;
  }
  for (auto& sl : c->slots) MVO_CUDA_TRY(c, sl.h_out.alloc(L.total));
  return MVO_OK;
}

// device -> pinned host copy of `rows` rows of `width` bytes whose source rows are `spitch` bytes apart
static cudaError_t copy_rows_d2h(void* dst, const void* src, size_t width, size_t spitch, size_t rows, cudaStream_t st) {
  if (spitch == width) return cudaMemcpyAsync(dst, src, width * rows, cudaMemcpyDeviceToHost, st);
  return cudaMemcpy2DAsync(dst, width, src, spitch, width, rows, cudaMemcpyDeviceToHost, st);
}

extern "C" {

static int group_enqueue(mvo_ctx* c, const uint8_t* images, int w, int h, int stride, int images_on_device, const double* K,
                         int slot) {
  if (!c) return MVO_ERR_INVALID;
  const int cn = c->grp_cn;
  if (!images || !K || stride < w * cn) {
    c->set_error("mvo_group_step: bad argument");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const int B = c->cfg.batch;
  c->trk_have_frame = false;   // the tracking-frame chain (mvo_group_track) does not survive a front-end step
  c->lk_hash[0] = c->lk_hash[1] = 0;   // ... nor does the single-call pyramid cache (the LK pyramids are rewritten)
  int rc = orb_prepare(c, w, h);
  if (rc) return rc;
  const OrbGeom& g = c->geom;
  const int cap = g.kp_cap;
  if (c->prev_kps.n < (size_t)B * cap) {
    MVO_CUDA_TRY(c, c->prev_kps.alloc((size_t)B * cap));
    MVO_CUDA_TRY(c, c->prev_kp_xy.alloc((size_t)B * cap));
    MVO_CUDA_TRY(c, c->prev_desc.alloc((size_t)B * cap * 32));
    MVO_CUDA_TRY(c, c->prev_kp_count.alloc(B));
    MVO_CUDA_TRY(c, c->d_results.alloc(B));
    c->have_prev = false;
  }
  if (c->lk_w != w || c->lk_h != h || c->lk_max_pts < cap || c->lk_cn != cn) {
    rc = lk_prepare(c, w, h, std::max(cap, c->lk_max_pts), cn);
    if (rc) return rc;
    c->have_prev = false;
  }
  rc = ransac_prepare(c, cap, 2000, RansacBufs::kLanes);
  if (rc) return rc;
  rc = pose_prepare(c);
  if (rc) return rc;
  rc = knn_prepare(c, cap);
  if (rc) return rc;
  RansacBufs& r = c->rs;
  for (auto& t : c->timers) t.used = false;
  const uint32_t om = c->out_mask;
  const OutLayout& L = c->out_layout;
  if (om) {
    rc = out_prepare(c, cap);
    if (rc) return rc;
    if (om & MVO_OUT_MODELS) MVO_CUDA_TRY(c, c->e_mask_keep.alloc((size_t)B * r.max_pts));
    if (om & MVO_OUT_CLOUD) MVO_CUDA_TRY(c, c->cloud.alloc((size_t)B * cap * 3));
  }
  uint8_t* const hout = c->slots[slot].h_out.p;
  c->slots[slot].had_prev = c->have_prev ? 1 : 0;

  // Fork / join: kNN only needs the descriptors, the three model searches only need the LK correspondences, so
  // they run beside each other on auxiliary CUDA streams (each search in its own RansacLane).
  struct Fork {
    mvo_ctx* c;
    Fork(mvo_ctx* c_, cudaStream_t s, int lane, cudaEvent_t after) : c(c_) {
      cudaStreamWaitEvent(s, after, 0);
      c->stream = s;
      c->rs.cur = lane;
    }
    ~Fork() {
      c->stream = c->main_stream;
      c->rs.cur = 0;
    }
  };

  STAGE_BEG(c, ST_TOTAL);
  // ---- ORB ----
  STAGE_BEG(c, ST_ORB);
  {
    // frames -> level 0 of the ORB pyramids.  Host frames go through a double-buffered staging area on the copy
    // stream (one DMA for the whole group), so the upload of step t+1 overlaps the kernels of step t.
    const uint8_t* src = images;
    int spitch = stride;
    long long sstride = (long long)h * stride;
    const int rowb = w * cn;     // bytes of one frame row
    if (!images_on_device && c->capturing) {
      src = c->slots[slot].stage.p;      // filled by mvo_group_step right before the graph launch
      spitch = rowb;
      sstride = (long long)h * rowb;
    } else if (!images_on_device) {
      GroupSlot& sl = c->slots[slot];
      MVO_CUDA_TRY(c, sl.stage.alloc((size_t)B * h * rowb));
      cudaStreamWaitEvent(c->copy_stream, sl.ev_free, 0);
      if (stride == rowb)   // contiguous frames: one flat DMA (a 2-D copy of 1241-byte rows runs far below PCIe speed)
        MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.stage.p, images, (size_t)B * h * rowb, cudaMemcpyHostToDevice, c->copy_stream));
      else
        MVO_CUDA_TRY(c, cudaMemcpy2DAsync(sl.stage.p, rowb, images, stride, rowb, (size_t)B * h, cudaMemcpyHostToDevice,
                                          c->copy_stream));
      cudaEventRecord(sl.ev_up, c->copy_stream);
      cudaStreamWaitEvent(c->main_stream, sl.ev_up, 0);
      src = sl.stage.p;
      spitch = rowb;
      sstride = (long long)h * rowb;
    }
    const LevelGeom& l0 = g.lv[0];
    int lk_pitch = 0;
    long long lk_fs = 0;
    uint8_t* lk0 = lk_level0(c, c->lk_cur, &lk_pitch, &lk_fs);
    // The LK pyramid slot about to be overwritten is the "previous" pyramid of the last enqueued step's tracker, which
    // runs on its own stream; the keypoint-position buffer ORB is about to fill is that tracker's input as well.
    WAIT_PREV_STEP(c, c->main_stream, c->ev_lk_done);
    if (cn == 1) {
      dim3 grid((w + 16 * 256 - 1) / (16 * 256), h, B);
      unpack_frames_kernel<<<grid, 256, 0, c->stream>>>(src, spitch, sstride, c->pyr.p + l0.off, l0.pitch, g.frame_stride,
                                                        lk0, lk_pitch, lk_fs, w, h);
    } else {
      dim3 grid((w + 4 * 256 - 1) / (4 * 256), h, B);
      MVO_CUDA_TRY(c, cudaMemsetAsync(c->lk_colour[c->lk_cur].p, 0, sizeof(int32_t) * B, c->stream));
      unpack_bgr_frames_kernel<<<grid, 256, 0, c->stream>>>(src, spitch, sstride, c->pyr.p + l0.off, l0.pitch,
                                                            g.frame_stride, lk0, lk_pitch, lk_fs, w, h,
                                                            c->lk_colour[c->lk_cur].p);
    }
    c->launches++;
    if (!images_on_device && !c->capturing) cudaEventRecord(c->slots[slot].ev_free, c->main_stream);
    cudaEventRecord(c->ev_unpack, c->main_stream);
  }
  const int cur = c->lk_cur, prev = cur ^ 1;
  // ---- LK pyramid of the new frame + tracker: needs the unpacked frame and the PREVIOUS frame's keypoints only, so it
  // runs beside ORB(t) on its own stream -- for a single stream ORB -> LK -> searches was the critical path, now it is
  // max(ORB, LK -> searches); in a group the small pyramid levels of ORB fill the tracker's tail and vice versa ----
  int32_t *res_h = nullptr, *res_f = nullptr, *ntri = nullptr;
  const int gb = (B + 127) / 128;
  if (c->have_prev) {
    // counts kept per stage
    MVO_CUDA_TRY(c, c->knn_counts.alloc((size_t)4 * B));
    res_h = c->knn_counts.p;
    res_f = res_h + B;
    ntri = res_f + B;
  }
  {
    Fork f(c, c->lk_stream, 0, c->ev_unpack);
    STAGE_BEG(c, ST_LK);
    rc = lk_build_pyramid(c, cur, nullptr, 0, 3);   // level 0 was written by the unpack kernel
    if (rc) return rc;
    if (c->have_prev) {
      WAIT_PREV_STEP(c, c->stream, c->ev_out_lk);   // the previous step's track outputs have left the LK buffers
      rc = lk_run(c, prev, cur, c->prev_kp_xy.p, c->prev_kp_count.p, cap, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p);
      if (rc) return rc;
      // the correspondence buffers feed the model searches of the previous step until its tail is done; from here on
      // this step's ORB + LK have overlapped them (software pipelining across the two steps in flight)
      WAIT_PREV_STEP(c, c->stream, c->ev_tail);
      lk_collect_kernel<<<B, 1024, 0, c->stream>>>(c->prev_kp_xy.p, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p,
                                                   c->prev_kp_count.p, cap, 30.0f, r.p1.p, r.p2.p, r.npts.p);
      c->launches++;
      if (om & MVO_OUT_TRACKS) {
        cudaEventRecord(c->ev_o_lk, c->stream);
        cudaStreamWaitEvent(c->out_stream, c->ev_o_lk, 0);
        MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.lk_xy, c->lk_pts_out.p, (size_t)B * cap * 8, cudaMemcpyDeviceToHost, c->out_stream));
        MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.lk_status, c->lk_status.p, (size_t)B * cap, cudaMemcpyDeviceToHost, c->out_stream));
        MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.lk_err, c->lk_err.p, (size_t)B * cap * 4, cudaMemcpyDeviceToHost, c->out_stream));
        cudaEventRecord(c->ev_out_lk, c->out_stream);
      }
    }
    cudaEventRecord(c->ev_lk_done, c->stream);
    STAGE_END(c, ST_LK);
    if (c->have_prev) {
      if (!c->capturing)   // (graph form: K was copied before the graph launch)
        MVO_CUDA_TRY(c, cudaMemcpyAsync(r.K.p, K, 72, cudaMemcpyHostToDevice, c->stream));
      fill_params_kernel<<<gb, 128, 0, c->stream>>>(r.lane[0].thr2.p, 1.0f, r.K.p, B);   // + replicate K to every stream
      c->launches++;
      cudaEventRecord(c->ev_fork[1], c->stream);
    }
  }

  // (the tracker is enqueued before ORB so that it starts as soon as the frame is unpacked, not after the host has
  // issued the ORB launches)
  // The model searches depend on the tracker only.  The essential-matrix chain is the longest dependency chain of a frame
  // (5-point solver -> recoverPose -> triangulation), the homography chain the second longest: they are issued before
  // the ~20 ORB launches so that a single stream does not wait for the host to get to them (F, the shortest, after ORB).
  if (c->have_prev) {
    {
      // ---- E RANSAC (K, conf 0.99, thr 1.0) -> recoverPose -> triangulate + chirality, lane 2 ----
      Fork f(c, c->aux_stream[1], 2, c->ev_fork[1]);
      STAGE_BEG(c, ST_E);
      const double t = 1.0 / ((K[0] + K[4]) / 2);
      fill_params_kernel<<<gb, 128, 0, c->stream>>>(r.ln().thr2.p, (float)(t * t), nullptr, B);
      c->launches++;
      rc = ransac_normalize(c);
      if (rc) return rc;
      rc = ransac_find(c, MVO_MODEL_E, 0.99);
      if (rc) return rc;
      STAGE_END(c, ST_E);
      if (om & MVO_OUT_MODELS)   // recoverPose rewrites the mask in place (mask_io): keep findEssentialMat's own
        MVO_CUDA_TRY(c, cudaMemcpyAsync(c->e_mask_keep.p, r.ln().mask.p, (size_t)B * r.max_pts, cudaMemcpyDeviceToDevice, c->stream));
      STAGE_BEG(c, ST_POSE);
      rc = pose_recover(c, true);
      if (rc) return rc;
      STAGE_END(c, ST_POSE);
      STAGE_BEG(c, ST_TRI);
      make_proj_kernel<<<gb, 128, 0, c->stream>>>(r.K.p, r.pose.p, r.proj.p, B);
      c->launches++;
      rc = pose_triangulate(c);
      if (rc) return rc;
      tri_count_kernel<<<B, kTriCountThreads, 0, c->stream>>>(r.X4.p, r.ln().mask.p, r.pose.p, r.npts.p, r.max_pts, ntri,
                                                 (om & MVO_OUT_CLOUD) ? c->cloud.p : nullptr, cap);
      c->launches++;
      STAGE_END(c, ST_TRI);
      cudaEventRecord(c->ev_join[1], c->stream);
    }
    {
      // ---- H RANSAC (thr 1.0), lane 0, on the tail stream (the join of all searches follows there, further down) ----
      Fork f(c, c->aux_stream[3], 0, c->ev_fork[1]);
      STAGE_BEG(c, ST_H);
      rc = ransac_find(c, MVO_MODEL_H, 0.995);
      if (rc) return rc;
      copy_i32_strided_kernel<<<gb, 128, 0, c->stream>>>(r.ln().result.p, 8, res_h, B);
      c->launches++;
      STAGE_END(c, ST_H);
    }
  }
  // the keypoint / descriptor buffers ORB is about to fill were the previous step's "prev" set: its kNN must be done
  WAIT_PREV_STEP(c, c->main_stream, c->ev_join[2]);
  rc = orb_run_detect(c, true);
  if (rc) return rc;
  {
    GroupSlot& sl = c->slots[slot];
    MVO_CUDA_TRY(c, sl.h_res.alloc(B));
    MVO_CUDA_TRY(c, sl.h_flags.alloc(B));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.h_flags.p, c->flags.p, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
    MVO_CUDA_TRY(c, sl.h_occ.alloc(2 * B));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.h_occ.p, c->occ.p, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
    if (om && c->have_prev)
      MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.prev_count, c->prev_kp_count.p, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  STAGE_END(c, ST_ORB);
  cudaEventRecord(c->ev_fork[0], c->main_stream);
  if (om & MVO_OUT_KEYPOINTS) {
    // keypoints + descriptors of the new frame leave on the output stream while the step goes on
    cudaEventRecord(c->ev_o_orb, c->main_stream);
    cudaStreamWaitEvent(c->out_stream, c->ev_o_orb, 0);
    MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.kps, c->kps.p, (size_t)B * cap * sizeof(mvo_keypoint), cudaMemcpyDeviceToHost,
                                    c->out_stream));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.desc, c->desc.p, (size_t)B * cap * 32, cudaMemcpyDeviceToHost, c->out_stream));
    cudaEventRecord(c->ev_out_orb, c->out_stream);
  }

  if (c->have_prev) {
    // ---- kNN + ratio: query = previous descriptors, train = new descriptors (src/tracker.cpp:190-191) ----
    Fork f(c, c->aux_stream[2], 0, c->ev_fork[0]);
    WAIT_PREV_STEP(c, c->stream, c->ev_tail);   // the previous step's gather still reads the match counters
    STAGE_BEG(c, ST_KNN);
    rc = knn_run(c, c->prev_desc.p, c->prev_kp_count.p, cap, cap, c->desc.p, c->kp_count.p, cap, cap, 0.7, B);
    if (rc) return rc;
    STAGE_END(c, ST_KNN);
    cudaEventRecord(c->ev_join[2], c->stream);
  }
  if (c->have_prev) {
    {
      // ---- F RANSAC (thr 1.0, conf 0.99), lane 1 ----
      Fork f(c, c->aux_stream[0], 1, c->ev_fork[1]);
      STAGE_BEG(c, ST_F);
      fill_params_kernel<<<gb, 128, 0, c->stream>>>(r.ln().thr2.p, 1.0f, nullptr, B);
      c->launches++;
      rc = ransac_find(c, MVO_MODEL_F, 0.99);
      if (rc) return rc;
      copy_i32_strided_kernel<<<gb, 128, 0, c->stream>>>(r.ln().result.p, 8, res_f, B);
      c->launches++;
      STAGE_END(c, ST_F);
      cudaEventRecord(c->ev_join[0], c->stream);
    }
    {
      // ---- the join of all searches and the result gather: tail stream (after the H search enqueued above) ----
      Fork f(c, c->aux_stream[3], 0, c->ev_fork[1]);
      for (int k = 0; k < 3; ++k) cudaStreamWaitEvent(c->stream, c->ev_join[k], 0);
      gather_results_kernel<<<gb, 128, 0, c->stream>>>(c->kp_count.p, c->knn_nmatch.p, r.npts.p, res_h, res_f,
                                                       r.lane[2].result.p, ntri, r.pose.p, 1, c->d_results.p, B);
      c->launches++;
      MVO_CUDA_TRY(c, cudaMemcpyAsync(c->slots[slot].h_res.p, c->d_results.p, (size_t)B * sizeof(mvo_frame_result),
                                      cudaMemcpyDeviceToHost, c->stream));
      // full outputs that the next step's kNN / searches overwrite: they leave before ev_tail releases those buffers
      if (om & MVO_OUT_MATCHES)
        MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.matches, c->knn_matches.p, (size_t)B * cap * sizeof(mvo_dmatch),
                                        cudaMemcpyDeviceToHost, c->stream));
      if (om & MVO_OUT_MODELS) {
        const uint8_t* msrc[4] = {r.lane[0].mask.p, r.lane[1].mask.p, c->e_mask_keep.p, r.lane[2].mask.p};
        for (int k = 0; k < 4; ++k)
          MVO_CUDA_TRY(c, copy_rows_d2h(hout + L.mask[k], msrc[k], (size_t)cap, (size_t)r.max_pts, (size_t)B, c->stream));
        for (int k = 0; k < 3; ++k)
          MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.models + (size_t)k * B * 72, r.lane[k].best_model.p, (size_t)B * 72,
                                          cudaMemcpyDeviceToHost, c->stream));
      }
      if (om & MVO_OUT_POINTS3D)
        MVO_CUDA_TRY(c, copy_rows_d2h(hout + L.x4, r.X4.p, (size_t)cap * 4, (size_t)r.max_pts * 4, (size_t)B * 4, c->stream));
      if (om & MVO_OUT_CLOUD)
        MVO_CUDA_TRY(c, cudaMemcpyAsync(hout + L.cloud, c->cloud.p, (size_t)B * cap * 12, cudaMemcpyDeviceToHost, c->stream));
      STAGE_END(c, ST_TOTAL);
      cudaEventRecord(c->ev_tail, c->stream);
      if (om & MVO_OUT_KEYPOINTS) cudaStreamWaitEvent(c->stream, c->ev_out_orb, 0);
      if (om & MVO_OUT_TRACKS) cudaStreamWaitEvent(c->stream, c->ev_out_lk, 0);
      cudaEventRecord(c->slots[slot].ev_done, c->stream);
    }
  } else {
    Fork f(c, c->aux_stream[3], 0, c->ev_fork[0]);
    cudaStreamWaitEvent(c->stream, c->ev_lk_done, 0);   // the step is done when its LK pyramid is built, too
    gather_results_kernel<<<(B + 127) / 128, 128, 0, c->stream>>>(c->kp_count.p, nullptr, nullptr, nullptr, nullptr,
                                                                 nullptr, nullptr, nullptr, 0, c->d_results.p, B);
    c->launches++;
    MVO_CUDA_TRY(c, cudaMemcpyAsync(c->slots[slot].h_res.p, c->d_results.p, (size_t)B * sizeof(mvo_frame_result),
                                    cudaMemcpyDeviceToHost, c->stream));
    STAGE_END(c, ST_TOTAL);
    cudaEventRecord(c->ev_tail, c->stream);
    if (om & MVO_OUT_KEYPOINTS) cudaStreamWaitEvent(c->stream, c->ev_out_orb, 0);
    cudaEventRecord(c->slots[slot].ev_done, c->stream);
  }
  MVO_CUDA_TRY(c, cudaGetLastError());
  // new frame becomes the previous one (host-side bookkeeping: the enqueued kernels already hold their pointers)
  std::swap(c->kps, c->prev_kps);
  std::swap(c->kp_xy, c->prev_kp_xy);
  std::swap(c->desc, c->prev_desc);
  std::swap(c->kp_count, c->prev_kp_count);
  c->lk_cur ^= 1;
  c->have_prev = true;
  return MVO_OK;
}

// wait for the step enqueued in `slot` and hand its results out
static int group_finish(mvo_ctx* c, int slot, mvo_frame_result* results) {
  GroupSlot& sl = c->slots[slot];
  const int B = c->cfg.batch;
  MVO_CUDA_TRY(c, cudaEventSynchronize(sl.ev_done));
  memcpy(results, sl.h_res.p, (size_t)B * sizeof(mvo_frame_result));
  c->out_slot = slot;
  c->occ_from_group = true;
  int flags0 = 0, first = -1;
  for (int b = 0; b < B; ++b) {
    if (sl.h_flags.p[b] && first < 0) first = b;
    flags0 |= sl.h_flags.p[b];
  }
  // same rule as the single-call path: an overflow of either list is an error for the step; the records of the other
  // streams are valid and mvo_group_outputs(...).flags says which streams were hit
  if (flags0 & 3) {
    c->set_error(std::string(flags0 & 1 ? "FAST candidate list overflow" : "keypoint capacity exceeded") + " (first on stream " +
                 std::to_string(first) + ")");
    return MVO_ERR_CAPACITY;
  }
  return MVO_OK;
}

// ---- CUDA-graph form of the synchronous step -----------------------------------------------------------------------
// What a graph is valid for: everything the enqueue code reads besides device memory contents.
static unsigned long long step_graph_key(const mvo_ctx* c, int w, int h) {
  unsigned long long k = 1469598103934665603ull;
  auto mix = [&k](unsigned long long v) { k = (k ^ v) * 1099511628211ull; };
  mix((unsigned long long)w); mix((unsigned long long)h); mix((unsigned long long)c->grp_cn); mix(c->out_mask);
  mix((unsigned long long)c->occupancy_div); mix((unsigned long long)c->dbg_lk_impl); mix((unsigned long long)c->dbg_knn_impl);
  mix((unsigned long long)c->dbg_h_refine_impl); mix((unsigned long long)c->dbg_e5_roots_impl);
  // which physical buffers are "current" and "previous" right now (they swap every frame)
  mix((unsigned long long)(uintptr_t)c->kps.p); mix((unsigned long long)(uintptr_t)c->prev_kps.p);
  mix((unsigned long long)(uintptr_t)c->desc.p); mix((unsigned long long)(uintptr_t)c->kp_xy.p);
  mix((unsigned long long)(uintptr_t)c->kp_count.p); mix((unsigned long long)c->lk_cur);
  mix((unsigned long long)(uintptr_t)c->main_stream);
  return k;
}
// After a capture the cross-step events were last "recorded" inside the capture; give them a real (already complete)
// record so that a later plain step can wait on them.  All streams are idle here (synchronous steps only).
static void rearm_events(mvo_ctx* c, int slot) {
  cudaEventRecord(c->ev_tail, c->aux_stream[3]);
  cudaEventRecord(c->ev_lk_done, c->lk_stream);
  cudaEventRecord(c->ev_unpack, c->main_stream);
  cudaEventRecord(c->ev_fork[0], c->main_stream);
  cudaEventRecord(c->ev_fork[1], c->lk_stream);
  cudaEventRecord(c->ev_join[0], c->aux_stream[0]);
  cudaEventRecord(c->ev_join[1], c->aux_stream[1]);
  cudaEventRecord(c->ev_join[2], c->aux_stream[2]);
  cudaEventRecord(c->ev_o_orb, c->main_stream);
  cudaEventRecord(c->ev_o_lk, c->lk_stream);
  cudaEventRecord(c->ev_out_orb, c->out_stream);
  cudaEventRecord(c->ev_out_lk, c->out_stream);
  cudaEventRecord(c->slots[slot].ev_done, c->aux_stream[3]);
}

// Small groups on host frames: the host's ~85 launches per step (0.35 - 0.45 ms), not the GPU's critical path (0.42 ms for
// one 1241 x 376 stream), bound the synchronous step.  The step is therefore captured once per buffer parity into a CUDA
// graph (every stream of the fork / join schedule above becomes a branch) and replayed with one launch; the frames and K
// are copied by plain asynchronous copies in front of the graph, so no node ever needs new parameters.
static int group_step_graph(mvo_ctx* c, const uint8_t* images, int w, int h, int stride, const double* K,
                            mvo_frame_result* results, bool* used) {
  *used = false;
  const int B = c->cfg.batch, cn = c->grp_cn;
  const size_t rowb = (size_t)w * cn, bytes = (size_t)B * h * rowb;
  if (!c->graph_enabled || !c->have_prev || bytes > (4u << 20) || c->steps_since_change < 1) return MVO_OK;
  GroupSlot& sl = c->slots[0];
  if (!sl.stage.p || sl.stage.n < bytes || !c->rs.K.p) return MVO_OK;
  const int parity = c->lk_cur & 1;
  mvo_ctx::StepGraph& gph = c->step_graph[parity];
  const unsigned long long key = step_graph_key(c, w, h);
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  // inputs of the step: plain copies in front of the graph (all earlier work is complete: synchronous steps)
  if ((size_t)stride == rowb)
    MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.stage.p, images, bytes, cudaMemcpyHostToDevice, c->main_stream));
  else
    MVO_CUDA_TRY(c, cudaMemcpy2DAsync(sl.stage.p, rowb, images, stride, rowb, (size_t)B * h, cudaMemcpyHostToDevice, c->main_stream));
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->rs.K.p, K, 72, cudaMemcpyHostToDevice, c->main_stream));
  if (!gph.exec || gph.key != key || gph.epoch != alloc_epoch()) {
    if (gph.exec) {
      cudaGraphExecDestroy(gph.exec);
      gph.exec = nullptr;
    }
    const unsigned long long epoch0 = alloc_epoch();
    const unsigned long long l0 = c->launches;
    cudaGraph_t graph = nullptr;
    MVO_CUDA_TRY(c, cudaStreamBeginCapture(c->main_stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    int rc = group_enqueue(c, images, w, h, stride, 0, K, 0);
    if (rc == MVO_OK) cudaStreamWaitEvent(c->main_stream, sl.ev_done, 0);   // join the tail (and through it every branch)
    c->capturing = false;
    const cudaError_t e = cudaStreamEndCapture(c->main_stream, &graph);
    if (rc != MVO_OK || e != cudaSuccess || !graph || alloc_epoch() != epoch0) {
      // (an allocation inside the capture, or a capture error: fall back to the plain step for this frame)
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      rearm_events(c, 0);
      c->steps_since_change = 0;
      c->graph_stats[2]++;
      if (rc != MVO_OK) return rc;
      // group_enqueue already advanced the bookkeeping (buffer swap): undo it, the plain path redoes the step
      std::swap(c->kps, c->prev_kps);
      std::swap(c->kp_xy, c->prev_kp_xy);
      std::swap(c->desc, c->prev_desc);
      std::swap(c->kp_count, c->prev_kp_count);
      c->lk_cur ^= 1;
      c->launches = l0;
      return MVO_OK;
    }
    const cudaError_t ei = cudaGraphInstantiate(&gph.exec, graph, 0);
    cudaGraphDestroy(graph);
    rearm_events(c, 0);
    if (ei != cudaSuccess) {
      gph.exec = nullptr;
      cudaGetLastError();
      std::swap(c->kps, c->prev_kps);
      std::swap(c->kp_xy, c->prev_kp_xy);
      std::swap(c->desc, c->prev_desc);
      std::swap(c->kp_count, c->prev_kp_count);
      c->lk_cur ^= 1;
      c->launches = l0;
      return MVO_OK;
    }
    gph.key = key;
    gph.epoch = epoch0;
    gph.launches = (int)(c->launches - l0);
    c->graph_stats[0]++;
  } else {
    // replay: the host-side bookkeeping group_enqueue does at its end
    c->trk_have_frame = false;
    c->lk_hash[0] = c->lk_hash[1] = 0;
    for (auto& t : c->timers) t.used = false;
    c->slots[0].had_prev = 1;
    std::swap(c->kps, c->prev_kps);
    std::swap(c->kp_xy, c->prev_kp_xy);
    std::swap(c->desc, c->prev_desc);
    std::swap(c->kp_count, c->prev_kp_count);
    c->lk_cur ^= 1;
    c->have_prev = true;
    c->launches += (unsigned long long)gph.launches;
    c->graph_stats[1]++;
  }
  MVO_CUDA_TRY(c, cudaGraphLaunch(gph.exec, c->main_stream));
  MVO_CUDA_TRY(c, cudaEventRecord(c->ev_graph_done, c->main_stream));
  MVO_CUDA_TRY(c, cudaEventSynchronize(c->ev_graph_done));
  *used = true;
  // results as group_finish hands them out (the slot's ev_done belongs to the graph: completion is ev_graph_done)
  memcpy(results, sl.h_res.p, (size_t)B * sizeof(mvo_frame_result));
  c->out_slot = 0;
  c->occ_from_group = true;
  int flags0 = 0, first = -1;
  for (int b = 0; b < B; ++b) {
    if (sl.h_flags.p[b] && first < 0) first = b;
    flags0 |= sl.h_flags.p[b];
  }
  if (flags0 & 3) {
    if ((flags0 & 1) && c->cand_scale < 16) {
      // the candidate lists are doubled for the steps that follow (the geometry is rebuilt, the previous frame forgotten)
      c->cand_scale *= 2;
      c->geom_w = c->geom_h = -1;
      c->have_prev = false;
    }
    c->set_error(std::string(flags0 & 1 ? "FAST candidate list overflow (capacity doubled for the next step)" : "keypoint capacity exceeded") +
                 " (first on stream " + std::to_string(first) + ")");
    return MVO_ERR_CAPACITY;
  }
  return MVO_OK;
}

int mvo_group_step(mvo_ctx* c, const uint8_t* images, int w, int h, int stride, int images_on_device, const double* K,
                   mvo_frame_result* results) {
  if (!c) return MVO_ERR_INVALID;
  if (!results) {
    c->set_error("mvo_group_step: bad argument");
    return MVO_ERR_INVALID;
  }
  if (c->q_count != 0) {
    c->set_error("mvo_group_step: submitted steps are still in flight (mvo_group_collect them first)");
    return MVO_ERR_INVALID;
  }
  if (!images_on_device && images && K && stride >= w * c->grp_cn && w == c->geom_w && h == c->geom_h) {
    bool used = false;
    const int rcg = group_step_graph(c, images, w, h, stride, K, results, &used);
    if (rcg != MVO_OK || used) return rcg;
  }
  const unsigned long long e0 = alloc_epoch();
  int rc = group_enqueue(c, images, w, h, stride, images_on_device, K, 0);
  if (rc) return rc;
  rc = group_finish(c, 0, results);
  c->steps_since_change = (alloc_epoch() == e0) ? c->steps_since_change + 1 : 0;
  return rc;
}

int mvo_group_submit(mvo_ctx* c, const uint8_t* images, int w, int h, int stride, int images_on_device, const double* K) {
  if (!c) return MVO_ERR_INVALID;
  if (c->q_count >= mvo_ctx::kSlots) {
    c->set_error("mvo_group_submit: two steps are already in flight");
    return MVO_ERR_CAPACITY;
  }
  const int slot = (c->q_head + c->q_count) % mvo_ctx::kSlots;
  int rc = group_enqueue(c, images, w, h, stride, images_on_device, K, slot);
  if (rc) return rc;
  c->q_count++;
  return MVO_OK;
}

int mvo_group_collect(mvo_ctx* c, mvo_frame_result* results) {
  if (!c) return MVO_ERR_INVALID;
  if (!results || c->q_count == 0) {
    c->set_error("mvo_group_collect: nothing in flight");
    return MVO_ERR_INVALID;
  }
  const int slot = c->q_head;
  c->q_head = (c->q_head + 1) % mvo_ctx::kSlots;
  c->q_count--;
  return group_finish(c, slot, results);
}

int mvo_group_reset(mvo_ctx* c) {
  if (!c) return MVO_ERR_INVALID;
  if (c->q_count != 0) {
    c->set_error("mvo_group_reset: submitted steps are still in flight");
    return MVO_ERR_INVALID;
  }
  c->have_prev = false;
  return MVO_OK;
}

int mvo_group_configure(mvo_ctx* c, const mvo_group_config* cfg) {
  if (!c || !cfg) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  if ((cfg->channels != 1 && cfg->channels != 3) || (cfg->outputs & ~(uint32_t)MVO_OUT_ALL)) {
    c->set_error("mvo_group_configure: channels must be 1 or 3, outputs a mask of MVO_OUT_*");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->main_stream));
  if (cfg->channels != c->grp_cn) {
    c->have_prev = false;
    c->trk_have_frame = false;
  }
  c->grp_cn = cfg->channels;
  c->out_mask = cfg->outputs;
  c->out_slot = -1;
  return MVO_OK;
}

int mvo_group_output_bytes(mvo_ctx* c, size_t* bytes) {
  if (!c || !bytes) return MVO_ERR_INVALID;
  const size_t B = (size_t)c->cfg.batch;
  const size_t cap = (size_t)(c->cfg.nfeatures + c->cfg.nfeatures / 4 + 64);
  size_t n = B * sizeof(mvo_frame_result) + B * 4;   // result records + overflow flags
  const uint32_t om = c->out_mask;
  if (om) n += B * 4;
  if (om & MVO_OUT_KEYPOINTS) n += B * cap * (sizeof(mvo_keypoint) + 32);
  if (om & MVO_OUT_MATCHES) n += B * cap * sizeof(mvo_dmatch);
  if (om & MVO_OUT_TRACKS) n += B * cap * 13;
  if (om & MVO_OUT_MODELS) n += B * cap * 4 + B * 27 * 8;
  if (om & MVO_OUT_POINTS3D) n += B * cap * 16;
  if (om & MVO_OUT_CLOUD) n += B * cap * 12;
  *bytes = n;
  return MVO_OK;
}

int mvo_group_outputs(mvo_ctx* c, int stream, mvo_stream_outputs* out) {
  if (!c || !out) return MVO_ERR_INVALID;
  if (stream < 0 || stream >= c->cfg.batch || c->out_slot < 0) {
    c->set_error("mvo_group_outputs: no finished step, or stream out of range");
    return MVO_ERR_INVALID;
  }
  const GroupSlot& sl = c->slots[c->out_slot];
  const OutLayout& L = c->out_layout;
  const uint32_t om = L.mask_bits;
  const mvo_frame_result& r = sl.h_res.p[stream];
  const size_t B = (size_t)c->cfg.batch, cap = (size_t)L.cap, b = (size_t)stream;
  const uint8_t* h = sl.h_out.p;
  memset(out, 0, sizeof(*out));
  out->n_keypoints = r.n_keypoints;
  out->flags = sl.h_flags.p[stream];
  if (sl.h_occ.p) {
    out->occupied_cells = sl.h_occ.p[2 * stream];
    out->total_cells = sl.h_occ.p[2 * stream + 1];
  }
  if (!om || !h) return MVO_OK;
  if (om & MVO_OUT_KEYPOINTS) {
    out->keypoints = reinterpret_cast<const mvo_keypoint*>(h + L.kps) + b * cap;
    out->descriptors = h + L.desc + b * cap * 32;
  }
  if (!sl.had_prev) return MVO_OK;
  out->n_matches = r.n_matches;
  out->n_tracked = r.n_tracked;
  out->n_prev = reinterpret_cast<const int32_t*>(h + L.prev_count)[stream];
  if (om & MVO_OUT_MATCHES) out->matches = reinterpret_cast<const mvo_dmatch*>(h + L.matches) + b * cap;
  if (om & MVO_OUT_TRACKS) {
    out->track_xy = reinterpret_cast<const float*>(h + L.lk_xy) + b * cap * 2;
    out->track_status = h + L.lk_status + b * cap;
    out->track_err = reinterpret_cast<const float*>(h + L.lk_err) + b * cap;
  }
  if (om & MVO_OUT_MODELS) {
    out->mask_h = h + L.mask[0] + b * cap;
    out->mask_f = h + L.mask[1] + b * cap;
    out->mask_e = h + L.mask[2] + b * cap;
    out->mask_pose = h + L.mask[3] + b * cap;
    const double* m = reinterpret_cast<const double*>(h + L.models);
    memcpy(out->H, m + (0 * B + b) * 9, 72);
    memcpy(out->F, m + (1 * B + b) * 9, 72);
    memcpy(out->E, m + (2 * B + b) * 9, 72);
  }
  if (om & MVO_OUT_POINTS3D) {
    out->X4 = reinterpret_cast<const float*>(h + L.x4) + b * 4 * cap;
    out->x4_stride = (int64_t)cap;
  }
  if (om & MVO_OUT_CLOUD) {
    out->cloud_xyz = reinterpret_cast<const float*>(h + L.cloud) + b * cap * 3;
    out->n_cloud = r.n_triangulated;
  }
  return MVO_OK;
}

// ---- tracking frame (Tracker::update's per-frame path) ---------------------------------------------------------------
static int track_prepare(mvo_ctx* c, int w, int h) {
  const size_t B = (size_t)c->cfg.batch;
  const int cap = std::max(c->cfg.max_points, c->cfg.nfeatures + c->cfg.nfeatures / 4 + 64);
  if (c->trk_cap != cap) {
    for (int k = 0; k < 2; ++k) {
      MVO_CUDA_TRY(c, c->trk_xy[k].alloc(B * cap));
      MVO_CUDA_TRY(c, c->trk_obj[k].alloc(B * cap * 3));
      MVO_CUDA_TRY(c, c->trk_n[k].alloc(B));
      MVO_CUDA_TRY(c, cudaMemsetAsync(c->trk_n[k].p, 0, B * 4, c->main_stream));
    }
    MVO_CUDA_TRY(c, c->trk_src.alloc(B * cap));
    MVO_CUDA_TRY(c, c->d_trk_res.alloc(B));
    MVO_CUDA_TRY(c, c->h_trk_res.alloc(B));
    c->trk_cap = cap;
    c->trk_cur = 0;
  }
  if (w > 0 && (c->lk_w != w || c->lk_h != h || c->lk_max_pts < cap || c->lk_cn != c->grp_cn)) {
    const int rc = lk_prepare(c, w, h, std::max(cap, c->lk_max_pts), c->grp_cn);
    if (rc) return rc;
    c->have_prev = false;
    c->trk_have_frame = false;
  }
  return MVO_OK;
}

int mvo_group_set_tracks(mvo_ctx* c, int stream, const float* xy, const float* xyz, int n) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  if (stream < 0 || stream >= c->cfg.batch || n < 0 || (n > 0 && (!xy || !xyz))) {
    c->set_error("mvo_group_set_tracks: bad argument");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  int rc = track_prepare(c, 0, 0);
  if (rc) return rc;
  if (n > c->trk_cap) {
    c->set_error("mvo_group_set_tracks: more observations than mvo_config.max_points");
    return MVO_ERR_CAPACITY;
  }
  const size_t o = (size_t)stream * c->trk_cap;
  const int k = c->trk_cur;
  if (n > 0) {
    MVO_CUDA_TRY(c, cudaMemcpyAsync(c->trk_xy[k].p + o, xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->main_stream));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(c->trk_obj[k].p + o * 3, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, c->main_stream));
  }
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->trk_n[k].p + stream, &n, 4, cudaMemcpyHostToDevice, c->main_stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->main_stream));
  return MVO_OK;
}

int mvo_group_track(mvo_ctx* c, const uint8_t* images, int w, int h, int stride, int images_on_device, const double* K,
                    mvo_track_result* results) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  const int cn = c->grp_cn;
  if (!images || !K || !results || stride < w * cn || w > c->cfg.max_width || h > c->cfg.max_height || w < 32 || h < 32) {
    c->set_error("mvo_group_track: bad argument");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const int B = c->cfg.batch;
  c->stream = c->main_stream;
  c->have_prev = false;   // the front-end chain (mvo_group_step) does not survive a tracking frame
  c->lk_hash[0] = c->lk_hash[1] = 0;
  int rc = track_prepare(c, w, h);
  if (rc) return rc;
  const int cap = c->trk_cap;
  rc = ransac_prepare(c, std::max(c->rs.max_pts, 64), std::max(c->rs.cap_iters, 1));   // the RNG table lives there
  if (rc) return rc;
  rc = pnp_prepare(c, std::max(cap, c->pnp.max_pts), std::max(100, c->pnp.cap_iters));
  if (rc) return rc;
  PnpBufs& p = c->pnp;
  StageTimer& tt = c->timers[ST_TOTAL];
  cudaEventRecord(tt.beg, c->stream);
  // ---- frames -> level 0 of the LK pyramid that becomes "next" ----
  {
    const int rowb = w * cn;
    const uint8_t* src = images;
    int spitch = stride;
    long long sstride = (long long)h * stride;
    if (!images_on_device) {
      GroupSlot& sl = c->slots[0];
      MVO_CUDA_TRY(c, sl.stage.alloc((size_t)B * h * rowb));
      if (stride == rowb)
        MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.stage.p, images, (size_t)B * h * rowb, cudaMemcpyHostToDevice, c->stream));
      else
        MVO_CUDA_TRY(c, cudaMemcpy2DAsync(sl.stage.p, rowb, images, stride, rowb, (size_t)B * h, cudaMemcpyHostToDevice, c->stream));
      src = sl.stage.p;
      spitch = rowb;
      sstride = (long long)h * rowb;
    }
    int lk_pitch = 0;
    long long lk_fs = 0;
    uint8_t* lk0 = lk_level0(c, c->lk_cur, &lk_pitch, &lk_fs);
    if (cn == 1) {
      dim3 grid((w + 16 * 256 - 1) / (16 * 256), h, B);
      unpack_frames_kernel<<<grid, 256, 0, c->stream>>>(src, spitch, sstride, lk0, lk_pitch, lk_fs, nullptr, 0, 0, w, h);
    } else {
      dim3 grid((w + 4 * 256 - 1) / (4 * 256), h, B);
      MVO_CUDA_TRY(c, cudaMemsetAsync(c->lk_colour[c->lk_cur].p, 0, sizeof(int32_t) * B, c->stream));
      unpack_bgr_frames_kernel<<<grid, 256, 0, c->stream>>>(src, spitch, sstride, nullptr, 0, 0, lk0, lk_pitch, lk_fs, w, h,
                                                            c->lk_colour[c->lk_cur].p);
    }
    c->launches++;
  }
  rc = lk_build_pyramid(c, c->lk_cur, nullptr, 0, 3);
  if (rc) return rc;
  const int gb = (B + 127) / 128;
  const int k0 = c->trk_cur, k1 = k0 ^ 1;
  if (c->trk_have_frame) {
    rc = lk_run(c, c->lk_cur ^ 1, c->lk_cur, c->trk_xy[k0].p, c->trk_n[k0].p, cap, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p);
    if (rc) return rc;
    track_collect_kernel<<<B, 1024, 0, c->stream>>>(c->lk_pts_out.p, c->lk_status.p, c->lk_err.p, c->trk_obj[k0].p,
                                                    c->trk_n[k0].p, cap, 30.0f, c->trk_xy[k1].p, c->trk_obj[k1].p,
                                                    c->trk_n[k1].p, c->trk_src.p, p.img.p, p.obj.p, p.npts.p, p.max_pts);
    MVO_CUDA_TRY(c, cudaMemcpyAsync(p.K.p, K, 72, cudaMemcpyHostToDevice, c->stream));
    replicate_k_kernel<<<gb, 128, 0, c->stream>>>(p.K.p, B);
    c->launches += 2;
    rc = pnp_run(c, 100, 8.0, 0.99);   // src/tracker.cpp:309
    if (rc) return rc;
  }
  gather_track_results_kernel<<<gb, 128, 0, c->stream>>>(c->trk_n[k0].p, c->trk_n[k1].p, p.result.p, p.pose_out.p,
                                                         c->trk_have_frame ? 1 : 0, c->d_trk_res.p, B);
  c->launches++;
  MVO_CUDA_TRY(c, cudaMemcpyAsync(c->h_trk_res.p, c->d_trk_res.p, (size_t)B * sizeof(mvo_track_result), cudaMemcpyDeviceToHost,
                                  c->stream));
  cudaEventRecord(tt.end, c->stream);
  tt.used = true;
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  MVO_CUDA_TRY(c, cudaGetLastError());
  memcpy(results, c->h_trk_res.p, (size_t)B * sizeof(mvo_track_result));
  if (c->trk_have_frame) c->trk_cur = k1;
  c->lk_cur ^= 1;
  c->trk_have_frame = true;
  return MVO_OK;
}

int mvo_group_get_tracks(mvo_ctx* c, int stream, float* xy, int32_t* src_idx, int32_t* pnp_inliers, int cap, int* n_tracked,
                         int* n_inliers) {
  if (!c) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  if (stream < 0 || stream >= c->cfg.batch || !c->h_trk_res.p || c->trk_cap <= 0) {
    c->set_error("mvo_group_get_tracks: no tracking frame yet, or stream out of range");
    return MVO_ERR_INVALID;
  }
  const mvo_track_result& r = c->h_trk_res.p[stream];
  if (n_tracked) *n_tracked = r.n_tracked;
  if (n_inliers) *n_inliers = r.n_pnp_inliers;
  if (((xy || src_idx) && r.n_tracked > cap) || (pnp_inliers && r.n_pnp_inliers > cap)) {
    c->set_error("mvo_group_get_tracks: caller buffers too small");
    return MVO_ERR_CAPACITY;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const size_t o = (size_t)stream * c->trk_cap;
  if (xy && r.n_tracked > 0)
    MVO_CUDA_TRY(c, cudaMemcpyAsync(xy, c->trk_xy[c->trk_cur].p + o, (size_t)r.n_tracked * 8, cudaMemcpyDeviceToHost, c->main_stream));
  if (src_idx && r.n_tracked > 0)
    MVO_CUDA_TRY(c, cudaMemcpyAsync(src_idx, c->trk_src.p + o, (size_t)r.n_tracked * 4, cudaMemcpyDeviceToHost, c->main_stream));
  if (pnp_inliers && r.n_pnp_inliers > 0)
    MVO_CUDA_TRY(c, cudaMemcpyAsync(pnp_inliers, c->pnp.inl_idx.p + (size_t)stream * c->pnp.max_pts, (size_t)r.n_pnp_inliers * 4,
                                    cudaMemcpyDeviceToHost, c->main_stream));
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->main_stream));
  return MVO_OK;
}

int mvo_stage_ms(mvo_ctx* c, const char* stage, float* ms) {
  if (!c || !stage || !ms) return MVO_ERR_INVALID;
  for (int s = 0; s < mvo_ctx::kNumStages; ++s) {
    if (strcmp(stage, kStageNames[s]) == 0) {
      if (!c->timers[s].used) {
        *ms = 0.f;
        return MVO_OK;
      }
      MVO_CUDA_TRY(c, cudaEventElapsedTime(ms, c->timers[s].beg, c->timers[s].end));
      return MVO_OK;
    }
  }
  c->set_error("mvo_stage_ms: unknown stage");
  return MVO_ERR_INVALID;
}

int mvo_stage_span_ms(mvo_ctx* c, const char* stage, float* beg_ms, float* end_ms) {
  if (!c || !stage || !beg_ms || !end_ms) return MVO_ERR_INVALID;
  for (int s = 0; s < mvo_ctx::kNumStages; ++s) {
    if (strcmp(stage, kStageNames[s]) == 0) {
      *beg_ms = *end_ms = 0.f;
      if (!c->timers[s].used || !c->timers[ST_TOTAL].used) return MVO_OK;
      MVO_CUDA_TRY(c, cudaEventElapsedTime(beg_ms, c->timers[ST_TOTAL].beg, c->timers[s].beg));
      MVO_CUDA_TRY(c, cudaEventElapsedTime(end_ms, c->timers[ST_TOTAL].beg, c->timers[s].end));
      return MVO_OK;
    }
  }
  c->set_error("mvo_stage_span_ms: unknown stage");
  return MVO_ERR_INVALID;
}

int mvo_graph_stats(mvo_ctx* c, uint64_t stats[3]) {
  if (!c || !stats) return MVO_ERR_INVALID;
  for (int i = 0; i < 3; ++i) stats[i] = c->graph_stats[i];
  return MVO_OK;
}

int mvo_debug_set(mvo_ctx* c, const char* key, int value) {
  if (!c || !key) return MVO_ERR_INVALID;
  if (strcmp(key, "lk_impl") == 0) c->dbg_lk_impl = value;
  else if (strcmp(key, "knn_impl") == 0) c->dbg_knn_impl = value;
  else if (strcmp(key, "h_refine_impl") == 0) c->dbg_h_refine_impl = value;
  else if (strcmp(key, "e5_roots_impl") == 0) c->dbg_e5_roots_impl = value;
  else if (strcmp(key, "pnp_refine_impl") == 0) c->dbg_pnp_refine_impl = value;
  else if (strcmp(key, "pnp_epnp_impl") == 0) c->dbg_pnp_epnp_impl = value;
  else if (strcmp(key, "pnp_rounds") == 0) c->dbg_pnp_rounds = value;
  else if (strcmp(key, "lk_bgr_gray") == 0) c->dbg_lk_bgr_gray = value;
  else if (strcmp(key, "graph") == 0) c->graph_enabled = value;
  else if (strcmp(key, "lk_ctas_per_sm") == 0) c->dbg_lk_ctas_per_sm = value;
  else if (strcmp(key, "cache") == 0) {
    c->cache_enabled = value;
    for (auto& e : c->dcache) e.hash = 0;
    c->lk_hash[0] = c->lk_hash[1] = 0;
  }
  else {
    c->set_error("mvo_debug_set: unknown key");
    return MVO_ERR_INVALID;
  }
  return MVO_OK;
}

// Re-run one stage on the state the last group step left behind: the latest frame's pyramid is lk_pyr[lk_cur ^ 1] and its
// keypoints / descriptors are the prev_* buffers (the step swapped them); the frame before is lk_pyr[lk_cur] / kps / desc.
int mvo_debug_time(mvo_ctx* c, const char* what, int reps, float* ms) {
  if (!c || !what || !ms || reps < 1) return MVO_ERR_INVALID;
  MVO_REQUIRE_IDLE(c);
  if (!c->have_prev || c->geom_w < 0) {
    c->set_error("mvo_debug_time: run at least two group steps first");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const int B = c->cfg.batch, cap = c->geom.kp_cap;
  c->stream = c->main_stream;
  MVO_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  StageTimer& t = c->timers[ST_TOTAL];
  int rc = MVO_OK;
  for (int r = -1; r < reps && rc == MVO_OK; ++r) {   // r == -1: warm-up
    if (r == 0) cudaEventRecord(t.beg, c->stream);
    if (strcmp(what, "lk_track") == 0)
      rc = lk_run(c, c->lk_cur ^ 1, c->lk_cur, c->prev_kp_xy.p, c->prev_kp_count.p, cap, c->lk_pts_out.p, c->lk_status.p,
                  c->lk_err.p);
    else if (strcmp(what, "lk_pyramid") == 0)
      rc = lk_build_pyramid(c, c->lk_cur, nullptr, 0, 3);
    else if (strcmp(what, "knn") == 0)
      rc = knn_run(c, c->prev_desc.p, c->prev_kp_count.p, cap, cap, c->desc.p, c->kp_count.p, cap, cap, 0.7, B);
    else if (strcmp(what, "orb") == 0)
      rc = orb_run_detect(c, true);
    else if (strcmp(what, "orb_dense") == 0)
      rc = orb_run_levels_only(c);
    else if (strcmp(what, "orb_levels") == 0)
      rc = orb_run_levels_fast(c);
    else {
      c->set_error("mvo_debug_time: unknown stage");
      return MVO_ERR_INVALID;
    }
  }
  if (rc) return rc;
  cudaEventRecord(t.end, c->stream);
  MVO_CUDA_TRY(c, cudaEventSynchronize(t.end));
  MVO_CUDA_TRY(c, cudaEventElapsedTime(ms, t.beg, t.end));
  *ms /= (float)reps;
  if (strncmp(what, "orb", 3) == 0) c->have_prev = false;   // the older frame's keypoint buffers were overwritten
  return MVO_OK;
}

}  // extern "C"
