// group.cu -- mvo_group_step: the whole front-end frame for a group of independent camera streams.
//
// One "front-end frame"  per stream: ORB(frame t) -> kNN+ratio(desc t-1, desc t) ->
// LK(keypoints t-1 -> frame t) -> H + F RANSAC -> E RANSAC -> recoverPose -> triangulate.  It chains the same
// kernels the single-call ABI uses (every kernel carries the stream index in its grid), keeps all
// intermediate data on the device and returns one small result record per stream.
// Call sites it stands for: Tracker::update (/root/reference/src/tracker.cpp:274-333) and
// Initializer::try_initializing (src/initializer.cpp:165-313).
#include "context.cuh"
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

// chirality count of the triangulated points (src/initializer.cpp:134-157): z > 0 in both cameras
__global__ void __launch_bounds__(256)
tri_count_kernel(const float* __restrict__ X4, const uint8_t* __restrict__ mask, const double* __restrict__ pose,
                 const int32_t* __restrict__ npts, int max_pts, int32_t* __restrict__ out) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = npts[b];
  const double* R = pose + b * 12;
  const double* t = R + 9;
  const float* X = X4 + (long long)b * 4 * max_pts;
  int c = 0;
  for (int i = tid; i < n; i += 256) {
    if (!mask[(long long)b * max_pts + i]) continue;
    const float w = X[3LL * max_pts + i];
    const float sc = w != 0.f ? 1.f / w : 1.f;        // convertPointsFromHomogeneous
    const float x = X[i] * sc, y = X[1LL * max_pts + i] * sc, z = X[2LL * max_pts + i] * sc;
    if (z <= 0) continue;
    const double z2 = R[6] * (double)x + R[7] * (double)y + R[8] * (double)z + t[2];
    if (z2 > 0) ++c;
  }
  c = warp_sum(c);
  __shared__ int s;
  if (tid == 0) s = 0;
  __syncthreads();
  if ((tid & 31) == 0 && c) atomicAdd(&s, c);
  __syncthreads();
  if (tid == 0) out[b] = s;
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

}  // namespace mvo

using namespace mvo;

#define STAGE_BEG(c, s) cudaEventRecord((c)->timers[s].beg, (c)->stream)
#define STAGE_END(c, s)                               \
  do {                                                \
    cudaEventRecord((c)->timers[s].end, (c)->stream); \
    (c)->timers[s].used = true;                       \
  } while (0)

extern "C" {

static int group_enqueue(mvo_ctx* c, const uint8_t* images, int w, int h, int stride, int images_on_device, const double* K,
                         int slot) {
  if (!c) return MVO_ERR_INVALID;
  if (!images || !K || stride < w) {
    c->set_error("mvo_group_step: bad argument");
    return MVO_ERR_INVALID;
  }
  MVO_CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  const int B = c->cfg.batch;
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
  if (c->lk_w != w || c->lk_h != h || c->lk_max_pts < cap || c->lk_cn != 1) {
    rc = lk_prepare(c, w, h, cap);
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
    if (!images_on_device) {
      GroupSlot& sl = c->slots[slot];
      MVO_CUDA_TRY(c, sl.stage.alloc((size_t)B * h * w));
      cudaStreamWaitEvent(c->copy_stream, sl.ev_free, 0);
      if (stride == w)   // contiguous frames: one flat DMA (a 2-D copy of 1241-byte rows runs far below PCIe speed)
        MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.stage.p, images, (size_t)B * h * w, cudaMemcpyHostToDevice, c->copy_stream));
      else
        MVO_CUDA_TRY(c, cudaMemcpy2DAsync(sl.stage.p, w, images, stride, w, (size_t)B * h, cudaMemcpyHostToDevice,
                                          c->copy_stream));
      cudaEventRecord(sl.ev_up, c->copy_stream);
      cudaStreamWaitEvent(c->main_stream, sl.ev_up, 0);
      src = sl.stage.p;
      spitch = w;
      sstride = (long long)h * w;
    }
    const LevelGeom& l0 = g.lv[0];
    dim3 grid((w + 16 * 256 - 1) / (16 * 256), h, B);
    int lk_pitch = 0;
    long long lk_fs = 0;
    uint8_t* lk0 = lk_level0(c, c->lk_cur, &lk_pitch, &lk_fs);
    unpack_frames_kernel<<<grid, 256, 0, c->stream>>>(src, spitch, sstride, c->pyr.p + l0.off, l0.pitch, g.frame_stride, lk0,
                                                      lk_pitch, lk_fs, w, h);
    c->launches++;
    if (!images_on_device) cudaEventRecord(c->slots[slot].ev_free, c->main_stream);
  }
  // the keypoint / descriptor buffers ORB is about to fill were the previous step's "prev" set: its kNN must be done
  cudaStreamWaitEvent(c->main_stream, c->ev_join[2], 0);
  rc = orb_run_detect(c, true);
  if (rc) return rc;
  {
    GroupSlot& sl = c->slots[slot];
    MVO_CUDA_TRY(c, sl.h_res.alloc(B));
    MVO_CUDA_TRY(c, sl.h_flags.alloc(B));
    MVO_CUDA_TRY(c, cudaMemcpyAsync(sl.h_flags.p, c->flags.p, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  STAGE_END(c, ST_ORB);
  cudaEventRecord(c->ev_fork[0], c->main_stream);

  const int cur = c->lk_cur, prev = cur ^ 1;
  if (c->have_prev) {
    // ---- kNN + ratio: query = previous descriptors, train = new descriptors (src/tracker.cpp:190-191) ----
    Fork f(c, c->aux_stream[2], 0, c->ev_fork[0]);
    cudaStreamWaitEvent(c->stream, c->ev_tail, 0);   // the previous step's gather still reads the match counters
    STAGE_BEG(c, ST_KNN);
    rc = knn_run(c, c->prev_desc.p, c->prev_kp_count.p, cap, cap, c->desc.p, c->kp_count.p, cap, cap, 0.7, B);
    if (rc) return rc;
    STAGE_END(c, ST_KNN);
    cudaEventRecord(c->ev_join[2], c->stream);
  }
  // ---- LK pyramid of the new frame (level 0 = ORB level 0, device to device) ----
  STAGE_BEG(c, ST_LK);
  rc = lk_build_pyramid(c, cur, nullptr, 0, 3);   // level 0 was written by the unpack kernel
  if (rc) return rc;
  if (c->have_prev) {
    rc = lk_run(c, prev, cur, c->prev_kp_xy.p, c->prev_kp_count.p, cap, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p);
    if (rc) return rc;
    // the correspondence buffers feed the model searches of the previous step until its tail is done; from here on
    // this step's ORB + LK have overlapped them (software pipelining across the two steps in flight)
    cudaStreamWaitEvent(c->main_stream, c->ev_tail, 0);
    lk_collect_kernel<<<B, 1024, 0, c->stream>>>(c->prev_kp_xy.p, c->lk_pts_out.p, c->lk_status.p, c->lk_err.p,
                                                 c->prev_kp_count.p, cap, 30.0f, r.p1.p, r.p2.p, r.npts.p);
    c->launches++;
  }
  STAGE_END(c, ST_LK);

  if (c->have_prev) {
    // counts kept per stage
    MVO_CUDA_TRY(c, c->knn_counts.alloc((size_t)4 * B));
    int32_t* res_h = c->knn_counts.p;
    int32_t* res_f = res_h + B;
    int32_t* ntri = res_f + B;
    const int gb = (B + 127) / 128;
    MVO_CUDA_TRY(c, cudaMemcpyAsync(r.K.p, K, 72, cudaMemcpyHostToDevice, c->stream));
    fill_params_kernel<<<gb, 128, 0, c->stream>>>(r.lane[0].thr2.p, 1.0f, r.K.p, B);   // + replicate K to every stream
    c->launches++;
    cudaEventRecord(c->ev_fork[1], c->main_stream);
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
      STAGE_BEG(c, ST_POSE);
      rc = pose_recover(c, true);
      if (rc) return rc;
      STAGE_END(c, ST_POSE);
      STAGE_BEG(c, ST_TRI);
      make_proj_kernel<<<gb, 128, 0, c->stream>>>(r.K.p, r.pose.p, r.proj.p, B);
      c->launches++;
      rc = pose_triangulate(c);
      if (rc) return rc;
      tri_count_kernel<<<B, 256, 0, c->stream>>>(r.X4.p, r.ln().mask.p, r.pose.p, r.npts.p, r.max_pts, ntri);
      c->launches++;
      STAGE_END(c, ST_TRI);
      cudaEventRecord(c->ev_join[1], c->stream);
    }
    {
      // ---- H RANSAC (thr 1.0), lane 0, then the join of all searches and the result gather: tail stream ----
      Fork f(c, c->aux_stream[3], 0, c->ev_fork[1]);
      STAGE_BEG(c, ST_H);
      rc = ransac_find(c, MVO_MODEL_H, 0.995);
      if (rc) return rc;
      copy_i32_strided_kernel<<<gb, 128, 0, c->stream>>>(r.ln().result.p, 8, res_h, B);
      c->launches++;
      STAGE_END(c, ST_H);
      for (int k = 0; k < 3; ++k) cudaStreamWaitEvent(c->stream, c->ev_join[k], 0);
      gather_results_kernel<<<gb, 128, 0, c->stream>>>(c->kp_count.p, c->knn_nmatch.p, r.npts.p, res_h, res_f,
                                                       r.lane[2].result.p, ntri, r.pose.p, 1, c->d_results.p, B);
      c->launches++;
      MVO_CUDA_TRY(c, cudaMemcpyAsync(c->slots[slot].h_res.p, c->d_results.p, (size_t)B * sizeof(mvo_frame_result),
                                      cudaMemcpyDeviceToHost, c->stream));
      STAGE_END(c, ST_TOTAL);
      cudaEventRecord(c->ev_tail, c->stream);
      cudaEventRecord(c->slots[slot].ev_done, c->stream);
    }
  } else {
    Fork f(c, c->aux_stream[3], 0, c->ev_fork[0]);
    gather_results_kernel<<<(B + 127) / 128, 128, 0, c->stream>>>(c->kp_count.p, nullptr, nullptr, nullptr, nullptr,
                                                                 nullptr, nullptr, nullptr, 0, c->d_results.p, B);
    c->launches++;
    MVO_CUDA_TRY(c, cudaMemcpyAsync(c->slots[slot].h_res.p, c->d_results.p, (size_t)B * sizeof(mvo_frame_result),
                                    cudaMemcpyDeviceToHost, c->stream));
    STAGE_END(c, ST_TOTAL);
    cudaEventRecord(c->ev_tail, c->stream);
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
  int flags0 = 0;
  for (int b = 0; b < B; ++b) flags0 |= sl.h_flags.p[b];
  if (flags0 & 1) {
    c->set_error("FAST candidate list overflow");
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
  int rc = group_enqueue(c, images, w, h, stride, images_on_device, K, 0);
  if (rc) return rc;
  return group_finish(c, 0, results);
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

int mvo_debug_set(mvo_ctx* c, const char* key, int value) {
  if (!c || !key) return MVO_ERR_INVALID;
  if (strcmp(key, "lk_impl") == 0) c->dbg_lk_impl = value;
  else if (strcmp(key, "knn_impl") == 0) c->dbg_knn_impl = value;
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
