// context.cuh -- the stream-group context behind the C ABI (one per group of camera streams).
#pragma once
#include "common.cuh"

struct StageTimer {
  cudaEvent_t beg = nullptr, end = nullptr;
  bool used = false;
};

// Scratch of one model search.  H, F and E searches of a group step run concurrently on three CUDA streams, each in
// its own lane; the single-call ABI uses lane 0.
struct RansacLane {
  mvo::DevBuf<uint8_t> mask;          // batch * max_pts
  mvo::DevBuf<int32_t> inl_idx;       // batch * max_pts
  mvo::DevBuf<int32_t> state;         // batch * 8 : sampler state
  mvo::DevBuf<float> thr2;            // batch : squared threshold as float
  mvo::DevBuf<uint16_t> att_next;     // batch * 32768 : end offset of the getSubset attempt starting at each draw
  mvo::DevBuf<uint8_t> att_ok;        // batch * 32768 : did that attempt pass checkSubset
  mvo::DevBuf<int32_t> subsets;       // batch * cap_iters * 8
  mvo::DevBuf<double> models;         // batch * cap_iters * 10 * 9
  mvo::DevBuf<double> e5_scratch;     // batch * cap_iters * 86 : 5-point solver intermediates
  mvo::DevBuf<int32_t> nmodels;       // batch * cap_iters
  mvo::DevBuf<int32_t> counts;        // batch * cap_iters * 10
  mvo::DevBuf<double> best_model;     // batch * 9
  mvo::DevBuf<int32_t> result;        // batch * 8
  int max_pts = 0, cap_iters = 0;
  bool ready = false;
  void release() {
    mask.release(); inl_idx.release(); state.release(); thr2.release(); att_next.release(); att_ok.release();
    subsets.release(); models.release(); e5_scratch.release(); nmodels.release(); counts.release();
    best_model.release(); result.release();
    max_pts = cap_iters = 0; ready = false;
  }
};

struct RansacBufs {
  static constexpr int kLanes = 3;
  mvo::DevBuf<uint32_t> rng;          // OpenCV MWC stream for the context seed
  int rng_len = 0;
  bool rng_ready = false;
  int max_pts = 0, cap_iters = 0;
  mvo::DevBuf<float2> p1, p2;         // batch * max_pts pixel correspondences
  mvo::DevBuf<double2> q1, q2;        // K-normalised (essential matrix)
  mvo::DevBuf<int32_t> npts;          // batch
  mvo::DevBuf<double> K;              // batch * 9
  RansacLane lane[kLanes];
  int cur = 0;                        // lane the host code is currently issuing work for
  RansacLane& ln() { return lane[cur]; }
  // pose
  mvo::DevBuf<double> cands;          // batch * 4 * 12 : (R | t) candidates
  mvo::DevBuf<uint8_t> cand_mask;     // batch * 4 * max_pts
  mvo::DevBuf<int32_t> cand_good;     // batch * 4
  mvo::DevBuf<double> pose;           // batch * 12 : R (9) + t (3)
  mvo::DevBuf<double> proj;           // batch * 24 : P0, P1
  mvo::DevBuf<float> X4;              // batch * 4 * max_pts
  void release() {
    rng.release(); p1.release(); p2.release(); q1.release(); q2.release(); npts.release(); K.release();
    for (auto& l : lane) l.release();
    cands.release(); cand_mask.release(); cand_good.release(); pose.release(); proj.release(); X4.release();
  }
};

// solvePnPRansac (pnp.cu)
struct PnpBufs {
  int max_pts = 0, cap_iters = 0;
  mvo::DevBuf<float> obj;             // batch * max_pts * 3 object points
  mvo::DevBuf<float2> img;            // batch * max_pts image points (pixels)
  mvo::DevBuf<double2> xn;            // K-normalised image points (rounded to float, as undistortPoints returns them)
  mvo::DevBuf<uint8_t> mask;          // batch * max_pts
  mvo::DevBuf<int32_t> inl_idx;       // batch * max_pts : ordered inlier indices of the winner
  mvo::DevBuf<int32_t> npts;          // batch
  mvo::DevBuf<double> K;              // batch * 9
  mvo::DevBuf<int32_t> subsets;       // batch * iters * 5
  mvo::DevBuf<double> models;         // batch * iters * 12 : R (9) | t (3)
  mvo::DevBuf<int32_t> ok, counts;    // batch * iters
  mvo::DevBuf<int32_t> bound;         // batch : hypotheses the adaptive loop can still reach after the first round
  mvo::DevBuf<double> best_model;     // batch * 12
  mvo::DevBuf<int32_t> result;        // batch * 8 : inliers, iterations run, winning iteration (-1: none)
  mvo::DevBuf<double> pose_out;       // batch * 8 : rvec, tvec, planar flag
  void release() {
    obj.release(); img.release(); xn.release(); mask.release(); inl_idx.release(); npts.release(); K.release();
    subsets.release(); models.release(); ok.release(); counts.release(); best_model.release(); result.release();
    pose_out.release();
    max_pts = cap_iters = 0;
  }
};

// one in-flight group step (mvo_group_submit / mvo_group_collect keep up to two)
struct GroupSlot {
  mvo::DevBuf<uint8_t> stage;                 // batch * h * w staged host frames
  mvo::PinBuf<mvo_frame_result> h_res;        // pinned result records
  mvo::PinBuf<int32_t> h_flags;
  mvo::PinBuf<int32_t> h_occ;           // batch * 2: occupied / total cells of the keypoint-distribution grid
  cudaEvent_t ev_up = nullptr, ev_free = nullptr, ev_done = nullptr;
  mvo::PinBuf<uint8_t> h_out;                 // full per-stream outputs (mvo_group_configure), sections of OutLayout
  int had_prev = 0;                           // the step enqueued in this slot had a previous frame
};

// sections of a slot's pinned output block (byte offsets; every section holds batch * cap entries)
struct OutLayout {
  size_t kps = 0, desc = 0, matches = 0, lk_xy = 0, lk_status = 0, lk_err = 0, mask[4] = {0, 0, 0, 0}, models = 0, x4 = 0,
         cloud = 0, prev_count = 0, total = 0;
  int cap = 0, batch = 0;
  uint32_t mask_bits = 0;
};

// Device-resident copy of a descriptor block (N x 32 bytes) keyed by its content: FeatureProcessor::find_matches is
// always called on blocks that were just produced by detect_and_compute or gathered again from the same observations
// (src/frame.cpp:50-64, src/keyframe.cpp:78-92, src/tracker.cpp:190-201) -- they do not have to cross PCIe again.
struct DescCacheEntry {
  uint64_t hash = 0, stamp = 0;
  int n = 0;
  mvo::DevBuf<uint8_t> buf;
};

struct mvo_ctx {
  mvo_config cfg{};
  cudaStream_t stream = nullptr;     // the stream host code currently issues on (main stream, or an aux stream inside a fork)
  cudaStream_t main_stream = nullptr;
  cudaStream_t aux_stream[4] = {nullptr, nullptr, nullptr, nullptr};   // group step: F, E(+pose), kNN, H + tail
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_tail = nullptr;       // all model searches + the result gather of the latest enqueued step are done
  cudaStream_t lk_stream = nullptr;    // group step: LK pyramid + tracker + correspondence list (runs beside ORB)
  cudaEvent_t ev_unpack = nullptr;     // main stream: level 0 of the new frame's ORB and LK pyramids is written
  cudaEvent_t ev_lk_done = nullptr;    // LK stream: the tracker of the latest enqueued step has read both pyramids
  bool own_stream = false;
  static constexpr int kSlots = 2;
  cudaStream_t copy_stream = nullptr;  // H2D of staged frames
  cudaStream_t out_stream = nullptr;   // D2H of the full per-stream outputs
  cudaEvent_t ev_o_orb = nullptr, ev_o_lk = nullptr;   // main stream: ORB outputs / LK outputs of the step are final
  cudaEvent_t ev_out_orb = nullptr, ev_out_lk = nullptr;   // out stream: their D2H copies are done
  // ---------------- group product surface (mvo_group_configure) ----------------
  int grp_cn = 1;                      // channels of the frames given to the group entry points
  uint32_t out_mask = 0;               // MVO_OUT_*
  OutLayout out_layout;
  int out_slot = -1;                   // slot whose outputs mvo_group_outputs hands out (last finished step)
  mvo::DevBuf<uint8_t> e_mask_keep;    // findEssentialMat's mask (recoverPose overwrites it in place)
  // ---------------- group tracking frame (mvo_group_track) ----------------
  mvo::DevBuf<float2> trk_xy[2];       // batch * trk_cap : observations of the latest / the new frame
  mvo::DevBuf<float> trk_obj[2];       // batch * trk_cap * 3 : their landmark coordinates
  mvo::DevBuf<int32_t> trk_n[2];       // batch
  mvo::DevBuf<int32_t> trk_src;        // batch * trk_cap : index of each new observation in the previous list
  mvo::DevBuf<mvo_track_result> d_trk_res;
  mvo::PinBuf<mvo_track_result> h_trk_res;
  int trk_cur = 0, trk_cap = 0;
  bool trk_have_frame = false;         // lk_pyr[lk_cur ^ 1] holds the latest frame of every stream
  GroupSlot slots[kSlots];
  int q_head = 0, q_count = 0;         // ring of submitted, not yet collected steps
  std::string err;
  uint64_t launches = 0;

  void set_error(const std::string& s) { err = s; }

  // ---------------- ORB ----------------
  mvo::OrbGeom geom{};
  int geom_w = -1, geom_h = -1;
  mvo::DevBuf<uint8_t> img_in;        // batch * h * stride staging for BGR input
  mvo::DevBuf<uint8_t> pyr, blur;     // batch * frame_stride
  mvo::DevBuf<uint32_t> xtab, ytab;   // INTER_LINEAR_EXACT coefficient tables (offset | c1 << 16)
  mvo::DevBuf<uint32_t> cand_xy;      // batch * cand_total : x | y << 16
  mvo::DevBuf<int32_t> cand_score;    // batch * cand_total : FAST score
  mvo::DevBuf<int32_t> cand_count;    // batch * 8
  mvo::DevBuf<uint32_t> cand_sel;     // batch * cand_total : candidates that survive retainBest(2 n_l)
  mvo::DevBuf<int32_t> sel_count;     // batch * 8
  mvo::DevBuf<uint32_t> hist;         // batch * 8 * 256 : FAST score histogram
  mvo::DevBuf<unsigned long long> c2_key, c2_key_sorted;  // Harris stage: sort keys
  mvo::DevBuf<float2> c2_ra, c2_ra_sorted;                // (response, angle)
  mvo::DevBuf<int32_t> c2_count;      // batch * 8
  mvo::DevBuf<mvo_keypoint> kps;      // batch * kp_cap
  mvo::DevBuf<float2> kp_xy;          // batch * kp_cap : keypoint positions (LK input of the next frame)
  mvo::DevBuf<uint8_t> desc;          // batch * kp_cap * 32
  mvo::DevBuf<uint8_t> kp_valid;      // batch * kp_cap (orb_compute hook)
  mvo::DevBuf<int32_t> kp_count;      // batch
  mvo::DevBuf<int32_t> flags;         // batch : bit0 = candidate overflow, bit1 = keypoint overflow
  mvo::PinBuf<uint8_t> h_stage;       // pinned staging for H2D / D2H
  size_t h_stage_bytes = 0;

  // previous-frame state for the group step (device resident)
  mvo::DevBuf<mvo_keypoint> prev_kps;
  mvo::DevBuf<float2> prev_kp_xy;
  mvo::DevBuf<mvo_frame_result> d_results;
  int lk_cur = 0;                     // which LK pyramid holds the current frame
  mvo::DevBuf<uint8_t> prev_desc;
  mvo::DevBuf<int32_t> prev_kp_count;
  bool have_prev = false;

  // ---------------- kNN ----------------
  mvo::DevBuf<uint8_t> knn_q, knn_t;          // staging for the host-pointer API
  mvo::DevBuf<uint32_t> knn_best;             // batch * maxq * 2 packed keys (dist << 22 | idx)
  mvo::DevBuf<mvo_dmatch> knn_matches;        // batch * maxq
  mvo::DevBuf<int32_t> knn_nmatch;            // batch
  mvo::DevBuf<int32_t> knn_counts;            // 2 ints (nq, nt) for the host-pointer API

  // ---------------- LK ----------------
  mvo::DevBuf<uint8_t> lk_pyr[2];             // ping-pong pyramids (levels 0..3) of prev / next frame
  mvo::DevBuf<float2> lk_pts_in, lk_pts_out;  // batch * lk_max_pts
  mvo::DevBuf<uint8_t> lk_status;
  mvo::DevBuf<float> lk_err;
  mvo::DevBuf<int32_t> lk_npts;               // batch
  int lk_w = 0, lk_h = 0, lk_max_pts = 0, lk_cn = 1;   // lk_cn: image planes per stream (1 gray, 3 BGR)
  // TMA descriptors (CUtensorMap, 128 bytes each) of every level of the two LK pyramid buffers: {x, y, stream} byte
  // tensors with 48 x 24 / 48 x 29 boxes (lk.cu: lk_track2_kernel stages its tiles with cp.async.bulk.tensor)
  alignas(64) unsigned char lk_tmaps[2][2][mvo::kLkLevels][128];   // [buffer][template | next-image box][level]
  bool lk_tmaps_ok = false;

  // ---------------- RANSAC / pose ----------------
  RansacBufs rs;
  PnpBufs pnp;
  int last_ransac_iters = 0;

  // ---------------- profiling / parity knobs (mvo_debug_set) ----------------
  int dbg_lk_impl = 2;     // 1: first-generation lk_track_kernel (in-tree cross-check), 2: lk_track2_kernel (persistent
                           // warps for large point sets), 3: lk_track2_kernel with one point per warp always
  mvo::DevBuf<int32_t> lk_work;   // work counters of the persistent LK kernels (two: BGR8 runs a team and a gray launch)
  mvo::DevBuf<int32_t> lk_colour[2];   // per pyramid buffer and stream: nonzero = the BGR8 frame has pixels with differing channels
  int dbg_lk_ctas_per_sm = 5;     // resident CTAs per SM of the persistent LK kernel (5 = all the shared memory of an SM)
  int cand_scale = 1;      // FAST candidate list capacity in units of (level pixels / 16); doubled after an overflow
  int occupancy_div = 50;  // keypoint-distribution grid cell size (config/params.yaml: initializer.occupancy_grid_div)
  mvo::DevBuf<int32_t> occ;            // batch * 2: occupied / total cells, written by orb_finalize_kernel
  int occ_single[2] = {0, 0};          // ... of the last single-call detect
  bool occ_from_group = false;         // mvo_orb_occupancy answers from the last collected group step
  mvo::DevBuf<float> cloud;            // group step: batch * cap * 3 packed ROS-axis points (MVO_OUT_CLOUD)
  mvo::DevBuf<uint8_t> pack_tmp;       // mvo_pack_pointcloud staging
  // ---- CUDA-graph form of the synchronous group step (small groups: the host's ~85 launches per step, not the GPU,
  // bound a single stream).  One executable graph per buffer parity (the keypoint / descriptor / pyramid buffers swap
  // every frame), re-captured when the geometry, the output mask, a knob or any allocation changes. ----
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    unsigned long long epoch = 0, key = 0;
    int launches = 0;
  };
  StepGraph step_graph[2];
  bool capturing = false;             // group_enqueue is being recorded into a graph: no waits on events of earlier steps
  int graph_enabled = 1;              // mvo_debug_set("graph", 0) switches the graph form off (per-stage timers need that)
  int steps_since_change = 0;         // plain steps since the last (re)allocation: the graph is captured after two
  cudaEvent_t ev_graph_done = nullptr;
  uint64_t graph_stats[3] = {0, 0, 0};   // captures, replays, capture fall-backs
  // ---- device-resident caches of the synchronous single-call path ----
  static constexpr int kDescCache = 4;
  DescCacheEntry dcache[kDescCache];
  uint64_t dcache_clock = 0;
  uint64_t lk_hash[2] = {0, 0};        // content hash of the image each LK pyramid slot was built from (0: none)
  int lk_slot_prev = 0, lk_slot_next = 1;   // slots the last mvo_lk_track call used (mvo_lk_get_level's which = 0 / 1)
  uint64_t cache_stats[4] = {0, 0, 0, 0};   // descriptor hits / misses, pyramid hits / misses
  int cache_enabled = 1;               // mvo_debug_set("cache", 0) switches both caches off
  int dbg_knn_impl = 0;    // kNN kernel choice (0 = default)
  int dbg_e5_roots_impl = 2;   // 1: derivative-level bracketing only (cross-check), 2: Ehrlich-Aberth iteration first, bracketing where it is not trusted
  int dbg_lk_bgr_gray = 1;     // 0: BGR8 streams with identical planes go through the three-warp teams as well (cross-check)
  int dbg_pnp_rounds = 0;      // 1: batched solvePnPRansac evaluates all hypotheses in one round (cross-check)
  int dbg_pnp_epnp_impl = 1;   // 12 x 12 Jacobi of pnp_epnp_kernel: 0 round 1 (cross-check), 1 one element pair per lane + short scalar chain
  int dbg_pnp_refine_impl = 2; // 1: first-generation initial pose of pnp_refine_kernel (cross-check), 2: block sums + warp LU
  int dbg_h_refine_impl = 2;   // 1: first-generation h_refine_kernel (cross-check), 2: h_refine2_kernel

  // ---------------- stage timing ----------------
  static constexpr int kNumStages = 10;
  StageTimer timers[kNumStages];
};

// single-call entry points share streams and scratch with the group pipeline: refuse to run while submitted steps are
// still in flight
#define MVO_REQUIRE_IDLE(c)                                                                          \
  do {                                                                                               \
    if ((c)->q_count != 0) {                                                                         \
      (c)->set_error("submitted group steps are still in flight (mvo_group_collect them first)");   \
      return MVO_ERR_INVALID;                                                                        \
    }                                                                                                \
  } while (0)

namespace mvo {

// host-side module entry points (defined in the respective .cu files)
int orb_prepare(mvo_ctx* c, int w, int h);
void orb_set_grid(mvo_ctx* c, int w, int h);
int orb_upload(mvo_ctx* c, const uint8_t* img, int w, int h, int stride, int channels, int on_device);
int orb_run_detect(mvo_ctx* c, bool want_desc);
int orb_run_levels_fast(mvo_ctx* c);
int orb_run_levels_only(mvo_ctx* c);
int orb_run_brief_given(mvo_ctx* c, int n);

int knn_prepare(mvo_ctx* c, int maxq);
// batched device-resident kNN: q/t are [batch][stride_rows][32]; counts on device
int knn_run(mvo_ctx* c, const uint8_t* q_dev, const int32_t* nq_dev, int q_stride_rows, int max_nq,
            const uint8_t* t_dev, const int32_t* nt_dev, int t_stride_rows, int max_nt, double ratio,
            int batch);

int upload_gray_rows(mvo_ctx* c, const uint8_t* host, int w, int h, int stride, int batch, uint8_t* dst, int dpitch,
                     long long dst_frame_stride);
int lk_prepare(mvo_ctx* c, int w, int h, int max_pts, int cn = 1);
int lk_build_pyramid(mvo_ctx* c, int which, const uint8_t* img, int stride, int on_device);
// level 0 of LK pyramid buffer `which` (gray): base pointer of stream 0, row pitch, distance between streams
uint8_t* lk_level0(mvo_ctx* c, int which, int* pitch, long long* frame_stride);
int lk_run(mvo_ctx* c, int prev_which, int next_which, const float2* pts_dev, const int32_t* npts_dev, int max_pts,
           float2* out_dev, uint8_t* status_dev, float* err_dev);

int ransac_prepare(mvo_ctx* c, int max_pts, int cap_iters, int lanes = 1);
int ransac_find(mvo_ctx* c, int model, double conf);
int ransac_normalize(mvo_ctx* c);
int ransac_sweep(mvo_ctx* c, int model, int m);
int pose_prepare(mvo_ctx* c);
int pnp_prepare(mvo_ctx* c, int max_pts, int iters);
int pnp_run(mvo_ctx* c, int iters, double reproj_err, double conf);
int pose_recover(mvo_ctx* c, bool use_mask);      // E in rs.best_model, points in rs.q1/q2, mask in rs.mask
int pose_triangulate(mvo_ctx* c);                 // P0/P1 in rs.proj, points in rs.p1/p2 -> rs.X4

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace mvo
