/*
 * monovo_b200.h -- C ABI of libmonovo_b200.so: the B200-native (sm_100a) implementation of the
 * ros2_mono_vo per-frame front-end hot path.
 *
 * This header is the drop-in boundary.  The reference reaches the hot path only through
 * OpenCV calls; each entry point below replaces exactly one of those call sites.  A maintainer of the
 * reference keeps FeatureProcessor / Initializer / Tracker unchanged and binds these functions inside
 * src/feature_processor.cpp and at the cv:: call sites of src/initializer.cpp / src/tracker.cpp
 * (see INTEGRATION.md for the binding code).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/torch/OpenCV types cross this boundary.
 *   - every function returns MVO_OK (0) or a negative mvo_status; nothing throws.
 *   - pointers are HOST memory unless the name ends in _dev; calls are synchronous with respect
 *     to the context's CUDA stream when they return (outputs are valid on return).
 *   - there is NO CPU fallback: without a CUDA device mvo_create fails with MVO_ERR_CUDA.
 *   - one context == one "stream group": `batch` independent camera streams processed in lock step
 *     on one GPU (batch = 1 is the single camera of the reference node).
 */
#ifndef MONOVO_B200_H_
#define MONOVO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVO_API __attribute__((visibility("default")))

typedef enum mvo_status {
  MVO_OK = 0,
  MVO_ERR_INVALID = -1,     /* bad argument (null pointer, size out of range, ...)         */
  MVO_ERR_CUDA = -2,        /* CUDA runtime error, see mvo_last_error()                      */
  MVO_ERR_CAPACITY = -3,    /* an internal or caller-provided capacity was exceeded          */
  MVO_ERR_UNSUPPORTED = -4, /* valid request that this build does not implement              */
  MVO_ERR_DEGENERATE = -5   /* input too small / degenerate for the requested model          */
} mvo_status;

typedef struct mvo_ctx mvo_ctx;

/* Layout-compatible with the fields of cv::KeyPoint that the reference reads (28 bytes). */
typedef struct mvo_keypoint {
  float x, y;      /* pt, full-resolution pixel coordinates (level coords * scale)  */
  float size;      /* 31 * scale                                                     */
  float angle;     /* degrees in [0,360)                                             */
  float response;  /* Harris response                                                */
  int32_t octave;  /* pyramid level                                                  */
  int32_t class_id;/* always -1                                                      */
} mvo_keypoint;

/* Layout-compatible with cv::DMatch (16 bytes). */
typedef struct mvo_dmatch {
  int32_t query_idx;
  int32_t train_idx;
  int32_t img_idx;   /* always 0 */
  float distance;    /* Hamming distance as float */
} mvo_dmatch;

typedef struct mvo_config {
  int32_t device;        /* CUDA device ordinal                                              */
  int32_t max_width;     /* largest image width the context will see                         */
  int32_t max_height;
  int32_t nfeatures;     /* cv::ORB::create(nfeatures); reference: src/mono_vo.cpp:16 (1000) */
  int32_t batch;         /* camera streams processed in lock step (>=1)                      */
  int32_t max_points;    /* capacity for LK / RANSAC point sets (0 -> 2*nfeatures)           */
  uint64_t ransac_seed;  /* 0 -> 0xFFFFFFFFFFFFFFFF == OpenCV's RNG seed inside find*()       */
  void* cuda_stream;     /* cudaStream_t to run on; NULL -> the context creates its own      */
} mvo_config;

/* ------------------------------------------------------------------------------------------------
 * context
 * replaces: FeatureProcessor::FeatureProcessor (src/feature_processor.cpp:5-10), which builds
 * cv::ORB::create(num_features) + cv::BFMatcher(NORM_HAMMING).
 */
MVO_API int mvo_create(mvo_ctx** out, const mvo_config* cfg);
MVO_API void mvo_destroy(mvo_ctx* ctx);
MVO_API const char* mvo_last_error(const mvo_ctx* ctx);   /* ctx may be NULL: last create error */
MVO_API const char* mvo_version(void);
MVO_API void* mvo_cuda_stream(mvo_ctx* ctx);               /* the cudaStream_t the ctx launches on */
MVO_API int mvo_batch(const mvo_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
MVO_API uint64_t mvo_launch_count(const mvo_ctx* ctx);

/* ------------------------------------------------------------------------------------------------
 * ORB
 * replaces: detector_->detectAndCompute(image, cv::noArray(), keypoints, descriptors)
 *           src/feature_processor.cpp:19-23  (callers src/frame.cpp:12)
 * img: h x w x channels u8, row stride `stride` bytes; channels 1 (gray) or 3 (BGR, as the node feeds).
 * kps / desc: caller buffers of capacity `cap` keypoints / cap*32 bytes.  *n_out = number written.
 * Keypoints come in the canonical order (octave, response desc, y, x); the SET equals cv::ORB's.
 * desc may be NULL (== FeatureProcessor::detect, src/feature_processor.cpp:12-17).
 */
MVO_API int mvo_orb_detect_and_compute(mvo_ctx* ctx, const uint8_t* img, int w, int h, int stride,
                                       int channels, mvo_keypoint* kps, uint8_t* desc, int cap,
                                       int* n_out);

/* Parity hook == cv::ORB::compute(image, keypoints, descriptors): descriptors for GIVEN keypoints
 * (pt, angle, octave honoured).  valid[i] = 0 for keypoints closer than 31 px (level coords) to the
 * level border (cv2 silently drops those); their descriptor rows are zeroed. */
MVO_API int mvo_orb_compute(mvo_ctx* ctx, const uint8_t* img, int w, int h, int stride, int channels,
                            const mvo_keypoint* kps_in, int n, uint8_t* desc, uint8_t* valid);

/* Debug / parity taps of the last ORB call on stream 0 of the group. */
MVO_API int mvo_orb_num_levels(void);
MVO_API int mvo_orb_level_size(mvo_ctx* ctx, int level, int* w, int* h);
/* copy pyramid level (blurred = 0: INTER_LINEAR_EXACT level; 1: the 7x7 sigma-2 blurred level) */
MVO_API int mvo_orb_get_level(mvo_ctx* ctx, int level, int blurred, uint8_t* out, int out_stride);
/* FAST+NMS candidates of a level after the 31-px edge filter: packed (x | y<<16), FAST score */
MVO_API int mvo_orb_get_fast(mvo_ctx* ctx, int level, uint32_t* xy, int32_t* score, int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------
 * Hamming kNN (k=2) + Lowe ratio
 * replaces: matcher_.knnMatch(d1, d2, knn, 2) + the ratio loop, src/feature_processor.cpp:25-40
 * q: nq x 32 u8 (query == descriptors1), t: nt x 32 u8 (train == descriptors2).
 * out: capacity nq; accepted matches in query order; *n_out = count.
 */
MVO_API int mvo_knn_ratio(mvo_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt,
                          double ratio, mvo_dmatch* out, int* n_out);
/* measured ceiling of the integer population-count pipe on this GPU (32-bit popc per second): the roofline
 * denominator of the matching kernel, which is popc-bound, not HBM-bound (bench.py) */
MVO_API int mvo_measure_popc_peak(mvo_ctx* ctx, double* popc_per_s);
/* raw top-2 (parity hook): idx/dist are nq x 2, -1 where the train set has fewer than 2 rows */
MVO_API int mvo_knn2(mvo_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt,
                     int32_t* idx, int32_t* dist);

/* ------------------------------------------------------------------------------------------------
 * pyramidal Lucas-Kanade
 * replaces: cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err) with all defaults
 *           (21x21 window, maxLevel 3, 30 iters / eps 0.01, minEigThreshold 1e-4)
 *           src/tracker.cpp:68-69
 */
MVO_API int mvo_lk_track(mvo_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h,
                         int stride, int channels, const float* prev_xy, int n, float* next_xy,
                         uint8_t* status, float* err);

/* Parity tap of the last mvo_lk_track call: pyramid level `level` (0 = the image) of the previous (which = 0) or next
 * (which = 1) image, channel plane `plane` (0 for gray; 0..2 = B, G, R for BGR input).  out may be NULL (size query). */
MVO_API int mvo_lk_get_level(mvo_ctx* ctx, int which, int level, int plane, uint8_t* out, int out_stride, int* w, int* h);

/* ------------------------------------------------------------------------------------------------
 * two-view geometry (deterministic parallel RANSAC; OpenCV's RNG stream and adaptive stop replayed)
 * p1/p2: n x 2 f32 interleaved (std::vector<cv::Point2f>).  mask: n bytes (0/1), may be NULL.
 */
/* replaces cv::findHomography(p1, p2, cv::RANSAC, thr, mask): src/initializer.cpp:82, src/tracker.cpp:243 */
MVO_API int mvo_find_homography(mvo_ctx* ctx, const float* p1, const float* p2, int n, double thr,
                                double H[9], uint8_t* mask, int* n_inliers);
/* replaces cv::findFundamentalMat(p1, p2, cv::FM_RANSAC, thr, conf, mask): src/initializer.cpp:87, src/tracker.cpp:248 */
MVO_API int mvo_find_fundamental(mvo_ctx* ctx, const float* p1, const float* p2, int n, double thr,
                                 double conf, double F[9], uint8_t* mask, int* n_inliers);
/* replaces cv::findEssentialMat(p1, p2, K, cv::RANSAC, conf, thr, mask): src/initializer.cpp:228-229 */
MVO_API int mvo_find_essential(mvo_ctx* ctx, const float* p1, const float* p2, int n,
                               const double K[9], double conf, double thr, double E[9],
                               uint8_t* mask, int* n_inliers);
/* replaces cv::recoverPose(E, p1, p2, K, R, t, mask): src/initializer.cpp:236.  mask_io may be NULL. */
MVO_API int mvo_recover_pose(mvo_ctx* ctx, const double E[9], const float* p1, const float* p2, int n,
                             const double K[9], double R[9], double t[3], uint8_t* mask_io, int* n_good);
/* replaces cv::triangulatePoints(P0, P1, p0, p1, X4): src/initializer.cpp:125, src/tracker.cpp:149.
 * X4 is 4 x n f32 row-major (the cv::Mat layout).  Each column is a unit vector, sign unspecified. */
MVO_API int mvo_triangulate(mvo_ctx* ctx, const double P0[12], const double P1[12], const float* p0,
                            const float* p1, int n, float* X4);

/* replaces cv::solvePnPRansac(points_3d, points_2d, K, d, rvec, tvec, false, iterations, reproj_err, confidence,
 * inliers): src/tracker.cpp:309 (the reference passes 100, 8.0, 0.99; SOLVEPNP_ITERATIVE; no extrinsic guess).
 * obj_xyz: n x 3 f32 (std::vector<cv::Point3f>), img_xy: n x 2 f32.  dist: n_dist distortion coefficients
 * (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]) or NULL.  Non-zero coefficients: the image points are undistorted on the
 * device (cv::undistortPoints, as OpenCV's minimal solver does) and the search runs in the distortion-free camera, i.e.
 * the 8 px threshold and the refinement are measured on undistorted pixels (OpenCV: on distorted ones); all-zero
 * coefficients (rectified images) take the bit-compatible path.  n < 6: MVO_ERR_DEGENERATE.  inliers: capacity n indices of the winning hypothesis' inliers (may be
 * NULL), *n_inliers their count.  rvec / tvec: the Levenberg-Marquardt pose on those inliers.
 * MVO_ERR_DEGENERATE == OpenCV returning false (no model). */
MVO_API int mvo_solve_pnp_ransac(mvo_ctx* ctx, const float* obj_xyz, const float* img_xy, int n, const double K[9],
                                 const double* dist, int n_dist, int iterations, double reproj_err,
                                 double confidence, double rvec[3], double tvec[3], int32_t* inliers,
                                 int* n_inliers);
/* parity hook: hypotheses of the last mvo_solve_pnp_ransac call -- subsets: iterations x 5, models: iterations x 12
 * (R row-major | t), counts: iterations (-1 where the minimal solver failed).  Any pointer may be NULL. */
MVO_API int mvo_pnp_get_hypotheses(mvo_ctx* ctx, int iterations, int32_t* subsets, double* models, int32_t* counts);
/* replaces cv::Rodrigues(rvec, R): src/tracker.cpp:315 (host arithmetic; R is 3x3 row-major) */
MVO_API int mvo_rodrigues(const double rvec[3], double R[9]);

/* model ids for the hypothesis sweep */
enum { MVO_MODEL_H = 0, MVO_MODEL_F = 1, MVO_MODEL_E = 2 };
/* C4 sweep: draw m minimal samples with the OpenCV RNG stream (seeded per ctx), solve, and score every
 * model over all n correspondences with no early exit.  counts: m x max_models ints (-1 = no model),
 * max_models = 1 (H), 3 (F), 10 (E).  models (optional): m x max_models x 9 doubles. */
MVO_API int mvo_score_hypotheses(mvo_ctx* ctx, int model, const float* p1, const float* p2, int n,
                                 const double K[9], double thr, int m, int32_t* sample_idx,
                                 int32_t* counts, double* models);

/* ------------------------------------------------------------------------------------------------
 * stream-group front-end step (bench / batched mode; SURVEY.md section 8d "front-end frame"):
 * for every stream b of the group: ORB(frame_t) -> kNN+ratio(desc_{t-1}, desc_t) -> LK(kps_{t-1} -> t)
 * -> H + F RANSAC -> E RANSAC -> recoverPose -> triangulate.  Images: batch x h x w u8 gray.
 */
typedef struct mvo_frame_result {
  int32_t n_keypoints;   /* ORB keypoints of frame t                    */
  int32_t n_matches;     /* accepted ratio-test matches (t-1 -> t)      */
  int32_t n_tracked;     /* LK points with status && err < 30           */
  int32_t score_h;       /* findHomography inliers                      */
  int32_t score_f;       /* findFundamentalMat inliers                  */
  int32_t n_inliers_e;   /* findEssentialMat inliers                    */
  int32_t n_pose_good;   /* recoverPose cheirality count                */
  int32_t n_triangulated;/* points in front of both cameras             */
  double R[9];
  double t[3];
} mvo_frame_result;

MVO_API int mvo_group_step(mvo_ctx* ctx, const uint8_t* images, int w, int h, int stride,
                           int images_on_device, const double K[9], mvo_frame_result* results);
/* Pipelined form of mvo_group_step: submit enqueues the upload (copy stream, double-buffered staging: pass PINNED host
 * frames) and the kernels of one step and returns at once; collect waits for the OLDEST submitted step and returns its
 * results.  Up to two steps may be in flight, so the H2D copy of step t+1 overlaps the kernels of step t.
 * BUFFER LIFETIME: the upload is asynchronous -- host frames passed to mvo_group_submit must stay valid and unmodified
 * until the mvo_group_collect of that step has returned (device frames: until then as well).
 * mvo_stage_ms reports the most recently enqueued step and is only meaningful when nothing else is in flight. */
MVO_API int mvo_group_submit(mvo_ctx* ctx, const uint8_t* images, int w, int h, int stride, int images_on_device,
                             const double K[9]);
MVO_API int mvo_group_collect(mvo_ctx* ctx, mvo_frame_result* results);
/* forget the previous frame of every stream (the next step only extracts features) */
MVO_API int mvo_group_reset(mvo_ctx* ctx);
/* per-stage device time (ms) of the last mvo_group_step, measured with CUDA events on the ctx stream.
 * names: "orb", "orb_dense" (the fused per-level kernels inside "orb"), "knn", "lk", "ransac_h", "ransac_f",
 * "ransac_e", "pose", "triangulate", "total" */
MVO_API int mvo_stage_ms(mvo_ctx* ctx, const char* stage, float* ms);
/* start / end of a stage of the last enqueued step relative to the start of that step (ms): the schedule of the step's
 * dependency graph over the context's CUDA streams (profiling aid; same validity rule as mvo_stage_ms) */
MVO_API int mvo_stage_span_ms(mvo_ctx* ctx, const char* stage, float* beg_ms, float* end_ms);

/* ------------------------------------------------------------------------------------------------
 * stream-group product surface: BGR8 input and the full per-stream outputs of a group step -- everything the
 * reference's data flow consumes (Frame::extract_observations needs keypoints + descriptors, src/frame.cpp:8-17; the
 * tracker needs the tracked points, src/tracker.cpp:70-77; landmark creation needs the masks and the 3-D points,
 * src/initializer.cpp:283-298).
 */
enum {
  MVO_OUT_KEYPOINTS = 1,  /* keypoints + descriptors of frame t                                   */
  MVO_OUT_MATCHES = 2,    /* ratio-test matches (t-1 -> t)                                        */
  MVO_OUT_TRACKS = 4,     /* LK next position / status / err of every keypoint of frame t-1       */
  MVO_OUT_MODELS = 8,     /* H, F, E and their inlier masks, the recoverPose mask                 */
  MVO_OUT_POINTS3D = 16,  /* triangulated homogeneous points                                      */
  MVO_OUT_CLOUD = 32,     /* the chirality-valid triangulated points, dehomogenised, in ROS axes, packed as the
                             sensor_msgs/PointCloud2 payload (12 bytes per point)                   */
  MVO_OUT_ALL = 63
};
typedef struct mvo_group_config {
  int32_t channels;   /* frames given to mvo_group_step / submit / track: 1 = gray8, 3 = BGR8 (the node feeds BGR8,
                         src/mono_vo.cpp:94: ORB runs on the fused BGR->gray conversion, LK on the three colour planes
                         exactly as cv::calcOpticalFlowPyrLK does on 3-channel Mats) */
  uint32_t outputs;   /* MVO_OUT_* mask: these outputs are copied to pinned host memory inside every step */
} mvo_group_config;
/* valid while no step is in flight; forgets the previous frame of every stream */
MVO_API int mvo_group_configure(mvo_ctx* ctx, const mvo_group_config* cfg);
/* bytes copied device -> host per step with the configured outputs (result records included) */
MVO_API int mvo_group_output_bytes(mvo_ctx* ctx, size_t* bytes);

/* Outputs of one stream for the step most recently returned by mvo_group_step / mvo_group_collect.  The pointers are
 * views into pinned host memory owned by the context (no copy); they stay valid until the next mvo_group_submit /
 * mvo_group_step / mvo_group_configure call.  Sections that were not requested (or have no previous frame) are NULL / 0.
 *   keypoints / descriptors : n_keypoints entries, canonical order (as mvo_orb_detect_and_compute)
 *   matches                 : n_matches entries, query = keypoints of frame t-1, train = keypoints of frame t
 *   track_*                 : n_prev entries, one per keypoint of frame t-1 (as mvo_lk_track returns them)
 *   mask_*                  : n_tracked entries over the ordered list of tracks with status && err < 30, i.e. the
 *                             p1 / p2 vectors of src/tracker.cpp:70-77; mask_pose is recoverPose's in/out mask (0 / 255)
 *   X4                      : 4 rows of n_tracked floats, row r at X4 + r * x4_stride (cv::triangulatePoints' 4 x N Mat)
 */
typedef struct mvo_stream_outputs {
  int32_t n_keypoints, n_matches, n_prev, n_tracked;
  const mvo_keypoint* keypoints;
  const uint8_t* descriptors;
  const mvo_dmatch* matches;
  const float* track_xy;
  const uint8_t* track_status;
  const float* track_err;
  const uint8_t* mask_h;
  const uint8_t* mask_f;
  const uint8_t* mask_e;
  const uint8_t* mask_pose;
  const float* X4;
  int64_t x4_stride;
  double H[9], F[9], E[9];
  int32_t flags;        /* bit0: FAST candidate list overflowed, bit1: keypoints truncated to the capacity */
  int32_t n_cloud;      /* points in cloud_xyz (== mvo_frame_result.n_triangulated)                        */
  const float* cloud_xyz;   /* MVO_OUT_CLOUD: n_cloud x (x, y, z) f32 = PointCloud2.data of points3d_to_pointcloud_msg
                               (src/utils.cpp:184-243) for the points with mask_pose set and z > 0 in both cameras,
                               in the order of the tracks                                                   */
  int32_t occupied_cells, total_cells;   /* keypoint distribution grid of frame t (see mvo_orb_occupancy)  */
} mvo_stream_outputs;
MVO_API int mvo_group_outputs(mvo_ctx* ctx, int stream, mvo_stream_outputs* out);

/* ------------------------------------------------------------------------------------------------
 * stream-group tracking frame == Tracker::update's per-frame path (src/tracker.cpp:274-316) for every stream:
 *   LK(previous frame -> new frame) on the stream's tracked observations (track_frame_with_optical_flow, :57-92),
 *   keep status && err < 30 in order, solvePnPRansac(landmarks, tracked points, K, 100, 8.0, 0.99) (:309).
 * The tracked set (image positions + landmark coordinates) stays on the device and is the next frame's input;
 * mvo_group_set_tracks replaces it for one stream (after initialisation / a new keyframe, src/tracker.cpp:176-235).
 * The first call after create / configure / set of a new size only stores the frame (Tracker::update :285-289).
 */
typedef struct mvo_track_result {
  int32_t n_prev;         /* observations tracked from                                    */
  int32_t n_tracked;      /* kept (status && err < 30): the new frame's observations      */
  int32_t n_pnp_inliers;  /* inliers of the winning solvePnPRansac hypothesis             */
  int32_t pnp_ok;         /* 1: rvec / tvec valid (solvePnPRansac returned true)          */
  double rvec[3];
  double tvec[3];
} mvo_track_result;
/* xy: n x 2 f32 pixel positions in the stream's latest frame, xyz: n x 3 f32 landmark coordinates */
MVO_API int mvo_group_set_tracks(mvo_ctx* ctx, int stream, const float* xy, const float* xyz, int n);
MVO_API int mvo_group_track(mvo_ctx* ctx, const uint8_t* images, int w, int h, int stride, int images_on_device,
                            const double K[9], mvo_track_result* results);
/* the stream's observations after the last mvo_group_track: position, index into the previous frame's observation
 * list (so the host can carry descriptor / landmark ids along, src/tracker.cpp:73-76), and the PnP inlier indices
 * (into the new list).  Any pointer may be NULL; cap = capacity of each array. */
MVO_API int mvo_group_get_tracks(mvo_ctx* ctx, int stream, float* xy, int32_t* src_idx, int32_t* pnp_inliers, int cap,
                                 int* n_tracked, int* n_inliers);

/* ------------------------------------------------------------------------------------------------
 * device-resident caches of the synchronous single-call path.  The reference gathers the same N x 32
 * descriptor block before every match (src/frame.cpp:50-64, src/keyframe.cpp:78-92) and tracks every frame against the
 * image that was the "next" image of the previous call (src/tracker.cpp:68-69, :331); both are recognised BY CONTENT
 * (64-bit hash of the host buffer, so a modified or re-allocated buffer can never alias a stale device copy):
 *   - descriptors returned by mvo_orb_detect_and_compute and blocks uploaded by mvo_knn_ratio / mvo_knn2 stay on the
 *     device (4 most recently used blocks); a later match on the same content skips the upload;
 *   - mvo_lk_track keeps both pyramids; an image whose pyramid is still resident is neither uploaded nor rebuilt.
 * stats: descriptor hits, descriptor misses, pyramid hits, pyramid misses since mvo_create.
 * mvo_debug_set("cache", 0) switches both caches off (A/B aid). */
MVO_API int mvo_cache_stats(mvo_ctx* ctx, uint64_t stats[4]);

/* ------------------------------------------------------------------------------------------------
 * by-products that save the reference's host loops.
 *
 * Keypoint-distribution grid == Initializer::good_keypoint_distribution (src/initializer.cpp:52-75): a grid of
 * (rows / div) x (cols / div) cells, a cell is occupied when a keypoint has r = (int)(pt.y / div), c = (int)(pt.x / div)
 * (flat index r * grid_cols + c, exactly the cv::Mat::at the reference performs; indices past the grid are ignored).
 * The counts are a by-product of the kernel that writes the final keypoints (no extra pass over them).  div defaults
 * to 50 (config/params.yaml: initializer.occupancy_grid_div); 0 switches the by-product off.  The reference's test is
 * (double)occupied / total > kp_distribution_thresh.
 * mvo_orb_occupancy: counts of the last mvo_orb_detect_and_compute call (stream 0) or of the group step most recently
 * returned by mvo_group_step / mvo_group_collect. */
MVO_API int mvo_set_occupancy_grid(mvo_ctx* ctx, int grid_div);
MVO_API int mvo_orb_occupancy(mvo_ctx* ctx, int stream, int* occupied_cells, int* total_cells);
/* points3d_to_pointcloud_msg's payload (src/utils.cpp:225-241) on the device: n points (x, y, z) f32 in the camera /
 * OpenCV frame -> n x 12 bytes (ROS x = z, ROS y = -x, ROS z = -y), i.e. sensor_msgs::msg::PointCloud2::data with
 * point_step 12, fields x / y / z FLOAT32 at offsets 0 / 4 / 8, height 1, width n, little endian, dense.
 * points_on_device / out_on_device: the buffers are device pointers (MVO_MEM_DEVICE) instead of host memory. */
MVO_API int mvo_pack_pointcloud(mvo_ctx* ctx, const float* points_xyz, int n, int points_on_device, uint8_t* out,
                                int out_on_device);

/* ------------------------------------------------------------------------------------------------
 * profiling / parity aids (not part of the reference-facing surface)
 * mvo_debug_set: "lk_impl" = 1 | 2 selects the first- / second-generation gray LK kernel (identical results; the old
 *                one is the in-tree cross-check), "knn_impl" likewise for the matching kernels, "h_refine_impl" / "e5_roots_impl"
 *                = 1 for the first-generation homography refinement / the bracketing root finder, "pnp_epnp_impl" = 0 /
 *                "pnp_refine_impl" = 1 for the round-1 EPnP eigen-solver / refinement start, "pnp_rounds" = 1 to evaluate all
 *                hypotheses of a batched solvePnPRansac in one round, "cache" = 0 and "graph" = 0 to switch the content
 *                caches / the CUDA-graph step off.
 * mvo_debug_time: re-runs one stage of the group pipeline `reps` times on the state left by the last mvo_group_step
 *                (at least two steps must have run) and returns the average device time per run in ms, measured with
 *                CUDA events on the context stream.  what = "lk_track" | "knn" | "orb" | "orb_levels" (the eight fused
 *                level kernels as a step launches them) | "orb_dense" (pyramid + blur only) | "lk_pyramid".
 */
MVO_API int mvo_debug_set(mvo_ctx* ctx, const char* key, int value);
/* CUDA-graph form of the synchronous mvo_group_step (groups whose frames are <= 4 MB, host frames): the step is captured
 * once per buffer parity and replayed with one launch -- the ~85 launches of a step, not the GPU, bound a single stream.
 * stats: graphs captured, steps replayed from a graph, captures abandoned (the plain step ran instead).
 * mvo_debug_set("graph", 0) switches the graph form off (mvo_stage_ms needs plain steps). */
MVO_API int mvo_graph_stats(mvo_ctx* ctx, uint64_t stats[3]);
MVO_API int mvo_debug_time(mvo_ctx* ctx, const char* what, int reps, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* MONOVO_B200_H_ */
