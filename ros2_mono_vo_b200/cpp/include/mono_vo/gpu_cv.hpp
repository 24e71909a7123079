// gpu_cv.hpp -- B200 replacements for the free cv:: functions the reference calls from Initializer / Tracker.
// Each function has the argument list of the OpenCV call it replaces (the subset of defaults the reference
// uses), so a call site changes by one token:  cv::findHomography(...)  ->  mono_vo::gpu::findHomography(ctx, ...).
//   src/tracker.cpp:68-69       cv::calcOpticalFlowPyrLK
//   src/initializer.cpp:82      cv::findHomography          src/tracker.cpp:243
//   src/initializer.cpp:87      cv::findFundamentalMat      src/tracker.cpp:248
//   src/initializer.cpp:228     cv::findEssentialMat
//   src/initializer.cpp:236     cv::recoverPose
//   src/initializer.cpp:125     cv::triangulatePoints       src/tracker.cpp:149
//   src/tracker.cpp:309         cv::solvePnPRansac          src/tracker.cpp:315  cv::Rodrigues
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "mono_vo/feature_processor.hpp"
#include "monovo_b200.h"

namespace mono_vo
{
namespace gpu
{
inline void check(mvo_ctx * c, int rc, const char * what)
{
  if (rc != MVO_OK) throw std::runtime_error(std::string(what) + ": " + mvo_last_error(c));
}

inline cv::Mat mat3x3(const double * m)
{
  cv::Mat r(3, 3, CV_64F);
  for (int i = 0; i < 9; ++i) r.at<double>(i / 3, i % 3) = m[i];
  return r;
}

// first coordinate of a point vector, NULL when it is empty (&v[0] on an empty vector is undefined behaviour)
inline const float * xy_ptr(const std::vector<cv::Point2f> & v) { return v.empty() ? nullptr : &v[0].x; }

inline void mat_to_array(const cv::Mat & m, double * out, int n)
{
  if (m.empty() || m.rows * m.cols < n) throw std::runtime_error("gpu_cv: matrix is empty or too small");
  for (int i = 0; i < n; ++i) out[i] = m.at<double>(i / m.cols, i % m.cols);
}

/// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err) with default arguments.
inline void calcOpticalFlowPyrLK(
  mvo_ctx * c, const cv::Mat & prev, const cv::Mat & next, const std::vector<cv::Point2f> & prev_pts,
  std::vector<cv::Point2f> & next_pts, std::vector<unsigned char> & status, std::vector<float> & err)
{
  const int n = static_cast<int>(prev_pts.size());
  next_pts.resize(n);
  status.resize(n);
  err.resize(n);
  check(c, mvo_lk_track(c, prev.data, next.data, prev.cols, prev.rows, static_cast<int>(prev.step), prev.channels(),
    n ? &prev_pts[0].x : nullptr, n, n ? &next_pts[0].x : nullptr, status.data(), err.data()), "calcOpticalFlowPyrLK");
}

/// cv::findHomography(p1, p2, cv::RANSAC, thr, mask)
inline cv::Mat findHomography(
  mvo_ctx * c, const std::vector<cv::Point2f> & p1, const std::vector<cv::Point2f> & p2, int /*method = RANSAC*/,
  double thr, std::vector<unsigned char> & mask)
{
  double H[9];
  int n_in = 0;
  mask.assign(p1.size(), 0);
  const int rc = mvo_find_homography(c, xy_ptr(p1), xy_ptr(p2), static_cast<int>(p1.size()), thr, H, mask.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return cv::Mat();   // OpenCV returns an empty matrix when no model is found
  check(c, rc, "findHomography");
  return mat3x3(H);
}

/// cv::findFundamentalMat(p1, p2, cv::FM_RANSAC, thr, conf, mask)
inline cv::Mat findFundamentalMat(
  mvo_ctx * c, const std::vector<cv::Point2f> & p1, const std::vector<cv::Point2f> & p2, int /*method*/, double thr,
  double conf, std::vector<unsigned char> & mask)
{
  double F[9];
  int n_in = 0;
  mask.assign(p1.size(), 0);
  const int rc = mvo_find_fundamental(c, xy_ptr(p1), xy_ptr(p2), static_cast<int>(p1.size()), thr, conf, F, mask.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return cv::Mat();
  check(c, rc, "findFundamentalMat");
  return mat3x3(F);
}

/// cv::findEssentialMat(p1, p2, K, cv::RANSAC, prob, threshold, mask)
inline cv::Mat findEssentialMat(
  mvo_ctx * c, const std::vector<cv::Point2f> & p1, const std::vector<cv::Point2f> & p2, const cv::Mat & K,
  int /*method*/, double prob, double threshold, std::vector<unsigned char> & mask)
{
  double k[9], E[9];
  mat_to_array(K, k, 9);
  int n_in = 0;
  mask.assign(p1.size(), 0);
  const int rc = mvo_find_essential(c, xy_ptr(p1), xy_ptr(p2), static_cast<int>(p1.size()), k, prob, threshold, E, mask.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return cv::Mat();
  check(c, rc, "findEssentialMat");
  return mat3x3(E);
}

/// cv::recoverPose(E, p1, p2, K, R, t, mask) -> number of points passing the cheirality check
inline int recoverPose(
  mvo_ctx * c, const cv::Mat & E, const std::vector<cv::Point2f> & p1, const std::vector<cv::Point2f> & p2,
  const cv::Mat & K, cv::Mat & R, cv::Mat & t, std::vector<unsigned char> & mask)
{
  double e[9], k[9], r[9], tt[3];
  mat_to_array(E, e, 9);
  mat_to_array(K, k, 9);
  int good = 0;
  check(c, mvo_recover_pose(c, e, xy_ptr(p1), xy_ptr(p2), static_cast<int>(p1.size()), k, r, tt,
    mask.size() == p1.size() ? mask.data() : nullptr, &good), "recoverPose");
  R = mat3x3(r);
  t = cv::Mat(3, 1, CV_64F);
  for (int i = 0; i < 3; ++i) t.at<double>(i, 0) = tt[i];
  return good;
}

/// cv::triangulatePoints(P0, P1, pts0, pts1, points4D): points4D is 4 x N CV_32F (row-major floats here)
inline void triangulatePoints(
  mvo_ctx * c, const cv::Mat & P0, const cv::Mat & P1, const std::vector<cv::Point2f> & pts0,
  const std::vector<cv::Point2f> & pts1, std::vector<float> & points4d)
{
  double p0[12], p1[12];
  mat_to_array(P0, p0, 12);
  mat_to_array(P1, p1, 12);
  points4d.assign(4 * pts0.size(), 0.f);
  check(c, mvo_triangulate(c, p0, p1, xy_ptr(pts0), xy_ptr(pts1), static_cast<int>(pts0.size()), points4d.data()), "triangulatePoints");
}

/// cv::solvePnPRansac(points_3d, points_2d, K, d, rvec, tvec, useExtrinsicGuess = false, iterationsCount, reprojectionError,
/// confidence, inliers) -- rvec / tvec become 3 x 1 CV_64F, inliers the indices of the winning hypothesis' inliers.
inline bool solvePnPRansac(
  mvo_ctx * c, const std::vector<cv::Point3f> & points_3d, const std::vector<cv::Point2f> & points_2d, const cv::Mat & K,
  const cv::Mat & dist, cv::Mat & rvec, cv::Mat & tvec, bool use_extrinsic_guess, int iterations, float reproj_err,
  double confidence, std::vector<int> & inliers)
{
  if (use_extrinsic_guess) throw std::runtime_error("solvePnPRansac: useExtrinsicGuess is not supported");
  double k[9], r[3], t[3];
  mat_to_array(K, k, 9);
  std::vector<double> d;
  for (int i = 0; i < dist.rows * dist.cols; ++i) d.push_back(dist.at<double>(i / dist.cols, i % dist.cols));
  const int n = static_cast<int>(points_3d.size());
  inliers.assign(n, 0);
  int n_in = 0;
  const int rc = mvo_solve_pnp_ransac(c, n ? &points_3d[0].x : nullptr, n ? &points_2d[0].x : nullptr, n, k,
    d.empty() ? nullptr : d.data(), static_cast<int>(d.size()), iterations, reproj_err, confidence, r, t, inliers.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) { inliers.clear(); return false; }   // OpenCV returns false when no model is found
  check(c, rc, "solvePnPRansac");
  inliers.resize(n_in);
  rvec = cv::Mat(3, 1, CV_64F);
  tvec = cv::Mat(3, 1, CV_64F);
  for (int i = 0; i < 3; ++i) { rvec.at<double>(i, 0) = r[i]; tvec.at<double>(i, 0) = t[i]; }
  return true;
}

/// cv::Rodrigues(rvec, R)
inline void Rodrigues(const cv::Mat & rvec, cv::Mat & R)
{
  double r[3] = {rvec.at<double>(0, 0), rvec.at<double>(1, 0), rvec.at<double>(2, 0)}, m[9];
  mvo_rodrigues(r, m);
  R = mat3x3(m);
}

/// the cv::Mat form the reference uses (src/initializer.cpp:124-125): points4D becomes 4 x N CV_32F
inline void triangulatePoints(
  mvo_ctx * c, const cv::Mat & P0, const cv::Mat & P1, const std::vector<cv::Point2f> & pts0,
  const std::vector<cv::Point2f> & pts1, cv::Mat & points4d)
{
  std::vector<float> x4;
  triangulatePoints(c, P0, P1, pts0, pts1, x4);
  const int n = static_cast<int>(pts0.size());
  points4d.create(4, n, CV_32F);
  for (int r = 0; r < 4; ++r)
    for (int i = 0; i < n; ++i) points4d.at<float>(r, i) = x4[static_cast<size_t>(r) * n + i];
}
}  // namespace gpu
}  // namespace mono_vo
