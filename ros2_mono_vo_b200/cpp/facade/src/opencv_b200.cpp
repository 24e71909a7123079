// opencv_b200.cpp -- implementation of the OpenCV-API facade (facade/include/opencv2/core.hpp): the hot cv:: functions
// forward to libmonovo_b200.so through the C ABI of include/monovo_b200.h; the cold helpers are plain host code.
// There is no CPU fallback for a hot function: a missing library / CUDA device surfaces as cv::Exception, exactly where
// OpenCV itself would throw.
#include <algorithm>
#include <mutex>

#include "monovo_b200.h"
#include "opencv2/core.hpp"

namespace cv {

// ---- context -----------------------------------------------------------------------------------------------------
namespace b200 {
namespace {
struct State {
  mvo_ctx* ctx = nullptr;
  int w = 0, h = 0, nfeatures = 0;
  unsigned long long launches_retired = 0;   // launches of contexts that were replaced by a larger one
  std::mutex mu;
};
State& state() {
  static State s;
  return s;
}
}  // namespace

mvo_ctx* context(int min_width, int min_height, int nfeatures) {
  State& s = state();
  std::lock_guard<std::mutex> lock(s.mu);
  const int want_n = nfeatures > 0 ? nfeatures : (s.nfeatures > 0 ? s.nfeatures : 1000);
  // max_width / max_height are limits the library checks image sizes against, not allocation sizes: the ABI's maximum,
  // so a larger image never replaces the context (and with it the cached pyramids / descriptor blocks).  Only a
  // different ORB feature count (a second cv::ORB::create(n) with another n) builds a new one.
  (void)min_width;
  (void)min_height;
  const int want_w = 16384, want_h = 16384;
  if (s.ctx && want_n == s.nfeatures) return s.ctx;
  if (s.ctx) {
    s.launches_retired += mvo_launch_count(s.ctx);
    mvo_destroy(s.ctx);
    s.ctx = nullptr;
  }
  mvo_config cfg{};
  cfg.device = 0;
  cfg.max_width = want_w;
  cfg.max_height = want_h;
  cfg.nfeatures = want_n;
  cfg.batch = 1;
  cfg.max_points = std::max(4 * want_n, 8192);   // LK / RANSAC point sets: landmarks tracked + matches
  if (mvo_create(&s.ctx, &cfg) != MVO_OK) throw Exception(std::string("monovo_b200: ") + mvo_last_error(nullptr));
  s.w = want_w;
  s.h = want_h;
  s.nfeatures = want_n;
  return s.ctx;
}

unsigned long long launch_count() {
  State& s = state();
  std::lock_guard<std::mutex> lock(s.mu);
  return s.launches_retired + (s.ctx ? mvo_launch_count(s.ctx) : 0ull);
}

void shutdown() {
  State& s = state();
  std::lock_guard<std::mutex> lock(s.mu);
  if (s.ctx) {
    s.launches_retired += mvo_launch_count(s.ctx);
    mvo_destroy(s.ctx);
  }
  s.ctx = nullptr;
  s.w = s.h = s.nfeatures = 0;
}
}  // namespace b200

namespace {
void check(mvo_ctx* c, int rc, const char* what) {
  if (rc != MVO_OK) throw Exception(std::string(what) + ": " + mvo_last_error(c));
}
Mat mat_from(const double* v, int r, int c) {
  Mat m(r, c, CV_64F);
  for (int i = 0; i < r * c; ++i) m.at<double>(i / c, i % c) = v[i];
  return m;
}
void to_array(const Mat& m, double* out, int r, int c, const char* what) {
  if (m.rows != r || m.cols != c || m.channels() != 1) throw Exception(std::string(what) + ": unexpected matrix size");
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) out[i * c + j] = m.getd(i, j);
}
const float* xy(const std::vector<Point2f>& p) { return p.empty() ? nullptr : &p[0].x; }
static_assert(sizeof(Point2f) == 8 && sizeof(Point3f) == 12, "point vectors are passed to the ABI as packed floats");
}  // namespace

// ---- cold helpers ------------------------------------------------------------------------------------------------
Mat Mat::eye(int r, int c, int type) {
  Mat m(r, c, type);
  for (int i = 0; i < std::min(r, c); ++i) {
    if (m.depth() == CV_64F) m.at<double>(i, i) = 1.0;
    else if (m.depth() == CV_32F) m.at<float>(i, i) = 1.f;
    else if (m.depth() == CV_32S) m.at<int>(i, i) = 1;
    else m.at<uchar>(i, i) = 1;
  }
  return m;
}

Mat Mat::t() const {
  if (channels() != 1) throw Exception("Mat::t: single-channel matrices only");
  Mat m(cols, rows, type_);
  const size_t es = elemSize();
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) std::memcpy(m.data + (size_t)c * m.step + (size_t)r * es, data + (size_t)r * step + (size_t)c * es, es);
  return m;
}

Mat operator*(const Mat& a, const Mat& b) {
  if (a.cols != b.rows || a.type() != CV_64F || b.type() != CV_64F) throw Exception("Mat * Mat: CV_64F matrices with matching inner size only");
  Mat m(a.rows, b.cols, CV_64F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      double s = 0;
      for (int k = 0; k < a.cols; ++k) s += a.at<double>(i, k) * b.at<double>(k, j);
      m.at<double>(i, j) = s;
    }
  return m;
}

Mat operator+(const Mat& a, const Mat& b) {
  if (a.rows != b.rows || a.cols != b.cols || a.type() != CV_64F || b.type() != CV_64F) throw Exception("Mat + Mat: CV_64F matrices of equal size only");
  Mat m(a.rows, a.cols, CV_64F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < a.cols; ++j) m.at<double>(i, j) = a.at<double>(i, j) + b.at<double>(i, j);
  return m;
}

std::ostream& operator<<(std::ostream& os, const Mat& m) {
  os << "[";
  for (int i = 0; i < m.rows; ++i) {
    for (int j = 0; j < m.cols; ++j) os << m.getd(i, j) << (j + 1 < m.cols ? ", " : "");
    os << (i + 1 < m.rows ? ";\n " : "]");
  }
  return os;
}

int countNonZero(const std::vector<uchar>& v) { return (int)std::count_if(v.begin(), v.end(), [](uchar x) { return x != 0; }); }

int countNonZero(const Mat& m) {
  int n = 0;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) n += m.getd(r, c) != 0.0;
  return n;
}

double norm(const Mat& m) {
  double s = 0;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) s += m.getd(r, c) * m.getd(r, c);
  return std::sqrt(s);
}

void hconcat(const Mat& a, const Mat& b, Mat& dst) {
  if (a.rows != b.rows || a.type() != b.type()) throw Exception("hconcat: row count / type mismatch");
  Mat m(a.rows, a.cols + b.cols, a.type());
  const size_t es = a.elemSize();
  for (int r = 0; r < a.rows; ++r) {
    std::memcpy(m.data + (size_t)r * m.step, a.data + (size_t)r * a.step, (size_t)a.cols * es);
    std::memcpy(m.data + (size_t)r * m.step + (size_t)a.cols * es, b.data + (size_t)r * b.step, (size_t)b.cols * es);
  }
  dst = m;
}

// OpenCV: scale = w != 0 ? 1 / w : 1 (float arithmetic for float input)
void convertPointsFromHomogeneous(const Mat& src, std::vector<Point3f>& dst) {
  if (src.cols != 4 || src.type() != CV_32F) throw Exception("convertPointsFromHomogeneous: N x 4 CV_32F input only");
  dst.resize(src.rows);
  for (int i = 0; i < src.rows; ++i) {
    const float* p = src.ptr<float>(i);
    const float s = p[3] != 0.f ? 1.f / p[3] : 1.f;
    dst[i] = Point3f(p[0] * s, p[1] * s, p[2] * s);
  }
}

void convertPointsFromHomogeneous(const Mat& src, Mat& dst) {
  std::vector<Point3f> v;
  convertPointsFromHomogeneous(src, v);
  dst = Mat((int)v.size(), 1, CV_32FC3);
  for (size_t i = 0; i < v.size(); ++i) dst.at<Point3f>((int)i, 0) = v[i];
}

// ---- cv::ORB / cv::BFMatcher -------------------------------------------------------------------------------------------
Ptr<ORB> ORB::create(int nfeatures) { return Ptr<ORB>(new ORB(nfeatures)); }

void ORB::run(const Mat& image, std::vector<KeyPoint>& keypoints, Mat* descriptors) {
  keypoints.clear();
  if (descriptors) *descriptors = Mat();
  if (image.empty()) return;
  if (image.depth() != CV_8U || (image.channels() != 1 && image.channels() != 3)) throw Exception("ORB: 8-bit gray or BGR images only");
  mvo_ctx* c = b200::context(image.cols, image.rows, nfeatures_);
  const int cap = nfeatures_ + nfeatures_ / 4 + 64;   // retainBest keeps ties: N can exceed nfeatures
  std::vector<mvo_keypoint> kps(cap);
  std::vector<uchar> desc(descriptors ? (size_t)cap * 32 : 0);
  int n = 0;
  check(c, mvo_orb_detect_and_compute(c, image.data, image.cols, image.rows, (int)image.step, image.channels(), kps.data(),
                                     descriptors ? desc.data() : nullptr, cap, &n), "ORB::detectAndCompute");
  keypoints.reserve(n);
  for (int i = 0; i < n; ++i) {
    const mvo_keypoint& k = kps[i];
    keypoints.emplace_back(Point2f(k.x, k.y), k.size, k.angle, k.response, k.octave, k.class_id);
  }
  if (descriptors && n > 0) {
    descriptors->create(n, 32, CV_8UC1);
    std::memcpy(descriptors->data, desc.data(), (size_t)n * 32);
  }
}

void ORB::detect(const Mat& image, std::vector<KeyPoint>& keypoints) { run(image, keypoints, nullptr); }

void ORB::detectAndCompute(const Mat& image, NoArray, std::vector<KeyPoint>& keypoints, Mat& descriptors) {
  run(image, keypoints, &descriptors);
}

BFMatcher::BFMatcher(int normType, bool crossCheck) : norm_(normType) {
  if (normType != NORM_HAMMING || crossCheck) throw Exception("BFMatcher: NORM_HAMMING without cross-check only (the reference's configuration)");
}

void BFMatcher::knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& matches, int k) const {
  if (k != 2) throw Exception("BFMatcher::knnMatch: k = 2 only (the reference's call)");
  matches.clear();
  const int nq = q.rows, nt = t.rows;
  if (nq == 0) return;
  matches.resize(nq);
  if (nt == 0) return;
  if (q.cols != 32 || t.cols != 32 || q.type() != CV_8UC1 || t.type() != CV_8UC1) throw Exception("BFMatcher::knnMatch: N x 32 CV_8U descriptors only");
  const Mat qd = q.isContinuous() ? q : q.clone(), td = t.isContinuous() ? t : t.clone();
  mvo_ctx* c = b200::context();
  std::vector<int32_t> idx((size_t)nq * 2), dist((size_t)nq * 2);
  check(c, mvo_knn2(c, qd.data, nq, td.data, nt, idx.data(), dist.data()), "BFMatcher::knnMatch");
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < 2; ++j)
      if (idx[2 * i + j] >= 0) matches[i].emplace_back(i, idx[2 * i + j], 0, (float)dist[2 * i + j]);
}

// ---- cv::calcOpticalFlowPyrLK ----------------------------------------------------------------------------------------------
void calcOpticalFlowPyrLK(const Mat& prev, const Mat& next, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                          std::vector<uchar>& status, std::vector<float>& err) {
  if (prev.rows != next.rows || prev.cols != next.cols || prev.type() != next.type()) throw Exception("calcOpticalFlowPyrLK: image size / type mismatch");
  const int n = (int)prevPts.size();
  nextPts.assign(n, Point2f());
  status.assign(n, 0);
  err.assign(n, 0.f);
  if (n == 0) return;
  mvo_ctx* c = b200::context(prev.cols, prev.rows);
  const Mat a = prev.step == next.step ? prev : prev.clone(), b = prev.step == next.step ? next : next.clone();
  check(c, mvo_lk_track(c, a.data, b.data, a.cols, a.rows, (int)a.step, a.channels(), xy(prevPts), n, &nextPts[0].x, status.data(),
                       err.data()), "calcOpticalFlowPyrLK");
}

// ---- two-view geometry ---------------------------------------------------------------------------------------------------------
Mat findHomography(const std::vector<Point2f>& p1, const std::vector<Point2f>& p2, int method, double thr, std::vector<uchar>& mask) {
  if (method != RANSAC) throw Exception("findHomography: cv::RANSAC only (the reference's call)");
  if (p1.size() != p2.size()) throw Exception("findHomography: point count mismatch");
  mask.assign(p1.size(), 0);
  if (p1.size() < 4) return Mat();
  mvo_ctx* c = b200::context();
  double H[9];
  int n_in = 0;
  const int rc = mvo_find_homography(c, xy(p1), xy(p2), (int)p1.size(), thr, H, mask.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return Mat();   // OpenCV: empty matrix when no model is found
  check(c, rc, "findHomography");
  return mat_from(H, 3, 3);
}

Mat findFundamentalMat(const std::vector<Point2f>& p1, const std::vector<Point2f>& p2, int method, double thr, double conf,
                       std::vector<uchar>& mask) {
  if (method != FM_RANSAC) throw Exception("findFundamentalMat: cv::FM_RANSAC only (the reference's call)");
  if (p1.size() != p2.size()) throw Exception("findFundamentalMat: point count mismatch");
  mask.assign(p1.size(), 0);
  if (p1.size() < 7) return Mat();
  mvo_ctx* c = b200::context();
  double F[9];
  int n_in = 0;
  const int rc = mvo_find_fundamental(c, xy(p1), xy(p2), (int)p1.size(), thr, conf, F, mask.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return Mat();
  check(c, rc, "findFundamentalMat");
  return mat_from(F, 3, 3);
}

Mat findEssentialMat(const std::vector<Point2f>& p1, const std::vector<Point2f>& p2, const Mat& K, int method, double prob,
                     double threshold, std::vector<uchar>& mask) {
  if (method != RANSAC) throw Exception("findEssentialMat: cv::RANSAC only (the reference's call)");
  if (p1.size() != p2.size()) throw Exception("findEssentialMat: point count mismatch");
  mask.assign(p1.size(), 0);
  if (p1.size() < 5) return Mat();
  double k[9], E[9];
  to_array(K, k, 3, 3, "findEssentialMat: cameraMatrix");
  mvo_ctx* c = b200::context();
  int n_in = 0;
  const int rc = mvo_find_essential(c, xy(p1), xy(p2), (int)p1.size(), k, prob, threshold, E, mask.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return Mat();
  check(c, rc, "findEssentialMat");
  return mat_from(E, 3, 3);
}

int recoverPose(const Mat& E, const std::vector<Point2f>& p1, const std::vector<Point2f>& p2, const Mat& K, Mat& R, Mat& t,
                std::vector<uchar>& mask) {
  double e[9], k[9], r[9], tt[3];
  to_array(E, e, 3, 3, "recoverPose: E");
  to_array(K, k, 3, 3, "recoverPose: cameraMatrix");
  if (mask.size() != p1.size()) mask.assign(p1.size(), 1);   // OpenCV: an empty mask means "all points"
  mvo_ctx* c = b200::context();
  int good = 0;
  check(c, mvo_recover_pose(c, e, xy(p1), xy(p2), (int)p1.size(), k, r, tt, mask.data(), &good), "recoverPose");
  R = mat_from(r, 3, 3);
  t = mat_from(tt, 3, 1);
  return good;
}

void triangulatePoints(const Mat& P0, const Mat& P1, const std::vector<Point2f>& pts0, const std::vector<Point2f>& pts1, Mat& points4D) {
  if (pts0.size() != pts1.size()) throw Exception("triangulatePoints: point count mismatch");
  double p0[12], p1[12];
  to_array(P0, p0, 3, 4, "triangulatePoints: projMatr1");
  to_array(P1, p1, 3, 4, "triangulatePoints: projMatr2");
  const int n = (int)pts0.size();
  points4D = Mat(4, n, CV_32F);
  if (n == 0) return;
  mvo_ctx* c = b200::context();
  check(c, mvo_triangulate(c, p0, p1, xy(pts0), xy(pts1), n, points4D.ptr<float>(0)), "triangulatePoints");
}

bool solvePnPRansac(const std::vector<Point3f>& obj, const std::vector<Point2f>& img, const Mat& K, const Mat& dist, Mat& rvec, Mat& tvec,
                    bool useExtrinsicGuess, int iterations, float reprojErr, double confidence, Mat& inliers) {
  if (useExtrinsicGuess) throw Exception("solvePnPRansac: useExtrinsicGuess is not supported (the reference passes false)");
  if (obj.size() != img.size()) throw Exception("solvePnPRansac: point count mismatch");
  double k[9], r[3], t[3];
  to_array(K, k, 3, 3, "solvePnPRansac: cameraMatrix");
  std::vector<double> d;
  for (int i = 0; i < dist.rows; ++i)
    for (int j = 0; j < dist.cols; ++j) d.push_back(dist.getd(i, j));
  const int n = (int)obj.size();
  std::vector<int32_t> idx(std::max(n, 1));
  int n_in = 0;
  inliers = Mat();
  mvo_ctx* c = b200::context();
  const int rc = mvo_solve_pnp_ransac(c, n ? &obj[0].x : nullptr, xy(img), n, k, d.empty() ? nullptr : d.data(), (int)d.size(), iterations,
                                      reprojErr, confidence, r, t, idx.data(), &n_in);
  if (rc == MVO_ERR_DEGENERATE) return false;   // OpenCV: false when no model is found
  check(c, rc, "solvePnPRansac");
  rvec = mat_from(r, 3, 1);
  tvec = mat_from(t, 3, 1);
  inliers = Mat(n_in, 1, CV_32SC1);
  for (int i = 0; i < n_in; ++i) inliers.at<int>(i, 0) = idx[i];
  return true;
}

void Rodrigues(const Mat& src, Mat& dst) {
  if (src.total() != 3) throw Exception("Rodrigues: rotation vector -> matrix only (the reference's call)");
  const double r[3] = {src.getd(0, 0), src.rows == 3 ? src.getd(1, 0) : src.getd(0, 1), src.rows == 3 ? src.getd(2, 0) : src.getd(0, 2)};
  double m[9];
  mvo_rodrigues(r, m);
  dst = mat_from(m, 3, 3);
}

}  // namespace cv
