// opencv2/core.hpp -- OpenCV-API facade over libmonovo_b200.so.
//
// Purpose: let the reference's OWN, UNCHANGED host sources (src/feature_processor.cpp, frame.cpp, keyframe.cpp,
// landmark.cpp, map.cpp, match_data.cpp, initializer.cpp, tracker.cpp and their headers) compile and run in an
// image that has neither OpenCV nor ROS 2, with every hot cv:: call (cv::ORB, cv::BFMatcher::knnMatch,
// cv::calcOpticalFlowPyrLK, cv::findHomography / findFundamentalMat / findEssentialMat, cv::recoverPose,
// cv::triangulatePoints, cv::solvePnPRansac) executed by the B200 library through its C ABI
// (include/monovo_b200.h).  It is the link-time form of the drop-in: nothing in the reference is edited, the hot
// functions are simply resolved by this facade instead of libopencv.  (With a real OpenCV installed a maintainer
// uses the one-token call-site edits of INTEGRATION.md section 1 instead; this facade is then not needed.)
//
// Only what those sources touch is provided: small value types (Point, KeyPoint, DMatch, Matx, Vec, Affine3d), a
// ref-counted dense Mat with the handful of operations used (zeros / eye / at / row / clone / copyTo / t / * / +
// / ROI / comma initialiser), and the hot functions (implemented in facade/src/opencv_b200.cpp).  The cold helpers
// (Mat arithmetic on 3x3 / 3x4 matrices, Affine3d algebra, convertPointsFromHomogeneous, countNonZero, hconcat) are
// plain host code: they are microsecond-scale bookkeeping in the reference as well.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define MVO_CV_FACADE 1

struct mvo_ctx;

namespace cv {

class Exception : public std::runtime_error {
public:
  explicit Exception(const std::string& what) : std::runtime_error(what) {}
};

template <typename T> using Ptr = std::shared_ptr<T>;

template <typename T> struct Point_ {
  T x{}, y{};
  Point_() = default;
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T> struct Point3_ {
  T x{}, y{}, z{};
  Point3_() = default;
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
using Point2f = Point_<float>;
using Point2d = Point_<double>;
using Point3f = Point3_<float>;
using Point3d = Point3_<double>;

struct Rect {
  int x = 0, y = 0, width = 0, height = 0;
  Rect() = default;
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
  KeyPoint() = default;
  KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 3.4028235e38f;
  DMatch() = default;
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};

// ---- small fixed matrices ------------------------------------------------------------------------
template <typename T, int M, int N> struct Matx {
  T val[M * N];
  Matx() { for (int i = 0; i < M * N; ++i) val[i] = T(0); }
  static Matx eye() {
    Matx m;
    for (int i = 0; i < (M < N ? M : N); ++i) m(i, i) = T(1);
    return m;
  }
  static Matx zeros() { return Matx(); }
  T& operator()(int r, int c) { return val[r * N + c]; }
  const T& operator()(int r, int c) const { return val[r * N + c]; }
  Matx<T, N, M> t() const {
    Matx<T, N, M> r;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) r(j, i) = (*this)(i, j);
    return r;
  }
};
template <typename T, int M, int N> Matx<T, M, N> operator-(const Matx<T, M, N>& a, const Matx<T, M, N>& b) {
  Matx<T, M, N> r;
  for (int i = 0; i < M * N; ++i) r.val[i] = a.val[i] - b.val[i];
  return r;
}
template <typename T, int M, int N> Matx<T, M, N> operator+(const Matx<T, M, N>& a, const Matx<T, M, N>& b) {
  Matx<T, M, N> r;
  for (int i = 0; i < M * N; ++i) r.val[i] = a.val[i] + b.val[i];
  return r;
}
template <typename T, int M, int K, int N> Matx<T, M, N> operator*(const Matx<T, M, K>& a, const Matx<T, K, N>& b) {
  Matx<T, M, N> r;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      T s = T(0);
      for (int k = 0; k < K; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
template <typename T, int M, int N> std::ostream& operator<<(std::ostream& os, const Matx<T, M, N>& m) {
  os << "[";
  for (int i = 0; i < M; ++i) {
    for (int j = 0; j < N; ++j) os << m(i, j) << (j + 1 < N ? ", " : "");
    os << (i + 1 < M ? ";\n " : "]");
  }
  return os;
}
template <typename T, int M, int N> double norm(const Matx<T, M, N>& m) {
  double s = 0;
  for (int i = 0; i < M * N; ++i) s += (double)m.val[i] * (double)m.val[i];
  return std::sqrt(s);
}

template <typename T, int N> struct Vec : public Matx<T, N, 1> {
  Vec() = default;
  Vec(T a, T b, T c) {
    static_assert(N == 3, "three-component constructor");
    this->val[0] = a;
    this->val[1] = b;
    this->val[2] = c;
  }
  Vec(const Matx<T, N, 1>& m) : Matx<T, N, 1>(m) {}
  static Vec all(T v) {
    Vec r;
    for (int i = 0; i < N; ++i) r.val[i] = v;
    return r;
  }
  T& operator[](int i) { return this->val[i]; }
  const T& operator[](int i) const { return this->val[i]; }
};
using Matx33d = Matx<double, 3, 3>;
using Matx34d = Matx<double, 3, 4>;
using Matx44d = Matx<double, 4, 4>;
using Vec3d = Vec<double, 3>;
using Vec3f = Vec<float, 3>;

// ---- dense matrix ----------------------------------------------------------------------------------
template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };

class Mat {
public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  size_t step = 0;

  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  /// wraps external memory (no ownership), like cv::Mat(rows, cols, type, data, step)
  Mat(int r, int c, int type, void* external, size_t stp = 0) : rows(r), cols(c), type_(type) {
    data = static_cast<uchar*>(external);
    step = stp ? stp : (size_t)c * elemSize();
  }
  template <typename T, int M, int N> explicit Mat(const Matx<T, M, N>& m) {
    create(M, N, DataType<T>::type);
    std::memcpy(data, m.val, sizeof(T) * M * N);
  }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;   // OpenCV keeps a matching allocation
    rows = r;
    cols = c;
    type_ = type;
    step = (size_t)c * elemSize();
    store_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 16], std::default_delete<uchar[]>());
    data = store_.get();
    std::memset(data, 0, (size_t)r * step);
  }
  template <typename S> void create(S r, int c, int type) { create(static_cast<int>(r), c, type); }

  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat eye(int r, int c, int type);

  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize1() const { return depth() == CV_64F ? 8 : (depth() == CV_32F || depth() == CV_32S) ? 4 : 1; }
  size_t elemSize() const { return (size_t)channels() * elemSize1(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
  size_t total() const { return (size_t)rows * cols; }

  Mat row(int r) const {
    Mat m = *this;
    m.rows = 1;
    m.data = data + (size_t)r * step;
    return m;
  }
  Mat operator()(const Rect& roi) const {
    Mat m = *this;
    m.rows = roi.height;
    m.cols = roi.width;
    m.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
    return m;
  }
  Mat clone() const {
    Mat m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
    return m;
  }
  /// copies into dst; a destination of matching size and type (e.g. a row view) is written in place
  void copyTo(Mat dst) const {
    if (dst.rows != rows || dst.cols != cols || dst.type() != type_) throw Exception("Mat::copyTo: this facade needs a pre-sized destination view");
    for (int r = 0; r < rows; ++r) std::memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
  }
  Mat t() const;

  template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  /// element i of a single-row / single-column (or continuous) matrix
  template <typename T> T& at(int i) {
    return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols);
  }
  template <typename T> const T& at(int i) const {
    return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols);
  }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }

  /// 3x1 / 1x3 CV_64F matrix -> Vec3d (cv::Affine3d(R, t) with a cv::Mat t relies on this conversion)
  template <typename T, int N> operator Vec<T, N>() const {
    if ((int)total() != N) throw Exception("Mat -> Vec: size mismatch");
    Vec<T, N> v;
    for (int i = 0; i < N; ++i) v[i] = (T)getd(i / cols, i % cols);
    return v;
  }
  template <typename T, int M, int N> operator Matx<T, M, N>() const {
    if (rows != M || cols != N) throw Exception("Mat -> Matx: size mismatch");
    Matx<T, M, N> m;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) m(i, j) = (T)getd(i, j);
    return m;
  }
  /// numeric read of a single-channel element as double (CV_64F / CV_32F / CV_32S / CV_8U)
  double getd(int r, int c) const {
    switch (depth()) {
      case CV_64F: return at<double>(r, c);
      case CV_32F: return at<float>(r, c);
      case CV_32S: return at<int>(r, c);
      default: return at<uchar>(r, c);
    }
  }

private:
  int type_ = 0;
  std::shared_ptr<uchar> store_;
};

Mat operator*(const Mat& a, const Mat& b);   // CV_64F matrix product
Mat operator+(const Mat& a, const Mat& b);   // CV_64F element-wise sum
std::ostream& operator<<(std::ostream& os, const Mat& m);

template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_ {
public:
  MatCommaInitializer_(Mat_<T>* m, T first);
  MatCommaInitializer_& operator,(T v);
  operator Mat() const;
private:
  Mat m_;
  int i_ = 0;
};
template <typename T> class Mat_ : public Mat {
public:
  Mat_() = default;
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  T& operator()(int r, int c) { return this->template at<T>(r, c); }
  const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
  MatCommaInitializer_<T> operator<<(T v) { return MatCommaInitializer_<T>(this, v); }
};
template <typename T> MatCommaInitializer_<T>::MatCommaInitializer_(Mat_<T>* m, T first) : m_(*m) {
  m_.at<T>(0) = first;
  i_ = 1;
}
template <typename T> MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(T v) {
  if (i_ >= (int)m_.total()) throw Exception("Mat_ comma initialiser: too many values");
  m_.at<T>(i_ / m_.cols, i_ % m_.cols) = v;
  ++i_;
  return *this;
}
template <typename T> MatCommaInitializer_<T>::operator Mat() const { return m_; }

// ---- rigid transform ---------------------------------------------------------------------------------
template <typename T> class Affine3 {
public:
  using Mat3 = Matx<T, 3, 3>;
  using Mat4 = Matx<T, 4, 4>;
  using Vec3 = Vec<T, 3>;
  Mat4 matrix;

  Affine3() : matrix(Mat4::eye()) {}
  Affine3(const Mat3& R, const Vec3& t = Vec3::all(0)) : matrix(Mat4::eye()) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) matrix(i, j) = R(i, j);
      matrix(i, 3) = t[i];
    }
  }
  /// 3x3 rotation matrix (the only cv::Mat form the reference passes) + translation
  explicit Affine3(const Mat& data, const Vec3& t = Vec3::all(0)) : Affine3(static_cast<Mat3>(data), t) {}
  static Affine3 Identity() { return Affine3(); }

  Mat3 rotation() const {
    Mat3 R;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R(i, j) = matrix(i, j);
    return R;
  }
  Vec3 translation() const { return Vec3(matrix(0, 3), matrix(1, 3), matrix(2, 3)); }
  /// rigid inverse is not assumed by OpenCV (it inverts the 4x4 numerically); for the rotations + translations
  /// the reference builds the closed form [R^T | -R^T t] is the same transform
  Affine3 inv() const {
    const Mat3 Rt = rotation().t();
    const Vec3 t = translation();
    Vec3 ti;
    for (int i = 0; i < 3; ++i) ti[i] = -(Rt(i, 0) * t[0] + Rt(i, 1) * t[1] + Rt(i, 2) * t[2]);
    return Affine3(Rt, ti);
  }
};
template <typename T> Affine3<T> operator*(const Affine3<T>& a, const Affine3<T>& b) {
  Affine3<T> r;
  r.matrix = a.matrix * b.matrix;
  return r;
}
template <typename T, typename P> Point3_<P> operator*(const Affine3<T>& a, const Point3_<P>& p) {
  const auto& m = a.matrix;
  return Point3_<P>((P)(m(0, 0) * p.x + m(0, 1) * p.y + m(0, 2) * p.z + m(0, 3)),
                    (P)(m(1, 0) * p.x + m(1, 1) * p.y + m(1, 2) * p.z + m(1, 3)),
                    (P)(m(2, 0) * p.x + m(2, 1) * p.y + m(2, 2) * p.z + m(2, 3)));
}
using Affine3d = Affine3<double>;

// ---- cold helpers ---------------------------------------------------------------------------------------
struct NoArray {};
inline NoArray noArray() { return NoArray(); }
int countNonZero(const std::vector<uchar>& v);
int countNonZero(const Mat& m);
void hconcat(const Mat& a, const Mat& b, Mat& dst);
/// src: N x 4 (or N x 3) single-channel float matrix of homogeneous points; dst: N x 1 CV_32FC3
void convertPointsFromHomogeneous(const Mat& src, Mat& dst);
void convertPointsFromHomogeneous(const Mat& src, std::vector<Point3f>& dst);
double norm(const Mat& m);

// ---- features2d: cv::ORB (src/feature_processor.cpp:5-23) and cv::BFMatcher (:25-40) ---------------------
enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };

class ORB {
public:
  /// cv::ORB::create(nfeatures) -- every other parameter keeps OpenCV's default, as in the reference
  static Ptr<ORB> create(int nfeatures = 500);
  void detect(const Mat& image, std::vector<KeyPoint>& keypoints);
  void detectAndCompute(const Mat& image, NoArray mask, std::vector<KeyPoint>& keypoints, Mat& descriptors);
private:
  explicit ORB(int n) : nfeatures_(n) {}
  void run(const Mat& image, std::vector<KeyPoint>& keypoints, Mat* descriptors);
  int nfeatures_;
};

class BFMatcher {
public:
  explicit BFMatcher(int normType = NORM_L2, bool crossCheck = false);
  /// k = 2 nearest train rows per query row by Hamming distance, ties to the lower train index
  void knnMatch(const Mat& queryDescriptors, const Mat& trainDescriptors, std::vector<std::vector<DMatch>>& matches, int k) const;
private:
  int norm_;
};

// ---- video: cv::calcOpticalFlowPyrLK with all defaults (src/tracker.cpp:68-69) --------------------------
void calcOpticalFlowPyrLK(const Mat& prevImg, const Mat& nextImg, const std::vector<Point2f>& prevPts,
                          std::vector<Point2f>& nextPts, std::vector<uchar>& status, std::vector<float>& err);

// ---- calib3d ----------------------------------------------------------------------------------------------
enum { LMEDS = 4, RANSAC = 8 };
enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 };

/// src/initializer.cpp:82, src/tracker.cpp:243
Mat findHomography(const std::vector<Point2f>& srcPoints, const std::vector<Point2f>& dstPoints, int method,
                   double ransacReprojThreshold, std::vector<uchar>& mask);
/// src/initializer.cpp:87, src/tracker.cpp:248
Mat findFundamentalMat(const std::vector<Point2f>& points1, const std::vector<Point2f>& points2, int method,
                       double ransacReprojThreshold, double confidence, std::vector<uchar>& mask);
/// src/initializer.cpp:228-229
Mat findEssentialMat(const std::vector<Point2f>& points1, const std::vector<Point2f>& points2, const Mat& cameraMatrix,
                     int method, double prob, double threshold, std::vector<uchar>& mask);
/// src/initializer.cpp:236
int recoverPose(const Mat& E, const std::vector<Point2f>& points1, const std::vector<Point2f>& points2,
                const Mat& cameraMatrix, Mat& R, Mat& t, std::vector<uchar>& mask);
/// src/initializer.cpp:125, src/tracker.cpp:149 -- points4D: 4 x N CV_32F
void triangulatePoints(const Mat& projMatr1, const Mat& projMatr2, const std::vector<Point2f>& projPoints1,
                       const std::vector<Point2f>& projPoints2, Mat& points4D);
/// src/tracker.cpp:309 -- inliers: N_in x 1 CV_32S
bool solvePnPRansac(const std::vector<Point3f>& objectPoints, const std::vector<Point2f>& imagePoints,
                    const Mat& cameraMatrix, const Mat& distCoeffs, Mat& rvec, Mat& tvec, bool useExtrinsicGuess,
                    int iterationsCount, float reprojectionError, double confidence, Mat& inliers);
/// src/tracker.cpp:315
void Rodrigues(const Mat& src, Mat& dst);

// ---- the B200 context behind the facade --------------------------------------------------------------------
namespace b200 {
/// The process-wide stream context the hot functions run on: created on first use, grown when a larger image or
/// another ORB feature count arrives.  There is no CPU fallback: without libmonovo_b200.so / a CUDA device this throws.
mvo_ctx* context(int min_width = 0, int min_height = 0, int nfeatures = 0);
/// kernels launched so far by the facade's context (evidence that the hot path ran on the GPU)
unsigned long long launch_count();
void shutdown();
}  // namespace b200

}  // namespace cv
