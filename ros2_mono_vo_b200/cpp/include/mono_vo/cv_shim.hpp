// cv_shim.hpp -- the handful of OpenCV / rclcpp types that cross the FeatureProcessor interface, for builds
// without OpenCV and ROS 2 (this container).  With the real libraries installed this header is not used:
// feature_processor.hpp includes <opencv2/opencv.hpp> and <rclcpp/logging.hpp> instead (__has_include).
// Only the members the reference touches at the hot-path boundary are provided (layout-free, value semantics).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define MVO_CV_SHIM 1
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)

namespace cv {

struct Point2f {
  float x = 0, y = 0;
  Point2f() = default;
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct Point3f {
  float x = 0, y = 0, z = 0;
  Point3f() = default;
  Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
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
  float distance = 0;
};

// minimal dense matrix: 8-bit (images, descriptors) or 64-bit float (3x3 / 3x4 / 3x1), ref-counted storage
class Mat {
public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* external, size_t stp = 0) : rows(r), cols(c), type_(type) {
    data = static_cast<uint8_t*>(external);
    step = stp ? stp : (size_t)c * elemSize();
  }
  void create(int r, int c, int type) {
    rows = r;
    cols = c;
    type_ = type;
    step = (size_t)c * elemSize();
    store_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * step + 1], std::default_delete<uint8_t[]>());
    data = store_.get();
    std::memset(data, 0, (size_t)r * step);
  }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return (size_t)channels() * (depth() == CV_64F ? 8 : depth() == CV_32F ? 4 : 1); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Mat row(int r) const {
    Mat m = *this;
    m.rows = 1;
    m.data = data + (size_t)r * step;
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
    return m;
  }
  template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }

private:
  int type_ = 0;
  std::shared_ptr<uint8_t> store_;
};

using uchar = unsigned char;
enum { RANSAC = 8, FM_RANSAC = 8 };

}  // namespace cv

namespace rclcpp {
class Logger {
public:
  explicit Logger(std::string n = "") : name_(std::move(n)) {}
  const char* get_name() const { return name_.c_str(); }
private:
  std::string name_;
};
inline Logger get_logger(const std::string& name) { return Logger(name); }
}  // namespace rclcpp
#ifndef RCLCPP_INFO
#define RCLCPP_INFO(logger, ...) ((void)(logger))
#define RCLCPP_WARN(logger, ...) ((void)(logger))
#endif
