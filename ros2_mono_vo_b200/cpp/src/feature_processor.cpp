// mono_vo::FeatureProcessor on top of the C ABI of libmonovo_b200.so (host side of the drop-in boundary).
#include "mono_vo/feature_processor.hpp"

#include <algorithm>
#include <stdexcept>
#include <string>

#include "monovo_b200.h"

namespace mono_vo
{
namespace
{
[[noreturn]] void fail(mvo_ctx * c, const char * what)
{
  throw std::runtime_error(std::string(what) + ": " + mvo_last_error(c));
}
static_assert(sizeof(mvo_keypoint) == 28, "mvo_keypoint must stay 7 x 4 bytes");
}  // namespace

FeatureProcessor::FeatureProcessor(int num_features, rclcpp::Logger logger)
: num_features_(num_features), logger_(logger)
{
}

FeatureProcessor::~FeatureProcessor()
{
  if (ctx_) mvo_destroy(ctx_);
}

mvo_ctx * FeatureProcessor::context(int width, int height) const
{
  (void)width;
  (void)height;
  if (ctx_) return ctx_;
  // One context for the life of the node.  max_width / max_height are only the limits the library checks image
  // sizes against (its buffers are sized for the images that actually arrive), so they are set to the ABI's maximum:
  // a larger image never destroys and re-creates the context (and with it the cached pyramids / descriptor blocks).
  mvo_config cfg{};
  cfg.device = 0;
  cfg.max_width = 16384;
  cfg.max_height = 16384;
  cfg.nfeatures = num_features_;
  cfg.batch = 1;
  if (mvo_create(&ctx_, &cfg) != MVO_OK) fail(nullptr, "FeatureProcessor: mvo_create");
  ctx_w_ = cfg.max_width;
  ctx_h_ = cfg.max_height;
  return ctx_;
}

mvo_ctx * FeatureProcessor::context() const
{
  return context(0, 0);
}

std::vector<cv::KeyPoint> FeatureProcessor::detect(const cv::Mat & image) const
{
  std::vector<cv::KeyPoint> keypoints;
  mvo_ctx * c = context(image.cols, image.rows);
  const int cap = num_features_ + num_features_ / 4 + 64;
  std::vector<mvo_keypoint> & kps = kp_scratch_;      // capacity-sized scratch lives with the object, not per call
  kps.resize(cap);
  int n = 0;
  if (mvo_orb_detect_and_compute(c, image.data, image.cols, image.rows, static_cast<int>(image.step),
      image.channels(), kps.data(), nullptr, cap, &n) != MVO_OK) fail(c, "FeatureProcessor::detect");
  keypoints.reserve(n);
  for (int i = 0; i < n; ++i) {
    const mvo_keypoint & k = kps[i];
    keypoints.emplace_back(cv::Point2f(k.x, k.y), k.size, k.angle, k.response, k.octave, k.class_id);
  }
  return keypoints;
}

void FeatureProcessor::detect_and_compute(
  const cv::Mat & image, std::vector<cv::KeyPoint> & keypoints, cv::Mat & descriptors) const
{
  mvo_ctx * c = context(image.cols, image.rows);
  const int cap = num_features_ + num_features_ / 4 + 64;
  std::vector<mvo_keypoint> & kps = kp_scratch_;
  kps.resize(cap);
  desc_scratch_.resize(static_cast<size_t>(cap) * 32);
  int n = 0;
  if (mvo_orb_detect_and_compute(c, image.data, image.cols, image.rows, static_cast<int>(image.step),
      image.channels(), kps.data(), desc_scratch_.data(), cap, &n) != MVO_OK) fail(c, "FeatureProcessor::detect_and_compute");
  keypoints.clear();
  keypoints.reserve(n);
  for (int i = 0; i < n; ++i) {
    const mvo_keypoint & k = kps[i];
    keypoints.emplace_back(cv::Point2f(k.x, k.y), k.size, k.angle, k.response, k.octave, k.class_id);
  }
  descriptors.create(n, 32, CV_8UC1);
  if (n > 0) {
    if (static_cast<size_t>(descriptors.step) == 32) {     // a fresh n x 32 CV_8UC1 Mat is dense: one copy
      std::copy(desc_scratch_.begin(), desc_scratch_.begin() + static_cast<size_t>(n) * 32, descriptors.ptr<unsigned char>(0));
    } else {
      for (int i = 0; i < n; ++i)
        std::copy(desc_scratch_.begin() + static_cast<size_t>(i) * 32, desc_scratch_.begin() + static_cast<size_t>(i + 1) * 32,
                  descriptors.ptr<unsigned char>(i));
    }
  }
}

std::vector<cv::DMatch> FeatureProcessor::find_matches(
  const cv::Mat & descriptors1, const cv::Mat & descriptors2, double lowes_distance_ratio) const
{
  mvo_ctx * c = context();
  const int nq = descriptors1.rows, nt = descriptors2.rows;
  // the ABI wants dense N x 32 rows; Frame::get_descriptors (src/frame.cpp:50-64) produces exactly that
  auto dense = [](const cv::Mat & m) {
    return (m.rows <= 1 || m.step == 32) ? m : m.clone();
  };
  const cv::Mat q = dense(descriptors1), t = dense(descriptors2);
  std::vector<mvo_dmatch> out(std::max(nq, 1));
  int n = 0;
  if (mvo_knn_ratio(c, q.data, nq, t.data, nt, lowes_distance_ratio, out.data(), &n) != MVO_OK)
    fail(c, "FeatureProcessor::find_matches");
  RCLCPP_INFO(logger_, "total matches: %d", nq);
  std::vector<cv::DMatch> good_matches(n);
  for (int i = 0; i < n; ++i) {
    good_matches[i].queryIdx = out[i].query_idx;
    good_matches[i].trainIdx = out[i].train_idx;
    good_matches[i].imgIdx = out[i].img_idx;
    good_matches[i].distance = out[i].distance;
  }
  RCLCPP_INFO(logger_, "good matches: %d", n);
  return good_matches;
}
}  // namespace mono_vo
