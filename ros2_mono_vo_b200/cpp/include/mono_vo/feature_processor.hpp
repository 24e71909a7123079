// B200-native drop-in for mono_vo::FeatureProcessor.
//
// Interface contract: the public members below are source compatible with the reference's class
// (/root/reference/include/mono_vo/feature_processor.hpp:14-31): same constructor defaults, same const methods,
// same argument meaning and container types, so Frame::extract_observations (src/frame.cpp:8-17),
// Initializer (src/initializer.cpp:187) and Tracker (src/tracker.cpp:186-191) compile against it unchanged.
// The implementation holds no cv::ORB / cv::BFMatcher: it forwards to libmonovo_b200.so through the C ABI of
// include/monovo_b200.h.  There is no CPU fallback: without the library or a CUDA device the calls throw.
#pragma once

#include <memory>
#include <vector>
#if __has_include(<opencv2/opencv.hpp>) && !defined(MVO_FORCE_CV_SHIM)
#include <opencv2/opencv.hpp>
#else
#include "mono_vo/cv_shim.hpp"
#endif
#if __has_include(<rclcpp/logging.hpp>) && !defined(MVO_FORCE_CV_SHIM)
#include <rclcpp/logging.hpp>
#endif

#include "monovo_b200.h"   // mvo_ctx, mvo_keypoint (the scratch members below)

namespace mono_vo
{
class FeatureProcessor
{
public:
  using Ptr = std::shared_ptr<FeatureProcessor>;

  FeatureProcessor(
    int num_features = 1000, rclcpp::Logger logger = rclcpp::get_logger("FeatureProcessor"));
  ~FeatureProcessor();
  FeatureProcessor(const FeatureProcessor &) = delete;
  FeatureProcessor & operator=(const FeatureProcessor &) = delete;

  /// ORB keypoints only (cv::ORB::detect).
  std::vector<cv::KeyPoint> detect(const cv::Mat & image) const;

  /// ORB keypoints + 32-byte rBRIEF descriptors (cv::ORB::detectAndCompute with noArray() mask).
  /// descriptors becomes an N x 32 CV_8U matrix; the keypoint SET, responses, angles and descriptors equal
  /// cv::ORB's bit for bit, the order is canonical (octave, response desc, y, x).
  void detect_and_compute(
    const cv::Mat & image, std::vector<cv::KeyPoint> & keypoints, cv::Mat & descriptors) const;

  /// Brute-force Hamming k=2 matching + Lowe ratio test: keeps the best match of query row i iff two
  /// neighbours exist and d0 < lowes_distance_ratio * d1 (float < double * float, evaluated in double).
  std::vector<cv::DMatch> find_matches(
    const cv::Mat & descriptors1, const cv::Mat & descriptors2, double lowes_distance_ratio) const;

  /// The stream context (for the geometry / LK entry points in mono_vo/gpu_cv.hpp); created on first use.
  mvo_ctx * context(int width, int height) const;
  /// The same context; for call sites that have no image at hand.
  mvo_ctx * context() const;

private:
  int num_features_;
  rclcpp::Logger logger_;
  mutable mvo_ctx * ctx_ = nullptr;
  mutable int ctx_w_ = 0, ctx_h_ = 0;
  // capacity-sized staging of detect / detect_and_compute (methods are const like the reference's: mutable scratch;
  // like the reference's cv::Ptr<cv::ORB>, one FeatureProcessor is used from one thread)
  mutable std::vector<mvo_keypoint> kp_scratch_;
  mutable std::vector<unsigned char> desc_scratch_;
};
}  // namespace mono_vo
