// vo_host_main.cpp -- drives the reference's own Initializer / Tracker / Map / Frame code (compiled UNCHANGED from
// /root/reference/src by oracle/build_ref_host.py against the OpenCV-API facade) over a recorded image sequence, the way
// MonoVO::image_callback does (/root/reference/src/mono_vo.cpp:83-153, minus the ROS publishing), and prints one JSON
// line per frame.  Every hot call those classes make (ORB, knnMatch, LK, H / F / E RANSAC, recoverPose, triangulation,
// solvePnPRansac) runs on the GPU through libmonovo_b200.so.
//
//   mono_vo_host --seq frames.bin [--params params.txt] [--nfeatures 1000]
//
// frames.bin: "MVOSEQ1\0", int32 n, h, w, channels, double K[9], then n * h * w * channels bytes.
// params.txt: lines "initializer.lowes_distance_ratio 0.7" (the node's parameter YAML, flattened).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <optional>
#include <string>
#include <vector>

#include "mono_vo/feature_processor.hpp"
#include "mono_vo/frame.hpp"
#include "mono_vo/initializer.hpp"
#include "mono_vo/map.hpp"
#include "mono_vo/ros_parameter_handler.hpp"
#include "mono_vo/tracker.hpp"

int main(int argc, char** argv)
{
  std::string seq_path, params_path;
  int nfeatures = 1000;   // /root/reference/src/mono_vo.cpp:16
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--seq" && i + 1 < argc) seq_path = argv[++i];
    else if (a == "--params" && i + 1 < argc) params_path = argv[++i];
    else if (a == "--nfeatures" && i + 1 < argc) nfeatures = std::atoi(argv[++i]);
    else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (seq_path.empty()) { std::fprintf(stderr, "usage: mono_vo_host --seq frames.bin [--params params.txt] [--nfeatures N]\n"); return 2; }

  std::ifstream in(seq_path, std::ios::binary);
  char magic[8] = {0};
  int32_t hdr[4] = {0, 0, 0, 0};
  double k[9];
  in.read(magic, 8);
  in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
  in.read(reinterpret_cast<char*>(k), sizeof(k));
  if (!in || std::memcmp(magic, "MVOSEQ1", 8) != 0) { std::fprintf(stderr, "bad sequence file\n"); return 2; }
  const int n = hdr[0], h = hdr[1], w = hdr[2], cn = hdr[3];
  const size_t fbytes = static_cast<size_t>(h) * w * cn;
  std::vector<unsigned char> pixels(fbytes * n);
  in.read(reinterpret_cast<char*>(pixels.data()), static_cast<std::streamsize>(pixels.size()));
  if (!in) { std::fprintf(stderr, "truncated sequence file\n"); return 2; }

  cv::Mat K(3, 3, CV_64F), d = cv::Mat::zeros(1, 5, CV_64F);
  for (int i = 0; i < 9; ++i) K.at<double>(i / 3, i % 3) = k[i];

  rclcpp::Node node("mono_vo");
  if (!params_path.empty()) {
    std::ifstream pf(params_path);
    std::string name;
    double value;
    while (pf >> name >> value) node.set_override(name, value);
  }

  try {
    auto map = std::make_shared<mono_vo::Map>(node.get_logger().get_child("map"));
    auto fp = std::make_shared<mono_vo::FeatureProcessor>(nfeatures, node.get_logger().get_child("feature_processor"));
    mono_vo::Initializer initializer(map, fp, node.get_logger().get_child("initializer"));
    mono_vo::Tracker tracker(map, fp, node.get_logger().get_child("tracker"));
    auto ih = mono_vo::RosParameterHandler(&node, "initializer");
    initializer.configure_parameters(ih);
    auto th = mono_vo::RosParameterHandler(&node, "tracker");
    tracker.configure_parameters(th);
    for (const auto& p : node.parameters()) std::printf("{\"param\": \"%s\", \"value\": %.9g}\n", p.first.c_str(), p.second);

    for (int f = 0; f < n; ++f) {
      const auto t0 = std::chrono::steady_clock::now();
      cv::Mat image(h, w, cn == 3 ? CV_8UC3 : CV_8UC1, pixels.data() + fbytes * f);
      mono_vo::Frame frame{image};
      std::optional<cv::Affine3d> pose;
      bool init_event = false;
      if (!initializer.is_initalized()) {
        std::optional<mono_vo::Frame> ref_frame = initializer.try_initializing(frame, K);
        if (ref_frame.has_value()) {
          tracker.update(ref_frame.value(), K, d);
          init_event = true;
          pose = ref_frame.value().pose_wc;
        }
      } else {
        pose = tracker.update(frame, K, d);
      }
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      std::printf("{\"frame\": %d, \"initialized\": %s, \"init_event\": %s, \"tracker_state\": %d, \"landmarks\": %zu, "
                  "\"keyframes\": %zu, \"ms\": %.3f, \"launches\": %llu, \"pose_wc\": ",
                  f, initializer.is_initalized() ? "true" : "false", init_event ? "true" : "false",
                  static_cast<int>(tracker.get_state()), map->num_landmarks(), map->num_keyframes(), ms, cv::b200::launch_count());
      if (pose.has_value()) {
        std::printf("[");
        for (int i = 0; i < 16; ++i) std::printf("%.12g%s", pose->matrix.val[i], i < 15 ? ", " : "]");
      } else {
        std::printf("null");
      }
      std::printf("}\n");
    }
  } catch (const std::exception& e) {
    std::printf("{\"error\": \"%s\"}\n", e.what());
    cv::b200::shutdown();
    return 1;
  }
  cv::b200::shutdown();
  return 0;
}
