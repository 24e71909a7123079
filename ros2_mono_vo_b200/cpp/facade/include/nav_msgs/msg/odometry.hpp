#pragma once
#include "geometry_msgs/msg/pose_stamped.hpp"
namespace nav_msgs { namespace msg {
struct PoseWithCovariance { geometry_msgs::msg::Pose pose; double covariance[36] = {0}; };
struct Odometry { std_msgs::msg::Header header; std::string child_frame_id; PoseWithCovariance pose; };
} }
