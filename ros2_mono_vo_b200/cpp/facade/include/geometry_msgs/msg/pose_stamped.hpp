// message stand-ins: only so that include/mono_vo/utils.hpp (declarations of the ROS conversion helpers, which
// initializer.hpp / tracker.hpp include) parses; src/utils.cpp and the node are not part of the facade build.
#pragma once
#include "std_msgs/msg/header.hpp"
namespace geometry_msgs { namespace msg {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::msg::Header header; Pose pose; };
} }
