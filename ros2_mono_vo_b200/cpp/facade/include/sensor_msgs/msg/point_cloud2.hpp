#pragma once
#include <vector>
#include "sensor_msgs/msg/point_field.hpp"
#include "std_msgs/msg/header.hpp"
namespace sensor_msgs { namespace msg {
struct PointCloud2 {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
} }
