#pragma once
#include <cstdint>
#include <string>
namespace sensor_msgs { namespace msg {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 0; uint32_t count = 0; };
} }
