#pragma once
#include <string>
namespace rcl_interfaces { namespace msg {
struct ParameterDescriptor { std::string description; };
} }
