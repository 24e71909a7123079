// rclcpp/rclcpp.hpp -- a parameter-holding stand-in for rclcpp::Node: exactly what mono_vo::RosParameterHandler
// (include/mono_vo/ros_parameter_handler.hpp) calls.  Overrides play the role of the parameter YAML
// (config/params.yaml: "initializer.lowes_distance_ratio: 0.7" ...).
#pragma once
#include <map>
#include <string>

#include "rcl_interfaces/msg/parameter_descriptor.hpp"
#include "rclcpp/logging.hpp"

namespace rclcpp {
class Node {
public:
  explicit Node(const std::string& name) : logger_(name) {}
  Logger get_logger() const { return logger_; }
  /// a value "from the YAML": wins over the default given to declare_parameter
  void set_override(const std::string& full_name, double v) { overrides_[full_name] = v; }
  template <typename T>
  T declare_parameter(const std::string& full_name, const T& default_value,
                      const rcl_interfaces::msg::ParameterDescriptor& = rcl_interfaces::msg::ParameterDescriptor()) {
    auto it = overrides_.find(full_name);
    values_[full_name] = it != overrides_.end() ? it->second : static_cast<double>(default_value);
    return static_cast<T>(values_[full_name]);
  }
  template <typename T> bool get_parameter(const std::string& full_name, T& out) const {
    auto it = values_.find(full_name);
    if (it == values_.end()) return false;
    out = static_cast<T>(it->second);
    return true;
  }
  const std::map<std::string, double>& parameters() const { return values_; }
private:
  Logger logger_;
  std::map<std::string, double> overrides_, values_;
};
}  // namespace rclcpp
