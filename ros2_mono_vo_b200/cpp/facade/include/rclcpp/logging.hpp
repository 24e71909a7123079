// rclcpp/logging.hpp -- the part of rclcpp's logging the reference's hot-path classes touch (Logger values and the
// printf-style RCLCPP_* macros), for the facade build without ROS 2.  Messages go to stderr when MVO_FACADE_LOG is set.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace rclcpp {
class Logger {
public:
  explicit Logger(std::string n = "") : name_(std::move(n)) {}
  const char* get_name() const { return name_.c_str(); }
  Logger get_child(const std::string& suffix) const { return Logger(name_.empty() ? suffix : name_ + "." + suffix); }
private:
  std::string name_;
};
inline Logger get_logger(const std::string& name) { return Logger(name); }

namespace detail {
inline bool log_enabled() {
  static const bool on = std::getenv("MVO_FACADE_LOG") != nullptr;
  return on;
}
__attribute__((format(printf, 3, 4))) inline void log(const Logger& lg, const char* level, const char* fmt, ...) {
  if (!log_enabled()) return;
  std::fprintf(stderr, "[%s] [%s]: ", level, lg.get_name());
  va_list ap;
  va_start(ap, fmt);
  std::vfprintf(stderr, fmt, ap);
  va_end(ap);
  std::fputc('\n', stderr);
}
}  // namespace detail
}  // namespace rclcpp

#define RCLCPP_DEBUG(logger, ...) ::rclcpp::detail::log(logger, "DEBUG", __VA_ARGS__)
#define RCLCPP_INFO(logger, ...) ::rclcpp::detail::log(logger, "INFO", __VA_ARGS__)
#define RCLCPP_WARN(logger, ...) ::rclcpp::detail::log(logger, "WARN", __VA_ARGS__)
#define RCLCPP_ERROR(logger, ...) ::rclcpp::detail::log(logger, "ERROR", __VA_ARGS__)
