// ros/ros.h stand-in (ROS is not installed in this image): enough of roscpp for the reference's offline drivers
// (gtsam/test_vro_imu_graph.cpp:62-71,476-514).  Private parameters come from the command line the way rosrun passes
// them (`_name:=value`) or from the environment (ROS_PARAM_<name>); ROS_INFO/WARN/ERROR print to stdout / stderr.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
namespace ros {
inline std::map<std::string, std::string>& param_table() { static std::map<std::string, std::string> t; return t; }
inline void init(int& argc, char** argv, const std::string& /*name*/, unsigned = 0) {
  for (int i = 1; i < argc; ++i) {
    const std::string a(argv[i]);
    const size_t p = a.find(":=");
    if (a.size() > 1 && a[0] == '_' && p != std::string::npos) param_table()[a.substr(1, p - 1)] = a.substr(p + 2);
  }
}
inline bool ok() { return true; }
inline void spinOnce() {}
inline void spin() {}
inline void shutdown() {}
class NodeHandle {
  static bool lookup(const std::string& name, std::string& out) {
    auto it = param_table().find(name);
    if (it != param_table().end()) { out = it->second; return true; }
    const char* e = std::getenv(("ROS_PARAM_" + name).c_str());
    if (e) { out = e; return true; }
    return false;
  }
  template <class T> static void parse(const std::string& s, T& v) { std::istringstream is(s); is >> v; }
  static void parse(const std::string& s, std::string& v) { v = s; }
  static void parse(const std::string& s, bool& v) { v = (s == "1" || s == "true" || s == "True" || s == "TRUE"); }
 public:
  NodeHandle(const std::string& = "") {}
  template <class T, class D> bool param(const std::string& name, T& var, const D& def) const {
    std::string s;
    if (lookup(name, s)) { parse(s, var); return true; }
    var = def;
    return false;
  }
  template <class T> bool getParam(const std::string& name, T& var) const { std::string s; if (!lookup(name, s)) return false; parse(s, var); return true; }
  template <class T> void setParam(const std::string& name, const T& v) const { std::ostringstream os; os << v; param_table()[name] = os.str(); }
  bool hasParam(const std::string& name) const { std::string s; return lookup(name, s); }
};
struct Time { double t = 0; static Time now() { return Time(); } double toSec() const { return t; } };
struct Rate { Rate(double) {} void sleep() {} };
}  // namespace ros
#define ROS_INFO(...) do { std::printf("[ INFO] "); std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_WARN(...) do { std::fprintf(stderr, "[ WARN] "); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ERROR] "); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_DEBUG(...) do { } while (0)
#define ROS_INFO_STREAM(x) do { std::ostringstream os_; os_ << x; std::printf("[ INFO] %s\n", os_.str().c_str()); } while (0)
#define ROS_WARN_STREAM(x) do { std::ostringstream os_; os_ << x; std::fprintf(stderr, "[ WARN] %s\n", os_.str().c_str()); } while (0)
#define ROS_ERROR_STREAM(x) do { std::ostringstream os_; os_ << x; std::fprintf(stderr, "[ERROR] %s\n", os_.str().c_str()); } while (0)
#define ROS_INFO_ONCE(...) ROS_INFO(__VA_ARGS__)
