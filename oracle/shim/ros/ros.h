// oracle/shim/ros/ros.h -- stand-in for the ROS client headers the reference's DDP translation unit
// includes only for wall-clock timing and console logging (ddp_optimizer.cpp:30,414-416).
// TEST INFRASTRUCTURE ONLY; used by `make -C oracle ref`.
#ifndef ORACLE_SHIM_ROS_H
#define ORACLE_SHIM_ROS_H
#include <chrono>
namespace ros {
struct Duration {
    double s;
    double toSec() const { return s; }
};
struct Time {
    double t;
    static Time now() {
        using namespace std::chrono;
        return Time{duration<double>(steady_clock::now().time_since_epoch()).count()};
    }
    Duration operator-(const Time &o) const { return Duration{t - o.t}; }
};
}  // namespace ros
#define ROS_WARN(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#endif
