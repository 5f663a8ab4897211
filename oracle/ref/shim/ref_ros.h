// ORACLE / reference pin (test infrastructure only). ROS logging macros as no-ops, ros::Time's
// nanosecond conversion, and the message structs the extracted text reads.
#ifndef MML_REF_ROS_H
#define MML_REF_ROS_H
#include <cstdint>
#include <memory>
#include <vector>
#define ROS_WARN_STREAM(x) do {} while (0)
#define ROS_INFO_STREAM(x) do {} while (0)
#define ROS_ERROR_STREAM(x) do {} while (0)
#define ROS_WARN(...) do {} while (0)
#define ROS_INFO(...) do {} while (0)
#define ROS_ASSERT(x) do {} while (0)
namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time& fromNSec(uint64_t t) { sec = (uint32_t)(t / 1000000000ull); nsec = (uint32_t)(t % 1000000000ull); return *this; }
  Time& fromSec(double t) { sec = (uint32_t)std::floor(t); nsec = (uint32_t)std::round((t - sec) * 1e9); return *this; }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
};
}  // namespace ros
namespace livox_ros_driver {
struct CustomPoint { uint32_t offset_time; float x, y, z; uint8_t reflectivity, tag, line; };
struct CustomMsg { std::vector<CustomPoint> points; uint32_t point_num = 0; };
using CustomMsgConstPtr = std::shared_ptr<const CustomMsg>;
}  // namespace livox_ros_driver
namespace sensor_msgs {
struct Imu {
  struct { ros::Time stamp; } header;
  struct V3 { double x = 0, y = 0, z = 0; } angular_velocity, linear_acceleration;
};
using ImuConstPtr = std::shared_ptr<const Imu>;
}  // namespace sensor_msgs
#endif
