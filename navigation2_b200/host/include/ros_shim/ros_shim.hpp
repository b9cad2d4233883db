// ros_shim.hpp — the smallest stand-ins for the ROS 2 / Nav2 types the MPPI controller plugin
// touches, for building and testing the host mirror where ROS 2 is not installed (this container:
// no /opt/ros, SURVEY.md "Facts").  With a real ROS 2 underlay these come from geometry_msgs,
// nav_msgs, nav2_core, nav2_costmap_2d and nav2_ros_common instead; the names,
// members and method signatures below are the reference's.
#ifndef ROS_SHIM__ROS_SHIM_HPP_
#define ROS_SHIM__ROS_SHIM_HPP_

#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace builtin_interfaces::msg {struct Time {int32_t sec{0}; uint32_t nanosec{0};};}
namespace std_msgs::msg {struct Header {builtin_interfaces::msg::Time stamp; std::string frame_id;};}

namespace geometry_msgs::msg
{
struct Point {double x{0}, y{0}, z{0};};
struct Quaternion {double x{0}, y{0}, z{0}, w{1};};
struct Vector3 {double x{0}, y{0}, z{0};};
struct Pose {Point position; Quaternion orientation;};
struct PoseStamped {std_msgs::msg::Header header; Pose pose;};
struct Twist {Vector3 linear; Vector3 angular;};
struct TwistStamped {std_msgs::msg::Header header; Twist twist;};
}  // namespace geometry_msgs::msg

namespace nav_msgs::msg
{
struct Path {std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses;};
// nav_msgs/msg/TrajectoryPoint.msg, nav_msgs/msg/Trajectory.msg (what "~/optimal_trajectory" carries)
struct TrajectoryPoint {std_msgs::msg::Header header; geometry_msgs::msg::Pose pose; geometry_msgs::msg::Twist velocity;};
struct Trajectory {std_msgs::msg::Header header; std::vector<TrajectoryPoint> points;};
}  // namespace nav_msgs::msg

namespace std_msgs::msg {struct ColorRGBA {float r{0}, g{0}, b{0}, a{0};};}
namespace visualization_msgs::msg
{
struct Marker
{
  static constexpr int32_t SPHERE = 2, LINE_STRIP = 4;
  static constexpr int32_t ADD = 0;
  std_msgs::msg::Header header;
  std::string ns;
  int32_t id{0};
  int32_t type{0};
  int32_t action{0};
  geometry_msgs::msg::Pose pose;
  geometry_msgs::msg::Vector3 scale;
  std_msgs::msg::ColorRGBA color;
  std::vector<geometry_msgs::msg::Point> points;
};
struct MarkerArray {std::vector<Marker> markers;};
}  // namespace visualization_msgs::msg
// nav2_msgs/msg/CriticsStats.msg
namespace nav2_msgs::msg
{
struct CriticsStats
{
  builtin_interfaces::msg::Time stamp;
  std::vector<std::string> critics;
  std::vector<bool> changed;
  std::vector<float> costs_sum;
};
}  // namespace nav2_msgs::msg

namespace tf2
{
// tf2::getYaw(quaternion) (geometry2 tf2/utils.h): yaw of a unit quaternion
inline double getYaw(const geometry_msgs::msg::Quaternion & q)
{
  const double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
  const double sarg = -2.0 * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
  if (sarg <= -0.99999) {return -2.0 * std::atan2(q.y, q.x);}
  if (sarg >= 0.99999) {return 2.0 * std::atan2(q.y, q.x);}
  return std::atan2(2.0 * (q.x * q.y + q.w * q.z), sqw + sqx - sqy - sqz);
}
inline geometry_msgs::msg::Quaternion quaternionFromYaw(double yaw)
{
  geometry_msgs::msg::Quaternion q;
  q.x = 0; q.y = 0; q.z = std::sin(yaw * 0.5); q.w = std::cos(yaw * 0.5);
  return q;
}
}  // namespace tf2

namespace nav2_core
{
// nav2_core/include/nav2_core/controller_exceptions.hpp:23-77
class ControllerException : public std::runtime_error
{
public:
  explicit ControllerException(const std::string & description) : std::runtime_error(description) {}
};
class NoValidControl : public ControllerException
{
public:
  explicit NoValidControl(const std::string & description) : ControllerException(description) {}
};
class InvalidPath : public ControllerException
{
public:
  explicit InvalidPath(const std::string & description) : ControllerException(description) {}
};
class ControllerTFError : public ControllerException
{
public:
  explicit ControllerTFError(const std::string & description) : ControllerException(description) {}
};

// nav2_core/include/nav2_core/goal_checker.hpp:66-124 (the member the MPPI critics call)
class GoalChecker
{
public:
  virtual ~GoalChecker() = default;
  virtual bool getTolerances(
    geometry_msgs::msg::Pose & pose_tolerance, geometry_msgs::msg::Twist & vel_tolerance,
    double & path_length_tolerance) = 0;
};
}  // namespace nav2_core

namespace nav2
{
using ParameterValue = std::variant<bool, int, double, std::string, std::vector<std::string>, std::vector<double>>;

// the slice of nav2::LifecycleNode the controller uses: a parameter store
class LifecycleNode
{
public:
  using WeakPtr = std::weak_ptr<LifecycleNode>;
  using SharedPtr = std::shared_ptr<LifecycleNode>;
  explicit LifecycleNode(const std::string & name) : name_(name) {}
  const std::string & get_name() const {return name_;}
  void declare_parameter(const std::string & name, const ParameterValue & value) {params_[name] = value;}
  void set_parameter(const std::string & name, const ParameterValue & value) {params_[name] = value;}
  bool has_parameter(const std::string & name) const {return params_.count(name) != 0;}
  const ParameterValue & get_parameter(const std::string & name) const {return params_.at(name);}

private:
  std::string name_;
  std::map<std::string, ParameterValue> params_;
};
// tf2_ros::Buffer stand-in: one planar transform per (target frame, source frame) pair; identity when none is set
struct TransformBuffer
{
  using SharedPtr = std::shared_ptr<TransformBuffer>;
  struct Planar {double x{0}, y{0}, yaw{0};};
  void setTransform(const std::string & target_frame, const std::string & source_frame, double x, double y, double yaw)
  {
    transforms_[target_frame + "<-" + source_frame] = Planar{x, y, yaw};
  }
  // pose expressed in source_frame -> target_frame (nav2_util::transformPoseInTargetFrame)
  bool transform(
    const geometry_msgs::msg::PoseStamped & in, geometry_msgs::msg::PoseStamped & out, const std::string & target_frame) const
  {
    out = in;
    out.header.frame_id = target_frame;
    if (in.header.frame_id == target_frame || in.header.frame_id.empty() || target_frame.empty()) {return true;}
    auto it = transforms_.find(target_frame + "<-" + in.header.frame_id);
    if (it == transforms_.end()) {return false;}
    const Planar & t = it->second;
    const double c = std::cos(t.yaw), s = std::sin(t.yaw);
    out.pose.position.x = t.x + c * in.pose.position.x - s * in.pose.position.y;
    out.pose.position.y = t.y + s * in.pose.position.x + c * in.pose.position.y;
    out.pose.orientation = tf2::quaternionFromYaw(tf2::getYaw(in.pose.orientation) + t.yaw);
    return true;
  }

private:
  std::map<std::string, Planar> transforms_;
};

// rclcpp_lifecycle::LifecyclePublisher stand-in: keeps the last message, counts what was published; subscribers are a number
template<typename MsgT>
class Publisher
{
public:
  explicit Publisher(std::string topic) : topic_(std::move(topic)) {}
  void on_activate() {active_ = true;}
  void on_deactivate() {active_ = false;}
  size_t get_subscription_count() const {return subscribers_;}
  void setSubscriptionCount(size_t n) {subscribers_ = n;}
  void publish(std::unique_ptr<MsgT> msg) {last_ = std::move(msg); ++published_;}
  const MsgT * last() const {return last_.get();}
  size_t publishedCount() const {return published_;}
  const std::string & topic() const {return topic_;}

private:
  std::string topic_;
  bool active_{false};
  size_t subscribers_{1};
  size_t published_{0};
  std::unique_ptr<MsgT> last_;
};
}  // namespace nav2

namespace nav2_costmap_2d
{
static constexpr unsigned char NO_INFORMATION = 255;
static constexpr unsigned char LETHAL_OBSTACLE = 254;
static constexpr unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;
static constexpr double NO_SPEED_LIMIT = 0.0;

// read side of nav2_costmap_2d::Costmap2D (costmap_2d.hpp:150-157,193,231-234,252,407-408)
class Costmap2D
{
public:
  using mutex_t = std::recursive_mutex;
  Costmap2D(unsigned int sx, unsigned int sy, double res, double ox, double oy, unsigned char def = 0)
  : size_x_(sx), size_y_(sy), resolution_(res), origin_x_(ox), origin_y_(oy), cells_(static_cast<size_t>(sx) * sy, def) {}
  unsigned char * getCharMap() {return cells_.data();}
  const unsigned char * getCharMap() const {return cells_.data();}
  unsigned int getSizeInCellsX() const {return size_x_;}
  unsigned int getSizeInCellsY() const {return size_y_;}
  double getResolution() const {return resolution_;}
  double getOriginX() const {return origin_x_;}
  double getOriginY() const {return origin_y_;}
  unsigned int getIndex(unsigned int mx, unsigned int my) const {return my * size_x_ + mx;}
  unsigned char getCost(unsigned int mx, unsigned int my) const {return cells_[getIndex(mx, my)];}
  void setCost(unsigned int mx, unsigned int my, unsigned char c) {cells_[getIndex(mx, my)] = c;}
  mutex_t * getMutex() {return &mutex_;}
  double getSizeInMetersX() const {return size_x_ * resolution_;}   // costmap_2d.cpp:558-566
  double getSizeInMetersY() const {return size_y_ * resolution_;}
  bool worldToMap(double wx, double wy, unsigned int & mx, unsigned int & my) const  // costmap_2d.cpp:292-305
  {
    if (wx < origin_x_ || wy < origin_y_) {return false;}
    mx = static_cast<unsigned int>((wx - origin_x_) / resolution_);
    my = static_cast<unsigned int>((wy - origin_y_) / resolution_);
    return mx < size_x_ && my < size_y_;
  }

private:
  unsigned int size_x_, size_y_;
  double resolution_, origin_x_, origin_y_;
  std::vector<unsigned char> cells_;
  mutex_t mutex_;
};

// what the MPPI critics ask LayeredCostmap / InflationLayer for
struct LayeredCostmap
{
  bool tracking_unknown{false};
  double inscribed_radius{0.0}, circumscribed_radius{0.0};
  bool isTrackingUnknown() const {return tracking_unknown;}
  double getInscribedRadius() const {return inscribed_radius;}
  double getCircumscribedRadius() const {return circumscribed_radius;}
};
struct InflationLayerInfo {bool present{false}; double cost_scaling_factor{0.0}, inflation_radius{0.0};};

class Costmap2DROS
{
public:
  Costmap2DROS(std::shared_ptr<Costmap2D> costmap, bool use_radius, std::vector<geometry_msgs::msg::Point> footprint)
  : costmap_(std::move(costmap)), use_radius_(use_radius), footprint_(std::move(footprint)) {}
  Costmap2D * getCostmap() {return costmap_.get();}
  bool getUseRadius() const {return use_radius_;}
  const std::vector<geometry_msgs::msg::Point> & getRobotFootprint() const {return footprint_;}
  LayeredCostmap * getLayeredCostmap() {return &layered_;}
  InflationLayerInfo & inflation() {return inflation_;}
  std::string getBaseFrameID() const {return "base_link";}
  std::string getGlobalFrameID() const {return "odom";}

private:
  std::shared_ptr<Costmap2D> costmap_;
  bool use_radius_;
  std::vector<geometry_msgs::msg::Point> footprint_;
  LayeredCostmap layered_;
  InflationLayerInfo inflation_;
};
}  // namespace nav2_costmap_2d

namespace nav2_core
{
// nav2_core/include/nav2_core/controller.hpp:59-148
class Controller
{
public:
  using Ptr = std::shared_ptr<Controller>;
  virtual ~Controller() = default;
  virtual void configure(
    const nav2::LifecycleNode::WeakPtr &, std::string name, nav2::TransformBuffer::SharedPtr,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS>) = 0;
  virtual void cleanup() = 0;
  virtual void activate() = 0;
  virtual void deactivate() = 0;
  virtual void newPathReceived(const nav_msgs::msg::Path & raw_global_path) = 0;
  virtual geometry_msgs::msg::TwistStamped computeVelocityCommands(
    const geometry_msgs::msg::PoseStamped & pose, const geometry_msgs::msg::Twist & velocity,
    nav2_core::GoalChecker * goal_checker, const nav_msgs::msg::Path & transformed_global_plan,
    const geometry_msgs::msg::PoseStamped & global_goal) = 0;
  virtual bool cancel() {return true;}
  virtual void setSpeedLimit(const double & speed_limit, const bool & percentage) = 0;
  virtual void reset() {}
};
}  // namespace nav2_core

#endif  // ROS_SHIM__ROS_SHIM_HPP_
