// mppi.hpp — C++ host mirror of nav2_mppi_controller's plugin surfaces on top of the C ABI
// (include/mppi_b200.h).  Class names, method names, argument meaning, ROS parameter names and
// error behaviour are the reference's:
//   nav2_mppi_controller::MPPIController   include/nav2_mppi_controller/controller.hpp:39-131
//   mppi::Optimizer                        include/nav2_mppi_controller/optimizer.hpp:59-333
//   mppi::ParametersHandler                include/nav2_mppi_controller/tools/parameters_handler.hpp:42-308
//   mppi::MotionModel (+ DiffDrive/Omni/Ackermann)  include/nav2_mppi_controller/motion_models.hpp:39-351
//   mppi::critics::CriticFunction (+ 11 critics)    include/nav2_mppi_controller/critic_function.hpp:44-114
//   mppi::CriticManager                    include/nav2_mppi_controller/critic_manager.hpp:43-116
// The classes hold configuration only; every per-trajectory number is computed by the CUDA
// kernels behind mppi_optimize / mppi_score_tensors.  There is no CPU implementation to fall to.
#ifndef NAV2_MPPI_CONTROLLER_B200__MPPI_HPP_
#define NAV2_MPPI_CONTROLLER_B200__MPPI_HPP_

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "mppi_b200.h"
#include "ros_shim/ros_shim.hpp"

namespace mppi
{

enum class ParameterType {Dynamic, Static};

// tools/parameters_handler.hpp: getParamGetter(ns)(setting, name, default, type); a mutex shared
// with the control loop; pre-callbacks validate a dynamic update, post-callbacks run after it (Optimizer::reset()).
// rcl_interfaces::msg::SetParametersResult
struct SetParametersResult {bool successful{true}; std::string reason;};

class ParametersHandler
{
public:
  using pre_callback_t = std::function<void (const std::string &, const nav2::ParameterValue &, SetParametersResult &)>;
  ParametersHandler(const nav2::LifecycleNode::WeakPtr & parent, std::string & name)
  : node_(parent), name_(name) {}
  std::mutex * getLock() {return &parameters_change_mutex_;}
  void addPostCallback(std::function<void()> cb) {post_callbacks_.push_back(std::move(cb));}
  // parameters_handler.hpp:109,222: validation that runs BEFORE a dynamic update is applied and may reject it
  void addPreCallback(const std::string & full_name, pre_callback_t cb) {pre_callbacks_[full_name] = std::move(cb);}
  void start() {}

  auto getParamGetter(const std::string & ns)
  {
    return [this, ns](auto & setting, const std::string & name, auto default_value,
             ParameterType param_type = ParameterType::Dynamic) {
             getParam(setting, ns.empty() ? name : ns + "." + name, std::move(default_value), param_type);
           };
  }

  template<typename SettingT, typename ParamT>
  void getParam(SettingT & setting, const std::string & full_name, ParamT default_value, ParameterType param_type = ParameterType::Dynamic)
  {
    auto node = node_.lock();
    if (!node) {throw std::runtime_error("ParametersHandler: node expired");}
    if (!node->has_parameter(full_name)) {node->declare_parameter(full_name, toValue(default_value));}
    assign(setting, node->get_parameter(full_name));
    if (param_type == ParameterType::Static) {static_names_[full_name] = true;}
  }

  // A dynamic parameter update as the reference handles it: validateParameterUpdatesCallback
  // (parameters_handler.cpp:59-93: static parameters and every registered pre-callback may reject it), then
  // updateParametersCallback (:95-127: under the mutex shared with the control loop, apply, run the post-callbacks).
  // The mirror applies the value to the node's parameter store; the optimiser's post-callback re-reads its
  // settings from there and reconfigures the device (Optimizer::reset() in the reference, optimizer.cpp:155).
  SetParametersResult setDynamic(const std::string & full_name, const nav2::ParameterValue & value)
  {
    SetParametersResult result;
    if (full_name.rfind(name_ + ".", 0) != 0) {return result;}   // not this plugin's parameter
    if (static_names_.count(full_name)) {
      result.successful = false;
      result.reason = "Rejected change to static parameter: " + full_name;
    }
    if (auto cb = pre_callbacks_.find(full_name); cb != pre_callbacks_.end()) {cb->second(full_name, value, result);}
    if (!result.successful) {return result;}
    std::lock_guard<std::mutex> lock(parameters_change_mutex_);
    node_.lock()->set_parameter(full_name, value);
    for (auto & cb : post_callbacks_) {cb();}
    return result;
  }

private:
  static nav2::ParameterValue toValue(bool v) {return v;}
  static nav2::ParameterValue toValue(int v) {return v;}
  static nav2::ParameterValue toValue(unsigned int v) {return static_cast<int>(v);}
  static nav2::ParameterValue toValue(float v) {return static_cast<double>(v);}
  static nav2::ParameterValue toValue(double v) {return v;}
  static nav2::ParameterValue toValue(const char * v) {return std::string(v);}
  static nav2::ParameterValue toValue(const std::string & v) {return v;}
  static nav2::ParameterValue toValue(const std::vector<std::string> & v) {return v;}
  static nav2::ParameterValue toValue(const std::vector<double> & v) {return v;}

  template<typename T>
  static void assign(T & setting, const nav2::ParameterValue & v)
  {
    std::visit(
      [&setting](const auto & x) {
        using X = std::decay_t<decltype(x)>;
        if constexpr (std::is_arithmetic_v<T> && std::is_arithmetic_v<X>) {
          setting = static_cast<T>(x);
        } else if constexpr (std::is_same_v<T, X>) {
          setting = x;
        } else {
          throw std::runtime_error("parameter has the wrong type");
        }
      }, v);
  }

  nav2::LifecycleNode::WeakPtr node_;
  std::string name_;
  std::mutex parameters_change_mutex_;
  std::vector<std::function<void()>> post_callbacks_;
  std::map<std::string, pre_callback_t> pre_callbacks_;
  std::map<std::string, bool> static_names_;
};

// ---- motion models (motion_models.xml) ----
// models/control_sequence.hpp:37-49
struct ControlSequence
{
  std::vector<float> vx, vy, wz;
  void reset(unsigned int time_steps) {vx.assign(time_steps, 0.0f); vy.assign(time_steps, 0.0f); wz.assign(time_steps, 0.0f);}
};

// models::State as the plugin API sees it (models/state.hpp:31-57): B x T column-major float tensors
namespace models
{
struct State
{
  uint32_t batch_size{0}, time_steps{0};
  std::vector<float> vx, vy, wz, cvx, cvy, cwz;   // element (b, t) at t * batch_size + b
  geometry_msgs::msg::PoseStamped pose;
  geometry_msgs::msg::Twist speed;
  float local_path_length{0.0f};
  // models/state.hpp:46-56
  void reset(unsigned int batch, unsigned int steps)
  {
    batch_size = batch; time_steps = steps;
    const size_t n = static_cast<size_t>(batch) * steps;
    for (auto * v : {&vx, &vy, &wz, &cvx, &cvy, &cwz}) {v->assign(n, 0.0f);}
  }
};
struct Trajectories   // models/trajectories.hpp:27-41
{
  std::vector<float> x, y, yaws;
  void reset(unsigned int batch_size, unsigned int time_steps)
  {
    const size_t n = static_cast<size_t>(batch_size) * time_steps;
    x.assign(n, 0.0f); y.assign(n, 0.0f); yaws.assign(n, 0.0f);
  }
};
struct Path   // models/path.hpp:27-42
{
  std::vector<float> x, y, yaws;
  void reset(unsigned int size) {x.assign(size, 0.0f); y.assign(size, 0.0f); yaws.assign(size, 0.0f);}
};
using ControlSequence = mppi::ControlSequence;
}  // namespace models

class MotionModel
{
public:
  virtual ~MotionModel() = default;
  virtual void initialize(ParametersHandler *, const std::string &) {}
  virtual bool isHolonomic() const = 0;
  virtual int32_t modelId() const = 0;
  virtual float getMinTurningRadius() const {return 0.0f;}
  // MotionModel::predict (motion_models.hpp:108-167) runs on the device for the three models of motion_models.xml
  // (kernel A); a subclass that overrides it cannot be honoured: Optimizer::setMotionModel only accepts the three
  // known plugin types and throws ControllerException otherwise, like pluginlib failing to load a class.
  virtual void predict(models::State &) {
    throw nav2_core::ControllerException("MotionModel::predict runs on the device; host-side motion models are not supported");
  }
  // MotionModel::applyConstraints (motion_models.hpp:173-175, 292-298): T elements, host side like the reference
  virtual void applyConstraints(ControlSequence & /*control_sequence*/) {}
};
class DiffDriveMotionModel : public MotionModel
{
public:
  bool isHolonomic() const override {return false;}
  int32_t modelId() const override {return MPPI_MODEL_DIFF_DRIVE;}
};
class OmniMotionModel : public MotionModel
{
public:
  bool isHolonomic() const override {return true;}
  int32_t modelId() const override {return MPPI_MODEL_OMNI;}
};
class AckermannMotionModel : public MotionModel
{
public:
  void initialize(ParametersHandler * param_handler, const std::string & plugin_name) override
  {
    auto getParam = param_handler->getParamGetter(plugin_name);
    getParam(min_turning_r_, "min_turning_r", 0.2f);  // motion_models.hpp:276
  }
  bool isHolonomic() const override {return false;}
  int32_t modelId() const override {return MPPI_MODEL_ACKERMANN;}
  float getMinTurningRadius() const override {return min_turning_r_;}
  void applyConstraints(ControlSequence & control_sequence) override  // motion_models.hpp:292-298
  {
    for (size_t i = 0; i < control_sequence.vx.size(); ++i) {
      const float wz_constrained = std::fabs(control_sequence.vx[i]) / min_turning_r_;
      control_sequence.wz[i] = std::min(std::max(control_sequence.wz[i], -wz_constrained), wz_constrained);
    }
  }

private:
  float min_turning_r_{0.2f};
};

// ---- critics (critics.xml) ----
namespace critics
{
// One class per reference critic.  on_configure reads "<name>.enabled" then the critic's own
// parameters (same names and code defaults as src/critics/<name>_critic.cpp::initialize()).
// CriticData (critic_data.hpp:36-56) with the tensors as plain column-major float arrays
struct CriticData
{
  const models::State & state;
  const models::Trajectories & trajectories;
  const nav_msgs::msg::Path & path;
  const geometry_msgs::msg::Pose & goal;
  std::vector<float> & costs;                 // [batch_size], critics add to it
  float & model_dt;
  bool fail_flag{false};
  nav2_core::GoalChecker * goal_checker{nullptr};
  std::vector<bool> trajectories_in_collision;
  int32_t furthest_reached_path_point{-1};    // < 0: not set (critic_data.hpp:53 std::optional)
};

class CriticFunction
{
public:
  virtual ~CriticFunction();
  // CriticFunction::score (critic_function.hpp:95): evaluates THIS critic on caller tensors.  The eleven critics of
  // critics.xml are scored by the device kernels (mppi_score_tensors on a one-critic handle kept by the instance);
  // a third-party critic written against Eigen CriticData has no device counterpart and is rejected when it is
  // loaded (CriticManager::on_configure throws ControllerException for unknown names).
  virtual void score(CriticData & data);
  void on_configure(
    nav2::LifecycleNode::WeakPtr parent, const std::string & parent_name, const std::string & name,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler);
  virtual void initialize() = 0;
  const std::string & getName() const {return name_;}
  const MppiCriticParams & params() const {return p_;}

protected:
  bool enabled_{true};
  std::string name_, parent_name_;
  nav2::LifecycleNode::WeakPtr parent_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  ParametersHandler * parameters_handler_{nullptr};
  MppiCriticParams p_{};
  MppiHandle * score_handle_{nullptr};        // one-critic handle behind score(), created on first use
  uint32_t score_batch_{0}, score_steps_{0};
};

#define MPPI_DECLARE_CRITIC(Name) \
  class Name : public CriticFunction {public: void initialize() override;}
MPPI_DECLARE_CRITIC(ConstraintCritic);
MPPI_DECLARE_CRITIC(CostCritic);
MPPI_DECLARE_CRITIC(GoalCritic);
MPPI_DECLARE_CRITIC(GoalAngleCritic);
MPPI_DECLARE_CRITIC(ObstaclesCritic);
MPPI_DECLARE_CRITIC(PathAlignCritic);
MPPI_DECLARE_CRITIC(PathAngleCritic);
MPPI_DECLARE_CRITIC(PathFollowCritic);
MPPI_DECLARE_CRITIC(PreferForwardCritic);
MPPI_DECLARE_CRITIC(TwirlingCritic);
MPPI_DECLARE_CRITIC(VelocityDeadbandCritic);
#undef MPPI_DECLARE_CRITIC
}  // namespace critics

// critic_manager.hpp: loads "mppi::critics::<name>" for every entry of the "critics" parameter
class CriticManager
{
public:
  void on_configure(
    nav2::LifecycleNode::WeakPtr parent, const std::string & name,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler);
  std::string getFullName(const std::string & name) const {return "mppi::critics::" + name;}
  const std::vector<std::unique_ptr<critics::CriticFunction>> & critics() const {return critics_;}
  const std::vector<std::string> & names() const {return critic_names_;}
  // The visualize half of evalTrajectoriesScores (critic_manager.cpp:79-120): the device keeps what every critic
  // added to every trajectory (MPPI_TENSOR_CRITIC_COSTS) and how far the critic loop got; this assembles
  // critic_costs_ and the nav2_msgs/CriticsStats message from them and publishes it on "~/critics_stats".
  void publishCriticsStats(MppiHandle * handle, uint32_t batch_size, const builtin_interfaces::msg::Time & stamp);
  const std::vector<std::pair<std::string, std::vector<float>>> & getCriticCosts() const {return critic_costs_;}
  nav2::Publisher<nav2_msgs::msg::CriticsStats> * criticsEffectPublisher() {return critics_effect_pub_.get();}

private:
  std::vector<std::unique_ptr<critics::CriticFunction>> critics_;
  std::vector<std::string> critic_names_;
  bool visualize_{false};
  std::unique_ptr<nav2::Publisher<nav2_msgs::msg::CriticsStats>> critics_effect_pub_;
  std::vector<std::pair<std::string, std::vector<float>>> critic_costs_;
};

// trajectory_visualizer.hpp:42-135: candidate trajectories coloured by cost + the optimal trajectory, as markers.
// The device already keeps exactly the samples add() draws (every trajectory_step-th candidate at every
// time_step-th point, MPPI_TENSOR_VIS_XY), so `trajectories` here is that strided tensor, not B x T.
class TrajectoryVisualizer
{
public:
  void on_configure(
    nav2::LifecycleNode::WeakPtr parent, const std::string & name, const std::string & frame_id,
    ParametersHandler * parameters_handler);
  void on_cleanup() {trajectories_publisher_.reset(); optimal_path_pub_.reset();}
  void on_activate() {trajectories_publisher_->on_activate(); optimal_path_pub_->on_activate();}
  void on_deactivate() {trajectories_publisher_->on_deactivate(); optimal_path_pub_->on_deactivate();}
  // add(optimal trajectory, namespace, stamp) — trajectory_visualizer.cpp:61-108; rows of (x, y, yaw)
  void add(
    const std::vector<std::array<float, 3>> & trajectory, const std::string & marker_namespace,
    const builtin_interfaces::msg::Time & cmd_stamp);
  // add(candidate trajectories, costs, collisions, stamp) — trajectory_visualizer.cpp:110-152.
  // samples_xy: [ceil(B / trajectory_step)][ceil(T / time_step)][2]; costs / collisions: per trajectory, [B]
  void add(
    const std::vector<float> & samples_xy, size_t n_trajectories, size_t n_points, const std::vector<float> & costs,
    const std::vector<bool> & collisions, const builtin_interfaces::msg::Time & stamp);
  void visualize();
  void reset();
  static std_msgs::msg::ColorRGBA costToColor(float normalized);   // trajectory_visualizer.cpp:189-202
  nav2::Publisher<visualization_msgs::msg::MarkerArray> * trajectoriesPublisher() {return trajectories_publisher_.get();}
  nav2::Publisher<nav_msgs::msg::Path> * optimalPathPublisher() {return optimal_path_pub_.get();}
  int trajectoryStep() const {return trajectory_step_;}
  int timeStep() const {return time_step_;}

private:
  std::string frame_id_;
  std::unique_ptr<nav2::Publisher<visualization_msgs::msg::MarkerArray>> trajectories_publisher_;
  std::unique_ptr<nav2::Publisher<nav_msgs::msg::Path>> optimal_path_pub_;
  std::unique_ptr<visualization_msgs::msg::MarkerArray> points_;
  std::unique_ptr<nav_msgs::msg::Path> optimal_path_;
  int trajectory_step_{5}, time_step_{3};
  int marker_id_{0};
};

// optimizer.hpp:59-333
class Optimizer
{
public:
  Optimizer() = default;
  ~Optimizer() {shutdown();}
  void initialize(
    nav2::LifecycleNode::WeakPtr parent, const std::string & name,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, nav2::TransformBuffer::SharedPtr tf_buffer,
    ParametersHandler * param_handler);
  void shutdown();
  // returns the command and the optimal trajectory as rows of (x, y, yaw)
  std::tuple<geometry_msgs::msg::TwistStamped, std::vector<std::array<float, 3>>> evalControl(
    const geometry_msgs::msg::PoseStamped & robot_pose, const geometry_msgs::msg::Twist & robot_speed,
    const nav_msgs::msg::Path & plan, const geometry_msgs::msg::Pose & goal, nav2_core::GoalChecker * goal_checker);
  void reset(bool reset_dynamic_speed_limits = true);
  void setSpeedLimit(double speed_limit, bool percentage);
  const ControlSequence & getOptimalControlSequence();
  bool isHolonomic() const {return motion_model_ && motion_model_->isHolonomic();}
  const MppiConfig & getSettings() const {return cfg_;}
  MppiHandle * handle() {return handle_;}
  // what MPPIController::visualize reads (controller.cpp:139-158): total costs (optimizer.hpp getCosts), per-critic
  // costs, collision flags, and the candidate trajectories at the visualizer's strides
  std::vector<float> getCosts(uint32_t robot = 0);
  std::vector<bool> getCollisionFlags(uint32_t robot = 0);
  std::vector<float> getVisualizedTrajectories(size_t & n_trajectories, size_t & n_points, uint32_t robot = 0);
  CriticManager & getCriticManager() {return critic_manager_;}
  // Fleet mode: one optimiser instance, n_robots independent robots sharing the
  // configuration; set before initialize().  evalControlFleet is evalControl for all of them in ONE
  // mppi_optimize_in_place call: element r of every argument belongs to robot r.
  void setFleetSize(uint32_t n_robots) {n_robots_ = n_robots;}
  struct FleetInput
  {
    geometry_msgs::msg::PoseStamped robot_pose;
    geometry_msgs::msg::Twist robot_speed;
    const nav_msgs::msg::Path * plan{nullptr};
    geometry_msgs::msg::Pose goal;
    nav2_costmap_2d::Costmap2D * costmap{nullptr};
  };
  struct FleetOutput
  {
    int status{MPPI_OK};                      // MPPI_ERR_NO_VALID_CONTROL where the reference would throw NoValidControl
    geometry_msgs::msg::TwistStamped cmd;
    std::vector<std::array<float, 3>> optimal_trajectory;
  };
  std::vector<FleetOutput> evalControlFleet(const std::vector<FleetInput> & inputs, nav2_core::GoalChecker * goal_checker);

  // Optimizer::isSpeedLimitActive (optimizer.cpp:218-229)
  bool isSpeedLimitActive() const {return speed_limit_active_;}

protected:
  void getParams();
  void fillDerivedSettings();
  void reconfigure();      // post-callback of a dynamic parameter update: re-read the settings, mppi_set_config
  void setMotionModel(const std::string & motion_model_name);
  void uploadCostmap();
  [[noreturn]] void throwStatus(int status, const char * what) const;

  nav2::LifecycleNode::WeakPtr parent_;
  std::string name_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  nav2_costmap_2d::Costmap2D * costmap_{nullptr};
  ParametersHandler * parameters_handler_{nullptr};
  std::shared_ptr<MotionModel> motion_model_;
  CriticManager critic_manager_;
  MppiConfig cfg_{};
  MppiHandle * handle_{nullptr};
  ControlSequence control_sequence_;
  uint32_t n_robots_{1};
  bool speed_limit_active_{false};
  bool callbacks_registered_{false};
};

}  // namespace mppi

namespace mppi::utils
{
// builtin_interfaces::msg::Time + rclcpp::Duration::from_seconds(seconds): the duration truncated to nanoseconds
inline builtin_interfaces::msg::Time addSeconds(const builtin_interfaces::msg::Time & stamp, double seconds)
{
  const int64_t total = static_cast<int64_t>(stamp.sec) * 1000000000LL + static_cast<int64_t>(stamp.nanosec) +
    static_cast<int64_t>(seconds * 1e9);
  builtin_interfaces::msg::Time out;
  int64_t sec = total / 1000000000LL, nsec = total % 1000000000LL;
  if (nsec < 0) {nsec += 1000000000LL; --sec;}
  out.sec = static_cast<int32_t>(sec);
  out.nanosec = static_cast<uint32_t>(nsec);
  return out;
}

// utils::toTrajectoryMsg (tools/utils.hpp:174-201): the optimal trajectory (rows of x, y, yaw) with the optimal control
// sequence's velocities, one point per time step, stamped header.stamp + i * model_dt
inline std::unique_ptr<nav_msgs::msg::Trajectory> toTrajectoryMsg(
  const std::vector<std::array<float, 3>> & trajectory, const ControlSequence & control_sequence, const double & model_dt,
  const std_msgs::msg::Header & header)
{
  auto trajectory_msg = std::make_unique<nav_msgs::msg::Trajectory>();
  trajectory_msg->header = header;
  trajectory_msg->points.resize(trajectory.size());
  for (size_t i = 0; i < trajectory.size(); ++i) {
    auto & curr_pt = trajectory_msg->points[i];
    curr_pt.header.frame_id = header.frame_id;
    curr_pt.header.stamp = addSeconds(header.stamp, static_cast<double>(i) * model_dt);
    curr_pt.pose.position.x = trajectory[i][0];
    curr_pt.pose.position.y = trajectory[i][1];
    curr_pt.pose.orientation = tf2::quaternionFromYaw(trajectory[i][2]);
    curr_pt.velocity.linear.x = control_sequence.vx[i];
    curr_pt.velocity.angular.z = control_sequence.wz[i];
    if (!control_sequence.vy.empty()) {curr_pt.velocity.linear.y = control_sequence.vy[i];}
  }
  return trajectory_msg;
}
}  // namespace mppi::utils

namespace nav2_mppi_controller
{

// controller.hpp:39-131 — the nav2_core::Controller plugin (pluginlib class "nav2_mppi_controller::MPPIController")
class MPPIController : public nav2_core::Controller
{
public:
  MPPIController() = default;
  void configure(
    const nav2::LifecycleNode::WeakPtr & parent, std::string name, nav2::TransformBuffer::SharedPtr tf,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros) override;
  void cleanup() override;
  void activate() override;
  void deactivate() override;
  void reset() override;
  geometry_msgs::msg::TwistStamped computeVelocityCommands(
    const geometry_msgs::msg::PoseStamped & robot_pose, const geometry_msgs::msg::Twist & robot_speed,
    nav2_core::GoalChecker * goal_checker, const nav_msgs::msg::Path & transformed_global_plan,
    const geometry_msgs::msg::PoseStamped & global_goal) override;
  void newPathReceived(const nav_msgs::msg::Path & raw_global_path) override;
  void setSpeedLimit(const double & speed_limit, const bool & percentage) override;
  mppi::Optimizer & optimizer() {return optimizer_;}
  const std::vector<std::array<float, 3>> & lastOptimalTrajectory() const {return optimal_trajectory_;}
  mppi::TrajectoryVisualizer & trajectoryVisualizer() {return trajectory_visualizer_;}
  // "~/optimal_trajectory" (controller.cpp:43-55): exists only with publish_optimal_trajectory
  nav2::Publisher<nav_msgs::msg::Trajectory> * optimalTrajectoryPublisher() {return opt_traj_pub_.get();}

protected:
  void visualize(const builtin_interfaces::msg::Time & cmd_stamp, const std::vector<std::array<float, 3>> & optimal_trajectory);

  std::string name_;
  nav2::LifecycleNode::WeakPtr parent_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  nav2::TransformBuffer::SharedPtr tf_buffer_;
  std::unique_ptr<mppi::ParametersHandler> parameters_handler_;
  mppi::Optimizer optimizer_;
  mppi::TrajectoryVisualizer trajectory_visualizer_;
  std::vector<std::array<float, 3>> optimal_trajectory_;
  bool visualize_{false};
  int critic_index_to_visualize_{0};
  bool publish_optimal_trajectory_{false};
  std::unique_ptr<nav2::Publisher<nav_msgs::msg::Trajectory>> opt_traj_pub_;
};
}  // namespace nav2_mppi_controller

#endif  // NAV2_MPPI_CONTROLLER_B200__MPPI_HPP_
