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
// with the control loop; post-callbacks run after a dynamic update (Optimizer::reset()).
class ParametersHandler
{
public:
  ParametersHandler(const nav2::LifecycleNode::WeakPtr & parent, std::string & name)
  : node_(parent), name_(name) {}
  std::mutex * getLock() {return &parameters_change_mutex_;}
  void addPostCallback(std::function<void()> cb) {post_callbacks_.push_back(std::move(cb));}
  void start() {}

  auto getParamGetter(const std::string & ns)
  {
    return [this, ns](auto & setting, const std::string & name, auto default_value,
             ParameterType = ParameterType::Dynamic) {
             getParam(setting, ns.empty() ? name : ns + "." + name, std::move(default_value));
           };
  }

  template<typename SettingT, typename ParamT>
  void getParam(SettingT & setting, const std::string & full_name, ParamT default_value)
  {
    auto node = node_.lock();
    if (!node) {throw std::runtime_error("ParametersHandler: node expired");}
    if (!node->has_parameter(full_name)) {node->declare_parameter(full_name, toValue(default_value));}
    assign(setting, node->get_parameter(full_name));
  }

  // dynamic update (parameters_handler.cpp:96): takes the mutex, sets, runs the post-callbacks
  void setDynamic(const std::string & full_name, const nav2::ParameterValue & value)
  {
    std::lock_guard<std::mutex> lock(parameters_change_mutex_);
    node_.lock()->set_parameter(full_name, value);
    for (auto & cb : post_callbacks_) {cb();}
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
};

// ---- motion models (motion_models.xml) ----
class MotionModel
{
public:
  virtual ~MotionModel() = default;
  virtual void initialize(ParametersHandler *, const std::string &) {}
  virtual bool isHolonomic() const = 0;
  virtual int32_t modelId() const = 0;
  virtual float getMinTurningRadius() const {return 0.0f;}
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

private:
  float min_turning_r_{0.2f};
};

// ---- critics (critics.xml) ----
namespace critics
{
// One class per reference critic.  on_configure reads "<name>.enabled" then the critic's own
// parameters (same names and code defaults as src/critics/<name>_critic.cpp::initialize()).
class CriticFunction
{
public:
  virtual ~CriticFunction() = default;
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

private:
  std::vector<std::unique_ptr<critics::CriticFunction>> critics_;
  std::vector<std::string> critic_names_;
};

struct ControlSequence {std::vector<float> vx, vy, wz;};

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

protected:
  void getParams();
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
};

}  // namespace mppi

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

protected:
  std::string name_;
  nav2::LifecycleNode::WeakPtr parent_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  nav2::TransformBuffer::SharedPtr tf_buffer_;
  std::unique_ptr<mppi::ParametersHandler> parameters_handler_;
  mppi::Optimizer optimizer_;
  std::vector<std::array<float, 3>> optimal_trajectory_;
  bool visualize_{false};
  int critic_index_to_visualize_{0};
  bool publish_optimal_trajectory_{false};
};
}  // namespace nav2_mppi_controller

#endif  // NAV2_MPPI_CONTROLLER_B200__MPPI_HPP_
