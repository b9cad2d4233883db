// mppi.cpp — host mirror of nav2_mppi_controller (see mppi.hpp).  Reference citations are relative to
// /root/reference/nav2_mppi_controller/.
#include "nav2_mppi_controller_b200/mppi.hpp"

#include <array>
#include <cstring>

namespace mppi
{

// ------------------------------------------------------------------ critics

namespace critics
{

void CriticFunction::on_configure(
  nav2::LifecycleNode::WeakPtr parent, const std::string & parent_name, const std::string & name,
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler)
{
  // critic_function.hpp:68-88
  parent_ = parent;
  name_ = name;
  parent_name_ = parent_name;
  costmap_ros_ = costmap_ros;
  parameters_handler_ = param_handler;
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(enabled_, "enabled", true);
  initialize();
  p_.enabled = enabled_ ? 1 : 0;
}

static void defaults(MppiCriticParams & p, int32_t type)
{
  if (mppi_default_critic_params(type, &p) != MPPI_OK) {
    throw nav2_core::ControllerException("unknown critic type");
  }
}

void ConstraintCritic::initialize()  // src/critics/constraint_critic.cpp:21-35
{
  defaults(p_, MPPI_CRITIC_CONSTRAINT);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 4.0f);
}

void CostCritic::initialize()  // src/critics/cost_critic.cpp:24-76
{
  defaults(p_, MPPI_CRITIC_COST);
  auto getParam = parameters_handler_->getParamGetter(name_);
  bool consider_footprint = false;
  std::string inflation_layer_name;
  getParam(consider_footprint, "consider_footprint", false);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 3.81f);
  getParam(p_.critical_cost, "critical_cost", 300.0f);
  getParam(p_.near_collision_cost, "near_collision_cost", 253);
  getParam(p_.collision_cost, "collision_cost", 1000000.0f);
  getParam(p_.near_goal_distance, "near_goal_distance", 0.5f);
  getParam(inflation_layer_name, "inflation_layer_name", std::string(""));
  getParam(p_.trajectory_point_step, "trajectory_point_step", 2);
  p_.consider_footprint = consider_footprint ? 1 : 0;
  if (costmap_ros_->getUseRadius() == consider_footprint && costmap_ros_->getUseRadius()) {
    throw nav2_core::ControllerException(
            "Considering footprint in CostCritic but no robot footprint provided in the "
            "costmap (robot radius used instead). Disable considering footprint.");
  }
}

void GoalCritic::initialize()  // src/critics/goal_critic.cpp:22-28
{
  defaults(p_, MPPI_CRITIC_GOAL);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 5.0f);
  getParam(p_.threshold_to_consider, "threshold_to_consider", 1.4f);
}

void GoalAngleCritic::initialize()  // src/critics/goal_angle_critic.cpp:20-27
{
  defaults(p_, MPPI_CRITIC_GOAL_ANGLE);
  auto getParam = parameters_handler_->getParamGetter(name_);
  bool symmetric = false;
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 3.0f);
  getParam(p_.threshold_to_consider, "threshold_to_consider", 0.5f);
  getParam(symmetric, "symmetric_yaw_tolerance", false);
  p_.symmetric_yaw_tolerance = symmetric ? 1 : 0;
}

void ObstaclesCritic::initialize()  // src/critics/obstacles_critic.cpp:24-61
{
  defaults(p_, MPPI_CRITIC_OBSTACLES);
  auto getParam = parameters_handler_->getParamGetter(name_);
  bool consider_footprint = false;
  std::string inflation_layer_name;
  getParam(consider_footprint, "consider_footprint", false);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.repulsion_weight, "repulsion_weight", 1.5f);
  getParam(p_.critical_weight, "critical_weight", 20.0f);
  getParam(p_.collision_cost, "collision_cost", 100000.0f);
  getParam(p_.collision_margin_distance, "collision_margin_distance", 0.10f);
  getParam(p_.near_goal_distance, "near_goal_distance", 0.5f);
  getParam(inflation_layer_name, "inflation_layer_name", std::string(""));
  p_.consider_footprint = consider_footprint ? 1 : 0;
  if (costmap_ros_->getUseRadius() == consider_footprint && costmap_ros_->getUseRadius()) {
    throw nav2_core::ControllerException(
            "Considering footprint in collision checking but no robot footprint provided in the costmap.");
  }
}

void PathAlignCritic::initialize()  // src/critics/path_align_critic.cpp:20-32
{
  defaults(p_, MPPI_CRITIC_PATH_ALIGN);
  auto getParam = parameters_handler_->getParamGetter(name_);
  bool use_path_orientations = false;
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 10.0f);
  getParam(p_.max_path_occupancy_ratio, "max_path_occupancy_ratio", 0.07f);
  getParam(p_.offset_from_furthest, "offset_from_furthest", 20);
  getParam(p_.trajectory_point_step, "trajectory_point_step", 4);
  getParam(p_.threshold_to_consider, "threshold_to_consider", 0.5f);
  getParam(use_path_orientations, "use_path_orientations", false);
  p_.use_path_orientations = use_path_orientations ? 1 : 0;
}

void PathAngleCritic::initialize()  // src/critics/path_angle_critic.cpp:23-54
{
  defaults(p_, MPPI_CRITIC_PATH_ANGLE);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.offset_from_furthest, "offset_from_furthest", 4);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 2.2f);
  getParam(p_.threshold_to_consider, "threshold_to_consider", 0.5f);
  getParam(p_.max_angle_to_furthest, "max_angle_to_furthest", 0.785398f);
  int mode = 0;
  getParam(mode, "mode", mode);
  if (mode < 0 || mode > 2) {
    throw nav2_core::ControllerException("Invalid path angle mode!");  // :95
  }
  p_.mode = mode;  // the reversing_allowed_ fix-up of :25-50 is applied by mppi_create from vx_min
}

void PathFollowCritic::initialize()  // src/critics/path_follow_critic.cpp:22-32
{
  defaults(p_, MPPI_CRITIC_PATH_FOLLOW);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.threshold_to_consider, "threshold_to_consider", 1.4f);
  getParam(p_.offset_from_furthest, "offset_from_furthest", 6);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 5.0f);
}

void PreferForwardCritic::initialize()  // src/critics/prefer_forward_critic.cpp:22-30
{
  defaults(p_, MPPI_CRITIC_PREFER_FORWARD);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 5.0f);
  getParam(p_.threshold_to_consider, "threshold_to_consider", 0.5f);
}

void TwirlingCritic::initialize()  // src/critics/twirling_critic.cpp:22-27
{
  defaults(p_, MPPI_CRITIC_TWIRLING);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 10.0f);
}

void VelocityDeadbandCritic::initialize()  // src/critics/velocity_deadband_critic.cpp:20-32
{
  defaults(p_, MPPI_CRITIC_VELOCITY_DEADBAND);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(p_.cost_power, "cost_power", 1);
  getParam(p_.cost_weight, "cost_weight", 35.0f);
  std::vector<double> deadband{0.0, 0.0, 0.0};
  getParam(deadband, "deadband_velocities", std::vector<double>{0.0, 0.0, 0.0});
  for (size_t i = 0; i < 3 && i < deadband.size(); ++i) {p_.deadband_velocities[i] = static_cast<float>(deadband[i]);}
}

}  // namespace critics

// ------------------------------------------------------------------ CriticManager

// stands in for pluginlib::ClassLoader<CriticFunction>::createUniqueInstance (critic_manager.cpp:44-69)
static std::unique_ptr<critics::CriticFunction> createCritic(const std::string & full_name)
{
  const std::string prefix = "mppi::critics::";
  const std::string n = full_name.rfind(prefix, 0) == 0 ? full_name.substr(prefix.size()) : full_name;
#define MPPI_MAKE(Name) if (n == #Name) {return std::make_unique<critics::Name>();}
  MPPI_MAKE(ConstraintCritic) MPPI_MAKE(CostCritic) MPPI_MAKE(GoalCritic) MPPI_MAKE(GoalAngleCritic)
  MPPI_MAKE(ObstaclesCritic) MPPI_MAKE(PathAlignCritic) MPPI_MAKE(PathAngleCritic) MPPI_MAKE(PathFollowCritic)
  MPPI_MAKE(PreferForwardCritic) MPPI_MAKE(TwirlingCritic) MPPI_MAKE(VelocityDeadbandCritic)
#undef MPPI_MAKE
  throw nav2_core::ControllerException("Failed to load critic plugin '" + full_name + "'");
}

void CriticManager::on_configure(
  nav2::LifecycleNode::WeakPtr parent, const std::string & name,
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler)
{
  auto getParam = param_handler->getParamGetter(name);
  getParam(critic_names_, "critics", std::vector<std::string>{}, ParameterType::Static);  // critic_manager.cpp:40
  critics_.clear();
  for (const auto & critic_name : critic_names_) {
    auto instance = createCritic(getFullName(critic_name));
    instance->on_configure(parent, name, name + "." + critic_name, costmap_ros, param_handler);
    critics_.push_back(std::move(instance));
  }
}

// ------------------------------------------------------------------ Optimizer

void Optimizer::throwStatus(int status, const char * what) const
{
  const std::string msg = std::string(what) + ": " + mppi_status_string(status) + " — " + mppi_last_error(handle_);
  if (status == MPPI_ERR_NO_VALID_CONTROL) {throw nav2_core::NoValidControl(msg);}
  if (status == MPPI_ERR_CONFIG) {throw nav2_core::ControllerException(msg);}
  throw std::runtime_error(msg);
}

void Optimizer::setMotionModel(const std::string & motion_model_name)  // optimizer.cpp:666-689
{
  const std::string plugin_ns = name_ + "." + motion_model_name;
  std::string plugin_type;
  auto getParam = parameters_handler_->getParamGetter(plugin_ns);
  const std::map<std::string, std::string> known = {
    {"diff_drive", "mppi::DiffDriveMotionModel"}, {"omni", "mppi::OmniMotionModel"}, {"ackermann", "mppi::AckermannMotionModel"},
    {"DiffDrive", "mppi::DiffDriveMotionModel"}, {"Omni", "mppi::OmniMotionModel"}, {"Ackermann", "mppi::AckermannMotionModel"}};
  auto it = known.find(motion_model_name);
  getParam(plugin_type, "plugin", it != known.end() ? it->second : std::string(""), ParameterType::Static);
  if (plugin_type == "mppi::DiffDriveMotionModel") {
    motion_model_ = std::make_shared<DiffDriveMotionModel>();
  } else if (plugin_type == "mppi::OmniMotionModel") {
    motion_model_ = std::make_shared<OmniMotionModel>();
  } else if (plugin_type == "mppi::AckermannMotionModel") {
    motion_model_ = std::make_shared<AckermannMotionModel>();
  } else {
    throw nav2_core::ControllerException(
            std::string("Failed to load motion model plugin '") + motion_model_name + "'");
  }
  motion_model_->initialize(parameters_handler_, plugin_ns);
}

void Optimizer::getParams()  // optimizer.cpp:77-161
{
  std::string motion_model_name;
  auto & s = cfg_;
  mppi_default_config(&s);
  auto getParam = parameters_handler_->getParamGetter(name_);
  auto getParentParam = parameters_handler_->getParamGetter("");
  getParam(s.model_dt, "model_dt", 0.05f);
  getParam(s.model_delay_vx, "model_delay_vx", 0.0f);
  getParam(s.model_delay_vy, "model_delay_vy", 0.0f);
  getParam(s.model_delay_wz, "model_delay_wz", 0.0f);
  bool clamp_raw = false, open_loop = false, regenerate = false, visualize = false;
  getParam(clamp_raw, "clamp_raw_controls", false);
  getParam(s.time_steps, "time_steps", 56);
  getParam(s.batch_size, "batch_size", 1000);
  getParam(s.iteration_count, "iteration_count", 1);
  getParam(s.temperature, "temperature", 0.3f);
  getParam(s.gamma, "gamma", 0.015f);
  getParam(s.vx_max, "vx_max", 0.5f);
  getParam(s.vx_min, "vx_min", -0.35f);
  getParam(s.vy_max, "vy_max", 0.5f);
  getParam(s.wz_max, "wz_max", 1.9f);
  getParam(s.ax_max, "ax_max", 3.0f);
  getParam(s.ax_min, "ax_min", -3.0f);
  getParam(s.ay_max, "ay_max", 3.0f);
  getParam(s.ay_min, "ay_min", -3.0f);
  getParam(s.az_max, "az_max", 3.5f);
  getParam(s.vx_std, "vx_std", 0.2f);
  getParam(s.vy_std, "vy_std", 0.2f);
  getParam(s.wz_std, "wz_std", 0.4f);
  getParam(s.retry_attempt_limit, "retry_attempt_limit", 1);
  getParam(open_loop, "open_loop", false);
  getParam(s.sgf_order, "sgf_order", 2);
  getParam(regenerate, "regenerate_noises", false);   // noise_generator.cpp:36
  getParam(visualize, "visualize", false);            // controller.cpp:40
  s.clamp_raw_controls = clamp_raw ? 1 : 0;
  s.open_loop = open_loop ? 1 : 0;
  s.regenerate_noises = regenerate ? 1 : 0;
  s.visualize = visualize ? 1 : 0;
  s.materialize_tensors = 0;  // the visualizer's strided samples come from MPPI_TENSOR_VIS_XY, not from B x T tensors
  {
    // TrajectoryVisualizer's own parameters (trajectory_visualizer.cpp:36-39) decide which samples the device keeps
    auto getVisParam = parameters_handler_->getParamGetter(name_ + ".TrajectoryVisualizer");
    int trajectory_step = 5, time_step = 3;
    getVisParam(trajectory_step, "trajectory_step", 5);
    getVisParam(time_step, "time_step", 3);
    s.vis_trajectory_step = static_cast<uint32_t>(std::max(trajectory_step, 1));
    s.vis_time_step = static_cast<uint32_t>(std::max(time_step, 1));
  }
  getParam(motion_model_name, "motion_model", std::string("diff_drive"));
  setMotionModel(motion_model_name);
  s.motion_model = motion_model_->modelId();
  s.min_turning_r = motion_model_->getMinTurningRadius() > 0.0f ? motion_model_->getMinTurningRadius() : 0.2f;
  parameters_handler_->addPostCallback([this]() {reset();});
  double controller_frequency = 0.0;
  getParentParam(controller_frequency, "controller_frequency", 0.0, ParameterType::Static);
  if (!(controller_frequency > 0.0)) {
    throw nav2_core::ControllerException("controller_frequency must be set on the parent node");
  }
  s.controller_frequency = controller_frequency;

  // TrajectoryValidator (optimal_trajectory_validator.hpp:87-103)
  std::string validator_plugin;
  auto getValidatorParam = parameters_handler_->getParamGetter(name_ + ".TrajectoryValidator");
  getValidatorParam(validator_plugin, "plugin", std::string("mppi::DefaultOptimalTrajectoryValidator"), ParameterType::Static);
  if (validator_plugin != "mppi::DefaultOptimalTrajectoryValidator") {
    throw nav2_core::ControllerException("Failed to load trajectory validator plugin '" + validator_plugin + "'");
  }
  bool validator_footprint = false;
  getValidatorParam(s.collision_lookahead_time, "collision_lookahead_time", 2.0f);
  getValidatorParam(validator_footprint, "consider_footprint", false);
  s.validator_enabled = 1;
  s.validator_consider_footprint = validator_footprint ? 1 : 0;
}

void Optimizer::initialize(
  nav2::LifecycleNode::WeakPtr parent, const std::string & name,
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, nav2::TransformBuffer::SharedPtr,
  ParametersHandler * param_handler)
{
  // optimizer.cpp:34-70
  parent_ = parent;
  name_ = name;
  costmap_ros_ = costmap_ros;
  costmap_ = costmap_ros_->getCostmap();
  parameters_handler_ = param_handler;
  getParams();
  critic_manager_.on_configure(parent_, name_, costmap_ros_, parameters_handler_);

  auto & s = cfg_;
  const auto & critics = critic_manager_.critics();
  if (critics.size() > MPPI_MAX_CRITICS) {throw nav2_core::ControllerException("too many critics");}
  s.n_critics = static_cast<uint32_t>(critics.size());
  for (size_t i = 0; i < critics.size(); ++i) {s.critics[i] = critics[i]->params();}

  // what the critics read from Costmap2DROS / LayeredCostmap / InflationLayer (cost_critic.cpp:85-129)
  s.use_radius = costmap_ros_->getUseRadius() ? 1 : 0;
  const auto & fp = costmap_ros_->getRobotFootprint();
  if (fp.size() > MPPI_MAX_FOOTPRINT) {throw nav2_core::ControllerException("footprint has too many points");}
  s.n_footprint = static_cast<uint32_t>(fp.size());
  for (size_t i = 0; i < fp.size(); ++i) {s.footprint_x[i] = fp[i].x; s.footprint_y[i] = fp[i].y;}
  s.inscribed_radius = costmap_ros_->getLayeredCostmap()->getInscribedRadius();
  s.circumscribed_radius = costmap_ros_->getLayeredCostmap()->getCircumscribedRadius();
  s.has_inflation_layer = costmap_ros_->inflation().present ? 1 : 0;
  s.inflation_cost_scaling_factor = costmap_ros_->inflation().cost_scaling_factor;
  s.inflation_radius = costmap_ros_->inflation().inflation_radius;
  s.track_unknown = costmap_ros_->getLayeredCostmap()->isTrackingUnknown() ? 1 : 0;

  if (handle_) {mppi_destroy(handle_); handle_ = nullptr;}
  const int rc = mppi_create(&s, &handle_);
  if (rc != MPPI_OK) {
    const std::string msg = std::string(mppi_status_string(rc)) + ": " + mppi_last_error(nullptr);
    if (rc == MPPI_ERR_CONFIG) {throw nav2_core::ControllerException(msg);}
    throw std::runtime_error("mppi_create failed (" + msg + "); the optimiser has no CPU fallback");
  }
  control_sequence_.vx.assign(s.time_steps, 0.0f);
  control_sequence_.vy.assign(s.time_steps, 0.0f);
  control_sequence_.wz.assign(s.time_steps, 0.0f);
}

void Optimizer::shutdown()  // optimizer.cpp:72-75
{
  if (handle_) {mppi_destroy(handle_); handle_ = nullptr;}
}

void Optimizer::reset(bool reset_dynamic_speed_limits)  // optimizer.cpp:183-211
{
  if (!handle_) {return;}
  const int rc = mppi_reset(handle_, UINT32_MAX, reset_dynamic_speed_limits ? 1 : 0);
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_reset");}
}

void Optimizer::setSpeedLimit(double speed_limit, bool percentage)  // optimizer.cpp:691-719
{
  const int rc = mppi_set_speed_limit(handle_, speed_limit, percentage ? 1 : 0);
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_set_speed_limit");}
}

void Optimizer::uploadCostmap()
{
  // the caller holds the costmap mutex (controller.cpp:106-107)
  const int rc = mppi_upload_costmap(
    handle_, 0, costmap_->getCharMap(), costmap_->getSizeInCellsX(), costmap_->getSizeInCellsY(),
    costmap_->getResolution(), costmap_->getOriginX(), costmap_->getOriginY());
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_upload_costmap");}
}

std::tuple<geometry_msgs::msg::TwistStamped, std::vector<std::array<float, 3>>> Optimizer::evalControl(
  const geometry_msgs::msg::PoseStamped & robot_pose, const geometry_msgs::msg::Twist & robot_speed,
  const nav_msgs::msg::Path & plan, const geometry_msgs::msg::Pose & goal, nav2_core::GoalChecker * goal_checker)
{
  // the caller holds the costmap mutex (controller.cpp:106-107): the costmap is read where it lies, like the
  // critics read it during evalControl (cost_critic.cpp:138-144), instead of being copied first
  const MppiCostmapView costmap_view{
    costmap_->getCharMap(), costmap_->getSizeInCellsX(), costmap_->getSizeInCellsY(),
    costmap_->getResolution(), costmap_->getOriginX(), costmap_->getOriginY()};
  const size_t n = plan.poses.size();
  std::vector<double> px(n), py(n), pyaw(n);
  for (size_t i = 0; i < n; ++i) {
    px[i] = plan.poses[i].pose.position.x;
    py[i] = plan.poses[i].pose.position.y;
    pyaw[i] = tf2::getYaw(plan.poses[i].pose.orientation);  // utils::toTensor (tools/utils.hpp:208-220)
  }
  MppiCycleIn in{};
  in.pose_x = robot_pose.pose.position.x;
  in.pose_y = robot_pose.pose.position.y;
  in.pose_yaw = tf2::getYaw(robot_pose.pose.orientation);
  in.speed_vx = robot_speed.linear.x;
  in.speed_vy = robot_speed.linear.y;
  in.speed_wz = robot_speed.angular.z;
  in.path_x = px.data(); in.path_y = py.data(); in.path_yaw = pyaw.data();
  in.n_path = static_cast<uint32_t>(n);
  in.goal_x = goal.position.x; in.goal_y = goal.position.y; in.goal_yaw = tf2::getYaw(goal.orientation);
  if (goal_checker != nullptr) {
    geometry_msgs::msg::Pose pose_tolerance;
    geometry_msgs::msg::Twist velocity_tolerance;
    double path_length_tolerance = 0.0;
    goal_checker->getTolerances(pose_tolerance, velocity_tolerance, path_length_tolerance);  // twirling_critic.cpp:39-44
    in.has_goal_checker = 1;
    in.goal_checker_xy_tolerance = pose_tolerance.position.x;
  }
  const uint32_t T = cfg_.time_steps;
  std::vector<float> traj(3 * T);
  MppiCycleOut out{};
  out.control_vx = control_sequence_.vx.data();
  out.control_vy = control_sequence_.vy.data();
  out.control_wz = control_sequence_.wz.data();
  out.optimal_trajectory = traj.data();
  const int rc = mppi_optimize_in_place(handle_, &costmap_view, &in, &out);
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_optimize_in_place");}
  if (out.status == MPPI_ERR_NO_VALID_CONTROL) {
    throw nav2_core::NoValidControl("Optimizer fail to compute path");  // optimizer.cpp:294
  }
  if (out.status != MPPI_OK) {throwStatus(out.status, "mppi_optimize");}
  geometry_msgs::msg::TwistStamped cmd;  // utils::toTwistStamped (tools/utils.hpp:144-172)
  cmd.header.frame_id = costmap_ros_->getBaseFrameID();
  cmd.header.stamp = plan.header.stamp;
  cmd.twist.linear.x = out.cmd_vx;
  cmd.twist.angular.z = out.cmd_wz;
  if (isHolonomic()) {cmd.twist.linear.y = out.cmd_vy;}
  std::vector<std::array<float, 3>> rows(T);
  for (uint32_t i = 0; i < T; ++i) {rows[i] = {traj[i], traj[T + i], traj[2 * T + i]};}
  return std::make_tuple(cmd, rows);
}

const ControlSequence & Optimizer::getOptimalControlSequence()  // optimizer.cpp:600-603
{
  const int rc = mppi_get_control_sequence(
    handle_, 0, control_sequence_.vx.data(), control_sequence_.vy.data(), control_sequence_.wz.data());
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_get_control_sequence");}
  return control_sequence_;
}

}  // namespace mppi

// ------------------------------------------------------------------ MPPIController

namespace nav2_mppi_controller
{

void MPPIController::configure(
  const nav2::LifecycleNode::WeakPtr & parent, std::string name, nav2::TransformBuffer::SharedPtr tf,
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros)
{
  // controller.cpp:26-57
  parent_ = parent;
  costmap_ros_ = costmap_ros;
  tf_buffer_ = tf;
  name_ = name;
  parameters_handler_ = std::make_unique<mppi::ParametersHandler>(parent, name_);
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(visualize_, "visualize", false);
  getParam(critic_index_to_visualize_, "critic_index_to_visualize", 0);
  getParam(publish_optimal_trajectory_, "publish_optimal_trajectory", false);
  optimizer_.initialize(parent_, name_, costmap_ros_, tf_buffer_, parameters_handler_.get());
}

void MPPIController::cleanup()  // controller.cpp:59-67
{
  optimizer_.shutdown();
  parameters_handler_.reset();
}

void MPPIController::activate()  // controller.cpp:69-78
{
  parameters_handler_->start();
}

void MPPIController::deactivate() {}  // controller.cpp:80-87

void MPPIController::reset()  // controller.cpp:88-91
{
  optimizer_.reset(false /*Don't reset zone-based speed limits between requests*/);
}

geometry_msgs::msg::TwistStamped MPPIController::computeVelocityCommands(
  const geometry_msgs::msg::PoseStamped & robot_pose, const geometry_msgs::msg::Twist & robot_speed,
  nav2_core::GoalChecker * goal_checker, const nav_msgs::msg::Path & transformed_global_plan,
  const geometry_msgs::msg::PoseStamped & global_goal)
{
  // controller.cpp:93-137: parameter lock, then the costmap's recursive mutex, for the whole optimisation
  std::lock_guard<std::mutex> param_lock(*parameters_handler_->getLock());
  nav2_costmap_2d::Costmap2D * costmap = costmap_ros_->getCostmap();
  std::unique_lock<nav2_costmap_2d::Costmap2D::mutex_t> costmap_lock(*(costmap->getMutex()));
  auto [cmd, optimal_trajectory] =
    optimizer_.evalControl(robot_pose, robot_speed, transformed_global_plan, global_goal.pose, goal_checker);
  optimal_trajectory_ = std::move(optimal_trajectory);
  return cmd;
}

void MPPIController::newPathReceived(const nav_msgs::msg::Path &) {}  // controller.cpp:160-162

void MPPIController::setSpeedLimit(const double & speed_limit, const bool & percentage)  // controller.cpp:164-167
{
  optimizer_.setSpeedLimit(speed_limit, percentage);
}

}  // namespace nav2_mppi_controller
