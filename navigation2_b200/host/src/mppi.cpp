// mppi.cpp — host mirror of nav2_mppi_controller (see mppi.hpp).  Reference citations are relative to
// /root/reference/nav2_mppi_controller/.
#include "nav2_mppi_controller_b200/mppi.hpp"

#include <algorithm>
#include <array>
#include <cstring>
#include <limits>

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

CriticFunction::~CriticFunction()
{
  if (score_handle_) {mppi_destroy(score_handle_);}
}

// CriticFunction::score on caller tensors (what test/critics_tests.cpp exercises): the critic's own parameter block
// on a one-critic device handle, mppi_score_tensors, results written back into CriticData.
void CriticFunction::score(CriticData & data)
{
  const uint32_t B = data.state.batch_size, T = data.state.time_steps;
  if (B == 0 || T == 0 || data.costs.size() != B) {throw nav2_core::ControllerException("CriticData tensors have no shape");}
  auto * costmap = costmap_ros_->getCostmap();
  if (!score_handle_ || score_batch_ != B || score_steps_ != T) {
    if (score_handle_) {mppi_destroy(score_handle_); score_handle_ = nullptr;}
    MppiConfig c{};
    mppi_default_config(&c);
    c.batch_size = B; c.time_steps = T; c.model_dt = data.model_dt;
    c.controller_frequency = 1.0 / static_cast<double>(data.model_dt);
    c.n_critics = 1; c.critics[0] = p_;
    c.visualize = 1;      // keep CriticData::trajectories_in_collision
    c.use_radius = costmap_ros_->getUseRadius() ? 1 : 0;
    const auto & fp = costmap_ros_->getRobotFootprint();
    c.n_footprint = static_cast<uint32_t>(std::min<size_t>(fp.size(), MPPI_MAX_FOOTPRINT));
    for (uint32_t i = 0; i < c.n_footprint; ++i) {c.footprint_x[i] = fp[i].x; c.footprint_y[i] = fp[i].y;}
    c.inscribed_radius = costmap_ros_->getLayeredCostmap()->getInscribedRadius();
    c.circumscribed_radius = costmap_ros_->getLayeredCostmap()->getCircumscribedRadius();
    c.has_inflation_layer = costmap_ros_->inflation().present ? 1 : 0;
    c.inflation_cost_scaling_factor = costmap_ros_->inflation().cost_scaling_factor;
    c.inflation_radius = costmap_ros_->inflation().inflation_radius;
    c.track_unknown = costmap_ros_->getLayeredCostmap()->isTrackingUnknown() ? 1 : 0;
    const int rc = mppi_create(&c, &score_handle_);
    if (rc != MPPI_OK) {
      score_handle_ = nullptr;
      throw nav2_core::ControllerException(std::string("CriticFunction::score: ") + mppi_status_string(rc) + ": " + mppi_last_error(nullptr));
    }
    score_batch_ = B; score_steps_ = T;
  }
  int rc = mppi_upload_costmap(
    score_handle_, 0, costmap->getCharMap(), costmap->getSizeInCellsX(), costmap->getSizeInCellsY(),
    costmap->getResolution(), costmap->getOriginX(), costmap->getOriginY());
  if (rc != MPPI_OK) {throw nav2_core::ControllerException(std::string("CriticFunction::score: ") + mppi_last_error(score_handle_));}
  const size_t n = data.path.poses.size();
  std::vector<double> px(n), py(n), pyaw(n);
  for (size_t i = 0; i < n; ++i) {
    px[i] = data.path.poses[i].pose.position.x;
    py[i] = data.path.poses[i].pose.position.y;
    pyaw[i] = tf2::getYaw(data.path.poses[i].pose.orientation);
  }
  MppiCycleIn in{};
  in.pose_x = data.state.pose.pose.position.x; in.pose_y = data.state.pose.pose.position.y;
  in.pose_yaw = tf2::getYaw(data.state.pose.pose.orientation);
  in.speed_vx = data.state.speed.linear.x; in.speed_vy = data.state.speed.linear.y; in.speed_wz = data.state.speed.angular.z;
  in.path_x = px.data(); in.path_y = py.data(); in.path_yaw = pyaw.data(); in.n_path = static_cast<uint32_t>(n);
  in.goal_x = data.goal.position.x; in.goal_y = data.goal.position.y; in.goal_yaw = tf2::getYaw(data.goal.orientation);
  in.has_local_path_length = 1;                      // the critic tests write state.local_path_length directly
  in.local_path_length = data.state.local_path_length;
  if (data.goal_checker != nullptr) {
    geometry_msgs::msg::Pose pose_tolerance;
    geometry_msgs::msg::Twist velocity_tolerance;
    double path_length_tolerance = 0.0;
    data.goal_checker->getTolerances(pose_tolerance, velocity_tolerance, path_length_tolerance);
    in.has_goal_checker = 1;
    in.goal_checker_xy_tolerance = pose_tolerance.position.x;
  }
  auto ptr = [&](const std::vector<float> & v) -> const float * {return v.size() == static_cast<size_t>(B) * T ? v.data() : nullptr;};
  std::vector<uint8_t> collisions(B, 0);
  int32_t fail = data.fail_flag ? 1 : 0, furthest = data.furthest_reached_path_point;
  rc = mppi_score_tensors(
    score_handle_, 0, &in, ptr(data.state.vx), ptr(data.state.vy), ptr(data.state.wz), ptr(data.trajectories.x),
    ptr(data.trajectories.y), ptr(data.trajectories.yaws), data.furthest_reached_path_point, data.costs.data(),
    collisions.data(), &fail, &furthest);
  if (rc != MPPI_OK) {throw nav2_core::ControllerException(std::string("CriticFunction::score: ") + mppi_last_error(score_handle_));}
  data.fail_flag = fail != 0;
  data.furthest_reached_path_point = furthest;
  if (data.trajectories_in_collision.size() == B) {
    for (uint32_t b = 0; b < B; ++b) {if (collisions[b]) {data.trajectories_in_collision[b] = true;}}
  }
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
  getParam(visualize_, "visualize", false);                                               // :41
  if (visualize_ && !critics_effect_pub_) {
    critics_effect_pub_ = std::make_unique<nav2::Publisher<nav2_msgs::msg::CriticsStats>>("~/critics_stats");  // :47-50
    critics_effect_pub_->on_activate();
  }
  critics_.clear();
  for (const auto & critic_name : critic_names_) {
    auto instance = createCritic(getFullName(critic_name));
    instance->on_configure(parent, name, name + "." + critic_name, costmap_ros, param_handler);
    critics_.push_back(std::move(instance));
  }
}

void CriticManager::publishCriticsStats(MppiHandle * handle, uint32_t batch_size, const builtin_interfaces::msg::Time & stamp)
{
  if (!visualize_ || !handle) {return;}
  const size_t n = critics_.size();
  std::vector<float> all(n * batch_size, 0.0f);
  if (n > 0 && mppi_get_tensor(handle, 0, MPPI_TENSOR_CRITIC_COSTS, all.data(), all.size() * sizeof(float)) != MPPI_OK) {
    throw std::runtime_error(std::string("mppi_get_tensor(CRITIC_COSTS): ") + mppi_last_error(handle));
  }
  uint32_t evaluated = static_cast<uint32_t>(n);
  mppi_get_critics_evaluated(handle, 0, &evaluated);
  auto stats_msg = std::make_unique<nav2_msgs::msg::CriticsStats>();
  critic_costs_.clear();
  critic_costs_.reserve(n);
  // critic_manager.cpp:89-117: only the critics the loop reached before fail_flag broke it are reported
  for (size_t i = 0; i < std::min<size_t>(evaluated, n); ++i) {
    std::vector<float> cost_diff(all.begin() + i * batch_size, all.begin() + (i + 1) * batch_size);
    float costs_sum = 0.0f;
    for (float v : cost_diff) {costs_sum += v;}
    stats_msg->critics.push_back(critic_names_[i]);
    stats_msg->costs_sum.push_back(costs_sum);
    stats_msg->changed.push_back(costs_sum != 0.0f);
    critic_costs_.emplace_back(critic_names_[i], std::move(cost_diff));
  }
  if (critics_effect_pub_) {
    stats_msg->stamp = stamp;
    critics_effect_pub_->publish(std::move(stats_msg));
  }
}

// ------------------------------------------------------------------ TrajectoryVisualizer

void TrajectoryVisualizer::on_configure(
  nav2::LifecycleNode::WeakPtr, const std::string & name, const std::string & frame_id, ParametersHandler * parameters_handler)
{
  // trajectory_visualizer.cpp:24-42
  frame_id_ = frame_id;
  trajectories_publisher_ = std::make_unique<nav2::Publisher<visualization_msgs::msg::MarkerArray>>("~/candidate_trajectories");
  optimal_path_pub_ = std::make_unique<nav2::Publisher<nav_msgs::msg::Path>>("~/optimal_path");
  auto getParam = parameters_handler->getParamGetter(name + ".TrajectoryVisualizer");
  getParam(trajectory_step_, "trajectory_step", 5);
  getParam(time_step_, "time_step", 3);
  reset();
}

void TrajectoryVisualizer::add(
  const std::vector<std::array<float, 3>> & trajectory, const std::string & marker_namespace,
  const builtin_interfaces::msg::Time & cmd_stamp)
{
  // trajectory_visualizer.cpp:61-108
  if (optimal_path_pub_->get_subscription_count() == 0 && trajectories_publisher_->get_subscription_count() == 0) {return;}
  const size_t size = trajectory.size();
  if (!size) {return;}
  optimal_path_->header.stamp = cmd_stamp;
  optimal_path_->header.frame_id = frame_id_;
  for (size_t i = 0; i < size; i++) {
    const float component = static_cast<float>(i) / static_cast<float>(size);
    visualization_msgs::msg::Marker marker;   // utils::createMarker (tools/utils.hpp:108-132)
    marker.header.frame_id = frame_id_;
    marker.ns = marker_namespace;
    marker.id = marker_id_++;
    marker.type = visualization_msgs::msg::Marker::SPHERE;
    marker.action = visualization_msgs::msg::Marker::ADD;
    marker.pose.position.x = trajectory[i][0];
    marker.pose.position.y = trajectory[i][1];
    marker.pose.position.z = 0.06;
    const bool last = i == size - 1;
    marker.scale.x = last ? 0.07 : 0.03; marker.scale.y = last ? 0.07 : 0.03; marker.scale.z = last ? 0.09 : 0.07;
    marker.color.r = 0.0f; marker.color.g = component; marker.color.b = component; marker.color.a = 1.0f;
    points_->markers.push_back(marker);
    geometry_msgs::msg::PoseStamped pose_stamped;
    pose_stamped.header.frame_id = frame_id_;
    pose_stamped.pose = marker.pose;
    pose_stamped.pose.orientation = tf2::quaternionFromYaw(trajectory[i][2]);
    optimal_path_->poses.push_back(pose_stamped);
  }
}

void TrajectoryVisualizer::add(
  const std::vector<float> & samples_xy, size_t n_trajectories, size_t n_points, const std::vector<float> & costs,
  const std::vector<bool> & collisions, const builtin_interfaces::msg::Time & stamp)
{
  // trajectory_visualizer.cpp:110-152
  if (trajectories_publisher_->get_subscription_count() == 0) {return;}
  const size_t n_rows = costs.size();
  if (n_trajectories == 0 || n_rows == 0 || samples_xy.size() < n_trajectories * n_points * 2) {return;}
  // normalise the costs excluding colliding trajectories
  float min_val = std::numeric_limits<float>::max();
  float max_val = std::numeric_limits<float>::lowest();
  for (size_t k = 0; k < costs.size(); ++k) {
    if (!collisions.empty() && k < collisions.size() && collisions[k]) {continue;}
    if (costs[k] < min_val) {min_val = costs[k];}
    if (costs[k] > max_val) {max_val = costs[k];}
  }
  if (max_val < min_val) {
    min_val = *std::min_element(costs.begin(), costs.end());
    max_val = *std::max_element(costs.begin(), costs.end());
  }
  const float range = max_val - min_val;
  for (size_t k = 0; k < n_trajectories; ++k) {
    const size_t i = k * static_cast<size_t>(trajectory_step_);   // the reference's `i += trajectory_step_`
    if (i >= n_rows) {break;}
    const float norm = (range > 0.0f) ? (costs[i] - min_val) / range : 0.0f;
    const bool in_collision = !collisions.empty() && i < collisions.size() && collisions[i];
    visualization_msgs::msg::Marker marker;   // addCostColoredTrajectory, :154-187
    marker.header.frame_id = frame_id_;
    marker.header.stamp = stamp;
    marker.ns = "Candidate Trajectories";
    marker.id = marker_id_++;
    marker.type = visualization_msgs::msg::Marker::LINE_STRIP;
    marker.action = visualization_msgs::msg::Marker::ADD;
    marker.pose.orientation.w = 1.0;
    marker.scale.x = 0.01;
    if (in_collision) {
      marker.color.r = 1.0f; marker.color.g = 0.0f; marker.color.b = 1.0f; marker.color.a = 0.6f;
    } else {
      marker.color = costToColor(norm);
    }
    marker.points.reserve(n_points);
    for (size_t j = 0; j < n_points; ++j) {
      geometry_msgs::msg::Point pt;
      pt.x = samples_xy[(k * n_points + j) * 2];
      pt.y = samples_xy[(k * n_points + j) * 2 + 1];
      pt.z = 0.03;
      marker.points.push_back(pt);
    }
    points_->markers.push_back(std::move(marker));
  }
}

std_msgs::msg::ColorRGBA TrajectoryVisualizer::costToColor(float normalized)
{
  normalized = std::clamp(normalized, 0.0f, 1.0f);
  std_msgs::msg::ColorRGBA c;
  if (normalized < 0.5f) {
    c.r = 2.0f * normalized; c.g = 1.0f;
  } else {
    c.r = 1.0f; c.g = 2.0f * (1.0f - normalized);
  }
  c.b = 0.0f; c.a = 0.8f;
  return c;
}

void TrajectoryVisualizer::reset()
{
  marker_id_ = 0;
  points_ = std::make_unique<visualization_msgs::msg::MarkerArray>();
  optimal_path_ = std::make_unique<nav_msgs::msg::Path>();
}

void TrajectoryVisualizer::visualize()
{
  if (trajectories_publisher_->get_subscription_count() > 0) {trajectories_publisher_->publish(std::move(points_));}
  if (optimal_path_pub_->get_subscription_count() > 0) {optimal_path_pub_->publish(std::move(optimal_path_));}
  reset();
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
  s.n_robots = n_robots_;
  auto getParam = parameters_handler_->getParamGetter(name_);
  auto getParentParam = parameters_handler_->getParamGetter("");
  if (!callbacks_registered_) {
    // Reject dynamic updates to the kinematic parameters while a speed limit is active (optimizer.cpp:85-103)
    auto kinematic_guard = [this](const std::string & param_name, const nav2::ParameterValue &, SetParametersResult & result) {
        if (isSpeedLimitActive()) {
          result.successful = false;
          if (!result.reason.empty()) {result.reason += "\n";}
          result.reason += "Rejected dynamic update to '" + param_name + "': speed limit is active. Clear the speed limit first.";
        }
      };
    for (const char * p : {"vx_max", "vx_min", "vy_max", "wz_max"}) {
      parameters_handler_->addPreCallback(name_ + "." + p, kinematic_guard);
    }
    parameters_handler_->addPostCallback([this]() {reconfigure();});   // optimizer.cpp:155 (reset() there)
    callbacks_registered_ = true;
  }
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
  fillDerivedSettings();
  auto & s = cfg_;

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

// the critics' parameter blocks and what they read from Costmap2DROS / LayeredCostmap / InflationLayer
void Optimizer::fillDerivedSettings()
{
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
}

// Post-callback of a dynamic parameter update (ParametersHandler::setDynamic holds the mutex the control loop takes):
// every setting is read again from the node's parameter store and the device is reconfigured; mppi_set_config ends
// with Optimizer::reset(), which is what the reference's post-callback is (optimizer.cpp:155).
void Optimizer::reconfigure()
{
  if (!handle_) {return;}
  getParams();
  fillDerivedSettings();
  const int rc = mppi_set_config(handle_, &cfg_);
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_set_config");}
  control_sequence_.vx.assign(cfg_.time_steps, 0.0f);
  control_sequence_.vy.assign(cfg_.time_steps, 0.0f);
  control_sequence_.wz.assign(cfg_.time_steps, 0.0f);
  speed_limit_active_ = false;   // reset() restores the base constraints
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
  // constraints differ from base_constraints exactly when a limit other than NO_SPEED_LIMIT / 100 % is set (:218-229)
  speed_limit_active_ = !(speed_limit == nav2_costmap_2d::NO_SPEED_LIMIT) &&
    !(percentage ? speed_limit == 100.0 : speed_limit == static_cast<double>(cfg_.vx_max));
}

std::vector<float> Optimizer::getCosts(uint32_t robot)
{
  std::vector<float> v(cfg_.batch_size);
  const int rc = mppi_get_tensor(handle_, robot, MPPI_TENSOR_COSTS, v.data(), v.size() * sizeof(float));
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_get_tensor(COSTS)");}
  return v;
}

std::vector<bool> Optimizer::getCollisionFlags(uint32_t robot)
{
  std::vector<uint8_t> raw(cfg_.batch_size);
  const int rc = mppi_get_tensor(handle_, robot, MPPI_TENSOR_COLLISIONS, raw.data(), raw.size());
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_get_tensor(COLLISIONS)");}
  return std::vector<bool>(raw.begin(), raw.end());
}

std::vector<float> Optimizer::getVisualizedTrajectories(size_t & n_trajectories, size_t & n_points, uint32_t robot)
{
  const uint32_t ts = std::max(cfg_.vis_trajectory_step, 1u), tt = std::max(cfg_.vis_time_step, 1u);
  n_trajectories = (cfg_.batch_size + ts - 1) / ts;
  n_points = (cfg_.time_steps + tt - 1) / tt;
  std::vector<float> v(n_trajectories * n_points * 2);
  const int rc = mppi_get_tensor(handle_, robot, MPPI_TENSOR_VIS_XY, v.data(), v.size() * sizeof(float));
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_get_tensor(VIS_XY)");}
  return v;
}

// Optimizer::evalControl for every robot of a fleet handle in one call
std::vector<Optimizer::FleetOutput> Optimizer::evalControlFleet(
  const std::vector<FleetInput> & inputs, nav2_core::GoalChecker * goal_checker)
{
  const uint32_t R = cfg_.n_robots, T = cfg_.time_steps;
  if (inputs.size() != R) {throw nav2_core::ControllerException("evalControlFleet: one input per robot of the fleet");}
  std::vector<MppiCostmapView> views(R);
  std::vector<MppiCycleIn> in(R);
  std::vector<MppiCycleOut> out(R);
  std::vector<std::vector<double>> px(R), py(R), pyaw(R);
  std::vector<std::vector<float>> traj(R, std::vector<float>(3 * T)), u(R, std::vector<float>(3 * T));
  double xy_tolerance = 0.0;
  if (goal_checker != nullptr) {
    geometry_msgs::msg::Pose pose_tolerance;
    geometry_msgs::msg::Twist velocity_tolerance;
    double path_length_tolerance = 0.0;
    goal_checker->getTolerances(pose_tolerance, velocity_tolerance, path_length_tolerance);
    xy_tolerance = pose_tolerance.position.x;
  }
  for (uint32_t r = 0; r < R; ++r) {
    const FleetInput & fi = inputs[r];
    if (!fi.plan || !fi.costmap) {throw nav2_core::ControllerException("evalControlFleet: robot without plan or costmap");}
    views[r] = MppiCostmapView{fi.costmap->getCharMap(), fi.costmap->getSizeInCellsX(), fi.costmap->getSizeInCellsY(),
      fi.costmap->getResolution(), fi.costmap->getOriginX(), fi.costmap->getOriginY()};
    const size_t n = fi.plan->poses.size();
    px[r].resize(n); py[r].resize(n); pyaw[r].resize(n);
    for (size_t i = 0; i < n; ++i) {
      px[r][i] = fi.plan->poses[i].pose.position.x;
      py[r][i] = fi.plan->poses[i].pose.position.y;
      pyaw[r][i] = tf2::getYaw(fi.plan->poses[i].pose.orientation);
    }
    MppiCycleIn & c = in[r];
    c = MppiCycleIn{};
    c.pose_x = fi.robot_pose.pose.position.x; c.pose_y = fi.robot_pose.pose.position.y;
    c.pose_yaw = tf2::getYaw(fi.robot_pose.pose.orientation);
    c.speed_vx = fi.robot_speed.linear.x; c.speed_vy = fi.robot_speed.linear.y; c.speed_wz = fi.robot_speed.angular.z;
    c.path_x = px[r].data(); c.path_y = py[r].data(); c.path_yaw = pyaw[r].data(); c.n_path = static_cast<uint32_t>(n);
    c.goal_x = fi.goal.position.x; c.goal_y = fi.goal.position.y; c.goal_yaw = tf2::getYaw(fi.goal.orientation);
    c.has_goal_checker = goal_checker != nullptr ? 1 : 0;
    c.goal_checker_xy_tolerance = xy_tolerance;
    out[r] = MppiCycleOut{};
    out[r].control_vx = u[r].data(); out[r].control_vy = u[r].data() + T; out[r].control_wz = u[r].data() + 2 * T;
    out[r].optimal_trajectory = traj[r].data();
  }
  const int rc = mppi_optimize_in_place(handle_, views.data(), in.data(), out.data());
  if (rc != MPPI_OK) {throwStatus(rc, "mppi_optimize_in_place");}
  std::vector<FleetOutput> results(R);
  for (uint32_t r = 0; r < R; ++r) {
    FleetOutput & fo = results[r];
    fo.status = out[r].status;
    fo.cmd.header.frame_id = costmap_ros_->getBaseFrameID();
    fo.cmd.header.stamp = inputs[r].plan->header.stamp;
    fo.cmd.twist.linear.x = out[r].cmd_vx;
    fo.cmd.twist.angular.z = out[r].cmd_wz;
    if (isHolonomic()) {fo.cmd.twist.linear.y = out[r].cmd_vy;}
    fo.optimal_trajectory.resize(T);
    for (uint32_t i = 0; i < T; ++i) {fo.optimal_trajectory[i] = {traj[r][i], traj[r][T + i], traj[r][2 * T + i]};}
  }
  return results;
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
  trajectory_visualizer_.on_configure(parent_, name_, costmap_ros_->getGlobalFrameID(), parameters_handler_.get());
  if (publish_optimal_trajectory_) {
    opt_traj_pub_ = std::make_unique<nav2::Publisher<nav_msgs::msg::Trajectory>>("~/optimal_trajectory");
  }
}

void MPPIController::cleanup()  // controller.cpp:59-67
{
  optimizer_.shutdown();
  trajectory_visualizer_.on_cleanup();
  parameters_handler_.reset();
  opt_traj_pub_.reset();
}

void MPPIController::activate()  // controller.cpp:69-78
{
  trajectory_visualizer_.on_activate();
  parameters_handler_->start();
  if (opt_traj_pub_) {opt_traj_pub_->on_activate();}
}

void MPPIController::deactivate()  // controller.cpp:80-87
{
  trajectory_visualizer_.on_deactivate();
  if (opt_traj_pub_) {opt_traj_pub_->on_deactivate();}
}

// controller.cpp:139-158, plus the CriticsStats message CriticManager publishes inside evalTrajectoriesScores
void MPPIController::visualize(
  const builtin_interfaces::msg::Time & cmd_stamp, const std::vector<std::array<float, 3>> & optimal_trajectory)
{
  auto & critic_manager = optimizer_.getCriticManager();
  critic_manager.publishCriticsStats(optimizer_.handle(), optimizer_.getSettings().batch_size, cmd_stamp);
  const auto & critic_costs = critic_manager.getCriticCosts();
  const std::vector<float> costs =
    (critic_index_to_visualize_ <= 0 || critic_index_to_visualize_ > static_cast<int>(critic_costs.size())) ?
    optimizer_.getCosts() : critic_costs[critic_index_to_visualize_ - 1].second;
  size_t n_trajectories = 0, n_points = 0;
  const std::vector<float> samples = optimizer_.getVisualizedTrajectories(n_trajectories, n_points);
  trajectory_visualizer_.add(samples, n_trajectories, n_points, costs, optimizer_.getCollisionFlags(), cmd_stamp);
  trajectory_visualizer_.add(optimal_trajectory, "Optimal Trajectory", cmd_stamp);
  trajectory_visualizer_.visualize();
}

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
  if (publish_optimal_trajectory_ && opt_traj_pub_ && opt_traj_pub_->get_subscription_count() > 0) {  // controller.cpp:117-129
    std_msgs::msg::Header trajectory_header;
    trajectory_header.stamp = cmd.header.stamp;
    trajectory_header.frame_id = costmap_ros_->getGlobalFrameID();
    opt_traj_pub_->publish(mppi::utils::toTrajectoryMsg(
        optimal_trajectory_, optimizer_.getOptimalControlSequence(), optimizer_.getSettings().model_dt, trajectory_header));
  }
  if (visualize_) {
    visualize(cmd.header.stamp, optimal_trajectory_);
  }
  return cmd;
}

void MPPIController::newPathReceived(const nav_msgs::msg::Path &) {}  // controller.cpp:160-162

void MPPIController::setSpeedLimit(const double & speed_limit, const bool & percentage)  // controller.cpp:164-167
{
  optimizer_.setSpeedLimit(speed_limit, percentage);
}

}  // namespace nav2_mppi_controller
