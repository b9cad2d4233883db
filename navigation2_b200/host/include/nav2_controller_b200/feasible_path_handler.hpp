// feasible_path_handler.hpp — the caller side of the hot path: what controller_server does
// with the global plan before every computeVelocityCommands call.  Mirror of nav2_controller::FeasiblePathHandler
// (nav2_controller/plugins/feasible_path_handler.cpp; citations below are to that file): setPlan, findPlanSegment
// (closest pose within max_robot_pose_search_dist, prune_distance ahead), transformLocalPlan (into the costmap
// frame, cut at the costmap border, prune what was passed).  Host C++, O(path length) per cycle, like the reference.
// Not mirrored: enforce_path_inversion / enforce_path_rotation (both default false, :66-69); asking for them throws.
#ifndef NAV2_CONTROLLER_B200__FEASIBLE_PATH_HANDLER_HPP_
#define NAV2_CONTROLLER_B200__FEASIBLE_PATH_HANDLER_HPP_

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <utility>

#include "ros_shim/ros_shim.hpp"

namespace nav2_controller
{

using PathIterator = std::vector<geometry_msgs::msg::PoseStamped>::iterator;   // nav2_core/path_handler.hpp
struct PathSegment {PathIterator start; PathIterator end;};

class FeasiblePathHandler
{
public:
  void initialize(
    const nav2::LifecycleNode::WeakPtr & parent, const std::string & plugin_name,
    const std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, nav2::TransformBuffer::SharedPtr tf)
  {
    // :48-93
    auto node = parent.lock();
    plugin_name_ = plugin_name;
    costmap_ros_ = costmap_ros;
    tf_ = tf;
    auto declare_or_get = [&](const std::string & name, nav2::ParameterValue def) {
        const std::string full = plugin_name + "." + name;
        if (!node->has_parameter(full)) {node->declare_parameter(full, def);}
        return node->get_parameter(full);
      };
    reject_unit_path_ = std::get<bool>(declare_or_get("reject_unit_path", false));
    max_robot_pose_search_dist_ = std::get<double>(declare_or_get("max_robot_pose_search_dist", getCostmapMaxExtent()));
    prune_distance_ = std::get<double>(declare_or_get("prune_distance", 2.0));
    const bool inversion = std::get<bool>(declare_or_get("enforce_path_inversion", false));
    const bool rotation = std::get<bool>(declare_or_get("enforce_path_rotation", false));
    if (inversion || rotation) {
      throw nav2_core::ControllerException("enforce_path_inversion / enforce_path_rotation are not mirrored");
    }
    if (max_robot_pose_search_dist_ < 0.0) {max_robot_pose_search_dist_ = std::numeric_limits<double>::max();}
  }

  double getCostmapMaxExtent() const  // :95-101
  {
    const auto * c = costmap_ros_->getCostmap();
    return std::max(c->getSizeInMetersX(), c->getSizeInMetersY()) / 2.0;
  }

  void setPlan(const nav_msgs::msg::Path & path)  // :124-133
  {
    std::lock_guard<std::mutex> lock(mutex_);
    global_plan_ = path;
    global_plan_up_to_constraint_ = global_plan_;
  }
  const nav_msgs::msg::Path & getPlan() const {return global_plan_;}
  size_t remainingPoses() const {return global_plan_up_to_constraint_.poses.size();}

  geometry_msgs::msg::PoseStamped transformToGlobalPlanFrame(const geometry_msgs::msg::PoseStamped & pose)  // :135-157
  {
    if (global_plan_up_to_constraint_.poses.empty()) {throw nav2_core::InvalidPath("Received plan with zero length");}
    if (reject_unit_path_ && global_plan_up_to_constraint_.poses.size() == 1) {
      throw nav2_core::InvalidPath("Received plan with length of one");
    }
    geometry_msgs::msg::PoseStamped robot_pose;
    if (!transform(pose, robot_pose, global_plan_up_to_constraint_.header.frame_id)) {
      throw nav2_core::ControllerTFError("Unable to transform robot pose into global plan's frame");
    }
    return robot_pose;
  }

  PathSegment findPlanSegment(const geometry_msgs::msg::PoseStamped & pose)  // :159-198
  {
    std::lock_guard<std::mutex> lock(mutex_);
    global_pose_ = transformToGlobalPlanFrame(pose);
    auto & poses = global_plan_up_to_constraint_.poses;
    auto closest_pose_upper_bound = firstAfterIntegratedDistance(poses.begin(), poses.end(), max_robot_pose_search_dist_);
    // nav2_util::geometry_utils::min_by (geometry_utils.hpp:109-126): `<=`, so the LAST of equal minima
    auto closest_point = poses.begin();
    if (poses.begin() == closest_pose_upper_bound) {
      closest_point = closest_pose_upper_bound;
    } else {
      double lowest = euclidean(global_pose_, *poses.begin());
      for (auto it = std::next(poses.begin()); it != closest_pose_upper_bound; ++it) {
        const double d = euclidean(global_pose_, *it);
        if (d <= lowest) {lowest = d; closest_point = it;}
      }
    }
    // do not prune the plan below 2 points: the end-of-path direction needs them
    if (poses.size() > 1 && closest_point == std::prev(poses.end())) {closest_point = std::prev(closest_point);}
    auto pruned_plan_end = firstAfterIntegratedDistance(closest_point, poses.end(), prune_distance_);
    return {closest_point, pruned_plan_end};
  }

  nav_msgs::msg::Path transformLocalPlan(const PathIterator & closest_point, const PathIterator & pruned_plan_end)  // :201-252
  {
    std::lock_guard<std::mutex> lock(mutex_);
    nav_msgs::msg::Path transformed_plan;
    transformed_plan.header.frame_id = costmap_ros_->getGlobalFrameID();
    transformed_plan.header.stamp = global_pose_.header.stamp;
    unsigned int mx, my;
    for (auto global_plan_pose = closest_point; global_plan_pose != pruned_plan_end; ++global_plan_pose) {
      geometry_msgs::msg::PoseStamped costmap_plan_pose;
      global_plan_pose->header.stamp = global_pose_.header.stamp;
      global_plan_pose->header.frame_id = global_plan_.header.frame_id;
      transform(*global_plan_pose, costmap_plan_pose, costmap_ros_->getGlobalFrameID());
      if (!costmap_ros_->getCostmap()->worldToMap(costmap_plan_pose.pose.position.x, costmap_plan_pose.pose.position.y, mx, my)) {
        break;   // the first pose outside the costmap ends the local plan
      }
      transformed_plan.poses.push_back(costmap_plan_pose);
    }
    // path pruning: what lies behind the closest pose is not looked at again
    global_plan_up_to_constraint_.poses.erase(global_plan_up_to_constraint_.poses.begin(), closest_point);
    if (transformed_plan.poses.empty()) {throw nav2_core::InvalidPath("Resulting plan has 0 poses in it.");}
    return transformed_plan;
  }

  geometry_msgs::msg::PoseStamped getTransformedGoal(const builtin_interfaces::msg::Time & stamp)  // :254-270
  {
    auto goal = global_plan_.poses.back();
    goal.header.frame_id = global_plan_.header.frame_id;
    goal.header.stamp = stamp;
    geometry_msgs::msg::PoseStamped transformed_goal;
    if (!transform(goal, transformed_goal, costmap_ros_->getGlobalFrameID())) {
      throw nav2_core::ControllerTFError("Unable to transform goal pose into costmap frame");
    }
    return transformed_goal;
  }

private:
  static double euclidean(const geometry_msgs::msg::PoseStamped & a, const geometry_msgs::msg::PoseStamped & b)
  {
    return std::hypot(a.pose.position.x - b.pose.position.x, a.pose.position.y - b.pose.position.y);
  }
  // nav2_util::geometry_utils::first_after_integrated_distance (geometry_utils.hpp:102-119)
  static PathIterator firstAfterIntegratedDistance(PathIterator begin, PathIterator end, double distance)
  {
    if (begin == end) {return end;}
    double dist = 0.0;
    for (auto it = begin; it != end - 1; it++) {
      dist += euclidean(*it, *(it + 1));
      if (dist > distance) {return it + 1;}
    }
    return end;
  }
  bool transform(const geometry_msgs::msg::PoseStamped & in, geometry_msgs::msg::PoseStamped & out, const std::string & frame) const
  {
    if (!tf_) {out = in; out.header.frame_id = frame; return true;}
    return tf_->transform(in, out, frame);
  }

  std::string plugin_name_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  nav2::TransformBuffer::SharedPtr tf_;
  nav_msgs::msg::Path global_plan_, global_plan_up_to_constraint_;
  geometry_msgs::msg::PoseStamped global_pose_;
  std::mutex mutex_;
  bool reject_unit_path_{false};
  double max_robot_pose_search_dist_{0.0}, prune_distance_{2.0};
};

}  // namespace nav2_controller

#endif  // NAV2_CONTROLLER_B200__FEASIBLE_PATH_HANDLER_HPP_
