// fleet_controller_server.hpp — the fleet-level caller.  The reference's controller_server
// owns ONE robot: per cycle it takes the robot pose, lets the path handler cut the local plan and calls the controller
// plugin (nav2_controller/src/controller_server.cpp:738-794 computeAndPublishVelocity).  A GPU serves many robots at
// once, so this is the same sequence for N robots with the N computeVelocityCommands calls gathered into ONE
// mppi_optimize_in_place (mppi::Optimizer::evalControlFleet): N costmaps read in place, one upload, one set of kernels.
#ifndef NAV2_CONTROLLER_B200__FLEET_CONTROLLER_SERVER_HPP_
#define NAV2_CONTROLLER_B200__FLEET_CONTROLLER_SERVER_HPP_

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "nav2_controller_b200/feasible_path_handler.hpp"
#include "nav2_mppi_controller_b200/mppi.hpp"

namespace nav2_controller
{

class FleetControllerServer
{
public:
  struct RobotCommand
  {
    bool valid{false};                 // false: this robot's cycle failed, `error` says why (the reference publishes a
    std::string error;                 // zero velocity and aborts the action, controller_server.cpp:616-690)
    geometry_msgs::msg::TwistStamped cmd;
    nav_msgs::msg::Path local_plan;    // what the path handler handed to the controller
  };

  // one costmap per robot; every robot shares the controller parameters under `name` (e.g. "FollowPath")
  void configure(
    const nav2::LifecycleNode::SharedPtr & node, const std::string & name,
    const std::vector<std::shared_ptr<nav2_costmap_2d::Costmap2DROS>> & costmaps, nav2::TransformBuffer::SharedPtr tf)
  {
    node_ = node;
    name_ = name;
    costmaps_ = costmaps;
    parameters_handler_ = std::make_unique<mppi::ParametersHandler>(node_, name_);
    optimizer_.setFleetSize(static_cast<uint32_t>(costmaps.size()));
    optimizer_.initialize(node_, name_, costmaps.front(), tf, parameters_handler_.get());
    path_handlers_.clear();
    for (auto & c : costmaps) {
      path_handlers_.emplace_back(std::make_unique<FeasiblePathHandler>());
      path_handlers_.back()->initialize(node_, "PathHandler", c, tf);
    }
  }
  size_t robots() const {return costmaps_.size();}
  void setPlan(size_t robot, const nav_msgs::msg::Path & path) {path_handlers_.at(robot)->setPlan(path);}
  mppi::Optimizer & optimizer() {return optimizer_;}
  FeasiblePathHandler & pathHandler(size_t robot) {return *path_handlers_.at(robot);}

  // computeAndPublishVelocity for the whole fleet: poses / speeds are indexed by robot
  std::vector<RobotCommand> computeVelocityCommands(
    const std::vector<geometry_msgs::msg::PoseStamped> & poses, const std::vector<geometry_msgs::msg::Twist> & speeds,
    nav2_core::GoalChecker * goal_checker)
  {
    const size_t R = costmaps_.size();
    std::vector<RobotCommand> out(R);
    std::vector<mppi::Optimizer::FleetInput> in(R);
    std::lock_guard<std::mutex> param_lock(*parameters_handler_->getLock());
    // every robot's costmap stays locked from the moment its path is cut to the end of the optimisation, like the
    // single-robot controller holds its costmap mutex around evalControl (nav2_mppi_controller/src/controller.cpp:104-107)
    std::vector<std::unique_lock<nav2_costmap_2d::Costmap2D::mutex_t>> locks;
    locks.reserve(R);
    for (size_t r = 0; r < R; ++r) {locks.emplace_back(*costmaps_[r]->getCostmap()->getMutex());}
    for (size_t r = 0; r < R; ++r) {
      try {
        auto segment = path_handlers_[r]->findPlanSegment(poses[r]);
        out[r].local_plan = path_handlers_[r]->transformLocalPlan(segment.start, segment.end);
        out[r].valid = true;
      } catch (const nav2_core::ControllerException & e) {
        out[r].error = e.what();
        // a robot without a usable plan still occupies its slot of the batch: it is fed its own pose as a one-pose plan
        out[r].local_plan.poses.assign(1, poses[r]);
      }
      in[r].robot_pose = poses[r];
      in[r].robot_speed = speeds[r];
      in[r].plan = &out[r].local_plan;
      in[r].goal = out[r].local_plan.poses.back().pose;
      if (out[r].valid) {
        try {in[r].goal = path_handlers_[r]->getTransformedGoal(poses[r].header.stamp).pose;} catch (const nav2_core::ControllerException &) {}
      }
      in[r].costmap = costmaps_[r]->getCostmap();
    }
    const auto results = optimizer_.evalControlFleet(in, goal_checker);
    for (size_t r = 0; r < R; ++r) {
      if (!out[r].valid) {continue;}
      if (results[r].status == MPPI_ERR_NO_VALID_CONTROL) {
        out[r].valid = false;
        out[r].error = "Optimizer fail to compute path";   // nav2_core::NoValidControl, optimizer.cpp:294
      } else if (results[r].status != MPPI_OK) {
        out[r].valid = false;
        out[r].error = mppi_status_string(results[r].status);
      } else {
        out[r].cmd = results[r].cmd;
      }
    }
    return out;
  }

private:
  nav2::LifecycleNode::SharedPtr node_;
  std::string name_;
  std::vector<std::shared_ptr<nav2_costmap_2d::Costmap2DROS>> costmaps_;
  std::unique_ptr<mppi::ParametersHandler> parameters_handler_;
  mppi::Optimizer optimizer_;
  std::vector<std::unique_ptr<FeasiblePathHandler>> path_handlers_;
};

}  // namespace nav2_controller

#endif  // NAV2_CONTROLLER_B200__FLEET_CONTROLLER_SERVER_HPP_
