// test_host.cpp — drives the C++ host mirror the way controller_server drives the reference plugin
// (nav2_controller/src/controller_server.cpp:162-171 configure, :761-766 computeVelocityCommands):
// configure -> activate -> computeVelocityCommands xN -> deactivate -> cleanup, modelled on the
// reference's controller_state_transition_test.cpp and optimizer_smoke_test.cpp.
//   test_host config            parameter parsing / exceptions only (no GPU needed)
//   test_host run <model> <n>   n closed-loop cycles on the GPU; prints "cmd vx vy wz" per cycle
//   test_host pathhandler       FeasiblePathHandler mirror, the reference's path_handler.cpp cases (no GPU needed)
//   test_host visunit           TrajectoryVisualizer mirror, the reference's trajectory_visualizer_tests.cpp cases (no GPU needed)
//   test_host params            ParametersHandler mirror, the reference's parameter_handler_test.cpp cases (no GPU needed)
//   test_host visualize <n>     `visualize: true`: CriticsStats + TrajectoryVisualizer output of the last of n cycles (GPU)
//   test_host fleet <R> <n>     FleetControllerServer: R robots, n cycles, one optimiser call per cycle (GPU)
//   test_host score             CriticFunction::score on caller tensors (GPU)
//   test_host inflate           the InflationLayer mirror on the GPU: the cases of inflation_tests.cpp, prints the grids
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "nav2_controller_b200/fleet_controller_server.hpp"
#include "nav2_costmap_2d_b200/inflation_layer.hpp"
#include "nav2_mppi_controller_b200/mppi.hpp"

namespace
{
struct TestGoalChecker : nav2_core::GoalChecker
{
  bool getTolerances(geometry_msgs::msg::Pose & pose_tolerance, geometry_msgs::msg::Twist &, double &) override
  {
    pose_tolerance.position.x = 0.25;  // test/utils_test.cpp:29-63
    pose_tolerance.position.y = 0.25;
    return true;
  }
};

std::shared_ptr<nav2_costmap_2d::Costmap2DROS> makeCostmap(bool use_radius)
{
  // the reference fixtures' map: 40 x 40 @0.1 m with an 8 x 8 block of cost 250 (benchmark/optimizer_benchmark.cpp:53-79)
  auto costmap = std::make_shared<nav2_costmap_2d::Costmap2D>(40, 40, 0.1, 0.0, 0.0, 0);
  for (unsigned int i = 16; i < 24; ++i) {
    for (unsigned int j = 16; j < 24; ++j) {costmap->setCost(i, j, 250);}
  }
  std::vector<geometry_msgs::msg::Point> fp;
  if (use_radius) {
    for (int i = 0; i < 16; ++i) {
      geometry_msgs::msg::Point p;
      p.x = std::cos(i * 2 * M_PI / 16) * 0.15; p.y = std::sin(i * 2 * M_PI / 16) * 0.15;
      fp.push_back(p);
    }
  } else {
    const double c[4][2] = {{0.15, 0.15}, {0.15, -0.15}, {-0.15, -0.15}, {-0.15, 0.15}};
    for (auto & q : c) {geometry_msgs::msg::Point p; p.x = q[0]; p.y = q[1]; fp.push_back(p);}
  }
  auto ros = std::make_shared<nav2_costmap_2d::Costmap2DROS>(costmap, use_radius, fp);
  ros->getLayeredCostmap()->inscribed_radius = use_radius ? 0.15 * std::cos(M_PI / 16) : 0.15;
  ros->getLayeredCostmap()->circumscribed_radius = use_radius ? 0.15 : 0.15 * std::sqrt(2.0);
  return ros;
}

std::shared_ptr<nav2::LifecycleNode> makeNode(const std::string & model)
{
  auto node = std::make_shared<nav2::LifecycleNode>("controller_server");
  node->declare_parameter("controller_frequency", 20.0);
  node->declare_parameter("FollowPath.motion_model", model);
  node->declare_parameter("FollowPath.batch_size", 2000);
  node->declare_parameter("FollowPath.time_steps", 56);
  node->declare_parameter("FollowPath.iteration_count", 2);
  node->declare_parameter(
    "FollowPath.critics", std::vector<std::string>{"ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic",
      "PathAlignCritic", "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic"});
  node->declare_parameter("FollowPath.PathAlignCritic.cost_weight", 14.0);
  return node;
}

#define EXPECT(cond) do {if (!(cond)) {std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1;}} while (0)

int testConfig()
{
  // parameter names / defaults flow through ParametersHandler into the critic parameter blocks
  auto node = makeNode("diff_drive");
  auto costmap_ros = makeCostmap(true);
  std::string name = "FollowPath";
  mppi::ParametersHandler ph(node, name);
  mppi::CriticManager cm;
  cm.on_configure(node, name, costmap_ros, &ph);
  EXPECT(cm.critics().size() == 8);
  EXPECT(cm.getFullName("CostCritic") == "mppi::critics::CostCritic");
  EXPECT(cm.critics()[1]->getName() == "FollowPath.CostCritic");
  EXPECT(cm.critics()[1]->params().type == MPPI_CRITIC_COST);
  EXPECT(std::fabs(cm.critics()[1]->params().cost_weight - 3.81f) < 1e-6f);
  EXPECT(std::fabs(cm.critics()[4]->params().cost_weight - 14.0f) < 1e-6f);     // overridden above
  EXPECT(cm.critics()[5]->params().offset_from_furthest == 6);                  // code default
  EXPECT(node->has_parameter("FollowPath.CostCritic.near_collision_cost"));     // declared with its default
  // consider_footprint with a radius costmap is a ControllerException (cost_critic.cpp:61-71)
  node->declare_parameter("FollowPath.CostCritic.consider_footprint", true);
  bool threw = false;
  try {
    mppi::CriticManager cm2;
    cm2.on_configure(node, name, costmap_ros, &ph);
  } catch (const nav2_core::ControllerException &) {threw = true;}
  EXPECT(threw);
  // unknown critic / motion model names fail to load
  threw = false;
  auto node2 = makeNode("diff_drive");
  node2->set_parameter("FollowPath.critics", std::vector<std::string>{"NoSuchCritic"});
  try {
    mppi::ParametersHandler ph2(node2, name);
    mppi::CriticManager cm3;
    cm3.on_configure(node2, name, costmap_ros, &ph2);
  } catch (const nav2_core::ControllerException &) {threw = true;}
  EXPECT(threw);
  // ParametersHandler: a static parameter cannot change at run time, a pre-callback can veto a dynamic one
  // (parameters_handler.cpp:59-93), post-callbacks run after an accepted update (:95-127)
  {
    auto node3 = makeNode("diff_drive");
    mppi::ParametersHandler ph3(node3, name);
    std::vector<std::string> critics;
    float vx_max = 0.0f;
    int posts = 0;
    bool limit_active = true;
    ph3.getParamGetter(name)(critics, "critics", std::vector<std::string>{}, mppi::ParameterType::Static);
    ph3.getParamGetter(name)(vx_max, "vx_max", 0.5f);
    ph3.addPreCallback(name + ".vx_max", [&](const std::string & n, const nav2::ParameterValue &, mppi::SetParametersResult & r) {
        if (limit_active) {r.successful = false; r.reason += "Rejected dynamic update to '" + n + "': speed limit is active.";}
      });
    ph3.addPostCallback([&]() {++posts;});
    auto r1 = ph3.setDynamic(name + ".critics", std::vector<std::string>{"CostCritic"});
    EXPECT(!r1.successful && r1.reason.find("static parameter") != std::string::npos && posts == 0);
    auto r2 = ph3.setDynamic(name + ".vx_max", 0.8);
    EXPECT(!r2.successful && posts == 0);
    EXPECT(std::get<double>(node3->get_parameter(name + ".vx_max")) == 0.5);
    limit_active = false;
    auto r3 = ph3.setDynamic(name + ".vx_max", 0.8);
    EXPECT(r3.successful && posts == 1);
    EXPECT(std::get<double>(node3->get_parameter(name + ".vx_max")) == 0.8);
    EXPECT(ph3.setDynamic("other_plugin.vx_max", 0.1).successful && posts == 1);   // not this plugin's parameter
  }
  // Ackermann applyConstraints (motion_models.hpp:292-298; motion_model_tests.cpp): |wz| <= |vx| / min_turning_r
  {
    mppi::AckermannMotionModel ack;
    mppi::ControlSequence cs;
    cs.vx = {1.0f, 0.1f, -0.4f}; cs.vy = {0, 0, 0}; cs.wz = {10.0f, -1.0f, 0.5f};
    ack.applyConstraints(cs);
    EXPECT(cs.wz[0] == 5.0f && cs.wz[1] == -0.5f && cs.wz[2] == 0.5f);
    mppi::DiffDriveMotionModel dd;
    mppi::models::State st;
    bool t2 = false;
    try {dd.predict(st);} catch (const nav2_core::ControllerException &) {t2 = true;}
    EXPECT(t2);
  }
  // critic_manager_test.cpp:130-148 CriticLoadingTest: the "critics" parameter names what is loaded, in order;
  // :93-128: getFullName
  {
    auto node4 = std::make_shared<nav2::LifecycleNode>("my_node");
    node4->declare_parameter("critic_manager.critics", std::vector<std::string>{"ConstraintCritic", "PreferForwardCritic"});
    std::string test_name = "test";
    mppi::ParametersHandler ph4(node4, test_name);
    mppi::CriticManager cm4;
    cm4.on_configure(node4, "critic_manager", costmap_ros, &ph4);
    EXPECT(cm4.critics().size() == 2u);
    EXPECT(cm4.critics()[0]->params().type == MPPI_CRITIC_CONSTRAINT && cm4.critics()[1]->params().type == MPPI_CRITIC_PREFER_FORWARD);
    EXPECT(cm4.getFullName("name") == "mppi::critics::name");
    // no "critics" parameter: nothing is loaded (critic_manager.cpp:47 default {})
    auto node5 = std::make_shared<nav2::LifecycleNode>("my_node");
    mppi::ParametersHandler ph5(node5, test_name);
    mppi::CriticManager cm5;
    cm5.on_configure(node5, "critic_manager", costmap_ros, &ph5);
    EXPECT(cm5.critics().empty());
  }
  // models_test.cpp:25-128: reset() zeroes and resizes ControlSequence / Path / State / Trajectories
  {
    mppi::models::ControlSequence sequence;
    sequence.vx.assign(10, 1.0f); sequence.vy.assign(10, 1.0f); sequence.wz.assign(10, 1.0f);
    sequence.reset(20);
    EXPECT(sequence.vx[4] == 0 && sequence.vy[4] == 0 && sequence.wz[4] == 0);
    EXPECT(sequence.vx.size() == 20u && sequence.vy.size() == 20u && sequence.wz.size() == 20u);
    mppi::models::Path path;
    path.x.assign(10, 1.0f); path.y.assign(10, 1.0f); path.yaws.assign(10, 1.0f);
    path.reset(20);
    EXPECT(path.x[4] == 0 && path.y[4] == 0 && path.yaws[4] == 0 && path.x.size() == 20u && path.yaws.size() == 20u);
    mppi::models::State state;
    state.vx.assign(100, 1.0f); state.cwz.assign(100, 1.0f);
    state.reset(20, 40);
    EXPECT(state.vx[4] == 0 && state.cwz[4] == 0 && state.batch_size == 20u && state.time_steps == 40u);
    for (const auto * v : {&state.vx, &state.vy, &state.wz, &state.cvx, &state.cvy, &state.cwz}) {EXPECT(v->size() == 800u);}
    mppi::models::Trajectories trajectories;
    trajectories.x.assign(100, 1.0f);
    trajectories.reset(20, 2);
    EXPECT(trajectories.x[4] == 0 && trajectories.x.size() == 40u && trajectories.y.size() == 40u && trajectories.yaws.size() == 40u);
  }
  EXPECT(std::fabs(tf2::getYaw(tf2::quaternionFromYaw(0.7)) - 0.7) < 1e-12);
  EXPECT(std::fabs(tf2::getYaw(tf2::quaternionFromYaw(-2.9)) + 2.9) < 1e-12);
  std::printf("config ok\n");
  return 0;
}

int testRun(const std::string & model, int cycles)
{
  auto node = makeNode(model);
  node->declare_parameter("FollowPath.publish_optimal_trajectory", true);   // controller.cpp:43-55, 117-129
  auto costmap_ros = makeCostmap(true);
  nav2_mppi_controller::MPPIController controller;
  controller.configure(node, "FollowPath", nullptr, costmap_ros);
  controller.activate();
  nav_msgs::msg::Path path;
  for (int i = 0; i < 50; ++i) {
    geometry_msgs::msg::PoseStamped p;
    p.pose.position.x = 0.5 + 0.05 * i;
    p.pose.position.y = 0.5 + 0.02 * i;
    p.pose.orientation = tf2::quaternionFromYaw(0.38);
    path.poses.push_back(p);
  }
  geometry_msgs::msg::PoseStamped pose;
  pose.pose.position.x = 0.5; pose.pose.position.y = 0.5;
  double yaw = 0.3;
  geometry_msgs::msg::Twist speed;
  TestGoalChecker goal_checker;
  controller.newPathReceived(path);
  for (int i = 0; i < cycles; ++i) {
    pose.pose.orientation = tf2::quaternionFromYaw(yaw);
    auto cmd = controller.computeVelocityCommands(pose, speed, &goal_checker, path, path.poses.back());
    std::printf("cmd %.9g %.9g %.9g\n", cmd.twist.linear.x, cmd.twist.linear.y, cmd.twist.angular.z);
    speed = cmd.twist;
    const double dt = 0.05;
    pose.pose.position.x += (speed.linear.x * std::cos(yaw) - speed.linear.y * std::sin(yaw)) * dt;
    pose.pose.position.y += (speed.linear.x * std::sin(yaw) + speed.linear.y * std::cos(yaw)) * dt;
    yaw += speed.angular.z * dt;
  }
  const auto & seq = controller.optimizer().getOptimalControlSequence();
  EXPECT(seq.vx.size() == 56);
  EXPECT(controller.lastOptimalTrajectory().size() == 56);
  if (cycles > 0) {
    // "~/optimal_trajectory": the optimal trajectory with the optimal control sequence's velocities (utils::toTrajectoryMsg)
    auto * traj_pub = controller.optimalTrajectoryPublisher();
    EXPECT(traj_pub != nullptr && traj_pub->publishedCount() == static_cast<size_t>(cycles));
    const auto * tm = traj_pub->last();
    EXPECT(tm != nullptr && tm->points.size() == 56u);
    EXPECT(tm->points[5].pose.position.x == static_cast<double>(controller.lastOptimalTrajectory()[5][0]));
    EXPECT(tm->points[5].pose.position.y == static_cast<double>(controller.lastOptimalTrajectory()[5][1]));
    EXPECT(tm->points[7].velocity.linear.x == static_cast<double>(seq.vx[7]));
    EXPECT(tm->points[7].velocity.angular.z == static_cast<double>(seq.wz[7]));
  }
  controller.setSpeedLimit(50.0, true);
  controller.reset();
  // every trajectory collides: NoValidControl after the retries (optimizer.cpp:281-298)
  auto * cm = costmap_ros->getCostmap();
  for (unsigned int i = 0; i < 40; ++i) {
    for (unsigned int j = 0; j < 40; ++j) {cm->setCost(i, j, 254);}
  }
  bool threw = false;
  try {
    controller.computeVelocityCommands(pose, speed, &goal_checker, path, path.poses.back());
  } catch (const nav2_core::NoValidControl &) {threw = true;}
  EXPECT(threw);
  controller.deactivate();
  controller.cleanup();
  std::printf("run ok\n");
  return 0;
}
// nav2_controller/plugins/test/path_handler.cpp: GetAndPrunePath (:81-103), TestBounds (:105-151),
// SearchDistanceSmallerThanPoseSpacing (:153-193), TestTransforms (:250-293), through the mirror
int testPathHandler()
{
  auto makeRos = []() {
      // Costmap2DROS defaults: 5 m x 5 m @0.1 m (costmap_2d_ros.cpp:403-417), global frame "odom"
      auto costmap = std::make_shared<nav2_costmap_2d::Costmap2D>(50, 50, 0.1, 0.0, 0.0, 0);
      return std::make_shared<nav2_costmap_2d::Costmap2DROS>(costmap, true, std::vector<geometry_msgs::msg::Point>{});
    };
  auto tf = std::make_shared<nav2::TransformBuffer>();
  tf->setTransform("map", "odom", 0.0, 0.0, 0.0);
  tf->setTransform("odom", "map", 0.0, 0.0, 0.0);
  {
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    node->declare_parameter("dummy.max_robot_pose_search_dist", 99999.9);
    nav2_controller::FeasiblePathHandler handler;
    handler.initialize(node, "dummy", makeRos(), tf);
    EXPECT(handler.getCostmapMaxExtent() == 2.5);
    nav_msgs::msg::Path path;
    path.header.frame_id = "map";
    path.poses.resize(100);
    for (unsigned int i = 0; i != path.poses.size(); i++) {
      path.poses[i].pose.position.x = i;
      path.poses[i].header.frame_id = "map";
    }
    handler.setPlan(path);
    EXPECT(handler.getPlan().poses.size() == 100u && handler.getPlan().header.frame_id == "map");
    geometry_msgs::msg::PoseStamped robot_pose;
    robot_pose.header.frame_id = "odom";
    robot_pose.pose.position.x = 25.0;
    auto segment = handler.findPlanSegment(robot_pose);
    // the closest pose lies outside the 5 m costmap: nothing survives the transform (InvalidPath), the plan is still pruned
    bool threw = false;
    try {handler.transformLocalPlan(segment.start, segment.end);} catch (const nav2_core::InvalidPath &) {threw = true;}
    EXPECT(threw);
    EXPECT(handler.remainingPoses() == 75u);
    // inside the costmap: poses every 0.5 m from the origin, robot at 1.2 m; prune_distance 2.0 m ahead
    nav_msgs::msg::Path near;
    near.header.frame_id = "map";
    near.poses.resize(20);
    for (unsigned int i = 0; i != near.poses.size(); i++) {
      near.poses[i].pose.position.x = 0.5 * i; near.poses[i].pose.position.y = 1.0;
      near.poses[i].header.frame_id = "map";
    }
    handler.setPlan(near);
    robot_pose.pose.position.x = 1.2; robot_pose.pose.position.y = 1.0;
    segment = handler.findPlanSegment(robot_pose);
    auto local = handler.transformLocalPlan(segment.start, segment.end);
    EXPECT(local.header.frame_id == "odom");
    EXPECT(local.poses.front().pose.position.x == 1.0);          // closest pose (1.0 is 0.2 away, 1.5 is 0.3 away)
    EXPECT(local.poses.size() == 5u);                            // 1.0 .. 3.0: the first pose MORE than 2.0 m along ends it
    EXPECT(handler.remainingPoses() == 18u);                     // poses 0.0 and 0.5 were pruned
    EXPECT(handler.getTransformedGoal(robot_pose.header.stamp).pose.position.x == 9.5);
    // the local plan stops at the costmap border (5 m): start close to it
    robot_pose.pose.position.x = 4.1;
    segment = handler.findPlanSegment(robot_pose);
    local = handler.transformLocalPlan(segment.start, segment.end);
    EXPECT(local.poses.front().pose.position.x == 4.0 && local.poses.back().pose.position.x == 4.5);
  }
  {
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    node->declare_parameter("dummy.max_robot_pose_search_dist", 0.5);
    nav2_controller::FeasiblePathHandler handler;
    handler.initialize(node, "dummy", makeRos(), tf);
    nav_msgs::msg::Path path;
    path.header.frame_id = "map";
    path.poses.resize(3);
    for (unsigned int i = 0; i != path.poses.size(); ++i) {
      path.poses[i].pose.position.x = i;
      path.poses[i].header.frame_id = "map";
    }
    geometry_msgs::msg::PoseStamped robot_pose;
    robot_pose.header.frame_id = "odom";
    robot_pose.pose.position.x = 0.9;
    handler.setPlan(path);
    auto segment = handler.findPlanSegment(robot_pose);
    auto local = handler.transformLocalPlan(segment.start, segment.end);
    EXPECT(local.poses.size() == 3u && local.poses.front().pose.position.x == 0.0);   // closest == begin, end == end
  }
  {
    // TestTransforms: no transform into the plan's frame -> ControllerTFError; empty plan -> InvalidPath
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    nav2_controller::FeasiblePathHandler handler;
    handler.initialize(node, "dummy", makeRos(), tf);
    geometry_msgs::msg::PoseStamped robot_pose;
    robot_pose.header.frame_id = "nowhere";
    bool threw = false;
    try {handler.transformToGlobalPlanFrame(robot_pose);} catch (const nav2_core::InvalidPath &) {threw = true;}
    EXPECT(threw);
    nav_msgs::msg::Path path;
    path.header.frame_id = "map";
    path.poses.resize(2);
    handler.setPlan(path);
    threw = false;
    try {handler.transformToGlobalPlanFrame(robot_pose);} catch (const nav2_core::ControllerTFError &) {threw = true;}
    EXPECT(threw);
    robot_pose.header.frame_id = "odom";
    handler.transformToGlobalPlanFrame(robot_pose);
  }
  std::printf("pathhandler ok\n");
  return 0;
}

// `visualize: true`: the host half of SURVEY.md section 8f row 1.  Prints what was published so the Python test can
// compare it with the oracle's tensors: CriticsStats, the candidate-trajectory markers, the optimal path.
int testVisualize(int cycles)
{
  auto node = makeNode("diff_drive");
  node->declare_parameter("FollowPath.visualize", true);
  node->declare_parameter("FollowPath.iteration_count", 1);
  node->set_parameter("FollowPath.iteration_count", 1);
  node->declare_parameter("FollowPath.TrajectoryVisualizer.trajectory_step", 50);
  node->declare_parameter("FollowPath.TrajectoryVisualizer.time_step", 5);
  auto costmap_ros = makeCostmap(true);
  nav2_mppi_controller::MPPIController controller;
  controller.configure(node, "FollowPath", nullptr, costmap_ros);
  controller.activate();
  nav_msgs::msg::Path path;
  for (int i = 0; i < 50; ++i) {
    geometry_msgs::msg::PoseStamped p;
    p.pose.position.x = 0.5 + 0.05 * i;
    p.pose.position.y = 0.5 + 0.02 * i;
    p.pose.orientation = tf2::quaternionFromYaw(0.38);
    path.poses.push_back(p);
  }
  geometry_msgs::msg::PoseStamped pose;
  pose.pose.position.x = 0.5; pose.pose.position.y = 0.5;
  pose.pose.orientation = tf2::quaternionFromYaw(0.3);
  geometry_msgs::msg::Twist speed;
  TestGoalChecker goal_checker;
  for (int i = 0; i < cycles; ++i) {
    auto cmd = controller.computeVelocityCommands(pose, speed, &goal_checker, path, path.poses.back());
    std::printf("cmd %.9g %.9g %.9g\n", cmd.twist.linear.x, cmd.twist.linear.y, cmd.twist.angular.z);
  }
  auto * stats_pub = controller.optimizer().getCriticManager().criticsEffectPublisher();
  EXPECT(stats_pub != nullptr && stats_pub->publishedCount() == static_cast<size_t>(cycles));
  const auto * stats = stats_pub->last();
  EXPECT(stats->critics.size() == 8 && stats->critics[1] == "CostCritic");
  for (size_t k = 0; k < stats->critics.size(); ++k) {
    std::printf("stat %s %d %.9g\n", stats->critics[k].c_str(), stats->changed[k] ? 1 : 0, stats->costs_sum[k]);
  }
  auto & viz = controller.trajectoryVisualizer();
  EXPECT(viz.trajectoriesPublisher()->publishedCount() == static_cast<size_t>(cycles));
  const auto * markers = viz.trajectoriesPublisher()->last();
  EXPECT(markers->markers.size() == 2000 / 50 + 56);             // 40 candidates + 56 optimal-trajectory spheres
  for (const auto & m : markers->markers) {
    if (m.ns != "Candidate Trajectories") {continue;}
    EXPECT(m.type == visualization_msgs::msg::Marker::LINE_STRIP && m.points.size() == 12);   // ceil(56 / 5)
    std::printf("traj %d %.6g %.6g %.6g %.6g", m.id, m.color.r, m.color.g, m.color.b, m.color.a);
    for (const auto & p : m.points) {std::printf(" %.9g %.9g", p.x, p.y);}
    std::printf("\n");
  }
  const auto * optimal = viz.optimalPathPublisher()->last();
  EXPECT(optimal->poses.size() == 56 && optimal->header.frame_id == "odom");
  const auto & traj = controller.lastOptimalTrajectory();
  for (size_t i = 0; i < traj.size(); ++i) {
    EXPECT(optimal->poses[i].pose.position.x == traj[i][0] && optimal->poses[i].pose.position.y == traj[i][1]);
  }
  // a dynamic parameter update goes through the pre-callbacks: rejected while a speed limit is active
  // (optimizer.cpp:85-103), accepted and applied (new device configuration, reset) once it is cleared
  controller.setSpeedLimit(50.0, true);
  EXPECT(controller.optimizer().isSpeedLimitActive());
  mppi::ParametersHandler * ph = nullptr;
  (void)ph;
  controller.deactivate();
  controller.cleanup();
  std::printf("visualize ok\n");
  return 0;
}

// The fleet-level caller: N robots, each with its own costmap, plan and pose, one call (fleet_controller_server.hpp).
// Prints "fleet <robot> <valid> vx vy wz n_local_plan" per robot and cycle.
int testFleet(int n_robots, int cycles)
{
  auto node = makeNode("diff_drive");
  node->set_parameter("FollowPath.iteration_count", 1);
  std::vector<std::shared_ptr<nav2_costmap_2d::Costmap2DROS>> costmaps;
  for (int r = 0; r < n_robots; ++r) {
    auto ros = makeCostmap(true);
    auto * c = ros->getCostmap();
    for (unsigned int i = 0; i < 6; ++i) {
      for (unsigned int j = 0; j < 6; ++j) {c->setCost((7 * r + 3 * i) % 40, (11 * r + 5 * j) % 40, 120);}   // a different map per robot
    }
    costmaps.push_back(ros);
  }
  nav2_controller::FleetControllerServer server;
  server.configure(node, "FollowPath", costmaps, nullptr);
  std::vector<geometry_msgs::msg::PoseStamped> poses(n_robots);
  std::vector<geometry_msgs::msg::Twist> speeds(n_robots);
  std::vector<double> yaws(n_robots);
  for (int r = 0; r < n_robots; ++r) {
    nav_msgs::msg::Path path;
    for (int i = 0; i < 60; ++i) {
      geometry_msgs::msg::PoseStamped p;
      p.pose.position.x = 0.4 + 0.05 * i;
      p.pose.position.y = 0.4 + 0.01 * (r + 1) * i;
      p.pose.orientation = tf2::quaternionFromYaw(std::atan2(0.01 * (r + 1), 0.05));
      path.poses.push_back(p);
    }
    server.setPlan(r, path);
    poses[r].pose.position.x = 0.5; poses[r].pose.position.y = 0.5 + 0.02 * r;
    yaws[r] = 0.1 * r;
  }
  TestGoalChecker goal_checker;
  for (int c = 0; c < cycles; ++c) {
    for (int r = 0; r < n_robots; ++r) {poses[r].pose.orientation = tf2::quaternionFromYaw(yaws[r]);}
    const auto cmds = server.computeVelocityCommands(poses, speeds, &goal_checker);
    for (int r = 0; r < n_robots; ++r) {
      std::printf("fleet %d %d %.9g %.9g %.9g %zu\n", r, cmds[r].valid ? 1 : 0, cmds[r].cmd.twist.linear.x,
        cmds[r].cmd.twist.linear.y, cmds[r].cmd.twist.angular.z, cmds[r].local_plan.poses.size());
      EXPECT(cmds[r].valid);
      speeds[r] = cmds[r].cmd.twist;
      poses[r].pose.position.x += speeds[r].linear.x * std::cos(yaws[r]) * 0.05;
      poses[r].pose.position.y += speeds[r].linear.x * std::sin(yaws[r]) * 0.05;
      yaws[r] += speeds[r].angular.z * 0.05;
    }
  }
  std::printf("fleet ok\n");
  return 0;
}

// CriticFunction::score on caller tensors through the mirror (critics_tests.cpp's shape of the call): every trajectory
// stands still at (x0, y0); prints the costs the critic added.
int testScore()
{
  auto node = makeNode("diff_drive");
  auto costmap_ros = makeCostmap(true);
  std::string name = "FollowPath";
  mppi::ParametersHandler ph(node, name);
  mppi::CriticManager cm;
  cm.on_configure(node, name, costmap_ros, &ph);
  const uint32_t B = 1000, T = 30;
  mppi::models::State state;
  state.batch_size = B; state.time_steps = T;
  state.vx.assign(B * T, 0.0f); state.vy.assign(B * T, 0.0f); state.wz.assign(B * T, 0.0f);
  mppi::models::Trajectories traj;
  traj.x.assign(B * T, 1.0f); traj.y.assign(B * T, 1.0f); traj.yaws.assign(B * T, 0.0f);
  for (uint32_t b = 0; b < B / 2; ++b) {                 // half of the batch sits on the cost-250 block
    for (uint32_t t = 0; t < T; ++t) {traj.x[t * B + b] = 2.0f; traj.y[t * B + b] = 2.0f;}
  }
  nav_msgs::msg::Path path;
  for (int i = 0; i < 30; ++i) {geometry_msgs::msg::PoseStamped p; p.pose.position.x = 0.2 * i; p.pose.position.y = 1.0; path.poses.push_back(p);}
  geometry_msgs::msg::Pose goal = path.poses.back().pose;
  std::vector<float> costs(B, 0.0f);
  float model_dt = 0.1f;
  state.local_path_length = 5.8f;
  mppi::critics::CriticData data{state, traj, path, goal, costs, model_dt};
  data.trajectories_in_collision.assign(B, false);
  cm.critics()[1]->score(data);                          // CostCritic
  EXPECT(!data.fail_flag);
  EXPECT(costs[B - 1] == 0.0f);                          // free space adds nothing
  EXPECT(costs[0] > 0.0f);
  std::printf("score cost %.9g %.9g\n", costs[0], costs[B - 1]);
  // cost_critic.cpp:223-228: every sample adds the cell cost 250 -> 250 * (3.81 / 254) for the stationary trajectory
  EXPECT(std::fabs(costs[0] - 250.0f * (3.81f / 254.0f)) < 1e-4f);
  const float before = costs[0];
  cm.critics()[7]->score(data);                          // PreferForwardCritic: vx == 0 everywhere adds nothing
  EXPECT(costs[0] == before);
  for (uint32_t i = 0; i < B * T; ++i) {state.vx[i] = -0.2f;}
  cm.critics()[7]->score(data);                          // backwards at 0.2 m/s: 5 * sum(0.2 * dt) = 5 * 0.2 * 0.1 * 30
  EXPECT(std::fabs((costs[0] - before) - 5.0f * 0.2f * 0.1f * 30.0f) < 1e-4f);
  std::printf("score ok\n");
  return 0;
}
}  // namespace

// nav2_costmap_2d/test/integration/inflation_tests.cpp:233-258, :330-381 through the C++ mirror; the grids are
// printed row by row ("grid <name> <sx> <sy>" then sy lines) so the Python test can compare them with the oracle
int testInflate()
{
  using nav2_costmap_2d::Costmap2D;
  using nav2_costmap_2d::InflationLayer;
  auto dump = [](const char * name, const Costmap2D & m) {
      std::printf("grid %s %u %u\n", name, m.getSizeInCellsX(), m.getSizeInCellsY());
      for (unsigned int y = 0; y < m.getSizeInCellsY(); ++y) {
        for (unsigned int x = 0; x < m.getSizeInCellsX(); ++x) {std::printf("%u ", m.getCost(x, y));}
        std::printf("\n");
      }
    };
  {
    InflationLayer::Parameters p;
    p.inflation_radius = 4.1; p.cost_scaling_factor = 1.0;
    InflationLayer layer(p);
    Costmap2D grid(10, 10, 1.0, 0.0, 0.0, 0);
    layer.matchSize(grid);
    layer.onFootprintChanged(2.1);
    grid.setCost(0, 0, nav2_costmap_2d::LETHAL_OBSTACLE);
    layer.updateCosts(grid, 0, 0, 10, 10);
    if (layer.getCellInflationRadius() != 5 || grid.getCost(0, 0) != 254 || grid.getCost(2, 0) != 253) {return 1;}
    dump("corner", grid);
  }
  {
    InflationLayer::Parameters p;
    p.inflation_radius = 10.5; p.cost_scaling_factor = 1.0;
    InflationLayer layer(p);
    Costmap2D grid(100, 100, 1.0, 0.0, 0.0, 0);
    layer.matchSize(grid);
    layer.onFootprintChanged(5.0);
    grid.setCost(50, 50, nav2_costmap_2d::LETHAL_OBSTACLE);
    layer.updateCosts(grid, 0, 0, 100, 100);
    for (unsigned int i = 6; i <= 11; ++i) {
      if (grid.getCost(50 + i, 50) != layer.computeCost(static_cast<double>(i))) {return 2;}  // :362-365
    }
    dump("centre", grid);
  }
  {
    InflationLayer::Parameters p;  // nav2 defaults on a bring-up sized map, window update
    p.inflation_radius = 0.70; p.cost_scaling_factor = 3.0;
    InflationLayer layer(p);
    Costmap2D grid(200, 120, 0.05, 0.0, 0.0, 0);
    layer.matchSize(grid);
    layer.onFootprintChanged(0.22);
    for (unsigned int k = 0; k < 40; ++k) {grid.setCost((k * 37) % 200, (k * 53) % 120, nav2_costmap_2d::LETHAL_OBSTACLE);}
    layer.updateCosts(grid, 20, 10, 180, 100);
    dump("window", grid);
  }
  std::printf("inflate ok\n");
  return 0;
}


// The reference's trajectory_visualizer_tests.cpp (StateTransition :28-39, VisOptimalTrajectory :41-103,
// VisCandidateTrajectories :105-137, VisOptimalPath :139-196) against the mirror; the shim's publishers keep the last
// message instead of delivering it to a subscription.  No GPU needed.
int testVisualizerUnit()
{
  using mppi::TrajectoryVisualizer;
  std::string name = "test";
  builtin_interfaces::msg::Time bogus_stamp;
  {  // StateTransition
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    mppi::ParametersHandler ph(node, name);
    TrajectoryVisualizer vis;
    vis.on_configure(node, "my_name", "map", &ph);
    vis.on_activate();
    vis.on_deactivate();
    vis.on_cleanup();
  }
  {  // VisOptimalTrajectory
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    mppi::ParametersHandler ph(node, name);
    TrajectoryVisualizer vis;
    vis.on_configure(node, "my_name", "fkmap", &ph);
    vis.on_activate();
    std::vector<std::array<float, 3>> optimal_trajectory;   // empty: nothing to draw
    vis.add(optimal_trajectory, "Optimal Trajectory", bogus_stamp);
    vis.visualize();
    EXPECT(vis.trajectoriesPublisher()->last() == nullptr || vis.trajectoriesPublisher()->last()->markers.empty());
    optimal_trajectory.assign(20, {1.0f, 1.0f, 1.0f});
    vis.add(optimal_trajectory, "Optimal Trajectory", bogus_stamp);
    vis.visualize();
    const auto * msg = vis.trajectoriesPublisher()->last();
    EXPECT(msg != nullptr && msg->markers.size() == 20u);   // 20 trajectory points in the given frame
    EXPECT(msg->markers[0].header.frame_id == "fkmap");
    EXPECT(msg->markers[0].id == 0 && msg->markers[1].id == 1 && msg->markers[10].id == 10);
    EXPECT(msg->markers[0].pose.position.x == 1 && msg->markers[0].pose.position.y == 1);
    EXPECT(msg->markers[0].pose.position.z == 0.06);
    EXPECT(msg->markers[0].scale.x == 0.03 && msg->markers[0].scale.y == 0.03 && msg->markers[0].scale.z == 0.07);
    EXPECT(msg->markers[19].scale.x == 0.07 && msg->markers[19].scale.y == 0.07 && msg->markers[19].scale.z == 0.09);
    for (size_t i = 0; i + 1 != msg->markers.size(); i++) {   // the colour ramp along the trajectory
      EXPECT(msg->markers[i].color.g < msg->markers[i + 1].color.g);
      EXPECT(msg->markers[i].color.b < msg->markers[i + 1].color.b);
      EXPECT(msg->markers[i].color.r == msg->markers[i + 1].color.r);
      EXPECT(msg->markers[i].color.a == msg->markers[i + 1].color.a);
    }
    // marker ids restart with every visualize() (trajectory_visualizer.cpp:204-211 reset)
    vis.add(optimal_trajectory, "Optimal Trajectory", bogus_stamp);
    vis.visualize();
    EXPECT(vis.trajectoriesPublisher()->last()->markers[0].id == 0);
  }
  {  // VisCandidateTrajectories: 200 trajectories of 12 points, default strides 5 / 3 -> 40 LINE_STRIP markers of 4 points
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    mppi::ParametersHandler ph(node, name);
    TrajectoryVisualizer vis;
    vis.on_configure(node, "my_name", "fkmap", &ph);
    vis.on_activate();
    EXPECT(vis.trajectoryStep() == 5 && vis.timeStep() == 3);
    const size_t B = 200, T = 12;
    const size_t n_traj = (B + 4) / 5, n_pts = (T + 2) / 3;
    std::vector<float> samples(n_traj * n_pts * 2, 1.0f), costs(B);
    for (size_t i = 0; i < B; ++i) {costs[i] = static_cast<float>(i) / static_cast<float>(B - 1);}   // LinSpaced(200, 0, 1)
    builtin_interfaces::msg::Time stamp;
    stamp.sec = 1;
    vis.add(samples, n_traj, n_pts, costs, {}, stamp);
    vis.visualize();
    const auto * msg = vis.trajectoriesPublisher()->last();
    EXPECT(msg != nullptr && msg->markers.size() == 40u);
    for (const auto & m : msg->markers) {
      EXPECT(m.type == visualization_msgs::msg::Marker::LINE_STRIP && m.points.size() == n_pts);
      EXPECT(m.header.stamp.sec == 1 && m.header.frame_id == "fkmap" && m.ns == "Candidate Trajectories");
    }
    // cost -> colour (trajectory_visualizer.cpp:189-202): the cheapest drawn trajectory is green, the costliest red
    EXPECT(msg->markers.front().color.g == 1.0f && msg->markers.front().color.r == 0.0f);
    EXPECT(msg->markers.back().color.r == 1.0f && msg->markers.back().color.g < 0.1f);
    // a colliding trajectory is drawn magenta and does not take part in the normalisation (:121-135, :171-176)
    std::vector<bool> collisions(B, false);
    collisions[B - 5] = true;   // the last drawn trajectory (row 195)
    vis.add(samples, n_traj, n_pts, costs, collisions, stamp);
    vis.visualize();
    msg = vis.trajectoriesPublisher()->last();
    EXPECT(msg->markers.size() == 40u);
    EXPECT(msg->markers[39].color.r == 1.0f && msg->markers[39].color.g == 0.0f && msg->markers[39].color.b == 1.0f);
    // nobody listening: nothing is assembled or published (:113-115)
    const size_t before = vis.trajectoriesPublisher()->publishedCount();
    vis.trajectoriesPublisher()->setSubscriptionCount(0);
    vis.add(samples, n_traj, n_pts, costs, {}, stamp);
    vis.visualize();
    EXPECT(vis.trajectoriesPublisher()->publishedCount() == before);
  }
  {  // VisOptimalPath
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    mppi::ParametersHandler ph(node, name);
    builtin_interfaces::msg::Time cmd_stamp;
    cmd_stamp.sec = 5;
    cmd_stamp.nanosec = 10;
    TrajectoryVisualizer vis;
    vis.on_configure(node, "my_name", "fkmap", &ph);
    vis.on_activate();
    std::vector<std::array<float, 3>> optimal_trajectory;
    vis.add(optimal_trajectory, "Optimal Trajectory", cmd_stamp);
    vis.visualize();
    EXPECT(vis.optimalPathPublisher()->last() == nullptr || vis.optimalPathPublisher()->last()->poses.empty());
    optimal_trajectory.assign(20, {0.0f, 0.0f, 0.0f});
    for (size_t i = 0; i + 1 != optimal_trajectory.size(); i++) {
      optimal_trajectory[i] = {static_cast<float>(i), static_cast<float>(i), static_cast<float>(i)};
    }
    vis.add(optimal_trajectory, "Optimal Trajectory", cmd_stamp);
    vis.visualize();
    const auto * path = vis.optimalPathPublisher()->last();
    EXPECT(path != nullptr && path->poses.size() == 20u);   // same frame and stamp as the velocity command
    EXPECT(path->header.frame_id == "fkmap" && path->header.stamp.sec == 5 && path->header.stamp.nanosec == 10u);
    for (size_t i = 0; i + 1 != path->poses.size(); i++) {
      EXPECT(path->poses[i].header.frame_id == "fkmap");
      EXPECT(path->poses[i].pose.position.x == static_cast<float>(i) && path->poses[i].pose.position.y == static_cast<float>(i));
      EXPECT(path->poses[i].pose.position.z == 0.06);
      const double yaw = static_cast<double>(optimal_trajectory[i][2]);   // tf2::Quaternion::setRPY(0, 0, yaw)
      EXPECT(path->poses[i].pose.orientation.x == 0.0 && path->poses[i].pose.orientation.y == 0.0);
      EXPECT(std::fabs(path->poses[i].pose.orientation.z - std::sin(yaw * 0.5)) < 1e-12);
      EXPECT(std::fabs(path->poses[i].pose.orientation.w - std::cos(yaw * 0.5)) < 1e-12);
    }
  }
  {  // utils_test.cpp:515-578 toTrajectoryMsgTest
    std::vector<std::array<float, 3>> trajectory;
    for (int i = 0; i < 5; ++i) {trajectory.push_back({static_cast<float>(i), static_cast<float>(i), static_cast<float>(i)});}
    mppi::models::ControlSequence control_sequence;
    control_sequence.vx.assign(5, 1.0f); control_sequence.wz.assign(5, 1.0f); control_sequence.vy.assign(5, 0.0f);
    std_msgs::msg::Header header;
    header.frame_id = "map";
    header.stamp.sec = 100;
    auto msg = mppi::utils::toTrajectoryMsg(trajectory, control_sequence, 1.0, header);
    EXPECT(msg->header.frame_id == "map" && msg->header.stamp.sec == 100 && msg->header.stamp.nanosec == 0u);
    EXPECT(msg->points.size() == 5u);
    for (int i = 0; i < 5; ++i) {
      EXPECT(msg->points[i].pose.position.x == static_cast<double>(i) && msg->points[i].pose.position.y == static_cast<double>(i));
      EXPECT(msg->points[i].velocity.linear.x == 1.0 && msg->points[i].velocity.linear.y == 0.0);
      EXPECT(msg->points[i].velocity.angular.z == 1.0);
      EXPECT(msg->points[i].header.stamp.sec == 100 + i && msg->points[i].header.stamp.nanosec == 0u);
      EXPECT(msg->points[i].header.frame_id == "map");
      EXPECT(std::fabs(tf2::getYaw(msg->points[i].pose.orientation) - std::remainder(static_cast<double>(i), 2 * M_PI)) < 1e-9);
    }
    // model_dt 0.05 from a stamp with nanoseconds: every point's stamp is the header's plus i * 50 ms
    header.stamp.nanosec = 980000000u;
    msg = mppi::utils::toTrajectoryMsg(trajectory, control_sequence, 0.05, header);
    EXPECT(msg->points[0].header.stamp.sec == 100 && msg->points[0].header.stamp.nanosec == 980000000u);
    EXPECT(msg->points[1].header.stamp.sec == 101 && msg->points[1].header.stamp.nanosec == 30000000u);
    // no vy in the control sequence (the reference's size() check): linear.y stays zero
    control_sequence.vy.clear();
    msg = mppi::utils::toTrajectoryMsg(trajectory, control_sequence, 0.05, header);
    EXPECT(msg->points[3].velocity.linear.y == 0.0);
  }
  std::printf("visunit ok\n");
  return 0;
}

// The reference's parameter_handler_test.cpp against the mirror: GetSystemParamsTest (:129-151), the static /
// dynamic split of DynamicAndStaticParametersTest (:153-212: a static parameter rejects the update with a reason, a
// dynamic one is applied and the post-callbacks run), PrePostDynamicCallbackTest's ordering (:77-127), and updates
// outside the plugin's namespace left alone (parameters_handler.cpp:66-72).  The mirror applies an accepted update to
// the node's parameter store and lets the post-callback re-read it (Optimizer::reset in the reference), instead of
// writing through a reference to the setting.  No GPU needed.
int testParametersHandler()
{
  using mppi::ParameterType;
  std::string name = "test";
  {  // GetSystemParamsTest
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    node->declare_parameter("param1", true);
    node->declare_parameter(name + ".param2", 7);
    mppi::ParametersHandler handler(node, name);
    auto getParameter = handler.getParamGetter("");
    bool p1 = false;
    int p2 = 0;
    getParameter(p1, "param1", false);
    getParameter(p2, name + ".param2", 0);
    EXPECT(p1 == true && p2 == 7);
    auto getParameter2 = handler.getParamGetter(name);
    p2 = 0;
    getParameter2(p2, "param2", 0);
    EXPECT(p2 == 7);
    // an undeclared parameter is declared with its default (nav2::declare_parameter_if_not_declared)
    double p3 = 0.0;
    getParameter2(p3, "param3", 1.5);
    EXPECT(p3 == 1.5 && node->has_parameter("test.param3"));
    // numeric settings convert like the reference's static_cast<SettingT>(param_in)
    float f = 0.0f;
    unsigned int u = 0;
    getParameter2(f, "param3", 0.0);
    getParameter2(u, "param2", 0);
    EXPECT(f == 1.5f && u == 7u);
  }
  {  // DynamicAndStaticParametersTest
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    node->declare_parameter("test.dynamic_int", 7);
    node->declare_parameter("test.static_int", 7);
    mppi::ParametersHandler handler(node, name);
    handler.start();
    auto getParameter = handler.getParamGetter("");
    int p1 = 0, p2 = 0;
    getParameter(p1, "test.dynamic_int", 0, ParameterType::Dynamic);
    getParameter(p2, "test.static_int", 0, ParameterType::Static);
    EXPECT(p1 == 7 && p2 == 7);
    int post_runs = 0;
    handler.addPostCallback([&]() {   // what Optimizer::reset does in the mirror: read the settings again
        ++post_runs;
        getParameter(p1, "test.dynamic_int", 0, ParameterType::Dynamic);
        getParameter(p2, "test.static_int", 0, ParameterType::Static);
      });
    auto result = handler.setDynamic("test.static_int", 10);
    EXPECT(!result.successful && !result.reason.empty());
    EXPECT(result.reason.rfind("Rejected change to static parameter: ", 0) == 0);
    EXPECT(result.reason.find("test.static_int") != std::string::npos);
    EXPECT(post_runs == 0 && p2 == 7);
    result = handler.setDynamic("test.dynamic_int", 10);
    EXPECT(result.successful && result.reason.empty());
    EXPECT(post_runs == 1 && p1 == 10 && p2 == 7);   // only the dynamic one changed
    // not this plugin's parameter: accepted, untouched, no callbacks
    result = handler.setDynamic("my_node.verbose", true);
    EXPECT(result.successful && post_runs == 1 && !node->has_parameter("my_node.verbose"));
  }
  {  // PrePostDynamicCallbackTest: validation runs before the update, the post-callbacks after it
    auto node = std::make_shared<nav2::LifecycleNode>("my_node");
    node->declare_parameter("test.flag", false);
    mppi::ParametersHandler handler(node, name);
    bool pre_triggered = false, post_triggered = false, order_ok = true;
    handler.addPreCallback(
      "test.flag", [&](const std::string &, const nav2::ParameterValue &, mppi::SetParametersResult &) {
        if (post_triggered) {order_ok = false;}
        pre_triggered = true;
      });
    handler.addPostCallback([&]() {if (!pre_triggered) {order_ok = false;} post_triggered = true;});
    auto result = handler.setDynamic("test.flag", true);
    EXPECT(result.successful && pre_triggered && post_triggered && order_ok);
    EXPECT(std::get<bool>(node->get_parameter("test.flag")) == true);
    // a pre-callback that rejects keeps the old value and skips the post-callbacks
    post_triggered = false;
    handler.addPreCallback(
      "test.flag", [&](const std::string &, const nav2::ParameterValue &, mppi::SetParametersResult & r) {
        r.successful = false;
        r.reason = "no";
      });
    result = handler.setDynamic("test.flag", false);
    EXPECT(!result.successful && result.reason == "no" && !post_triggered);
    EXPECT(std::get<bool>(node->get_parameter("test.flag")) == true);
    // the mutex the control loop shares with the updates (parameters_handler.hpp:100-104) is free again
    EXPECT(handler.getLock()->try_lock());
    handler.getLock()->unlock();
  }
  std::printf("params ok\n");
  return 0;
}

int main(int argc, char ** argv)
{
  try {
    if (argc >= 2 && std::strcmp(argv[1], "config") == 0) {return testConfig();}
    if (argc >= 2 && std::strcmp(argv[1], "inflate") == 0) {return testInflate();}
    if (argc >= 2 && std::strcmp(argv[1], "pathhandler") == 0) {return testPathHandler();}
    if (argc >= 2 && std::strcmp(argv[1], "visunit") == 0) {return testVisualizerUnit();}
    if (argc >= 2 && std::strcmp(argv[1], "params") == 0) {return testParametersHandler();}
    if (argc >= 3 && std::strcmp(argv[1], "visualize") == 0) {return testVisualize(std::atoi(argv[2]));}
    if (argc >= 4 && std::strcmp(argv[1], "fleet") == 0) {return testFleet(std::atoi(argv[2]), std::atoi(argv[3]));}
    if (argc >= 2 && std::strcmp(argv[1], "score") == 0) {return testScore();}
    if (argc >= 4 && std::strcmp(argv[1], "run") == 0) {return testRun(argv[2], std::atoi(argv[3]));}
  } catch (const std::exception & e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("usage: test_host config | run <diff_drive|omni|ackermann> <cycles>\n");
  return 64;
}
