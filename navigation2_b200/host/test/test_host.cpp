// test_host.cpp — drives the C++ host mirror the way controller_server drives the reference plugin
// (nav2_controller/src/controller_server.cpp:162-171 configure, :761-766 computeVelocityCommands):
// configure -> activate -> computeVelocityCommands xN -> deactivate -> cleanup, modelled on the
// reference's controller_state_transition_test.cpp and optimizer_smoke_test.cpp.
//   test_host config            parameter parsing / exceptions only (no GPU needed)
//   test_host run <model> <n>   n closed-loop cycles on the GPU; prints "cmd vx vy wz" per cycle
//   test_host inflate           the InflationLayer mirror on the GPU: the cases of inflation_tests.cpp, prints the grids
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

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
  EXPECT(std::fabs(tf2::getYaw(tf2::quaternionFromYaw(0.7)) - 0.7) < 1e-12);
  EXPECT(std::fabs(tf2::getYaw(tf2::quaternionFromYaw(-2.9)) + 2.9) < 1e-12);
  std::printf("config ok\n");
  return 0;
}

int testRun(const std::string & model, int cycles)
{
  auto node = makeNode(model);
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

int main(int argc, char ** argv)
{
  try {
    if (argc >= 2 && std::strcmp(argv[1], "config") == 0) {return testConfig();}
    if (argc >= 2 && std::strcmp(argv[1], "inflate") == 0) {return testInflate();}
    if (argc >= 4 && std::strcmp(argv[1], "run") == 0) {return testRun(argv[2], std::atoi(argv[3]));}
  } catch (const std::exception & e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("usage: test_host config | run <diff_drive|omni|ackermann> <cycles>\n");
  return 64;
}
