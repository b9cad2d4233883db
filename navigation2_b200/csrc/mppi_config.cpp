// mppi_config.cpp — host-only helpers of the C ABI: parameter defaults, names, status strings.
// Defaults are the reference's *code* defaults (not the bring-up yaml), each cited below; paths
// relative to /root/reference/nav2_mppi_controller/.
#include <cmath>
#include <cstring>

#include "mppi_b200.h"

extern "C" {

int mppi_default_config(MppiConfig * cfg)
{
  if (!cfg) {return MPPI_ERR_INVALID_ARGUMENT;}
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = MPPI_ABI_VERSION;
  cfg->n_robots = 1;
  // src/optimizer.cpp:105-129
  cfg->model_dt = 0.05f;
  cfg->model_delay_vx = 0.0f;
  cfg->model_delay_vy = 0.0f;
  cfg->model_delay_wz = 0.0f;
  cfg->clamp_raw_controls = 0;
  cfg->time_steps = 56;
  cfg->batch_size = 1000;
  cfg->iteration_count = 1;
  cfg->temperature = 0.3f;
  cfg->gamma = 0.015f;
  cfg->vx_max = 0.5f;
  cfg->vx_min = -0.35f;
  cfg->vy_max = 0.5f;
  cfg->wz_max = 1.9f;
  cfg->ax_max = 3.0f;
  cfg->ax_min = -3.0f;
  cfg->ay_max = 3.0f;
  cfg->ay_min = -3.0f;
  cfg->az_max = 3.5f;
  cfg->vx_std = 0.2f;
  cfg->vy_std = 0.2f;
  cfg->wz_std = 0.4f;
  cfg->retry_attempt_limit = 1;
  cfg->open_loop = 0;
  cfg->sgf_order = 2;
  cfg->controller_frequency = 20.0;  // nav2_controller default; == 1/model_dt so shifting is ON
  cfg->motion_model = MPPI_MODEL_DIFF_DRIVE;  // optimizer.cpp:150
  cfg->min_turning_r = 0.2f;                  // motion_models.hpp:276
  cfg->regenerate_noises = 0;                 // src/noise_generator.cpp:36
  cfg->noise_seed = 0x9E3779B97F4A7C15ull;
  cfg->n_critics = 0;                         // src/critic_manager.cpp:40 (empty list)
  cfg->visualize = 0;                         // src/controller.cpp:40
  cfg->materialize_tensors = 0;
  cfg->vis_trajectory_step = 5;               // src/trajectory_visualizer.cpp:38
  cfg->vis_time_step = 3;                     // src/trajectory_visualizer.cpp:39
  cfg->validator_enabled = 1;                 // optimizer.cpp:56-58 default validator plugin
  cfg->collision_lookahead_time = 2.0f;       // optimal_trajectory_validator.hpp:87
  cfg->validator_consider_footprint = 0;      // :98
  cfg->use_radius = 1;
  cfg->n_footprint = 0;
  cfg->inscribed_radius = 0.0;
  cfg->circumscribed_radius = 0.0;
  cfg->has_inflation_layer = 0;
  cfg->track_unknown = 0;
  cfg->shard_rank = 0;
  cfg->shard_world = 1;
  cfg->debug_cells = 0;
  return MPPI_OK;
}

int mppi_default_critic_params(int32_t critic_type, MppiCriticParams * p)
{
  if (!p || critic_type < 0 || critic_type >= MPPI_CRITIC_TYPE_COUNT) {
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  std::memset(p, 0, sizeof(*p));
  p->type = critic_type;
  p->enabled = 1;     // include/nav2_mppi_controller/critic_function.hpp:80-81
  p->cost_power = 1;
  switch (critic_type) {
    case MPPI_CRITIC_CONSTRAINT:  // src/critics/constraint_critic.cpp:21-35
      p->cost_weight = 4.0f;
      break;
    case MPPI_CRITIC_COST:  // src/critics/cost_critic.cpp:24-36
      p->consider_footprint = 0;
      p->cost_weight = 3.81f;
      p->critical_cost = 300.0f;
      p->near_collision_cost = 253;
      p->collision_cost = 1000000.0f;
      p->near_goal_distance = 0.5f;
      p->trajectory_point_step = 2;
      break;
    case MPPI_CRITIC_GOAL:  // src/critics/goal_critic.cpp:22-28
      p->cost_weight = 5.0f;
      p->threshold_to_consider = 1.4f;
      break;
    case MPPI_CRITIC_GOAL_ANGLE:  // src/critics/goal_angle_critic.cpp:20-27
      p->cost_weight = 3.0f;
      p->threshold_to_consider = 0.5f;
      p->symmetric_yaw_tolerance = 0;
      break;
    case MPPI_CRITIC_OBSTACLES:  // src/critics/obstacles_critic.cpp:24-35
      p->consider_footprint = 0;
      p->repulsion_weight = 1.5f;
      p->critical_weight = 20.0f;
      p->collision_cost = 100000.0f;
      p->collision_margin_distance = 0.10f;
      p->near_goal_distance = 0.5f;
      break;
    case MPPI_CRITIC_PATH_ALIGN:  // src/critics/path_align_critic.cpp:20-32
      p->cost_weight = 10.0f;
      p->max_path_occupancy_ratio = 0.07f;
      p->offset_from_furthest = 20;
      p->trajectory_point_step = 4;
      p->threshold_to_consider = 0.5f;
      p->use_path_orientations = 0;
      break;
    case MPPI_CRITIC_PATH_ANGLE:  // src/critics/path_angle_critic.cpp:23-54
      p->offset_from_furthest = 4;
      p->cost_weight = 2.2f;
      p->threshold_to_consider = 0.5f;
      p->max_angle_to_furthest = 0.785398f;
      p->mode = 0;
      break;
    case MPPI_CRITIC_PATH_FOLLOW:  // src/critics/path_follow_critic.cpp:22-32
      p->threshold_to_consider = 1.4f;
      p->offset_from_furthest = 6;
      p->cost_weight = 5.0f;
      break;
    case MPPI_CRITIC_PREFER_FORWARD:  // src/critics/prefer_forward_critic.cpp:22-30
      p->cost_weight = 5.0f;
      p->threshold_to_consider = 0.5f;
      break;
    case MPPI_CRITIC_TWIRLING:  // src/critics/twirling_critic.cpp:22-27
      p->cost_weight = 10.0f;
      break;
    case MPPI_CRITIC_VELOCITY_DEADBAND:  // src/critics/velocity_deadband_critic.cpp:20-32
      p->cost_weight = 35.0f;
      p->deadband_velocities[0] = 0.0f;
      p->deadband_velocities[1] = 0.0f;
      p->deadband_velocities[2] = 0.0f;
      break;
    default:
      return MPPI_ERR_INVALID_ARGUMENT;
  }
  return MPPI_OK;
}

// critics.xml class names without the "mppi::critics::" prefix (critic_manager.cpp:70-73)
static const char * const kCriticNames[MPPI_CRITIC_TYPE_COUNT] = {
  "ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic",
  "PathAlignCritic", "PathAngleCritic", "PathFollowCritic", "PreferForwardCritic",
  "TwirlingCritic", "VelocityDeadbandCritic"};

int mppi_critic_type_from_name(const char * name)
{
  if (!name) {return -1;}
  const char * prefix = "mppi::critics::";
  if (std::strncmp(name, prefix, std::strlen(prefix)) == 0) {name += std::strlen(prefix);}
  for (int i = 0; i < MPPI_CRITIC_TYPE_COUNT; ++i) {
    if (std::strcmp(name, kCriticNames[i]) == 0) {return i;}
  }
  return -1;
}

const char * mppi_critic_name(int32_t critic_type)
{
  if (critic_type < 0 || critic_type >= MPPI_CRITIC_TYPE_COUNT) {return "";}
  return kCriticNames[critic_type];
}

const char * mppi_status_string(int status)
{
  switch (status) {
    case MPPI_OK: return "ok";
    case MPPI_ERR_INVALID_ARGUMENT: return "invalid argument";
    case MPPI_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
    case MPPI_ERR_NO_VALID_CONTROL: return "no valid control (nav2_core::NoValidControl)";
    case MPPI_ERR_CONFIG: return "configuration error (nav2_core::ControllerException)";
    case MPPI_ERR_NOT_READY: return "not ready";
    case MPPI_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

size_t mppi_sizeof_config(void) {return sizeof(MppiConfig);}
size_t mppi_sizeof_critic_params(void) {return sizeof(MppiCriticParams);}
size_t mppi_sizeof_cycle_in(void) {return sizeof(MppiCycleIn);}
size_t mppi_sizeof_cycle_out(void) {return sizeof(MppiCycleOut);}
size_t mppi_sizeof_optimizer_state(void) {return sizeof(MppiOptimizerState);}
size_t mppi_sizeof_costmap_view(void) {return sizeof(MppiCostmapView);}

// nav2_costmap_2d/include/nav2_costmap_2d/inflation_layer.hpp:121-135
uint8_t mppi_inflation_compute_cost(
  double distance_cells, double resolution, double inscribed_radius, double cost_scaling_factor)
{
  unsigned char cost = 0;
  if (distance_cells == 0) {
    cost = 254;
  } else if (distance_cells * resolution <= inscribed_radius) {
    cost = 253;
  } else {
    const double factor =
      exp(-1.0 * cost_scaling_factor * (distance_cells * resolution - inscribed_radius));
    cost = static_cast<unsigned char>((253 - 1) * factor);
  }
  return cost;
}

}  // extern "C"
