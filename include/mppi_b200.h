/*
 * mppi_b200.h — C ABI of the B200-native MPPI trajectory optimiser.
 *
 * This is the drop-in boundary for the hot path of nav2_mppi_controller
 * (mppi::Optimizer::evalControl and everything below it).  The reference has no
 * C ABI; each entry point below names the reference C++ interface it replaces (file:line under
 * /root/reference/nav2_mppi_controller unless another package is named).  A C++ shim with the
 * reference's class names (navigation2_b200/host/) and the Python binding
 * (navigation2_b200/capi.py) both sit on top of exactly these symbols.
 *
 * Conventions
 *  - plain C, POD only: no ROS, Eigen, torch or C++ types cross this boundary;
 *  - every function returns an MppiStatus (0 = OK, negative = error); no exceptions.  The C++
 *    shim maps MPPI_ERR_NO_VALID_CONTROL -> nav2_core::NoValidControl and
 *    MPPI_ERR_CONFIG -> nav2_core::ControllerException (nav2_core/controller_exceptions.hpp:23-77);
 *  - all B x T tensors are float32, column-major with the batch index contiguous:
 *    element (b, t) lives at t * B + b — the layout of the reference's Eigen::ArrayXXf
 *    (models/state.hpp:31-57, models/trajectories.hpp:27-41);
 *  - one handle = one optimiser instance = n_robots independent robots sharing one MppiConfig
 *    (n_robots = 1 is the reference's single controller; > 1 is fleet mode);
 *  - a handle is not thread-safe; the reference serialises on the ParametersHandler mutex
 *    (src/controller.cpp:104) and so does the shim.
 *  - there is NO CPU fallback: every entry point that computes fails with MPPI_ERR_CUDA when no
 *    sm_100-class device is usable.
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_ABI_VERSION 1
#define MPPI_MAX_CRITICS 16
#define MPPI_MAX_FOOTPRINT 32
#define MPPI_MAX_DELAY_STEPS 16

typedef enum MppiStatus
{
  MPPI_OK = 0,
  MPPI_ERR_INVALID_ARGUMENT = -1,
  MPPI_ERR_CUDA = -2,
  /* nav2_core::NoValidControl: retry limit exceeded (optimizer.cpp:292-295) or validator FAILURE (:251-253) */
  MPPI_ERR_NO_VALID_CONTROL = -3,
  /* nav2_core::ControllerException: configuration errors (optimizer.cpp:178,683; cost_critic.cpp:67;
   * obstacles_critic.cpp:56; path_angle_critic.cpp:95) */
  MPPI_ERR_CONFIG = -4,
  MPPI_ERR_NOT_READY = -5, /* optimise called before a costmap was uploaded */
  MPPI_ERR_UNSUPPORTED = -6
} MppiStatus;

/* motion_models.xml: mppi::DiffDriveMotionModel / OmniMotionModel / AckermannMotionModel */
typedef enum MppiMotionModel
{
  MPPI_MODEL_DIFF_DRIVE = 0,
  MPPI_MODEL_OMNI = 1,
  MPPI_MODEL_ACKERMANN = 2
} MppiMotionModel;

/* critics.xml: mppi::critics::<Name>Critic, in the file's order */
typedef enum MppiCriticType
{
  MPPI_CRITIC_CONSTRAINT = 0,
  MPPI_CRITIC_COST = 1,
  MPPI_CRITIC_GOAL = 2,
  MPPI_CRITIC_GOAL_ANGLE = 3,
  MPPI_CRITIC_OBSTACLES = 4,
  MPPI_CRITIC_PATH_ALIGN = 5,
  MPPI_CRITIC_PATH_ANGLE = 6,
  MPPI_CRITIC_PATH_FOLLOW = 7,
  MPPI_CRITIC_PREFER_FORWARD = 8,
  MPPI_CRITIC_TWIRLING = 9,
  MPPI_CRITIC_VELOCITY_DEADBAND = 10,
  MPPI_CRITIC_TYPE_COUNT = 11
} MppiCriticType;

/*
 * One configured critic.  A flat superset of every critic's ROS parameters;
 * fields a critic does not declare are ignored for it.  Names are the ROS parameter names.
 * mppi_default_critic_params() fills the code defaults of src/critics/<name>_critic.cpp::initialize().
 */
typedef struct MppiCriticParams
{
  int32_t type;                    /* MppiCriticType */
  int32_t enabled;                 /* critic_function.hpp:80-81 "<name>.enabled" */
  uint32_t cost_power;             /* all critics */
  float cost_weight;               /* all but Obstacles; CostCritic: the ROS value (3.81), divided by 254 internally (cost_critic.cpp:39) */
  float threshold_to_consider;     /* Goal, GoalAngle, PathAlign, PathAngle, PathFollow, PreferForward */
  int32_t consider_footprint;      /* Cost, Obstacles */
  float critical_cost;             /* Cost */
  uint32_t near_collision_cost;    /* Cost */
  float collision_cost;            /* Cost (1e6), Obstacles (1e5) */
  float near_goal_distance;        /* Cost, Obstacles */
  uint32_t trajectory_point_step;  /* Cost (2), PathAlign (4) */
  float repulsion_weight;          /* Obstacles */
  float critical_weight;           /* Obstacles */
  float collision_margin_distance; /* Obstacles */
  int32_t symmetric_yaw_tolerance; /* GoalAngle */
  float max_path_occupancy_ratio;  /* PathAlign */
  uint32_t offset_from_furthest;   /* PathAlign (20), PathAngle (4), PathFollow (6) */
  int32_t use_path_orientations;   /* PathAlign */
  float max_angle_to_furthest;     /* PathAngle */
  int32_t mode;                    /* PathAngle: 0 forward pref., 1 no directional pref., 2 feasible path orientations */
  float deadband_velocities[3];    /* VelocityDeadband: vx, vy, wz */
} MppiCriticParams;

/*
 * Everything Optimizer::getParams (src/optimizer.cpp:77-161), MotionModel::initialize,
 * NoiseGenerator::initialize (src/noise_generator.cpp:36), the validator
 * (optimal_trajectory_validator.hpp:87-103) and the critics read at configure time, plus the
 * quantities the critics derive from Costmap2DROS / LayeredCostmap / InflationLayer
 * (cost_critic.cpp:85-129, obstacles_critic.cpp:70-118; SURVEY.md §8b "Derived").
 */
typedef struct MppiConfig
{
  uint32_t abi_version;            /* MPPI_ABI_VERSION */
  uint32_t n_robots;               /* 1 = reference behaviour; >1 = fleet of independent robots */

  /* --- optimizer.cpp:105-129 (ROS parameter names) --- */
  uint32_t batch_size;
  uint32_t time_steps;
  uint32_t iteration_count;
  float model_dt;
  float model_delay_vx;
  float model_delay_vy;
  float model_delay_wz;
  int32_t clamp_raw_controls;
  float temperature;
  float gamma;
  float vx_max, vx_min, vy_max, wz_max;
  float ax_max, ax_min, ay_max, ay_min, az_max;
  float vx_std, vy_std, wz_std;
  uint32_t retry_attempt_limit;
  int32_t open_loop;
  uint32_t sgf_order;
  double controller_frequency;     /* parent node parameter, optimizer.cpp:157-160 */
  int32_t motion_model;            /* MppiMotionModel, from "motion_model" */
  float min_turning_r;             /* "<name>.ackermann.min_turning_r" (motion_models.hpp:276) */

  /* --- noise_generator.cpp:36 --- */
  int32_t regenerate_noises;
  uint64_t noise_seed;             /* Philox4x32-10 key; the reference's minstd_rand0 stream is unpinned */

  /* --- critic_manager.cpp:40-41 "critics" (order is evaluation order), "visualize" --- */
  uint32_t n_critics;
  MppiCriticParams critics[MPPI_MAX_CRITICS];
  int32_t visualize;               /* keep per-critic cost arrays + collision flags (critic_manager.cpp:79-120) */
  int32_t materialize_tensors;     /* write state.v*, state.cv*, trajectories.x/y/yaws to HBM every iteration so
                                      mppi_get_tensor can return what the plugin API exposes (CriticData refs) */
  /* TrajectoryVisualizer "trajectory_step" / "time_step" (trajectory_visualizer.cpp:38-39, defaults 5 / 3): with
   * `visualize` set, every trajectory_step-th candidate trajectory is kept at every time_step-th point
   * (MPPI_TENSOR_VIS_XY) — exactly what TrajectoryVisualizer::add reads (:144,177) — without materialising the
   * B x T tensors.  0 = keep no samples. */
  uint32_t vis_trajectory_step;
  uint32_t vis_time_step;

  /* --- TrajectoryValidator (optimal_trajectory_validator.hpp:87-103) --- */
  int32_t validator_enabled;       /* 0 = no validation (always SUCCESS) */
  float collision_lookahead_time;
  int32_t validator_consider_footprint;

  /* --- derived from Costmap2DROS / LayeredCostmap / InflationLayer at configure --- */
  int32_t use_radius;              /* Costmap2DROS::getUseRadius(): robot_radius instead of a footprint polygon */
  uint32_t n_footprint;            /* padded footprint polygon, Costmap2DROS::getRobotFootprint() */
  double footprint_x[MPPI_MAX_FOOTPRINT];
  double footprint_y[MPPI_MAX_FOOTPRINT];
  double inscribed_radius;         /* LayeredCostmap::getInscribedRadius() */
  double circumscribed_radius;     /* LayeredCostmap::getCircumscribedRadius() */
  int32_t has_inflation_layer;
  double inflation_cost_scaling_factor; /* InflationLayer::getCostScalingFactor() */
  double inflation_radius;         /* InflationLayer::getInflationRadius() */
  int32_t track_unknown;           /* LayeredCostmap::isTrackingUnknown() */

  /* --- batch sharding across GPUs; world 1 = no sharding --- */
  uint32_t shard_rank;
  uint32_t shard_world;            /* batch_size is the GLOBAL batch; this rank owns batch_size / shard_world rows */

  /* --- debug --- */
  int32_t debug_cells;             /* record costmap cell indices of every CostCritic / ObstaclesCritic lookup */
} MppiConfig;

/*
 * Per-robot inputs of one Optimizer::evalControl call (src/optimizer.cpp:230-236):
 * robot_pose, robot_speed, plan, goal, goal_checker.  Quaternions are reduced to yaw by the
 * caller with tf2::getYaw (the shim does it); everything stays double, as in the ROS messages.
 */
typedef struct MppiCycleIn
{
  double pose_x, pose_y, pose_yaw;          /* robot_pose.pose */
  double speed_vx, speed_vy, speed_wz;      /* robot_speed (measured twist) */
  const double * path_x;                    /* plan.poses[i].pose.position.x, n_path entries */
  const double * path_y;
  const double * path_yaw;                  /* tf2::getYaw(plan.poses[i].pose.orientation) */
  uint32_t n_path;                          /* may be 0 (SURVEY.md §9 hazard): path/goal critics are skipped */
  double goal_x, goal_y, goal_yaw;          /* `goal` argument (kept for CriticData::goal parity) */
  int32_t has_goal_checker;                 /* goal_checker != nullptr */
  double goal_checker_xy_tolerance;         /* pose_tolerance.position.x from GoalChecker::getTolerances */
  /* models::State::local_path_length is normally calculate_path_length(plan) (optimizer.cpp:333).  The
   * reference's critic tests write the field directly; set has_local_path_length to do the same. */
  int32_t has_local_path_length;
  double local_path_length;
} MppiCycleIn;

typedef enum MppiValidation
{
  MPPI_VALIDATION_SUCCESS = 0,
  MPPI_VALIDATION_SOFT_RESET = 1,
  MPPI_VALIDATION_FAILURE = 2
} MppiValidation;

/*
 * Per-robot outputs of one evalControl call: the TwistStamped command
 * (optimizer.cpp:647-664), the optimal trajectory (getOptimizedTrajectory :582-598) and the
 * control sequence (getOptimalControlSequence :600-603).  Array pointers may be NULL.
 */
typedef struct MppiCycleOut
{
  int32_t status;                   /* MppiStatus for this robot */
  float cmd_vx, cmd_vy, cmd_wz;     /* control_sequence(offset), offset = shift ? 1 : 0 */
  int32_t fail_flag;                /* CriticData::fail_flag after the last optimize() */
  int32_t validation;               /* MppiValidation of the last attempt */
  uint32_t retries;                 /* fallback() attempts consumed by this call */
  int32_t furthest_reached_path_point; /* -1 if no critic asked for it */
  float * control_vx;               /* [time_steps] sequence as it stood when the command was read (before shift) */
  float * control_vy;
  float * control_wz;
  float * optimal_trajectory;       /* [3 * time_steps]: x[T], y[T], yaw[T] (Eigen ArrayXXf(T,3), column-major) */
} MppiCycleOut;

/* mppi_get_tensor selectors.  B x T float unless noted. */
typedef enum MppiTensor
{
  MPPI_TENSOR_VX = 0, MPPI_TENSOR_VY = 1, MPPI_TENSOR_WZ = 2,          /* models::State::vx/vy/wz */
  MPPI_TENSOR_CVX = 3, MPPI_TENSOR_CVY = 4, MPPI_TENSOR_CWZ = 5,       /* models::State::cvx/cvy/cwz */
  MPPI_TENSOR_X = 6, MPPI_TENSOR_Y = 7, MPPI_TENSOR_YAW = 8,           /* models::Trajectories */
  MPPI_TENSOR_NOISE_VX = 9, MPPI_TENSOR_NOISE_VY = 10, MPPI_TENSOR_NOISE_WZ = 11, /* NoiseGenerator::noises_* */
  MPPI_TENSOR_COSTS = 12,           /* float[B]: costs_ after the last iteration (critics + gamma terms) */
  MPPI_TENSOR_CRITIC_COSTS = 13,    /* float[n_critics * B]: cost added by each critic (CriticManager::critic_costs_) */
  MPPI_TENSOR_COLLISIONS = 14,      /* uint8[B]: CriticData::trajectories_in_collision */
  MPPI_TENSOR_COST_CELLS = 15,      /* uint32[n_s * B]: CostCritic cell index per strided sample, 0xFFFFFFFF off-map, 0xFFFFFFFE not visited (debug_cells) */
  MPPI_TENSOR_OBSTACLE_CELLS = 16,  /* uint32[T * B]: ObstaclesCritic cell index per step, same sentinels (debug_cells) */
  MPPI_TENSOR_SOFTMAX = 17,         /* float[B]: unnormalised softmax weights exp(-(c - min)/temperature) */
  MPPI_TENSOR_VIS_XY = 19,          /* float[ceil(B / vis_trajectory_step)][ceil(T / vis_time_step)][2]: x, y of the candidate
                                       trajectories the visualizer draws (visualize + vis_trajectory_step > 0) */
  MPPI_TENSOR_PHASE_CLOCKS = 18     /* int64[16 * 16 * 32]: SM clock at the phase boundaries of the fused kernel, per cluster rank and warp
                                       (tracing; enabled by the environment variable MPPI_B200_PHASE_CLOCKS=1) */
} MppiTensor;

/*
 * Everything mppi::Optimizer carries from one evalControl call to the next, besides the noise:
 * control_sequence_, control_history_ (optimizer.hpp:309-310), last_command_vel_ (:331), the
 * motion model's command-history ring buffers (motion_models.hpp:246-248) and fallback()'s retry
 * counter (optimizer.cpp:283).  The reference has no checkpoint/resume; this is
 * the hand-over used to migrate a robot between handles / GPUs and to start two implementations
 * from bit-identical state in the parity tests.
 */
typedef struct MppiOptimizerState
{
  double last_command_vel[3];                    /* vx, vy, wz */
  float control_history[12];                     /* 4 x (vx, vy, wz), oldest first */
  float cmd_history_vx[MPPI_MAX_DELAY_STEPS];    /* newest last; only offsetSteps(delay) entries are used */
  float cmd_history_vy[MPPI_MAX_DELAY_STEPS];
  float cmd_history_wz[MPPI_MAX_DELAY_STEPS];
  uint32_t fallback_counter;
} MppiOptimizerState;

typedef struct MppiHandle MppiHandle;

/* ---- configuration helpers (no GPU needed) ---- */

/* Code defaults of Optimizer::getParams (optimizer.cpp:105-150), NoiseGenerator (noise_generator.cpp:36)
 * and the validator (optimal_trajectory_validator.hpp:87,98); no critics, 1 robot, no sharding. */
int mppi_default_config(MppiConfig * cfg);
/* Code defaults of src/critics/<type>_critic.cpp::initialize(). */
int mppi_default_critic_params(int32_t critic_type, MppiCriticParams * out);
/* "ConstraintCritic" ... -> MppiCriticType, -1 if unknown (critic_manager.cpp:70-73 getFullName). */
int mppi_critic_type_from_name(const char * name);
const char * mppi_critic_name(int32_t critic_type);
const char * mppi_status_string(int status);
size_t mppi_sizeof_config(void);
size_t mppi_sizeof_critic_params(void);
size_t mppi_sizeof_cycle_in(void);
size_t mppi_sizeof_cycle_out(void);
size_t mppi_sizeof_optimizer_state(void);
size_t mppi_sizeof_costmap_view(void);
/* InflationLayer::computeCost (nav2_costmap_2d/include/nav2_costmap_2d/inflation_layer.hpp:121-135) */
uint8_t mppi_inflation_compute_cost(
  double distance_cells, double resolution, double inscribed_radius, double cost_scaling_factor);

/* ---- lifecycle ---- */

/* Optimizer::initialize (optimizer.cpp:34-70) + MPPIController::configure (controller.cpp:26-57):
 * validates the configuration (MPPI_ERR_CONFIG where the reference throws ControllerException),
 * allocates device tensors, streams and pinned staging, draws the first noise, reset(). */
int mppi_create(const MppiConfig * cfg, MppiHandle ** out);
/* Optimizer::shutdown (optimizer.cpp:72-75) + MPPIController::cleanup (controller.cpp:59-67). */
int mppi_destroy(MppiHandle * h);
/* Dynamic parameter update: ParametersHandler post-callback -> Optimizer::reset() (optimizer.cpp:155);
 * batch_size / time_steps may change, device buffers are re-allocated. */
int mppi_set_config(MppiHandle * h, const MppiConfig * cfg);
/* Optimizer::reset(bool) (optimizer.cpp:183-211) for one robot, or all robots if robot == UINT32_MAX. */
int mppi_reset(MppiHandle * h, uint32_t robot, int reset_dynamic_speed_limits);
/* Optimizer::setSpeedLimit (optimizer.cpp:691-719); speed_limit == 0.0 is nav2_costmap_2d::NO_SPEED_LIMIT. */
int mppi_set_speed_limit(MppiHandle * h, double speed_limit, int percentage);
/* Message of the last error on this handle (or of the last failed mppi_create if h is NULL). */
const char * mppi_last_error(const MppiHandle * h);

/* ---- per-cycle data ---- */

/* Replaces the critics' direct reads of Costmap2D::getCharMap()/getOriginX()/getResolution()/
 * getSizeInCells*() (cost_critic.cpp:138-144; nav2_costmap_2d/src/costmap_2d.cpp:265-305).
 * `cells` is row-major uint8, idx = my * size_x + mx (costmap_2d.hpp:231-234).  The bytes are
 * copied to pinned staging (the caller's buffer is free again on return) and sent by pinned
 * cudaMemcpyAsync: on a side stream right away when the handle runs the general kernels, by the
 * next mppi_optimize on its own stream — window or whole map, see mppi_optimize_in_place — when
 * its cycles are a single launch.  Call while holding the costmap mutex, as the reference does
 * (controller.cpp:106-107). */
int mppi_upload_costmap(
  MppiHandle * h, uint32_t robot, const uint8_t * cells, uint32_t size_x, uint32_t size_y,
  double resolution, double origin_x, double origin_y);

/* The same hand-over for a costmap that already lives in DEVICE memory — the output of the inflation pass
 * (costmap_inflation_update_costs_device, include/costmap_inflation_b200.h; reference producer:
 * nav2_costmap_2d/plugins/inflation_layer.cpp:283-348, consumer: cost_critic.cpp:138-144).  `device_cells` is
 * row-major uint8 with `row_pitch` bytes between rows; it is copied device-to-device into the handle's own storage on
 * the handle's stream, ordered after everything enqueued so far on `producer_stream` (a cudaStream_t; NULL = the
 * caller has synchronised).  Nothing crosses PCIe and the host never sees the cells: path validity
 * (utils::findPathCosts) is then evaluated on the device.  Replaced by the next mppi_upload_costmap /
 * mppi_optimize_in_place for that robot. */
int mppi_set_costmap_device(
  MppiHandle * h, uint32_t robot, const uint8_t * device_cells, uint32_t size_x, uint32_t size_y, size_t row_pitch,
  double resolution, double origin_x, double origin_y, void * producer_stream);

/* Reference-noise injection: replaces NoiseGenerator::noises_vx_/vy_/wz_ (noise_generator.hpp:96-98).
 * Each pointer is B_local x T column-major host memory; nvy may be NULL (zeros, as for
 * non-holonomic models).  Injected noise survives reset() until mppi_generate_noise is called. */
int mppi_set_noise(
  MppiHandle * h, uint32_t robot, const float * noises_vx, const float * noises_vy,
  const float * noises_wz);
/* NoiseGenerator::generateNoisedControls (noise_generator.cpp:108-119) with Philox4x32-10 + Box-Muller,
 * keyed by (seed, robot, epoch) and the GLOBAL trajectory index so results do not depend on sharding. */
int mppi_generate_noise(MppiHandle * h, uint32_t robot);

/* Optimizer::evalControl (optimizer.cpp:230-270) for every robot of the handle: prepare,
 * do { optimize; validate } while (fallback), read the command, shift the control sequence.
 * `in` and `out` are arrays of n_robots.  Host buffers in, host buffers out; H2D/D2H inside.
 * Returns MPPI_OK if the call ran; per-robot results are in out[i].status. */
int mppi_optimize(MppiHandle * h, const MppiCycleIn * in, MppiCycleOut * out);

/* One robot's costmap where it lies in the caller's memory: what a critic sees through
 * Costmap2D::getCharMap() / getSizeInCellsX() / getResolution() / getOriginX() while it scores
 * (cost_critic.cpp:138-144).  Row-major uint8, idx = my * size_x + mx. */
typedef struct MppiCostmapView
{
  const uint8_t * cells;
  uint32_t size_x, size_y;
  double resolution, origin_x, origin_y;
} MppiCostmapView;

/* mppi_upload_costmap for every robot + mppi_optimize in one call, the way the reference does it:
 * evalControl reads the costmap in place, under the costmap mutex the caller holds
 * (controller.cpp:106-107).  `costmaps` is an array of n_robots views that must stay valid and unchanged
 * until the call returns.  Knowing the robot pose at the moment it sees the costmap lets the library
 * send only the cells the trajectories can plausibly reach (a ~10 KB window instead of a 160 KB map
 * at the reference configuration); the kernel detects a lookup outside that window, commits nothing,
 * and the cycle is repeated with the whole map (mppi_window_miss_count).  Results are identical to
 * the two-call sequence. */
int mppi_optimize_in_place(
  MppiHandle * h, const MppiCostmapView * costmaps, const MppiCycleIn * in, MppiCycleOut * out);
/* Cycles that had to be repeated with the whole costmap since the handle was created. */
uint64_t mppi_window_miss_count(const MppiHandle * h);
/* Bytes of costmaps, cycle inputs and costmap windows this handle has copied host -> device since it was
 * created (bench.py derives "h2d_bytes_per_step" from it). */
uint64_t mppi_h2d_byte_count(const MppiHandle * h);

/* The scoring pass alone, on caller-supplied tensors: CriticManager::evalTrajectoriesScores
 * (critic_manager.cpp:76-121) with CriticData{state, trajectories, path, goal, costs, ...} given
 * explicitly — the interface the reference's critics_tests.cpp exercises.  Host pointers,
 * B x T column-major; vy may be NULL.  furthest_reached_path_point < 0 means "not set"
 * (critic_data.hpp:53).  costs is in/out (critics add to it); collisions / fail_flag may be NULL. */
int mppi_score_tensors(
  MppiHandle * h, uint32_t robot, const MppiCycleIn * in,
  const float * vx, const float * vy, const float * wz,
  const float * x, const float * y, const float * yaw,
  int32_t furthest_reached_path_point,
  float * costs, uint8_t * collisions, int32_t * fail_flag, int32_t * furthest_out);

/* Copies one tensor of the last optimize()/score call to host memory (dst_bytes must match). */
int mppi_get_tensor(MppiHandle * h, uint32_t robot, int32_t which, void * dst, size_t dst_bytes);
/* How many entries of the critic list the last optimize() iteration evaluated before fail_flag stopped the loop
 * (CriticManager::evalTrajectoriesScores, critic_manager.cpp:89-93): the critics CriticsStats reports (:101-117). */
int mppi_get_critics_evaluated(MppiHandle * h, uint32_t robot, uint32_t * n_evaluated);
/* Current control sequence (after the last evalControl's shift), 3 x T floats: vx, vy, wz. */
int mppi_get_control_sequence(MppiHandle * h, uint32_t robot, float * vx, float * vy, float * wz);
int mppi_set_control_sequence(
  MppiHandle * h, uint32_t robot, const float * vx, const float * vy, const float * wz);
/* Save / restore the carried state (control sequence excluded: use the two calls above). */
int mppi_get_state(MppiHandle * h, uint32_t robot, MppiOptimizerState * state);
int mppi_set_state(MppiHandle * h, uint32_t robot, const MppiOptimizerState * state);

/* ---- resident / measurement path (bench.py `value`, ncu) ---- */

/* Re-enqueue the optimize() iterations of the last mppi_optimize inputs on `cuda_stream`
 * (a cudaStream_t, NULL = the handle's stream) with everything already resident in HBM:
 * no H2D, no D2H, no host sync.  The control sequence is restored first so every call does
 * identical work. */
int mppi_enqueue_resident(MppiHandle * h, void * cuda_stream);
/* `steps` such replays on the handle's own stream, each bracketed by CUDA events, after `warmup` untimed
 * ones; ms_per_step receives `steps` device times.  flush_bytes > 0 flushes L2 before every replay by
 * overwriting a buffer of that size (bench.py "value": larger than the 126 MB L2); 0 = hot L2. */
int mppi_time_resident(MppiHandle * h, int steps, int warmup, size_t flush_bytes, float * ms_per_step);
/* The scoring pass alone (kernel A in score_tensors mode + kernel B) on the state / trajectory tensors the
 * last mppi_optimize materialised (materialize_tensors = 1), everything resident: the kernel the
 * reference's critics correspond to, for roofline measurement. */
int mppi_enqueue_score_resident(MppiHandle * h, void * cuda_stream);
/* mppi_enqueue_resident replays the inputs of the last mppi_optimize; keeping the pristine copies it
 * needs costs three small device copies per cycle, so it is opt-in. */
int mppi_set_resident_capture(MppiHandle * h, int enabled);
/* `cycles` back-to-back control cycles, each timed with steady_clock on the calling thread: the
 * end-to-end latency of this C ABI with host buffers (H2D costmap + cycle inputs, kernels, D2H
 * result).  in_place = 0: (mppi_upload_costmap for every robot; mppi_optimize) pairs; in_place != 0:
 * mppi_optimize_in_place with the same costmap for every robot.  per_call_seconds receives `cycles`
 * entries. */
int mppi_time_e2e(
  MppiHandle * h, const uint8_t * cells, uint32_t size_x, uint32_t size_y, double resolution, double origin_x,
  double origin_y, const MppiCycleIn * in, MppiCycleOut * out, int cycles, int in_place, double * per_call_seconds);
/* The same measurement with one costmap view per robot (a fleet whose robots each own a map): `cycles`
 * mppi_optimize_in_place calls, each timed with steady_clock on the calling thread. */
int mppi_time_e2e_views(
  MppiHandle * h, const MppiCostmapView * costmaps, const MppiCycleIn * in, MppiCycleOut * out, int cycles,
  double * per_call_seconds);
/* Kernels launched by this handle since creation (bench.py "gpu_launches"). */
uint64_t mppi_launch_count(const MppiHandle * h);
/* CUDA-event time of the dominant (rollout+score) kernel, averaged since the last call with
 * reset != 0; enabled by mppi_set_kernel_timing(h, 1).  Milliseconds per launch. */
int mppi_set_kernel_timing(MppiHandle * h, int enabled);
int mppi_kernel_time_ms(MppiHandle * h, int reset, double * ms_per_launch, uint64_t * launches);
/* Same for kernel C (softmax weights + weighted control sums), the HBM-bound kernel of large batches. */
int mppi_kernel_c_time_ms(MppiHandle * h, int reset, double * ms_per_launch, uint64_t * launches);
/* The stream the handle launches on (cudaStream_t), for event timing by the caller. */
void * mppi_stream(MppiHandle * h);
/* Launch on the caller's stream instead (e.g. the stream torch.distributed orders its NCCL
 * collectives against); NULL restores the handle's own stream. */
int mppi_set_stream(MppiHandle * h, void * cuda_stream);
int mppi_device_synchronize(MppiHandle * h);

/* Device self-test (no reference equivalent).  The fused kernel divides by the costmap resolution
 * and takes square roots with straight-line sequences that must equal the IEEE operators the
 * reference uses (CostCritic::worldToMapFloat, cost_critic.hpp:100-113; utils.hpp:307-312).  Runs
 * both on n host operand pairs (a[i] / b[i], sqrt(|a[i]|)) and counts results that differ from `/`
 * and sqrtf (must be 0) and operands the sequences decline (handled by the operators themselves).
 * Errors are reported through mppi_last_error(NULL). */
int mppi_selftest_exact_math(
  const float * a, const float * b, int n, unsigned int * mismatches, unsigned int * flagged);

/* ---- batch sharding across GPUs: one process per GPU, collectives by the caller ---- */

/* Exchange buffers in device memory.  ar1: int32[ar1_words * n_robots], combine with MAX across ranks
 * (furthest point, negated-ordered repulsive min, any-trajectory-free flags).  ar2: this rank's slot of
 * float[ar2_words * n_robots]: (cost min, softmax sum, weighted control sums [3T]); all-gather the slots
 * into `ar2_all` (world x ar2_words x n_robots), which the finalize kernel reads. */
int mppi_shard_buffers(
  MppiHandle * h, void ** ar1, size_t * ar1_words, void ** ar2_slot, size_t * ar2_words,
  void ** ar2_all);
int mppi_shard_begin(MppiHandle * h, const MppiCycleIn * in);   /* prepare + H2D */
int mppi_shard_rollout(MppiHandle * h);                         /* kernel A: rollout + per-trajectory critics; then all-reduce ar1 (MAX) */
int mppi_shard_score(MppiHandle * h);                           /* kernels B, C: path critics, local softmax partials; then all-gather ar2 */
int mppi_shard_update(MppiHandle * h, int last_iteration);      /* kernel D: combine slots, filter, constraints */
int mppi_shard_end(MppiHandle * h, MppiCycleOut * out);         /* D2H + host state machine */
/* The same cycle with both exchanges done on the device over peer memory (NVLink) instead of by a host-launched
 * collective: every rank owns a mailbox its peers write into directly.  Setup, once per handle: each rank exports its
 * mailbox (mppi_shard_ipc_handle, a 64-byte CUDA IPC handle), the caller all-gathers the handles (torch.distributed)
 * and gives every rank all of them, rank-major (mppi_shard_open_peers).  Then mppi_shard_cycle, called by all ranks
 * with the same inputs, is a whole evalControl: the kernels and the exchange kernels of an iteration are enqueued
 * back to back, nothing goes through the host in between.  A peer that does not show up within 2 s fails the cycle
 * (MPPI_ERR_CUDA) instead of hanging it. */
int mppi_shard_ipc_handle(MppiHandle * h, unsigned char * handle64);
int mppi_shard_open_peers(MppiHandle * h, const unsigned char * handles /* [shard_world][64] */);
int mppi_shard_cycle(MppiHandle * h, const MppiCycleIn * in, MppiCycleOut * out);
/* `cycles` back-to-back mppi_shard_cycle calls (every rank calls it with the same arguments), each timed with
 * steady_clock on the calling thread: the end-to-end latency of a sharded cycle free of binding overhead. */
int mppi_time_shard_cycles(MppiHandle * h, const MppiCycleIn * in, MppiCycleOut * out, int cycles, double * per_call_seconds);

#ifdef __cplusplus
}
#endif

#endif  /* MPPI_B200_H_ */
