// mppi_device.h — POD structures shared by the host runtime (mppi_capi.cu) and the kernels
// (mppi_kernels.cu).  Everything here lives in device global memory or is passed by value as a
// kernel argument.  Names follow the reference's domain: trajectories (batch rows), time steps
// (columns), critics, costmap cells, path points.
#ifndef NAVIGATION2_B200_CSRC_MPPI_DEVICE_H_
#define NAVIGATION2_B200_CSRC_MPPI_DEVICE_H_

#include <stdint.h>

#include "mppi_b200.h"

#define MPPI_MAX_PATH_POINTS 2048
#define MPPI_BLOCK_A 256        // rollout+score kernel, throughput mode: threads per block (one trajectory each)
#define MPPI_BLOCK_B 128        // path-critics kernel: threads per block (eight resident blocks per SM)
#define MPPI_BLOCK_C 256        // softmax partial kernel
#define MPPI_TCHUNK 8           // time steps reduced together in the softmax partial kernel
#define MPPI_TCHUNK_A 8         // time steps per bulk-async stage of the rollout kernel
#define MPPI_MAX_STAGES 16      // ring slots (mbarriers) of the rollout kernel's noise pipeline
#define MPPI_MAX_PARTIAL_BLOCKS 296  // 2 x 148 SMs
#define MPPI_FUSED_G 128        // fused small-batch kernel: trajectories per CTA
#define MPPI_FUSED_THREADS 512  // 4 roles x MPPI_FUSED_G
#define MPPI_PHASE_CLOCK_WORDS (MPPI_FUSED_MAX_CLUSTER * (MPPI_FUSED_THREADS / 32) * 32)  // per robot
#define MPPI_FUSED_MAX_CLUSTER 16
#define MPPI_FINALIZE_MAX_SEGMENTS 6  // k_finalize: segments a column of block partials is summed in
#define MPPI_MAX_WINDOW_HALF 96   // cells either side of the robot staged in shared memory
#define MPPI_MAX_WINDOW_BYTES (((2 * MPPI_MAX_WINDOW_HALF + 1 + 15 + 15) & ~15) * (2 * MPPI_MAX_WINDOW_HALF + 1))

// slots of the per-robot reduction words (int32); the first MPPI_AR1_WORDS are what batch
// sharding all-reduces with MAX across ranks
enum
{
  RED_FURTHEST = 0,    // max over trajectories of the per-trajectory closest reachable path index
  RED_REP_MIN = 1,     // ObstaclesCritic repulsive_cost.minCoeff(), stored as ~ordered(float) so MAX == min
  RED_COST_FREE = 2,   // 1 if any trajectory does NOT collide in CostCritic (fail_flag = !this)
  RED_OBS_FREE = 3,    // same for ObstaclesCritic
  MPPI_AR1_WORDS = 4,
  RED_COST_MIN = 4,    // min total cost over the local rows, ~ordered(float)
  RED_MAP_MISS = 5,    // general kernels, window-only cycle: some lookup fell outside the uploaded costmap window
  RED_WORDS = 8
};

struct DevCritic
{
  int32_t type;
  int32_t enabled;
  uint32_t power;
  float weight;                 // CostCritic: cost_weight / 254 (cost_critic.cpp:39)
  int32_t consider_footprint;
  float critical_cost;
  float near_collision_cost;    // as float, compared like cost_critic.cpp:214
  float collision_cost;
  uint32_t step;                // trajectory_point_step
  float repulsion_weight;
  float critical_weight;
  float collision_margin_distance;
  int32_t symmetric_yaw_tolerance;
  float max_path_occupancy_ratio;
  uint32_t offset_from_furthest;
  int32_t use_path_orientations;
  float max_angle_to_furthest;
  int32_t mode;                 // PathAngle mode after the reversing_allowed_ fix-up
  float deadband[3];            // fabs()'d
};

// Static per configuration; uploaded by mppi_create / mppi_set_config.
struct alignas(16) DevStatic
{
  DevCritic critics[MPPI_MAX_CRITICS];
  int32_t n_critics;
  int32_t critic_of_type[MPPI_CRITIC_TYPE_COUNT];  // index into critics[] or -1
  int32_t n_footprint;
  int32_t track_unknown;
  double footprint_x[MPPI_MAX_FOOTPRINT];
  double footprint_y[MPPI_MAX_FOOTPRINT];
  // ObstaclesCritic::distanceToObstacle (obstacles_critic.cpp:105-118) for every uint8 cost,
  // [0] = point cost (radius subtracted), [1] = footprint cost; computed on the host with libm log
  float obstacle_dist_lut[2][256];
  float obstacle_inflation_radius;   // inflation_radius_ (0 => no repulsion, :165-167)
  float obstacle_scale_factor;
  float param_vx_max, param_vx_min, param_vy_max;  // ConstraintCritic reads the parent parameters
  float min_turning_r;
  int32_t motion_model;
  // MotionModel::predict deltas (motion_models.hpp:111-115)
  float max_delta_vx, min_delta_vx, max_delta_vy, min_delta_vy, max_delta_wz;
  int32_t offset_vx, offset_vy, offset_wz;  // delay offsets (:160-162)
  int32_t clamp_raw_controls;
  float model_dt;
  float controller_period;
  int32_t shift_control_sequence;
  float gamma_vx, gamma_wz, gamma_vy;       // gamma / std^2 (optimizer.cpp:612,618,625)
  int32_t use_gamma_wz;                     // sampling_std.wz > 0
  float inv_temp;                           // 1 / temperature
  uint32_t sgf_order;
  float ax_max, ax_min, ay_max, ay_min, az_max;
  int32_t validator_enabled;
  int32_t validator_consider_footprint;
  uint32_t validator_samples;               // traj_samples_to_evaluate_
};

// Per robot, per evalControl call; uploaded in one H2D copy.  The last fields are updated by the
// kernels between iterations of the same call.
struct alignas(16) DevCycle
{
  double pose_x_d, pose_y_d, pose_yaw_d;
  double origin_x_d, origin_y_d, resolution_d;
  float pose_x, pose_y, pose_yaw;
  float init_cos, init_sin;
  float speed_vx, speed_vy, speed_wz;       // state.speed narrowed to float (column 0 of v*)
  float local_path_length;
  float origin_x_f, origin_y_f, resolution_f;
  uint32_t size_x, size_y, pitch;
  uint32_t win_packet_off;                  // byte offset of this robot's rows in KArgs::win_packets (window-only cycles)
  int32_t win_x0, win_y0, win_w, win_h;     // costmap window staged in shared memory
  int32_t n_path;
  uint32_t active_mask;                     // bit k: critics[k] passes its enabled / threshold gates
  uint32_t near_goal_mask;                  // bit k: local_path_length < near_goal_distance (Cost, Obstacles)
  float possible_collision_cost[MPPI_MAX_CRITICS];
  float goal_x, goal_y, goal_yaw, sym_goal_yaw;  // last path pose (goal_critic.cpp:41, goal_angle_critic.cpp:42-47)
  float c_vx_max, c_vx_min, c_vy, c_wz;     // settings_.constraints (speed limit applied)
  float hist_vx[MPPI_MAX_DELAY_STEPS], hist_vy[MPPI_MAX_DELAY_STEPS], hist_wz[MPPI_MAX_DELAY_STEPS];
  int32_t track_collisions;
  int32_t need_furthest;                    // some active critic calls setPathFurthestPointIfNotSet
  int32_t need_path_valid;
  int32_t valid_on_device;                  // the costmap exists only in device memory (mppi_set_costmap_device): the kernels
                                            // evaluate utils::findPathCosts themselves instead of reading path_valid
  float w2m_eps;                            // how far (in cells) the float estimate of the double world -> map quotient can be off
                                            // for this robot's map (>= 1: the estimate is never trusted); see obstacle_cell()
  // ---- mutable within the call ----
  int32_t furthest_cached;                  // CriticData::furthest_reached_path_point, -1 = unset
  int32_t fail_flag;                        // CriticData::fail_flag (sticky until prepare())
  int32_t run;                              // 0: this robot is idle in this attempt (fleet retries, score_tensors)
  uint32_t seq;                             // stamped into DevOutHeader::seq by the last iteration: the host polls for it
};

// Written by the finalize kernel, copied D2H once per call.  Followed in memory by
// float u[3][T] (vx, vy, wz) and float traj[3][T] (x, y, yaw).
struct DevOutHeader
{
  float cmd_vx, cmd_vy, cmd_wz;
  int32_t fail_flag;
  int32_t validation;
  int32_t furthest;
  float cost_min;
  float softmax_sum;
  int32_t map_miss;   // a lookup fell outside the uploaded costmap window while the full map was not resident:
                      // nothing was committed (control sequence, filter history); the host uploads the map and reruns
  uint32_t seq;       // DevCycle::seq of the cycle these results belong to
  int32_t critics_evaluated;  // how many entries of the critic list CriticManager::evalTrajectoriesScores walked before
                              // fail_flag stopped it (critic_manager.cpp:89-93): n_critics if it never did
  int32_t pad[1];
};

struct KArgs
{
  int32_t B;            // local trajectories per robot
  int32_t T;
  int32_t R;
  int32_t b_offset;     // global index of local row 0 (batch sharding)
  int32_t n_terms;      // n_critics + 3 gamma slots
  int32_t n_pa;         // PathAlign strided samples per trajectory
  int32_t pa_step;
  int32_t pa_stride;    // floats per trajectory row of pa_samples: 3 * n_pa rounded up to a multiple of 4
  int32_t max_path;     // stride of the path arrays
  int32_t world;        // shard world size
  int32_t rank;
  int32_t n_partial_blocks;
  int32_t materialize;
  int32_t last_iteration;
  int32_t do_shift;
  int32_t zero_costs;   // first iteration of an evalControl call: costs_ start from zero (optimizer.cpp:335) without a memset
  // noise [R][T][B]
  const float * noise_vx;
  const float * noise_vy;
  const float * noise_wz;
  // clamped controls [R][T][B], only when clamp_raw_controls (read back by the softmax kernel)
  float * cvx; float * cvy; float * cwz;
  // state / trajectories [R][T][B]: outputs when materialize, inputs for score_tensors
  float * vx; float * vy; float * wz;
  float * x; float * y; float * yaw;
  float * u;            // [R][3][T] control sequence
  float * ctrl_hist;    // [R][4][3]
  float * terms;        // [R][n_terms][B]
  float * obs_raw;      // [R][B]
  float * obs_rep;      // [R][B]
  float * last_xyz;     // [R][3][B]
  float * pa_samples;   // [R][B][pa_stride]: per trajectory x0 y0 x1 y1 ... then yaw0 yaw1 ... (a thread's row is contiguous:
                        // kernel A stores with immediate offsets, kernel B fetches it with 16-byte loads)
  float * costs;        // [R][B] running total (accumulates over iterations)
  float * softmax;      // [R][B]
  float * critic_costs; // [R][n_critics][B] or null (visualize)
  uint8_t * collisions; // [R][B]
  uint32_t * dbg_cost_cells;      // [R][n_s][B] or null
  uint32_t * dbg_obstacle_cells;  // [R][T][B] or null
  int32_t * red;        // [R][RED_WORDS]
  float * partials;     // [R][n_partial_blocks][1 + 3T]
  float * ar2_slot;     // [R][2 + 3T] this rank's (min, sum, W)
  const float * ar2_all;// [world][R][2 + 3T]
  // Batch sharding with both exchanges folded into the kernels (peer memory over NVLink, one robot): the last block of
  // kernel A pushes the reduction words into every rank's mailbox, every block of kernel B waits for its peers' words and
  // reduces them, the finalize kernel pushes this rank's softmax slot, waits and combines.  xch_peers == nullptr: no
  // folded exchange (single GPU, or the caller runs collectives between the mppi_shard_* phases).
  // Mailbox layout, identical on every rank: uint32 seq[2 kinds][2 parities][world], then per kind a payload area
  // [2 parities][world][stride bytes].  Exchange number n of a kind uses parity n & 1 and publishes n + 1.
  unsigned char * const * xch_peers;   // [world] mailbox base pointers (peers' opened with CUDA IPC)
  uint32_t * xch_state;                // device words: [0] / [1] exchanges of kind 0 / 1 completed, [2] kernel A blocks done
  int32_t * xch_error;                 // mapped host word, set when a peer did not show up within the timeout
  uint32_t xch_words_base, xch_words_stride, xch_slots_base, xch_slots_stride;  // bytes
  const uint8_t * costmap;  // [R][map_stride]
  size_t map_stride;
  // Single-launch cycles upload only the window of the costmap the trajectories can plausibly reach, packed
  // row by row (win_w x win_h bytes per robot, next to the cycle block: one copy).  map_resident = 0 then:
  // `costmap` is stale, a lookup outside the window raises DevOutHeader::map_miss instead of reading it.
  // TrajectoryVisualizer samples (MPPI_TENSOR_VIS_XY): [R][vis_n_traj][vis_n_pts][2], or null
  float * vis_xy;
  int32_t vis_traj_step, vis_time_step, vis_n_traj, vis_n_pts;
  const uint8_t * win_packets;  // robot r's rows start at DevCycle::win_packet_off
  int32_t map_resident;
  int32_t stamp_results;   // unused since results carry their own tags (kept: part of the plan key)
  const float * path;   // [R][4][max_path]: x, y, yaw, integrated distance
  const uint8_t * path_valid;  // [R][max_path]
  DevCycle * cyc;       // [R]
  // Where an iteration READS its inputs from; the live buffers above are where it WRITES.  They are the same
  // buffers except on the first iteration of a resident replay (pristine copies: no restore copies needed) and
  // of a single-launch cycle (the pinned host block, read over PCIe: no H2D copy node ahead of the kernel).
  const DevCycle * cyc_in;
  const float * u_in;
  const float * hist_in;
  const DevStatic * st;
  long long * phase_clocks;  // [R][MPPI_FUSED_MAX_CLUSTER][warps][32] SM clock at each phase boundary of the fused kernel, or null
  // Result blocks in mapped pinned host memory, [R][out_words] 8-byte words: word k holds the 32-bit word k of the plain
  // layout (DevOutHeader, then the control sequence [3][T], then the optimal trajectory [3][T]) in .x and the
  // sequence number of the cycle that wrote it (DevCycle::seq) in .y.  The host takes a result once every word carries
  // the number it waits for, so the kernels need no fence and no ordering between their stores.
  uint2 * out_tag;
  size_t out_words;
  int32_t debug_flags;  // MPPI_B200_DEBUG_FLAGS (measurements only): bit 0 = the fused kernel rehearses its finalize once
  int32_t pad_debug;
};

#endif  // NAVIGATION2_B200_CSRC_MPPI_DEVICE_H_
