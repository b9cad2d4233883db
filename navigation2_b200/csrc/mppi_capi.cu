// mppi_capi.cu — host runtime behind the C ABI of include/mppi_b200.h.
//
// Owns the device tensors, the streams and the pinned staging of one optimiser instance and runs
// the host-side state machine of mppi::Optimizer (prepare / fallback / reset / speed limit,
// src/optimizer.cpp) around the kernels of mppi_kernels.cu.  Everything numeric per trajectory
// happens on the GPU; the host only prepares O(path length) scalars per cycle.
// There is no CPU fallback: without a usable device every compute entry point fails.
//
// Reference citations are relative to /root/reference/nav2_mppi_controller/.

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mppi_b200.h"
#include "mppi_device.h"
#include "mppi_kernels.h"
#include "mppi_math.h"

namespace
{

thread_local std::string g_create_error;

// fleets of at least this many robots share the per-robot host work of a cycle (Optimizer::prepare, window packing,
// result unpacking) out over the host cores with OpenMP
// Host threads for the per-robot preparation of a fleet cycle (window packing, prepare_robot): MPPI_B200_HOST_THREADS,
// else the machine's hardware threads, at most 16.  Given explicitly to every parallel region: launchers such as
// torchrun export OMP_NUM_THREADS=1, which would silently serialise a fleet's host side.
static int mppi_host_threads()
{
  static const int n = [] {
      const char * e = std::getenv("MPPI_B200_HOST_THREADS");
      int v = e ? std::atoi(e) : 0;
      if (v <= 0) {v = static_cast<int>(std::thread::hardware_concurrency());}
      return std::max(1, std::min(v, 16));
    }();
  return n;
}

#ifndef MPPI_HOST_PARALLEL_ROBOTS
#define MPPI_HOST_PARALLEL_ROBOTS 32
#endif

struct HostRobot
{
  // Optimizer::last_command_vel_ (doubles of the float command)
  double last_cmd_vx = 0, last_cmd_vy = 0, last_cmd_wz = 0;
  // MotionModel::cmd_history_* ring buffers (motion_models.hpp:246-248), newest last
  std::vector<float> hist_vx, hist_vy, hist_wz;
  size_t fallback_counter = 0;  // optimizer.cpp:283 (function-static there; per robot here)
  // costmap geometry of the last upload
  bool map_valid = false;
  uint32_t size_x = 0, size_y = 0, pitch = 0;
  double resolution = 0, origin_x = 0, origin_y = 0;
  const uint8_t * map_host = nullptr;  // the robot's slab of the pinned staging buffer (16-byte-pitched rows); path
                                       // validity is evaluated on the host from it, O(n_path) lookups
  bool noise_injected = false;
  uint32_t noise_epoch = 0;
  cudaEvent_t stage_event = nullptr;   // completion of a slab copy this robot led (upload_stale_maps)
  int stage_owner = 0;                 // the robot whose stage_event covers the last copy out of THIS robot's slab
  bool stage_pending = false;
  bool device_stale = false;     // the host-side map is newer than the device copy
  size_t map_bytes = 0;          // pitch * size_y
  // where the host reads this robot's cells from (path validity, window packets): the staging slab after
  // mppi_upload_costmap, the caller's own buffer during mppi_optimize_in_place
  const uint8_t * map_src = nullptr;
  uint32_t map_src_pitch = 0;
  bool staged_current = false;   // the staging slab holds the costmap of the current cycle
  bool device_only = false;      // the costmap was handed over in device memory (mppi_set_costmap_device): no host copy exists
};

#define CU_CHECK(h, call) \
  do { \
    cudaError_t e__ = (call); \
    if (e__ != cudaSuccess) { \
      (h)->error = std::string(#call) + ": " + cudaGetErrorString(e__); \
      return MPPI_ERR_CUDA; \
    } \
  } while (0)

}  // namespace

struct MppiHandle
{
  MppiConfig cfg{};
  std::string error;
  int R = 1, B = 0, Bglobal = 0, T = 0;
  bool holo = false;
  bool shift = false;
  float controller_period = 0.0f;
  // settings_.base_constraints / constraints (speed limit), optimizer.cpp:115-123,152
  float base_vx_max = 0, base_vx_min = 0, base_vy = 0, base_wz = 0;
  float c_vx_max = 0, c_vx_min = 0, c_vy = 0, c_wz = 0;
  float ax_max = 0, ax_min = 0, ay_max = 0, ay_min = 0, az_max = 0;
  int off_vx = 0, off_vy = 0, off_wz = 0;
  uint32_t validator_samples = 0;
  int n_pa = 0, pa_step = 1, cc_ns = 0;
  int critic_of_type[MPPI_CRITIC_TYPE_COUNT];
  int path_angle_mode = 0;
  std::vector<HostRobot> robots;
  DevStatic h_static{};

  cudaStream_t stream = nullptr, own_stream = nullptr, side = nullptr;
  cudaEvent_t map_event = nullptr;
  bool map_event_pending = false;

  // device buffers
  float * d_noise[3] = {nullptr, nullptr, nullptr};   // vx, vy, wz
  // regenerate_noises: the tensors of the NEXT cycle are drawn on the side stream while the host is between cycles
  // (the reference's noise thread, noise_generator.cpp:98-106); the two sets swap roles every cycle
  float * d_noise_next[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t noise_event = nullptr;
  bool noise_prefetched = false;
  bool noise_prefetch = true;      // MPPI_B200_NO_NOISE_PREFETCH=1: draw at the start of the cycle on its own stream
  uint32_t noise_flip = 0;
  float * d_cv[3] = {nullptr, nullptr, nullptr};
  float * d_state[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // vx vy wz x y yaw
  float * d_u = nullptr, * d_u_saved = nullptr;
  float * d_hist = nullptr, * d_hist_saved = nullptr;
  float * d_terms = nullptr, * d_obs_raw = nullptr, * d_obs_rep = nullptr, * d_last = nullptr;
  float * d_pa = nullptr, * d_costs = nullptr, * d_softmax = nullptr, * d_critic_costs = nullptr;
  uint8_t * d_collisions = nullptr;
  uint32_t * d_dbg_cost = nullptr, * d_dbg_obs = nullptr;
  int32_t * d_red = nullptr;
  float * d_vis_xy = nullptr;     // TrajectoryVisualizer samples [R][vis_n_traj][vis_n_pts][2]
  int vis_n_traj = 0, vis_n_pts = 0;
  long long * d_phase_clocks = nullptr;
  bool defer_map_upload = false;  // the last cycle was single-launch: mppi_optimize moves map + cycle block in one copy
  bool single_copy = true;        // MPPI_B200_NO_SINGLE_COPY=1: side-stream map upload + copy node, as on the general path
  float * d_partials = nullptr, * d_slot = nullptr, * d_slot_all = nullptr;
  uint8_t * d_map = nullptr;
  size_t map_stride = 0;
  DevStatic * d_static = nullptr;
  uint8_t * d_cycle_block = nullptr, * d_cycle_saved = nullptr;   // [DevCycle R][path R*4*max_path f32][valid R*max_path]
  uint2 * d_out_tag = nullptr;    // device alias of h_out_tag
  size_t out_stride = 0;
  int n_partial_blocks = 1;
  int max_path = 0;
  // pinned host staging
  // One pinned and one device arena with the same layout: [R costmap slabs of map_stride bytes][cycle block].
  // h_map_stage / d_map and h_cycle_block / d_cycle_block point into them, so a single-launch cycle moves the
  // freshly written costmap and the cycle block host -> device with ONE copy ahead of ONE kernel.
  uint8_t * h_arena = nullptr, * d_arena = nullptr;
  // ... followed by the window packets of a single-launch cycle (tightly packed, DevCycle::win_packet_off)
  uint8_t * h_packets = nullptr, * d_packets = nullptr;
  size_t packets_capacity = 0;
  size_t packets_used = 0;
  int window_half_override = 0;   // MPPI_B200_WINDOW_HALF=<cells>: shrink the staged window (tests force window misses)
  // batch sharding over peer memory (k_shard_exchange): this rank's mailbox, the peers' mailboxes opened with CUDA IPC
  unsigned char * d_mailbox = nullptr;
  size_t mailbox_bytes = 0, mailbox_words_base = 0, mailbox_slots_base = 0, mailbox_words_stride = 0, mailbox_slots_stride = 0;
  std::vector<void *> peer_mailboxes;     // [world]; own entry = d_mailbox
  unsigned char ** d_peers = nullptr;     // the same pointers on the device
  int32_t * d_xerror = nullptr, * h_xerror = nullptr;
  uint32_t * d_xstate = nullptr;          // device words: exchanges completed per kind [2] + kernel A's block ticket (KArgs::xch_state)
  bool peers_open = false;
  int shard_iteration = 0;                // iterations launched since mppi_shard_begin
  bool shard_failed = false;              // an exchange timed out: the ranks are out of step for good, the handle is dead
  uint8_t * d_flush = nullptr;    // mppi_time_resident: the buffer overwritten to flush L2
  size_t flush_bytes = 0;
  std::vector<cudaEvent_t> resident_events;
  uint32_t cycle_seq = 0;         // attempts launched; stamped into the result header by the kernels
  bool poll_results = true;       // MPPI_B200_NO_POLL=1: wait for the stream instead of polling the result stamp
  uint64_t h2d_bytes = 0;         // bytes of costmaps, cycle blocks and window packets copied host -> device so far
  uint64_t window_misses = 0;     // window-only attempts that had to be repeated with the whole map
  bool zero_copy = false;         // MPPI_B200_ZERO_COPY=1: single-launch window-only cycles read the cycle block and the window rows
                                  // straight from the pinned host arena (no cudaMemcpyAsync ahead of the kernel)
  bool window_only = true;        // MPPI_B200_NO_WINDOW_ONLY=1: single-launch cycles upload the whole map (A/B measurements)
  bool window_only_general = true;  // MPPI_B200_NO_WINDOW_GENERAL=1: the general kernels always get whole maps
  uint8_t * h_cycle_block = nullptr;
  size_t cycle_block_bytes = 0;
  uint8_t * h_out = nullptr;      // plain result blocks [R][out_stride], filled by unpack_results from h_out_tag
  uint2 * h_out_tag = nullptr;    // mapped pinned: what the kernels write ([R][out_stride / 4] tagged words, KArgs::out_tag)
  uint8_t * h_map_stage = nullptr;

  bool host_timers = false;    // MPPI_B200_HOST_TIMERS=1: accumulate host-side phase times, printed by mppi_destroy
  double ht[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t ht_n = 0;
  bool allow_fused = true;     // MPPI_B200_NO_FUSED=1 forces the general four-kernel path
  uint64_t fused_checked_key = ~0ull;
  bool fused_checked_ok = false;
  bool have_cycle = false;     // a cycle block is resident (mppi_enqueue_resident allowed)
  struct LaunchPlanBox * resident_plan = nullptr;  // the launch plan of the captured cycle (the host cycle block's run flags are
                                                   // cleared when a cycle finishes: a plan made afterwards would see no active critic)
  bool resident_capture = false;  // keep pristine copies of the cycle inputs for mppi_enqueue_resident
  bool use_graphs = true;      // MPPI_B200_NO_GRAPH=1 launches node by node
  struct CycleGraph {uint64_t key; cudaGraphExec_t exec;};
  std::vector<CycleGraph> graphs;  // one instantiated graph per launch-plan signature
  bool from_tensors_loaded = false;
  uint64_t launches = 0;
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;    // around kernel A (or the fused kernel)
  size_t timing_used = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events_c;  // around kernel C (softmax partials)
  size_t timing_used_c = 0;

  DevCycle * hcyc(int r) {return reinterpret_cast<DevCycle *>(h_cycle_block) + r;}
  float * hpath(int r) {
    return reinterpret_cast<float *>(h_cycle_block + sizeof(DevCycle) * R) + static_cast<size_t>(r) * 4 * max_path;
  }
  uint8_t * hvalid(int r) {
    return h_cycle_block + sizeof(DevCycle) * R + sizeof(float) * 4 * max_path * R + static_cast<size_t>(r) * max_path;
  }
};

namespace
{

void drop_graphs(MppiHandle * h);

void free_device(MppiHandle * h)
{
  auto F = [](auto *& p) {if (p) {cudaFree(p); p = nullptr;}};
  for (auto & p : h->d_noise) {F(p);}
  for (auto & p : h->d_noise_next) {F(p);}
  h->noise_prefetched = false;
  for (auto & p : h->d_cv) {F(p);}
  for (auto & p : h->d_state) {F(p);}
  F(h->d_u); F(h->d_u_saved); F(h->d_hist); F(h->d_hist_saved); F(h->d_terms); F(h->d_obs_raw);
  F(h->d_obs_rep); F(h->d_last); F(h->d_pa); F(h->d_costs); F(h->d_softmax); F(h->d_critic_costs);
  F(h->d_collisions); F(h->d_dbg_cost); F(h->d_dbg_obs); F(h->d_red); F(h->d_partials); F(h->d_slot);
  F(h->d_vis_xy); F(h->d_phase_clocks); F(h->d_slot_all); F(h->d_static); F(h->d_cycle_saved);
  h->d_out_tag = nullptr;  // alias of the mapped h_out_tag
  // the arena (costmaps + cycle block) outlives a reconfiguration: mppi_destroy frees it
  if (h->h_out_tag) {cudaFreeHost(h->h_out_tag); h->h_out_tag = nullptr;}
  std::free(h->h_out); h->h_out = nullptr;
}

// MotionModel::offsetSteps (motion_models.hpp:221-227) and predict's own rounding (:160-162)
int offset_steps(float delay, float model_dt)
{
  if (delay <= 0.0f || model_dt <= 0.0f) {return 0;}
  return static_cast<int>(std::floor((delay / model_dt) + 0.5));
}

// InflationLayer::computeCost via the C ABI helper
float find_circumscribed_cost(const MppiHandle * h, const HostRobot & rb, bool cost_critic)
{
  // cost_critic.cpp:85-129 / obstacles_critic.cpp:70-103
  double result = -1.0;
  const double circum_radius = h->cfg.circumscribed_radius;
  if (static_cast<float>(circum_radius) == 0.0f) {
    return 0.0f;  // circumscribed_radius_ starts at 0: early return of the cached 0
  }
  if (h->cfg.has_inflation_layer) {
    if (cost_critic && h->cfg.inflation_radius < circum_radius) {
      return 0.0f;
    }
    result = mppi_inflation_compute_cost(
      circum_radius / rb.resolution, rb.resolution, h->cfg.inscribed_radius,
      h->cfg.inflation_cost_scaling_factor);
  }
  return static_cast<float>(result);
}

int validate_and_derive(MppiHandle * h, const MppiConfig & c, std::string & err)
{
  if (c.abi_version != MPPI_ABI_VERSION) {err = "abi_version mismatch"; return MPPI_ERR_INVALID_ARGUMENT;}
  if (c.n_robots < 1 || c.batch_size < 1 || c.time_steps < 2 || c.iteration_count < 1) {
    err = "n_robots, batch_size, iteration_count must be >= 1 and time_steps >= 2";
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  if (c.n_critics > MPPI_MAX_CRITICS) {err = "too many critics"; return MPPI_ERR_INVALID_ARGUMENT;}
  if (c.motion_model < 0 || c.motion_model > 2) {
    err = "Failed to load motion model plugin";  // optimizer.cpp:683
    return MPPI_ERR_CONFIG;
  }
  if (c.shard_world < 1 || c.shard_rank >= c.shard_world || c.batch_size % c.shard_world != 0) {
    err = "batch_size must be divisible by shard_world and shard_rank < shard_world";
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  if (!(c.model_dt > 0.0f) || !(c.controller_frequency > 0.0) || !(c.temperature > 0.0f)) {
    err = "model_dt, controller_frequency and temperature must be positive";
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  h->cfg = c;
  h->R = static_cast<int>(c.n_robots);
  h->Bglobal = static_cast<int>(c.batch_size);
  h->B = static_cast<int>(c.batch_size / c.shard_world);
  h->T = static_cast<int>(c.time_steps);
  h->holo = c.motion_model == MPPI_MODEL_OMNI;
  if (h->cfg.sgf_order < 1 || h->cfg.sgf_order > 2) {h->cfg.sgf_order = 2;}  // optimizer.cpp:130-133

  // optimizer.cpp:135-148 sign fix-ups
  h->ax_max = std::fabs(c.ax_max);
  h->ax_min = c.ax_min > 0.0f ? -1.0f * c.ax_min : c.ax_min;
  h->ay_max = c.ay_max;
  h->ay_min = c.ay_min > 0.0f ? -1.0f * c.ay_min : c.ay_min;
  h->az_max = c.az_max;
  h->base_vx_max = c.vx_max; h->base_vx_min = c.vx_min; h->base_vy = c.vy_max; h->base_wz = c.wz_max;
  h->c_vx_max = c.vx_max; h->c_vx_min = c.vx_min; h->c_vy = c.vy_max; h->c_wz = c.wz_max;

  // optimizer.cpp:157-181 controller period / setOffset
  h->controller_period = static_cast<float>(1.0 / c.controller_frequency);
  {
    constexpr double eps = 1e-6;
    const double controller_period = h->controller_period;
    h->shift = false;
    if ((controller_period + eps) < c.model_dt) {
      // "Controller period is less then model dt, consider setting it equal"
    } else if (std::abs(controller_period - c.model_dt) < eps) {
      h->shift = true;
    } else {
      err = "Controller period more then model dt, set it equal to model dt";
      return MPPI_ERR_CONFIG;
    }
  }
  h->off_vx = offset_steps(c.model_delay_vx, c.model_dt);
  h->off_vy = offset_steps(c.model_delay_vy, c.model_dt);
  h->off_wz = offset_steps(c.model_delay_wz, c.model_dt);
  if (h->off_vx > MPPI_MAX_DELAY_STEPS || h->off_vy > MPPI_MAX_DELAY_STEPS || h->off_wz > MPPI_MAX_DELAY_STEPS) {
    err = "model_delay_* exceeds MPPI_MAX_DELAY_STEPS model steps";
    return MPPI_ERR_UNSUPPORTED;
  }
  // optimal_trajectory_validator.hpp:87-97
  h->validator_samples = static_cast<uint32_t>(c.collision_lookahead_time / c.model_dt);
  if (h->validator_samples > c.time_steps) {h->validator_samples = c.time_steps;}

  // critics (critic_manager.cpp:44-69 + each initialize())
  DevStatic & s = h->h_static;
  std::memset(&s, 0, sizeof(s));
  for (int i = 0; i < MPPI_CRITIC_TYPE_COUNT; ++i) {s.critic_of_type[i] = -1; h->critic_of_type[i] = -1;}
  s.n_critics = static_cast<int32_t>(c.n_critics);
  h->n_pa = 0; h->pa_step = 1; h->cc_ns = 0;
  for (uint32_t k = 0; k < c.n_critics; ++k) {
    const MppiCriticParams & p = c.critics[k];
    if (p.type < 0 || p.type >= MPPI_CRITIC_TYPE_COUNT) {
      err = "unknown critic type";  // pluginlib would fail to load the class
      return MPPI_ERR_CONFIG;
    }
    if (s.critic_of_type[p.type] >= 0) {
      err = "a critic type may be listed only once";
      return MPPI_ERR_UNSUPPORTED;
    }
    s.critic_of_type[p.type] = static_cast<int32_t>(k);
    h->critic_of_type[p.type] = static_cast<int>(k);
    DevCritic & d = s.critics[k];
    d.type = p.type;
    d.enabled = p.enabled;
    d.power = p.cost_power;
    d.weight = p.type == MPPI_CRITIC_COST ? p.cost_weight / 254.0f : p.cost_weight;  // cost_critic.cpp:39
    d.consider_footprint = p.consider_footprint;
    d.critical_cost = p.critical_cost;
    d.near_collision_cost = static_cast<float>(p.near_collision_cost);
    d.collision_cost = p.collision_cost;
    d.step = p.trajectory_point_step;
    d.repulsion_weight = p.repulsion_weight;
    d.critical_weight = p.critical_weight;
    d.collision_margin_distance = p.collision_margin_distance;
    d.symmetric_yaw_tolerance = p.symmetric_yaw_tolerance;
    d.max_path_occupancy_ratio = p.max_path_occupancy_ratio;
    d.offset_from_furthest = p.offset_from_furthest;
    d.use_path_orientations = p.use_path_orientations;
    d.max_angle_to_furthest = p.max_angle_to_furthest;
    d.mode = p.mode;
    d.deadband[0] = std::fabs(p.deadband_velocities[0]);
    d.deadband[1] = std::fabs(p.deadband_velocities[1]);
    d.deadband[2] = std::fabs(p.deadband_velocities[2]);
    if (p.type == MPPI_CRITIC_COST || p.type == MPPI_CRITIC_OBSTACLES) {
      // cost_critic.cpp:57-71, obstacles_critic.cpp:46-60
      if ((c.use_radius != 0) == (p.consider_footprint != 0) && c.use_radius) {
        err = "Considering footprint in collision checking but no robot footprint provided in the costmap";
        return MPPI_ERR_CONFIG;
      }
      if (p.consider_footprint && c.n_footprint < 3) {
        err = "consider_footprint needs a footprint polygon of at least 3 points";
        return MPPI_ERR_CONFIG;
      }
    }
    if ((p.type == MPPI_CRITIC_COST || p.type == MPPI_CRITIC_PATH_ALIGN) && p.trajectory_point_step < 1) {
      err = "trajectory_point_step must be >= 1";
      return MPPI_ERR_INVALID_ARGUMENT;
    }
    if (p.type == MPPI_CRITIC_COST) {
      h->cc_ns = (h->T - 1) / static_cast<int>(p.trajectory_point_step) + 1;
    }
    if (p.type == MPPI_CRITIC_PATH_ALIGN) {
      h->pa_step = static_cast<int>(p.trajectory_point_step);
      h->n_pa = (h->T - 1) / h->pa_step + 1;
    }
    if (p.type == MPPI_CRITIC_PATH_ANGLE) {
      // path_angle_critic.cpp:25-50
      bool reversing_allowed = true;
      if (std::fabs(c.vx_min) < 1e-6f) {reversing_allowed = false;}
      int mode = p.mode;
      if (!reversing_allowed && mode == 1) {mode = 0;}
      d.mode = mode;
    }
  }
  if (c.validator_enabled && c.validator_consider_footprint && c.n_footprint < 3) {
    err = "validator consider_footprint needs a footprint polygon";
    return MPPI_ERR_CONFIG;
  }
  if (c.n_footprint > MPPI_MAX_FOOTPRINT) {err = "footprint too large"; return MPPI_ERR_INVALID_ARGUMENT;}
  s.n_footprint = static_cast<int32_t>(c.n_footprint);
  for (uint32_t i = 0; i < c.n_footprint; ++i) {s.footprint_x[i] = c.footprint_x[i]; s.footprint_y[i] = c.footprint_y[i];}
  s.track_unknown = c.track_unknown;
  // ObstaclesCritic::distanceToObstacle (obstacles_critic.cpp:105-118) tabulated per uint8 cost.
  // `log(cost.cost)` there is the C double log applied to a float.
  {
    // inflation_scale_factor_ / inflation_radius_ are only filled inside findCircumscribedCost when a
    // layer exists and the circumscribed radius differs from the cached 0 (obstacles_critic.cpp:75-91)
    const bool have = c.has_inflation_layer && static_cast<float>(c.circumscribed_radius) != 0.0f;
    const float scale_factor = have ? static_cast<float>(c.inflation_cost_scaling_factor) : 0.0f;
    const float min_radius = static_cast<float>(c.inscribed_radius);
    s.obstacle_scale_factor = scale_factor;
    s.obstacle_inflation_radius = have ? static_cast<float>(c.inflation_radius) : 0.0f;
    for (int v = 0; v < 256; ++v) {
      const float cost = static_cast<float>(v);
      float dist_to_obj = (scale_factor * min_radius - log(cost) + log(253.0f)) / scale_factor;
      s.obstacle_dist_lut[1][v] = dist_to_obj;
      dist_to_obj -= min_radius;
      s.obstacle_dist_lut[0][v] = dist_to_obj;
    }
  }
  s.param_vx_max = c.vx_max; s.param_vx_min = c.vx_min; s.param_vy_max = c.vy_max;
  s.min_turning_r = c.min_turning_r;
  s.motion_model = c.motion_model;
  s.max_delta_vx = c.model_dt * h->ax_max;
  s.min_delta_vx = c.model_dt * h->ax_min;
  s.max_delta_vy = c.model_dt * h->ay_max;
  s.min_delta_vy = c.model_dt * h->ay_min;
  s.max_delta_wz = c.model_dt * h->az_max;
  s.offset_vx = h->off_vx; s.offset_vy = h->off_vy; s.offset_wz = h->off_wz;
  s.clamp_raw_controls = c.clamp_raw_controls;
  s.model_dt = c.model_dt;
  s.controller_period = h->controller_period;
  s.shift_control_sequence = h->shift ? 1 : 0;
  s.gamma_vx = c.gamma / (c.vx_std * c.vx_std);
  s.gamma_wz = c.gamma / (c.wz_std * c.wz_std);
  s.gamma_vy = c.gamma / (c.vy_std * c.vy_std);
  s.use_gamma_wz = c.wz_std > 0.0f ? 1 : 0;
  s.inv_temp = 1.0f / c.temperature;
  s.sgf_order = h->cfg.sgf_order;
  s.ax_max = h->ax_max; s.ax_min = h->ax_min; s.ay_max = h->ay_max; s.ay_min = h->ay_min; s.az_max = h->az_max;
  s.validator_enabled = c.validator_enabled;
  s.validator_consider_footprint = c.validator_consider_footprint;
  s.validator_samples = h->validator_samples;
  return MPPI_OK;
}

int allocate(MppiHandle * h)
{
  drop_graphs(h);
  free_device(h);
  const size_t R = h->R, B = h->B, T = h->T;
  const size_t bt = R * B * T;
  const bool mat = h->cfg.materialize_tensors != 0;
  const int n_critics = static_cast<int>(h->cfg.n_critics);
  for (int i = 0; i < 3; ++i) {CU_CHECK(h, cudaMalloc(&h->d_noise[i], bt * sizeof(float)));}
  CU_CHECK(h, cudaMemset(h->d_noise[1], 0, bt * sizeof(float)));
  if (h->cfg.regenerate_noises && h->noise_prefetch) {
    for (int i = 0; i < 3; ++i) {CU_CHECK(h, cudaMalloc(&h->d_noise_next[i], bt * sizeof(float)));}
    CU_CHECK(h, cudaMemset(h->d_noise_next[1], 0, bt * sizeof(float)));
    if (!h->noise_event) {CU_CHECK(h, cudaEventCreateWithFlags(&h->noise_event, cudaEventDisableTiming));}
  }
  if (mat || h->cfg.clamp_raw_controls) {
    for (int i = 0; i < 3; ++i) {
      CU_CHECK(h, cudaMalloc(&h->d_cv[i], bt * sizeof(float)));
      CU_CHECK(h, cudaMemset(h->d_cv[i], 0, bt * sizeof(float)));
    }
  }
  if (mat) {
    for (int i = 0; i < 6; ++i) {
      CU_CHECK(h, cudaMalloc(&h->d_state[i], bt * sizeof(float)));
      CU_CHECK(h, cudaMemset(h->d_state[i], 0, bt * sizeof(float)));
    }
  }
  CU_CHECK(h, cudaMalloc(&h->d_u, R * 3 * T * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_u_saved, R * 3 * T * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_hist, R * 12 * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_hist_saved, R * 12 * sizeof(float)));
  CU_CHECK(h, cudaMemset(h->d_u, 0, R * 3 * T * sizeof(float)));
  CU_CHECK(h, cudaMemset(h->d_hist, 0, R * 12 * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_terms, R * (n_critics + 3) * B * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_obs_raw, R * B * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_obs_rep, R * B * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_last, R * 3 * B * sizeof(float)));
  // one row per trajectory plus a spare row (where the tail threads of k_rollout_lean's last block park their stores)
  CU_CHECK(h, cudaMalloc(&h->d_pa, std::max<size_t>(4, (R * B + 1) * static_cast<size_t>((3 * h->n_pa + 3) & ~3)) * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_costs, R * B * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_softmax, R * B * sizeof(float)));
  CU_CHECK(h, cudaMemset(h->d_costs, 0, R * B * sizeof(float)));
  CU_CHECK(h, cudaMemset(h->d_softmax, 0, R * B * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_collisions, R * B));
  CU_CHECK(h, cudaMemset(h->d_collisions, 0, R * B));
  if (h->cfg.visualize && n_critics > 0) {
    CU_CHECK(h, cudaMalloc(&h->d_critic_costs, R * n_critics * B * sizeof(float)));
    CU_CHECK(h, cudaMemset(h->d_critic_costs, 0, R * n_critics * B * sizeof(float)));
  }
  h->vis_n_traj = 0; h->vis_n_pts = 0;
  if (h->cfg.visualize && h->cfg.vis_trajectory_step > 0 && h->cfg.vis_time_step > 0) {
    h->vis_n_traj = static_cast<int>((B + h->cfg.vis_trajectory_step - 1) / h->cfg.vis_trajectory_step);
    h->vis_n_pts = static_cast<int>((T + h->cfg.vis_time_step - 1) / h->cfg.vis_time_step);
    const size_t n = static_cast<size_t>(R) * h->vis_n_traj * h->vis_n_pts * 2;
    CU_CHECK(h, cudaMalloc(&h->d_vis_xy, n * sizeof(float)));
    CU_CHECK(h, cudaMemset(h->d_vis_xy, 0, n * sizeof(float)));
  }
  if (h->cfg.debug_cells) {
    if (h->cc_ns > 0) {
      CU_CHECK(h, cudaMalloc(&h->d_dbg_cost, R * h->cc_ns * B * sizeof(uint32_t)));
      CU_CHECK(h, cudaMemset(h->d_dbg_cost, 0xFF, R * h->cc_ns * B * sizeof(uint32_t)));
    }
    if (h->critic_of_type[MPPI_CRITIC_OBSTACLES] >= 0) {
      CU_CHECK(h, cudaMalloc(&h->d_dbg_obs, bt * sizeof(uint32_t)));
      CU_CHECK(h, cudaMemset(h->d_dbg_obs, 0xFF, bt * sizeof(uint32_t)));
    }
  }
  if (const char * e = std::getenv("MPPI_B200_NO_SINGLE_COPY"); e && e[0] == '1') {h->single_copy = false;}
  if (const char * e = std::getenv("MPPI_B200_NO_WINDOW_ONLY"); e && e[0] == '1') {h->window_only = false;}
  if (const char * e = std::getenv("MPPI_B200_NO_WINDOW_GENERAL"); e && e[0] == '1') {h->window_only_general = false;}
  if (const char * e = std::getenv("MPPI_B200_NO_POLL"); e && e[0] == '1') {h->poll_results = false;}
  if (const char * e = std::getenv("MPPI_B200_ZERO_COPY"); e) {h->zero_copy = e[0] == '1';}
  if (const char * e = std::getenv("MPPI_B200_NO_NOISE_PREFETCH"); e && e[0] == '1') {h->noise_prefetch = false;}
  if (const char * e = std::getenv("MPPI_B200_WINDOW_HALF"); e && std::atoi(e) > 0) {h->window_half_override = std::atoi(e);}
  if (const char * e = std::getenv("MPPI_B200_PHASE_CLOCKS"); e && e[0] == '1') {
    CU_CHECK(h, cudaMalloc(&h->d_phase_clocks, R * MPPI_PHASE_CLOCK_WORDS * sizeof(long long)));
    CU_CHECK(h, cudaMemset(h->d_phase_clocks, 0, R * MPPI_PHASE_CLOCK_WORDS * sizeof(long long)));
  }
  CU_CHECK(h, cudaMalloc(&h->d_red, R * RED_WORDS * sizeof(int32_t)));
  {
    std::vector<int32_t> init(R * RED_WORDS, 0);
    for (size_t r = 0; r < R; ++r) {
      init[r * RED_WORDS + RED_REP_MIN] = INT32_MIN;
      init[r * RED_WORDS + RED_COST_MIN] = INT32_MIN;
    }
    CU_CHECK(h, cudaMemcpy(h->d_red, init.data(), init.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  {
    const size_t tile = static_cast<size_t>(MPPI_BLOCK_C) * ((B % 4 == 0) ? 4 : 1);  // trajectories per block pass
    h->n_partial_blocks = static_cast<int>(std::min<size_t>((B + tile - 1) / tile, MPPI_MAX_PARTIAL_BLOCKS * 2));
  }
  if (R > 1) {
    // fleet: keep the total grid near two waves of 148 SMs
    const int per_robot = std::max<int>(1, static_cast<int>(MPPI_MAX_PARTIAL_BLOCKS / R));
    h->n_partial_blocks = std::max(1, std::min(h->n_partial_blocks, per_robot));
  }
  CU_CHECK(h, cudaMalloc(&h->d_partials, R * h->n_partial_blocks * (1 + 3 * T) * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_slot, R * (2 + 3 * T) * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_slot_all, static_cast<size_t>(h->cfg.shard_world) * R * (2 + 3 * T) * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_static, sizeof(DevStatic)));
  CU_CHECK(h, cudaMemcpy(h->d_static, &h->h_static, sizeof(DevStatic), cudaMemcpyHostToDevice));
  if (h->max_path < 256) {h->max_path = 256;}
  h->out_stride = (sizeof(DevOutHeader) + 6 * T * sizeof(float) + 15) & ~static_cast<size_t>(15);
  // result blocks live in mapped pinned host memory: the finalize kernel's ~350 words per robot go straight over
  // PCIe as posted 8-byte writes, each carrying the cycle's sequence number (KArgs::out_tag): no D2H copy node, no
  // fence in the kernel.  unpack_results() moves a finished block into the plain layout the host code reads.
  CU_CHECK(h, cudaHostAlloc(&h->h_out_tag, R * (h->out_stride / 4) * sizeof(uint2), cudaHostAllocMapped));
  std::memset(h->h_out_tag, 0, R * (h->out_stride / 4) * sizeof(uint2));
  CU_CHECK(h, cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_out_tag), h->h_out_tag, 0));
  h->h_out = static_cast<uint8_t *>(std::calloc(R, h->out_stride));
  if (!h->h_out) {h->error = "out of host memory"; return MPPI_ERR_CUDA;}
  return MPPI_OK;
}

size_t cycle_block_size(int R, int path_cap)
{
  const size_t n = sizeof(DevCycle) * R + sizeof(float) * 4 * path_cap * R + static_cast<size_t>(path_cap) * R;
  return (n + 15) & ~static_cast<size_t>(15);
}

// (Re)allocates the arenas when a costmap larger than map_stride or a path longer than max_path arrives.
// Costmaps already staged are carried over on the host side and marked stale on the device side.
int ensure_arena(MppiHandle * h, size_t need_stride, int need_path)
{
  const bool have = h->h_arena != nullptr && h->d_cycle_saved != nullptr;
  if (have && need_stride <= h->map_stride && need_path <= h->max_path) {return MPPI_OK;}
  int cap = std::max(h->max_path, 1);
  while (cap < need_path) {cap *= 2;}
  if (cap > MPPI_MAX_PATH_POINTS) {
    h->error = "path longer than MPPI_MAX_PATH_POINTS";
    return MPPI_ERR_UNSUPPORTED;
  }
  const size_t stride = std::max(h->map_stride, need_stride);
  const bool same_shape = h->h_arena != nullptr && stride == h->map_stride && cap == h->max_path;
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  CU_CHECK(h, cudaStreamSynchronize(h->side));
  drop_graphs(h);
  const size_t block = cycle_block_size(h->R, cap);
  if (!same_shape) {
    uint8_t * nh = nullptr, * nd = nullptr;
    // window packets: room for every robot's largest window (single-launch cycles and the general kernels both take
    // window-only uploads)
    const size_t packets = static_cast<size_t>(h->R) * MPPI_MAX_WINDOW_BYTES;
    const size_t total = stride * h->R + block + packets;
    CU_CHECK(h, cudaMallocHost(&nh, total));
    std::memset(nh, 0, total);
    if (cudaMalloc(&nd, total) != cudaSuccess || cudaMemset(nd, 0, total) != cudaSuccess) {
      cudaFreeHost(nh);
      if (nd) {cudaFree(nd);}
      h->error = "out of device memory for the costmap / cycle arena";
      return MPPI_ERR_CUDA;
    }
    for (int r = 0; r < h->R; ++r) {
      HostRobot & rb = h->robots[r];
      if (rb.map_valid && rb.map_host) {
        std::memcpy(nh + stride * r, rb.map_host, rb.map_bytes);
        if (rb.map_src == rb.map_host) {rb.map_src = nh + stride * r;}
        rb.map_host = nh + stride * r;
        rb.device_stale = true;
      } else if (rb.map_valid && rb.device_only && h->d_arena) {
        // a map that exists only on the device moves with the arena
        if (cudaMemcpy(nd + stride * r, h->d_arena + h->map_stride * r, rb.map_bytes, cudaMemcpyDeviceToDevice) != cudaSuccess) {
          rb.map_valid = false;
        }
      }
      rb.stage_pending = false;  // both streams are idle
    }
    if (h->h_arena) {cudaFreeHost(h->h_arena);}
    if (h->d_arena) {cudaFree(h->d_arena);}
    h->h_arena = nh; h->d_arena = nd;
    h->map_stride = stride;
    h->max_path = cap;
    h->cycle_block_bytes = block;
    h->h_map_stage = nh; h->d_map = nd;
    h->h_cycle_block = nh + stride * h->R;
    h->d_cycle_block = nd + stride * h->R;
    h->h_packets = h->h_cycle_block + block;
    h->d_packets = h->d_cycle_block + block;
    h->packets_capacity = packets;
  }
  if (h->d_cycle_saved) {cudaFree(h->d_cycle_saved); h->d_cycle_saved = nullptr;}
  CU_CHECK(h, cudaMalloc(&h->d_cycle_saved, block));
  h->have_cycle = false;
  return MPPI_OK;
}

int ensure_cycle_block(MppiHandle * h, int need_path)
{
  return ensure_arena(h, h->map_stride, need_path);
}

// next = false: into the tensors the coming kernels read, on the handle's stream; next = true: into the other set,
// on the side stream (prefetch for the following cycle)
int generate_noise(MppiHandle * h, int robot, bool next = false)
{
  // one robot at a time: pointers offset to that robot's slab, R = 1, robot index in the counter
  const size_t slab = static_cast<size_t>(h->B) * h->T;
  HostRobot & rb = h->robots[robot];
  // robot index and epoch are folded into the Philox key so each (robot, reset) draws a fresh tensor
  const uint64_t seed = h->cfg.noise_seed ^ (static_cast<uint64_t>(robot) * 0x9E3779B97F4A7C15ull);
  float ** set = next ? h->d_noise_next : h->d_noise;
  CU_CHECK(h, mppi_launch_philox(
      set[0] + robot * slab, set[1] + robot * slab, set[2] + robot * slab,
      h->B, h->T, 1, static_cast<int>(h->cfg.shard_rank) * h->B, seed, rb.noise_epoch,
      h->cfg.vx_std, h->cfg.vy_std, h->cfg.wz_std, h->holo, next ? h->side : h->stream));
  rb.noise_epoch++;
  rb.noise_injected = false;
  h->launches++;
  return MPPI_OK;
}

// Optimizer::reset (optimizer.cpp:183-211) for one robot
int reset_robot(MppiHandle * h, int r, bool reset_dynamic_speed_limits)
{
  const size_t T = h->T;
  CU_CHECK(h, cudaMemsetAsync(h->d_u + r * 3 * T, 0, 3 * T * sizeof(float), h->stream));
  CU_CHECK(h, cudaMemsetAsync(h->d_hist + r * 12, 0, 12 * sizeof(float), h->stream));
  CU_CHECK(h, cudaMemsetAsync(h->d_costs + static_cast<size_t>(r) * h->B, 0, h->B * sizeof(float), h->stream));
  HostRobot & rb = h->robots[r];
  rb.last_cmd_vx = rb.last_cmd_vy = rb.last_cmd_wz = 0.0;
  if (reset_dynamic_speed_limits) {
    h->c_vx_max = h->base_vx_max; h->c_vx_min = h->base_vx_min; h->c_vy = h->base_vy; h->c_wz = h->base_wz;
  }
  // motion_model_->setConstraints resizes, clearCommandHistory zeroes (motion_models.hpp:78-80,97-102)
  rb.hist_vx.assign(h->off_vx, 0.0f);
  rb.hist_vy.assign(h->off_vy, 0.0f);
  rb.hist_wz.assign(h->off_wz, 0.0f);
  // noise_generator_.reset (noise_generator.cpp:77-96): regenerate unless reference noise is injected
  if (!rb.noise_injected) {
    int rc = generate_noise(h, r);
    if (rc != MPPI_OK) {return rc;}
  }
  return MPPI_OK;
}

void push_one(std::vector<float> & buf, float v)  // motion_models.hpp:232-237
{
  if (buf.empty()) {return;}
  std::rotate(buf.begin(), buf.begin() + 1, buf.end());
  buf.back() = v;
}

// Fill the per-robot cycle block: Optimizer::prepare (optimizer.cpp:300-343) plus the scalar
// gates every critic's score() evaluates before touching tensors.
int prepare_robot(MppiHandle * h, int r, const MppiCycleIn & in, int preset_furthest, bool from_tensors)
{
  HostRobot & rb = h->robots[r];
  if (!rb.map_valid) {h->error = "no costmap uploaded for robot"; return MPPI_ERR_NOT_READY;}
  if (in.n_path > 0 && (!in.path_x || !in.path_y || !in.path_yaw)) {
    h->error = "path arrays are NULL";
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  DevCycle & cy = *h->hcyc(r);
  std::memset(&cy, 0, sizeof(cy));
  const MppiConfig & c = h->cfg;

  // optimizer.cpp:307-330 latency-compensating speed prediction (double)
  double sp_vx = in.speed_vx, sp_vy = in.speed_vy, sp_wz = in.speed_wz;
  if (c.open_loop) {
    sp_vx = rb.last_cmd_vx; sp_vy = rb.last_cmd_vy; sp_wz = rb.last_cmd_wz;
  } else {
    const double dt = h->controller_period;
    sp_vx = std::clamp(rb.last_cmd_vx, in.speed_vx + dt * h->ax_min, in.speed_vx + dt * h->ax_max);
    sp_wz = std::clamp(rb.last_cmd_wz, in.speed_wz - dt * h->az_max, in.speed_wz + dt * h->az_max);
    if (h->holo) {
      sp_vy = std::clamp(rb.last_cmd_vy, in.speed_vy + dt * h->ay_min, in.speed_vy + dt * h->ay_max);
    }
  }
  cy.speed_vx = static_cast<float>(sp_vx);
  cy.speed_vy = static_cast<float>(sp_vy);
  cy.speed_wz = static_cast<float>(sp_wz);
  cy.pose_x_d = in.pose_x; cy.pose_y_d = in.pose_y; cy.pose_yaw_d = in.pose_yaw;
  cy.pose_x = static_cast<float>(in.pose_x);
  cy.pose_y = static_cast<float>(in.pose_y);
  cy.pose_yaw = static_cast<float>(in.pose_yaw);
  mppi_sincosf(cy.pose_yaw, &cy.init_sin, &cy.init_cos);

  // nav2_util/geometry_utils.hpp:155-165 calculate_path_length
  double path_length = 0.0;
  if (in.n_path >= 2) {
    for (size_t i = 0; i + 1 < in.n_path; ++i) {
      path_length += std::hypot(in.path_x[i] - in.path_x[i + 1], in.path_y[i] - in.path_y[i + 1]);
    }
  }
  if (in.has_local_path_length) {path_length = in.local_path_length;}
  cy.local_path_length = static_cast<float>(path_length);
  const int n = static_cast<int>(in.n_path);
  cy.n_path = n;

  // utils::toTensor (tools/utils.hpp:208-220) + integrated distances (utils.hpp:299-304)
  float * px = h->hpath(r), * py = px + h->max_path, * pyaw = py + h->max_path, * pd = pyaw + h->max_path;
  for (int i = 0; i < n; ++i) {
    px[i] = static_cast<float>(in.path_x[i]);
    py[i] = static_cast<float>(in.path_y[i]);
    pyaw[i] = static_cast<float>(in.path_yaw[i]);
  }
  if (n > 0) {pd[0] = 0.0f;}
  for (int i = 1; i < n; ++i) {
    const float dx = px[i] - px[i - 1];
    const float dy = py[i] - py[i - 1];
    pd[i] = pd[i - 1] + sqrtf(dx * dx + dy * dy);
  }

  // costmap geometry (cost_critic.cpp:139-144 refreshes it every score())
  cy.origin_x_d = rb.origin_x; cy.origin_y_d = rb.origin_y; cy.resolution_d = rb.resolution;
  cy.origin_x_f = static_cast<float>(rb.origin_x);
  cy.origin_y_f = static_cast<float>(rb.origin_y);
  cy.resolution_f = static_cast<float>(rb.resolution);
  cy.size_x = rb.size_x; cy.size_y = rb.size_y; cy.pitch = rb.pitch;
  {
    // Bound on |float estimate - double quotient| of Costmap2D::worldToMap for points on this map (kernel A's
    // obstacle_cell): the float origin and resolution are off by <= 2^-24 relative, the float subtraction and division
    // round once each; in cells that is < 2^-24 (|origin| / res + 4 size).  Four times that, plus a floor.
    const double cells = std::max(std::fabs(rb.origin_x), std::fabs(rb.origin_y)) / rb.resolution + 4.0 * std::max(rb.size_x, rb.size_y);
    const double eps = std::ldexp(cells, -22) + 1e-6;
    cy.w2m_eps = eps < 0.2 ? static_cast<float>(eps) : 2.0f;
  }

  // utils::findPathCosts (tools/utils.hpp:358-387): validity of path points 0 .. n-2
  uint8_t * valid = h->hvalid(r);
  std::memset(valid, 0, h->max_path);
  cy.valid_on_device = rb.device_only ? 1 : 0;   // no host copy of the map: the kernels evaluate the same rule
  for (int i = 0; !rb.device_only && i + 1 < n; ++i) {
    const double wx = px[i], wy = py[i];
    bool ok = false;
    if (!(wx < rb.origin_x || wy < rb.origin_y)) {
      const unsigned int mx = static_cast<unsigned int>((wx - rb.origin_x) / rb.resolution);
      const unsigned int my = static_cast<unsigned int>((wy - rb.origin_y) / rb.resolution);
      if (mx < rb.size_x && my < rb.size_y) {
        const uint8_t cost = rb.map_src[static_cast<size_t>(my) * rb.map_src_pitch + mx];
        if (cost == 254 || cost == 253) {
          ok = false;
        } else if (cost == 255) {
          ok = c.track_unknown != 0;
        } else {
          ok = true;
        }
      }
    }
    valid[i] = ok ? 1 : 0;
  }

  // critic gates (the scalar early-returns at the top of each score())
  const float lpl = cy.local_path_length;
  uint32_t active = 0, near_goal = 0;
  bool need_furthest = false;
  for (uint32_t k = 0; k < c.n_critics; ++k) {
    const MppiCriticParams & p = c.critics[k];
    bool on = p.enabled != 0;
    switch (p.type) {
      case MPPI_CRITIC_CONSTRAINT: break;
      case MPPI_CRITIC_COST:
      case MPPI_CRITIC_OBSTACLES:
        if (lpl < p.near_goal_distance) {near_goal |= 1u << k;}
        cy.possible_collision_cost[k] = find_circumscribed_cost(h, rb, p.type == MPPI_CRITIC_COST);
        break;
      case MPPI_CRITIC_GOAL:        // goal_critic.cpp:37-39
      case MPPI_CRITIC_GOAL_ANGLE:  // goal_angle_critic.cpp:38-40
        on = on && !(lpl > p.threshold_to_consider) && n > 0;
        break;
      case MPPI_CRITIC_PATH_ALIGN:  // path_align_critic.cpp:42-44
      case MPPI_CRITIC_PATH_ANGLE:  // path_angle_critic.cpp:64-66
        on = on && !(lpl < p.threshold_to_consider) && n > 0;
        if (on) {need_furthest = true;}
        break;
      case MPPI_CRITIC_PATH_FOLLOW:  // path_follow_critic.cpp:40-42
        on = on && !(n < 2 || lpl < p.threshold_to_consider);
        if (on) {need_furthest = true;}
        break;
      case MPPI_CRITIC_PREFER_FORWARD:  // prefer_forward_critic.cpp:42-44
        on = on && !(lpl < p.threshold_to_consider);
        break;
      case MPPI_CRITIC_TWIRLING:  // twirling_critic.cpp:39-48
        if (in.has_goal_checker && lpl < in.goal_checker_xy_tolerance) {on = false;}
        break;
      case MPPI_CRITIC_VELOCITY_DEADBAND: break;
      default: break;
    }
    if (on) {active |= 1u << k;}
  }
  cy.active_mask = active;
  cy.near_goal_mask = near_goal;
  cy.need_furthest = need_furthest ? 1 : 0;
  cy.need_path_valid = 1;
  if (n > 0) {
    // utils::getLastPathPose (utils.hpp:227-245): float path point -> double Pose -> float in the critics
    cy.goal_x = px[n - 1];
    cy.goal_y = py[n - 1];
    cy.goal_yaw = pyaw[n - 1];
    cy.sym_goal_yaw = static_cast<float>(mppi_angles_normalize_angle(cy.goal_yaw + M_PI));  // goal_angle_critic.cpp:47
  }
  cy.c_vx_max = h->c_vx_max; cy.c_vx_min = h->c_vx_min; cy.c_vy = h->c_vy; cy.c_wz = h->c_wz;
  for (size_t i = 0; i < rb.hist_vx.size(); ++i) {cy.hist_vx[i] = rb.hist_vx[i];}
  for (size_t i = 0; i < rb.hist_vy.size(); ++i) {cy.hist_vy[i] = rb.hist_vy[i];}
  for (size_t i = 0; i < rb.hist_wz.size(); ++i) {cy.hist_wz[i] = rb.hist_wz[i];}
  cy.track_collisions = (c.visualize || from_tensors) ? 1 : 0;
  cy.furthest_cached = preset_furthest >= 0 ? preset_furthest : -1;
  cy.fail_flag = 0;
  cy.run = 1;  // run

  // costmap window staged in shared memory: the cells the trajectories can plausibly reach
  {
    const double reach_v = std::max({std::fabs(static_cast<double>(h->c_vx_max)),
          std::fabs(static_cast<double>(h->c_vx_min)), h->holo ? static_cast<double>(h->c_vy) : 0.0});
    const double reach = reach_v * h->T * c.model_dt * 1.25 + c.circumscribed_radius + 4.0 * rb.resolution;
    int half = static_cast<int>(std::ceil(reach / rb.resolution));
    half = std::min(std::max(half, 8), MPPI_MAX_WINDOW_HALF);
    if (h->window_half_override > 0) {half = std::min(h->window_half_override, MPPI_MAX_WINDOW_HALF);}
    const int cx = static_cast<int>(std::floor((in.pose_x - rb.origin_x) / rb.resolution));
    const int cyy = static_cast<int>(std::floor((in.pose_y - rb.origin_y) / rb.resolution));
    int x0 = std::max(0, cx - half) & ~15;
    int x1 = std::min(static_cast<int>(rb.pitch), ((cx + half + 1) + 15) & ~15);
    int y0 = std::max(0, cyy - half);
    int y1 = std::min(static_cast<int>(rb.size_y), cyy + half + 1);
    if (x1 <= x0 || y1 <= y0 || x0 >= static_cast<int>(rb.pitch) || y0 >= static_cast<int>(rb.size_y)) {
      x0 = 0; x1 = 0; y0 = 0; y1 = 0;  // robot far off the map: global lookups only
    }
    cy.win_x0 = x0; cy.win_y0 = y0; cy.win_w = x1 - x0; cy.win_h = y1 - y0;
  }
  return MPPI_OK;
}

KArgs make_args(MppiHandle * h)
{
  KArgs a{};
  a.B = h->B; a.T = h->T; a.R = h->R;
  a.b_offset = static_cast<int>(h->cfg.shard_rank) * h->B;
  a.n_terms = static_cast<int>(h->cfg.n_critics) + 3;
  a.n_pa = h->n_pa; a.pa_step = h->pa_step; a.pa_stride = (3 * h->n_pa + 3) & ~3;
  a.max_path = h->max_path;
  a.world = static_cast<int>(h->cfg.shard_world); a.rank = static_cast<int>(h->cfg.shard_rank);
  a.n_partial_blocks = h->n_partial_blocks;
  a.materialize = h->cfg.materialize_tensors;
  a.last_iteration = 0;
  a.do_shift = h->shift ? 1 : 0;
  a.noise_vx = h->d_noise[0]; a.noise_vy = h->d_noise[1]; a.noise_wz = h->d_noise[2];
  a.cvx = h->d_cv[0]; a.cvy = h->d_cv[1]; a.cwz = h->d_cv[2];
  a.vx = h->d_state[0]; a.vy = h->d_state[1]; a.wz = h->d_state[2];
  a.x = h->d_state[3]; a.y = h->d_state[4]; a.yaw = h->d_state[5];
  a.u = h->d_u; a.ctrl_hist = h->d_hist;
  a.terms = h->d_terms; a.obs_raw = h->d_obs_raw; a.obs_rep = h->d_obs_rep; a.last_xyz = h->d_last;
  a.pa_samples = h->d_pa; a.costs = h->d_costs; a.softmax = h->d_softmax;
  a.critic_costs = h->d_critic_costs;
  a.collisions = h->d_collisions;
  a.dbg_cost_cells = h->d_dbg_cost; a.dbg_obstacle_cells = h->d_dbg_obs;
  a.red = h->d_red; a.partials = h->d_partials;
  a.ar2_slot = h->d_slot; a.ar2_all = h->cfg.shard_world == 1 ? h->d_slot : h->d_slot_all;
  if (h->peers_open && h->cfg.shard_world > 1) {
    a.xch_peers = h->d_peers;
    a.xch_state = h->d_xstate;
    a.xch_error = h->d_xerror;
    a.xch_words_base = static_cast<uint32_t>(h->mailbox_words_base);
    a.xch_words_stride = static_cast<uint32_t>(h->mailbox_words_stride);
    a.xch_slots_base = static_cast<uint32_t>(h->mailbox_slots_base);
    a.xch_slots_stride = static_cast<uint32_t>(h->mailbox_slots_stride);
  }
  a.costmap = h->d_map; a.map_stride = h->map_stride;
  a.win_packets = h->d_packets; a.map_resident = 1;
  a.stamp_results = 0;
  a.vis_xy = h->d_vis_xy;
  a.vis_traj_step = static_cast<int>(h->cfg.vis_trajectory_step); a.vis_time_step = static_cast<int>(h->cfg.vis_time_step);
  a.vis_n_traj = h->vis_n_traj; a.vis_n_pts = h->vis_n_pts;
  a.path = reinterpret_cast<const float *>(h->d_cycle_block + sizeof(DevCycle) * h->R);
  a.path_valid = h->d_cycle_block + sizeof(DevCycle) * h->R + sizeof(float) * 4 * h->max_path * h->R;
  a.cyc = reinterpret_cast<DevCycle *>(h->d_cycle_block);
  a.cyc_in = a.cyc; a.u_in = a.u; a.hist_in = a.ctrl_hist;
  a.st = h->d_static;
  a.phase_clocks = h->d_phase_clocks;
  a.out_tag = h->d_out_tag; a.out_words = h->out_stride / 4;
  {
    static const int dbg = [] {const char * e = std::getenv("MPPI_B200_DEBUG_FLAGS"); return e ? std::atoi(e) : 0;}();
    a.debug_flags = dbg; a.pad_debug = 0;
  }
  return a;
}

struct LaunchPlan
{
  KSmemA sm_a;
  KSmemA sm_tensors;  // score_tensors: no noise ring
  KPlanA ka;
  int block_a;
  int block_b;
  int path_cap;
  bool fused;          // one-cluster-per-robot single-launch path (small batches, lean configuration)
  KSmemF sm_f;
  int cluster_size;
};

}  // namespace
struct LaunchPlanBox {LaunchPlan plan;};
namespace
{

LaunchPlan make_plan(MppiHandle * h)
{
  LaunchPlan p{};
  int max_n = 0, win_bytes = 0;
  bool need = false;
  for (int r = 0; r < h->R; ++r) {
    const DevCycle & cy = *h->hcyc(r);
    max_n = std::max(max_n, cy.n_path);
    need = need || cy.need_furthest;
    win_bytes = std::max(win_bytes, cy.win_w * cy.win_h);
  }
  p.path_cap = std::max(max_n, 1);
  // small batches: narrower blocks spread the trajectories over more SMs (latency-bound regime)
  // (up to four 64-thread blocks per SM before switching to 256-thread blocks: 16384 trajectories are 256 small blocks
  // on all 148 SMs instead of 64 large ones on 64 of them)
  p.block_a = (static_cast<size_t>(h->B) * h->R <= 148u * 64u * 4u) ? 64 : MPPI_BLOCK_A;
  const int lag = std::max({std::max(h->off_vx, 1) - 1, std::max(h->off_wz, 1) - 1,
        h->holo ? std::max(h->off_vy, 1) - 1 : 0});
  const int ring = lag > 0 ? lag + 1 : 0;
  const bool obstacles = h->critic_of_type[MPPI_CRITIC_OBSTACLES] >= 0;
  // noise ring: every time-step chunk in flight if it fits in ~40 KB, else a 4-deep ring.
  // Bulk copies need 16-byte aligned column segments: B % 4 == 0 (blocks start at multiples of 64).
  const int axes = h->holo ? 3 : 2;
  const int n_chunks = (h->T + MPPI_TCHUNK_A - 1) / MPPI_TCHUNK_A;
  const size_t stage_bytes = sizeof(float) * axes * MPPI_TCHUNK_A * p.block_a;
  int stages = static_cast<int>(std::min<size_t>(n_chunks, std::max<size_t>(2, (40 * 1024) / stage_bytes)));
  stages = std::min(stages, MPPI_MAX_STAGES);
  if (h->B % 4 != 0) {stages = 0;}
  p.sm_a = mppi_smem_layout_a(h->T, need ? p.path_cap : 0, obstacles, ring, p.block_a, win_bytes, axes, stages);
  p.sm_tensors = mppi_smem_layout_a(h->T, need ? p.path_cap : 0, obstacles, 0, p.block_a, win_bytes, axes, 0);
  p.block_b = MPPI_BLOCK_B;
  // kernel instantiation: union of the active critics' types over the robots of this call
  uint32_t types = 0;
  for (int r = 0; r < h->R; ++r) {
    const DevCycle & cy = *h->hcyc(r);
    if (!cy.run) {continue;}
    for (uint32_t k = 0; k < h->cfg.n_critics; ++k) {
      if ((cy.active_mask >> k) & 1u) {types |= 1u << h->cfg.critics[k].type;}
    }
  }
  p.ka.block = p.block_a;
  p.ka.holo = h->holo;
  p.ka.from_tensors = false;
  p.ka.full = h->cfg.materialize_tensors || h->cfg.debug_cells || h->cfg.visualize || h->cfg.clamp_raw_controls || lag > 0;
  p.ka.crit_types = types;
  const int kc = h->critic_of_type[MPPI_CRITIC_COST], kp = h->critic_of_type[MPPI_CRITIC_PATH_ALIGN];
  p.ka.cc_step = kc >= 0 ? static_cast<int>(h->cfg.critics[kc].trajectory_point_step) : 2;
  p.ka.pa_step = kp >= 0 ? static_cast<int>(h->cfg.critics[kp].trajectory_point_step) : 4;
  p.ka.model = h->cfg.motion_model;
  {
    bool ok = !(kc >= 0 && h->cfg.critics[kc].consider_footprint) && !(kp >= 0 && h->cfg.critics[kp].use_path_orientations);
    for (int r = 0; r < h->R && ok; ++r) {
      const float res = static_cast<float>(h->robots[r].resolution);
      ok = res >= 0x1p-60f && res <= 0x1p60f;
      // the heading stays far inside mppi_sincosf's range over the whole horizon: |wz(t)| <= |wz(0)| + t * max_delta_wz
      // (MotionModel::predict clamps every step, whatever the noise), so |yaw(t)| <= |yaw(0)| + T dt (|wz(0)| + T max_delta_wz)
      const DevCycle & cy = *h->hcyc(r);
      const double horizon = static_cast<double>(h->T) * h->cfg.model_dt;
      const double bound = std::fabs(static_cast<double>(cy.pose_yaw)) +
        horizon * (std::fabs(static_cast<double>(cy.speed_wz)) + h->T * static_cast<double>(h->h_static.max_delta_wz));
      ok = ok && bound < 1.0e6;   // NaN fails the comparison
    }
    p.ka.lean_ok = ok;
  }
  // fused path: the whole batch of a robot fits one thread-block cluster and nothing asks for the
  // full-feature kernels (tensor materialisation, debug dumps, visualisation, clamp_raw, input delay)
  p.fused = false;
  // (`visualize` alone is fine: the fused kernel writes per-critic costs, collision flags and the visualizer's
  // strided trajectory samples itself)
  const bool fused_excluded = h->cfg.materialize_tensors || h->cfg.debug_cells || h->cfg.clamp_raw_controls || lag > 0;
  if (h->allow_fused && h->cfg.shard_world == 1 && !fused_excluded && h->B <= MPPI_FUSED_G * MPPI_FUSED_MAX_CLUSTER) {
    int cs = 1;
    while (cs * MPPI_FUSED_G < h->B) {cs *= 2;}
    // The fused kernel trades throughput for latency (one 512-thread CTA per SM, serial phases): it is
    // the right tool while all clusters fit one wave of the 148 SMs.  Larger fleets are throughput-bound
    // and go through the general thread-per-trajectory kernels with a (blocks, robots) grid.
    const bool one_wave = static_cast<long long>(cs) * h->R <= 148;
    p.cluster_size = cs;
    // the fused kernel copies the path arrays 16 bytes at a time: capacity rounded up to 16 points (max_path is a
    // power of two >= 256, the cycle block is zero-initialised, entries past n_path are never read as path points)
    const int fused_cap = std::min(h->max_path, (p.path_cap + 15) & ~15);
    p.sm_f = mppi_smem_layout_f(h->T, h->holo, fused_cap, static_cast<int>(h->cfg.n_critics) + 3, obstacles, win_bytes, h->n_pa);
    if (p.sm_f.total <= 227u * 1024u) {
      const uint64_t key = (static_cast<uint64_t>(p.sm_f.total) << 8) | static_cast<uint64_t>(cs);
      if (h->fused_checked_key != key) {
        h->fused_checked_key = key;
        h->fused_checked_ok = mppi_fused_supported(p.sm_f, h->holo, cs);
      }
      p.fused = h->fused_checked_ok && one_wave;
    }
  }
  return p;
}

// Where the first iteration of a cycle reads its inputs from (KArgs::cyc_in / u_in / hist_in / path).
enum class CycleSource
{
  LIVE,      // the device cycle block, uploaded ahead of the kernels
  PRISTINE   // the saved copies of a resident replay: every replay starts from the same state without restore copies
};

int enqueue_iterations(
  MppiHandle * h, cudaStream_t stream, const LaunchPlan & plan, bool fresh_costs = false, CycleSource source = CycleSource::LIVE,
  bool window_only = false, bool zero_copy = false)
{
  KArgs a = make_args(h);
  a.map_resident = window_only ? 0 : 1;
  a.stamp_results = (h->poll_results && source == CycleSource::LIVE) ? 1 : 0;
  if (zero_copy) {
    // the kernel's first phase pulls the cycle block, the path arrays and the packed window rows over PCIe from the
    // pinned host arena (same layout as the device arena; pinned memory is device-accessible at its own address)
    a.cyc_in = reinterpret_cast<const DevCycle *>(h->h_cycle_block);
    a.path = reinterpret_cast<const float *>(h->h_cycle_block + sizeof(DevCycle) * h->R);
    a.path_valid = h->h_cycle_block + sizeof(DevCycle) * h->R + sizeof(float) * 4 * h->max_path * h->R;
    a.win_packets = h->h_packets;
  }
  const KArgs live = a;
  const int iters = static_cast<int>(h->cfg.iteration_count);
  // the path arrays never change within a cycle: they are read from the source block by every iteration
  if (source == CycleSource::PRISTINE) {
    const uint8_t * block = h->d_cycle_saved;
    a.path = reinterpret_cast<const float *>(block + sizeof(DevCycle) * h->R);
    a.path_valid = block + sizeof(DevCycle) * h->R + sizeof(float) * 4 * h->max_path * h->R;
  }
  for (int it = 0; it < iters; ++it) {
    a.last_iteration = (it == iters - 1) ? 1 : 0;
    a.zero_costs = (fresh_costs && it == 0) ? 1 : 0;
    if (it == 0 && source == CycleSource::PRISTINE) {
      a.cyc_in = reinterpret_cast<const DevCycle *>(h->d_cycle_saved);
      a.u_in = h->d_u_saved;
      a.hist_in = h->d_hist_saved;
    } else {
      a.cyc_in = live.cyc_in; a.u_in = live.u_in; a.hist_in = live.hist_in;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
      if (h->timing_used == h->timing_events.size()) {
        cudaEvent_t x, y;
        CU_CHECK(h, cudaEventCreate(&x));
        CU_CHECK(h, cudaEventCreate(&y));
        h->timing_events.emplace_back(x, y);
      }
      e0 = h->timing_events[h->timing_used].first;
      e1 = h->timing_events[h->timing_used].second;
      h->timing_used++;
      CU_CHECK(h, cudaEventRecord(e0, stream));
    }
    if (plan.fused) {
      CU_CHECK(h, mppi_launch_fused(a, plan.sm_f, h->holo, plan.cluster_size, stream));
      if (h->timing) {CU_CHECK(h, cudaEventRecord(e1, stream));}
      h->launches += 1;
      continue;
    }
    CU_CHECK(h, mppi_launch_rollout_score(a, plan.sm_a, plan.ka, stream));
    if (h->timing) {CU_CHECK(h, cudaEventRecord(e1, stream));}
    CU_CHECK(h, mppi_launch_path_score(a, plan.path_cap, h->holo, false, plan.block_b, stream));
    cudaEvent_t c0 = nullptr, c1 = nullptr;
    if (h->timing) {
      if (h->timing_used_c == h->timing_events_c.size()) {
        cudaEvent_t x, y;
        CU_CHECK(h, cudaEventCreate(&x));
        CU_CHECK(h, cudaEventCreate(&y));
        h->timing_events_c.emplace_back(x, y);
      }
      c0 = h->timing_events_c[h->timing_used_c].first;
      c1 = h->timing_events_c[h->timing_used_c].second;
      h->timing_used_c++;
      CU_CHECK(h, cudaEventRecord(c0, stream));
    }
    CU_CHECK(h, mppi_launch_softmax_partial(a, h->holo, stream));
    if (h->timing) {CU_CHECK(h, cudaEventRecord(c1, stream));}
    CU_CHECK(h, mppi_launch_finalize(a, h->holo, a.xch_peers ? 3 : 0, stream));
    h->launches += 4;
  }
  return MPPI_OK;
}

void drop_graphs(MppiHandle * h)
{
  for (auto & g : h->graphs) {cudaGraphExecDestroy(g.exec);}
  h->graphs.clear();
}

// Everything that is baked into a captured cycle graph besides fixed device / pinned addresses.
uint64_t plan_key(const MppiHandle * h, const LaunchPlan & p)
{
  uint64_t k = 1469598103934665603ull;
  auto mix = [&](uint64_t v) {k = (k ^ v) * 1099511628211ull;};
  mix(p.fused); mix(p.cluster_size); mix(p.sm_f.total); mix(p.sm_f.path_cap);
  mix(p.sm_a.total); mix(p.sm_a.path_cap); mix(p.sm_a.noise_stages); mix(p.sm_a.off_win); mix(p.sm_a.off_ring);
  mix(p.block_a); mix(p.block_b); mix(p.path_cap);
  mix(p.ka.block); mix(p.ka.holo); mix(p.ka.full); mix(p.ka.crit_types); mix(p.ka.cc_step); mix(p.ka.pa_step);
  mix(p.ka.model); mix(p.ka.lean_ok); mix(p.sm_a.off_cc_lut);
  mix(h->cycle_block_bytes); mix(h->max_path); mix(h->map_stride); mix(h->cfg.visualize); mix(h->single_copy); mix(h->noise_flip); mix(h->poll_results);
  mix(h->peers_open); mix(reinterpret_cast<uint64_t>(h->d_peers));
  return k;
}

int upload_cycle_block(MppiHandle * h, cudaStream_t stream)
{
  CU_CHECK(h, cudaMemcpyAsync(h->d_cycle_block, h->h_cycle_block, h->cycle_block_bytes, cudaMemcpyHostToDevice, stream));
  h->h2d_bytes += h->cycle_block_bytes;
  return MPPI_OK;
}

// pinned staging -> device on the side stream for every robot whose device copy is behind its staging slab
int upload_stale_maps(MppiHandle * h)
{
  // neighbouring robots whose slabs are full (map_bytes == map_stride, the usual fleet of equal maps) go up with
  // one copy: 512 robots staged by mppi_optimize_in_place are one 20 MB transfer, not 512 enqueues
  bool any = false;
  int r = 0;
  while (r < h->R) {
    HostRobot & first = h->robots[r];
    if (!first.device_stale || !first.map_valid) {++r; continue;}
    int r1 = r + 1;
    size_t bytes = first.map_bytes;
    if (first.map_bytes == h->map_stride) {
      while (r1 < h->R && h->robots[r1].device_stale && h->robots[r1].map_valid && h->robots[r1].map_bytes == h->map_stride) {++r1;}
      bytes = h->map_stride * static_cast<size_t>(r1 - r);
    }
    CU_CHECK(h, cudaMemcpyAsync(
        h->d_map + h->map_stride * r, h->h_map_stage + h->map_stride * r, bytes, cudaMemcpyHostToDevice, h->side));
    h->h2d_bytes += bytes;
    // one event per copy, owned by the first robot of the group: a robot's staging slab may be overwritten again
    // once the event of the copy that last read it has completed
    if (!first.stage_event) {CU_CHECK(h, cudaEventCreateWithFlags(&first.stage_event, cudaEventDisableTiming));}
    CU_CHECK(h, cudaEventRecord(first.stage_event, h->side));
    for (int q = r; q < r1; ++q) {
      h->robots[q].stage_pending = true;
      h->robots[q].stage_owner = r;
      h->robots[q].device_stale = false;
    }
    any = true;
    r = r1;
  }
  if (any) {
    CU_CHECK(h, cudaEventRecord(h->map_event, h->side));
    h->map_event_pending = true;
  }
  return MPPI_OK;
}

// Single-launch cycles: everything the kernel needs from the host in as few copies as possible on `stream`.
// With one robot the stale costmap slab and the cycle block are adjacent in the arena: one copy.
// Window-only cycles: the rows of each robot's costmap window, packed behind the cycle block (one copy for both).
int fill_window_packets(MppiHandle * h)
{
  size_t off = 0;
  for (int r = 0; r < h->R; ++r) {
    DevCycle & cy = *h->hcyc(r);
    cy.win_packet_off = static_cast<uint32_t>(off);
    if (!cy.run || cy.win_w <= 0 || cy.win_h <= 0) {continue;}
    const size_t bytes = static_cast<size_t>(cy.win_w) * cy.win_h;
    if (off + bytes > h->packets_capacity) {h->error = "window packets overflow"; return MPPI_ERR_UNSUPPORTED;}
    off += bytes;
  }
  h->packets_used = off;
  // the rows of every robot's window, packed; robots are independent (a fleet's packing is shared out over the host cores)
#pragma omp parallel for schedule(static) num_threads(mppi_host_threads()) if (h->R >= MPPI_HOST_PARALLEL_ROBOTS)
  for (int r = 0; r < h->R; ++r) {
    const HostRobot & rb = h->robots[r];
    const DevCycle & cy = *h->hcyc(r);
    if (!cy.run || cy.win_w <= 0 || cy.win_h <= 0) {continue;}
    uint8_t * dst = h->h_packets + cy.win_packet_off;
    // the window's right edge is rounded up to 16 cells and may run past size_x into the row padding
    const int copy_w = std::min<int>(cy.win_w, static_cast<int>(rb.size_x) - cy.win_x0);
    for (int row = 0; row < cy.win_h; ++row) {
      const uint8_t * src = rb.map_src + static_cast<size_t>(cy.win_y0 + row) * rb.map_src_pitch + cy.win_x0;
      std::memcpy(dst + static_cast<size_t>(row) * cy.win_w, src, copy_w);
      if (copy_w < cy.win_w) {std::memset(dst + static_cast<size_t>(row) * cy.win_w + copy_w, 0, cy.win_w - copy_w);}
    }
  }
  return MPPI_OK;
}

int upload_cycle_and_packets(MppiHandle * h, cudaStream_t stream)
{
  const size_t block_off = h->map_stride * h->R;
  CU_CHECK(h, cudaMemcpyAsync(
      h->d_arena + block_off, h->h_arena + block_off, h->cycle_block_bytes + h->packets_used, cudaMemcpyHostToDevice, stream));
  h->h2d_bytes += h->cycle_block_bytes + h->packets_used;
  return MPPI_OK;
}

int upload_arena(MppiHandle * h, cudaStream_t stream)
{
  const size_t block_off = h->map_stride * h->R;
  if (h->R == 1) {
    HostRobot & rb = h->robots[0];
    const size_t from = (rb.device_stale && rb.map_valid) ? 0 : block_off;
    CU_CHECK(h, cudaMemcpyAsync(
        h->d_arena + from, h->h_arena + from, block_off + h->cycle_block_bytes - from, cudaMemcpyHostToDevice, stream));
    h->h2d_bytes += block_off + h->cycle_block_bytes - from;
    rb.device_stale = false;
    return MPPI_OK;
  }
  for (int r = 0; r < h->R; ++r) {
    HostRobot & rb = h->robots[r];
    if (!rb.device_stale || !rb.map_valid) {continue;}
    CU_CHECK(h, cudaMemcpyAsync(
        h->d_arena + h->map_stride * r, h->h_arena + h->map_stride * r, rb.map_bytes, cudaMemcpyHostToDevice, stream));
    h->h2d_bytes += rb.map_bytes;
    rb.device_stale = false;
  }
  CU_CHECK(h, cudaMemcpyAsync(h->d_arena + block_off, h->h_arena + block_off, h->cycle_block_bytes, cudaMemcpyHostToDevice, stream));
  h->h2d_bytes += h->cycle_block_bytes;
  return MPPI_OK;
}

int wait_map(MppiHandle * h, cudaStream_t stream)
{
  int rc = upload_stale_maps(h);
  if (rc != MPPI_OK) {return rc;}
  if (h->map_event_pending) {
    CU_CHECK(h, cudaStreamWaitEvent(stream, h->map_event, 0));
    h->map_event_pending = false;
  }
  return MPPI_OK;
}

// Robot r's result block as the kernels publish it (KArgs::out_tag): 8-byte words, value in the low half, the
// sequence number of the cycle that wrote it in the high half.
static inline const volatile uint64_t * tagged_block(const MppiHandle * h, int r)
{
  return reinterpret_cast<const volatile uint64_t *>(h->h_out_tag) + static_cast<size_t>(r) * (h->out_stride / 4);
}

// True once every word of robot r's result block carries `seq`: the header first (a window miss publishes nothing
// else), then the control sequence and the optimal trajectory.
static bool results_complete(const MppiHandle * h, int r, uint32_t seq)
{
  const volatile uint64_t * w = tagged_block(h, r);
  constexpr int HW = static_cast<int>(sizeof(DevOutHeader) / 4);
  for (int i = 0; i < HW; ++i) {
    if (static_cast<uint32_t>(w[i] >> 32) != seq) {return false;}
  }
  if (static_cast<uint32_t>(w[offsetof(DevOutHeader, map_miss) / 4]) != 0u) {return true;}
  const int n = HW + 6 * h->T;
  for (int i = HW; i < n; ++i) {
    if (static_cast<uint32_t>(w[i] >> 32) != seq) {return false;}
  }
  return true;
}

// Moves robot r's published words into the plain block (h_out: DevOutHeader, control sequence, optimal trajectory)
// that the rest of the host code reads.
static void unpack_results(MppiHandle * h, int r)
{
  const volatile uint64_t * w = tagged_block(h, r);
  uint32_t * dst = reinterpret_cast<uint32_t *>(h->h_out + static_cast<size_t>(r) * h->out_stride);
  const size_t n = h->out_stride / 4;
  for (size_t i = 0; i < n; ++i) {dst[i] = static_cast<uint32_t>(w[i]);}
}

// fallback() (optimizer.cpp:281-298) + the tail of evalControl (:261-269) for one robot after an
// attempt.  Returns 1 if this robot must run another attempt.
int finish_attempt(MppiHandle * h, int r, MppiCycleOut & out, bool & retry)
{
  const uint8_t * o = h->h_out + static_cast<size_t>(r) * h->out_stride;
  const DevOutHeader * hdr = reinterpret_cast<const DevOutHeader *>(o);
  HostRobot & rb = h->robots[r];
  DevCycle & cy = *h->hcyc(r);
  const int T = h->T;
  out.fail_flag = hdr->fail_flag;
  out.validation = hdr->validation;
  out.furthest_reached_path_point = hdr->furthest;
  const bool fail = hdr->fail_flag != 0 || hdr->validation != MPPI_VALIDATION_SUCCESS;
  retry = false;
  if (fail) {
    int rc = reset_robot(h, r, false);
    if (rc != MPPI_OK) {return rc;}
    if (++rb.fallback_counter > h->cfg.retry_attempt_limit) {
      rb.fallback_counter = 0;
      out.status = MPPI_ERR_NO_VALID_CONTROL;  // "Optimizer fail to compute path"
      cy.run = 0;
      return MPPI_OK;
    }
    out.retries++;
    // fail_flag and furthest_reached_path_point stay as they are: prepare() is not called again
    cy.fail_flag = hdr->fail_flag;
    cy.furthest_cached = hdr->furthest;
    cy.run = 1;
    retry = true;
    return MPPI_OK;
  }
  rb.fallback_counter = 0;
  cy.run = 0;
  const float * u = reinterpret_cast<const float *>(o + sizeof(DevOutHeader));
  const float * traj = u + 3 * T;
  out.cmd_vx = hdr->cmd_vx;
  out.cmd_vy = hdr->cmd_vy;
  out.cmd_wz = hdr->cmd_wz;
  // getControlFromSequenceAsTwist: pushCommandHistory + last_command_vel_ (optimizer.cpp:657, :263)
  push_one(rb.hist_vx, hdr->cmd_vx);
  push_one(rb.hist_vy, hdr->cmd_vy);
  push_one(rb.hist_wz, hdr->cmd_wz);
  rb.last_cmd_vx = hdr->cmd_vx;
  rb.last_cmd_vy = h->holo ? hdr->cmd_vy : 0.0;
  rb.last_cmd_wz = hdr->cmd_wz;
  if (out.control_vx) {std::memcpy(out.control_vx, u, sizeof(float) * T);}
  if (out.control_vy) {std::memcpy(out.control_vy, u + T, sizeof(float) * T);}
  if (out.control_wz) {std::memcpy(out.control_wz, u + 2 * T, sizeof(float) * T);}
  if (out.optimal_trajectory) {std::memcpy(out.optimal_trajectory, traj, sizeof(float) * 3 * T);}
  out.status = MPPI_OK;
  return MPPI_OK;
}

int check_handle(const MppiHandle * h)
{
  return h ? MPPI_OK : MPPI_ERR_INVALID_ARGUMENT;
}

}  // namespace

extern "C" {

const char * mppi_last_error(const MppiHandle * h)
{
  return h ? h->error.c_str() : g_create_error.c_str();
}

int mppi_create(const MppiConfig * cfg, MppiHandle ** out)
{
  if (!cfg || !out) {g_create_error = "null argument"; return MPPI_ERR_INVALID_ARGUMENT;}
  *out = nullptr;
  auto * h = new MppiHandle();
  std::string err;
  int rc = validate_and_derive(h, *cfg, err);
  if (rc != MPPI_OK) {g_create_error = err; delete h; return rc;}
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    delete h;
    return MPPI_ERR_CUDA;
  }
  auto fail = [&](int code) {
      g_create_error = h->error;
      free_device(h);
      if (h->stream) {cudaStreamDestroy(h->stream);}
      if (h->side) {cudaStreamDestroy(h->side);}
      if (h->map_event) {cudaEventDestroy(h->map_event);}
      delete h;
      return code;
    };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
    cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
    cudaEventCreateWithFlags(&h->map_event, cudaEventDisableTiming) != cudaSuccess)
  {
    h->error = "stream/event creation failed";
    return fail(MPPI_ERR_CUDA);
  }
  h->own_stream = h->stream;
  {
    const char * e = std::getenv("MPPI_B200_NO_FUSED");
    h->allow_fused = !(e && e[0] == '1');
    const char * g = std::getenv("MPPI_B200_NO_GRAPH");
    h->use_graphs = !(g && g[0] == '1');
    const char * t = std::getenv("MPPI_B200_HOST_TIMERS");
    h->host_timers = t && t[0] == '1';
  }
  h->robots.assign(h->R, HostRobot());
  rc = allocate(h);
  if (rc != MPPI_OK) {return fail(rc);}
  rc = ensure_cycle_block(h, 1);
  if (rc != MPPI_OK) {return fail(rc);}
  for (int r = 0; r < h->R; ++r) {
    rc = reset_robot(h, r, true);
    if (rc != MPPI_OK) {return fail(rc);}
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
    h->error = "device synchronisation failed after reset";
    return fail(MPPI_ERR_CUDA);
  }
  *out = h;
  return MPPI_OK;
}

int mppi_destroy(MppiHandle * h)
{
  if (!h) {return MPPI_OK;}
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->side);
  if (h->host_timers && h->ht_n) {
    const double n = static_cast<double>(h->ht_n);
    std::fprintf(stderr, "[mppi host timers, us/cycle over %llu cycles] upload_costmap %.1f  prepare %.1f  plan %.1f  "
      "enqueue %.1f  sync %.1f  finish %.1f\n", static_cast<unsigned long long>(h->ht_n), 1e6 * h->ht[0] / n,
      1e6 * h->ht[1] / n, 1e6 * h->ht[2] / n, 1e6 * h->ht[3] / n, 1e6 * h->ht[4] / n, 1e6 * h->ht[5] / n);
  }
  drop_graphs(h);
  free_device(h);
  if (h->d_arena) {cudaFree(h->d_arena);}
  if (h->h_arena) {cudaFreeHost(h->h_arena);}
  for (auto & p : h->timing_events) {cudaEventDestroy(p.first); cudaEventDestroy(p.second);}
  for (auto & p : h->timing_events_c) {cudaEventDestroy(p.first); cudaEventDestroy(p.second);}
  for (auto & e : h->resident_events) {cudaEventDestroy(e);}
  for (size_t p = 0; p < h->peer_mailboxes.size(); ++p) {
    if (h->peer_mailboxes[p] && h->peer_mailboxes[p] != h->d_mailbox) {cudaIpcCloseMemHandle(h->peer_mailboxes[p]);}
  }
  if (h->d_mailbox) {cudaFree(h->d_mailbox);}
  if (h->d_peers) {cudaFree(h->d_peers);}
  if (h->d_xstate) {cudaFree(h->d_xstate);}
  if (h->h_xerror) {cudaFreeHost(h->h_xerror);}
  if (h->d_flush) {cudaFree(h->d_flush);}
  delete h->resident_plan;
  for (auto & rb : h->robots) {if (rb.stage_event) {cudaEventDestroy(rb.stage_event);}}
  cudaEventDestroy(h->map_event);
  if (h->noise_event) {cudaEventDestroy(h->noise_event);}
  cudaStreamDestroy(h->own_stream);
  cudaStreamDestroy(h->side);
  delete h;
  return MPPI_OK;
}

int mppi_set_config(MppiHandle * h, const MppiConfig * cfg)
{
  if (!h || !cfg) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (cfg->n_robots != static_cast<uint32_t>(h->R)) {
    h->error = "n_robots cannot change on a live handle";
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  const MppiConfig old = h->cfg;
  if (h->d_mailbox &&
    (cfg->time_steps != old.time_steps || cfg->shard_world != old.shard_world || cfg->shard_rank != old.shard_rank))
  {
    // the mailbox layout (slot stride, peers' IPC mappings) was exchanged for the old shape: a peer writing a larger
    // payload at the old stride would overrun its neighbour's slot
    h->error = "time_steps / shard_world / shard_rank cannot change once the exchange mailbox exists: create a new handle";
    return MPPI_ERR_UNSUPPORTED;
  }
  std::string err;
  int rc = validate_and_derive(h, *cfg, err);
  if (rc != MPPI_OK) {
    h->error = err;
    std::string e2;
    validate_and_derive(h, old, e2);
    return rc;
  }
  for (auto & rb : h->robots) {
    if (old.batch_size != cfg->batch_size || old.time_steps != cfg->time_steps || old.shard_world != cfg->shard_world) {
      rb.noise_injected = false;  // shapes changed: injected tensors no longer fit
    }
  }
  rc = allocate(h);
  if (rc != MPPI_OK) {return rc;}
  h->have_cycle = false;
  // the ParametersHandler post-callback is Optimizer::reset() (optimizer.cpp:155)
  for (int r = 0; r < h->R; ++r) {
    if (h->robots[r].noise_injected) {h->robots[r].noise_injected = false;}
    rc = reset_robot(h, r, true);
    if (rc != MPPI_OK) {return rc;}
  }
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  return MPPI_OK;
}

int mppi_reset(MppiHandle * h, uint32_t robot, int reset_dynamic_speed_limits)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  for (int r = 0; r < h->R; ++r) {
    if (robot != UINT32_MAX && static_cast<int>(robot) != r) {continue;}
    int rc = reset_robot(h, r, reset_dynamic_speed_limits != 0);
    if (rc != MPPI_OK) {return rc;}
  }
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  return MPPI_OK;
}

int mppi_set_speed_limit(MppiHandle * h, double speed_limit, int percentage)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  // optimizer.cpp:691-719; NO_SPEED_LIMIT == 0.0 (nav2_costmap_2d/costmap_filters/filter_values.hpp)
  if (speed_limit == 0.0) {
    h->c_vx_max = h->base_vx_max; h->c_vx_min = h->base_vx_min; h->c_vy = h->base_vy; h->c_wz = h->base_wz;
  } else {
    const double ratio = percentage ? speed_limit / 100.0 : speed_limit / h->base_vx_max;
    h->c_vx_max = h->base_vx_max * ratio;
    h->c_vx_min = h->base_vx_min * ratio;
    h->c_vy = h->base_vy * ratio;
    h->c_wz = h->base_wz * ratio;
  }
  return MPPI_OK;
}

// Copies a costmap into the robot's pinned staging slab (16-byte-pitched rows) and makes it the host-side source.
int stage_map(
  MppiHandle * h, uint32_t robot, const uint8_t * cells, uint32_t size_x, uint32_t size_y,
  double resolution, double origin_x, double origin_y)
{
  const uint32_t pitch = (size_x + 15u) & ~15u;
  const size_t need = static_cast<size_t>(pitch) * size_y;
  if (need > h->map_stride) {
    int arc = ensure_arena(h, need, h->max_path);
    if (arc != MPPI_OK) {return arc;}
  }
  HostRobot & rb = h->robots[robot];
  // the previous async copy out of THIS robot's staging slab must have drained before it is overwritten
  if (rb.stage_pending) {
    CU_CHECK(h, cudaEventSynchronize(h->robots[rb.stage_owner].stage_event));
    rb.stage_pending = false;
  }
  uint8_t * stage = h->h_map_stage + h->map_stride * robot;
  if (pitch == size_x) {
    std::memcpy(stage, cells, need);
  } else {
    for (uint32_t y = 0; y < size_y; ++y) {
      std::memcpy(stage + static_cast<size_t>(y) * pitch, cells + static_cast<size_t>(y) * size_x, size_x);
    }
  }
  rb.map_host = stage;
  rb.device_only = false;
  rb.staged_current = true;
  rb.map_src = stage; rb.map_src_pitch = pitch;
  rb.size_x = size_x; rb.size_y = size_y; rb.pitch = pitch;
  rb.resolution = resolution; rb.origin_x = origin_x; rb.origin_y = origin_y;
  rb.map_valid = true;
  rb.map_bytes = need;
  rb.device_stale = true;
  return MPPI_OK;
}

int mppi_upload_costmap(
  MppiHandle * h, uint32_t robot, const uint8_t * cells, uint32_t size_x, uint32_t size_y,
  double resolution, double origin_x, double origin_y)
{
  if (!h || !cells || robot >= static_cast<uint32_t>(h->R) || size_x == 0 || size_y == 0 || !(resolution > 0.0)) {
    if (h) {h->error = "bad costmap arguments";}
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  const auto ht0 = std::chrono::steady_clock::now();
  int rc = stage_map(h, robot, cells, size_x, size_y, resolution, origin_x, origin_y);
  if (rc != MPPI_OK) {return rc;}
  // General path: pinned cudaMemcpyAsync on the side stream right away, overlapping the host-side preparation of
  // the cycle; consumers wait on map_event.  A handle whose cycles are single-launch (fused) defers it:
  // mppi_optimize then sends what the kernel needs with ONE copy on its own stream ahead of the ONE kernel —
  // the costmap window next to the cycle block, or the whole map next to it (two copies on two streams plus an
  // event wait cost ~15 us more per cycle, measured).
  if (!h->defer_map_upload) {
    rc = upload_stale_maps(h);
    if (rc != MPPI_OK) {return rc;}
  }
  if (h->host_timers) {h->ht[0] += std::chrono::duration<double>(std::chrono::steady_clock::now() - ht0).count();}
  return MPPI_OK;
}

// The costmap handed over where a device-side producer left it (the inflation pass, costmap_inflation_update_costs_device):
// one device-to-device 2-D copy into the robot's slab of the arena, ordered after the producer's stream by an event.
// No byte of the map crosses PCIe; the host keeps no copy, so path validity (utils::findPathCosts) is evaluated by the
// kernels (DevCycle::valid_on_device) and single-launch cycles read their window from the resident map.
int mppi_set_costmap_device(
  MppiHandle * h, uint32_t robot, const uint8_t * device_cells, uint32_t size_x, uint32_t size_y, size_t row_pitch,
  double resolution, double origin_x, double origin_y, void * producer_stream)
{
  if (!h || !device_cells || robot >= static_cast<uint32_t>(h->R) || size_x == 0 || size_y == 0 || row_pitch < size_x ||
    !(resolution > 0.0))
  {
    if (h) {h->error = "bad costmap arguments";}
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  const uint32_t pitch = (size_x + 15u) & ~15u;
  const size_t need = static_cast<size_t>(pitch) * size_y;
  if (need > h->map_stride) {
    int arc = ensure_arena(h, need, h->max_path);
    if (arc != MPPI_OK) {return arc;}
  }
  HostRobot & rb = h->robots[robot];
  if (producer_stream != nullptr && static_cast<cudaStream_t>(producer_stream) != h->stream) {
    if (!rb.stage_event) {CU_CHECK(h, cudaEventCreateWithFlags(&rb.stage_event, cudaEventDisableTiming));}
    CU_CHECK(h, cudaEventRecord(rb.stage_event, static_cast<cudaStream_t>(producer_stream)));
    CU_CHECK(h, cudaStreamWaitEvent(h->stream, rb.stage_event, 0));
  }
  uint8_t * dst = h->d_map + h->map_stride * robot;
  if (pitch != size_x) {CU_CHECK(h, cudaMemsetAsync(dst, 0, need, h->stream));}   // row padding reads as free space
  CU_CHECK(h, cudaMemcpy2DAsync(dst, pitch, device_cells, row_pitch, size_x, size_y, cudaMemcpyDeviceToDevice, h->stream));
  rb.size_x = size_x; rb.size_y = size_y; rb.pitch = pitch;
  rb.resolution = resolution; rb.origin_x = origin_x; rb.origin_y = origin_y;
  rb.map_bytes = need;
  rb.map_valid = true;
  rb.device_stale = false;      // the device copy IS the map
  rb.device_only = true;
  rb.staged_current = false;
  rb.map_src = nullptr; rb.map_src_pitch = 0; rb.map_host = nullptr;
  rb.stage_pending = false;
  return MPPI_OK;
}

int mppi_set_noise(
  MppiHandle * h, uint32_t robot, const float * noises_vx, const float * noises_vy, const float * noises_wz)
{
  if (!h || robot >= static_cast<uint32_t>(h->R) || !noises_vx || !noises_wz) {return MPPI_ERR_INVALID_ARGUMENT;}
  const size_t slab = static_cast<size_t>(h->B) * h->T;
  CU_CHECK(h, cudaMemcpyAsync(h->d_noise[0] + robot * slab, noises_vx, slab * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if (noises_vy) {
    CU_CHECK(h, cudaMemcpyAsync(h->d_noise[1] + robot * slab, noises_vy, slab * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  } else {
    CU_CHECK(h, cudaMemsetAsync(h->d_noise[1] + robot * slab, 0, slab * sizeof(float), h->stream));
  }
  CU_CHECK(h, cudaMemcpyAsync(h->d_noise[2] + robot * slab, noises_wz, slab * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  h->robots[robot].noise_injected = true;
  return MPPI_OK;
}

int mppi_generate_noise(MppiHandle * h, uint32_t robot)
{
  if (!h || robot >= static_cast<uint32_t>(h->R)) {return MPPI_ERR_INVALID_ARGUMENT;}
  int rc = generate_noise(h, static_cast<int>(robot));
  if (rc != MPPI_OK) {return rc;}
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  return MPPI_OK;
}

namespace
{
int optimize_impl(MppiHandle * h, const MppiCostmapView * views, const MppiCycleIn * in, MppiCycleOut * out);
}

int mppi_optimize(MppiHandle * h, const MppiCycleIn * in, MppiCycleOut * out)
{
  if (!h || !in || !out) {return MPPI_ERR_INVALID_ARGUMENT;}
  return optimize_impl(h, nullptr, in, out);
}

int mppi_optimize_in_place(MppiHandle * h, const MppiCostmapView * costmaps, const MppiCycleIn * in, MppiCycleOut * out)
{
  if (!h || !costmaps || !in || !out) {return MPPI_ERR_INVALID_ARGUMENT;}
  for (int r = 0; r < h->R; ++r) {
    const MppiCostmapView & v = costmaps[r];
    if (!v.cells || v.size_x == 0 || v.size_y == 0 || !(v.resolution > 0.0)) {
      h->error = "bad costmap arguments";
      return MPPI_ERR_INVALID_ARGUMENT;
    }
  }
  const int rc = optimize_impl(h, costmaps, in, out);
  // the caller's buffers are only borrowed for the duration of the call
  for (auto & rb : h->robots) {
    if (rb.map_src != rb.map_host) {rb.map_src = rb.map_host; rb.map_src_pitch = rb.pitch; rb.map_valid = rb.map_host != nullptr && rb.staged_current;}
  }
  return rc;
}

namespace
{
int optimize_impl(MppiHandle * h, const MppiCostmapView * views, const MppiCycleIn * in, MppiCycleOut * out)
{
  if (h->cfg.shard_world != 1 && !h->peers_open) {
    h->error = "mppi_optimize on a sharded handle without open peer mailboxes: use mppi_shard_open_peers + mppi_shard_cycle, "
      "or the mppi_shard_* phases with collectives in between";
    return MPPI_ERR_UNSUPPORTED;
  }
  int max_n = 1;
  for (int r = 0; r < h->R; ++r) {max_n = std::max<int>(max_n, in[r].n_path);}
  int rc = ensure_cycle_block(h, max_n);
  if (rc != MPPI_OK) {return rc;}
  bool prefetch_noise = false;
  if (h->cfg.regenerate_noises) {
    // noise thread analogue (noise_generator.cpp:98-106): fresh noise for this cycle.  Normally it was drawn on
    // the side stream after the previous cycle was launched and only has to be swapped in; injected noise
    // (mppi_set_noise) lives in one set only, so a handle with any injected robot draws in line instead.
    bool any_injected = false;
    for (int r = 0; r < h->R; ++r) {any_injected = any_injected || h->robots[r].noise_injected;}
    prefetch_noise = h->d_noise_next[0] != nullptr && !any_injected;
    if (prefetch_noise && h->noise_prefetched) {
      CU_CHECK(h, cudaStreamWaitEvent(h->stream, h->noise_event, 0));
      for (int i = 0; i < 3; ++i) {std::swap(h->d_noise[i], h->d_noise_next[i]);}
      h->noise_flip ^= 1u;
    } else {
      for (int r = 0; r < h->R; ++r) {
        if (!h->robots[r].noise_injected) {
          rc = generate_noise(h, r);
          if (rc != MPPI_OK) {return rc;}
        }
      }
    }
    h->noise_prefetched = false;
  }
  using clk = std::chrono::steady_clock;
  auto tick = [&](int slot, clk::time_point & t0) {
      if (h->host_timers) {
        const auto t1 = clk::now();
        h->ht[slot] += std::chrono::duration<double>(t1 - t0).count();
        t0 = t1;
      }
    };
  clk::time_point tp = clk::now();
  if (views) {
    // mppi_optimize_in_place: the costmaps are read where they lie (path validity, window packets); they are
    // staged and uploaded whole only if this cycle turns out to need more than the window
    for (int r = 0; r < h->R; ++r) {
      const MppiCostmapView & v = views[r];
      HostRobot & rb = h->robots[r];
      const uint32_t pitch = (v.size_x + 15u) & ~15u;
      const size_t need = static_cast<size_t>(pitch) * v.size_y;
      if (need > h->map_stride) {
        rc = ensure_arena(h, need, h->max_path);
        if (rc != MPPI_OK) {return rc;}
      }
      rb.size_x = v.size_x; rb.size_y = v.size_y; rb.pitch = pitch;
      rb.resolution = v.resolution; rb.origin_x = v.origin_x; rb.origin_y = v.origin_y;
      rb.map_bytes = need;
      rb.map_src = v.cells; rb.map_src_pitch = v.size_x;
      rb.map_valid = true;
      rb.device_stale = true;
      rb.device_only = false;
      rb.staged_current = false;
    }
  }
  {
    int prc = MPPI_OK;
#pragma omp parallel for schedule(static) num_threads(mppi_host_threads()) if (h->R >= MPPI_HOST_PARALLEL_ROBOTS)
    for (int r = 0; r < h->R; ++r) {
      out[r].retries = 0;
      out[r].status = MPPI_OK;
      const int one = prepare_robot(h, r, in[r], -1, false);   // writes only robot r's slots of the cycle block
      if (one != MPPI_OK) {
#pragma omp critical
        prc = one;
      }
    }
    if (prc != MPPI_OK) {return prc;}
  }
  tick(1, tp);
  const LaunchPlan plan = make_plan(h);
  tick(2, tp);
  // every consumer but the window-only launch needs the whole map on the device: stage what was only borrowed
  auto stage_borrowed = [&]() -> int {
      for (int r = 0; views && r < h->R; ++r) {
        if (h->robots[r].staged_current) {continue;}
        const MppiCostmapView & v = views[r];
        const int src = stage_map(h, static_cast<uint32_t>(r), v.cells, v.size_x, v.size_y, v.resolution, v.origin_x, v.origin_y);
        if (src != MPPI_OK) {return src;}
      }
      return MPPI_OK;
    };
  bool first = true;
  bool force_resident = false;  // a window-only attempt missed: same attempt again with the whole map
  std::vector<char> parked(h->R, 0);
  while (true) {
    // One attempt = H2D of the cycle block, the optimize() iterations, D2H of the result block.  The
    // common case (first attempt, own stream, no per-kernel timing) replays a captured CUDA graph:
    // one launch call instead of one per node.  Addresses are fixed; the plan signature keys the cache.
    h->cycle_seq++;
    for (int r = 0; r < h->R; ++r) {h->hcyc(r)->seq = h->cycle_seq;}
    const bool graphable = h->use_graphs && first && !h->resident_capture && !h->timing && h->stream == h->own_stream;
    const bool single_copy = graphable && plan.fused && h->single_copy;
    // window-only: ship the rows of the costmap the trajectories can plausibly reach instead of the map (the
    // kernel reports a lookup outside them and commits nothing; the attempt is then repeated with the whole map)
    bool any_device_only = false;
    for (int r = 0; r < h->R; ++r) {any_device_only = any_device_only || h->robots[r].device_only;}
    bool host_maps = true;   // every robot's cells are readable on the host (staging slab or the caller's buffer)
    for (int r = 0; r < h->R; ++r) {host_maps = host_maps && h->robots[r].map_src != nullptr;}
    // The general kernels take window-only uploads too (kernel A stages the packed rows, the finalize kernel's
    // validator reads them, a lookup outside raises the robot's RED_MAP_MISS word and nothing is committed): a
    // fleet cycle then moves ~9 KB per robot instead of staging and uploading 40 KB maps.
    const bool window_general = graphable && !plan.fused && h->window_only_general && views != nullptr;
    const bool window_only = (single_copy || window_general) && h->window_only && !force_resident &&
      h->cfg.iteration_count == 1 && h->packets_capacity > 0 && !any_device_only && host_maps && h->cfg.shard_world == 1;
    h->defer_map_upload = single_copy;
    const bool zero_copy = window_only && single_copy && h->zero_copy;
    if (window_only) {
      rc = fill_window_packets(h);
      if (rc == MPPI_OK && !zero_copy) {rc = upload_cycle_and_packets(h, h->stream);}
      if (zero_copy) {h->h2d_bytes += h->cycle_block_bytes + h->packets_used;}  // the same bytes cross PCIe, pulled by the kernel
    } else {
      rc = stage_borrowed();
      if (rc == MPPI_OK) {rc = single_copy ? upload_arena(h, h->stream) : wait_map(h, h->stream);}
    }
    if (rc != MPPI_OK) {return rc;}
    if (graphable) {
      const uint64_t key = plan_key(h, plan) ^ (window_only ? 0x9E3779B97F4A7C15ull : 0ull) ^ (single_copy ? 0x7F4A7C159E3779B9ull : 0ull) ^
        (zero_copy ? 0x3C6EF372FE94F82Bull : 0ull);
      cudaGraphExec_t exec = nullptr;
      for (auto & g : h->graphs) {
        if (g.key == key) {exec = g.exec; break;}
      }
      if (!exec) {
        const uint64_t launches_before = h->launches;
        CU_CHECK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        // single-launch plans: the cycle block already went up with the map (upload_arena); a copy node inside the
        // graph costs ~11 us of latency (measured), a stream-ordered copy ahead of a kernel-only graph ~5 us
        // The cycle block goes up with a stream-ordered copy AHEAD of a kernel-only graph (below), never as a copy node
        // inside it: a copy node costs ~11 us of latency, a stream-ordered copy ~5 us (measured, tools/experiments).
        int crc = MPPI_OK;
        if (crc == MPPI_OK && h->cfg.visualize && !plan.fused) {
          if (cudaMemsetAsync(h->d_collisions, 0, static_cast<size_t>(h->B) * h->R, h->stream) != cudaSuccess) {crc = MPPI_ERR_CUDA;}
          if (h->d_critic_costs &&
            cudaMemsetAsync(h->d_critic_costs, 0, sizeof(float) * h->cfg.n_critics * h->B * h->R, h->stream) != cudaSuccess)
          {
            crc = MPPI_ERR_CUDA;
          }
        }
        // the finalize kernel writes the result block straight into mapped pinned host memory: no D2H node
        if (crc == MPPI_OK) {
          crc = enqueue_iterations(h, h->stream, plan, true, CycleSource::LIVE, window_only, zero_copy);
        }
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        h->launches = launches_before;  // capture enqueued nothing; launches are counted per replay below
        if (crc != MPPI_OK || ce != cudaSuccess || !graph) {
          if (graph) {cudaGraphDestroy(graph);}
          const cudaError_t last = cudaGetLastError();
          // still correct (node-by-node launches, whole maps) but slower, and a launch that cannot be captured is a bug:
          // say so instead of degrading silently
          std::fprintf(stderr, "[mppi_b200] warning: capturing the cycle's graph failed (status %d, %s / %s): this handle "
                       "falls back to node-by-node launches\n", crc, cudaGetErrorString(ce), cudaGetErrorString(last));
          h->use_graphs = false;  // fall back to node-by-node launches for this handle
          continue;
        }
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
          cudaGetLastError();
          std::fprintf(stderr, "[mppi_b200] warning: instantiating the cycle's graph failed (%s): this handle falls back to "
                       "node-by-node launches\n", cudaGetErrorString(ce));
          h->use_graphs = false;
          continue;
        }
        if (h->graphs.size() >= 16) {drop_graphs(h);}
        h->graphs.push_back({key, exec});
      }
      if (!single_copy && !window_only) {
        rc = upload_cycle_block(h, h->stream);
        if (rc != MPPI_OK) {return rc;}
      }
      CU_CHECK(h, cudaGraphLaunch(exec, h->stream));
      h->launches += static_cast<uint64_t>(h->cfg.iteration_count) * (plan.fused ? 1 : 4);
      first = false;
    } else {
      rc = upload_cycle_block(h, h->stream);
      if (rc != MPPI_OK) {return rc;}
      if (first) {
        if (h->resident_capture) {
          if (!h->resident_plan) {h->resident_plan = new LaunchPlanBox();}
          h->resident_plan->plan = plan;
          // pristine copies for mppi_enqueue_resident
          CU_CHECK(h, cudaMemcpyAsync(h->d_cycle_saved, h->d_cycle_block, h->cycle_block_bytes, cudaMemcpyDeviceToDevice, h->stream));
          CU_CHECK(h, cudaMemcpyAsync(h->d_u_saved, h->d_u, sizeof(float) * 3 * h->T * h->R, cudaMemcpyDeviceToDevice, h->stream));
          CU_CHECK(h, cudaMemcpyAsync(h->d_hist_saved, h->d_hist, sizeof(float) * 12 * h->R, cudaMemcpyDeviceToDevice, h->stream));
        }
        CU_CHECK(h, cudaMemsetAsync(h->d_costs, 0, sizeof(float) * h->B * h->R, h->stream));
        first = false;
      }
      if (h->cfg.visualize && !plan.fused) {
        CU_CHECK(h, cudaMemsetAsync(h->d_collisions, 0, static_cast<size_t>(h->B) * h->R, h->stream));
        if (h->d_critic_costs) {
          CU_CHECK(h, cudaMemsetAsync(h->d_critic_costs, 0, sizeof(float) * h->cfg.n_critics * h->B * h->R, h->stream));
        }
      }
      rc = enqueue_iterations(h, h->stream, plan);
      if (rc != MPPI_OK) {return rc;}
    }
    if (prefetch_noise && !h->noise_prefetched) {
      // the other set was last read by the previous cycle, which has completed: draw the next cycle's noise into it
      for (int r = 0; r < h->R; ++r) {
        rc = generate_noise(h, r, true);
        if (rc != MPPI_OK) {return rc;}
      }
      CU_CHECK(h, cudaEventRecord(h->noise_event, h->side));
      h->noise_prefetched = true;
    }
    tick(3, tp);
    // Every word of a result block (mapped pinned memory) carries the sequence number of the cycle that wrote it:
    // polling for a complete set returns as soon as the last store has crossed PCIe, microseconds before the stream
    // would report the kernel retired, with no fence in the kernel.  Stream order still protects everything that
    // follows; a cycle that is not complete within 20 ms (an error, or a very large batch) falls back to waiting for
    // the stream.
    bool stamped = false;
    if (h->poll_results) {
      const auto p0 = std::chrono::steady_clock::now();
      for (unsigned spins = 0; !stamped; ++spins) {
        stamped = true;
        for (int r = 0; r < h->R && stamped; ++r) {
          if (!h->hcyc(r)->run) {continue;}
          stamped = results_complete(h, r, h->cycle_seq);
        }
        if (!stamped && (spins & 255u) == 255u &&
          std::chrono::steady_clock::now() - p0 > std::chrono::milliseconds(20)) {break;}
      }
    }
    if (!stamped) {CU_CHECK(h, cudaStreamSynchronize(h->stream));}
    for (int r = 0; r < h->R; ++r) {
      if (h->hcyc(r)->run) {unpack_results(h, r);}
    }
    tick(4, tp);
    // A window-only attempt may have missed for SOME robots of a fleet (a lookup outside the uploaded window: that
    // robot committed nothing).  The others are finished now — their results stand, they are not run again — and
    // the missed ones repeat the same attempt with the whole map.  Robots whose fallback() asks for another attempt
    // are parked until no fresh rerun is pending (a fresh rerun starts from zero costs, a retry accumulates,
    // optimizer.cpp:281-298 / :335).
    bool any_missed = false, any_retry = false;
    for (int r = 0; r < h->R; ++r) {
      if (!h->hcyc(r)->run) {continue;}  // finished in an earlier attempt, or parked
      const DevOutHeader * hdr = reinterpret_cast<const DevOutHeader *>(h->h_out + static_cast<size_t>(r) * h->out_stride);
      if (window_only && hdr->map_miss) {
        any_missed = true;
        if (std::getenv("MPPI_B200_DEBUG_MISS")) {std::fprintf(stderr, "[mppi] window miss: robot %d seq %u\n", r, h->cycle_seq);}
        continue;
      }
      bool retry = false;
      rc = finish_attempt(h, r, out[r], retry);
      if (rc != MPPI_OK) {return rc;}
      if (retry) {parked[r] = 1; h->hcyc(r)->run = 0;}
    }
    tick(5, tp);
    if (any_missed) {
      h->window_misses++;
      force_resident = true;
      first = true;   // nothing was committed for the robots still running: this is still their first attempt
      continue;
    }
    for (int r = 0; r < h->R; ++r) {
      if (parked[r]) {parked[r] = 0; h->hcyc(r)->run = 1; any_retry = true;}
    }
    if (!any_retry) {break;}
  }
  h->ht_n++;
  h->have_cycle = h->resident_capture;
  return MPPI_OK;
}
}  // namespace

uint64_t mppi_window_miss_count(const MppiHandle * h)
{
  return h ? h->window_misses : 0;
}

uint64_t mppi_h2d_byte_count(const MppiHandle * h)
{
  return h ? h->h2d_bytes : 0;
}

int mppi_set_resident_capture(MppiHandle * h, int enabled)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  h->resident_capture = enabled != 0;
  if (!h->resident_capture) {h->have_cycle = false;}
  return MPPI_OK;
}

// Times `cycles` back-to-back (mppi_upload_costmap; mppi_optimize) pairs for robot 0 .. n_robots-1 on the
// calling thread with steady_clock: the end-to-end latency of the C ABI with host buffers, free of any
// binding overhead.  per_call_seconds receives `cycles` entries.
int mppi_time_e2e(
  MppiHandle * h, const uint8_t * cells, uint32_t size_x, uint32_t size_y, double resolution, double origin_x,
  double origin_y, const MppiCycleIn * in, MppiCycleOut * out, int cycles, int in_place, double * per_call_seconds)
{
  if (!h || !cells || !in || !out || !per_call_seconds || cycles < 1) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (in_place) {
    std::vector<MppiCostmapView> views(h->R, MppiCostmapView{cells, size_x, size_y, resolution, origin_x, origin_y});
    for (int i = 0; i < cycles; ++i) {
      const auto t0 = std::chrono::steady_clock::now();
      int rc = mppi_optimize_in_place(h, views.data(), in, out);
      if (rc != MPPI_OK) {return rc;}
      per_call_seconds[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return MPPI_OK;
  }
  // MPPI_B200_E2E_SKIP_UPLOAD=1 (measurement aid): time mppi_optimize alone against the map already on the device
  const char * skip_env = std::getenv("MPPI_B200_E2E_SKIP_UPLOAD");
  const bool skip_upload = skip_env && skip_env[0] == '1';
  for (int i = 0; i < cycles; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < h->R && !(skip_upload && h->robots[r].map_valid); ++r) {
      int rc = mppi_upload_costmap(h, static_cast<uint32_t>(r), cells, size_x, size_y, resolution, origin_x, origin_y);
      if (rc != MPPI_OK) {return rc;}
    }
    int rc = mppi_optimize(h, in, out);
    if (rc != MPPI_OK) {return rc;}
    const auto t1 = std::chrono::steady_clock::now();
    per_call_seconds[i] = std::chrono::duration<double>(t1 - t0).count();
  }
  return MPPI_OK;
}

// The same with one costmap view per robot (fleets whose robots each own a map), always in place.
int mppi_time_e2e_views(
  MppiHandle * h, const MppiCostmapView * costmaps, const MppiCycleIn * in, MppiCycleOut * out, int cycles,
  double * per_call_seconds)
{
  if (!h || !costmaps || !in || !out || !per_call_seconds || cycles < 1) {return MPPI_ERR_INVALID_ARGUMENT;}
  for (int i = 0; i < cycles; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = mppi_optimize_in_place(h, costmaps, in, out);
    if (rc != MPPI_OK) {return rc;}
    per_call_seconds[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return MPPI_OK;
}

int mppi_score_tensors(
  MppiHandle * h, uint32_t robot, const MppiCycleIn * in,
  const float * vx, const float * vy, const float * wz,
  const float * x, const float * y, const float * yaw,
  int32_t furthest_reached_path_point,
  float * costs, uint8_t * collisions, int32_t * fail_flag, int32_t * furthest_out)
{
  if (!h || !in || !vx || !wz || !x || !y || !yaw || !costs || robot >= static_cast<uint32_t>(h->R)) {
    return MPPI_ERR_INVALID_ARGUMENT;
  }
  int rc = ensure_cycle_block(h, std::max<int>(1, in->n_path));
  if (rc != MPPI_OK) {return rc;}
  const size_t bt_all = static_cast<size_t>(h->R) * h->B * h->T;
  for (int i = 0; i < 6; ++i) {
    if (!h->d_state[i]) {
      CU_CHECK(h, cudaMalloc(&h->d_state[i], bt_all * sizeof(float)));
      CU_CHECK(h, cudaMemset(h->d_state[i], 0, bt_all * sizeof(float)));
    }
  }
  const size_t slab = static_cast<size_t>(h->B) * h->T;
  const float * src[6] = {vx, vy, wz, x, y, yaw};
  for (int i = 0; i < 6; ++i) {
    float * dst = h->d_state[i] + slab * robot;
    if (src[i]) {
      CU_CHECK(h, cudaMemcpyAsync(dst, src[i], slab * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    } else {
      CU_CHECK(h, cudaMemsetAsync(dst, 0, slab * sizeof(float), h->stream));
    }
  }
  // only `robot` runs: every other robot's cycle gets run = 0
  for (int r = 0; r < h->R; ++r) {
    if (static_cast<uint32_t>(r) == robot) {
      rc = prepare_robot(h, r, *in, furthest_reached_path_point, true);
      if (rc != MPPI_OK) {return rc;}
    } else {
      h->hcyc(r)->run = 0;
    }
  }
  const LaunchPlan plan = make_plan(h);
  rc = upload_cycle_block(h, h->stream);
  if (rc != MPPI_OK) {return rc;}
  CU_CHECK(h, cudaMemcpyAsync(h->d_costs + static_cast<size_t>(robot) * h->B, costs, sizeof(float) * h->B, cudaMemcpyHostToDevice, h->stream));
  CU_CHECK(h, cudaMemsetAsync(h->d_collisions + static_cast<size_t>(robot) * h->B, 0, h->B, h->stream));
  rc = wait_map(h, h->stream);
  if (rc != MPPI_OK) {return rc;}
  KArgs a = make_args(h);
  KPlanA ka = plan.ka;
  ka.from_tensors = true;
  CU_CHECK(h, mppi_launch_rollout_score(a, plan.sm_tensors, ka, h->stream));
  CU_CHECK(h, mppi_launch_path_score(a, plan.path_cap, h->holo, true, plan.block_b, h->stream));
  h->launches += 2;
  CU_CHECK(h, cudaMemcpyAsync(costs, h->d_costs + static_cast<size_t>(robot) * h->B, sizeof(float) * h->B, cudaMemcpyDeviceToHost, h->stream));
  if (collisions) {
    CU_CHECK(h, cudaMemcpyAsync(collisions, h->d_collisions + static_cast<size_t>(robot) * h->B, h->B, cudaMemcpyDeviceToHost, h->stream));
  }
  // reset the reduction words this call dirtied (the finalize kernel does it on the normal path)
  {
    int32_t init[RED_WORDS] = {0, INT32_MIN, 0, 0, INT32_MIN, 0, 0, 0};
    CU_CHECK(h, cudaStreamSynchronize(h->stream));
    CU_CHECK(h, cudaMemcpy(h->d_red + robot * RED_WORDS, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  unpack_results(h, static_cast<int>(robot));
  const DevOutHeader * hdr = reinterpret_cast<const DevOutHeader *>(h->h_out + h->out_stride * robot);
  if (fail_flag) {*fail_flag = hdr->fail_flag;}
  if (furthest_out) {*furthest_out = hdr->furthest;}
  h->have_cycle = false;
  return MPPI_OK;
}

int mppi_get_tensor(MppiHandle * h, uint32_t robot, int32_t which, void * dst, size_t dst_bytes)
{
  if (!h || !dst || robot >= static_cast<uint32_t>(h->R)) {return MPPI_ERR_INVALID_ARGUMENT;}
  const size_t B = h->B, T = h->T, slab = B * T;
  const void * src = nullptr;
  size_t bytes = 0;
  auto need = [&](const void * p, const char * what) -> int {
      if (!p) {h->error = std::string(what) + " is not kept (enable materialize_tensors / visualize / debug_cells)"; return MPPI_ERR_UNSUPPORTED;}
      return MPPI_OK;
    };
  int rc = MPPI_OK;
  switch (which) {
    case MPPI_TENSOR_VX: rc = need(h->d_state[0], "state.vx"); src = h->d_state[0] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_VY: rc = need(h->d_state[1], "state.vy"); src = h->d_state[1] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_WZ: rc = need(h->d_state[2], "state.wz"); src = h->d_state[2] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_X: rc = need(h->d_state[3], "trajectories.x"); src = h->d_state[3] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_Y: rc = need(h->d_state[4], "trajectories.y"); src = h->d_state[4] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_YAW: rc = need(h->d_state[5], "trajectories.yaws"); src = h->d_state[5] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_CVX: rc = need(h->d_cv[0], "state.cvx"); src = h->d_cv[0] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_CVY: rc = need(h->d_cv[1], "state.cvy"); src = h->d_cv[1] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_CWZ: rc = need(h->d_cv[2], "state.cwz"); src = h->d_cv[2] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_NOISE_VX: src = h->d_noise[0] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_NOISE_VY: src = h->d_noise[1] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_NOISE_WZ: src = h->d_noise[2] + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_COSTS: src = h->d_costs + B * robot; bytes = B * 4; break;
    case MPPI_TENSOR_SOFTMAX: src = h->d_softmax + B * robot; bytes = B * 4; break;
    case MPPI_TENSOR_CRITIC_COSTS:
      rc = need(h->d_critic_costs, "critic costs");
      src = h->d_critic_costs + static_cast<size_t>(h->cfg.n_critics) * B * robot; bytes = h->cfg.n_critics * B * 4; break;
    case MPPI_TENSOR_COLLISIONS: src = h->d_collisions + B * robot; bytes = B; break;
    case MPPI_TENSOR_COST_CELLS:
      rc = need(h->d_dbg_cost, "cost cells");
      src = h->d_dbg_cost + static_cast<size_t>(h->cc_ns) * B * robot; bytes = static_cast<size_t>(h->cc_ns) * B * 4; break;
    case MPPI_TENSOR_OBSTACLE_CELLS:
      rc = need(h->d_dbg_obs, "obstacle cells"); src = h->d_dbg_obs + slab * robot; bytes = slab * 4; break;
    case MPPI_TENSOR_VIS_XY:
      rc = need(h->d_vis_xy, "visualizer samples (visualize with vis_trajectory_step / vis_time_step > 0)");
      bytes = static_cast<size_t>(h->vis_n_traj) * h->vis_n_pts * 2 * sizeof(float);
      src = h->d_vis_xy + static_cast<size_t>(robot) * h->vis_n_traj * h->vis_n_pts * 2; break;
    case MPPI_TENSOR_PHASE_CLOCKS:
      rc = need(h->d_phase_clocks, "phase clocks (set MPPI_B200_PHASE_CLOCKS=1)");
      src = h->d_phase_clocks + static_cast<size_t>(robot) * MPPI_PHASE_CLOCK_WORDS;
      bytes = MPPI_PHASE_CLOCK_WORDS * sizeof(long long); break;
    default: h->error = "unknown tensor"; return MPPI_ERR_INVALID_ARGUMENT;
  }
  if (rc != MPPI_OK) {return rc;}
  if (bytes != dst_bytes) {h->error = "dst_bytes does not match the tensor size"; return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  CU_CHECK(h, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

int mppi_get_critics_evaluated(MppiHandle * h, uint32_t robot, uint32_t * n_evaluated)
{
  if (!h || robot >= static_cast<uint32_t>(h->R) || !n_evaluated) {return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  const DevOutHeader * hdr = reinterpret_cast<const DevOutHeader *>(h->h_out + static_cast<size_t>(robot) * h->out_stride);
  *n_evaluated = static_cast<uint32_t>(std::max(0, std::min<int>(hdr->critics_evaluated, static_cast<int>(h->cfg.n_critics))));
  return MPPI_OK;
}

int mppi_get_control_sequence(MppiHandle * h, uint32_t robot, float * vx, float * vy, float * wz)
{
  if (!h || robot >= static_cast<uint32_t>(h->R)) {return MPPI_ERR_INVALID_ARGUMENT;}
  const size_t T = h->T;
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  const float * u = h->d_u + robot * 3 * T;
  if (vx) {CU_CHECK(h, cudaMemcpy(vx, u, T * 4, cudaMemcpyDeviceToHost));}
  if (vy) {CU_CHECK(h, cudaMemcpy(vy, u + T, T * 4, cudaMemcpyDeviceToHost));}
  if (wz) {CU_CHECK(h, cudaMemcpy(wz, u + 2 * T, T * 4, cudaMemcpyDeviceToHost));}
  return MPPI_OK;
}

int mppi_set_control_sequence(MppiHandle * h, uint32_t robot, const float * vx, const float * vy, const float * wz)
{
  if (!h || robot >= static_cast<uint32_t>(h->R)) {return MPPI_ERR_INVALID_ARGUMENT;}
  const size_t T = h->T;
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  float * u = h->d_u + robot * 3 * T;
  if (vx) {CU_CHECK(h, cudaMemcpy(u, vx, T * 4, cudaMemcpyHostToDevice));}
  if (vy) {CU_CHECK(h, cudaMemcpy(u + T, vy, T * 4, cudaMemcpyHostToDevice));}
  if (wz) {CU_CHECK(h, cudaMemcpy(u + 2 * T, wz, T * 4, cudaMemcpyHostToDevice));}
  return MPPI_OK;
}

int mppi_get_state(MppiHandle * h, uint32_t robot, MppiOptimizerState * state)
{
  if (!h || !state || robot >= static_cast<uint32_t>(h->R)) {return MPPI_ERR_INVALID_ARGUMENT;}
  std::memset(state, 0, sizeof(*state));
  const HostRobot & rb = h->robots[robot];
  state->last_command_vel[0] = rb.last_cmd_vx;
  state->last_command_vel[1] = rb.last_cmd_vy;
  state->last_command_vel[2] = rb.last_cmd_wz;
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  CU_CHECK(h, cudaMemcpy(state->control_history, h->d_hist + robot * 12, 12 * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < rb.hist_vx.size(); ++i) {state->cmd_history_vx[i] = rb.hist_vx[i];}
  for (size_t i = 0; i < rb.hist_vy.size(); ++i) {state->cmd_history_vy[i] = rb.hist_vy[i];}
  for (size_t i = 0; i < rb.hist_wz.size(); ++i) {state->cmd_history_wz[i] = rb.hist_wz[i];}
  state->fallback_counter = static_cast<uint32_t>(rb.fallback_counter);
  return MPPI_OK;
}

int mppi_set_state(MppiHandle * h, uint32_t robot, const MppiOptimizerState * state)
{
  if (!h || !state || robot >= static_cast<uint32_t>(h->R)) {return MPPI_ERR_INVALID_ARGUMENT;}
  HostRobot & rb = h->robots[robot];
  rb.last_cmd_vx = state->last_command_vel[0];
  rb.last_cmd_vy = state->last_command_vel[1];
  rb.last_cmd_wz = state->last_command_vel[2];
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  CU_CHECK(h, cudaMemcpy(h->d_hist + robot * 12, state->control_history, 12 * sizeof(float), cudaMemcpyHostToDevice));
  for (size_t i = 0; i < rb.hist_vx.size(); ++i) {rb.hist_vx[i] = state->cmd_history_vx[i];}
  for (size_t i = 0; i < rb.hist_vy.size(); ++i) {rb.hist_vy[i] = state->cmd_history_vy[i];}
  for (size_t i = 0; i < rb.hist_wz.size(); ++i) {rb.hist_wz[i] = state->cmd_history_wz[i];}
  rb.fallback_counter = state->fallback_counter;
  return MPPI_OK;
}

int mppi_enqueue_resident(MppiHandle * h, void * cuda_stream)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (!h->have_cycle) {h->error = "mppi_enqueue_resident needs a preceding mppi_optimize"; return MPPI_ERR_NOT_READY;}
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  // Every replay does identical work: the first iteration reads the state the cycle started from (cycle block,
  // control sequence, filter history) from the pristine copies and starts from zero costs, so a replay is the
  // optimize() kernels and nothing else: no restore copies, no memset.
  int rc = wait_map(h, s);
  if (rc != MPPI_OK) {return rc;}
  if (!h->resident_plan) {h->error = "no captured cycle"; return MPPI_ERR_NOT_READY;}
  return enqueue_iterations(h, s, h->resident_plan->plan, true, CycleSource::PRISTINE);
}

// `steps` resident replays on the handle's stream, each bracketed by CUDA events, after `warmup` untimed ones.
// flush_bytes > 0: before every replay a buffer of flush_bytes is overwritten (cudaMemsetAsync), so the replay starts
// with a cold L2.  (Reading a second buffer afterwards, to leave no dirty lines behind, measured 1 us slower, not faster.)
int mppi_time_resident(MppiHandle * h, int steps, int warmup, size_t flush_bytes, float * ms_per_step)
{
  if (check_handle(h) != MPPI_OK || steps < 1 || warmup < 0 || !ms_per_step) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (!h->have_cycle) {h->error = "mppi_time_resident needs resident capture and a preceding mppi_optimize"; return MPPI_ERR_NOT_READY;}
  if (flush_bytes > 0 && h->flush_bytes < flush_bytes) {
    if (h->d_flush) {cudaFree(h->d_flush); h->d_flush = nullptr;}
    CU_CHECK(h, cudaMalloc(&h->d_flush, flush_bytes));
    CU_CHECK(h, cudaMemset(h->d_flush, 1, flush_bytes));
    h->flush_bytes = flush_bytes;
  }
  while (h->resident_events.size() < static_cast<size_t>(2 * steps)) {
    cudaEvent_t e;
    CU_CHECK(h, cudaEventCreate(&e));
    h->resident_events.push_back(e);
  }
  cudaStream_t s = h->stream;
  for (int i = -warmup; i < steps; ++i) {
    if (flush_bytes > 0) {
      CU_CHECK(h, cudaMemsetAsync(h->d_flush, 0, flush_bytes, s));
    }
    if (i >= 0) {CU_CHECK(h, cudaEventRecord(h->resident_events[2 * i], s));}
    int rc = mppi_enqueue_resident(h, nullptr);
    if (rc != MPPI_OK) {return rc;}
    if (i >= 0) {CU_CHECK(h, cudaEventRecord(h->resident_events[2 * i + 1], s));}
  }
  CU_CHECK(h, cudaStreamSynchronize(s));
  for (int i = 0; i < steps; ++i) {
    CU_CHECK(h, cudaEventElapsedTime(&ms_per_step[i], h->resident_events[2 * i], h->resident_events[2 * i + 1]));
  }
  return MPPI_OK;
}

int mppi_enqueue_score_resident(MppiHandle * h, void * cuda_stream)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (!h->have_cycle || !h->d_state[0]) {
    h->error = "mppi_enqueue_score_resident needs materialize_tensors, resident capture and a preceding mppi_optimize";
    return MPPI_ERR_NOT_READY;
  }
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
  if (int wrc = wait_map(h, s); wrc != MPPI_OK) {return wrc;}
  CU_CHECK(h, cudaMemcpyAsync(h->d_cycle_block, h->d_cycle_saved, h->cycle_block_bytes, cudaMemcpyDeviceToDevice, s));
  CU_CHECK(h, cudaMemsetAsync(h->d_costs, 0, sizeof(float) * h->B * h->R, s));
  const LaunchPlan plan = h->resident_plan ? h->resident_plan->plan : make_plan(h);
  KArgs a = make_args(h);
  a.materialize = 0;
  KPlanA ka = plan.ka;
  ka.from_tensors = true;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing) {
    if (h->timing_used == h->timing_events.size()) {
      cudaEvent_t x, y;
      CU_CHECK(h, cudaEventCreate(&x));
      CU_CHECK(h, cudaEventCreate(&y));
      h->timing_events.emplace_back(x, y);
    }
    e0 = h->timing_events[h->timing_used].first;
    e1 = h->timing_events[h->timing_used].second;
    h->timing_used++;
    CU_CHECK(h, cudaEventRecord(e0, s));
  }
  CU_CHECK(h, mppi_launch_rollout_score(a, plan.sm_tensors, ka, s));
  if (h->timing) {CU_CHECK(h, cudaEventRecord(e1, s));}
  CU_CHECK(h, mppi_launch_path_score(a, plan.path_cap, h->holo, true, plan.block_b, s));
  // the reduction words this pass dirtied are reset by the next mppi_optimize's finalize kernel; do it here too
  static const int32_t init[RED_WORDS] = {0, INT32_MIN, 0, 0, INT32_MIN, 0, 0, 0};
  for (int r = 0; r < h->R; ++r) {
    CU_CHECK(h, cudaMemcpyAsync(h->d_red + r * RED_WORDS, init, sizeof(init), cudaMemcpyHostToDevice, s));
  }
  h->launches += 2;
  return MPPI_OK;
}

uint64_t mppi_launch_count(const MppiHandle * h) {return h ? h->launches : 0;}

int mppi_set_kernel_timing(MppiHandle * h, int enabled)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  h->timing = enabled != 0;
  h->timing_used = 0;
  h->timing_used_c = 0;
  return MPPI_OK;
}

int mppi_kernel_time_ms(MppiHandle * h, int reset, double * ms_per_launch, uint64_t * launches)
{
  if (!h || !ms_per_launch) {return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaDeviceSynchronize());
  double total = 0.0;
  for (size_t i = 0; i < h->timing_used; ++i) {
    float ms = 0.0f;
    CU_CHECK(h, cudaEventElapsedTime(&ms, h->timing_events[i].first, h->timing_events[i].second));
    total += ms;
  }
  *ms_per_launch = h->timing_used ? total / static_cast<double>(h->timing_used) : 0.0;
  if (launches) {*launches = h->timing_used;}
  if (reset) {h->timing_used = 0;}
  return MPPI_OK;
}

int mppi_kernel_c_time_ms(MppiHandle * h, int reset, double * ms_per_launch, uint64_t * launches)
{
  if (!h || !ms_per_launch) {return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaDeviceSynchronize());
  double total = 0.0;
  for (size_t i = 0; i < h->timing_used_c; ++i) {
    float ms = 0.0f;
    CU_CHECK(h, cudaEventElapsedTime(&ms, h->timing_events_c[i].first, h->timing_events_c[i].second));
    total += ms;
  }
  *ms_per_launch = h->timing_used_c ? total / static_cast<double>(h->timing_used_c) : 0.0;
  if (launches) {*launches = h->timing_used_c;}
  if (reset) {h->timing_used_c = 0;}
  return MPPI_OK;
}

void * mppi_stream(MppiHandle * h) {return h ? static_cast<void *>(h->stream) : nullptr;}

int mppi_set_stream(MppiHandle * h, void * cuda_stream)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return MPPI_OK;
}

int mppi_device_synchronize(MppiHandle * h)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  CU_CHECK(h, cudaStreamSynchronize(h->side));
  return MPPI_OK;
}

int mppi_selftest_exact_math(const float * a, const float * b, int n, unsigned int * mismatches, unsigned int * flagged)
{
  if (!a || !b || n < 1 || !mismatches || !flagged) {return MPPI_ERR_INVALID_ARGUMENT;}
  float * d_a = nullptr, * d_b = nullptr;
  unsigned int * d_cnt = nullptr;
  unsigned int cnt[2] = {0u, 0u};
  const size_t bytes = sizeof(float) * static_cast<size_t>(n);
  cudaError_t e = cudaMalloc(&d_a, bytes);
  if (e == cudaSuccess) {e = cudaMalloc(&d_b, bytes);}
  if (e == cudaSuccess) {e = cudaMalloc(&d_cnt, sizeof(cnt));}
  if (e == cudaSuccess) {e = cudaMemcpy(d_a, a, bytes, cudaMemcpyHostToDevice);}
  if (e == cudaSuccess) {e = cudaMemcpy(d_b, b, bytes, cudaMemcpyHostToDevice);}
  if (e == cudaSuccess) {e = cudaMemset(d_cnt, 0, sizeof(cnt));}
  if (e == cudaSuccess) {e = mppi_launch_selftest_exact_math(d_a, d_b, n, d_cnt, d_cnt + 1, nullptr);}
  if (e == cudaSuccess) {e = cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost);}
  cudaFree(d_a); cudaFree(d_b); cudaFree(d_cnt);
  if (e != cudaSuccess) {
    g_create_error = std::string("mppi_selftest_exact_math: ") + cudaGetErrorString(e);
    return MPPI_ERR_CUDA;
  }
  *mismatches = cnt[0];
  *flagged = cnt[1];
  return MPPI_OK;
}

// ---------------------------------------------------------------- batch sharding phases

int mppi_shard_buffers(
  MppiHandle * h, void ** ar1, size_t * ar1_words, void ** ar2_slot, size_t * ar2_words, void ** ar2_all)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (ar1) {*ar1 = h->d_red;}
  if (ar1_words) {*ar1_words = static_cast<size_t>(RED_WORDS) * h->R;}
  if (ar2_slot) {*ar2_slot = h->d_slot;}
  if (ar2_words) {*ar2_words = static_cast<size_t>(2 + 3 * h->T) * h->R;}
  if (ar2_all) {*ar2_all = h->d_slot_all;}
  return MPPI_OK;
}

namespace
{
size_t align256(size_t v) {return (v + 255) & ~static_cast<size_t>(255);}

int ensure_mailbox(MppiHandle * h)
{
  if (h->d_mailbox) {return MPPI_OK;}
  const size_t world = h->cfg.shard_world;
  h->mailbox_words_stride = align256(sizeof(int32_t) * RED_WORDS * h->R);
  h->mailbox_slots_stride = align256(sizeof(float) * (2 + 3 * h->T) * h->R);
  h->mailbox_words_base = align256(sizeof(uint32_t) * 2 * 2 * world);
  h->mailbox_slots_base = h->mailbox_words_base + 2 * world * h->mailbox_words_stride;
  h->mailbox_bytes = h->mailbox_slots_base + 2 * world * h->mailbox_slots_stride;
  CU_CHECK(h, cudaMalloc(&h->d_mailbox, h->mailbox_bytes));
  CU_CHECK(h, cudaMemset(h->d_mailbox, 0, h->mailbox_bytes));
  CU_CHECK(h, cudaMalloc(&h->d_peers, sizeof(unsigned char *) * world));
  CU_CHECK(h, cudaMalloc(&h->d_xstate, sizeof(uint32_t) * 4));
  CU_CHECK(h, cudaMemset(h->d_xstate, 0, sizeof(uint32_t) * 4));
  // the timeout flag lives in mapped pinned memory: the host reads it after the cycle without another copy
  CU_CHECK(h, cudaHostAlloc(&h->h_xerror, sizeof(int32_t), cudaHostAllocMapped));
  *h->h_xerror = 0;
  CU_CHECK(h, cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_xerror), h->h_xerror, 0));
  return MPPI_OK;
}

}  // namespace

int mppi_shard_ipc_handle(MppiHandle * h, unsigned char * handle64)
{
  if (check_handle(h) != MPPI_OK || !handle64) {return MPPI_ERR_INVALID_ARGUMENT;}
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  int rc = ensure_mailbox(h);
  if (rc != MPPI_OK) {return rc;}
  cudaIpcMemHandle_t ipc;
  CU_CHECK(h, cudaIpcGetMemHandle(&ipc, h->d_mailbox));
  std::memcpy(handle64, &ipc, sizeof(ipc));
  return MPPI_OK;
}

int mppi_shard_open_peers(MppiHandle * h, const unsigned char * handles)
{
  if (check_handle(h) != MPPI_OK || !handles) {return MPPI_ERR_INVALID_ARGUMENT;}
  int rc = ensure_mailbox(h);
  if (rc != MPPI_OK) {return rc;}
  const uint32_t world = h->cfg.shard_world, rank = h->cfg.shard_rank;
  h->peer_mailboxes.assign(world, nullptr);
  for (uint32_t p = 0; p < world; ++p) {
    if (p == rank) {h->peer_mailboxes[p] = h->d_mailbox; continue;}
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, handles + 64 * static_cast<size_t>(p), sizeof(ipc));
    CU_CHECK(h, cudaIpcOpenMemHandle(&h->peer_mailboxes[p], ipc, cudaIpcMemLazyEnablePeerAccess));
  }
  CU_CHECK(h, cudaMemcpy(h->d_peers, h->peer_mailboxes.data(), sizeof(void *) * world, cudaMemcpyHostToDevice));
  h->peers_open = true;
  drop_graphs(h);
  return MPPI_OK;
}

int mppi_shard_begin(MppiHandle * h, const MppiCycleIn * in)
{
  if (!h || !in) {return MPPI_ERR_INVALID_ARGUMENT;}
  int max_n = 1;
  for (int r = 0; r < h->R; ++r) {max_n = std::max<int>(max_n, in[r].n_path);}
  int rc = ensure_cycle_block(h, max_n);
  if (rc != MPPI_OK) {return rc;}
  for (int r = 0; r < h->R; ++r) {
    rc = prepare_robot(h, r, in[r], -1, false);
    if (rc != MPPI_OK) {return rc;}
  }
  rc = upload_cycle_block(h, h->stream);
  if (rc != MPPI_OK) {return rc;}
  h->shard_iteration = 0;  // the first iteration starts from zero costs (KArgs::zero_costs, no memset)
  return wait_map(h, h->stream);
}

int mppi_shard_rollout(MppiHandle * h)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  const LaunchPlan plan = make_plan(h);
  KArgs a = make_args(h);
  a.xch_peers = nullptr;  // the caller runs the collectives between the phases
  CU_CHECK(h, mppi_launch_rollout_score(a, plan.sm_a, plan.ka, h->stream));
  h->launches += 1;
  return MPPI_OK;
}

int mppi_shard_score(MppiHandle * h)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  const LaunchPlan plan = make_plan(h);
  KArgs a = make_args(h);
  a.xch_peers = nullptr;
  a.zero_costs = h->shard_iteration == 0 ? 1 : 0;
  CU_CHECK(h, mppi_launch_path_score(a, plan.path_cap, h->holo, false, plan.block_b, h->stream));
  CU_CHECK(h, mppi_launch_softmax_partial(a, h->holo, h->stream));
  CU_CHECK(h, mppi_launch_finalize(a, h->holo, 1, h->stream));
  h->launches += 3;
  return MPPI_OK;
}

int mppi_shard_update(MppiHandle * h, int last_iteration)
{
  if (check_handle(h) != MPPI_OK) {return MPPI_ERR_INVALID_ARGUMENT;}
  KArgs a = make_args(h);
  a.xch_peers = nullptr;
  a.last_iteration = last_iteration ? 1 : 0;
  CU_CHECK(h, mppi_launch_finalize(a, h->holo, 2, h->stream));
  h->launches += 1;
  h->shard_iteration++;
  return MPPI_OK;
}

// A whole sharded evalControl with the two exchanges done on the device over peer memory: every rank calls it with
// the same inputs.  It is mppi_optimize's own cycle (same host state machine, same captured graph of four kernels per
// iteration, result block polled in mapped memory); the exchanges are folded into kernels A / B / D (KArgs::xch_*).
int mppi_shard_cycle(MppiHandle * h, const MppiCycleIn * in, MppiCycleOut * out)
{
  if (!h || !in || !out) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (!h->peers_open && h->cfg.shard_world > 1) {h->error = "mppi_shard_cycle needs mppi_shard_open_peers first"; return MPPI_ERR_NOT_READY;}
  if (h->shard_failed) {
    h->error = "an earlier exchange timed out: the ranks are out of step, destroy and re-create the sharded handles";
    return MPPI_ERR_CUDA;
  }
  if (h->R != 1 && h->cfg.shard_world > 1) {h->error = "batch sharding drives one robot per handle"; return MPPI_ERR_UNSUPPORTED;}
  const int rc = optimize_impl(h, nullptr, in, out);
  if (h->h_xerror && *static_cast<volatile int32_t *>(h->h_xerror)) {
    // The kernels after a timed-out exchange ran on stale payloads and committed their result on this rank, and the
    // peers are one exchange out of step: nothing on this handle can be trusted any more.  It is marked dead (every
    // later cycle fails at once instead of waiting 2 s per exchange); the caller re-creates the handles on all ranks.
    *h->h_xerror = 0;
    h->shard_failed = true;
    h->error = "a peer did not reach an exchange within 2 s; the sharded handle is dead, re-create it on every rank";
    return MPPI_ERR_CUDA;
  }
  return rc;
}

// `cycles` back-to-back mppi_shard_cycle calls, each timed with steady_clock on the calling thread (all ranks call it)
int mppi_time_shard_cycles(MppiHandle * h, const MppiCycleIn * in, MppiCycleOut * out, int cycles, double * per_call_seconds)
{
  if (!h || !in || !out || !per_call_seconds || cycles < 1) {return MPPI_ERR_INVALID_ARGUMENT;}
  for (int i = 0; i < cycles; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = mppi_shard_cycle(h, in, out);
    if (rc != MPPI_OK) {return rc;}
    per_call_seconds[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return MPPI_OK;
}

int mppi_shard_end(MppiHandle * h, MppiCycleOut * out)
{
  if (!h || !out) {return MPPI_ERR_INVALID_ARGUMENT;}
  CU_CHECK(h, cudaStreamSynchronize(h->stream));
  int result = MPPI_OK;
  for (int r = 0; r < h->R; ++r) {
    bool retry = false;
    out[r].status = MPPI_OK;
    unpack_results(h, r);
    int rc = finish_attempt(h, r, out[r], retry);
    if (rc != MPPI_OK) {return rc;}
    if (retry) {result = 1;}  // caller re-uploads (mppi_shard_retry) and repeats the phases
  }
  if (result == 1) {
    int rc = upload_cycle_block(h, h->stream);
    if (rc != MPPI_OK) {return rc;}
  }
  return result;
}

}  // extern "C"
