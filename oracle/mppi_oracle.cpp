// =====================================================================================
// mppi_oracle.cpp — CPU restatement of nav2_mppi_controller's optimize() hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under navigation2_b200/ (the product) includes, links,
// or calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs load it, as the checker and as the timed CPU baseline.
//
// What it is: a from-scratch, single-threaded C++17 restatement (no Eigen, no ROS) of the
// reference algorithm, operation by operation, on the same column-major (t*B + b) float
// tensors.  Each function cites the reference file:line it follows; paths are relative to
// /root/reference/.  The reference itself cannot be compiled here (no ROS 2, Eigen, angles,
// tf2 — SURVEY.md §8c), so:
//
//   PARITY PINNING: this oracle is checked (tests/test_oracle_*.py) against every closed-form
//   expectation the reference's own tests hold for this path (critics_tests.cpp,
//   optimizer_unit_tests.cpp, motion_model_tests.cpp, utils_test.cpp,
//   footprint_collision_checker_test.cpp), and its Bresenham restatement is checked against the
//   reference's own nav2_util/line_iterator.hpp compiled from where it lies (oracle/_ref).
//   UNPINNED upstream (no reference test fixes them): CostCritic / ObstaclesCritic scores, the
//   softmax / updateControlSequence numeric result, the Savitzky-Golay filter's exact output and
//   the RNG stream.  For those this restatement IS the oracle; it says "parity unpinned" for them.
//
// Third-party arithmetic that is not under /root/reference (named + restated):
//   * Eigen3 (un-vendored, any 3.3/3.4): cos/sin/exp/sqrt packets, reductions, GEMV.  Restated
//     as sequential float loops; sin/cos through include/mppi_math.h (shared with the GPU so cell
//     indices can be compared bit-exactly); exp through libm expf.
//   * ros/angles (branch ros2): normalize_angle / shortest_angular_distance in double, restated in
//     include/mppi_math.h from the published header.
//   * tf2::getYaw / setRPY: callers pass yaw directly.
//   * libstdc++ minstd_rand0 + normal_distribution<float>: stream is unpinned; noise is injected.
//
// Build: oracle/Makefile (parity build -ffp-contract=off; timing build -O3 -mavx2 -mfma).
// =====================================================================================

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <optional>
#include <random>
#include <string>
#include <vector>

#include "mppi_b200.h"
#include "mppi_math.h"

namespace oracle
{

// ----- nav2_costmap_2d/include/nav2_costmap_2d/cost_values.hpp:70-74 -----
static constexpr unsigned char NO_INFORMATION = 255;
static constexpr unsigned char LETHAL_OBSTACLE = 254;
static constexpr unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;

// Column-major B x T float tensor, element (b, t) at t*B + b (Eigen::ArrayXXf).
struct Tensor
{
  int rows = 0, cols = 0;
  std::vector<float> d;
  void setZero(int r, int c) {rows = r; cols = c; d.assign(static_cast<size_t>(r) * c, 0.0f);}
  float & operator()(int b, int t) {return d[static_cast<size_t>(t) * rows + b];}
  float operator()(int b, int t) const {return d[static_cast<size_t>(t) * rows + b];}
  float * col(int t) {return d.data() + static_cast<size_t>(t) * rows;}
  const float * col(int t) const {return d.data() + static_cast<size_t>(t) * rows;}
};

// ----- nav2_costmap_2d::Costmap2D read side (src/costmap_2d.cpp:265-305) -----
struct Costmap
{
  std::vector<uint8_t> cells;
  unsigned int size_x = 0, size_y = 0;
  double resolution = 0.0, origin_x = 0.0, origin_y = 0.0;
  bool valid = false;

  unsigned char getCost(unsigned int mx, unsigned int my) const {return cells[my * size_x + mx];}
  unsigned char getCost(unsigned int idx) const {return cells[idx];}
  // costmap_2d.cpp:292-305
  bool worldToMap(double wx, double wy, unsigned int & mx, unsigned int & my) const
  {
    if (wx < origin_x || wy < origin_y) {
      return false;
    }
    mx = static_cast<unsigned int>((wx - origin_x) / resolution);
    my = static_cast<unsigned int>((wy - origin_y) / resolution);
    if (mx < size_x && my < size_y) {
      return true;
    }
    return false;
  }
};

// ----- nav2_util/include/nav2_util/line_iterator.hpp:56-124 (Bresenham) -----
struct LineIterator
{
  int x0_, y0_, x1_, y1_, x_, y_, deltax_, deltay_, curpixel_;
  int xinc1_, xinc2_, yinc1_, yinc2_, den_, num_, numadd_, numpixels_;
  LineIterator(int x0, int y0, int x1, int y1)
  : x0_(x0), y0_(y0), x1_(x1), y1_(y1), x_(x0), y_(y0),
    deltax_(std::abs(x1 - x0)), deltay_(std::abs(y1 - y0)), curpixel_(0)
  {
    if (x1_ >= x0_) {xinc1_ = 1; xinc2_ = 1;} else {xinc1_ = -1; xinc2_ = -1;}
    if (y1_ >= y0_) {yinc1_ = 1; yinc2_ = 1;} else {yinc1_ = -1; yinc2_ = -1;}
    if (deltax_ >= deltay_) {
      xinc1_ = 0; yinc2_ = 0;
      den_ = deltax_; num_ = deltax_ / 2; numadd_ = deltay_; numpixels_ = deltax_;
    } else {
      xinc2_ = 0; yinc1_ = 0;
      den_ = deltay_; num_ = deltay_ / 2; numadd_ = deltax_; numpixels_ = deltay_;
    }
  }
  bool isValid() const {return curpixel_ <= numpixels_;}
  void advance()
  {
    num_ += numadd_;
    if (num_ >= den_) {num_ -= den_; x_ += xinc1_; y_ += yinc1_;}
    x_ += xinc2_;
    y_ += yinc2_;
    curpixel_++;
  }
};

// ----- nav2_costmap_2d/src/footprint_collision_checker.cpp:48-144 -----
struct FootprintCollisionChecker
{
  const Costmap * costmap = nullptr;

  double pointCost(int x, int y) const  // :117-120
  {
    return static_cast<double>(costmap->getCost(x, y));
  }
  double lineCost(int x0, int x1, int y0, int y1) const  // :88-107
  {
    double line_cost = 0.0;
    double point_cost = -1.0;
    for (LineIterator line(x0, y0, x1, y1); line.isValid(); line.advance()) {
      point_cost = pointCost(line.x_, line.y_);
      if (point_cost == static_cast<double>(LETHAL_OBSTACLE)) {
        return point_cost;
      }
      if (line_cost < point_cost) {
        line_cost = point_cost;
      }
    }
    return line_cost;
  }
  double footprintCost(const std::vector<std::array<double, 2>> & footprint) const  // :48-86
  {
    unsigned int x0, x1 = 0, y0, y1 = 0;
    double footprint_cost = 0.0;
    if (!costmap->worldToMap(footprint[0][0], footprint[0][1], x0, y0)) {
      return static_cast<double>(LETHAL_OBSTACLE);
    }
    unsigned int xstart = x0, ystart = y0;
    for (unsigned int i = 0; i < footprint.size() - 1; ++i) {
      if (!costmap->worldToMap(footprint[i + 1][0], footprint[i + 1][1], x1, y1)) {
        return static_cast<double>(LETHAL_OBSTACLE);
      }
      footprint_cost = std::max(lineCost(x0, x1, y0, y1), footprint_cost);
      x0 = x1;
      y0 = y1;
      if (footprint_cost == static_cast<double>(LETHAL_OBSTACLE)) {
        return footprint_cost;
      }
    }
    return std::max(lineCost(xstart, x1, ystart, y1), footprint_cost);
  }
  double footprintCostAtPose(
    double x, double y, double theta, const std::vector<std::array<double, 2>> & footprint) const  // :128-144
  {
    double cos_th = cos(theta);
    double sin_th = sin(theta);
    std::vector<std::array<double, 2>> oriented;
    oriented.reserve(footprint.size());
    for (unsigned int i = 0; i < footprint.size(); ++i) {
      std::array<double, 2> p;
      p[0] = x + (footprint[i][0] * cos_th - footprint[i][1] * sin_th);
      p[1] = y + (footprint[i][0] * sin_th + footprint[i][1] * cos_th);
      oriented.push_back(p);
    }
    return footprintCost(oriented);
  }
};

// nav2_costmap_2d/include/nav2_costmap_2d/inflation_layer.hpp:121-135
static unsigned char computeCost(
  double distance, double resolution, double inscribed_radius, double cost_scaling_factor)
{
  unsigned char cost = 0;
  if (distance == 0) {
    cost = LETHAL_OBSTACLE;
  } else if (distance * resolution <= inscribed_radius) {
    cost = INSCRIBED_INFLATED_OBSTACLE;
  } else {
    double factor = exp(-1.0 * cost_scaling_factor * (distance * resolution - inscribed_radius));
    cost = static_cast<unsigned char>((INSCRIBED_INFLATED_OBSTACLE - 1) * factor);
  }
  return cost;
}

// ----- models (include/nav2_mppi_controller/models/*.hpp) -----
struct Twist {double vx = 0, vy = 0, wz = 0;};
struct State
{
  Tensor vx, vy, wz, cvx, cvy, cwz;
  double pose_x = 0, pose_y = 0, pose_yaw = 0;
  Twist speed;
  float local_path_length = 0.0f;
  void reset(unsigned int B, unsigned int T)  // state.hpp:49-58
  {
    vx.setZero(B, T); vy.setZero(B, T); wz.setZero(B, T);
    cvx.setZero(B, T); cvy.setZero(B, T); cwz.setZero(B, T);
  }
};
struct Trajectories
{
  Tensor x, y, yaws;
  void reset(unsigned int B, unsigned int T) {x.setZero(B, T); y.setZero(B, T); yaws.setZero(B, T);}
};
struct ControlSequence
{
  std::vector<float> vx, vy, wz;
  void reset(unsigned int T) {vx.assign(T, 0.0f); vy.assign(T, 0.0f); wz.assign(T, 0.0f);}
};
struct Control {float vx, vy, wz;};
struct Path
{
  std::vector<float> x, y, yaws;
  void reset(unsigned int n) {x.assign(n, 0.0f); y.assign(n, 0.0f); yaws.assign(n, 0.0f);}
};
struct ControlConstraints {float vx_max, vx_min, vy, wz, ax_max, ax_min, ay_min, ay_max, az_max;};

// critic_data.hpp:39-55
struct CriticData
{
  const State & state;
  const Trajectories & trajectories;
  const Path & path;
  std::vector<float> & costs;
  float & model_dt;
  bool fail_flag = false;
  bool has_goal_checker = false;
  double goal_checker_xy_tolerance = 0.0;
  int motion_model = MPPI_MODEL_DIFF_DRIVE;
  float min_turning_r = 0.0f;
  std::optional<std::vector<bool>> path_pts_valid;
  std::optional<size_t> furthest_reached_path_point;
  std::vector<bool> trajectories_in_collision;
  // debug: cell indices looked up by Cost / Obstacles critics (0xFFFFFFFF off-map, 0xFFFFFFFE not visited)
  std::vector<uint32_t> * cost_cells = nullptr;
  std::vector<uint32_t> * obstacle_cells = nullptr;
};

// ----- tools/utils.hpp -----

// utils.hpp:292-341
static size_t findPathFurthestReachedPoint(const CriticData & data)
{
  const int traj_cols = data.trajectories.x.cols;
  const int n_rows = data.trajectories.x.rows;
  const int n_cols = static_cast<int>(data.path.x.size());

  std::vector<float> path_integrated_dists(n_cols, 0.0f);
  for (int i = 1; i < n_cols; ++i) {
    const float dx = data.path.x[i] - data.path.x[i - 1];
    const float dy = data.path.y[i] - data.path.y[i - 1];
    path_integrated_dists[i] = path_integrated_dists[i - 1] + sqrtf(dx * dx + dy * dy);
  }

  // rowwise().sum() of sqrt(dx^2 + dy^2) over consecutive columns: sequential in t
  std::vector<float> traj_integrated_dists(n_rows, 0.0f);
  for (int b = 0; b < n_rows; ++b) {
    float acc = 0.0f;
    for (int t = 1; t < traj_cols; ++t) {
      const float dx = data.trajectories.x(b, t) - data.trajectories.x(b, t - 1);
      const float dy = data.trajectories.y(b, t) - data.trajectories.y(b, t - 1);
      const float seg = sqrtf(dx * dx + dy * dy);
      acc = (t == 1) ? seg : acc + seg;
    }
    traj_integrated_dists[b] = acc;
  }

  int max_idx = 0;
  for (int i = 0; i < n_rows; ++i) {
    int max_reachable_idx = static_cast<int>(
      std::lower_bound(
        path_integrated_dists.begin(), path_integrated_dists.end(),
        traj_integrated_dists[i]) - path_integrated_dists.begin());
    max_reachable_idx = std::min(max_reachable_idx, n_cols - 1);
    if (max_reachable_idx < 0) {continue;}

    const float ex = data.trajectories.x(i, traj_cols - 1);
    const float ey = data.trajectories.y(i, traj_cols - 1);
    int eucl_idx = 0;
    float best = 0.0f;
    for (int j = 0; j <= max_reachable_idx; ++j) {
      const float dx = data.path.x[j] - ex;
      const float dy = data.path.y[j] - ey;
      const float dsq = dx * dx + dy * dy;
      if (j == 0 || dsq < best) {best = dsq; eucl_idx = j;}
    }
    max_idx = std::max(max_idx, eucl_idx);
    if (max_idx == n_cols - 1) {
      break;
    }
  }
  return static_cast<size_t>(max_idx);
}

static void setPathFurthestPointIfNotSet(CriticData & data)  // utils.hpp:347-352
{
  if (!data.furthest_reached_path_point) {
    data.furthest_reached_path_point = findPathFurthestReachedPoint(data);
  }
}

// utils.hpp:358-387
static std::vector<bool> pathPointsValid(const Path & path, const Costmap & costmap, bool tracking_unknown)
{
  unsigned int map_x, map_y;
  const size_t path_segments_count = path.x.size() - 1;
  std::vector<bool> path_pts_valid(path_segments_count, false);
  for (unsigned int idx = 0; idx < path_segments_count; idx++) {
    if (!costmap.worldToMap(path.x[idx], path.y[idx], map_x, map_y)) {
      path_pts_valid[idx] = false;
      continue;
    }
    switch (costmap.getCost(map_x, map_y)) {
      case (LETHAL_OBSTACLE):
        path_pts_valid[idx] = false;
        continue;
      case (INSCRIBED_INFLATED_OBSTACLE):
        path_pts_valid[idx] = false;
        continue;
      case (NO_INFORMATION):
        path_pts_valid[idx] = tracking_unknown ? true : false;
        continue;
    }
    path_pts_valid[idx] = true;
  }
  return path_pts_valid;
}

static void findPathCosts(CriticData & data, const Costmap & costmap, bool tracking_unknown)
{
  data.path_pts_valid = pathPointsValid(data.path, costmap, tracking_unknown);
}

static void setPathCostsIfNotSet(CriticData & data, const Costmap & costmap, bool tracking_unknown)
{
  if (!data.path_pts_valid) {  // utils.hpp:393-400
    findPathCosts(data, costmap, tracking_unknown);
  }
}

// utils.hpp:410-427 (pose yaw is passed in place of the quaternion)
static float posePointAngle(
  double pose_x_d, double pose_y_d, double pose_yaw_d, double point_x, double point_y,
  bool forward_preference)
{
  float pose_x = pose_x_d;
  float pose_y = pose_y_d;
  float pose_yaw = pose_yaw_d;
  float yaw = atan2f(point_y - pose_y, point_x - pose_x);
  if (!forward_preference) {
    return std::min(
      fabs(mppi_angles_shortest_angular_distance(yaw, pose_yaw)),
      fabs(
        mppi_angles_shortest_angular_distance(
          yaw, mppi_angles_normalize_angle(pose_yaw + MPPI_PIF))));
  }
  return fabs(mppi_angles_shortest_angular_distance(yaw, pose_yaw));
}

// utils.hpp:437-452
static float posePointAngle(
  double pose_x_d, double pose_y_d, double pose_yaw_d, double point_x, double point_y,
  double point_yaw)
{
  float pose_x = static_cast<float>(pose_x_d);
  float pose_y = static_cast<float>(pose_y_d);
  float pose_yaw = static_cast<float>(pose_yaw_d);
  float yaw = atan2f(static_cast<float>(point_y) - pose_y, static_cast<float>(point_x) - pose_x);
  if (fabs(mppi_angles_shortest_angular_distance(yaw, static_cast<float>(point_yaw))) > MPPI_PIF_2) {
    yaw = mppi_angles_normalize_angle(yaw + MPPI_PIF);
  }
  return fabs(mppi_angles_shortest_angular_distance(yaw, pose_yaw));
}

// utils.hpp:460-542
static void savitskyGolayFilter(
  ControlSequence & control_sequence, std::array<Control, 4> & control_history,
  unsigned int sgf_order, bool shift_control_sequence)
{
  std::array<float, 9> filter;
  if (sgf_order == 1) {
    filter = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f};
    for (auto & f : filter) {f /= 9.0f;}
  } else {
    filter = {-21.0f, 14.0f, 39.0f, 54.0f, 59.0f, 54.0f, 39.0f, 14.0f, -21.0f};
    for (auto & f : filter) {f /= 231.0f;}
  }

  const unsigned int num_sequences = control_sequence.vx.size() - 1;
  if (num_sequences < 20) {
    return;
  }

  auto applyFilter = [&](const std::array<float, 9> & data) -> float {
      float s = data[0] * filter[0];
      for (int i = 1; i < 9; ++i) {s += data[i] * filter[i];}
      return s;
    };

  auto applyFilterOverAxis =
    [&](std::vector<float> & sequence, const std::vector<float> & initial_sequence,
    const float hist_0, const float hist_1, const float hist_2, const float hist_3) -> void
    {
      float pt_m4 = hist_0, pt_m3 = hist_1, pt_m2 = hist_2, pt_m1 = hist_3;
      float pt = initial_sequence[0];
      float pt_p1 = initial_sequence[1], pt_p2 = initial_sequence[2];
      float pt_p3 = initial_sequence[3], pt_p4 = initial_sequence[4];
      for (unsigned int idx = 0; idx != num_sequences; idx++) {
        sequence[idx] = applyFilter({pt_m4, pt_m3, pt_m2, pt_m1, pt, pt_p1, pt_p2, pt_p3, pt_p4});
        pt_m4 = pt_m3; pt_m3 = pt_m2; pt_m2 = pt_m1; pt_m1 = pt;
        pt = pt_p1; pt_p1 = pt_p2; pt_p2 = pt_p3; pt_p3 = pt_p4;
        if (idx + 5 < num_sequences) {
          pt_p4 = initial_sequence[idx + 5];
        } else {
          pt_p4 = initial_sequence[num_sequences];
        }
      }
    };

  const ControlSequence initial = control_sequence;
  applyFilterOverAxis(
    control_sequence.vx, initial.vx, control_history[0].vx, control_history[1].vx,
    control_history[2].vx, control_history[3].vx);
  applyFilterOverAxis(
    control_sequence.vy, initial.vy, control_history[0].vy, control_history[1].vy,
    control_history[2].vy, control_history[3].vy);
  applyFilterOverAxis(
    control_sequence.wz, initial.wz, control_history[0].wz, control_history[1].wz,
    control_history[2].wz, control_history[3].wz);

  unsigned int offset = shift_control_sequence ? 1 : 0;
  control_history[0] = control_history[1];
  control_history[1] = control_history[2];
  control_history[2] = control_history[3];
  control_history[3] = {
    control_sequence.vx[offset], control_sequence.vy[offset], control_sequence.wz[offset]};
}

// utils.hpp:550-567
static unsigned int findClosestPathPt(
  const std::vector<float> & vec, const float dist, const unsigned int init = 0u)
{
  float distim1 = init != 0u ? vec[init] : 0.0f;
  float disti = 0.0f;
  const unsigned int size = vec.size();
  for (unsigned int i = init + 1; i != size; i++) {
    disti = vec[i];
    if (disti > dist) {
      if (i > 0 && dist - distim1 < disti - dist) {
        return i - 1;
      }
      return i;
    }
    distim1 = disti;
  }
  return size - 1;
}

// utils.hpp:582-608 on a vector (1-D case)
static void shiftByOnePlace(std::vector<float> & e, int direction)
{
  const int size = static_cast<int>(e.size());
  if (size <= 1) {return;}
  if (direction == 1) {
    for (int i = size - 1; i > 0; --i) {e[i] = e[i - 1];}
  } else {
    for (int i = 0; i < size - 1; ++i) {e[i] = e[i + 1];}
  }
}

// ----- motion_models.hpp:39-252 -----
struct MotionModel
{
  int type = MPPI_MODEL_DIFF_DRIVE;
  float min_turning_r = 0.2f;
  float model_dt_ = 0.0f;
  float model_delay_vx_ = 0.0f, model_delay_vy_ = 0.0f, model_delay_wz_ = 0.0f;
  bool clamp_raw_controls_ = false;
  std::vector<float> cmd_history_vx_, cmd_history_vy_, cmd_history_wz_;
  ControlConstraints control_constraints_{0, 0, 0, 0, 0, 0, 0, 0, 0};

  bool isHolonomic() const {return type == MPPI_MODEL_OMNI;}

  std::size_t offsetSteps(float delay) const  // :221-227
  {
    if (delay <= 0.0f || model_dt_ <= 0.0f) {
      return 0u;
    }
    return static_cast<std::size_t>(std::floor(delay / model_dt_ + 0.5f));
  }

  void setConstraints(  // :67-81
    const ControlConstraints & c, float model_dt, float dvx, float dvy, float dwz, bool clamp_raw)
  {
    control_constraints_ = c;
    model_dt_ = model_dt;
    model_delay_vx_ = dvx; model_delay_vy_ = dvy; model_delay_wz_ = dwz;
    clamp_raw_controls_ = clamp_raw;
    cmd_history_vx_.resize(offsetSteps(model_delay_vx_), 0.0f);
    cmd_history_vy_.resize(offsetSteps(model_delay_vy_), 0.0f);
    cmd_history_wz_.resize(offsetSteps(model_delay_wz_), 0.0f);
  }
  static void pushOne(std::vector<float> & buf, float v)  // :232-237
  {
    if (buf.empty()) {return;}
    std::rotate(buf.begin(), buf.begin() + 1, buf.end());
    buf.back() = v;
  }
  void pushCommandHistory(float vx, float vy, float wz)  // :87-92
  {
    pushOne(cmd_history_vx_, vx); pushOne(cmd_history_vy_, vy); pushOne(cmd_history_wz_, wz);
  }
  void clearCommandHistory()  // :97-102
  {
    std::fill(cmd_history_vx_.begin(), cmd_history_vx_.end(), 0.0f);
    std::fill(cmd_history_vy_.begin(), cmd_history_vy_.end(), 0.0f);
    std::fill(cmd_history_wz_.begin(), cmd_history_wz_.end(), 0.0f);
  }

  void predict(State & state) const  // :108-167
  {
    const bool is_holo = isHolonomic();
    float max_delta_vx = model_dt_ * control_constraints_.ax_max;
    float min_delta_vx = model_dt_ * control_constraints_.ax_min;
    float max_delta_vy = model_dt_ * control_constraints_.ay_max;
    float min_delta_vy = model_dt_ * control_constraints_.ay_min;
    float max_delta_wz = model_dt_ * control_constraints_.az_max;
    unsigned int n_cols = state.vx.cols;
    const int B = state.vx.rows;

    for (unsigned int i = 1; i < n_cols; i++) {
      float * vx_i = state.vx.col(i);
      const float * vx_p = state.vx.col(i - 1);
      float * cvx_p = state.cvx.col(i - 1);
      for (int b = 0; b < B; ++b) {
        const float lower = (vx_p[b] > 0) ? vx_p[b] + min_delta_vx : vx_p[b] - max_delta_vx;
        const float upper = (vx_p[b] > 0) ? vx_p[b] + max_delta_vx : vx_p[b] - min_delta_vx;
        vx_i[b] = std::min(std::max(cvx_p[b], lower), upper);
        if (clamp_raw_controls_) {cvx_p[b] = vx_i[b];}
      }
      float * wz_i = state.wz.col(i);
      const float * wz_p = state.wz.col(i - 1);
      float * cwz_p = state.cwz.col(i - 1);
      for (int b = 0; b < B; ++b) {
        wz_i[b] = std::min(std::max(cwz_p[b], wz_p[b] - max_delta_wz), wz_p[b] + max_delta_wz);
        if (clamp_raw_controls_) {cwz_p[b] = wz_i[b];}
      }
      if (is_holo) {
        float * vy_i = state.vy.col(i);
        const float * vy_p = state.vy.col(i - 1);
        float * cvy_p = state.cvy.col(i - 1);
        for (int b = 0; b < B; ++b) {
          const float lower = (vy_p[b] > 0) ? vy_p[b] + min_delta_vy : vy_p[b] - max_delta_vy;
          const float upper = (vy_p[b] > 0) ? vy_p[b] + max_delta_vy : vy_p[b] - min_delta_vy;
          vy_i[b] = std::min(std::max(cvy_p[b], lower), upper);
          if (clamp_raw_controls_) {cvy_p[b] = vy_i[b];}
        }
      }
    }

    const unsigned int offset_vx = std::floor((model_delay_vx_ / model_dt_) + 0.5);
    const unsigned int offset_vy = std::floor((model_delay_vy_ / model_dt_) + 0.5);
    const unsigned int offset_wz = std::floor((model_delay_wz_ / model_dt_) + 0.5);
    if (offset_vx > 0u || offset_wz > 0u || (is_holo && offset_vy > 0u)) {
      applyDelayShift(state, is_holo, offset_vx, offset_vy, offset_wz);
    }
  }

  void applyDelayShift(  // :187-216
    State & state, bool is_holo, unsigned int offset_vx, unsigned int offset_vy,
    unsigned int offset_wz) const
  {
    auto shift = [](Tensor & velocities, unsigned int offset, const std::vector<float> & history) {
        const unsigned int cols = static_cast<unsigned int>(velocities.cols);
        if (offset == 0u || cols == 0u) {
          return;
        }
        const int B = velocities.rows;
        for (unsigned int k = (offset < cols) ? cols - offset : 0u; k > 0; --k) {
          std::memcpy(velocities.col(offset + k - 1), velocities.col(k), sizeof(float) * B);
        }
        const unsigned int end = std::min(offset, cols);
        for (unsigned int j = 1; j < end; ++j) {
          std::fill(velocities.col(j), velocities.col(j) + B, history[j]);
        }
      };
    shift(state.vx, offset_vx, cmd_history_vx_);
    shift(state.wz, offset_wz, cmd_history_wz_);
    if (is_holo) {
      shift(state.vy, offset_vy, cmd_history_vy_);
    }
  }

  void applyConstraints(ControlSequence & cs) const  // Ackermann :292-298; others no-op
  {
    if (type != MPPI_MODEL_ACKERMANN) {return;}
    for (size_t i = 0; i < cs.vx.size(); ++i) {
      const float wz_constrained = fabsf(cs.vx[i]) / min_turning_r;
      cs.wz[i] = std::min(std::max(cs.wz[i], -wz_constrained), wz_constrained);
    }
  }
};

// ----- settings (models/optimizer_settings.hpp) -----
struct OptimizerSettings
{
  ControlConstraints base_constraints{0, 0, 0, 0, 0, 0, 0, 0, 0};
  ControlConstraints constraints{0, 0, 0, 0, 0, 0, 0, 0, 0};
  float std_vx = 0, std_vy = 0, std_wz = 0;
  float model_dt = 0, model_delay_vx = 0, model_delay_vy = 0, model_delay_wz = 0;
  bool clamp_raw_controls = false;
  float controller_period = 0, temperature = 0, gamma = 0;
  unsigned int batch_size = 0, time_steps = 0, iteration_count = 0;
  bool shift_control_sequence = false;
  size_t retry_attempt_limit = 0;
  bool open_loop = false;
  unsigned int sgf_order = 2;
};

// What the critics read from Costmap2DROS / LayeredCostmap / InflationLayer.
struct CostmapRosInfo
{
  bool use_radius = true;
  std::vector<std::array<double, 2>> footprint;
  double inscribed_radius = 0, circumscribed_radius = 0;
  bool has_inflation_layer = false;
  double cost_scaling_factor = 0, inflation_radius = 0;
  bool track_unknown = false;
};

struct CollisionCost {float cost = 0; bool using_footprint = false;};  // critic_function.hpp:34-38

// ----- critics (src/critics/*.cpp) -----
struct Critic
{
  MppiCriticParams p;
  // Cost / Obstacles derived state
  float possible_collision_cost_ = 0.0f;
  float circumscribed_radius_ = 0.0f, circumscribed_cost_ = 0.0f;
  float inflation_scale_factor_ = 0.0f, inflation_radius_ = 0.0f;
  bool reversing_allowed_ = true;  // PathAngle
  int path_angle_mode_ = 0;
  float weight_ = 0.0f;  // CostCritic: cost_weight / 254
};

struct Optimizer;

struct CriticContext
{
  const Costmap * costmap;
  const CostmapRosInfo * ros;
  FootprintCollisionChecker checker;
  float param_vx_max, param_vx_min, param_vy_max;  // ConstraintCritic reads the parent params
};

// constraint_critic.cpp:37-95
static void scoreConstraint(const Critic & c, CriticData & data, const CriticContext & ctx)
{
  if (!c.p.enabled) {return;}
  const int B = data.state.vx.rows, T = data.state.vx.cols;
  const float vx_max_ = ctx.param_vx_max, vx_min_ = ctx.param_vx_min, vy_max_ = ctx.param_vy_max;
  const unsigned int power_ = c.p.cost_power;
  const float weight_ = c.p.cost_weight;
  for (int b = 0; b < B; ++b) {
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float vx = data.state.vx(b, t);
      float term = std::max(vx - vx_max_, 0.0f) + std::max(vx_min_ - vx, 0.0f);
      if (data.motion_model == MPPI_MODEL_OMNI) {
        term = term + std::max(fabsf(data.state.vy(b, t)) - vy_max_, 0.0f);
      } else if (data.motion_model == MPPI_MODEL_ACKERMANN) {
        const float wz_safe = std::max(fabsf(data.state.wz(b, t)), 1e-6f);
        term = term + std::max(data.min_turning_r - (fabsf(vx) / wz_safe), 0.0f);
      }
      const float v = term * data.model_dt;
      sum = (t == 0) ? v : sum + v;
    }
    if (power_ > 1u) {
      data.costs[b] += mppi_powu(sum * weight_, power_);
    } else {
      data.costs[b] += sum * weight_;
    }
  }
}

// cost_critic.cpp:85-129
static float costCriticFindCircumscribedCost(Critic & c, const CriticContext & ctx)
{
  double result = -1.0;
  const double circum_radius = ctx.ros->circumscribed_radius;
  if (static_cast<float>(circum_radius) == c.circumscribed_radius_) {
    return c.circumscribed_cost_;
  }
  if (ctx.ros->has_inflation_layer) {
    const double resolution = ctx.costmap->resolution;
    double inflation_radius = ctx.ros->inflation_radius;
    if (inflation_radius < circum_radius) {
      result = 0.0;
      return result;
    }
    result = computeCost(
      circum_radius / resolution, resolution, ctx.ros->inscribed_radius,
      ctx.ros->cost_scaling_factor);
  }
  c.circumscribed_radius_ = static_cast<float>(circum_radius);
  c.circumscribed_cost_ = static_cast<float>(result);
  return c.circumscribed_cost_;
}

// cost_critic.hpp:59-81
static bool costCriticInCollision(
  const Critic & c, const CriticContext & ctx, float cost, float x, float y, float theta)
{
  float score_cost = cost;
  if (c.p.consider_footprint &&
    (cost >= c.possible_collision_cost_ || c.possible_collision_cost_ < 1.0f))
  {
    score_cost = static_cast<float>(ctx.checker.footprintCostAtPose(
        static_cast<double>(x), static_cast<double>(y), static_cast<double>(theta),
        ctx.ros->footprint));
  }
  switch (static_cast<unsigned char>(score_cost)) {
    case (LETHAL_OBSTACLE):
      return true;
    case (INSCRIBED_INFLATED_OBSTACLE):
      return c.p.consider_footprint ? false : true;
    case (NO_INFORMATION):
      return ctx.ros->track_unknown ? false : true;
  }
  return false;
}

// cost_critic.cpp:131-231
static void scoreCost(Critic & c, CriticData & data, const CriticContext & ctx)
{
  if (!c.p.enabled) {return;}
  const Costmap * costmap = ctx.costmap;
  const float origin_x_ = static_cast<float>(costmap->origin_x);
  const float origin_y_ = static_cast<float>(costmap->origin_y);
  const float resolution_ = static_cast<float>(costmap->resolution);
  const unsigned int size_x_ = costmap->size_x, size_y_ = costmap->size_y;

  auto worldToMapFloat = [&](float wx, float wy, unsigned int & mx, unsigned int & my) {
      // cost_critic.hpp:100-113
      if (wx < origin_x_ || wy < origin_y_) {
        return false;
      }
      mx = static_cast<unsigned int>((wx - origin_x_) / resolution_);
      my = static_cast<unsigned int>((wy - origin_y_) / resolution_);
      if (mx < size_x_ && my < size_y_) {
        return true;
      }
      return false;
    };

  if (c.p.consider_footprint) {
    c.possible_collision_cost_ = costCriticFindCircumscribedCost(c, ctx);
  }

  bool near_goal = false;
  if (data.state.local_path_length < c.p.near_goal_distance) {
    near_goal = true;
  }

  const int B = data.trajectories.x.rows;
  const int T = data.trajectories.x.cols;
  std::vector<float> repulsive_cost(B, 0.0f);
  bool all_trajectories_collide = true;
  auto & collisions = data.trajectories_in_collision;
  const bool track_collisions = !collisions.empty();

  const int step = static_cast<int>(c.p.trajectory_point_step);
  int strided_traj_cols = static_cast<int>(floor((T - 1) / step)) + 1;
  if (data.cost_cells) {
    data.cost_cells->assign(static_cast<size_t>(strided_traj_cols) * B, 0xFFFFFFFEu);
  }

  for (int i = 0; i < B; ++i) {
    bool trajectory_collide = false;
    float pose_cost = 0.0f;
    float & traj_cost = repulsive_cost[i];

    for (int j = 0; j < strided_traj_cols; j++) {
      float Tx = data.trajectories.x(i, j * step);
      float Ty = data.trajectories.y(i, j * step);
      unsigned int x_i = 0u, y_i = 0u;
      if (!worldToMapFloat(Tx, Ty, x_i, y_i)) {
        pose_cost = 255.0f;
        if (data.cost_cells) {(*data.cost_cells)[static_cast<size_t>(j) * B + i] = 0xFFFFFFFFu;}
      } else {
        if (data.cost_cells) {
          (*data.cost_cells)[static_cast<size_t>(j) * B + i] = y_i * size_x_ + x_i;
        }
        pose_cost = static_cast<float>(costmap->getCost(y_i * size_x_ + x_i));
        if (pose_cost < 1.0f) {
          continue;
        }
      }

      if (costCriticInCollision(c, ctx, pose_cost, Tx, Ty, data.trajectories.yaws(i, j * step))) {
        traj_cost = c.p.collision_cost;
        trajectory_collide = true;
        if (track_collisions) {collisions[i] = true;}
        break;
      }

      if (pose_cost >= static_cast<float>(c.p.near_collision_cost)) {
        traj_cost += c.p.critical_cost;
      } else if (!near_goal) {
        traj_cost += pose_cost;
      }
    }
    all_trajectories_collide &= trajectory_collide;
  }

  const float scale = c.weight_ / static_cast<float>(strided_traj_cols);
  for (int i = 0; i < B; ++i) {
    if (c.p.cost_power > 1u) {
      data.costs[i] += mppi_powu(repulsive_cost[i] * scale, c.p.cost_power);
    } else {
      data.costs[i] += repulsive_cost[i] * scale;
    }
  }
  data.fail_flag = all_trajectories_collide;
}

// obstacles_critic.cpp:70-103
static float obstaclesFindCircumscribedCost(Critic & c, const CriticContext & ctx)
{
  double result = -1.0;
  const double circum_radius = ctx.ros->circumscribed_radius;
  if (static_cast<float>(circum_radius) == c.circumscribed_radius_) {
    return c.circumscribed_cost_;
  }
  if (ctx.ros->has_inflation_layer) {
    const double resolution = ctx.costmap->resolution;
    result = computeCost(
      circum_radius / resolution, resolution, ctx.ros->inscribed_radius,
      ctx.ros->cost_scaling_factor);
    c.inflation_scale_factor_ = static_cast<float>(ctx.ros->cost_scaling_factor);
    c.inflation_radius_ = static_cast<float>(ctx.ros->inflation_radius);
  }
  c.circumscribed_radius_ = static_cast<float>(circum_radius);
  c.circumscribed_cost_ = static_cast<float>(result);
  return c.circumscribed_cost_;
}

// obstacles_critic.cpp:105-118
static float obstaclesDistanceToObstacle(
  const Critic & c, const CriticContext & ctx, const CollisionCost & cost)
{
  const float scale_factor = c.inflation_scale_factor_;
  const float min_radius = ctx.ros->inscribed_radius;
  float dist_to_obj = (scale_factor * min_radius - log(cost.cost) + log(253.0f)) / scale_factor;
  if (!cost.using_footprint) {
    dist_to_obj -= min_radius;
  }
  return dist_to_obj;
}

// obstacles_critic.cpp:206-222
static bool obstaclesInCollision(const Critic & c, const CriticContext & ctx, float cost)
{
  switch (static_cast<unsigned char>(cost)) {
    case (LETHAL_OBSTACLE):
      return true;
    case (INSCRIBED_INFLATED_OBSTACLE):
      return c.p.consider_footprint ? false : true;
    case (NO_INFORMATION):
      return ctx.ros->track_unknown ? false : true;
  }
  return false;
}

// obstacles_critic.cpp:224-245
static CollisionCost obstaclesCostAtPose(
  const Critic & c, const CriticContext & ctx, float x, float y, float theta, uint32_t * cell_out)
{
  CollisionCost collision_cost;
  float & cost = collision_cost.cost;
  collision_cost.using_footprint = false;
  unsigned int x_i, y_i;
  if (!ctx.costmap->worldToMap(x, y, x_i, y_i)) {
    cost = NO_INFORMATION;
    if (cell_out) {*cell_out = 0xFFFFFFFFu;}
    return collision_cost;
  }
  if (cell_out) {*cell_out = y_i * ctx.costmap->size_x + x_i;}
  cost = ctx.checker.pointCost(x_i, y_i);
  if (c.p.consider_footprint &&
    (cost >= c.possible_collision_cost_ || c.possible_collision_cost_ < 1.0f))
  {
    cost = static_cast<float>(ctx.checker.footprintCostAtPose(x, y, theta, ctx.ros->footprint));
    collision_cost.using_footprint = true;
  }
  return collision_cost;
}

// obstacles_critic.cpp:120-199
static void scoreObstacles(Critic & c, CriticData & data, const CriticContext & ctx)
{
  if (!c.p.enabled) {return;}
  if (c.p.consider_footprint) {
    c.possible_collision_cost_ = obstaclesFindCircumscribedCost(c, ctx);
  }
  bool near_goal = false;
  if (data.state.local_path_length < c.p.near_goal_distance) {
    near_goal = true;
  }
  const unsigned int traj_len = data.trajectories.x.cols;
  const unsigned int batch_size = data.trajectories.x.rows;
  std::vector<float> raw_cost(batch_size, 0.0f), repulsive_cost(batch_size, 0.0f);
  bool all_trajectories_collide = true;
  auto & collisions = data.trajectories_in_collision;
  const bool track_collisions = !collisions.empty();
  if (data.obstacle_cells) {
    data.obstacle_cells->assign(static_cast<size_t>(traj_len) * batch_size, 0xFFFFFFFEu);
  }

  for (unsigned int i = 0; i != batch_size; i++) {
    bool trajectory_collide = false;
    float traj_cost = 0.0f;
    const auto & traj = data.trajectories;
    CollisionCost pose_cost;
    raw_cost[i] = 0.0f;
    repulsive_cost[i] = 0.0f;

    for (unsigned int j = 0; j != traj_len; j++) {
      uint32_t * cell =
        data.obstacle_cells ? &(*data.obstacle_cells)[static_cast<size_t>(j) * batch_size + i] :
        nullptr;
      pose_cost = obstaclesCostAtPose(c, ctx, traj.x(i, j), traj.y(i, j), traj.yaws(i, j), cell);
      if (pose_cost.cost < 1.0f) {continue;}

      if (obstaclesInCollision(c, ctx, pose_cost.cost)) {
        trajectory_collide = true;
        break;
      }
      if (c.inflation_radius_ == 0.0f || c.inflation_scale_factor_ == 0.0f) {
        continue;
      }
      const float dist_to_obj = obstaclesDistanceToObstacle(c, ctx, pose_cost);
      if (dist_to_obj < c.p.collision_margin_distance) {
        traj_cost += (c.p.collision_margin_distance - dist_to_obj);
      }
      if (!near_goal) {
        repulsive_cost[i] += c.inflation_radius_ - dist_to_obj;
      }
    }
    if (!trajectory_collide) {all_trajectories_collide = false;}
    raw_cost[i] = trajectory_collide ? c.p.collision_cost : traj_cost;
    if (trajectory_collide && track_collisions) {collisions[i] = true;}
  }

  float min_rep = repulsive_cost.empty() ? 0.0f : repulsive_cost[0];
  for (float v : repulsive_cost) {min_rep = std::min(min_rep, v);}
  for (unsigned int i = 0; i < batch_size; ++i) {
    const float rep_norm = (repulsive_cost[i] - min_rep) / traj_len;
    const float v = (c.p.critical_weight * raw_cost[i]) + (c.p.repulsion_weight * rep_norm);
    if (c.p.cost_power > 1u) {
      data.costs[i] += mppi_powu(v, c.p.cost_power);
    } else {
      data.costs[i] += v;
    }
  }
  data.fail_flag = all_trajectories_collide;
}

// goal_critic.cpp:35-56
static void scoreGoal(const Critic & c, CriticData & data)
{
  if (!c.p.enabled || data.state.local_path_length > c.p.threshold_to_consider) {
    return;
  }
  if (data.path.x.empty()) {return;}  // SURVEY.md §9 hazard: reference indexes size-1 (UB)
  const size_t last = data.path.x.size() - 1;
  // getLastPathPose stores the float path point into double Pose fields; Eigen narrows the scalar
  // back to float when it meets the float array.
  const float goal_x = static_cast<float>(static_cast<double>(data.path.x[last]));
  const float goal_y = static_cast<float>(static_cast<double>(data.path.y[last]));
  const int B = data.trajectories.x.rows, T = data.trajectories.x.cols;
  for (int b = 0; b < B; ++b) {
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float dx = data.trajectories.x(b, t) - goal_x;
      const float dy = data.trajectories.y(b, t) - goal_y;
      const float d = sqrtf(dx * dx + dy * dy);
      sum = (t == 0) ? d : sum + d;
    }
    const float mean = sum / static_cast<float>(T);
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(mean * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += mean * c.p.cost_weight;
    }
  }
}

// goal_angle_critic.cpp:36-64
static void scoreGoalAngle(const Critic & c, CriticData & data)
{
  if (!c.p.enabled || data.state.local_path_length > c.p.threshold_to_consider) {
    return;
  }
  if (data.path.x.empty()) {return;}
  const size_t last = data.path.x.size() - 1;
  // getLastPathPose -> setRPY(0,0,yaw) -> tf2::getYaw returns the same yaw (to double rounding)
  const float goal_yaw = static_cast<float>(static_cast<double>(data.path.yaws[last]));
  const int B = data.trajectories.yaws.rows, T = data.trajectories.yaws.cols;
  const bool sym = c.p.symmetric_yaw_tolerance != 0;
  const float symmetric_goal_yaw =
    static_cast<float>(mppi_angles_normalize_angle(goal_yaw + M_PI));
  for (int b = 0; b < B; ++b) {
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float yaw = data.trajectories.yaws(b, t);
      float d = fabsf(mppi_shortest_angular_distancef(yaw, goal_yaw));
      if (sym) {
        const float d2 = fabsf(mppi_shortest_angular_distancef(yaw, symmetric_goal_yaw));
        d = std::min(d, d2);
      }
      sum = (t == 0) ? d : sum + d;
    }
    const float mean = sum / static_cast<float>(T);
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(mean * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += mean * c.p.cost_weight;
    }
  }
}

// path_align_critic.cpp:40-159
static void scorePathAlign(const Critic & c, CriticData & data, const CriticContext & ctx)
{
  if (!c.p.enabled || data.state.local_path_length < c.p.threshold_to_consider) {
    return;
  }
  if (data.path.x.empty()) {return;}
  setPathFurthestPointIfNotSet(data);
  const size_t path_segments_count = *data.furthest_reached_path_point;
  float path_segments_flt = static_cast<float>(path_segments_count);
  if (path_segments_count < c.p.offset_from_furthest) {
    return;
  }
  setPathCostsIfNotSet(data, *ctx.costmap, ctx.ros->track_unknown);
  std::vector<bool> & path_pts_valid = *data.path_pts_valid;
  float invalid_ctr = 0.0f;
  for (size_t i = 0; i < path_segments_count; i++) {
    if (!path_pts_valid[i]) {invalid_ctr += 1.0f;}
    if (invalid_ctr / path_segments_flt > c.p.max_path_occupancy_ratio && invalid_ctr > 2.0f) {
      return;
    }
  }
  if (path_segments_count == 0) {return;}  // offset_from_furthest == 0 corner: nothing to align to

  const size_t batch_size = data.trajectories.x.rows;
  std::vector<float> cost(batch_size, 0.0f);
  std::vector<float> path_integrated_distances(path_segments_count, 0.0f);
  struct Pose2D {float x, y, theta;};
  std::vector<Pose2D> path(path_segments_count);
  float dx = 0.0f, dy = 0.0f;
  for (unsigned int i = 1; i != path_segments_count; i++) {
    auto & pose = path[i - 1];
    pose.x = data.path.x[i - 1];
    pose.y = data.path.y[i - 1];
    pose.theta = data.path.yaws[i - 1];
    dx = data.path.x[i] - pose.x;
    dy = data.path.y[i] - pose.y;
    path_integrated_distances[i] = path_integrated_distances[i - 1] + sqrtf(dx * dx + dy * dy);
  }
  auto & final_pose = path[path_segments_count - 1];
  final_pose.x = data.path.x[path_segments_count - 1];
  final_pose.y = data.path.y[path_segments_count - 1];
  final_pose.theta = data.path.yaws[path_segments_count - 1];

  float summed_path_dist = 0.0f, dyaw = 0.0f;
  unsigned int num_samples = 0u;
  unsigned int path_pt = 0u;
  float traj_integrated_distance = 0.0f;
  const int step = static_cast<int>(c.p.trajectory_point_step);
  const int T = data.trajectories.x.cols;
  const int traj_sampled_size = static_cast<int>(floor((T - 1) / step)) + 1;

  for (size_t t = 0; t < batch_size; ++t) {
    summed_path_dist = 0.0f;
    num_samples = 0u;
    traj_integrated_distance = 0.0f;
    path_pt = 0u;
    float Tx_m1 = data.trajectories.x(t, 0);
    float Ty_m1 = data.trajectories.y(t, 0);
    for (int p = 1; p < traj_sampled_size; p++) {
      const float Tx = data.trajectories.x(t, p * step);
      const float Ty = data.trajectories.y(t, p * step);
      dx = Tx - Tx_m1;
      dy = Ty - Ty_m1;
      Tx_m1 = Tx;
      Ty_m1 = Ty;
      traj_integrated_distance += sqrtf(dx * dx + dy * dy);
      path_pt = findClosestPathPt(path_integrated_distances, traj_integrated_distance, path_pt);
      if (path_pts_valid[path_pt]) {
        const auto & pose = path[path_pt];
        dx = pose.x - Tx;
        dy = pose.y - Ty;
        num_samples++;
        if (c.p.use_path_orientations) {
          dyaw = mppi_angles_shortest_angular_distance(pose.theta, data.trajectories.yaws(t, p * step));
          summed_path_dist += sqrtf(dx * dx + dy * dy + dyaw * dyaw);
        } else {
          summed_path_dist += sqrtf(dx * dx + dy * dy);
        }
      }
    }
    if (num_samples > 0u) {
      cost[t] = summed_path_dist / static_cast<float>(num_samples);
    } else {
      cost[t] = 0.0f;
    }
  }
  for (size_t b = 0; b < batch_size; ++b) {
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(cost[b] * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += cost[b] * c.p.cost_weight;
    }
  }
}

// path_angle_critic.cpp:62-148.  Returns MPPI_ERR_CONFIG for an invalid mode (:95 throws).
static int scorePathAngle(const Critic & c, CriticData & data)
{
  if (!c.p.enabled || data.state.local_path_length < c.p.threshold_to_consider) {
    return MPPI_OK;
  }
  if (data.path.x.empty()) {return MPPI_OK;}
  setPathFurthestPointIfNotSet(data);
  auto offsetted_idx = std::min(
    *data.furthest_reached_path_point + static_cast<size_t>(c.p.offset_from_furthest),
    static_cast<size_t>(data.path.x.size()) - 1);
  const float goal_x = data.path.x[offsetted_idx];
  const float goal_y = data.path.y[offsetted_idx];
  const float goal_yaw = data.path.yaws[offsetted_idx];
  const double px = data.state.pose_x, py = data.state.pose_y, pyaw = data.state.pose_yaw;

  switch (c.path_angle_mode_) {
    case 0:
      if (posePointAngle(px, py, pyaw, goal_x, goal_y, true) < c.p.max_angle_to_furthest) {
        return MPPI_OK;
      }
      break;
    case 1:
      if (posePointAngle(px, py, pyaw, goal_x, goal_y, false) < c.p.max_angle_to_furthest) {
        return MPPI_OK;
      }
      break;
    case 2:
      if (posePointAngle(px, py, pyaw, goal_x, goal_y, static_cast<double>(goal_yaw)) <
        c.p.max_angle_to_furthest)
      {
        return MPPI_OK;
      }
      break;
    default:
      return MPPI_ERR_CONFIG;
  }

  const int B = data.trajectories.y.rows;
  const int last_idx = data.trajectories.y.cols - 1;
  for (int b = 0; b < B; ++b) {
    const float diff_y = goal_y - data.trajectories.y(b, last_idx);
    const float diff_x = goal_x - data.trajectories.x(b, last_idx);
    const float yaw_between = atan2f(diff_y, diff_x);
    const float last_yaw = data.trajectories.yaws(b, last_idx);
    float v;
    if (c.path_angle_mode_ == 0) {
      v = fabsf(mppi_shortest_angular_distancef(last_yaw, yaw_between));
    } else if (c.path_angle_mode_ == 1) {
      // utils.hpp:617-631
      const float yaws = fabsf(mppi_shortest_angular_distancef(last_yaw, yaw_between));
      const float corrected = yaws < MPPI_PIF_2 ?
        yaw_between : static_cast<float>(mppi_angles_normalize_angle(yaw_between + MPPI_PIF));
      v = fabsf(mppi_shortest_angular_distancef(last_yaw, corrected));
    } else {
      // utils.hpp:639-651
      const float corrected =
        fabs(mppi_angles_normalize_angle(yaw_between - goal_yaw)) < MPPI_PIF_2 ?
        yaw_between : static_cast<float>(mppi_angles_normalize_angle(yaw_between + MPPI_PIF));
      v = fabsf(mppi_shortest_angular_distancef(last_yaw, corrected));
    }
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(v * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += v * c.p.cost_weight;
    }
  }
  return MPPI_OK;
}

// path_follow_critic.cpp:34-75
static void scorePathFollow(const Critic & c, CriticData & data, const CriticContext & ctx)
{
  if (!c.p.enabled) {return;}
  if (data.path.x.size() < 2 || data.state.local_path_length < c.p.threshold_to_consider) {
    return;
  }
  setPathFurthestPointIfNotSet(data);
  setPathCostsIfNotSet(data, *ctx.costmap, ctx.ros->track_unknown);
  const size_t path_size = data.path.x.size() - 1;
  auto offsetted_idx = std::min(
    *data.furthest_reached_path_point + static_cast<size_t>(c.p.offset_from_furthest), path_size);
  bool valid = false;
  while (!valid && offsetted_idx < path_size - 1) {
    valid = (*data.path_pts_valid)[offsetted_idx];
    if (!valid) {
      offsetted_idx++;
    }
  }
  const float path_x = data.path.x[offsetted_idx];
  const float path_y = data.path.y[offsetted_idx];
  const int B = data.trajectories.x.rows;
  const int last = data.trajectories.x.cols - 1;
  for (int b = 0; b < B; ++b) {
    const float delta_x = data.trajectories.x(b, last) - path_x;
    const float delta_y = data.trajectories.y(b, last) - path_y;
    const float d = sqrtf(delta_x * delta_x + delta_y * delta_y);
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(d * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += d * c.p.cost_weight;
    }
  }
}

// prefer_forward_critic.cpp:36-54
static void scorePreferForward(const Critic & c, CriticData & data)
{
  if (!c.p.enabled) {return;}
  if (data.state.local_path_length < c.p.threshold_to_consider) {
    return;
  }
  const int B = data.state.vx.rows, T = data.state.vx.cols;
  for (int b = 0; b < B; ++b) {
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float v = std::max(-data.state.vx(b, t), 0.0f) * data.model_dt;
      sum = (t == 0) ? v : sum + v;
    }
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(sum * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += sum * c.p.cost_weight;
    }
  }
}

// twirling_critic.cpp:33-55
static void scoreTwirling(const Critic & c, CriticData & data)
{
  if (!c.p.enabled) {return;}
  if (data.has_goal_checker) {
    if (data.state.local_path_length < data.goal_checker_xy_tolerance) {
      return;
    }
  }
  const int B = data.state.wz.rows, T = data.state.wz.cols;
  for (int b = 0; b < B; ++b) {
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float v = fabsf(data.state.wz(b, t));
      sum = (t == 0) ? v : sum + v;
    }
    const float mean = sum / static_cast<float>(T);
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(mean * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += mean * c.p.cost_weight;
    }
  }
}

// velocity_deadband_critic.cpp:41-72
static void scoreVelocityDeadband(const Critic & c, CriticData & data)
{
  if (!c.p.enabled) {return;}
  const int B = data.state.vx.rows, T = data.state.vx.cols;
  const bool holo = data.motion_model == MPPI_MODEL_OMNI;
  const float db0 = fabs(c.p.deadband_velocities[0]);
  const float db1 = fabs(c.p.deadband_velocities[1]);
  const float db2 = fabs(c.p.deadband_velocities[2]);
  for (int b = 0; b < B; ++b) {
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
      float term = std::max(db0 - fabsf(data.state.vx(b, t)), 0.0f);
      if (holo) {
        term = term + std::max(db1 - fabsf(data.state.vy(b, t)), 0.0f);
      }
      term = term + std::max(db2 - fabsf(data.state.wz(b, t)), 0.0f);
      const float v = term * data.model_dt;
      sum = (t == 0) ? v : sum + v;
    }
    if (c.p.cost_power > 1u) {
      data.costs[b] += mppi_powu(sum * c.p.cost_weight, c.p.cost_power);
    } else {
      data.costs[b] += sum * c.p.cost_weight;
    }
  }
}

// ----- optimal_trajectory_validator.hpp:117-158 -----
static int validateTrajectory(
  const std::vector<float> & traj /* x[T], y[T], yaw[T] */, unsigned int T,
  unsigned int traj_samples_to_evaluate, bool consider_footprint, const Costmap & costmap,
  const CostmapRosInfo & ros)
{
  FootprintCollisionChecker checker{&costmap};
  for (size_t i = 0; i < traj_samples_to_evaluate; ++i) {
    const double x = static_cast<double>(traj[i]);
    const double y = static_cast<double>(traj[T + i]);
    if (consider_footprint) {
      const double theta = static_cast<double>(traj[2 * T + i]);
      double footprint_cost = checker.footprintCostAtPose(x, y, theta, ros.footprint);
      if (footprint_cost == static_cast<double>(LETHAL_OBSTACLE)) {
        return MPPI_VALIDATION_SOFT_RESET;
      }
    } else {
      unsigned int x_i = 0u, y_i = 0u;
      if (!costmap.worldToMap(x, y, x_i, y_i)) {
        continue;
      }
      unsigned char cost = costmap.getCost(x_i, y_i);
      if (cost == INSCRIBED_INFLATED_OBSTACLE || cost == LETHAL_OBSTACLE) {
        return MPPI_VALIDATION_SOFT_RESET;
      }
    }
  }
  return MPPI_VALIDATION_SUCCESS;
}

// ----- src/optimizer.cpp -----
struct Optimizer
{
  MppiConfig cfg;
  OptimizerSettings settings_;
  MotionModel motion_model_;
  State state_;
  ControlSequence control_sequence_;
  std::array<Control, 4> control_history_{};
  Trajectories generated_trajectories_;
  Path path_;
  std::vector<float> costs_;
  Twist last_command_vel_;
  float model_dt_ref_ = 0.0f;
  Tensor noises_vx_, noises_vy_, noises_wz_;
  bool noise_injected_ = false;
  uint64_t noise_epoch_ = 0;
  std::vector<Critic> critics_;
  Costmap costmap_;
  CostmapRosInfo ros_;
  size_t fallback_counter_ = 0;  // optimizer.cpp:283 function-static, per instance here
  unsigned int traj_samples_to_evaluate_ = 0;
  // per-call
  bool fail_flag_ = false;
  std::optional<size_t> furthest_;
  std::optional<std::vector<bool>> path_pts_valid_;
  bool has_goal_checker_ = false;
  double goal_checker_tol_ = 0.0;
  std::vector<std::vector<float>> critic_costs_;
  std::vector<bool> collisions_;
  std::vector<uint32_t> cost_cells_, obstacle_cells_;
  std::vector<float> softmaxes_;
  std::string error_;

  bool isHolonomic() const {return motion_model_.isHolonomic();}

  int configure(const MppiConfig & c)  // optimizer.cpp:34-70, 77-161
  {
    cfg = c;
    auto & s = settings_;
    s.model_dt = c.model_dt;
    s.model_delay_vx = c.model_delay_vx;
    s.model_delay_vy = c.model_delay_vy;
    s.model_delay_wz = c.model_delay_wz;
    s.clamp_raw_controls = c.clamp_raw_controls != 0;
    s.time_steps = c.time_steps;
    s.batch_size = c.batch_size;
    s.iteration_count = c.iteration_count;
    s.temperature = c.temperature;
    s.gamma = c.gamma;
    s.base_constraints.vx_max = c.vx_max;
    s.base_constraints.vx_min = c.vx_min;
    s.base_constraints.vy = c.vy_max;
    s.base_constraints.wz = c.wz_max;
    s.base_constraints.ax_max = c.ax_max;
    s.base_constraints.ax_min = c.ax_min;
    s.base_constraints.ay_max = c.ay_max;
    s.base_constraints.ay_min = c.ay_min;
    s.base_constraints.az_max = c.az_max;
    s.std_vx = c.vx_std; s.std_vy = c.vy_std; s.std_wz = c.wz_std;
    s.retry_attempt_limit = c.retry_attempt_limit;
    s.open_loop = c.open_loop != 0;
    s.sgf_order = c.sgf_order;
    if (s.sgf_order < 1 || s.sgf_order > 2) {s.sgf_order = 2;}  // :130-133
    s.base_constraints.ax_max = fabs(s.base_constraints.ax_max);  // :135
    if (s.base_constraints.ax_min > 0.0) {s.base_constraints.ax_min = -1.0 * s.base_constraints.ax_min;}
    if (s.base_constraints.ay_min > 0.0) {s.base_constraints.ay_min = -1.0 * s.base_constraints.ay_min;}
    s.constraints = s.base_constraints;

    if (c.motion_model < 0 || c.motion_model > 2) {error_ = "bad motion model"; return MPPI_ERR_CONFIG;}
    motion_model_.type = c.motion_model;
    motion_model_.min_turning_r = c.min_turning_r;
    motion_model_.setConstraints(
      s.constraints, s.model_dt, s.model_delay_vx, s.model_delay_vy, s.model_delay_wz,
      s.clamp_raw_controls);

    s.controller_period = static_cast<float>(1.0 / c.controller_frequency);  // :159
    // setOffset :163-181
    {
      constexpr double eps = 1e-6;
      const double controller_period = s.controller_period;
      s.shift_control_sequence = false;
      if ((controller_period + eps) < s.model_dt) {
        // warn only
      } else if (std::abs(controller_period - s.model_dt) < eps) {
        s.shift_control_sequence = true;
      } else {
        error_ = "Controller period more then model dt, set it equal to model dt";
        return MPPI_ERR_CONFIG;
      }
    }

    ros_.use_radius = c.use_radius != 0;
    ros_.footprint.clear();
    for (uint32_t i = 0; i < c.n_footprint; ++i) {
      ros_.footprint.push_back({c.footprint_x[i], c.footprint_y[i]});
    }
    ros_.inscribed_radius = c.inscribed_radius;
    ros_.circumscribed_radius = c.circumscribed_radius;
    ros_.has_inflation_layer = c.has_inflation_layer != 0;
    ros_.cost_scaling_factor = c.inflation_cost_scaling_factor;
    ros_.inflation_radius = c.inflation_radius;
    ros_.track_unknown = c.track_unknown != 0;

    // CriticManager::on_configure / loadCritics (critic_manager.cpp:44-69) + each initialize()
    critics_.clear();
    for (uint32_t i = 0; i < c.n_critics; ++i) {
      Critic k;
      k.p = c.critics[i];
      if (k.p.type == MPPI_CRITIC_COST) {
        k.weight_ = k.p.cost_weight / 254.0f;  // cost_critic.cpp:39
        if (ros_.use_radius == (k.p.consider_footprint != 0) && ros_.use_radius) {
          error_ = "Considering footprint in CostCritic but no robot footprint provided";
          return MPPI_ERR_CONFIG;  // cost_critic.cpp:61-71
        }
      }
      if (k.p.type == MPPI_CRITIC_OBSTACLES) {
        if (ros_.use_radius == (k.p.consider_footprint != 0) && ros_.use_radius) {
          error_ = "Considering footprint in ObstaclesCritic but no robot footprint provided";
          return MPPI_ERR_CONFIG;  // obstacles_critic.cpp:50-60
        }
      }
      if (k.p.type == MPPI_CRITIC_PATH_ANGLE) {
        // path_angle_critic.cpp:25-50
        const float vx_min = c.vx_min;
        k.reversing_allowed_ = true;
        if (fabs(vx_min) < 1e-6f) {
          k.reversing_allowed_ = false;
        } else if (vx_min < 0.0f) {
          k.reversing_allowed_ = true;
        }
        k.path_angle_mode_ = k.p.mode;
        if (!k.reversing_allowed_ && k.path_angle_mode_ == 1) {
          k.path_angle_mode_ = 0;
        }
      }
      critics_.push_back(k);
    }

    // validator initialize (optimal_trajectory_validator.hpp:87-97)
    traj_samples_to_evaluate_ = c.collision_lookahead_time / s.model_dt;
    if (traj_samples_to_evaluate_ > s.time_steps) {
      traj_samples_to_evaluate_ = s.time_steps;
    }
    reset(true);
    return MPPI_OK;
  }

  // Critic initialize() calls findCircumscribedCost against the costmap (cost_critic.cpp:47-48,
  // obstacles_critic.cpp:36-37).  Here the costmap arrives after configure, so this runs on upload.
  void initCriticsCostmap()
  {
    CriticContext ctx = makeContext();
    for (auto & k : critics_) {
      if (k.p.type == MPPI_CRITIC_COST) {
        k.circumscribed_radius_ = 0.0f; k.circumscribed_cost_ = 0.0f;
        k.possible_collision_cost_ = costCriticFindCircumscribedCost(k, ctx);
      } else if (k.p.type == MPPI_CRITIC_OBSTACLES) {
        k.circumscribed_radius_ = 0.0f; k.circumscribed_cost_ = 0.0f;
        k.possible_collision_cost_ = obstaclesFindCircumscribedCost(k, ctx);
      }
    }
  }

  CriticContext makeContext() const
  {
    CriticContext ctx;
    ctx.costmap = &costmap_;
    ctx.ros = &ros_;
    ctx.checker.costmap = &costmap_;
    ctx.param_vx_max = cfg.vx_max;
    ctx.param_vx_min = cfg.vx_min;
    ctx.param_vy_max = cfg.vy_max;
    return ctx;
  }

  void generateNoisedControls()  // noise_generator.cpp:108-119 (engine: SURVEY.md §8d mt19937(42))
  {
    auto & s = settings_;
    std::mt19937 generator(42u + static_cast<uint32_t>(noise_epoch_++));
    std::normal_distribution<float> ndvx(0.0f, s.std_vx), ndvy(0.0f, s.std_vy), ndwz(0.0f, s.std_wz);
    noises_vx_.setZero(s.batch_size, s.time_steps);
    noises_vy_.setZero(s.batch_size, s.time_steps);
    noises_wz_.setZero(s.batch_size, s.time_steps);
    for (auto & v : noises_vx_.d) {v = ndvx(generator);}
    for (auto & v : noises_wz_.d) {v = ndwz(generator);}
    if (isHolonomic()) {
      for (auto & v : noises_vy_.d) {v = ndvy(generator);}
    }
  }

  void reset(bool reset_dynamic_speed_limits = true)  // optimizer.cpp:183-211
  {
    state_.reset(settings_.batch_size, settings_.time_steps);
    control_sequence_.reset(settings_.time_steps);
    control_history_[0] = {0.0f, 0.0f, 0.0f};
    control_history_[1] = {0.0f, 0.0f, 0.0f};
    control_history_[2] = {0.0f, 0.0f, 0.0f};
    control_history_[3] = {0.0f, 0.0f, 0.0f};
    last_command_vel_ = Twist();
    if (reset_dynamic_speed_limits) {
      settings_.constraints = settings_.base_constraints;
    }
    costs_.assign(settings_.batch_size, 0.0f);
    generated_trajectories_.reset(settings_.batch_size, settings_.time_steps);
    // noise_generator_.reset (noise_generator.cpp:77-96): zero then regenerate; injected noise is kept
    if (!noise_injected_ ||
      noises_vx_.rows != static_cast<int>(settings_.batch_size) ||
      noises_vx_.cols != static_cast<int>(settings_.time_steps))
    {
      noise_injected_ = false;
      generateNoisedControls();
    }
    motion_model_.setConstraints(
      settings_.constraints, settings_.model_dt, settings_.model_delay_vx, settings_.model_delay_vy,
      settings_.model_delay_wz, settings_.clamp_raw_controls);
    motion_model_.clearCommandHistory();
    traj_samples_to_evaluate_ = cfg.collision_lookahead_time / settings_.model_dt;
    if (traj_samples_to_evaluate_ > settings_.time_steps) {
      traj_samples_to_evaluate_ = settings_.time_steps;
    }
  }

  void setSpeedLimit(double speed_limit, bool percentage)  // optimizer.cpp:691-719
  {
    auto & s = settings_;
    if (speed_limit == 0.0) {  // nav2_costmap_2d::NO_SPEED_LIMIT
      s.constraints.vx_max = s.base_constraints.vx_max;
      s.constraints.vx_min = s.base_constraints.vx_min;
      s.constraints.vy = s.base_constraints.vy;
      s.constraints.wz = s.base_constraints.wz;
    } else {
      double ratio = percentage ? speed_limit / 100.0 : speed_limit / s.base_constraints.vx_max;
      s.constraints.vx_max = s.base_constraints.vx_max * ratio;
      s.constraints.vx_min = s.base_constraints.vx_min * ratio;
      s.constraints.vy = s.base_constraints.vy * ratio;
      s.constraints.wz = s.base_constraints.wz * ratio;
    }
    motion_model_.setConstraints(
      settings_.constraints, settings_.model_dt, settings_.model_delay_vx, settings_.model_delay_vy,
      settings_.model_delay_wz, settings_.clamp_raw_controls);
  }

  void prepare(const MppiCycleIn & in)  // optimizer.cpp:300-343
  {
    Twist robot_speed{in.speed_vx, in.speed_vy, in.speed_wz};
    if (settings_.open_loop) {
      state_.speed = last_command_vel_;
    } else {
      const auto & c = settings_.constraints;
      const double dt = settings_.controller_period;
      state_.speed = robot_speed;
      state_.speed.vx = std::clamp(
        last_command_vel_.vx, robot_speed.vx + dt * c.ax_min, robot_speed.vx + dt * c.ax_max);
      state_.speed.wz = std::clamp(
        last_command_vel_.wz, robot_speed.wz - dt * c.az_max, robot_speed.wz + dt * c.az_max);
      if (isHolonomic()) {
        state_.speed.vy = std::clamp(
          last_command_vel_.vy, robot_speed.vy + dt * c.ay_min, robot_speed.vy + dt * c.ay_max);
      }
    }
    state_.pose_x = in.pose_x; state_.pose_y = in.pose_y; state_.pose_yaw = in.pose_yaw;
    // nav2_util/geometry_utils.hpp:155-165 calculate_path_length (double, std::hypot)
    double path_length = 0.0;
    if (in.n_path >= 2) {
      for (size_t idx = 0; idx < in.n_path - 1; ++idx) {
        path_length += std::hypot(in.path_x[idx] - in.path_x[idx + 1], in.path_y[idx] - in.path_y[idx + 1]);
      }
    }
    if (in.has_local_path_length) {path_length = in.local_path_length;}
    state_.local_path_length = path_length;
    // utils::toTensor utils.hpp:208-220
    path_.reset(in.n_path);
    for (size_t i = 0; i < in.n_path; ++i) {
      path_.x[i] = in.path_x[i];
      path_.y[i] = in.path_y[i];
      path_.yaws[i] = in.path_yaw[i];
    }
    costs_.assign(settings_.batch_size, 0.0f);
    fail_flag_ = false;
    has_goal_checker_ = in.has_goal_checker != 0;
    goal_checker_tol_ = in.goal_checker_xy_tolerance;
    furthest_.reset();
    path_pts_valid_.reset();
  }

  void applyControlSequenceInterIterationConstraints()  // optimizer.cpp:368-402
  {
    auto & s = settings_;
    float first_dt = s.controller_period;
    float max_delta_vx = first_dt * s.constraints.ax_max;
    float min_delta_vx = first_dt * s.constraints.ax_min;
    float max_delta_vy = first_dt * s.constraints.ay_max;
    float min_delta_vy = first_dt * s.constraints.ay_min;
    float max_delta_wz = first_dt * s.constraints.az_max;
    float speed_vx = static_cast<float>(state_.speed.vx);
    float speed_wz = static_cast<float>(state_.speed.wz);
    if (s.shift_control_sequence) {
      control_sequence_.vx[0] = speed_vx;
      control_sequence_.wz[0] = speed_wz;
      if (isHolonomic()) {
        control_sequence_.vy[0] = static_cast<float>(state_.speed.vy);
      }
    } else {
      control_sequence_.vx[0] = mppi_clamp_velocity_by_accel(
        speed_vx, control_sequence_.vx[0], min_delta_vx, max_delta_vx);
      control_sequence_.wz[0] = mppi_clamp_velocity_by_accel(
        speed_wz, control_sequence_.wz[0], -max_delta_wz, max_delta_wz);
      if (isHolonomic()) {
        float speed_vy = static_cast<float>(state_.speed.vy);
        control_sequence_.vy[0] = mppi_clamp_velocity_by_accel(
          speed_vy, control_sequence_.vy[0], min_delta_vy, max_delta_vy);
      }
    }
  }

  void setNoisedControls()  // noise_generator.cpp:66-75
  {
    const int B = settings_.batch_size, T = settings_.time_steps;
    for (int t = 0; t < T; ++t) {
      for (int b = 0; b < B; ++b) {
        state_.cvx(b, t) = noises_vx_(b, t) + control_sequence_.vx[t];
        state_.cvy(b, t) = noises_vy_(b, t) + control_sequence_.vy[t];
        state_.cwz(b, t) = noises_wz_(b, t) + control_sequence_.wz[t];
      }
    }
  }

  void updateInitialStateVelocities()  // optimizer.cpp:473-481
  {
    const int B = settings_.batch_size;
    std::fill(state_.vx.col(0), state_.vx.col(0) + B, static_cast<float>(state_.speed.vx));
    std::fill(state_.wz.col(0), state_.wz.col(0) + B, static_cast<float>(state_.speed.wz));
    if (isHolonomic()) {
      std::fill(state_.vy.col(0), state_.vy.col(0) + B, static_cast<float>(state_.speed.vy));
    }
  }

  // optimizer.cpp:539-580
  void integrateStateVelocities(Trajectories & trajectories, const State & state) const
  {
    auto initial_yaw = static_cast<float>(state.pose_yaw);
    const int n_cols = trajectories.yaws.cols;
    const int B = trajectories.yaws.rows;
    const float dt = settings_.model_dt;
    const bool holo = isHolonomic();

    std::vector<float> last_yaws(B, initial_yaw);
    for (int i = 0; i != n_cols; i++) {
      const float * wz = state.wz.col(i);
      float * yaws = trajectories.yaws.col(i);
      for (int b = 0; b < B; ++b) {
        last_yaws[b] += wz[b] * dt;
        yaws[b] = last_yaws[b];
      }
    }
    // cos/sin of all yaws shifted right by one column, column 0 = cos/sin(initial_yaw) (:552-557).
    // dx, dy (:559-565); last_x/last_y accumulate (:567-579)
    float c0, s0;
    mppi_sincosf(initial_yaw, &s0, &c0);
    std::vector<float> last_x(B, static_cast<float>(state.pose_x));
    std::vector<float> last_y(B, static_cast<float>(state.pose_y));
    for (int i = 0; i != n_cols; i++) {
      const float * vx = state.vx.col(i);
      const float * vy = state.vy.col(i);
      const float * yaw_prev = (i > 0) ? trajectories.yaws.col(i - 1) : nullptr;
      float * X = trajectories.x.col(i);
      float * Y = trajectories.y.col(i);
      for (int b = 0; b < B; ++b) {
        float c = c0, s = s0;
        if (yaw_prev) {mppi_sincosf(yaw_prev[b], &s, &c);}
        float dx = vx[b] * c;
        float dy = vx[b] * s;
        if (holo) {
          dx = dx - vy[b] * s;
          dy = dy + vy[b] * c;
        }
        last_x[b] += dx * dt;
        last_y[b] += dy * dt;
        X[b] = last_x[b];
        Y[b] = last_y[b];
      }
    }
  }

  // optimizer.cpp:489-537 (single sequence) via getOptimizedTrajectory :582-598
  std::vector<float> getOptimizedTrajectory() const
  {
    const unsigned int T = settings_.time_steps;
    const bool holo = isHolonomic();
    std::vector<float> traj(3 * T, 0.0f);
    float * traj_x = traj.data(), * traj_y = traj.data() + T, * traj_yaws = traj.data() + 2 * T;
    float initial_yaw = static_cast<float>(state_.pose_yaw);
    const float dt = settings_.model_dt;
    if (T == 0) {return traj;}
    float last_yaw = initial_yaw;
    for (size_t i = 0; i != T; i++) {
      last_yaw += control_sequence_.wz[i] * dt;
      traj_yaws[i] = last_yaw;
    }
    float last_x = state_.pose_x;
    float last_y = state_.pose_y;
    for (size_t i = 0; i != T; i++) {
      float c, s;
      mppi_sincosf(i == 0 ? initial_yaw : traj_yaws[i - 1], &s, &c);
      float dx = control_sequence_.vx[i] * c;
      float dy = control_sequence_.vx[i] * s;
      if (holo) {
        dx = dx - control_sequence_.vy[i] * s;
        dy = dy + control_sequence_.vy[i] * c;
      }
      last_x += dx * dt;
      last_y += dy * dt;
      traj_x[i] = last_x;
      traj_y[i] = last_y;
    }
    return traj;
  }

  void generateNoisedTrajectories()  // optimizer.cpp:359-366
  {
    applyControlSequenceInterIterationConstraints();
    setNoisedControls();
    updateInitialStateVelocities();
    motion_model_.predict(state_);
    integrateStateVelocities(generated_trajectories_, state_);
  }

  // critic_manager.cpp:76-121
  int evalTrajectoriesScores(CriticData & data)
  {
    const bool visualize = cfg.visualize != 0;
    if (visualize) {
      data.trajectories_in_collision.assign(data.costs.size(), false);
    }
    critic_costs_.assign(critics_.size(), std::vector<float>(data.costs.size(), 0.0f));
    CriticContext ctx = makeContext();
    for (size_t i = 0; i < critics_.size(); ++i) {
      if (data.fail_flag) {
        break;
      }
      std::vector<float> costs_before = data.costs;
      Critic & k = critics_[i];
      switch (k.p.type) {
        case MPPI_CRITIC_CONSTRAINT: scoreConstraint(k, data, ctx); break;
        case MPPI_CRITIC_COST: scoreCost(k, data, ctx); break;
        case MPPI_CRITIC_GOAL: scoreGoal(k, data); break;
        case MPPI_CRITIC_GOAL_ANGLE: scoreGoalAngle(k, data); break;
        case MPPI_CRITIC_OBSTACLES: scoreObstacles(k, data, ctx); break;
        case MPPI_CRITIC_PATH_ALIGN: scorePathAlign(k, data, ctx); break;
        case MPPI_CRITIC_PATH_ANGLE: {
            int rc = scorePathAngle(k, data);
            if (rc != MPPI_OK) {error_ = "Invalid path angle mode!"; return rc;}
            break;
          }
        case MPPI_CRITIC_PATH_FOLLOW: scorePathFollow(k, data, ctx); break;
        case MPPI_CRITIC_PREFER_FORWARD: scorePreferForward(k, data); break;
        case MPPI_CRITIC_TWIRLING: scoreTwirling(k, data); break;
        case MPPI_CRITIC_VELOCITY_DEADBAND: scoreVelocityDeadband(k, data); break;
        default: error_ = "unknown critic"; return MPPI_ERR_CONFIG;
      }
      for (size_t b = 0; b < data.costs.size(); ++b) {
        critic_costs_[i][b] = data.costs[b] - costs_before[b];
      }
    }
    return MPPI_OK;
  }

  CriticData makeCriticData(const State & st, const Trajectories & tr)
  {
    CriticData data{st, tr, path_, costs_, model_dt_ref_};
    model_dt_ref_ = settings_.model_dt;
    data.fail_flag = fail_flag_;
    data.has_goal_checker = has_goal_checker_;
    data.goal_checker_xy_tolerance = goal_checker_tol_;
    data.motion_model = motion_model_.type;
    data.min_turning_r = motion_model_.min_turning_r;
    data.furthest_reached_path_point = furthest_;
    data.path_pts_valid = path_pts_valid_;
    if (cfg.debug_cells) {
      data.cost_cells = &cost_cells_;
      data.obstacle_cells = &obstacle_cells_;
    }
    return data;
  }
  void storeCriticData(CriticData & data)
  {
    fail_flag_ = data.fail_flag;
    furthest_ = data.furthest_reached_path_point;
    path_pts_valid_ = data.path_pts_valid;
    collisions_ = data.trajectories_in_collision;
  }

  void applyControlSequenceConstraints()  // optimizer.cpp:404-464
  {
    auto & s = settings_;
    motion_model_.applyConstraints(control_sequence_);
    float first_dt = s.controller_period;
    float max_delta_vx = first_dt * s.constraints.ax_max;
    float min_delta_vx = first_dt * s.constraints.ax_min;
    float max_delta_vy = first_dt * s.constraints.ay_max;
    float min_delta_vy = first_dt * s.constraints.ay_min;
    float max_delta_wz = first_dt * s.constraints.az_max;
    float vx_last = static_cast<float>(state_.speed.vx);
    float wz_last = static_cast<float>(state_.speed.wz);
    float vy_last = isHolonomic() ? static_cast<float>(state_.speed.vy) : 0.0f;
    if (s.shift_control_sequence) {
      control_sequence_.vx[0] = vx_last;
      control_sequence_.wz[0] = wz_last;
      if (isHolonomic()) {
        control_sequence_.vy[0] = vy_last;
      }
    }
    for (unsigned int i = 0; i != control_sequence_.vx.size(); i++) {
      if (i == 1) {
        max_delta_vx = s.model_dt * s.constraints.ax_max;
        min_delta_vx = s.model_dt * s.constraints.ax_min;
        max_delta_vy = s.model_dt * s.constraints.ay_max;
        min_delta_vy = s.model_dt * s.constraints.ay_min;
        max_delta_wz = s.model_dt * s.constraints.az_max;
      }
      float & vx_curr = control_sequence_.vx[i];
      vx_curr = mppi_clampf(s.constraints.vx_min, s.constraints.vx_max, vx_curr);
      vx_curr = mppi_clamp_velocity_by_accel(vx_last, vx_curr, min_delta_vx, max_delta_vx);
      vx_last = vx_curr;
      float & wz_curr = control_sequence_.wz[i];
      wz_curr = mppi_clampf(-s.constraints.wz, s.constraints.wz, wz_curr);
      wz_curr = mppi_clamp_velocity_by_accel(wz_last, wz_curr, -max_delta_wz, max_delta_wz);
      wz_last = wz_curr;
      if (isHolonomic()) {
        float & vy_curr = control_sequence_.vy[i];
        vy_curr = mppi_clampf(-s.constraints.vy, s.constraints.vy, vy_curr);
        vy_curr = mppi_clamp_velocity_by_accel(vy_last, vy_curr, min_delta_vy, max_delta_vy);
        vy_last = vy_curr;
      }
    }
    motion_model_.applyConstraints(control_sequence_);
  }

  void updateControlSequence()  // optimizer.cpp:605-645
  {
    const bool is_holo = isHolonomic();
    auto & s = settings_;
    const int B = s.batch_size, T = s.time_steps;

    auto addGamma = [&](const Tensor & cv, const std::vector<float> & u, float gamma_axis) {
        for (int b = 0; b < B; ++b) {
          float sum = 0.0f;
          for (int t = 0; t < T; ++t) {
            const float v = (cv(b, t) - u[t]) * u[t];
            sum = (t == 0) ? v : sum + v;
          }
          costs_[b] += gamma_axis * sum;
        }
      };
    const float gamma_vx = s.gamma / (s.std_vx * s.std_vx);
    addGamma(state_.cvx, control_sequence_.vx, gamma_vx);
    if (s.std_wz > 0.0f) {
      const float gamma_wz = s.gamma / (s.std_wz * s.std_wz);
      addGamma(state_.cwz, control_sequence_.wz, gamma_wz);
    }
    if (is_holo) {
      const float gamma_vy = s.gamma / (s.std_vy * s.std_vy);
      addGamma(state_.cvy, control_sequence_.vy, gamma_vy);
    }

    float min_cost = costs_[0];
    for (float c : costs_) {min_cost = std::min(min_cost, c);}
    const float inv_temp = 1.0f / s.temperature;
    softmaxes_.assign(B, 0.0f);
    // Sums over the batch are accumulated in double (SURVEY.md §8e "Accumulation order"): Eigen's
    // vectorised reduction order is not reproducible, and 2^20-term float sums drift.
    double sum = 0.0;
    for (int b = 0; b < B; ++b) {
      softmaxes_[b] = expf(-inv_temp * (costs_[b] - min_cost));
      sum += softmaxes_[b];
    }
    auto weighted = [&](const Tensor & cv, std::vector<float> & u) {
        for (int t = 0; t < T; ++t) {
          double acc = 0.0;
          const float * col = cv.col(t);
          for (int b = 0; b < B; ++b) {acc += static_cast<double>(col[b]) * softmaxes_[b];}
          u[t] = static_cast<float>(acc / sum);
        }
      };
    weighted(state_.cvx, control_sequence_.vx);
    weighted(state_.cwz, control_sequence_.wz);
    if (is_holo) {
      weighted(state_.cvy, control_sequence_.vy);
    }
    savitskyGolayFilter(control_sequence_, control_history_, s.sgf_order, s.shift_control_sequence);
    applyControlSequenceConstraints();
  }

  int optimize()  // optimizer.cpp:272-279
  {
    for (size_t i = 0; i < settings_.iteration_count; ++i) {
      generateNoisedTrajectories();
      CriticData data = makeCriticData(state_, generated_trajectories_);
      int rc = evalTrajectoriesScores(data);
      storeCriticData(data);
      if (rc != MPPI_OK) {return rc;}
      updateControlSequence();
    }
    return MPPI_OK;
  }

  // optimizer.cpp:281-298; returns 1 = retry, 0 = done, <0 = NoValidControl
  int fallback(bool fail)
  {
    if (!fail) {
      fallback_counter_ = 0;
      return 0;
    }
    reset(false);
    if (++fallback_counter_ > settings_.retry_attempt_limit) {
      fallback_counter_ = 0;
      return MPPI_ERR_NO_VALID_CONTROL;
    }
    return 1;
  }

  void shiftControlSequence()  // optimizer.cpp:345-357
  {
    auto size = control_sequence_.vx.size();
    shiftByOnePlace(control_sequence_.vx, -1);
    shiftByOnePlace(control_sequence_.wz, -1);
    control_sequence_.vx[size - 1] = control_sequence_.vx[size - 2];
    control_sequence_.wz[size - 1] = control_sequence_.wz[size - 2];
    if (isHolonomic()) {
      shiftByOnePlace(control_sequence_.vy, -1);
      control_sequence_.vy[size - 1] = control_sequence_.vy[size - 2];
    }
  }

  int evalControl(const MppiCycleIn & in, MppiCycleOut & out)  // optimizer.cpp:230-270
  {
    if (!costmap_.valid) {return MPPI_ERR_NOT_READY;}
    prepare(in);
    std::vector<float> optimal_trajectory;
    bool trajectory_valid = true;
    out.retries = 0;
    out.validation = MPPI_VALIDATION_SUCCESS;
    const unsigned int T = settings_.time_steps;
    while (true) {
      int rc = optimize();
      if (rc != MPPI_OK) {out.status = rc; return rc;}
      optimal_trajectory = getOptimizedTrajectory();
      int v = MPPI_VALIDATION_SUCCESS;
      if (cfg.validator_enabled) {
        v = validateTrajectory(
          optimal_trajectory, T, traj_samples_to_evaluate_, cfg.validator_consider_footprint != 0,
          costmap_, ros_);
      }
      out.validation = v;
      trajectory_valid = (v == MPPI_VALIDATION_SUCCESS);
      out.fail_flag = fail_flag_ ? 1 : 0;
      int fb = fallback(fail_flag_ || !trajectory_valid);
      if (fb < 0) {out.status = fb; return fb;}
      if (fb == 0) {break;}
      out.retries++;
    }
    // getControlFromSequenceAsTwist :647-664
    unsigned int offset = settings_.shift_control_sequence ? 1 : 0;
    float vx = control_sequence_.vx[offset];
    float wz = control_sequence_.wz[offset];
    float vy = isHolonomic() ? control_sequence_.vy[offset] : 0.0f;
    motion_model_.pushCommandHistory(vx, vy, wz);
    out.cmd_vx = vx; out.cmd_vy = vy; out.cmd_wz = wz;
    last_command_vel_.vx = vx;
    last_command_vel_.vy = isHolonomic() ? vy : 0.0;
    last_command_vel_.wz = wz;
    out.furthest_reached_path_point = furthest_ ? static_cast<int32_t>(*furthest_) : -1;
    if (out.control_vx) {std::memcpy(out.control_vx, control_sequence_.vx.data(), sizeof(float) * T);}
    if (out.control_vy) {std::memcpy(out.control_vy, control_sequence_.vy.data(), sizeof(float) * T);}
    if (out.control_wz) {std::memcpy(out.control_wz, control_sequence_.wz.data(), sizeof(float) * T);}
    if (out.optimal_trajectory) {
      std::memcpy(out.optimal_trajectory, optimal_trajectory.data(), sizeof(float) * 3 * T);
    }
    if (settings_.shift_control_sequence) {
      shiftControlSequence();
    }
    out.status = MPPI_OK;
    return MPPI_OK;
  }
};

}  // namespace oracle

// =====================================================================================
// C entry points (loaded with ctypes by tests/ and bench.py)
// =====================================================================================
using oracle::Optimizer;

extern "C" {

int oracle_create(const MppiConfig * cfg, void ** out)
{
  if (!cfg || !out) {return MPPI_ERR_INVALID_ARGUMENT;}
  auto * o = new Optimizer();
  int rc = o->configure(*cfg);
  if (rc != MPPI_OK) {delete o; *out = nullptr; return rc;}
  *out = o;
  return MPPI_OK;
}

int oracle_destroy(void * h)
{
  delete static_cast<Optimizer *>(h);
  return MPPI_OK;
}

int oracle_reset(void * h, int reset_dynamic_speed_limits)
{
  static_cast<Optimizer *>(h)->reset(reset_dynamic_speed_limits != 0);
  return MPPI_OK;
}

int oracle_set_speed_limit(void * h, double speed_limit, int percentage)
{
  static_cast<Optimizer *>(h)->setSpeedLimit(speed_limit, percentage != 0);
  return MPPI_OK;
}

int oracle_upload_costmap(
  void * h, const uint8_t * cells, uint32_t size_x, uint32_t size_y, double resolution,
  double origin_x, double origin_y)
{
  auto * o = static_cast<Optimizer *>(h);
  o->costmap_.cells.assign(cells, cells + static_cast<size_t>(size_x) * size_y);
  o->costmap_.size_x = size_x;
  o->costmap_.size_y = size_y;
  o->costmap_.resolution = resolution;
  o->costmap_.origin_x = origin_x;
  o->costmap_.origin_y = origin_y;
  const bool first = !o->costmap_.valid;
  o->costmap_.valid = true;
  if (first) {o->initCriticsCostmap();}
  return MPPI_OK;
}

int oracle_set_noise(void * h, const float * nvx, const float * nvy, const float * nwz)
{
  auto * o = static_cast<Optimizer *>(h);
  const int B = o->settings_.batch_size, T = o->settings_.time_steps;
  const size_t n = static_cast<size_t>(B) * T;
  o->noises_vx_.setZero(B, T); o->noises_vy_.setZero(B, T); o->noises_wz_.setZero(B, T);
  std::memcpy(o->noises_vx_.d.data(), nvx, n * sizeof(float));
  if (nvy) {std::memcpy(o->noises_vy_.d.data(), nvy, n * sizeof(float));}
  std::memcpy(o->noises_wz_.d.data(), nwz, n * sizeof(float));
  o->noise_injected_ = true;
  return MPPI_OK;
}

int oracle_optimize(void * h, const MppiCycleIn * in, MppiCycleOut * out)
{
  return static_cast<Optimizer *>(h)->evalControl(*in, *out);
}

// Times `cycles` back-to-back evalControl calls on the same inputs; returns seconds per call
// (steady_clock) in *sec_per_call.  Used by bench.py's cpu_baseline / --impl reference legs.
int oracle_time_optimize(void * h, const MppiCycleIn * in, int cycles, double * sec_per_call)
{
  auto * o = static_cast<Optimizer *>(h);
  MppiCycleOut out{};
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < cycles; ++i) {
    int rc = o->evalControl(*in, out);
    if (rc != MPPI_OK) {return rc;}
  }
  auto t1 = std::chrono::steady_clock::now();
  *sec_per_call = std::chrono::duration<double>(t1 - t0).count() / cycles;
  return MPPI_OK;
}

int oracle_score_tensors(
  void * h, const MppiCycleIn * in, const float * vx, const float * vy, const float * wz,
  const float * x, const float * y, const float * yaw, int32_t furthest_reached_path_point,
  float * costs, uint8_t * collisions, int32_t * fail_flag, int32_t * furthest_out)
{
  auto * o = static_cast<Optimizer *>(h);
  if (!o->costmap_.valid) {return MPPI_ERR_NOT_READY;}
  const int B = o->settings_.batch_size, T = o->settings_.time_steps;
  const size_t n = static_cast<size_t>(B) * T;
  o->prepare(*in);
  auto fill = [&](oracle::Tensor & t, const float * src) {
      t.setZero(B, T);
      if (src) {std::memcpy(t.d.data(), src, n * sizeof(float));}
    };
  fill(o->state_.vx, vx); fill(o->state_.vy, vy); fill(o->state_.wz, wz);
  fill(o->generated_trajectories_.x, x); fill(o->generated_trajectories_.y, y);
  fill(o->generated_trajectories_.yaws, yaw);
  std::memcpy(o->costs_.data(), costs, sizeof(float) * B);
  if (furthest_reached_path_point >= 0) {o->furthest_ = static_cast<size_t>(furthest_reached_path_point);}
  oracle::CriticData data = o->makeCriticData(o->state_, o->generated_trajectories_);
  if (collisions) {data.trajectories_in_collision.assign(B, false);}
  int rc = o->evalTrajectoriesScores(data);
  o->storeCriticData(data);
  std::memcpy(costs, o->costs_.data(), sizeof(float) * B);
  if (collisions) {
    for (int b = 0; b < B; ++b) {collisions[b] = data.trajectories_in_collision[b] ? 1 : 0;}
  }
  if (fail_flag) {*fail_flag = data.fail_flag ? 1 : 0;}
  if (furthest_out) {*furthest_out = o->furthest_ ? static_cast<int32_t>(*o->furthest_) : -1;}
  return rc;
}

int oracle_get_tensor(void * h, int32_t which, void * dst, size_t dst_bytes)
{
  auto * o = static_cast<Optimizer *>(h);
  const void * src = nullptr;
  size_t bytes = 0;
  std::vector<float> flat;
  std::vector<uint8_t> flags;
  auto T = [&](const oracle::Tensor & t) {src = t.d.data(); bytes = t.d.size() * sizeof(float);};
  switch (which) {
    case MPPI_TENSOR_VX: T(o->state_.vx); break;
    case MPPI_TENSOR_VY: T(o->state_.vy); break;
    case MPPI_TENSOR_WZ: T(o->state_.wz); break;
    case MPPI_TENSOR_CVX: T(o->state_.cvx); break;
    case MPPI_TENSOR_CVY: T(o->state_.cvy); break;
    case MPPI_TENSOR_CWZ: T(o->state_.cwz); break;
    case MPPI_TENSOR_X: T(o->generated_trajectories_.x); break;
    case MPPI_TENSOR_Y: T(o->generated_trajectories_.y); break;
    case MPPI_TENSOR_YAW: T(o->generated_trajectories_.yaws); break;
    case MPPI_TENSOR_NOISE_VX: T(o->noises_vx_); break;
    case MPPI_TENSOR_NOISE_VY: T(o->noises_vy_); break;
    case MPPI_TENSOR_NOISE_WZ: T(o->noises_wz_); break;
    case MPPI_TENSOR_COSTS: src = o->costs_.data(); bytes = o->costs_.size() * sizeof(float); break;
    case MPPI_TENSOR_CRITIC_COSTS:
      for (auto & v : o->critic_costs_) {flat.insert(flat.end(), v.begin(), v.end());}
      src = flat.data(); bytes = flat.size() * sizeof(float); break;
    case MPPI_TENSOR_COLLISIONS:
      flags.assign(o->settings_.batch_size, 0);
      for (size_t b = 0; b < o->collisions_.size(); ++b) {flags[b] = o->collisions_[b] ? 1 : 0;}
      src = flags.data(); bytes = flags.size(); break;
    case MPPI_TENSOR_COST_CELLS:
      src = o->cost_cells_.data(); bytes = o->cost_cells_.size() * sizeof(uint32_t); break;
    case MPPI_TENSOR_OBSTACLE_CELLS:
      src = o->obstacle_cells_.data(); bytes = o->obstacle_cells_.size() * sizeof(uint32_t); break;
    case MPPI_TENSOR_SOFTMAX:
      src = o->softmaxes_.data(); bytes = o->softmaxes_.size() * sizeof(float); break;
    default: return MPPI_ERR_INVALID_ARGUMENT;
  }
  if (bytes != dst_bytes) {return MPPI_ERR_INVALID_ARGUMENT;}
  if (bytes) {std::memcpy(dst, src, bytes);}
  return MPPI_OK;
}

int oracle_get_control_sequence(void * h, float * vx, float * vy, float * wz)
{
  auto * o = static_cast<Optimizer *>(h);
  const size_t T = o->settings_.time_steps;
  if (vx) {std::memcpy(vx, o->control_sequence_.vx.data(), T * sizeof(float));}
  if (vy) {std::memcpy(vy, o->control_sequence_.vy.data(), T * sizeof(float));}
  if (wz) {std::memcpy(wz, o->control_sequence_.wz.data(), T * sizeof(float));}
  return MPPI_OK;
}

int oracle_set_control_sequence(void * h, const float * vx, const float * vy, const float * wz)
{
  auto * o = static_cast<Optimizer *>(h);
  const size_t T = o->settings_.time_steps;
  if (vx) {std::memcpy(o->control_sequence_.vx.data(), vx, T * sizeof(float));}
  if (vy) {std::memcpy(o->control_sequence_.vy.data(), vy, T * sizeof(float));}
  if (wz) {std::memcpy(o->control_sequence_.wz.data(), wz, T * sizeof(float));}
  return MPPI_OK;
}

int oracle_get_state(void * h, MppiOptimizerState * state)
{
  auto * o = static_cast<Optimizer *>(h);
  std::memset(state, 0, sizeof(*state));
  state->last_command_vel[0] = o->last_command_vel_.vx;
  state->last_command_vel[1] = o->last_command_vel_.vy;
  state->last_command_vel[2] = o->last_command_vel_.wz;
  for (int i = 0; i < 4; ++i) {
    state->control_history[3 * i] = o->control_history_[i].vx;
    state->control_history[3 * i + 1] = o->control_history_[i].vy;
    state->control_history[3 * i + 2] = o->control_history_[i].wz;
  }
  const auto & m = o->motion_model_;
  for (size_t i = 0; i < m.cmd_history_vx_.size(); ++i) {state->cmd_history_vx[i] = m.cmd_history_vx_[i];}
  for (size_t i = 0; i < m.cmd_history_vy_.size(); ++i) {state->cmd_history_vy[i] = m.cmd_history_vy_[i];}
  for (size_t i = 0; i < m.cmd_history_wz_.size(); ++i) {state->cmd_history_wz[i] = m.cmd_history_wz_[i];}
  state->fallback_counter = static_cast<uint32_t>(o->fallback_counter_);
  return MPPI_OK;
}

int oracle_set_state(void * h, const MppiOptimizerState * state)
{
  auto * o = static_cast<Optimizer *>(h);
  o->last_command_vel_.vx = state->last_command_vel[0];
  o->last_command_vel_.vy = state->last_command_vel[1];
  o->last_command_vel_.wz = state->last_command_vel[2];
  for (int i = 0; i < 4; ++i) {
    o->control_history_[i] = {state->control_history[3 * i], state->control_history[3 * i + 1],
      state->control_history[3 * i + 2]};
  }
  auto & m = o->motion_model_;
  for (size_t i = 0; i < m.cmd_history_vx_.size(); ++i) {m.cmd_history_vx_[i] = state->cmd_history_vx[i];}
  for (size_t i = 0; i < m.cmd_history_vy_.size(); ++i) {m.cmd_history_vy_[i] = state->cmd_history_vy[i];}
  for (size_t i = 0; i < m.cmd_history_wz_.size(); ++i) {m.cmd_history_wz_[i] = state->cmd_history_wz[i];}
  o->fallback_counter_ = state->fallback_counter;
  return MPPI_OK;
}

const char * oracle_last_error(void * h)
{
  return static_cast<Optimizer *>(h)->error_.c_str();
}

// ----- unit-level entry points for the known-answer tests transcribed from the reference -----

// MotionModel::predict on caller tensors (in place), motion_models.hpp:108-167.
int oracle_predict(
  void * h, float * vx, float * vy, float * wz, float * cvx, float * cvy, float * cwz)
{
  auto * o = static_cast<Optimizer *>(h);
  const int B = o->settings_.batch_size, T = o->settings_.time_steps;
  const size_t n = static_cast<size_t>(B) * T;
  oracle::State & st = o->state_;
  st.reset(B, T);
  std::memcpy(st.vx.d.data(), vx, n * 4); std::memcpy(st.vy.d.data(), vy, n * 4);
  std::memcpy(st.wz.d.data(), wz, n * 4); std::memcpy(st.cvx.d.data(), cvx, n * 4);
  std::memcpy(st.cvy.d.data(), cvy, n * 4); std::memcpy(st.cwz.d.data(), cwz, n * 4);
  o->motion_model_.predict(st);
  std::memcpy(vx, st.vx.d.data(), n * 4); std::memcpy(vy, st.vy.d.data(), n * 4);
  std::memcpy(wz, st.wz.d.data(), n * 4); std::memcpy(cvx, st.cvx.d.data(), n * 4);
  std::memcpy(cvy, st.cvy.d.data(), n * 4); std::memcpy(cwz, st.cwz.d.data(), n * 4);
  return MPPI_OK;
}

int oracle_push_command_history(void * h, float vx, float vy, float wz)
{
  static_cast<Optimizer *>(h)->motion_model_.pushCommandHistory(vx, vy, wz);
  return MPPI_OK;
}

// Optimizer::integrateStateVelocities(Trajectories&, State&), optimizer.cpp:539-580.
int oracle_integrate(
  void * h, double pose_x, double pose_y, double pose_yaw, const float * vx, const float * vy,
  const float * wz, float * x, float * y, float * yaw)
{
  auto * o = static_cast<Optimizer *>(h);
  const int B = o->settings_.batch_size, T = o->settings_.time_steps;
  const size_t n = static_cast<size_t>(B) * T;
  oracle::State & st = o->state_;
  st.reset(B, T);
  st.pose_x = pose_x; st.pose_y = pose_y; st.pose_yaw = pose_yaw;
  std::memcpy(st.vx.d.data(), vx, n * 4);
  if (vy) {std::memcpy(st.vy.d.data(), vy, n * 4);}
  std::memcpy(st.wz.d.data(), wz, n * 4);
  o->generated_trajectories_.reset(B, T);
  o->integrateStateVelocities(o->generated_trajectories_, st);
  std::memcpy(x, o->generated_trajectories_.x.d.data(), n * 4);
  std::memcpy(y, o->generated_trajectories_.y.d.data(), n * 4);
  std::memcpy(yaw, o->generated_trajectories_.yaws.d.data(), n * 4);
  return MPPI_OK;
}

// applyControlSequenceConstraints on the current control sequence with a given state speed
// (optimizer.cpp:404-464); `inter_iteration` != 0 runs :368-402 instead.
int oracle_apply_constraints(void * h, double speed_vx, double speed_vy, double speed_wz, int inter_iteration)
{
  auto * o = static_cast<Optimizer *>(h);
  o->state_.speed = {speed_vx, speed_vy, speed_wz};
  if (inter_iteration) {o->applyControlSequenceInterIterationConstraints();} else {o->applyControlSequenceConstraints();}
  return MPPI_OK;
}

int oracle_shift_control_sequence(void * h)
{
  static_cast<Optimizer *>(h)->shiftControlSequence();
  return MPPI_OK;
}

// savitskyGolayFilter on the current control sequence with an explicit history (utils.hpp:460-542).
int oracle_savitsky_golay(void * h, float * history12 /* 4 x (vx,vy,wz), in/out */)
{
  auto * o = static_cast<Optimizer *>(h);
  std::array<oracle::Control, 4> hist;
  for (int i = 0; i < 4; ++i) {hist[i] = {history12[3 * i], history12[3 * i + 1], history12[3 * i + 2]};}
  oracle::savitskyGolayFilter(
    o->control_sequence_, hist, o->settings_.sgf_order, o->settings_.shift_control_sequence);
  for (int i = 0; i < 4; ++i) {
    history12[3 * i] = hist[i].vx; history12[3 * i + 1] = hist[i].vy; history12[3 * i + 2] = hist[i].wz;
  }
  return MPPI_OK;
}

int oracle_get_settings(void * h, float * out16)
{
  auto * o = static_cast<Optimizer *>(h);
  const auto & c = o->settings_.constraints;
  float v[16] = {c.vx_max, c.vx_min, c.vy, c.wz, c.ax_max, c.ax_min, c.ay_max, c.ay_min, c.az_max,
    o->settings_.shift_control_sequence ? 1.0f : 0.0f, o->settings_.controller_period,
    static_cast<float>(o->traj_samples_to_evaluate_), 0, 0, 0, 0};
  std::memcpy(out16, v, sizeof(v));
  return MPPI_OK;
}

float oracle_pose_point_angle(
  double pose_x, double pose_y, double pose_yaw, double point_x, double point_y,
  int forward_preference)
{
  return oracle::posePointAngle(pose_x, pose_y, pose_yaw, point_x, point_y, forward_preference != 0);
}

float oracle_pose_point_angle_yaw(
  double pose_x, double pose_y, double pose_yaw, double point_x, double point_y, double point_yaw)
{
  return oracle::posePointAngle(pose_x, pose_y, pose_yaw, point_x, point_y, point_yaw);
}

float oracle_normalize_angle(float a) {return mppi_normalize_anglef(a);}
float oracle_shortest_angular_distance(float from, float to) {return mppi_shortest_angular_distancef(from, to);}
double oracle_angles_normalize_angle(double a) {return mppi_angles_normalize_angle(a);}
void oracle_sincosf(float x, float * s, float * c) {mppi_sincosf(x, s, c);}

unsigned int oracle_find_closest_path_pt(const float * vec, unsigned int n, float dist, unsigned int init)
{
  std::vector<float> v(vec, vec + n);
  return oracle::findClosestPathPt(v, dist, init);
}

// Bresenham cells of a line (line_iterator.hpp); returns the number of cells written (<= cap).
int oracle_line_cells(int x0, int y0, int x1, int y1, int * xs, int * ys, int cap)
{
  int n = 0;
  for (oracle::LineIterator line(x0, y0, x1, y1); line.isValid(); line.advance()) {
    if (n < cap) {xs[n] = line.x_; ys[n] = line.y_;}
    ++n;
  }
  return n;
}

// FootprintCollisionChecker on an explicit costmap (footprint_collision_checker.cpp).
double oracle_footprint_cost_at_pose(
  const uint8_t * cells, uint32_t size_x, uint32_t size_y, double resolution, double origin_x,
  double origin_y, double x, double y, double theta, const double * fx, const double * fy, uint32_t n)
{
  oracle::Costmap cm;
  cm.cells.assign(cells, cells + static_cast<size_t>(size_x) * size_y);
  cm.size_x = size_x; cm.size_y = size_y; cm.resolution = resolution;
  cm.origin_x = origin_x; cm.origin_y = origin_y; cm.valid = true;
  oracle::FootprintCollisionChecker chk{&cm};
  std::vector<std::array<double, 2>> fp;
  for (uint32_t i = 0; i < n; ++i) {fp.push_back({fx[i], fy[i]});}
  return chk.footprintCostAtPose(x, y, theta, fp);
}

double oracle_line_cost(
  const uint8_t * cells, uint32_t size_x, uint32_t size_y, int x0, int x1, int y0, int y1)
{
  oracle::Costmap cm;
  cm.cells.assign(cells, cells + static_cast<size_t>(size_x) * size_y);
  cm.size_x = size_x; cm.size_y = size_y; cm.resolution = 1.0; cm.valid = true;
  oracle::FootprintCollisionChecker chk{&cm};
  return chk.lineCost(x0, x1, y0, y1);
}

uint8_t oracle_compute_cost(
  double distance_cells, double resolution, double inscribed_radius, double cost_scaling_factor)
{
  return oracle::computeCost(distance_cells, resolution, inscribed_radius, cost_scaling_factor);
}

int oracle_world_to_map(
  double wx, double wy, double origin_x, double origin_y, double resolution, uint32_t size_x,
  uint32_t size_y, uint32_t * mx, uint32_t * my)
{
  oracle::Costmap cm;
  cm.size_x = size_x; cm.size_y = size_y; cm.resolution = resolution;
  cm.origin_x = origin_x; cm.origin_y = origin_y;
  unsigned int a = 0, b = 0;
  bool ok = cm.worldToMap(wx, wy, a, b);
  *mx = a; *my = b;
  return ok ? 1 : 0;
}

// utils::findPathCosts (utils.hpp:358-387) on a bare costmap: valid[i] for the n - 1 path segments' first points
int oracle_find_path_costs(
  const uint8_t * cells, uint32_t size_x, uint32_t size_y, double resolution, double origin_x, double origin_y,
  const float * px, const float * py, int n, int tracking_unknown, uint8_t * valid)
{
  if (n < 1) {return 0;}
  oracle::Costmap cm;
  cm.cells.assign(cells, cells + static_cast<size_t>(size_x) * size_y);
  cm.size_x = size_x; cm.size_y = size_y; cm.resolution = resolution;
  cm.origin_x = origin_x; cm.origin_y = origin_y; cm.valid = true;
  oracle::Path path;
  path.x.assign(px, px + n); path.y.assign(py, py + n); path.yaws.assign(n, 0.0f);
  const std::vector<bool> v = oracle::pathPointsValid(path, cm, tracking_unknown != 0);
  for (size_t i = 0; i < v.size(); ++i) {valid[i] = v[i] ? 1 : 0;}
  return static_cast<int>(v.size());
}

}  // extern "C"
