// mppi_kernels.h — launch interface between the host runtime (mppi_capi.cu) and mppi_kernels.cu.
#ifndef NAVIGATION2_B200_CSRC_MPPI_KERNELS_H_
#define NAVIGATION2_B200_CSRC_MPPI_KERNELS_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "mppi_device.h"

// dynamic shared-memory layout of the rollout+score kernel (byte offsets)
struct KSmemA
{
  uint32_t off_u;      // float u[3][T]
  uint32_t off_path;   // float path x / y / integrated distance, path_cap each
  uint32_t off_lut;    // float obstacle distance LUT [2][256]
  uint32_t off_cc_lut; // float CostCritic score per cell cost [256] (k_rollout_lean)
  uint32_t off_ring;   // float delay ring [3][ring_depth][block]
  uint32_t off_win;    // uint8 costmap window
  uint32_t off_noise;  // float noise ring [noise_stages][axes][MPPI_TCHUNK_A][block], filled by cp.async.bulk
  int32_t noise_stages;  // 0 = read noise straight from global memory
  int32_t path_cap;
  int32_t ring_depth;
  uint32_t total;
};

struct KSmemB
{
  int32_t path_cap;
  uint32_t off_pa;     // float [pa_stride][block]: the block's PathAlign sample rows, transposed
  uint32_t total;
};

// dynamic shared-memory layout of the fused one-cluster-per-robot kernel (byte offsets)
struct KSmemF
{
  uint32_t off_arr;    // float [5 or 6][T][MPPI_FUSED_G]: vx, wz, yaw, x, y (, vy)
  uint32_t off_u;      // float u[3][T]
  uint32_t off_path;   // float path x / y / yaw / integrated distance + uint8 validity, path_cap each
  uint32_t off_terms;  // float [n_terms + 2][MPPI_FUSED_G]
  uint32_t off_w;      // float softmax weights [MPPI_FUSED_G]
  uint32_t off_xch;    // FusedExchange + float W[3][T]: read by the other CTAs through DSMEM
  uint32_t off_lut;    // float obstacle distance LUT [2][256]
  uint32_t off_fin;    // finalize scratch, 11 T floats
  uint32_t off_pa;     // PathAlign scratch [2][n_pa][MPPI_FUSED_G] (n_pa <= T / 4 + 1 at the smallest sampling step the fused path takes)
  uint32_t off_win;    // uint8 costmap window
  int32_t path_cap;
  int32_t has_lut;   // the obstacle LUT is allocated (ObstaclesCritic configured)
  uint32_t total;
};

KSmemF mppi_smem_layout_f(int T, bool holo, int path_cap, int n_terms, bool obstacles, int win_bytes, int n_pa);
cudaError_t mppi_launch_fused(const KArgs & a, const KSmemF & sm, bool holo, int cluster_size, cudaStream_t stream);
bool mppi_fused_supported(const KSmemF & sm, bool holo, int cluster_size);

KSmemA mppi_smem_layout_a(
  int T, int path_cap, bool obstacles, int ring_depth, int block, int win_bytes, int noise_axes, int noise_stages);

// what the host knows about this launch; selects the kernel instantiation
struct KPlanA
{
  int block;             // 64 (latency mode, small batches) or MPPI_BLOCK_A
  bool holo;
  bool from_tensors;     // score_tensors: state / trajectories come from caller tensors
  bool full;             // materialize / debug cells / collision flags / clamp_raw_controls / input delay
  uint32_t crit_types;   // union of CRIT bits (1 << MppiCriticType) of the critics active this call
  int cc_step, pa_step;  // trajectory_point_step of CostCritic / PathAlignCritic (0 if absent)
  int model;             // MppiMotionModel
  bool lean_ok;          // nothing in the configuration rules out k_rollout_lean (point collision checks in CostCritic,
                         // no path orientations in PathAlign, costmap resolution inside exact_div's range)
};

cudaError_t mppi_launch_rollout_score(
  const KArgs & a, const KSmemA & sm, const KPlanA & plan, cudaStream_t stream);
cudaError_t mppi_launch_path_score(
  const KArgs & a, int path_cap, bool holo, bool from_tensors, int block, cudaStream_t stream);
cudaError_t mppi_launch_softmax_partial(const KArgs & a, bool holo, cudaStream_t stream);
// shard_stage: 0 single GPU, 1 reduce partials into this rank's slot, 2 combine slots + finalize,
// 3 reduce + exchange over peer memory (KArgs::xch_*) + combine + finalize in one launch
cudaError_t mppi_launch_finalize(const KArgs & a, bool holo, int shard_stage, cudaStream_t stream);
cudaError_t mppi_launch_philox(
  float * nvx, float * nvy, float * nwz, int B, int T, int R, int b_offset, uint64_t seed,
  uint32_t epoch, float std_vx, float std_vy, float std_wz, bool holo, cudaStream_t stream);

cudaError_t mppi_launch_selftest_exact_math(
  const float * a, const float * b, int n, unsigned int * mismatches, unsigned int * flagged, cudaStream_t stream);

#endif  // NAVIGATION2_B200_CSRC_MPPI_KERNELS_H_
