// Costmap inflation on the device: C ABI of include/costmap_inflation_b200.h and its kernel (sm_100a).
//
// What the reference does (nav2_costmap_2d/plugins/inflation_layer.cpp:283-348 + :237-279): mark LETHAL cells
// (and NO_INFORMATION ones with inflate_around_unknown) as seeds, run an exact Euclidean distance transform
// over the update window grown by the inflation radius, and for every cell within `cell_inflation_radius_`
// cells of a seed look the cost up in a distance table and merge it into the master grid.
//
// The device formulation keeps every quantity an integer, so the result is exact by construction:
//   * the squared distance of a cell to its nearest seed is  min over dx of  dx^2 + g(x + dx, y)^2  where g is the
//     distance to the nearest seed in the same column; seeds further than R = cell_inflation_radius_ never matter
//     (such cells are skipped, inflation_layer.cpp:259-261: sqrtf is monotone and sqrtf(R^2) == R), so both searches
//     stop at R and a tile only needs a halo of R cells;
//   * distance -> cost goes through a table indexed by the SQUARED distance, built on the host from the reference's
//     own chain  sqrtf -> d * 100 + 0.5f -> cost_lut_  (inflation_layer.cpp:266-269), so the kernel has no float math.
// One CTA produces a 64 x 32 tile: stage (64 + 2R) x (32 + 2R) raw cells in shared memory, two vertical sweeps
// (one thread per column and direction), then one horizontal search per output cell.  Tiles whose staged region
// holds no seed (most of a sparse map) stop after the load and write nothing.
//
// In place is safe without a second buffer: a cell is written only by the CTA that owns it, after that CTA read it,
// and the update never changes whether a cell is a seed (a cost of 254 is only looked up at distance 0, i.e. for a
// seed; with inflate_around_unknown an unknown seed becomes 254, still a seed), so halo reads racing with a
// neighbour's writes see the same seeds either way.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "costmap_inflation_b200.h"

namespace
{

constexpr int TILE_W = 64;
constexpr int TILE_H = 32;
constexpr int THREADS = 256;
constexpr int ROWS_PER_PASS = THREADS / TILE_W;  // 4
constexpr unsigned FAR_SQ = 60000u;              // "no seed within R" in the uint16 squared-distance rows (R^2 <= 14400)
constexpr int LUT_PRECISION = 100;               // InflationLayer::COST_LUT_PRECISION, inflation_layer.hpp:231
constexpr uint8_t NO_INFORMATION = 255, LETHAL_OBSTACLE = 254, INSCRIBED_INFLATED_OBSTACLE = 253, FREE_SPACE = 0;

struct InflateArgs
{
  uint8_t * master;
  size_t map_stride;
  const uint8_t * cost_by_sq;  // [R*R + 1]
  int size_x, size_y;
  int min_i, min_j, max_i, max_j;  // clamped update window
  int radius;
  int inflate_unknown, inflate_around_unknown;
};

struct SmemPlan
{
  int raw_pitch, raw_rows, g_pitch;
  size_t off_raw, off_down, off_up, off_sq, off_lut, total;
};

SmemPlan plan_smem(int radius)
{
  SmemPlan p;
  p.raw_pitch = (TILE_W + 2 * radius + 3) & ~3;
  p.raw_rows = TILE_H + 2 * radius;
  p.g_pitch = p.raw_pitch;
  size_t o = 0;
  p.off_raw = o; o += static_cast<size_t>(p.raw_pitch) * p.raw_rows;
  p.off_down = o; o += static_cast<size_t>(p.g_pitch) * TILE_H;
  p.off_up = o; o += static_cast<size_t>(p.g_pitch) * TILE_H;
  o = (o + 15) & ~static_cast<size_t>(15);
  p.off_sq = o; o += static_cast<size_t>(p.g_pitch) * TILE_H * sizeof(uint16_t);
  p.off_lut = o; o += static_cast<size_t>(radius) * radius + 1;
  p.total = (o + 15) & ~static_cast<size_t>(15);
  return p;
}

__global__ void __launch_bounds__(THREADS) k_inflate(const InflateArgs a, const SmemPlan sp)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_row_has_seed[TILE_H];
  uint8_t * s_raw = smem + sp.off_raw;
  uint8_t * s_down = smem + sp.off_down;
  uint8_t * s_up = smem + sp.off_up;
  uint16_t * s_sq = reinterpret_cast<uint16_t *>(smem + sp.off_sq);
  uint8_t * s_lut = smem + sp.off_lut;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = a.radius;
  const int x0 = a.min_i + blockIdx.x * TILE_W, y0 = a.min_j + blockIdx.y * TILE_H;
  uint8_t * master = a.master + static_cast<size_t>(blockIdx.z) * a.map_stride;
  const int region_w = TILE_W + 2 * R, region_h = TILE_H + 2 * R;
  const bool around_unknown = a.inflate_around_unknown != 0;

  // ---- stage the tile and its halo (cells outside the map hold no seed) ----
  int seen_seed = 0;
  for (int ry = warp; ry < region_h; ry += THREADS / 32) {
    const int gy = y0 - R + ry;
    const bool row_in = gy >= 0 && gy < a.size_y;
    const uint8_t * src = master + static_cast<size_t>(row_in ? gy : 0) * a.size_x;
    for (int rx = lane; rx < region_w; rx += 32) {
      const int gx = x0 - R + rx;
      uint8_t c = FREE_SPACE;
      if (row_in && gx >= 0 && gx < a.size_x) {c = __ldg(src + gx);}
      s_raw[ry * sp.raw_pitch + rx] = c;
      seen_seed |= (c == LETHAL_OBSTACLE) | (around_unknown & (c == NO_INFORMATION));
    }
  }
  if (tid < TILE_H) {s_row_has_seed[tid] = 0;}
  if (!__syncthreads_or(seen_seed)) {return;}  // nothing within R of this tile: every cell is skipped (:259-261)
  for (int i = tid; i <= R * R; i += THREADS) {s_lut[i] = __ldg(a.cost_by_sq + i);}

  // ---- vertical sweeps: distance to the nearest seed above (down sweep) and below (up sweep), capped at 255 ----
  for (int job = tid; job < 2 * region_w; job += THREADS) {
    const bool upward = job >= region_w;
    const int col = upward ? job - region_w : job;
    uint8_t * dst = upward ? s_up : s_down;
    unsigned dist = 255u;
    for (int step = 0; step < region_h; ++step) {
      const int y = upward ? region_h - 1 - step : step;
      const uint8_t c = s_raw[y * sp.raw_pitch + col];
      const bool seed = (c == LETHAL_OBSTACLE) | (around_unknown & (c == NO_INFORMATION));
      dist = seed ? 0u : min(dist + 1u, 255u);
      const int out_row = y - R;
      if (out_row >= 0 && out_row < TILE_H) {dst[out_row * sp.g_pitch + col] = static_cast<uint8_t>(dist);}
    }
  }
  __syncthreads();

  // ---- squared column distances (FAR_SQ beyond R) and which tile rows can reach a seed at all ----
  for (int r = warp; r < TILE_H; r += THREADS / 32) {
    int any = 0;
    for (int col = lane; col < region_w; col += 32) {
      const unsigned g = min(static_cast<unsigned>(s_down[r * sp.g_pitch + col]), static_cast<unsigned>(s_up[r * sp.g_pitch + col]));
      const bool near = g <= static_cast<unsigned>(R);
      s_sq[r * sp.g_pitch + col] = static_cast<uint16_t>(near ? g * g : FAR_SQ);
      any |= near ? 1 : 0;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {s_row_has_seed[r] = any;}
  }
  __syncthreads();

  // ---- horizontal search, cost lookup and the merge rule of applyInflation (:263-276) ----
  const int c = tid & (TILE_W - 1);
  const int gx = x0 + c;
  const unsigned r_sq = static_cast<unsigned>(R * R);
  const bool inflate_unknown = a.inflate_unknown != 0;
  for (int r = tid / TILE_W; r < TILE_H; r += ROWS_PER_PASS) {
    const int gy = y0 + r;
    if (!s_row_has_seed[r] || gx >= a.max_i || gy >= a.max_j) {continue;}
    const uint16_t * row = s_sq + r * sp.g_pitch + c;  // row[k] is column gx - R + k
    unsigned best = FAR_SQ;
    unsigned dx_sq = r_sq;
    int delta = -(2 * R - 1);  // (dx + 1)^2 - dx^2 at dx = -R
#pragma unroll 4
    for (int k = 0; k <= 2 * R; ++k) {
      best = min(best, static_cast<unsigned>(row[k]) + dx_sq);
      dx_sq += delta;
      delta += 2;
    }
    if (best > r_sq) {continue;}
    const uint8_t cost = s_lut[best];
    const uint8_t before = s_raw[(r + R) * sp.raw_pitch + c + R];
    const bool fill_unknown = before == NO_INFORMATION &&
      (inflate_unknown ? cost > FREE_SPACE : cost >= INSCRIBED_INFLATED_OBSTACLE);
    const uint8_t after = fill_unknown ? cost : max(before, cost);
    if (after != before) {master[static_cast<size_t>(gy) * a.size_x + gx] = after;}
  }
}

}  // namespace

struct CostmapInflation
{
  CostmapInflationConfig cfg{};
  uint32_t radius = 0;
  std::vector<uint8_t> lut;         // cost_lut_
  std::vector<uint8_t> cost_by_sq;  // [R*R + 1]
  uint8_t * d_cost_by_sq = nullptr;
  cudaStream_t stream = nullptr;
  uint8_t * d_rows = nullptr;       // device copy of the rows a host update touches
  uint8_t * h_rows = nullptr;       // pinned staging
  size_t rows_cap = 0;
  SmemPlan plan{};
  uint64_t launches = 0;
  std::string error;
};

namespace
{

#define INFL_CU(h, call) \
  do { \
    const cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) { \
      (h)->error = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return COSTMAP_INFLATION_ERR_CUDA; \
    } \
  } while (0)

// inflation_layer.hpp:121-135
uint8_t compute_cost(const CostmapInflationConfig & c, double distance_cells)
{
  if (distance_cells == 0) {return LETHAL_OBSTACLE;}
  if (distance_cells * c.resolution <= c.inscribed_radius) {return INSCRIBED_INFLATED_OBSTACLE;}
  const double factor = std::exp(-1.0 * c.cost_scaling_factor * (distance_cells * c.resolution - c.inscribed_radius));
  return static_cast<uint8_t>((INSCRIBED_INFLATED_OBSTACLE - 1) * factor);
}

struct Window {int min_i, min_j, max_i, max_j;};

bool clamp_window(uint32_t size_x, uint32_t size_y, int min_i, int min_j, int max_i, int max_j, Window & w)
{
  w.min_i = std::max(0, min_i);
  w.min_j = std::max(0, min_j);
  w.max_i = std::min(static_cast<int>(size_x), max_i);
  w.max_j = std::min(static_cast<int>(size_y), max_j);
  return w.max_i > w.min_i && w.max_j > w.min_j;
}

int launch(
  CostmapInflation * h, uint8_t * d_master, uint32_t size_x, uint32_t size_y, uint32_t n_maps, size_t map_stride,
  const Window & w, cudaStream_t stream)
{
  InflateArgs a;
  a.master = d_master;
  a.map_stride = map_stride;
  a.cost_by_sq = h->d_cost_by_sq;
  a.size_x = static_cast<int>(size_x);
  a.size_y = static_cast<int>(size_y);
  a.min_i = w.min_i; a.min_j = w.min_j; a.max_i = w.max_i; a.max_j = w.max_j;
  a.radius = static_cast<int>(h->radius);
  a.inflate_unknown = h->cfg.inflate_unknown;
  a.inflate_around_unknown = h->cfg.inflate_around_unknown;
  const dim3 grid((w.max_i - w.min_i + TILE_W - 1) / TILE_W, (w.max_j - w.min_j + TILE_H - 1) / TILE_H, n_maps);
  k_inflate<<<grid, THREADS, h->plan.total, stream>>>(a, h->plan);
  INFL_CU(h, cudaGetLastError());
  ++h->launches;
  return COSTMAP_INFLATION_OK;
}

}  // namespace

extern "C" {

const char * costmap_inflation_last_error(const CostmapInflation * h) {return h ? h->error.c_str() : "null handle";}

int costmap_inflation_create(const CostmapInflationConfig * config, CostmapInflation ** out)
{
  if (!config || !out || !(config->resolution > 0.0) || !(config->inflation_radius >= 0.0)) {
    return COSTMAP_INFLATION_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  // costmap_2d.cpp:254-258 cellDistance
  const double cells = std::max(0.0, std::ceil(config->inflation_radius / config->resolution));
  if (cells > COSTMAP_INFLATION_MAX_CELL_RADIUS) {return COSTMAP_INFLATION_ERR_RADIUS_TOO_LARGE;}
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {return COSTMAP_INFLATION_ERR_CUDA;}  // no CPU path

  CostmapInflation * h = new CostmapInflation;
  h->cfg = *config;
  h->radius = static_cast<uint32_t>(cells);
  const uint32_t R = h->radius;
  if (R > 0) {
    // computeCaches, inflation_layer.cpp:351-366
    const uint32_t last = R * LUT_PRECISION + 1;
    h->lut.resize(last + 1);
    for (uint32_t i = 0; i <= last; ++i) {h->lut[i] = compute_cost(h->cfg, static_cast<double>(i) / LUT_PRECISION);}
    // applyInflation's distance -> table slot chain (:266-269) evaluated once per possible squared distance
    h->cost_by_sq.resize(static_cast<size_t>(R) * R + 1);
    for (uint32_t sq = 0; sq <= R * R; ++sq) {
      const float distance_cells = std::sqrt(static_cast<float>(sq));
      const float scaled = distance_cells * static_cast<float>(LUT_PRECISION);
      const uint32_t slot = std::min(last, static_cast<uint32_t>(scaled + 0.5f));
      h->cost_by_sq[sq] = h->lut[slot];
    }
  } else {
    h->cost_by_sq.assign(1, LETHAL_OBSTACLE);
  }
  h->plan = plan_smem(static_cast<int>(R));
  auto fail = [&](cudaError_t e, const char * what) {
      (void)e; (void)what;
      costmap_inflation_destroy(h);
      return COSTMAP_INFLATION_ERR_CUDA;
    };
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {return fail(e, "stream");}
  e = cudaMalloc(&h->d_cost_by_sq, h->cost_by_sq.size());
  if (e != cudaSuccess) {return fail(e, "table");}
  e = cudaMemcpy(h->d_cost_by_sq, h->cost_by_sq.data(), h->cost_by_sq.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {return fail(e, "table copy");}
  e = cudaFuncSetAttribute(k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->plan.total));
  if (e != cudaSuccess) {return fail(e, "shared memory");}
  *out = h;
  return COSTMAP_INFLATION_OK;
}

int costmap_inflation_destroy(CostmapInflation * h)
{
  if (!h) {return COSTMAP_INFLATION_ERR_INVALID_ARGUMENT;}
  if (h->stream) {cudaStreamSynchronize(h->stream);}
  if (h->d_cost_by_sq) {cudaFree(h->d_cost_by_sq);}
  if (h->d_rows) {cudaFree(h->d_rows);}
  if (h->h_rows) {cudaFreeHost(h->h_rows);}
  if (h->stream) {cudaStreamDestroy(h->stream);}
  delete h;
  return COSTMAP_INFLATION_OK;
}

uint32_t costmap_inflation_cell_radius(const CostmapInflation * h) {return h ? h->radius : 0;}

int costmap_inflation_lut(const CostmapInflation * h, uint8_t * dst, int cap)
{
  if (!h || (!dst && cap > 0)) {return COSTMAP_INFLATION_ERR_INVALID_ARGUMENT;}
  const int n = static_cast<int>(h->lut.size());
  if (cap > 0 && n > 0) {std::memcpy(dst, h->lut.data(), static_cast<size_t>(std::min(n, cap)));}
  return n;
}

uint64_t costmap_inflation_launch_count(const CostmapInflation * h) {return h ? h->launches : 0;}

int costmap_inflation_update_costs_device(
  CostmapInflation * h, uint8_t * device_master, uint32_t size_x, uint32_t size_y, uint32_t n_maps,
  size_t map_stride, int32_t min_i, int32_t min_j, int32_t max_i, int32_t max_j, void * cuda_stream)
{
  if (!h || !device_master || size_x == 0 || size_y == 0 || n_maps == 0 || n_maps > 65535u ||
    (n_maps > 1 && map_stride < static_cast<size_t>(size_x) * size_y))
  {
    if (h) {h->error = "bad arguments";}
    return COSTMAP_INFLATION_ERR_INVALID_ARGUMENT;
  }
  if (h->radius == 0) {return COSTMAP_INFLATION_OK;}  // inflation_layer.cpp:289
  Window w;
  if (!clamp_window(size_x, size_y, min_i, min_j, max_i, max_j, w)) {return COSTMAP_INFLATION_OK;}
  return launch(h, device_master, size_x, size_y, n_maps, map_stride, w,
           cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream);
}

int costmap_inflation_update_costs(
  CostmapInflation * h, uint8_t * master, uint32_t size_x, uint32_t size_y,
  int32_t min_i, int32_t min_j, int32_t max_i, int32_t max_j)
{
  if (!h || !master || size_x == 0 || size_y == 0) {
    if (h) {h->error = "bad arguments";}
    return COSTMAP_INFLATION_ERR_INVALID_ARGUMENT;
  }
  if (h->radius == 0) {return COSTMAP_INFLATION_OK;}
  Window w;
  if (!clamp_window(size_x, size_y, min_i, min_j, max_i, max_j, w)) {return COSTMAP_INFLATION_OK;}
  // whole rows of the padded region go up; the kernel addresses them as a map of `n_rows` rows
  const int pad = static_cast<int>(h->radius);
  const int row0 = std::max(0, w.min_j - pad), row1 = std::min(static_cast<int>(size_y), w.max_j + pad);
  const size_t n_rows = static_cast<size_t>(row1 - row0), bytes = n_rows * size_x;
  if (bytes > h->rows_cap) {
    if (h->d_rows) {cudaFree(h->d_rows); h->d_rows = nullptr;}
    if (h->h_rows) {cudaFreeHost(h->h_rows); h->h_rows = nullptr;}
    h->rows_cap = 0;
    INFL_CU(h, cudaMalloc(&h->d_rows, bytes));
    INFL_CU(h, cudaMallocHost(&h->h_rows, bytes));
    h->rows_cap = bytes;
  }
  std::memcpy(h->h_rows, master + static_cast<size_t>(row0) * size_x, bytes);
  INFL_CU(h, cudaMemcpyAsync(h->d_rows, h->h_rows, bytes, cudaMemcpyHostToDevice, h->stream));
  Window local = w;
  local.min_j -= row0;
  local.max_j -= row0;
  const int rc = launch(h, h->d_rows, size_x, static_cast<uint32_t>(n_rows), 1, bytes, local, h->stream);
  if (rc != COSTMAP_INFLATION_OK) {return rc;}
  const size_t back_off = static_cast<size_t>(local.min_j) * size_x;
  const size_t back_bytes = static_cast<size_t>(local.max_j - local.min_j) * size_x;
  INFL_CU(h, cudaMemcpyAsync(h->h_rows + back_off, h->d_rows + back_off, back_bytes, cudaMemcpyDeviceToHost, h->stream));
  INFL_CU(h, cudaStreamSynchronize(h->stream));
  for (int j = w.min_j; j < w.max_j; ++j) {
    std::memcpy(master + static_cast<size_t>(j) * size_x + w.min_i,
      h->h_rows + static_cast<size_t>(j - row0) * size_x + w.min_i, static_cast<size_t>(w.max_i - w.min_i));
  }
  return COSTMAP_INFLATION_OK;
}

int costmap_inflation_time_resident(
  CostmapInflation * h, const uint8_t * host_maps, uint32_t size_x, uint32_t size_y, uint32_t n_maps,
  int steps, int warmup, size_t flush_bytes, float * ms_per_pass)
{
  if (!h || !host_maps || !ms_per_pass || size_x == 0 || size_y == 0 || n_maps == 0 || steps <= 0 || warmup < 0) {
    if (h) {h->error = "bad arguments";}
    return COSTMAP_INFLATION_ERR_INVALID_ARGUMENT;
  }
  const size_t stride = static_cast<size_t>(size_x) * size_y, bytes = stride * n_maps;
  uint8_t * d_pristine = nullptr, * d_work = nullptr, * d_flush = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
      if (d_pristine) {cudaFree(d_pristine);}
      if (d_work) {cudaFree(d_work);}
      if (d_flush) {cudaFree(d_flush);}
      if (e0) {cudaEventDestroy(e0);}
      if (e1) {cudaEventDestroy(e1);}
    };
#define INFL_TRY(call) \
  do { \
    const cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) { \
      h->error = std::string(#call) + ": " + cudaGetErrorString(e_); \
      cleanup(); \
      return COSTMAP_INFLATION_ERR_CUDA; \
    } \
  } while (0)
  INFL_TRY(cudaMalloc(&d_pristine, bytes));
  INFL_TRY(cudaMalloc(&d_work, bytes));
  if (flush_bytes) {INFL_TRY(cudaMalloc(&d_flush, flush_bytes));}
  INFL_TRY(cudaEventCreate(&e0));
  INFL_TRY(cudaEventCreate(&e1));
  INFL_TRY(cudaMemcpy(d_pristine, host_maps, bytes, cudaMemcpyHostToDevice));
  const Window w{0, 0, static_cast<int>(size_x), static_cast<int>(size_y)};
  double total_ms = 0.0;
  for (int it = 0; it < warmup + steps; ++it) {
    INFL_TRY(cudaMemcpyAsync(d_work, d_pristine, bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (d_flush) {INFL_TRY(cudaMemsetAsync(d_flush, it & 0xff, flush_bytes, h->stream));}
    INFL_TRY(cudaEventRecord(e0, h->stream));
    if (h->radius > 0) {
      const int rc = launch(h, d_work, size_x, size_y, n_maps, stride, w, h->stream);
      if (rc != COSTMAP_INFLATION_OK) {cleanup(); return rc;}
    }
    INFL_TRY(cudaEventRecord(e1, h->stream));
    INFL_TRY(cudaEventSynchronize(e1));
    float ms = 0.0f;
    INFL_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (it >= warmup) {total_ms += ms;}
  }
#undef INFL_TRY
  cleanup();
  *ms_per_pass = static_cast<float>(total_ms / steps);
  return COSTMAP_INFLATION_OK;
}

}  // extern "C"
