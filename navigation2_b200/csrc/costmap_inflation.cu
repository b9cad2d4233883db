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
// (one thread per column), then one horizontal search per output cell (two cells per thread, packed add-then-min).  Tiles whose staged region
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
constexpr int THREADS = 256;        // CTA size for small grids (one tile's latency matters); large grids use 128
static_assert(COSTMAP_INFLATION_MAX_CELL_RADIUS * COSTMAP_INFLATION_MAX_CELL_RADIUS < 40000, "squared radius must stay below FAR_SQ");
constexpr unsigned FAR_SQ = 40000u;              // "no seed within R" in the uint16 squared distances: R^2 <= 14400 < FAR_SQ, FAR_SQ + R^2 < 65536
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
  int aligned_rows;  // master, map_stride and size_x are multiples of 4: rows can be read as 32-bit words
};

struct SmemPlan
{
  int raw_pitch, raw_rows, g_pitch;
  size_t off_raw, off_up, off_sq, off_dx, off_lut, total;
};

SmemPlan plan_smem(int radius)
{
  SmemPlan p;
  p.raw_pitch = (TILE_W + 2 * radius + 3 + 3) & ~3;  // + up to 3 bytes of alignment shift
  p.raw_rows = TILE_H + 2 * radius;
  p.g_pitch = p.raw_pitch;
  size_t o = 0;
  p.off_raw = o; o += static_cast<size_t>(p.raw_pitch) * p.raw_rows;
  p.off_up = o; o += static_cast<size_t>(p.g_pitch) * TILE_H;
  o = (o + 15) & ~static_cast<size_t>(15);
  p.off_sq = o; o += static_cast<size_t>(p.g_pitch) * (TILE_H / 2) * sizeof(uint32_t);
  p.off_dx = o; o += static_cast<size_t>(2 * radius + 1) * sizeof(uint32_t);
  p.off_lut = o; o += static_cast<size_t>(radius) * radius + 1;
  p.total = (o + 15) & ~static_cast<size_t>(15);
  return p;
}

// CR > 0: the cell radius is known at compile time (loops unroll, dx^2 become immediates); CR == 0: a.radius.
// NT: threads per CTA (a multiple of TILE_W).
template<int CR, int NT>
__global__ void __launch_bounds__(NT) k_inflate(const InflateArgs a, const SmemPlan sp)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_row_has_seed[TILE_H / 2];
  __shared__ int s_word_has_seed[(TILE_W + 2 * COSTMAP_INFLATION_MAX_CELL_RADIUS + 6) / 4 + 1];  // per 4 staged columns
  __shared__ int s_col_lo, s_col_hi;  // staged columns that hold a squared distance below FAR_SQ
  uint8_t * s_raw = smem + sp.off_raw;
  uint8_t * s_up = smem + sp.off_up;
  uint32_t * s_sq = reinterpret_cast<uint32_t *>(smem + sp.off_sq);
  uint32_t * s_dx = reinterpret_cast<uint32_t *>(smem + sp.off_dx);
  uint8_t * s_lut = smem + sp.off_lut;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = CR > 0 ? CR : a.radius;
  const int x0 = a.min_i + blockIdx.x * TILE_W, y0 = a.min_j + blockIdx.y * TILE_H;
  uint8_t * master = a.master + static_cast<size_t>(blockIdx.z) * a.map_stride;
  const int region_w = TILE_W + 2 * R, region_h = TILE_H + 2 * R;
  const bool around_unknown = a.inflate_around_unknown != 0;
  const unsigned seed_span = around_unknown ? 1u : 0u;  // seed <=> cell - 254 <= span (254, or 254 and 255)

  // ---- stage the tile and its halo (cells outside the map hold no seed) ----
  // With 4-byte aligned rows the region is fetched as 32-bit words starting at the aligned column `xa`; the
  // staged rows then begin `shift` bytes before the region's first column.
  const bool words = a.aligned_rows != 0;
  const int xa = words ? ((x0 - R) & ~3) : (x0 - R);
  const int shift = (x0 - R) - xa;
  unsigned seen_seed = 0;
  for (int i = tid; i < (region_w + 6) / 4 + 1; i += NT) {s_word_has_seed[i] = words ? 0 : 1;}
  if (tid < TILE_H / 2) {s_row_has_seed[tid] = 0;}
  if (tid == 0) {s_col_lo = region_w; s_col_hi = -1;}
  __syncthreads();
  if (words) {
    const int n_words = (region_w + shift + 3) >> 2;
    // a byte is a seed iff (byte & keep) == 0xFE: keep = 0xFF matches 254 only, keep = 0xFE matches 254 and 255
    const unsigned keep = around_unknown ? 0xFEFEFEFEu : 0xFFFFFFFFu;
    for (int w = lane; w < n_words; w += 32) {
      const int gx = xa + 4 * w;  // size_x is a multiple of 4: a word is inside the row or outside it
      const bool col_in = gx >= 0 && gx < a.size_x;
      int gy = y0 - R + warp;
      const uint8_t * src = master + static_cast<ptrdiff_t>(gy) * a.size_x + gx;
      uint32_t * dst = reinterpret_cast<uint32_t *>(s_raw + warp * sp.raw_pitch) + w;
      const size_t src_step = static_cast<size_t>(NT / 32) * a.size_x;
      const int dst_step = (NT / 32) * (sp.raw_pitch >> 2);
      unsigned word_seen = 0u;
#pragma unroll 4
      for (int ry = warp; ry < region_h; ry += NT / 32) {
        uint32_t cells = 0u;
        if (col_in && static_cast<unsigned>(gy) < static_cast<unsigned>(a.size_y)) {
          cells = __ldg(reinterpret_cast<const uint32_t *>(src));
        }
        *dst = cells;
        const unsigned v = (cells ^ 0xFEFEFEFEu) & keep;             // zero byte <=> seed
        word_seen |= (v - 0x01010101u) & ~v & 0x80808080u;           // non-zero iff some byte of v is zero
        gy += NT / 32; src += src_step; dst += dst_step;
      }
      seen_seed |= word_seen;
      if (word_seen != 0u) {s_word_has_seed[w] = 1;}
    }
  } else {
    for (int ry = warp; ry < region_h; ry += NT / 32) {
      const int gy = y0 - R + ry;
      const bool row_in = gy >= 0 && gy < a.size_y;
      const uint8_t * src = master + static_cast<size_t>(row_in ? gy : 0) * a.size_x;
      for (int rx = lane; rx < region_w; rx += 32) {
        const int gx = x0 - R + rx;
        uint8_t c = FREE_SPACE;
        if (row_in && gx >= 0 && gx < a.size_x) {c = __ldg(src + gx);}
        s_raw[ry * sp.raw_pitch + rx] = c;
        seen_seed |= (static_cast<unsigned>(c) - 254u <= seed_span) ? 1u : 0u;
      }
    }
  }
  if (!__syncthreads_or(seen_seed != 0u)) {return;}  // nothing within R of this tile: every cell is skipped (:259-261)
  for (int i = tid; i <= R * R; i += NT) {s_lut[i] = __ldg(a.cost_by_sq + i);}
  if (CR == 0) {
    for (int i = tid; i <= 2 * R; i += NT) {
      const unsigned dx_sq = static_cast<unsigned>((i - R) * (i - R));
      s_dx[i] = dx_sq | (dx_sq << 16);
    }
  }

  // ---- vertical pass, one thread per column: distance to the nearest seed below (up sweep over the halo below
  //      the tile, then the tile rows), then to the nearest seed above (down sweep), combined into squared column
  //      distances (FAR_SQ beyond R); tile rows r and r + 16 share one 32-bit word ----
  uint16_t * s_sq16 = reinterpret_cast<uint16_t *>(s_sq);
  // "no seed yet": with a compiled-in radius the sweeps stay far below 256 and need no clamp before the byte store
  const unsigned far_start = CR > 0 ? static_cast<unsigned>(CR) + 1u : 2u * COSTMAP_INFLATION_MAX_CELL_RADIUS + TILE_H + 8u;
  for (int col = tid; col < region_w; col += NT) {
    if (!s_word_has_seed[(col + shift) >> 2]) {  // no seed in these four staged columns: nothing to sweep
      uint16_t * sq = s_sq16 + (col << 1);
#pragma unroll 4
      for (int r = 0; r < TILE_H / 2; ++r) {
        sq[0] = static_cast<uint16_t>(FAR_SQ);
        sq[1] = static_cast<uint16_t>(FAR_SQ);
        sq += 2 * sp.g_pitch;
      }
      continue;
    }
    bool column_near = false;
    const uint8_t * cell = s_raw + (region_h - 1) * sp.raw_pitch + col + shift;
    unsigned dist = far_start;
    for (int i = 0; i < R; ++i) {  // halo below the tile
      dist = (static_cast<unsigned>(*cell) - 254u <= seed_span) ? 0u : dist + 1u;
      cell -= sp.raw_pitch;
    }
    uint8_t * up = s_up + (TILE_H - 1) * sp.g_pitch + col;
#pragma unroll 4
    for (int i = 0; i < TILE_H; ++i) {
      dist = (static_cast<unsigned>(*cell) - 254u <= seed_span) ? 0u : dist + 1u;
      *up = static_cast<uint8_t>(CR > 0 ? dist : min(dist, 255u));
      cell -= sp.raw_pitch;
      up -= sp.g_pitch;
    }
    cell = s_raw + col + shift;
    dist = far_start;
    for (int i = 0; i < R; ++i) {  // halo above the tile
      dist = (static_cast<unsigned>(*cell) - 254u <= seed_span) ? 0u : dist + 1u;
      cell += sp.raw_pitch;
    }
    up = s_up + col;
    // tile rows r and r + TILE_H / 2 are the two halves of word (r, col): one running pointer per half
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint16_t * sq = s_sq16 + (col << 1) + half;
      int * row_flag = s_row_has_seed;
#pragma unroll 4
      for (int r = 0; r < TILE_H / 2; ++r) {
        dist = (static_cast<unsigned>(*cell) - 254u <= seed_span) ? 0u : dist + 1u;
        const unsigned g = min(dist, static_cast<unsigned>(*up));
        const bool near = g <= static_cast<unsigned>(R);
        *sq = static_cast<uint16_t>(near ? g * g : FAR_SQ);
        if (near) {*row_flag = 1;}
        column_near |= near;
        cell += sp.raw_pitch;
        up += sp.g_pitch;
        sq += 2 * sp.g_pitch;
        ++row_flag;
      }
    }
    if (column_near) {
      atomicMin(&s_col_lo, col);
      atomicMax(&s_col_hi, col);
    }
  }
  __syncthreads();

  // ---- horizontal search (one add-then-min instruction per tap for two cells), cost lookup and the merge rule
  //      of applyInflation (:263-276) ----
  const int c = tid & (TILE_W - 1);
  const int gx = x0 + c;
  const unsigned r_sq = static_cast<unsigned>(R * R);
  const bool inflate_unknown = a.inflate_unknown != 0;
  // the taps of this cell are staged columns c .. c + 2R: none of them is near a seed -> every row of it is skipped
  if (gx >= a.max_i || c + 2 * R < s_col_lo || c > s_col_hi) {return;}
  for (int r = tid / TILE_W; r < TILE_H / 2; r += (NT / TILE_W)) {
    if (!s_row_has_seed[r]) {continue;}
    const uint32_t * row = s_sq + r * sp.g_pitch + c;  // row[k] is column gx - R + k
    unsigned best = FAR_SQ | (FAR_SQ << 16);
    if (CR > 0) {
#pragma unroll
      for (int k = 0; k <= 2 * CR; ++k) {
        const unsigned dx_sq = static_cast<unsigned>((k - CR) * (k - CR));
        best = __viaddmin_u16x2(row[k], dx_sq | (dx_sq << 16), best);  // per half: min(column distance^2 + dx^2, best)
      }
    } else {
#pragma unroll 4
      for (int k = 0; k <= 2 * R; ++k) {
        best = __viaddmin_u16x2(row[k], s_dx[k], best);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const unsigned sq = half ? best >> 16 : best & 0xffffu;
      const int tr = r + half * (TILE_H / 2);
      const int gy = y0 + tr;
      if (sq > r_sq || gy >= a.max_j) {continue;}
      const uint8_t cost = s_lut[sq];
      const uint8_t before = s_raw[(tr + R) * sp.raw_pitch + c + R + shift];
      const bool fill_unknown = before == NO_INFORMATION &&
        (inflate_unknown ? cost > FREE_SPACE : cost >= INSCRIBED_INFLATED_OBSTACLE);
      const uint8_t after = fill_unknown ? cost : max(before, cost);
      if (after != before) {master[static_cast<size_t>(gy) * a.size_x + gx] = after;}
    }
  }
}

// Small grids (fewer tiles than two waves of eight CTAs per SM could hold) run 256-thread CTAs: one tile's latency is
// what matters.  Large grids run 128-thread CTAs: the vertical pass occupies one thread per column (92 at R = 14),
// so smaller CTAs leave fewer warps idle at its barrier (8192 x 8192: 0.209 -> 0.187 ms, measured).
constexpr unsigned LARGE_GRID_TILES = 2u * 8u * 148u;

template<int CR>
cudaError_t launch_inflate(dim3 grid, size_t smem, cudaStream_t stream, const InflateArgs & a, const SmemPlan & sp)
{
  if (static_cast<size_t>(grid.x) * grid.y * grid.z >= LARGE_GRID_TILES) {
    k_inflate<CR, THREADS / 2><<<grid, THREADS / 2, smem, stream>>>(a, sp);
  } else {
    k_inflate<CR, THREADS><<<grid, THREADS, smem, stream>>>(a, sp);
  }
  return cudaGetLastError();
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
  a.aligned_rows = (reinterpret_cast<uintptr_t>(d_master) % 4 == 0 && map_stride % 4 == 0 && size_x % 4 == 0) ? 1 : 0;
  const dim3 grid((w.max_i - w.min_i + TILE_W - 1) / TILE_W, (w.max_j - w.min_j + TILE_H - 1) / TILE_H, n_maps);
  // the radii of nav2's defaults (0.55 m and 0.70 m at 0.05 m per cell) are compiled in; any other goes through a.radius
  switch (h->radius) {
    case 11: INFL_CU(h, launch_inflate<11>(grid, h->plan.total, stream, a, h->plan)); break;
    case 14: INFL_CU(h, launch_inflate<14>(grid, h->plan.total, stream, a, h->plan)); break;
    default: INFL_CU(h, launch_inflate<0>(grid, h->plan.total, stream, a, h->plan)); break;
  }
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
  e = cudaFuncSetAttribute(k_inflate<0, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->plan.total));
  if (e != cudaSuccess) {return fail(e, "shared memory");}
  e = cudaFuncSetAttribute(k_inflate<0, THREADS / 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->plan.total));
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
