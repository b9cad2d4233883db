/*
 * C ABI of the B200 costmap inflation pass (the map is produced on the device).
 *
 * Replaces, for the master grid of a nav2_costmap_2d::LayeredCostmap, what
 * nav2_costmap_2d::InflationLayer does in
 *   updateCosts      nav2_costmap_2d/plugins/inflation_layer.cpp:283-348  (seed mask, distance transform)
 *   applyInflation   nav2_costmap_2d/plugins/inflation_layer.cpp:237-279  (distance -> cost, merge rule)
 *   computeCaches    nav2_costmap_2d/plugins/inflation_layer.cpp:351-366  (distance -> cost table)
 *   computeCost      nav2_costmap_2d/include/nav2_costmap_2d/inflation_layer.hpp:121-135
 * Results are bit-identical to the reference's for maps up to 2048 cells on a side (beyond that the reference's
 * float parabola intersections may themselves deviate from the exact Euclidean transform by an ulp; this
 * library always produces the exact one).  Plain pointers and sizes only; every call returns 0 or a negative
 * status; no CPU fallback: without a CUDA device `costmap_inflation_create` fails.
 */
#ifndef COSTMAP_INFLATION_B200_H_
#define COSTMAP_INFLATION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COSTMAP_INFLATION_OK 0
#define COSTMAP_INFLATION_ERR_INVALID_ARGUMENT (-1)
#define COSTMAP_INFLATION_ERR_CUDA (-2)
#define COSTMAP_INFLATION_ERR_RADIUS_TOO_LARGE (-3)

/* largest cell_inflation_radius the kernel stages in shared memory (inflation_radius / resolution, rounded up) */
#define COSTMAP_INFLATION_MAX_CELL_RADIUS 120

/* the parameters InflationLayer reads (inflation_layer.cpp:82-100) plus what it derives from the layered
 * costmap: resolution (matchSize, :139-146) and the footprint's inscribed radius (onFootprintChanged, :207-219) */
typedef struct CostmapInflationConfig
{
  double resolution;            /* m per cell */
  double inflation_radius;      /* m; cells = ceil(inflation_radius / resolution) (costmap_2d.cpp:254-258) */
  double inscribed_radius;      /* m */
  double cost_scaling_factor;
  int32_t inflate_unknown;
  int32_t inflate_around_unknown;
} CostmapInflationConfig;

typedef struct CostmapInflation CostmapInflation;

int costmap_inflation_create(const CostmapInflationConfig * config, CostmapInflation ** out);
int costmap_inflation_destroy(CostmapInflation * h);
const char * costmap_inflation_last_error(const CostmapInflation * h);

/* cell_inflation_radius_ (inflation_layer.cpp:144) */
uint32_t costmap_inflation_cell_radius(const CostmapInflation * h);
/* cost_lut_ (inflation_layer.cpp:360-365): copies min(cap, length) entries, returns the length */
int costmap_inflation_lut(const CostmapInflation * h, uint8_t * dst, int cap);

/* InflationLayer::updateCosts(master_grid, min_i, min_j, max_i, max_j) on a master grid in HOST memory
 * (row-major uint8, index = my * size_x + mx, costmap_2d.hpp:231-234), updated in place: the rows the window
 * and its padding touch go up with one copy, one kernel runs, the window's rows come back with one copy. */
int costmap_inflation_update_costs(
  CostmapInflation * h, uint8_t * master, uint32_t size_x, uint32_t size_y,
  int32_t min_i, int32_t min_j, int32_t max_i, int32_t max_j);

/* The same on `n_maps` grids already in DEVICE memory (`map_stride` bytes apart, each size_x * size_y), in place,
 * enqueued on `cuda_stream` (a cudaStream_t; NULL = the object's own stream) without synchronising. */
int costmap_inflation_update_costs_device(
  CostmapInflation * h, uint8_t * device_master, uint32_t size_x, uint32_t size_y, uint32_t n_maps,
  size_t map_stride, int32_t min_i, int32_t min_j, int32_t max_i, int32_t max_j, void * cuda_stream);

/* Measurement: `steps` whole-map passes over `n_maps` device-resident grids, each pass restored from a pristine
 * copy first (the restore is outside the timed events), after `warmup` untimed passes and with `flush_bytes` of
 * L2 overwritten before every timed pass (0 = none).  Writes the mean kernel time per pass in milliseconds. */
int costmap_inflation_time_resident(
  CostmapInflation * h, const uint8_t * host_maps, uint32_t size_x, uint32_t size_y, uint32_t n_maps,
  int steps, int warmup, size_t flush_bytes, float * ms_per_pass);

/* kernels this object has launched so far */
uint64_t costmap_inflation_launch_count(const CostmapInflation * h);

#ifdef __cplusplus
}
#endif

#endif  /* COSTMAP_INFLATION_B200_H_ */
