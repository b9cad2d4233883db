// CPU ORACLE of the costmap inflation pass.  TEST INFRASTRUCTURE, not product code.
//
// Restates, function by function, what nav2_costmap_2d's InflationLayer does to the master grid in
// updateCosts(), so that tests/ can check the CUDA kernel (navigation2_b200/csrc/costmap_inflation.cu)
// bit for bit.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
//
//   reference                                                          here
//   include/nav2_costmap_2d/distance_transform.hpp:66-104  (1-D pass)   lower_envelope_1d
//   include/nav2_costmap_2d/distance_transform.hpp:120-176 (2-D pass)   euclidean_distance_map
//   include/nav2_costmap_2d/inflation_layer.hpp:121-135    computeCost  cost_at_distance
//   plugins/inflation_layer.cpp:351-366                    computeCaches  distance_lut
//   plugins/inflation_layer.cpp:237-279                    applyInflation  (second half of update_costs)
//   plugins/inflation_layer.cpp:283-348                    updateCosts  update_costs
//   src/costmap_2d.cpp:254-258                             cellDistance  cell_distance
//
// Pinning: tests/test_oracle_inflation.py transcribes the closed-form expectations of
// nav2_costmap_2d/test/integration/inflation_tests.cpp and, when oracle/_ref/libref_distance_transform.so
// exists (built from the reference's own header where it lies, Makefile target `ref`), compares
// euclidean_distance_map with the reference's DistanceTransform::distanceTransform2D on random masks.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

namespace
{

constexpr float kFar = FLT_MAX;          // DistanceTransform::DT_INF, distance_transform.hpp:51
constexpr int kLutPrecision = 100;       // InflationLayer::COST_LUT_PRECISION, inflation_layer.hpp:231
constexpr uint8_t kNoInformation = 255;  // cost_values.hpp:70-74
constexpr uint8_t kLethal = 254;
constexpr uint8_t kInscribed = 253;
constexpr uint8_t kFree = 0;

// where the parabolas rooted at p and q (q > p) cross: float arithmetic in the reference's order
inline float crossing(const float * f, int p, int q)
{
  return (f[q] - f[p] + static_cast<float>(q * q - p * p)) / (2.0f * static_cast<float>(q - p));
}

// distance_transform.hpp:66-104: squared 1-D distance transform of the sampled function f
void lower_envelope_1d(const float * f, float * d, int n, std::vector<int> & root, std::vector<float> & from)
{
  if (n <= 0) {return;}
  int top = 0;
  root[0] = 0;
  from[0] = -kFar;
  from[1] = kFar;
  for (int q = 1; q < n; ++q) {
    float s = crossing(f, root[top], q);
    while (s <= from[top]) {
      --top;
      s = crossing(f, root[top], q);
    }
    ++top;
    root[top] = q;
    from[top] = s;
    from[top + 1] = kFar;
  }
  int seg = 0;
  for (int q = 0; q < n; ++q) {
    while (from[seg + 1] < static_cast<float>(q)) {++seg;}
    const int off = q - root[seg];
    d[q] = static_cast<float>(off * off) + f[root[seg]];
  }
}

// distance_transform.hpp:120-176: columns, then rows, then the element-wise square root.  `img` is row-major.
void euclidean_distance_map(std::vector<float> & img, int height, int width)
{
  const int longest = std::max(height, width);
  std::vector<float> f(longest), d(longest), from(longest + 1);
  std::vector<int> root(longest);
  for (int x = 0; x < width; ++x) {
    for (int y = 0; y < height; ++y) {f[y] = img[static_cast<size_t>(y) * width + x];}
    lower_envelope_1d(f.data(), d.data(), height, root, from);
    for (int y = 0; y < height; ++y) {img[static_cast<size_t>(y) * width + x] = d[y];}
  }
  for (int y = 0; y < height; ++y) {
    float * row = img.data() + static_cast<size_t>(y) * width;
    for (int x = 0; x < width; ++x) {f[x] = row[x];}
    lower_envelope_1d(f.data(), d.data(), width, root, from);
    for (int x = 0; x < width; ++x) {row[x] = d[x];}
  }
  for (float & v : img) {v = std::sqrt(v);}
}

// inflation_layer.hpp:121-135
uint8_t cost_at_distance(double distance_cells, double resolution, double inscribed_radius, double cost_scaling_factor)
{
  if (distance_cells == 0) {return kLethal;}
  if (distance_cells * resolution <= inscribed_radius) {return kInscribed;}
  const double factor = std::exp(-1.0 * cost_scaling_factor * (distance_cells * resolution - inscribed_radius));
  return static_cast<uint8_t>((kInscribed - 1) * factor);
}

// costmap_2d.cpp:254-258
unsigned int cell_distance(double world_dist, double resolution)
{
  return static_cast<unsigned int>(std::max(0.0, std::ceil(world_dist / resolution)));
}

// inflation_layer.cpp:351-366
std::vector<uint8_t> distance_lut(
  unsigned int cell_radius, double resolution, double inscribed_radius, double cost_scaling_factor)
{
  std::vector<uint8_t> lut;
  if (cell_radius == 0) {return lut;}
  const unsigned int last = cell_radius * kLutPrecision + 1;
  lut.resize(last + 1);
  for (unsigned int i = 0; i <= last; ++i) {
    lut[i] = cost_at_distance(static_cast<double>(i) / kLutPrecision, resolution, inscribed_radius, cost_scaling_factor);
  }
  return lut;
}

}  // namespace

extern "C" {

struct OracleInflationConfig
{
  double resolution, inflation_radius, inscribed_radius, cost_scaling_factor;
  int32_t inflate_unknown, inflate_around_unknown;
};

unsigned int oracle_inflation_cell_radius(const OracleInflationConfig * c)
{
  return cell_distance(c->inflation_radius, c->resolution);
}

// the distance -> cost table of computeCaches(); returns its length (0 when the cell radius is 0)
int oracle_inflation_lut(const OracleInflationConfig * c, uint8_t * dst, int cap)
{
  const std::vector<uint8_t> lut = distance_lut(
    oracle_inflation_cell_radius(c), c->resolution, c->inscribed_radius, c->cost_scaling_factor);
  for (int i = 0; i < static_cast<int>(lut.size()) && i < cap; ++i) {dst[i] = lut[i];}
  return static_cast<int>(lut.size());
}

// Euclidean distance (in cells) of every cell of a row-major 0/1 mask to the nearest 1
void oracle_distance_map(const uint8_t * obstacle_mask, int height, int width, float * out)
{
  std::vector<float> img(static_cast<size_t>(height) * width);
  for (size_t i = 0; i < img.size(); ++i) {img[i] = obstacle_mask[i] ? 0.0f : kFar;}
  euclidean_distance_map(img, height, width);
  std::copy(img.begin(), img.end(), out);
}

// InflationLayer::updateCosts (inflation_layer.cpp:283-348) followed by applyInflation (:237-279), in place
int oracle_inflation_update_costs(
  const OracleInflationConfig * c, uint8_t * master, unsigned int size_x, unsigned int size_y,
  int min_i, int min_j, int max_i, int max_j)
{
  const unsigned int cell_radius = oracle_inflation_cell_radius(c);
  if (cell_radius == 0) {return 0;}  // :289
  const std::vector<uint8_t> lut = distance_lut(cell_radius, c->resolution, c->inscribed_radius, c->cost_scaling_factor);

  min_i = std::max(0, min_i);
  min_j = std::max(0, min_j);
  max_i = std::min(static_cast<int>(size_x), max_i);
  max_j = std::min(static_cast<int>(size_y), max_j);

  // region of interest: the update window grown by the inflation radius (:303-311)
  const int pad = static_cast<int>(cell_radius);
  const int roi_min_i = std::max(0, min_i - pad), roi_min_j = std::max(0, min_j - pad);
  const int roi_max_i = std::min(static_cast<int>(size_x), max_i + pad);
  const int roi_max_j = std::min(static_cast<int>(size_y), max_j + pad);
  const int roi_w = roi_max_i - roi_min_i, roi_h = roi_max_j - roi_min_j;
  if (roi_w <= 0 || roi_h <= 0) {return 0;}

  std::vector<float> dist(static_cast<size_t>(roi_h) * roi_w);
  for (int y = 0; y < roi_h; ++y) {
    for (int x = 0; x < roi_w; ++x) {
      const uint8_t cell = master[static_cast<size_t>(y + roi_min_j) * size_x + (x + roi_min_i)];
      const bool seed = cell == kLethal || (c->inflate_around_unknown && cell == kNoInformation);  // :326-334
      dist[static_cast<size_t>(y) * roi_w + x] = seed ? 0.0f : kFar;
    }
  }
  euclidean_distance_map(dist, roi_h, roi_w);

  const float radius_f = static_cast<float>(cell_radius);
  const unsigned int lut_last = static_cast<unsigned int>(lut.size() - 1);
  for (int j = min_j; j < max_j; ++j) {
    for (int i = min_i; i < max_i; ++i) {
      const float d = dist[static_cast<size_t>(j - roi_min_j) * roi_w + (i - roi_min_i)];
      if (d > radius_f) {continue;}
      const size_t idx = static_cast<size_t>(j) * size_x + i;
      const uint8_t before = master[idx];
      const unsigned int slot = std::min(lut_last, static_cast<unsigned int>(d * kLutPrecision + 0.5f));
      const uint8_t cost = lut[slot];
      const bool fill_unknown = before == kNoInformation && (c->inflate_unknown ? cost > kFree : cost >= kInscribed);
      master[idx] = fill_unknown ? cost : std::max(before, cost);
    }
  }
  return 0;
}

// best-of-`repeats` wall time of update_costs over the whole map, restoring the input between runs
double oracle_inflation_time(
  const OracleInflationConfig * c, const uint8_t * master, unsigned int size_x, unsigned int size_y, int repeats);

}  // extern "C"

#include <chrono>
double oracle_inflation_time(
  const OracleInflationConfig * c, const uint8_t * master, unsigned int size_x, unsigned int size_y, int repeats)
{
  std::vector<uint8_t> work(static_cast<size_t>(size_x) * size_y);
  double best = 1e30;
  for (int r = 0; r < repeats; ++r) {
    std::copy(master, master + work.size(), work.begin());
    const auto t0 = std::chrono::steady_clock::now();
    oracle_inflation_update_costs(c, work.data(), size_x, size_y, 0, 0, static_cast<int>(size_x), static_cast<int>(size_y));
    best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  return best;
}
