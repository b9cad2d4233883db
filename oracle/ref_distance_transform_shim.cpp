// Thin C wrapper around the REFERENCE's own Felzenszwalb-Huttenlocher distance transform, compiled from
// where it lies (-I/root/reference/nav2_costmap_2d/include, with oracle/ref_shims standing in for the absent
// <Eigen/Core>; see Makefile target `ref`).  No reference source is copied into this repository.
// Test infrastructure: tests/test_oracle_inflation.py pins oracle/inflation_oracle.cpp against it.
#include <cstdint>

#include "nav2_costmap_2d/distance_transform.hpp"

extern "C" void ref_distance_map(const uint8_t * obstacle_mask, int height, int width, float * out)
{
  nav2_costmap_2d::MatrixXfRM img(height, width);
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      img(y, x) = obstacle_mask[static_cast<size_t>(y) * width + x] ? 0.0f : nav2_costmap_2d::DistanceTransform::DT_INF;
    }
  }
  nav2_costmap_2d::DistanceTransform::distanceTransform2D(img, height, width);
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {out[static_cast<size_t>(y) * width + x] = img(y, x);}
  }
}
