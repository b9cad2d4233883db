// inflation_layer.hpp — C++ host mirror of nav2_costmap_2d::InflationLayer's cost pass on top of the C ABI
// (include/costmap_inflation_b200.h).  Same member names and meaning as the reference
// (nav2_costmap_2d/include/nav2_costmap_2d/inflation_layer.hpp:77-245, plugins/inflation_layer.cpp): parameters
// `enabled`, `inflation_radius`, `cost_scaling_factor`, `inflate_unknown`, `inflate_around_unknown` (:82-100),
// matchSize() (:139-146), onFootprintChanged() (:207-219), updateCosts() (:283-348), computeCost() (:121-135 of the
// header), cellDistance() (:183-186).  What is not mirrored: the Layer base class plumbing (updateBounds, dynamic
// parameter callbacks, the layered-costmap registration), which stays the reference's own code in a ROS 2 build.
#ifndef NAV2_COSTMAP_2D_B200__INFLATION_LAYER_HPP_
#define NAV2_COSTMAP_2D_B200__INFLATION_LAYER_HPP_

#include <algorithm>
#include <cmath>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "costmap_inflation_b200.h"
#include "ros_shim/ros_shim.hpp"

namespace nav2_costmap_2d
{

class InflationLayer
{
public:
  // inflation_layer.cpp:82-100 (defaults 0.55 m, 10.0, false, false)
  struct Parameters
  {
    bool enabled{true};
    double inflation_radius{0.55};
    double cost_scaling_factor{10.0};
    bool inflate_unknown{false};
    bool inflate_around_unknown{false};
  };

  InflationLayer() = default;
  explicit InflationLayer(const Parameters & p) : params_(p) {}
  ~InflationLayer() {release();}
  InflationLayer(const InflationLayer &) = delete;
  InflationLayer & operator=(const InflationLayer &) = delete;

  // matchSize (:139-146): resolution of the master grid, then the caches
  void matchSize(const Costmap2D & master_grid)
  {
    std::lock_guard<Costmap2D::mutex_t> guard(mutex_);
    resolution_ = master_grid.getResolution();
    cell_inflation_radius_ = cellDistance(params_.inflation_radius);
    computeCaches();
  }

  // onFootprintChanged (:207-219): the layered costmap's inscribed radius, then the caches
  void onFootprintChanged(double inscribed_radius)
  {
    std::lock_guard<Costmap2D::mutex_t> guard(mutex_);
    inscribed_radius_ = inscribed_radius;
    cell_inflation_radius_ = cellDistance(params_.inflation_radius);
    computeCaches();
  }

  // updateCosts (:283-348), on the device
  void updateCosts(Costmap2D & master_grid, int min_i, int min_j, int max_i, int max_j)
  {
    std::lock_guard<Costmap2D::mutex_t> guard(mutex_);
    if (!params_.enabled || cell_inflation_radius_ == 0) {return;}
    if (!device_) {throw std::runtime_error("InflationLayer: matchSize() has not been called");}
    const int rc = costmap_inflation_update_costs(
      device_, master_grid.getCharMap(), master_grid.getSizeInCellsX(), master_grid.getSizeInCellsY(),
      min_i, min_j, max_i, max_j);
    if (rc != COSTMAP_INFLATION_OK) {
      throw std::runtime_error(std::string("InflationLayer: ") + costmap_inflation_last_error(device_));
    }
  }

  // computeCost (inflation_layer.hpp:121-135)
  unsigned char computeCost(double distance) const
  {
    if (distance == 0) {return LETHAL_OBSTACLE;}
    if (distance * resolution_ <= inscribed_radius_) {return INSCRIBED_INFLATED_OBSTACLE;}
    const double factor = std::exp(-1.0 * params_.cost_scaling_factor * (distance * resolution_ - inscribed_radius_));
    return static_cast<unsigned char>((INSCRIBED_INFLATED_OBSTACLE - 1) * factor);
  }

  // cellDistance (inflation_layer.hpp:183-186 -> costmap_2d.cpp:254-258)
  unsigned int cellDistance(double world_dist) const
  {
    return static_cast<unsigned int>(std::max(0.0, std::ceil(world_dist / resolution_)));
  }

  double getCostScalingFactor() const {return params_.cost_scaling_factor;}
  double getInflationRadius() const {return params_.inflation_radius;}
  unsigned int getCellInflationRadius() const {return cell_inflation_radius_;}
  Costmap2D::mutex_t * getMutex() {return &mutex_;}

private:
  // computeCaches (:351-366): the device object owns cost_lut_
  void computeCaches()
  {
    release();
    if (cell_inflation_radius_ == 0) {return;}
    CostmapInflationConfig c{resolution_, params_.inflation_radius, inscribed_radius_, params_.cost_scaling_factor,
      params_.inflate_unknown ? 1 : 0, params_.inflate_around_unknown ? 1 : 0};
    const int rc = costmap_inflation_create(&c, &device_);
    if (rc == COSTMAP_INFLATION_ERR_RADIUS_TOO_LARGE) {
      throw std::runtime_error("InflationLayer: inflation radius above 120 cells is not supported on the device");
    }
    if (rc != COSTMAP_INFLATION_OK) {
      throw std::runtime_error("InflationLayer: no CUDA device (there is no CPU path)");
    }
    if (costmap_inflation_cell_radius(device_) != cell_inflation_radius_) {
      throw std::runtime_error("InflationLayer: cell radius mismatch between host and device");
    }
  }
  void release()
  {
    if (device_) {costmap_inflation_destroy(device_); device_ = nullptr;}
  }

  Parameters params_;
  double resolution_{0.05};
  double inscribed_radius_{0.0};
  unsigned int cell_inflation_radius_{0};
  CostmapInflation * device_{nullptr};
  Costmap2D::mutex_t mutex_;
};

}  // namespace nav2_costmap_2d

#endif  // NAV2_COSTMAP_2D_B200__INFLATION_LAYER_HPP_
