// OccGrid — laser-scan occupancy grid (reference include/f110-mpc/occupancy_grid.h:13-47).  The grid is a
// column-major float matrix like the reference's Eigen::MatrixXf so `data()` can be handed to the device
// collision check unchanged.
#pragma once
#include <utility>
#include <vector>
#include "msgs.h"

class OccGrid {
 public:
  explicit OccGrid(const f110::Params& params);
  virtual ~OccGrid() = default;

  std::pair<int, int> WorldToOccupancy(std::pair<float, float> point) const { return WorldToOccupancy(point.first, point.second); }
  std::pair<int, int> WorldToOccupancy(float x, float y) const;  // returns (col, row)
  std::pair<float, float> OccupancyToWorld(int row, int col) const;
  std::pair<float, float> OccupancyToWorld(std::pair<int, int> grid_point) const { return OccupancyToWorld(grid_point.second, grid_point.first); }
  std::pair<float, float> PolarToCartesian(float range, float angle) const;
  bool IsOccupied(float x_ind, float y_ind) const;
  void FillOccGrid(const geometry_msgs::Pose& pose_msg, const sensor_msgs::LaserScan& scan_msg);
  bool InGrid(int col, int row) const;
  bool InGrid(std::pair<int, int> grid_point) const { return InGrid(grid_point.first, grid_point.second); }
  bool CartesianInGrid(float x, float y) const { return InGrid(WorldToOccupancy(x, y)); }
  bool CartesianInGrid(std::pair<float, float> p) const { return CartesianInGrid(p.first, p.second); }
  int size() const { return size_; }

  // device hand-off
  const float* data() const { return grid_.data(); }
  int blocks() const { return grid_blocks_; }
  float discrete() const { return discrete_; }
  std::pair<float, float> offset() const { return occ_offset_; }

 private:
  int size_;
  float discrete_;
  int grid_blocks_;
  float dilation_;
  std::pair<float, float> occ_offset_{0.f, 0.f};
  std::vector<float> grid_;  // cell (row, col) at row + col * grid_blocks_
};
