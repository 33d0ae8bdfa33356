#include "occupancy_grid.h"
#include <algorithm>
#include <cmath>

OccGrid::OccGrid(const f110::Params& prm) : size_(prm.occ_size), discrete_(prm.occ_discrete), dilation_(prm.occ_dilation) {
  grid_blocks_ = size_ / discrete_;  // int <- int / float (occupancy_grid.cpp:9): 100 for 10 / 0.1f
  grid_.assign(static_cast<std::size_t>(grid_blocks_) * grid_blocks_, 0.f);
}

std::pair<int, int> OccGrid::WorldToOccupancy(float x, float y) const {
  // float arithmetic, then truncation toward zero (so an index in (-1, 0) lands on cell 0)
  const int col = (x - occ_offset_.first) / discrete_ + grid_blocks_ / 2;
  const int row = (y - occ_offset_.second) / discrete_ + grid_blocks_ / 2;
  return {col, row};
}

std::pair<float, float> OccGrid::OccupancyToWorld(int row, int col) const {
  return {discrete_ * (col - grid_blocks_ / 2) + occ_offset_.first, discrete_ * (row - grid_blocks_ / 2) + occ_offset_.second};
}

std::pair<float, float> OccGrid::PolarToCartesian(float range, float angle) const {
  return {range * std::cos(angle), range * std::sin(angle)};
}

bool OccGrid::InGrid(int col, int row) const { return col >= 0 && col < grid_blocks_ && row >= 0 && row < grid_blocks_; }

bool OccGrid::IsOccupied(float x_ind, float y_ind) const {
  // the reference indexes grid_(x_ind, y_ind) with float indices: first = row, second = column
  return grid_[static_cast<std::size_t>(static_cast<long>(x_ind)) + static_cast<std::size_t>(static_cast<long>(y_ind)) * grid_blocks_] != 0.f;
}

void OccGrid::FillOccGrid(const geometry_msgs::Pose& pose, const sensor_msgs::LaserScan& scan) {
  std::fill(grid_.begin(), grid_.end(), 0.f);
  const auto& o = pose.orientation;
  const float yaw = std::atan2(2 * o.w * o.z, 1 - 2 * o.z * o.z);
  // the grid is centred 0.275 m ahead of the car (occupancy_grid.cpp:63-64): double arithmetic, stored as float
  occ_offset_.first = pose.position.x + 0.275 * std::cos(yaw);
  occ_offset_.second = pose.position.y + 0.275 * std::sin(yaw);
  int beams = (scan.angle_max - scan.angle_min) / scan.angle_increment + 1;
  beams = std::min(beams, static_cast<int>(scan.ranges.size()));
  for (int i = 0; i < beams; ++i) {
    const float bearing = scan.angle_min + i * scan.angle_increment + yaw;
    std::pair<float, float> hit = PolarToCartesian(scan.ranges[i], bearing);
    hit.first += occ_offset_.first;
    hit.second += occ_offset_.second;
    // dilation stamp: float loop counters, exactly as the reference iterates them
    for (float dx = -dilation_; dx <= dilation_; dx += discrete_)
      for (float dy = -dilation_; dy <= dilation_; dy += discrete_) {
        const std::pair<int, int> cell = WorldToOccupancy(hit.first + dx, hit.second + dy);
        if (InGrid(cell)) grid_[static_cast<std::size_t>(cell.second) + static_cast<std::size_t>(cell.first) * grid_blocks_] = 1.f;
      }
  }
}
