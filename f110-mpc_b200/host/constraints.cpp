#include "constraints.h"
#include <cmath>

Constraints::Constraints(const f110::Params& prm)
    : d_(prm.state_lims), ftg_thresh_(prm.follow_gap_thresh), umax_val_(prm.umax), umin_val_(prm.umin),
      divider_(prm.fov_divider), buffer_(prm.buffer) {
  x_max_ = f110::Vector{INFTY, INFTY, INFTY};
  x_min_ = f110::Vector{-INFTY, -INFTY, -INFTY};
  u_max_ = f110::Vector{umax_val_, 0.43f};   // speed, steering: floats widened into the vector (constraints.cpp:18-19)
  u_min_ = f110::Vector{umin_val_, -0.43f};  // constraints.cpp:20-21
  l1_ = f110::Vector(3);
  l2_ = f110::Vector(3);
}

void Constraints::SetXLims(State s) {
  x_max_(0) = s.x() + d_; x_max_(1) = s.y() + d_;
  x_min_(0) = s.x() - d_; x_min_(1) = s.y() - d_;
}

bool Constraints::FindHalfSpaces(State& state, sensor_msgs::LaserScan& scan) {
  // number of beams from the float expression the reference uses (constraints.cpp:118)
  int beams = (scan.angle_max - scan.angle_min) / scan.angle_increment + 1;
  if (beams > static_cast<int>(scan.ranges.size())) beams = static_cast<int>(scan.ranges.size());
  const float half_fov = 1.571f / divider_;
  // Run-length scan for the widest run of "far" beams.  Kept literally: `hi` survives the end of a run, so a
  // new run is credited the stale hi until its second beam, and a one-beam run never registers (SURVEY a13').
  int widest = -1, lo = -1, hi = -1;
  bool inside = false;
  best_lo_ = 0; best_hi_ = 0;
  for (int i = 0; i < beams; ++i) {
    const float bearing = scan.angle_min + i * scan.angle_increment;
    if (!(bearing > -half_fov && bearing < half_fov)) continue;
    if (scan.ranges[i] > ftg_thresh_) {
      if (inside) hi = i; else { lo = i; inside = true; }
    } else {
      inside = false;
      if (hi - lo > widest) { widest = hi - lo; best_hi_ = hi; best_lo_ = lo; }
    }
    if (hi - lo > widest) { widest = hi - lo; best_hi_ = hi; best_lo_ = lo; }
  }
  if (best_hi_ - best_lo_ > 2 * buffer_) {  // shrink wide gaps by `buffer` beams per side (int <- float arithmetic)
    best_hi_ = best_hi_ - buffer_;
    best_lo_ = best_lo_ + buffer_;
  }
  const int nr = static_cast<int>(scan.ranges.size());
  if (best_lo_ < 0 || best_hi_ < 0 || best_lo_ >= nr || best_hi_ >= nr) return false;

  const double px = state.x(), py = state.y();
  const float heading = state.ori();
  const float ang_lo = scan.angle_min + best_lo_ * scan.angle_increment + heading;
  const float ang_hi = scan.angle_min + best_hi_ * scan.angle_increment + heading;
  // float * cosf(float) + double, narrowed into the float pair (constraints.cpp:182-189)
  p1_.first = scan.ranges[best_lo_] * std::cos(ang_lo) + px;
  p1_.second = scan.ranges[best_lo_] * std::sin(ang_lo) + py;
  p2_.first = scan.ranges[best_hi_] * std::cos(ang_hi) + px;
  p2_.second = scan.ranges[best_hi_] * std::sin(ang_hi) + py;
  p_.first = px;
  p_.second = py;
  // line through (p, edge) as a x + b y + c = 0, oriented so the OTHER edge lies on the >= 0 side; float math
  auto line_through = [this](const std::pair<float, float>& edge, const std::pair<float, float>& other, f110::Vector& out) {
    float a = p_.second - edge.second;
    float b = edge.first - p_.first;
    float c = p_.first * edge.second - p_.second * edge.first;
    if (a * other.first + b * other.second + c < 0) { a = -a; b = -b; c = -c; }
    out(0) = a; out(1) = b; out(2) = c + 0.5;
  };
  line_through(p1_, p2_, l1_);
  line_through(p2_, p1_, l2_);
  return true;
}
