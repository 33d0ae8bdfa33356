#include "model.h"
#include <cmath>

namespace {
const double kRolloutWheelbase = 0.35;  // CAR_LENGTH of the roll-out model (model.cpp:2); Linearize uses 0.3302f
}

Model::Model() : A_(3, 3), B_(3, 2), C_(3, 1) {}

void Model::Linearize(State& S, Input& I, double dt) {
  const float wheelbase = 0.3302f;  // model.cpp:32 — a float, widened in every expression below
  const double th = S.ori(), v = I.v(), de = I.steer_ang();
  const double inv_cos2 = std::pow(std::cos(de), -2);
  A_ = f110::Matrix(3, 3);
  B_ = f110::Matrix(3, 2);
  C_ = f110::Matrix(3, 1);
  for (int d = 0; d < 3; ++d) A_(d, d) = 1;
  A_(0, 2) = -1 * v * std::sin(th) * dt;
  A_(1, 2) = v * std::cos(th) * dt;
  B_(0, 0) = std::cos(th) * dt;
  B_(1, 0) = std::sin(th) * dt;
  B_(2, 0) = std::tan(de) * dt / wheelbase;
  B_(2, 1) = v * inv_cos2 * dt / wheelbase;
  C_(0, 0) = v * th * std::sin(th) * dt;
  C_(1, 0) = -1 * v * th * std::cos(th) * dt;
  C_(2, 0) = -1 * de * v * inv_cos2 * dt / wheelbase;
}

void Model::simulate_dynamics(State& state, Input& input, double dt, State& new_state) {
  const double rate[3] = {input.v() * std::cos(state.ori()), input.v() * std::sin(state.ori()),
                          std::tan(input.steer_ang()) * input.v() / kRolloutWheelbase};
  new_state.set_x(state.x() + rate[0] * dt);
  new_state.set_y(state.y() + rate[1] * dt);
  new_state.set_ori(state.ori() + rate[2] * dt);
}
