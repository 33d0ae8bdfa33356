#include "mpc.h"
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace {
f110_mpc_config config_from(const f110::Params& prm, const Constraints& con) {
  f110_mpc_config c;
  f110_mpc_default_config(&c);
  c.horizon = prm.horizon;
  c.gap_mode = prm.gap_mode;
  c.dt = prm.dt;  // the float dt_ widens to double where the reference passes it to Linearize (mpc.cpp:73)
  c.q[0] = prm.q0; c.q[1] = prm.q1; c.q[2] = prm.q2;
  c.r[0] = prm.r0; c.r[1] = prm.r1;
  c.u_des[0] = prm.des_vel; c.u_des[1] = prm.des_steer;
  for (int j = 0; j < 2; ++j) { c.u_min[j] = con.u_min()(j); c.u_max[j] = con.u_max()(j); }
  if (prm.steer_rate_max > 0.0) { c.rate_rows = 1; c.rate_delta = prm.steer_rate_max * c.dt; }
  if (prm.state_box) { c.state_rows = 1; c.state_lim = prm.state_lims; }   // Constraints::SetXLims' d (constraints.cpp:10, 108-114)
  return c;
}

// record = x0 | (v, steer) | l1 | l2 | ref[0..N-1]; a trajectory shorter than the horizon repeats its last state
void fill_record(double* rec, int N, const State& x0, const Input& in, const std::vector<State>& desired,
                 const f110::Vector& l1, const f110::Vector& l2) {
  rec[0] = x0.x(); rec[1] = x0.y(); rec[2] = x0.ori();
  rec[3] = in.v(); rec[4] = in.steer_ang();
  for (int j = 0; j < 3; ++j) { rec[5 + j] = l1(j); rec[8 + j] = l2(j); }
  const int have = static_cast<int>(desired.size());
  for (int k = 0; k < N; ++k) {
    const State& s = desired[k < have ? k : have - 1];
    rec[11 + 3 * k] = s.x(); rec[12 + 3 * k] = s.y(); rec[13 + 3 * k] = s.ori();
  }
}
}  // namespace

MPC::MPC(const f110::Params& prm, int device)
    : horizon_(prm.horizon), dt_(prm.dt), constraints_(prm),
      cost_(f110::Matrix::Diagonal({prm.q0, prm.q1, prm.q2}), f110::Matrix::Diagonal({prm.r0, prm.r1})),
      desired_input_(prm.des_vel, prm.des_steer) {
  num_inputs_ = input_size_ * horizon_;                                       // mpc.cpp:26
  num_states_ = state_size_ * (horizon_ + 1);                                 // mpc.cpp:27
  num_variables_ = num_states_ + num_inputs_;                                 // mpc.cpp:28
  num_constraints_ = num_states_ + 2 * (horizon_ + 1) + num_inputs_;          // mpc.cpp:29: dynamics + gap + input box
  config_ = config_from(prm, constraints_);
  num_constraints_ = f110_mpc_num_rows(&config_);                             // + N steering-rate rows when enabled
  QPsolution_.assign(num_variables_, 0.0);
  QPdual_.assign(num_constraints_, 0.0);
  record_.assign(f110_mpc_record_doubles(horizon_), 0.0);
  f110_solver_default_settings(&settings_);  // OSQP defaults + warm start, the reference's configuration (mpc.cpp:98-99)
  const int rc = f110_mpc_create(&config_, &settings_, 1, device, &solver_);
  if (rc != F110_OK) throw std::runtime_error(std::string("MPC: f110_mpc_create failed: ") + f110_last_error());
}

MPC::~MPC() { f110_mpc_destroy(solver_); }

void MPC::Update(State current_state, Input input, std::vector<State>& desired_state_trajectory) {
  current_state_ = current_state;
  desired_state_trajectory_ = desired_state_trajectory;
  model_.Linearize(current_state_, input, dt_);                       // kept for the A()/B()/C() accessors
  constraints_.set_state(current_state_);
  if (!scan_msg_.ranges.empty()) constraints_.FindHalfSpaces(current_state_, scan_msg_);
  if (desired_state_trajectory_.empty()) {
    // the reference indexes an empty vector here (project.cpp:184 clears it first) — defined: skip the cycle
    std::fprintf(stderr, "MPC::Update: empty desired trajectory, cycle skipped\n");
    return;
  }
  fill_record(record_.data(), horizon_, current_state_, input, desired_state_trajectory_, constraints_.l1(), constraints_.l2());
  std::vector<double> x(num_variables_), y(num_constraints_);
  double u0[2];
  int32_t status = F110_UNSOLVED, iters = 0;
  const int rc = f110_mpc_solve_host(solver_, 1, record_.data(), static_cast<int>(record_.size()), x.data(), y.data(), u0, &status, &iters);
  last_status_ = status;
  last_iters_ = iters;
  if (rc != F110_OK || status != F110_SOLVED) {
    // OsqpEigen::Solver::solve() returns false for an error or any status but "solved" (mpc.cpp:133-136)
    std::fprintf(stderr, "solve failed (%s)\n", rc != F110_OK ? f110_last_error() : "status");
    return;
  }
  std::lock_guard<std::mutex> lock(result_mutex_);
  QPsolution_ = x;
  QPdual_ = y;
  UpdateSolvedTrajectory();
}

void MPC::UpdateSolvedTrajectory() {
  solved_trajectory_.clear();
  for (int i = num_states_; i < num_variables_ - 1; i += 2) {  // mpc.cpp:148
    const double v = QPsolution_[i], angle = QPsolution_[i + 1];
    if (std::isnan(v) || std::isnan(angle)) return;  // leaves a truncated trajectory, like the reference
    solved_trajectory_.emplace_back(v, angle);
  }
}

std::vector<Input> MPC::solved_trajectory() {
  std::lock_guard<std::mutex> lock(result_mutex_);
  return solved_trajectory_;
}

BatchMPC::BatchMPC(const f110::Params& prm, int max_batch, int device, bool warm_start) : max_batch_(max_batch) {
  Constraints con(prm);
  config_ = config_from(prm, con);
  f110_solver_default_settings(&settings_);
  settings_.warm_start = warm_start ? 1 : 0;
  const int rc = f110_mpc_create(&config_, &settings_, max_batch, device, &solver_);
  if (rc != F110_OK) throw std::runtime_error(std::string("BatchMPC: f110_mpc_create failed: ") + f110_last_error());
  records_.assign(static_cast<std::size_t>(max_batch) * record_doubles(), 0.0);
  u0_.assign(2 * static_cast<std::size_t>(max_batch), 0.0);
  status_.assign(max_batch, F110_UNSOLVED);
  iters_.assign(max_batch, 0);
}

BatchMPC::BatchMPC(const f110::Params& prm, int max_batch, const std::vector<int>& devices, int shard_unit)
    : max_batch_(max_batch), shard_unit_(shard_unit) {
  Constraints con(prm);
  config_ = config_from(prm, con);
  f110_solver_default_settings(&settings_);
  settings_.warm_start = 0;
  const int rc = f110_mpc_create_multi(&config_, &settings_, max_batch, devices.data(), static_cast<int>(devices.size()), &multi_);
  if (rc != F110_OK) throw std::runtime_error(std::string("BatchMPC: f110_mpc_create_multi failed: ") + f110_last_error());
  records_.assign(static_cast<std::size_t>(max_batch) * record_doubles(), 0.0);
  u0_.assign(2 * static_cast<std::size_t>(max_batch), 0.0);
  status_.assign(max_batch, F110_UNSOLVED);
  iters_.assign(max_batch, 0);
}

BatchMPC::~BatchMPC() {
  f110_mpc_destroy(solver_);
  f110_mpc_destroy_multi(multi_);
}

void BatchMPC::SetProblem(int b, const State& x0, const Input& in, const std::vector<State>& desired, const f110::Vector& l1,
                          const f110::Vector& l2) {
  fill_record(records_.data() + static_cast<std::size_t>(b) * record_doubles(), config_.horizon, x0, in, desired, l1, l2);
}

int BatchMPC::Solve(int count) {
  if (multi_)
    return f110_mpc_solve_multi_host(multi_, count, shard_unit_, records_.data(), record_doubles(), u0_.data(), status_.data(), iters_.data());
  return f110_mpc_solve_host(solver_, count, records_.data(), record_doubles(), nullptr, nullptr, u0_.data(), status_.data(), iters_.data());
}
