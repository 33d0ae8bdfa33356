// MPC — same public interface as the reference class (include/f110-mpc/mpc.h:18-37), ROS-free.  Where the
// reference drives an OsqpEigen::Solver member (mpc.h:63, mpc.cpp:81-142) this class marshals the cycle's data
// into one parameter record and calls the CUDA solver through the C ABI (include/f110_mpc_b200.h) with a batch
// of one; the QP matrices are never built on the host.  BatchMPC is the same thing for many independent cars.
#pragma once
#include <mutex>
#include <vector>
#include "../../include/f110_mpc_b200.h"
#include "constraints.h"
#include "cost.h"
#include "model.h"
#include "msgs.h"
#include "state.h"

class MPC {
 public:
  explicit MPC(const f110::Params& params, int device = 0);
  virtual ~MPC();
  MPC(const MPC&) = delete;
  MPC& operator=(const MPC&) = delete;

  // One MPC cycle from the latest state and the last applied input, tracking desired_state_trajectory
  // (mpc.cpp:69-143).  On a failed solve the previous solved trajectory is kept, as in the reference.
  void Update(State current_state, Input input, std::vector<State>& desired_state_trajectory);
  void UpdateScan(const sensor_msgs::LaserScan& scan_msg) { scan_msg_ = scan_msg; }  // mpc.cpp:64-67

  Constraints constraints() const { return constraints_; }
  float dt() const { return dt_; }
  int horizon() const { return horizon_; }
  std::vector<Input> solved_trajectory();  // copy taken under a lock (the reference reads it unsynchronised)

  // extras the reference has no equivalent for (OsqpEigen exposes them, mpc.cpp never reads them)
  int last_status() const { return last_status_; }
  int last_iterations() const { return last_iters_; }
  const std::vector<double>& last_primal() const { return QPsolution_; }
  const std::vector<double>& last_dual() const { return QPdual_; }
  int num_variables() const { return num_variables_; }
  int num_constraints() const { return num_constraints_; }
  Model& model() { return model_; }
  Cost& cost() { return cost_; }
  f110_solver_settings& settings() { return settings_; }  // effective at the next (re)creation only

 private:
  int horizon_, input_size_ = 2, state_size_ = 3;
  int num_states_, num_inputs_, num_variables_, num_constraints_;
  float dt_;
  Constraints constraints_;
  Model model_;
  Cost cost_;
  State current_state_;
  Input desired_input_;
  std::vector<State> desired_state_trajectory_;
  sensor_msgs::LaserScan scan_msg_;
  std::vector<double> QPsolution_, QPdual_, record_;
  std::vector<Input> solved_trajectory_;
  std::mutex result_mutex_;
  int last_status_ = F110_UNSOLVED, last_iters_ = 0;
  f110_mpc_config config_;
  f110_solver_settings settings_;
  f110_mpc_solver* solver_ = nullptr;
  void UpdateSolvedTrajectory();  // mpc.cpp:145-159
};

// Many independent cars / candidate paths at once: slot b of every call is one persistent "solver instance"
// (its warm start lives on the device).
class BatchMPC {
 public:
  BatchMPC(const f110::Params& params, int max_batch, int device = 0, bool warm_start = false);
  // Several GPUs of one box behind the same interface: Solve() cuts the batch into contiguous shards of whole `shard_unit`s (QPs
  // that belong together, e.g. one scenario's lane x path QPs), one shard per device; the chosen controls of every device land on
  // devices[0] through the solve kernels' own peer stores and come back in one copy.  Cold start (slots move with the sharding).
  BatchMPC(const f110::Params& params, int max_batch, const std::vector<int>& devices, int shard_unit = 1);
  ~BatchMPC();
  int num_devices() const { return multi_ ? f110_mpc_multi_devices(multi_) : 1; }
  BatchMPC(const BatchMPC&) = delete;
  BatchMPC& operator=(const BatchMPC&) = delete;
  int horizon() const { return config_.horizon; }
  int record_doubles() const { return f110_mpc_record_doubles(config_.horizon); }
  // Fill the record of slot b exactly as MPC::Update would for that car.
  void SetProblem(int b, const State& current_state, const Input& input, const std::vector<State>& desired,
                  const f110::Vector& l1, const f110::Vector& l2);
  // Solve slots [0, count); returns 0 or an F110_ERR_* code.
  int Solve(int count);
  Input first_input(int b) const { return Input(u0_[2 * b], u0_[2 * b + 1]); }
  int status(int b) const { return status_[b]; }
  int iterations(int b) const { return iters_[b]; }
  std::vector<double>& records() { return records_; }
 private:
  f110_mpc_config config_;
  f110_solver_settings settings_;
  f110_mpc_solver* solver_ = nullptr;
  f110_mpc_multi* multi_ = nullptr;
  int max_batch_, shard_unit_ = 1;
  std::vector<double> records_, u0_;
  std::vector<int32_t> status_, iters_;
};
