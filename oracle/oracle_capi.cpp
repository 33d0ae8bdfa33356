// ORACLE — TEST INFRASTRUCTURE ONLY (see osqp_restated.hpp header). PARITY UNPINNED.
// C entry points so tests/, smoke() and bench.py's cpu_baseline / --impl reference legs can drive
// the CPU restatement through ctypes.  Flat double arrays carry config/settings so this library
// shares no struct layout with the product.
//
//   cfg[18]      = N, dt, Q0,Q1,Q2, R0,R1, udes0,udes1, umin0,umin1, umax0,umax1, gap_mode, rate_rows, rate_delta, state_rows, state_lim
//   settings[15] = rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, max_iter,
//                  check_termination, scaling, adaptive_rho, adaptive_rho_interval,
//                  adaptive_rho_tolerance, warm_start, scaled_termination
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "f110_ref.hpp"

using namespace f110_ref;
namespace osq = osqp_restated;

static MpcConfig cfg_from(const double* c) {
  MpcConfig k;
  k.N = (int)c[0]; k.dt = c[1];
  for (int i = 0; i < 3; ++i) k.Q[i] = c[2 + i];
  for (int i = 0; i < 2; ++i) { k.R[i] = c[5 + i]; k.u_des[i] = c[7 + i]; k.u_min[i] = c[9 + i]; k.u_max[i] = c[11 + i]; }
  k.gap_mode = (int)c[13];
  k.rate_rows = (int)c[14]; k.rate_delta = c[15];
  k.state_rows = (int)c[16]; k.state_lim = c[17];
  return k;
}
static osq::Settings settings_from(const double* s) {
  osq::Settings t;
  t.rho = s[0]; t.sigma = s[1]; t.alpha = s[2]; t.eps_abs = s[3]; t.eps_rel = s[4];
  t.eps_prim_inf = s[5]; t.eps_dual_inf = s[6]; t.max_iter = (int)s[7]; t.check_termination = (int)s[8];
  t.scaling = (int)s[9]; t.adaptive_rho = (int)s[10]; t.adaptive_rho_interval = (int)s[11];
  t.adaptive_rho_tolerance = s[12]; t.warm_start = (int)s[13]; t.scaled_termination = (int)s[14];
  return t;
}

struct MpcBatch {
  MpcConfig cfg;
  osq::Settings settings;
  int B = 0, nthreads = 1;
  // cold mode: one (structure, solver) per thread; warm mode: one per QP, persistent
  std::vector<std::unique_ptr<QpData>> thread_qp;
  std::vector<std::unique_ptr<osq::Solver>> thread_solver;
  std::vector<char> thread_ready;
  std::vector<std::unique_ptr<QpData>> qp;
  std::vector<std::unique_ptr<osq::Solver>> solver;
  std::vector<char> qp_ready;
};

extern "C" {

int orc_qp_record_doubles(int N) { return qp_record_doubles(N); }
void orc_default_cfg(double* c) {
  MpcConfig k;
  c[0] = k.N; c[1] = k.dt;
  for (int i = 0; i < 3; ++i) c[2 + i] = k.Q[i];
  for (int i = 0; i < 2; ++i) { c[5 + i] = k.R[i]; c[7 + i] = k.u_des[i]; c[9 + i] = k.u_min[i]; c[11 + i] = k.u_max[i]; }
  c[13] = k.gap_mode; c[14] = k.rate_rows; c[15] = k.rate_delta; c[16] = k.state_rows; c[17] = k.state_lim;
}
void orc_default_settings(double* s) {
  osq::Settings t;
  s[0] = t.rho; s[1] = t.sigma; s[2] = t.alpha; s[3] = t.eps_abs; s[4] = t.eps_rel; s[5] = t.eps_prim_inf;
  s[6] = t.eps_dual_inf; s[7] = t.max_iter; s[8] = t.check_termination; s[9] = t.scaling; s[10] = t.adaptive_rho;
  s[11] = t.adaptive_rho_interval; s[12] = t.adaptive_rho_tolerance; s[13] = t.warm_start; s[14] = t.scaled_termination;
}

// ---- generic dense-in QP (closed-form / scipy cross-check tests) -----------------------------------
// P: n x n row-major symmetric; A: m x n row-major.  info_out[8] = iter, status, obj, pri, dua,
// rho_estimate, rho_updates, n_factor.
int orc_osqp_dense(int n, int m, const double* P, const double* q, const double* A, const double* l,
                   const double* u, const double* settings, double* x, double* y, double* info_out) {
  osq::Csc Pc, Ac;
  Pc.nrow = Pc.ncol = n; Pc.p.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i <= j; ++i)
      if (P[(size_t)i * n + j] != 0.0) { Pc.i.push_back(i); Pc.x.push_back(P[(size_t)i * n + j]); }
    Pc.p[j + 1] = (int)Pc.x.size();
  }
  Ac.nrow = m; Ac.ncol = n; Ac.p.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i)
      if (A[(size_t)i * n + j] != 0.0) { Ac.i.push_back(i); Ac.x.push_back(A[(size_t)i * n + j]); }
    Ac.p[j + 1] = (int)Ac.x.size();
  }
  osq::Solver s;
  if (s.setup(Pc, q, Ac, l, u, settings_from(settings), nullptr)) return 1;
  if (s.solve()) return 2;
  std::copy(s.sol_x.begin(), s.sol_x.end(), x);
  std::copy(s.sol_y.begin(), s.sol_y.end(), y);
  info_out[0] = s.info.iter; info_out[1] = s.info.status; info_out[2] = s.info.obj_val; info_out[3] = s.info.pri_res;
  info_out[4] = s.info.dua_res; info_out[5] = s.info.rho_estimate; info_out[6] = s.info.rho_updates; info_out[7] = s.info.n_factor;
  return 0;
}

// ---- MPC QP: dense assembly for cross-checks ---------------------------------------------------------
// Pd n x n, Ad m x n row-major, plus q, l, u.  Returns 0; dims via orc_mpc_dims.
void orc_mpc_dims(int N, int* n, int* m) { *n = 5 * N + 3; *m = 7 * N + 5; }
int orc_mpc_rows(const double* cfg) { MpcConfig k = cfg_from(cfg); return 7 * k.N + 5 + (k.rate_rows ? k.N : 0) + (k.state_rows ? 3 * (k.N + 1) : 0); }
int orc_mpc_assemble_dense(const double* cfg, const double* rec, double* Pd, double* q, double* Ad, double* l, double* u) {
  MpcConfig k = cfg_from(cfg);
  QpData d;
  qp_build_structure(k, &d);
  qp_fill_values(k, rec, &d);
  std::fill(Pd, Pd + (size_t)d.n * d.n, 0.0);
  std::fill(Ad, Ad + (size_t)d.m * d.n, 0.0);
  for (int j = 0; j < d.n; ++j)
    for (int p = d.P.p[j]; p < d.P.p[j + 1]; ++p) { Pd[(size_t)d.P.i[p] * d.n + j] = d.P.x[p]; Pd[(size_t)j * d.n + d.P.i[p]] = d.P.x[p]; }
  for (int j = 0; j < d.n; ++j)
    for (int p = d.A.p[j]; p < d.A.p[j + 1]; ++p) Ad[(size_t)d.A.i[p] * d.n + j] = d.A.x[p];
  std::copy(d.q.begin(), d.q.end(), q);
  std::copy(d.l.begin(), d.l.end(), l);
  std::copy(d.u.begin(), d.u.end(), u);
  return 0;
}
int orc_mpc_nnz(const double* cfg, int* nnzP, int* nnzA, int* nnzL) {
  MpcConfig k = cfg_from(cfg);
  QpData d;
  qp_build_structure(k, &d);
  std::vector<double> rec(qp_record_doubles(k.N), 0.0);
  rec[3] = 4.5;
  qp_fill_values(k, rec.data(), &d);
  osq::Solver s;
  osq::Settings st;
  s.setup(d.P, d.q.data(), d.A, d.l.data(), d.u.data(), st, d.perm.data());
  *nnzP = d.P.nnz(); *nnzA = d.A.nnz(); *nnzL = s.kkt_nnzL();
  return 0;
}

// ---- MPC batch handle -----------------------------------------------------------------------------------
void* orc_mpc_create(const double* cfg, const double* settings, int B, int nthreads) {
  MpcBatch* h = new MpcBatch();
  h->cfg = cfg_from(cfg);
  h->settings = settings_from(settings);
  h->B = B;
  int hw = (int)std::thread::hardware_concurrency();
  h->nthreads = nthreads > 0 ? nthreads : (hw > 0 ? hw : 1);
  h->thread_qp.resize(h->nthreads); h->thread_solver.resize(h->nthreads); h->thread_ready.assign(h->nthreads, 0);
  h->qp.resize(B); h->solver.resize(B); h->qp_ready.assign(B, 0);
  return h;
}
void orc_mpc_destroy(void* hv) { delete (MpcBatch*)hv; }
int orc_mpc_threads(void* hv) { return ((MpcBatch*)hv)->nthreads; }

// Solve B QPs.  warm = 0: every QP is a fresh problem (scaling, rho = settings.rho, cold start).
// warm = 1: reference steady-state path per QP slot — first call = setup (mpc.cpp:98-129), later calls =
// updateGradient / updateLinearConstraintsMatrix / updateBounds (mpc.cpp:83-94) then solve() with the
// iterates and rho kept from the previous call.
// Outputs (any may be null): x[B*n], y[B*m], status[B], iters[B], rho_updates[B],
// extra[B*4] = obj, pri_res, dua_res, rho at exit.  Returns wall seconds of the solve loop (<0 on error).
double orc_mpc_solve(void* hv, const double* recs, int stride, int count, int warm, double* x, double* y,
                     int* status, int* iters, int* rho_updates, double* extra) {
  MpcBatch* h = (MpcBatch*)hv;
  if (count > h->B) return -1.0;
  const int n = 5 * h->cfg.N + 3, m = 7 * h->cfg.N + 5 + (h->cfg.rate_rows ? h->cfg.N : 0) + (h->cfg.state_rows ? 3 * (h->cfg.N + 1) : 0);
  std::atomic<int> next(0);
  std::atomic<int> err(0);
  auto worker = [&](int tid) {
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= count) break;
      const double* rec = recs + (size_t)b * stride;
      osq::Solver* s;
      if (!warm) {
        if (!h->thread_ready[tid]) {
          h->thread_qp[tid].reset(new QpData());
          qp_build_structure(h->cfg, h->thread_qp[tid].get());
          h->thread_solver[tid].reset(new osq::Solver());
        }
        QpData* d = h->thread_qp[tid].get();
        s = h->thread_solver[tid].get();
        qp_fill_values(h->cfg, rec, d);
        if (!h->thread_ready[tid]) {
          osq::Settings st = h->settings; st.warm_start = 0;
          if (s->setup(d->P, d->q.data(), d->A, d->l.data(), d->u.data(), st, d->perm.data())) { err = 1; break; }
          h->thread_ready[tid] = 1;
        } else {
          if (s->resetup(d->q.data(), d->A.x.data(), d->l.data(), d->u.data())) { err = 1; break; }
        }
      } else {
        if (!h->qp_ready[b]) {
          h->qp[b].reset(new QpData());
          qp_build_structure(h->cfg, h->qp[b].get());
          h->solver[b].reset(new osq::Solver());
        }
        QpData* d = h->qp[b].get();
        s = h->solver[b].get();
        qp_fill_values(h->cfg, rec, d);
        if (!h->qp_ready[b]) {
          osq::Settings st = h->settings; st.warm_start = 1;
          if (s->setup(d->P, d->q.data(), d->A, d->l.data(), d->u.data(), st, d->perm.data())) { err = 1; break; }
          h->qp_ready[b] = 1;
        } else {
          s->update_lin_cost(d->q.data());
          if (s->update_A(d->A.x.data())) { err = 1; break; }
          if (s->update_bounds(d->l.data(), d->u.data())) { err = 1; break; }
        }
      }
      if (s->solve()) { err = 2; break; }
      if (x) std::copy(s->sol_x.begin(), s->sol_x.end(), x + (size_t)b * n);
      if (y) std::copy(s->sol_y.begin(), s->sol_y.end(), y + (size_t)b * m);
      if (status) status[b] = s->info.status;
      if (iters) iters[b] = s->info.iter;
      if (rho_updates) rho_updates[b] = s->info.rho_updates;
      if (extra) {
        extra[4 * (size_t)b + 0] = s->info.obj_val; extra[4 * (size_t)b + 1] = s->info.pri_res;
        extra[4 * (size_t)b + 2] = s->info.dua_res; extra[4 * (size_t)b + 3] = s->rho();
      }
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  int nt = std::min(h->nthreads, std::max(count, 1));
  if (nt <= 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (err) return -(double)err;
  return std::chrono::duration<double>(t1 - t0).count();
}

// scaling vectors + iterates of one warm slot (for state-parity tests)
int orc_mpc_get_scaling(void* hv, int slot, double* D, double* E, double* c) {
  MpcBatch* h = (MpcBatch*)hv;
  if (slot >= h->B || !h->qp_ready[slot]) return 1;
  osq::Solver* s = h->solver[slot].get();
  std::copy(s->D.begin(), s->D.end(), D);
  std::copy(s->E.begin(), s->E.end(), E);
  *c = s->c;
  return 0;
}

// ---- pipeline pieces ---------------------------------------------------------------------------------------
void orc_linearize(double ori, double v, double steer, double dt, double* A, double* B, double* C) {
  linearize(ori, v, steer, dt, A, B, C);
}
void orc_simulate_dynamics(const double* s, const double* in, double dt, double* out) {
  State3 a{s[0], s[1], s[2]};
  Input2 i{in[0], in[1]};
  State3 o = simulate_dynamics(a, i, dt);
  out[0] = o.x; out[1] = o.y; out[2] = o.ori;
}
// out: P x S x 3 (x, y, ori); returns P = steer_discrete + 1
int orc_traj_table(double steer_max, int steer_discrete, int traj_discrete, double speed_max, double dt, double* out) {
  Params p;
  p.steer_max = steer_max; p.steer_discrete = steer_discrete; p.traj_discrete = traj_discrete;
  p.speed_max = speed_max; p.dt_plan = dt;
  auto t = generate_traj_table(p);
  for (size_t i = 0; i < t.size(); ++i)
    for (size_t j = 0; j < t[i].size(); ++j) {
      out[(i * traj_discrete + j) * 3 + 0] = t[i][j].x;
      out[(i * traj_discrete + j) * 3 + 1] = t[i][j].y;
      out[(i * traj_discrete + j) * 3 + 2] = t[i][j].ori;
    }
  return (int)t.size();
}
// pose7 = px, py, pz, qx, qy, qz, qw
static Pose pose_from(const double* p) {
  Pose q; q.px = p[0]; q.py = p[1]; q.pz = p[2]; q.qx = p[3]; q.qy = p[4]; q.qz = p[5]; q.qw = p[6];
  return q;
}
void orc_car_to_world_R(const double* pose7, double* R4) {
  Mat3 r = car_to_world_rotation(pose_from(pose7));
  R4[0] = r.m[0][0]; R4[1] = r.m[0][1]; R4[2] = r.m[1][0]; R4[3] = r.m[1][1];
}
float orc_car_orientation(const double* pose7) { return car_orientation(pose_from(pose7)); }
// grid: blocks*blocks floats, Eigen column-major (row + col*blocks); offset[2]; returns blocks
int orc_fill_grid(int occ_size, float discrete, float dilation, const double* pose7, float angle_min, float angle_max,
                  float angle_inc, const float* ranges, int n_ranges, float* grid, float* offset) {
  Params p; p.occ_size = occ_size; p.occ_discrete = discrete; p.occ_dilation = dilation;
  OccGrid g(p);
  Scan s; s.angle_min = angle_min; s.angle_max = angle_max; s.angle_increment = angle_inc; s.ranges = ranges; s.n_ranges = n_ranges;
  g.FillOccGrid(pose_from(pose7), s);
  std::copy(g.grid_.begin(), g.grid_.end(), grid);
  offset[0] = g.occ_offset_.first; offset[1] = g.occ_offset_.second;
  return g.grid_blocks_;
}
int orc_grid_blocks(int occ_size, float discrete) { return (int)(occ_size / discrete); }
// table_xy: P x S x 2 doubles
void orc_collision_check(const float* grid, int blocks, float discrete, const float* offset, const double* R4,
                         const double* pose_xy, const double* table_xy, int P, int S, uint8_t* valid,
                         int* free_count, float* end_world) {
  Params p; p.occ_discrete = discrete;
  OccGrid g(p);
  g.discrete_ = discrete; g.grid_blocks_ = blocks;
  g.grid_.assign(grid, grid + (size_t)blocks * blocks);
  g.occ_offset_.first = offset[0]; g.occ_offset_.second = offset[1];
  CheckResult r = collision_check(g, R4, pose_xy[0], pose_xy[1], table_xy, P, S);
  for (int i = 0; i < P; ++i) { valid[i] = r.valid[i]; free_count[i] = r.free_count[i]; end_world[2 * i] = r.end_world[2 * i]; end_world[2 * i + 1] = r.end_world[2 * i + 1]; }
}
int orc_select_best(const uint8_t* valid, const float* end_world, int P, double gx, double gy) {
  CheckResult r;
  r.valid.assign(valid, valid + P); r.end_world.assign(end_world, end_world + 2 * P);
  return select_best_path(r, gx, gy);
}
// wp_xy: W x 2 floats (already float-parsed CSV columns 0,1); out_ori: W doubles (float-valued)
void orc_waypoint_headings(const float* wp_xy, int W, double* out_ori) {
  std::vector<std::pair<float, float>> t(W);
  for (int i = 0; i < W; ++i) t[i] = {wp_xy[2 * i], wp_xy[2 * i + 1]};
  auto wp = waypoints_from_xy(t);
  for (int i = 0; i < W; ++i) out_ori[i] = wp[i].ori;
}
int orc_best_global_idx(const float* wp_xy, int W, const double* pose7, float lookahead) {
  std::vector<State3> wp(W);
  for (int i = 0; i < W; ++i) { wp[i].x = wp_xy[2 * i]; wp[i].y = wp_xy[2 * i + 1]; }
  return get_best_global_idx(wp, pose_from(pose7), lookahead);
}
// returns 1 when half-planes were produced, 0 for the reference's undefined no-gap case
int orc_find_half_spaces(float thresh, float divider, float buffer, const double* state3, float angle_min, float angle_max,
                         float angle_inc, const float* ranges, int n_ranges, double* l1, double* l2, int* lohi) {
  Params p; p.follow_gap_thresh = thresh; p.fov_divider = divider; p.buffer = buffer;
  Scan s; s.angle_min = angle_min; s.angle_max = angle_max; s.angle_increment = angle_inc; s.ranges = ranges; s.n_ranges = n_ranges;
  State3 st{state3[0], state3[1], state3[2]};
  HalfSpaces h;
  bool ok = find_half_spaces(p, st, s, &h);
  lohi[0] = h.best_lo; lohi[1] = h.best_hi;
  if (!ok) return 0;
  for (int i = 0; i < 3; ++i) { l1[i] = h.l1[i]; l2[i] = h.l2[i]; }
  return 1;
}

}  // extern "C"
