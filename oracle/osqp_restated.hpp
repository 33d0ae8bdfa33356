// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.
//
// PARITY UNPINNED: the arithmetic restated here lives in OSQP (un-vendored, un-pinned
// third-party dependency of the reference: /root/reference/CMakeLists.txt:20-22,56-60,
// `find_package(OsqpEigen REQUIRED)`); the reference holds no tests or golden vectors
// (SURVEY.md §4, §8c).  This file restates the *published* OSQP algorithm
// (Stellato et al., "OSQP: an operator splitting solver for quadratic programs",
// Math. Prog. Comp. 2020) with the v0.6.x default constants and control flow, as it is
// driven from the reference call sites /root/reference/src/mpc.cpp:81-142.
// It is pinned by (a) closed-form QPs, (b) KKT residual checks, (c) an independent
// scipy cross-check (tests/test_oracle_*.py) — not by the OSQP binary, which is absent.
//
// Generic sparse formulation (CSC P upper-triangular, CSC A), quasi-definite KKT
//   [P + sigma I, A^T; A, -diag(1/rho)]
// factored by an up-looking sparse LDL^T under a caller-supplied permutation.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace osqp_restated {

// ---- constants (OSQP v0.6.x `constants.h`) ---------------------------------------
constexpr double OSQP_INFTY = 1e30;
constexpr double RHO_MIN = 1e-6;
constexpr double RHO_MAX = 1e6;
constexpr double RHO_EQ_OVER_RHO_INEQ = 1e3;
constexpr double RHO_TOL = 1e-4;
constexpr double MIN_SCALING = 1e-4;
constexpr double MAX_SCALING = 1e4;

enum Status : int {
  SOLVED = 1,
  SOLVED_INACCURATE = 2,
  PRIMAL_INFEASIBLE_INACCURATE = 3,
  DUAL_INFEASIBLE_INACCURATE = 4,
  MAX_ITER_REACHED = -2,
  PRIMAL_INFEASIBLE = -3,
  DUAL_INFEASIBLE = -4,
  NON_CVX = -7,
  UNSOLVED = -10,
};

struct Settings {
  double rho = 0.1;
  double sigma = 1e-6;
  double alpha = 1.6;
  double eps_abs = 1e-3;
  double eps_rel = 1e-3;
  double eps_prim_inf = 1e-4;
  double eps_dual_inf = 1e-4;
  int max_iter = 4000;
  int check_termination = 25;
  int scaling = 10;
  int adaptive_rho = 1;
  // OSQP's default 0 means "pick from setup/solve time ratio" — timing dependent.
  // The oracle fixes it: 25 = the value OSQP lands on for small problems.
  int adaptive_rho_interval = 25;
  double adaptive_rho_tolerance = 5.0;
  int warm_start = 1;
  int scaled_termination = 0;
};

struct Csc {
  int nrow = 0, ncol = 0;
  std::vector<int> p, i;
  std::vector<double> x;
  int nnz() const { return (int)x.size(); }
};

struct Info {
  int iter = 0;
  int status = UNSOLVED;
  double obj_val = 0, pri_res = 0, dua_res = 0, rho_estimate = 0;
  int rho_updates = 0;
  int n_factor = 0;  // numeric factorisations done in this solve()+preceding updates
};

// ---- sparse up-looking LDL^T (Davis' LDL algorithm; what QDLDL implements) ------
class SparseLDL {
 public:
  // K: upper-triangular CSC of the (already permuted) symmetric matrix.
  void symbolic(const Csc& K) {
    n_ = K.ncol;
    parent_.assign(n_, -1);
    lnz_.assign(n_, 0);
    flag_.assign(n_, -1);
    for (int k = 0; k < n_; ++k) {
      flag_[k] = k;
      for (int p = K.p[k]; p < K.p[k + 1]; ++p) {
        int i = K.i[p];
        if (i >= k) continue;
        for (; flag_[i] != k; i = parent_[i]) {
          if (parent_[i] == -1) parent_[i] = k;
          lnz_[i]++;
          flag_[i] = k;
        }
      }
    }
    lp_.assign(n_ + 1, 0);
    for (int k = 0; k < n_; ++k) lp_[k + 1] = lp_[k] + lnz_[k];
    li_.assign(lp_[n_], 0);
    lx_.assign(lp_[n_], 0.0);
    d_.assign(n_, 0.0);
    dinv_.assign(n_, 0.0);
    y_.assign(n_, 0.0);
    pattern_.assign(n_, 0);
    cnt_.assign(n_, 0);
  }
  int nnzL() const { return lp_.empty() ? 0 : lp_[n_]; }
  // returns false on a zero pivot
  bool numeric(const Csc& K) {
    std::fill(cnt_.begin(), cnt_.end(), 0);
    std::fill(flag_.begin(), flag_.end(), -1);
    for (int k = 0; k < n_; ++k) {
      y_[k] = 0.0;
      int top = n_;
      flag_[k] = k;
      for (int p = K.p[k]; p < K.p[k + 1]; ++p) {
        int i = K.i[p];
        if (i > k) continue;
        y_[i] += K.x[p];
        int len = 0;
        for (; flag_[i] != k; i = parent_[i]) {
          pattern_[len++] = i;
          flag_[i] = k;
        }
        while (len > 0) pattern_[--top] = pattern_[--len];
      }
      double dk = y_[k];
      y_[k] = 0.0;
      for (; top < n_; ++top) {
        int i = pattern_[top];
        double yi = y_[i];
        y_[i] = 0.0;
        int p2 = lp_[i] + cnt_[i];
        for (int p = lp_[i]; p < p2; ++p) y_[li_[p]] -= lx_[p] * yi;
        double lki = yi * dinv_[i];
        dk -= lki * yi;
        li_[p2] = k;
        lx_[p2] = lki;
        cnt_[i]++;
      }
      if (dk == 0.0) return false;
      d_[k] = dk;
      dinv_[k] = 1.0 / dk;
    }
    return true;
  }
  // in-place solve of L D L^T x = b
  void solve(double* x) const {
    for (int j = 0; j < n_; ++j) {
      double xj = x[j];
      for (int p = lp_[j]; p < lp_[j + 1]; ++p) x[li_[p]] -= lx_[p] * xj;
    }
    for (int j = 0; j < n_; ++j) x[j] *= dinv_[j];
    for (int j = n_ - 1; j >= 0; --j) {
      double xj = x[j];
      for (int p = lp_[j]; p < lp_[j + 1]; ++p) xj -= lx_[p] * x[li_[p]];
      x[j] = xj;
    }
  }

 private:
  int n_ = 0;
  std::vector<int> parent_, lnz_, flag_, lp_, li_, pattern_, cnt_;
  std::vector<double> lx_, d_, dinv_, y_;
};

// ---- the solver ------------------------------------------------------------------
class Solver {
 public:
  int n = 0, m = 0;
  Settings settings;
  Info info;
  // unscaled solution (OSQP `solution->x`, `solution->y`)
  std::vector<double> sol_x, sol_y;
  // scaling (OSQP `work->scaling`)
  std::vector<double> D, E, Dinv, Einv;
  double c = 1.0, cinv = 1.0;

  // osqp_setup(): copy data, scale, classify rows, build + factor KKT.
  // perm (size n+m, perm[k] = original KKT index placed at position k) may be null.
  int setup(const Csc& P_triu, const double* q, const Csc& A, const double* l, const double* u,
            const Settings& s, const int* perm) {
    settings = s;
    n = P_triu.ncol;
    m = A.nrow;
    P_ = P_triu;
    A_ = A;
    P_orig_x_ = P_triu.x;
    rho0_ = s.rho;
    q_.assign(q, q + n);
    l_.assign(l, l + m);
    u_.assign(u, u + m);
    x_.assign(n, 0.0); x_prev_.assign(n, 0.0); delta_x_.assign(n, 0.0);
    Px_.assign(n, 0.0); Aty_.assign(n, 0.0);
    z_.assign(m, 0.0); z_prev_.assign(m, 0.0); y_.assign(m, 0.0); delta_y_.assign(m, 0.0);
    Ax_.assign(m, 0.0); Adelta_x_.assign(m, 0.0); Atdelta_y_.assign(n, 0.0); Pdelta_x_.assign(n, 0.0);
    xz_tilde_.assign(n + m, 0.0);
    rho_vec_.assign(m, 0.0); rho_inv_vec_.assign(m, 0.0); constr_type_.assign(m, 0);
    D.assign(n, 1.0); Dinv.assign(n, 1.0); E.assign(m, 1.0); Einv.assign(m, 1.0);
    D_temp_.assign(n, 0.0); D_temp_A_.assign(n, 0.0); E_temp_.assign(m, 0.0);
    sol_x.assign(n, 0.0); sol_y.assign(m, 0.0);
    c = cinv = 1.0;
    if (settings.scaling) scale_data();
    set_rho_vec();
    build_kkt_structure(perm);
    fill_kkt_values();
    ldl_.symbolic(K_);
    if (!ldl_.numeric(K_)) return 1;
    info = Info();
    info.n_factor = 1;
    return 0;
  }

  // Same result as a fresh setup() on a problem with the same sparsity pattern (new q, A values,
  // l, u), but re-using the symbolic factorisation.  Used for cold batch throughput so the CPU
  // baseline is not charged for ordering/etree work the reference pays only once (mpc.cpp:98-129).
  int resetup(const double* q, const double* Ax_new, const double* l, const double* u) {
    P_.x = P_orig_x_;
    std::copy(Ax_new, Ax_new + A_.nnz(), A_.x.begin());
    q_.assign(q, q + n);
    l_.assign(l, l + m);
    u_.assign(u, u + m);
    settings.rho = rho0_;
    c = cinv = 1.0;
    if (settings.scaling) scale_data();
    else {
      std::fill(D.begin(), D.end(), 1.0); std::fill(Dinv.begin(), Dinv.end(), 1.0);
      std::fill(E.begin(), E.end(), 1.0); std::fill(Einv.begin(), Einv.end(), 1.0);
    }
    set_rho_vec();
    fill_kkt_values();
    if (!ldl_.numeric(K_)) return 1;
    cold_start();
    info = Info();
    info.n_factor = 1;
    return 0;
  }

  // osqp_update_lin_cost()
  void update_lin_cost(const double* q_new) {
    q_.assign(q_new, q_new + n);
    if (settings.scaling) {
      for (int j = 0; j < n; ++j) q_[j] = q_[j] * D[j];
      for (int j = 0; j < n; ++j) q_[j] *= c;
    }
    reset_info();
  }

  // osqp_update_A() with all values replaced (what OsqpEigen::updateLinearConstraintsMatrix
  // ends up doing for a fixed sparsity pattern): unscale, overwrite, re-Ruiz, refactor.
  int update_A(const double* Ax_new) {
    if (settings.scaling) unscale_data();
    std::copy(Ax_new, Ax_new + A_.nnz(), A_.x.begin());
    if (settings.scaling) scale_data();
    fill_kkt_values();
    if (!ldl_.numeric(K_)) return 1;
    reset_info();
    info.n_factor++;
    return 0;
  }

  // osqp_update_bounds(): scale, then update_rho_vec() (refactor only if a row changed class)
  int update_bounds(const double* l_new, const double* u_new) {
    for (int i = 0; i < m; ++i)
      if (l_new[i] > u_new[i]) return 1;
    l_.assign(l_new, l_new + m);
    u_.assign(u_new, u_new + m);
    if (settings.scaling) {
      for (int i = 0; i < m; ++i) { l_[i] *= E[i]; u_[i] *= E[i]; }
    }
    reset_info();
    return update_rho_vec();
  }

  // cold_start(): x = z = y = 0
  void cold_start() {
    std::fill(x_.begin(), x_.end(), 0.0);
    std::fill(z_.begin(), z_.end(), 0.0);
    std::fill(y_.begin(), y_.end(), 0.0);
  }

  // raw (scaled) iterates: OSQP keeps these across solves when warm_start = 1
  const std::vector<double>& iter_x() const { return x_; }
  const std::vector<double>& iter_z() const { return z_; }
  const std::vector<double>& iter_y() const { return y_; }
  void set_iterates(const double* x, const double* z, const double* y) {
    std::copy(x, x + n, x_.begin()); std::copy(z, z + m, z_.begin()); std::copy(y, y + m, y_.begin());
  }
  double rho() const { return settings.rho; }
  int kkt_nnzL() const { return ldl_.nnzL(); }
  const std::vector<double>& rho_vec() const { return rho_vec_; }

  // osqp_solve()
  int solve() {
    int iter;
    bool can_check = false;
    if (!settings.warm_start) cold_start();
    info.status = UNSOLVED;
    info.rho_updates = 0;
    bool broke = false;
    for (iter = 1; iter <= settings.max_iter; ++iter) {
      x_.swap(x_prev_);
      z_.swap(z_prev_);
      update_xz_tilde();
      update_x();
      update_z();
      update_y();
      can_check = settings.check_termination && (iter % settings.check_termination == 0);
      if (can_check) {
        update_info(iter);
        if (check_termination(false)) { broke = true; break; }
      }
      if (settings.adaptive_rho && settings.adaptive_rho_interval &&
          (iter % settings.adaptive_rho_interval == 0)) {
        if (!can_check) update_info(iter);
        if (adapt_rho()) return 1;
      }
    }
    if (!broke) iter = settings.max_iter;  // loop ran to completion
    if (!can_check) {
      update_info(iter);
      check_termination(false);
    }
    if (info.status == UNSOLVED) {
      if (!check_termination(true)) info.status = MAX_ITER_REACHED;
    }
    info.iter = iter;
    info.rho_estimate = compute_rho_estimate();
    store_solution();
    return 0;
  }

 private:
  Csc P_, A_;
  std::vector<double> P_orig_x_;
  double rho0_ = 0.1;
  std::vector<double> q_, l_, u_;
  std::vector<double> x_, x_prev_, delta_x_, Px_, Aty_, Atdelta_y_, Pdelta_x_;
  std::vector<double> z_, z_prev_, y_, delta_y_, Ax_, Adelta_x_;
  std::vector<double> xz_tilde_, rho_vec_, rho_inv_vec_;
  std::vector<int> constr_type_;
  std::vector<double> D_temp_, D_temp_A_, E_temp_;
  // KKT
  Csc K_;
  std::vector<int> perm_, pinv_;
  std::vector<int> P_to_K_, A_to_K_, sig_to_K_, rho_to_K_;
  std::vector<char> P_is_diag_;
  SparseLDL ldl_;
  std::vector<double> rhs_perm_;

  void reset_info() { info.status = UNSOLVED; info.rho_updates = 0; }

  // ---- linear algebra helpers ------------------------------------------------------
  static double norm_inf(const std::vector<double>& v) {
    double r = 0.0;
    for (double a : v) { double b = std::fabs(a); if (b > r) r = b; }
    return r;
  }
  static double scaled_norm_inf(const std::vector<double>& S, const std::vector<double>& v) {
    double r = 0.0;
    for (size_t k = 0; k < v.size(); ++k) { double b = std::fabs(S[k] * v[k]); if (b > r) r = b; }
    return r;
  }
  void mat_vec_A(const std::vector<double>& x, std::vector<double>& y) const {  // y = A x
    std::fill(y.begin(), y.end(), 0.0);
    for (int j = 0; j < A_.ncol; ++j)
      for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) y[A_.i[p]] += A_.x[p] * x[j];
  }
  void mat_tvec_A(const std::vector<double>& x, std::vector<double>& y) const {  // y = A^T x
    for (int j = 0; j < A_.ncol; ++j) {
      double s = 0.0;
      for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) s += A_.x[p] * x[A_.i[p]];
      y[j] = s;
    }
  }
  void mat_vec_P(const std::vector<double>& x, std::vector<double>& y) const {  // y = P x, P sym triu
    std::fill(y.begin(), y.end(), 0.0);
    for (int j = 0; j < P_.ncol; ++j)
      for (int p = P_.p[j]; p < P_.p[j + 1]; ++p) {
        int i = P_.i[p];
        y[i] += P_.x[p] * x[j];
        if (i != j) y[j] += P_.x[p] * x[i];
      }
  }

  // ---- scaling (OSQP scaling.c: scale_data / unscale_data) ----------------------------
  static void limit_scaling(double* v, int len) {
    for (int k = 0; k < len; ++k) {
      v[k] = v[k] < MIN_SCALING ? 1.0 : v[k];
      v[k] = v[k] > MAX_SCALING ? MAX_SCALING : v[k];
    }
  }
  void inf_norm_cols_sym_triu(const Csc& M, std::vector<double>& e) const {
    std::fill(e.begin(), e.end(), 0.0);
    for (int j = 0; j < M.ncol; ++j)
      for (int p = M.p[j]; p < M.p[j + 1]; ++p) {
        int i = M.i[p];
        double a = std::fabs(M.x[p]);
        if (a > e[j]) e[j] = a;
        if (i != j && a > e[i]) e[i] = a;
      }
  }
  void scale_data() {
    c = 1.0;
    std::fill(D.begin(), D.end(), 1.0); std::fill(Dinv.begin(), Dinv.end(), 1.0);
    std::fill(E.begin(), E.end(), 1.0); std::fill(Einv.begin(), Einv.end(), 1.0);
    for (int it = 0; it < settings.scaling; ++it) {
      // column inf-norms of KKT = [P A'; A 0]
      inf_norm_cols_sym_triu(P_, D_temp_);
      std::fill(D_temp_A_.begin(), D_temp_A_.end(), 0.0);
      std::fill(E_temp_.begin(), E_temp_.end(), 0.0);
      for (int j = 0; j < n; ++j)
        for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) {
          double a = std::fabs(A_.x[p]);
          if (a > D_temp_A_[j]) D_temp_A_[j] = a;
          if (a > E_temp_[A_.i[p]]) E_temp_[A_.i[p]] = a;
        }
      for (int j = 0; j < n; ++j) D_temp_[j] = std::max(D_temp_[j], D_temp_A_[j]);
      limit_scaling(D_temp_.data(), n);
      limit_scaling(E_temp_.data(), m);
      for (int j = 0; j < n; ++j) D_temp_[j] = 1.0 / std::sqrt(D_temp_[j]);
      for (int i = 0; i < m; ++i) E_temp_[i] = 1.0 / std::sqrt(E_temp_[i]);
      // P <- D P D ; A <- E A D ; q <- D q
      for (int j = 0; j < n; ++j)
        for (int p = P_.p[j]; p < P_.p[j + 1]; ++p) { P_.x[p] *= D_temp_[P_.i[p]]; }
      for (int j = 0; j < n; ++j)
        for (int p = P_.p[j]; p < P_.p[j + 1]; ++p) { P_.x[p] *= D_temp_[j]; }
      for (int j = 0; j < n; ++j)
        for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) { A_.x[p] *= E_temp_[A_.i[p]]; }
      for (int j = 0; j < n; ++j)
        for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) { A_.x[p] *= D_temp_[j]; }
      for (int j = 0; j < n; ++j) q_[j] *= D_temp_[j];
      for (int j = 0; j < n; ++j) D[j] *= D_temp_[j];
      for (int i = 0; i < m; ++i) E[i] *= E_temp_[i];
      // cost normalisation
      inf_norm_cols_sym_triu(P_, D_temp_);
      double c_temp = 0.0;
      for (int j = 0; j < n; ++j) c_temp += D_temp_[j];
      c_temp /= (double)n;
      double inf_norm_q = norm_inf(q_);
      limit_scaling(&inf_norm_q, 1);
      c_temp = std::max(c_temp, inf_norm_q);
      limit_scaling(&c_temp, 1);
      c_temp = 1.0 / c_temp;
      for (double& v : P_.x) v *= c_temp;
      for (double& v : q_) v *= c_temp;
      c *= c_temp;
    }
    cinv = 1.0 / c;
    for (int j = 0; j < n; ++j) Dinv[j] = 1.0 / D[j];
    for (int i = 0; i < m; ++i) Einv[i] = 1.0 / E[i];
    for (int i = 0; i < m; ++i) { l_[i] *= E[i]; u_[i] *= E[i]; }
  }
  void unscale_data() {
    for (double& v : P_.x) v *= cinv;
    for (int j = 0; j < n; ++j)
      for (int p = P_.p[j]; p < P_.p[j + 1]; ++p) P_.x[p] *= Dinv[P_.i[p]];
    for (int j = 0; j < n; ++j)
      for (int p = P_.p[j]; p < P_.p[j + 1]; ++p) P_.x[p] *= Dinv[j];
    for (int j = 0; j < n; ++j) q_[j] *= cinv;
    for (int j = 0; j < n; ++j) q_[j] *= Dinv[j];
    for (int j = 0; j < n; ++j)
      for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) A_.x[p] *= Einv[A_.i[p]];
    for (int j = 0; j < n; ++j)
      for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) A_.x[p] *= Dinv[j];
    for (int i = 0; i < m; ++i) { l_[i] *= Einv[i]; u_[i] *= Einv[i]; }
  }

  // ---- rho vector (OSQP auxil.c: set_rho_vec / update_rho_vec) ----------------------
  void set_rho_vec() {
    settings.rho = std::min(std::max(settings.rho, RHO_MIN), RHO_MAX);
    for (int i = 0; i < m; ++i) {
      if (l_[i] < -OSQP_INFTY * MIN_SCALING && u_[i] > OSQP_INFTY * MIN_SCALING) {
        constr_type_[i] = -1;
        rho_vec_[i] = RHO_MIN;
      } else if (u_[i] - l_[i] < RHO_TOL) {
        constr_type_[i] = 1;
        rho_vec_[i] = RHO_EQ_OVER_RHO_INEQ * settings.rho;
      } else {
        constr_type_[i] = 0;
        rho_vec_[i] = settings.rho;
      }
      rho_inv_vec_[i] = 1.0 / rho_vec_[i];
    }
  }
  int update_rho_vec() {
    bool changed = false;
    for (int i = 0; i < m; ++i) {
      if (l_[i] < -OSQP_INFTY * MIN_SCALING && u_[i] > OSQP_INFTY * MIN_SCALING) {
        if (constr_type_[i] != -1) {
          constr_type_[i] = -1; rho_vec_[i] = RHO_MIN; rho_inv_vec_[i] = 1.0 / RHO_MIN; changed = true;
        }
      } else if (u_[i] - l_[i] < RHO_TOL) {
        if (constr_type_[i] != 1) {
          constr_type_[i] = 1; rho_vec_[i] = RHO_EQ_OVER_RHO_INEQ * settings.rho;
          rho_inv_vec_[i] = 1.0 / rho_vec_[i]; changed = true;
        }
      } else {
        if (constr_type_[i] != 0) {
          constr_type_[i] = 0; rho_vec_[i] = settings.rho; rho_inv_vec_[i] = 1.0 / settings.rho; changed = true;
        }
      }
    }
    if (changed) {
      update_kkt_rho();
      if (!ldl_.numeric(K_)) return 1;
      info.n_factor++;
    }
    return 0;
  }
  int update_rho(double rho_new) {  // osqp_update_rho()
    settings.rho = std::min(std::max(rho_new, RHO_MIN), RHO_MAX);
    for (int i = 0; i < m; ++i) {
      if (constr_type_[i] == 0) {
        rho_vec_[i] = settings.rho; rho_inv_vec_[i] = 1.0 / settings.rho;
      } else if (constr_type_[i] == 1) {
        rho_vec_[i] = RHO_EQ_OVER_RHO_INEQ * settings.rho; rho_inv_vec_[i] = 1.0 / rho_vec_[i];
      }
    }
    update_kkt_rho();
    if (!ldl_.numeric(K_)) return 1;
    info.n_factor++;
    return 0;
  }

  // ---- KKT assembly under a permutation ------------------------------------------------
  void build_kkt_structure(const int* perm) {
    const int N = n + m;
    perm_.resize(N); pinv_.resize(N);
    for (int k = 0; k < N; ++k) perm_[k] = perm ? perm[k] : k;
    for (int k = 0; k < N; ++k) pinv_[perm_[k]] = k;
    struct Ent { int r, c, src, idx; };
    std::vector<Ent> ents;
    // P (upper) + sigma on the diagonal; diagonal of P may be structurally absent
    std::vector<char> has_diag(n, 0);
    P_is_diag_.assign(P_.nnz(), 0);
    for (int j = 0; j < n; ++j)
      for (int p = P_.p[j]; p < P_.p[j + 1]; ++p) {
        int i = P_.i[p];
        if (i == j) { has_diag[j] = 1; P_is_diag_[p] = 1; }
        int a = pinv_[i], b = pinv_[j];
        ents.push_back({std::min(a, b), std::max(a, b), 0, p});
      }
    for (int j = 0; j < n; ++j) ents.push_back({pinv_[j], pinv_[j], 1, j});  // sigma
    for (int j = 0; j < n; ++j)
      for (int p = A_.p[j]; p < A_.p[j + 1]; ++p) {
        int a = pinv_[n + A_.i[p]], b = pinv_[j];
        ents.push_back({std::min(a, b), std::max(a, b), 2, p});
      }
    for (int i = 0; i < m; ++i) ents.push_back({pinv_[n + i], pinv_[n + i], 3, i});  // -1/rho
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) {
      return a.c != b.c ? a.c < b.c : a.r < b.r;
    });
    K_.nrow = K_.ncol = N;
    K_.p.assign(N + 1, 0); K_.i.clear(); K_.x.clear();
    P_to_K_.assign(P_.nnz(), -1); A_to_K_.assign(A_.nnz(), -1);
    sig_to_K_.assign(n, -1); rho_to_K_.assign(m, -1);
    int last_r = -1, last_c = -1;
    for (const Ent& e : ents) {
      if (e.r != last_r || e.c != last_c) {
        K_.i.push_back(e.r); K_.x.push_back(0.0); K_.p[e.c + 1]++;
        last_r = e.r; last_c = e.c;
      }
      int pos = (int)K_.i.size() - 1;
      if (e.src == 0) P_to_K_[e.idx] = pos;
      else if (e.src == 1) sig_to_K_[e.idx] = pos;
      else if (e.src == 2) A_to_K_[e.idx] = pos;
      else rho_to_K_[e.idx] = pos;
    }
    for (int k = 0; k < N; ++k) K_.p[k + 1] += K_.p[k];
    rhs_perm_.assign(N, 0.0);
  }
  void fill_kkt_values() {
    std::fill(K_.x.begin(), K_.x.end(), 0.0);
    for (int p = 0; p < P_.nnz(); ++p) K_.x[P_to_K_[p]] += P_.x[p];
    for (int j = 0; j < n; ++j) K_.x[sig_to_K_[j]] += settings.sigma;
    for (int p = 0; p < A_.nnz(); ++p) K_.x[A_to_K_[p]] += A_.x[p];
    for (int i = 0; i < m; ++i) K_.x[rho_to_K_[i]] = -rho_inv_vec_[i];
  }
  void update_kkt_rho() {
    for (int i = 0; i < m; ++i) K_.x[rho_to_K_[i]] = -rho_inv_vec_[i];
  }
  void kkt_solve(std::vector<double>& b) {
    const int N = n + m;
    for (int k = 0; k < N; ++k) rhs_perm_[k] = b[perm_[k]];
    ldl_.solve(rhs_perm_.data());
    for (int k = 0; k < N; ++k) b[perm_[k]] = rhs_perm_[k];
  }

  // ---- ADMM steps (OSQP auxil.c) -------------------------------------------------------
  void update_xz_tilde() {
    for (int j = 0; j < n; ++j) xz_tilde_[j] = settings.sigma * x_prev_[j] - q_[j];
    for (int i = 0; i < m; ++i) xz_tilde_[n + i] = z_prev_[i] - rho_inv_vec_[i] * y_[i];
    kkt_solve(xz_tilde_);
    for (int i = 0; i < m; ++i)
      xz_tilde_[n + i] = z_prev_[i] + rho_inv_vec_[i] * (xz_tilde_[n + i] - y_[i]);
  }
  void update_x() {
    const double a = settings.alpha;
    for (int j = 0; j < n; ++j) x_[j] = a * xz_tilde_[j] + (1.0 - a) * x_prev_[j];
    for (int j = 0; j < n; ++j) delta_x_[j] = x_[j] - x_prev_[j];
  }
  void update_z() {
    const double a = settings.alpha;
    for (int i = 0; i < m; ++i)
      z_[i] = a * xz_tilde_[n + i] + (1.0 - a) * z_prev_[i] + rho_inv_vec_[i] * y_[i];
    for (int i = 0; i < m; ++i) z_[i] = std::min(std::max(z_[i], l_[i]), u_[i]);
  }
  void update_y() {
    const double a = settings.alpha;
    for (int i = 0; i < m; ++i) {
      delta_y_[i] = rho_vec_[i] * (a * xz_tilde_[n + i] + (1.0 - a) * z_prev_[i] - z_[i]);
      y_[i] += delta_y_[i];
    }
  }

  // ---- residuals / termination ------------------------------------------------------------
  double compute_pri_res() {
    mat_vec_A(x_, Ax_);
    for (int i = 0; i < m; ++i) z_prev_[i] = Ax_[i] - z_[i];
    if (settings.scaling && !settings.scaled_termination) return scaled_norm_inf(Einv, z_prev_);
    return norm_inf(z_prev_);
  }
  double compute_dua_res() {
    mat_vec_P(x_, Px_);
    mat_tvec_A(y_, Aty_);
    for (int j = 0; j < n; ++j) x_prev_[j] = q_[j] + Px_[j] + Aty_[j];
    if (settings.scaling && !settings.scaled_termination) return cinv * scaled_norm_inf(Dinv, x_prev_);
    return norm_inf(x_prev_);
  }
  double compute_pri_tol(double eps_abs, double eps_rel) const {
    double mx;
    if (settings.scaling && !settings.scaled_termination)
      mx = std::max(scaled_norm_inf(Einv, z_), scaled_norm_inf(Einv, Ax_));
    else
      mx = std::max(norm_inf(z_), norm_inf(Ax_));
    return eps_abs + eps_rel * mx;
  }
  double compute_dua_tol(double eps_abs, double eps_rel) const {
    double mx;
    if (settings.scaling && !settings.scaled_termination) {
      mx = scaled_norm_inf(Dinv, q_);
      mx = std::max(mx, scaled_norm_inf(Dinv, Aty_));
      mx = std::max(mx, scaled_norm_inf(Dinv, Px_));
      mx *= cinv;
    } else {
      mx = std::max(norm_inf(q_), std::max(norm_inf(Aty_), norm_inf(Px_)));
    }
    return eps_abs + eps_rel * mx;
  }
  void update_info(int iter) {
    info.iter = iter;
    // objective
    double obj = 0.0;
    mat_vec_P(x_, Px_);
    for (int j = 0; j < n; ++j) obj += 0.5 * x_[j] * Px_[j] + q_[j] * x_[j];
    if (settings.scaling) obj *= cinv;
    info.obj_val = obj;
    info.pri_res = m ? compute_pri_res() : 0.0;
    info.dua_res = compute_dua_res();
  }
  bool is_primal_infeasible(double eps_prim_inf) {
    // project delta_y onto the polar of the recession cone of [l, u]
    for (int i = 0; i < m; ++i) {
      if (u_[i] > OSQP_INFTY * MIN_SCALING) {
        if (l_[i] < -OSQP_INFTY * MIN_SCALING) delta_y_[i] = 0.0;
        else delta_y_[i] = std::min(delta_y_[i], 0.0);
      } else if (l_[i] < -OSQP_INFTY * MIN_SCALING) {
        delta_y_[i] = std::max(delta_y_[i], 0.0);
      }
    }
    double norm_delta_y;
    if (settings.scaling && !settings.scaled_termination) norm_delta_y = scaled_norm_inf(E, delta_y_);
    else norm_delta_y = norm_inf(delta_y_);
    if (norm_delta_y > eps_prim_inf) {
      double ineq_lhs = 0.0;
      for (int i = 0; i < m; ++i)
        ineq_lhs += u_[i] * std::max(delta_y_[i], 0.0) + l_[i] * std::min(delta_y_[i], 0.0);
      if (ineq_lhs < -eps_prim_inf * norm_delta_y) {
        mat_tvec_A(delta_y_, Atdelta_y_);
        double nr = (settings.scaling && !settings.scaled_termination) ? scaled_norm_inf(Dinv, Atdelta_y_)
                                                                       : norm_inf(Atdelta_y_);
        return nr < eps_prim_inf * norm_delta_y;
      }
    }
    return false;
  }
  bool is_dual_infeasible(double eps_dual_inf) {
    double norm_delta_x, cost_scaling;
    if (settings.scaling && !settings.scaled_termination) {
      norm_delta_x = scaled_norm_inf(D, delta_x_);
      cost_scaling = c;
    } else {
      norm_delta_x = norm_inf(delta_x_);
      cost_scaling = 1.0;
    }
    if (norm_delta_x > eps_dual_inf) {
      double qdx = 0.0;
      for (int j = 0; j < n; ++j) qdx += q_[j] * delta_x_[j];
      if (qdx < -cost_scaling * eps_dual_inf * norm_delta_x) {
        mat_vec_P(delta_x_, Pdelta_x_);
        double np = (settings.scaling && !settings.scaled_termination) ? scaled_norm_inf(Dinv, Pdelta_x_)
                                                                       : norm_inf(Pdelta_x_);
        if (np < cost_scaling * eps_dual_inf * norm_delta_x) {
          mat_vec_A(delta_x_, Adelta_x_);
          if (settings.scaling && !settings.scaled_termination)
            for (int i = 0; i < m; ++i) Adelta_x_[i] *= Einv[i];
          for (int i = 0; i < m; ++i) {
            if ((u_[i] < OSQP_INFTY * MIN_SCALING && Adelta_x_[i] > eps_dual_inf * norm_delta_x) ||
                (l_[i] > -OSQP_INFTY * MIN_SCALING && Adelta_x_[i] < -eps_dual_inf * norm_delta_x))
              return false;
          }
          return true;
        }
      }
    }
    return false;
  }
  bool check_termination(bool approximate) {
    double eps_abs = settings.eps_abs, eps_rel = settings.eps_rel;
    double eps_prim_inf = settings.eps_prim_inf, eps_dual_inf = settings.eps_dual_inf;
    bool prim_res_check = false, dual_res_check = false, prim_inf_check = false, dual_inf_check = false;
    if (info.pri_res > OSQP_INFTY || info.dua_res > OSQP_INFTY) {
      info.status = NON_CVX;
      info.obj_val = std::numeric_limits<double>::quiet_NaN();
      return true;
    }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_prim_inf *= 10; eps_dual_inf *= 10; }
    if (m == 0) {
      prim_res_check = true;
    } else {
      double eps_prim = compute_pri_tol(eps_abs, eps_rel);
      if (info.pri_res < eps_prim) prim_res_check = true;
      else prim_inf_check = is_primal_infeasible(eps_prim_inf);
    }
    double eps_dual = compute_dua_tol(eps_abs, eps_rel);
    if (info.dua_res < eps_dual) dual_res_check = true;
    else dual_inf_check = is_dual_infeasible(eps_dual_inf);
    if (prim_res_check && dual_res_check) {
      info.status = approximate ? SOLVED_INACCURATE : SOLVED;
      return true;
    }
    if (prim_inf_check) {
      info.status = approximate ? PRIMAL_INFEASIBLE_INACCURATE : PRIMAL_INFEASIBLE;
      info.obj_val = OSQP_INFTY;
      return true;
    }
    if (dual_inf_check) {
      info.status = approximate ? DUAL_INFEASIBLE_INACCURATE : DUAL_INFEASIBLE;
      info.obj_val = -OSQP_INFTY;
      return true;
    }
    return false;
  }

  // ---- adaptive rho -----------------------------------------------------------------------
  double compute_rho_estimate() const {
    // z_prev_ holds (Ax - z), x_prev_ holds (q + Px + A'y): scaled residual vectors
    double pri_res = norm_inf(z_prev_);
    double dua_res = norm_inf(x_prev_);
    double prim_norm = std::max(norm_inf(z_), norm_inf(Ax_));
    pri_res /= (prim_norm + 1e-10);
    double dual_norm = std::max(norm_inf(q_), std::max(norm_inf(Aty_), norm_inf(Px_)));
    dua_res /= (dual_norm + 1e-10);
    double est = settings.rho * std::sqrt(pri_res / (dua_res + 1e-10));
    return std::min(std::max(est, RHO_MIN), RHO_MAX);
  }
  int adapt_rho() {
    double rho_new = compute_rho_estimate();
    info.rho_estimate = rho_new;
    if (rho_new > settings.rho * settings.adaptive_rho_tolerance ||
        rho_new < settings.rho / settings.adaptive_rho_tolerance) {
      int e = update_rho(rho_new);
      info.rho_updates++;
      return e;
    }
    return 0;
  }

  void store_solution() {
    bool has_solution = !(info.status == PRIMAL_INFEASIBLE || info.status == PRIMAL_INFEASIBLE_INACCURATE ||
                          info.status == DUAL_INFEASIBLE || info.status == DUAL_INFEASIBLE_INACCURATE ||
                          info.status == NON_CVX);
    if (has_solution) {
      for (int j = 0; j < n; ++j) sol_x[j] = settings.scaling ? D[j] * x_[j] : x_[j];
      for (int i = 0; i < m; ++i) sol_y[i] = settings.scaling ? cinv * (E[i] * y_[i]) : y_[i];
    } else {
      const double nan = std::numeric_limits<double>::quiet_NaN();
      std::fill(sol_x.begin(), sol_x.end(), nan);
      std::fill(sol_y.begin(), sol_y.end(), nan);
      cold_start();  // OSQP resets the iterates so a later warm start is sane
    }
  }
};

}  // namespace osqp_restated
