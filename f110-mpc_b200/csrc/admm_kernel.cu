// Batched OSQP-ADMM solve of the f110-mpc tracking QP — one warp per QP, one horizon stage per lane.
//
// What it replaces: the OsqpEigen/OSQP call sequence of MPC::Update (reference src/mpc.cpp:81-142) for
// the QP that MPC::MPC / Create*/Update* build (mpc.cpp:26-35, 208-306) from Model::Linearize
// (model.cpp:30-59).  The algorithm is OSQP's (Ruiz equilibration, rho vector by row class, relaxed
// ADMM, residual test every `check_termination` iterations, adaptive rho) — the same iterates as the
// CPU oracle in exact arithmetic — but the linear algebra is re-derived for the horizon structure:
//
//  * The scaled problem (D, E, c from Ruiz) is iterated in UNSCALED coordinates with diagonal metrics
//        sigma_j = sigma / (c d_j^2),   rho_i = rho_bar_i e_i^2 / c,
//    which is algebraically the same iteration, and keeps the dynamics LTI (A, B are 6 numbers).
//  * The quasi-definite KKT solve  [P+Sigma, A'; A, -1/rho] is condensed: z~ = A x~ exactly, so each
//    iteration solves (P + Sigma + A' R A) w = Sigma w_prev - q + A'(R z_prev - y).  Inputs u_k are
//    eliminated per stage (2x2), leaving a symmetric block-tridiagonal system in x_0..x_N with 3x3 blocks.
//  * That system is solved by parallel cyclic reduction across the lanes of the warp: log2(N+1) levels,
//    every lane busy at every level.  The PCR multipliers are computed once per rho (factor step) and kept
//    in shared memory; each iteration only applies them to the right-hand side.
//  * x/z/y update, projection onto [l,u], residual norms and the termination test are fused in the same
//    kernel; all reductions are warp shuffles.  Nothing but the parameter record is read from HBM and
//    nothing but the solution is written.
//
// Lane k owns stage k: x_k(3), u_k(2) (k<N), dynamics rows k (3), gap rows k (2), input-box rows k (2).
#include "admm_kernel.cuh"

namespace f110 {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr double OSQP_INFTY = 1e30;
constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ_OVER_RHO_INEQ = 1e3, RHO_TOL = 1e-4;
constexpr double MIN_SCALING = 1e-4, MAX_SCALING = 1e4;
constexpr double INF_THRESH = OSQP_INFTY * MIN_SCALING;  // 1e26
constexpr double HUGE_BOUND = 1e300;

constexpr int MAX_LEVELS = 5;
// shared memory per warp, in doubles (element-major, lane fastest -> conflict free)
constexpr int SM_COEF = MAX_LEVELS * 18 * 32;  // alpha(9), gamma(9) per level
constexpr int SM_BINV = 6 * 32;                // final symmetric block inverse
constexpr int SM_DE = 12 * 32;                 // dx3 du2 ed3 eg2 eb2 (only read at checks)
constexpr int SM_PREV = 12 * 32;               // x,u,y before the last iteration (infeasibility tests)
constexpr int SM_PER_WARP = SM_COEF + SM_BINV + SM_DE + SM_PREV;

enum : int {
  ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF_INACC = 3, ST_DINF_INACC = 4,
  ST_MAX_ITER = -2, ST_PINF = -3, ST_DINF = -4, ST_NON_CVX = -7, ST_UNSOLVED = -10
};

__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  return v > MAX_SCALING ? MAX_SCALING : v;
}
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// ---- 3x3 helpers (row-major double[9]) ------------------------------------------------------------
__device__ __forceinline__ void mm3(const double* a, const double* b, double* c) {  // c = a b
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// inverse of a symmetric positive definite 3x3 via LDL^T (reads the lower triangle)
__device__ __forceinline__ void inv_spd3(const double* a, double* inv) {
  const double d0 = a[0], i0 = 1.0 / d0;
  const double l10 = a[3] * i0, l20 = a[6] * i0;
  const double d1 = a[4] - l10 * a[3], i1 = 1.0 / d1;
  const double l21 = (a[7] - l20 * a[3]) * i1;
  const double d2 = a[8] - l20 * a[6] - l21 * (a[7] - l20 * a[3]), i2 = 1.0 / d2;
  // Linv = [[1,0,0],[-l10,1,0],[l10*l21-l20,-l21,1]]
  const double m10 = -l10, m20 = l10 * l21 - l20, m21 = -l21;
  // inv = Linv' Dinv Linv
  inv[0] = i0 + m10 * m10 * i1 + m20 * m20 * i2;
  inv[1] = inv[3] = m10 * i1 + m20 * m21 * i2;
  inv[2] = inv[6] = m20 * i2;
  inv[4] = i1 + m21 * m21 * i2;
  inv[5] = inv[7] = m21 * i2;
  inv[8] = i2;
}

struct Model {  // Model::Linearize output (model.cpp:30-59): A = I + [0 0 a02; 0 0 a12; 0 0 0], B = [b00 0; b10 0; b20 b21]
  double a02, a12, b00, b10, b20, b21;
};
__device__ __forceinline__ void A_mul(const Model& m, const double* v, double* o) {   // o = A v
  o[0] = v[0] + m.a02 * v[2]; o[1] = v[1] + m.a12 * v[2]; o[2] = v[2];
}
__device__ __forceinline__ void At_mul(const Model& m, const double* v, double* o) {  // o = A' v
  o[0] = v[0]; o[1] = v[1]; o[2] = m.a02 * v[0] + m.a12 * v[1] + v[2];
}
__device__ __forceinline__ void B_mul(const Model& m, const double* h, double* o) {   // o = B h
  o[0] = m.b00 * h[0]; o[1] = m.b10 * h[0]; o[2] = m.b20 * h[0] + m.b21 * h[1];
}
__device__ __forceinline__ void Bt_mul(const Model& m, const double* v, double* o) {  // o = B' v
  o[0] = m.b00 * v[0] + m.b10 * v[1] + m.b20 * v[2]; o[1] = m.b21 * v[2];
}

// Everything one lane keeps for its stage.
struct Stage {
  // problem data
  double bd[3];            // dynamics rhs (l = u): -x_cur at k = 0, -C at k >= 1      (mpc.cpp:299,305)
  double gm[6];            // gap rows 2x3: ones at k = 0, [l1a l1b 0; l2a l2b 0] after (mpc.cpp:237-241, 260-272)
  double gl[2], gu[2];     // gap bounds                                              (mpc.cpp:279-300)
  double bl[2], bu[2];     // input box                                               (mpc.cpp:281,290)
  double qx[3];            // -Q ref_k                                                 (mpc.cpp:225,228)
  // iterates (unscaled)
  double x[3], u[2];
  double zd[3], zg[2], zb[2];
  double yd[3], yg[2], yb[2];
  // metric
  double sx[3], su[2];                 // sigma_j
  double rd[3], rg[2], rb[2];          // rho_i
  double id[3], ig[2], ib[2];          // 1 / rho_i
  // input elimination
  double wi[3];            // inverse of W_k = R + Sigma_u + rho_box + B' R_{k+1} B   (00, 01, 11)
  double rdn[3];           // rho of the NEXT stage's dynamics rows
};

}  // namespace

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) admm_kernel(const KParams p) {
  extern __shared__ double smem_all[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int qp = blockIdx.x * WARPS + warp;
  if (qp >= p.B) return;
  double* sm_coef = smem_all + (size_t)warp * SM_PER_WARP;
  double* sm_binv = sm_coef + SM_COEF;
  double* sm_de = sm_binv + SM_BINV;
  double* sm_prev = sm_de + SM_DE;

  const int N = p.N;
  const int k = lane;
  const bool act = k <= N;   // lane owns a stage
  const bool actu = k < N;   // stage has an input (and box rows, and a successor)
  const bool hasp = act && k > 0;  // stage has a predecessor (lanes above N must stay identically zero)
  const int nvar = 5 * N + 3;
  int nlev = 0;
  while ((1 << nlev) <= N && nlev < MAX_LEVELS) ++nlev;

  // ---------------- load the parameter record, linearise, stack -----------------------------------------
  const double* rec = p.recs + (size_t)qp * p.stride;
  const double x0[3] = {rec[0], rec[1], rec[2]};
  const double vlin = rec[3], slin = rec[4];
  Model md;
  double Cv[3];
  {
    // Model::Linearize, model.cpp:42-55 (operation order kept)
    const double L = p.wheelbase, dt = p.dt;
    const double so = sin(x0[2]), co = cos(x0[2]);
    const double cs = cos(slin);
    const double pw = pow(cs, -2.0);
    md.a02 = -1.0 * vlin * so * dt;
    md.a12 = vlin * co * dt;
    md.b00 = co * dt;
    md.b10 = so * dt;
    md.b20 = tan(slin) * dt / L;
    md.b21 = vlin * pw * dt / L;
    Cv[0] = vlin * x0[2] * so * dt;
    Cv[1] = -1.0 * vlin * x0[2] * co * dt;
    Cv[2] = -1.0 * slin * vlin * pw * dt / L;
  }
  const double aA[9] = {1.0, 0.0, fabs(md.a02), 0.0, 1.0, fabs(md.a12), 0.0, 0.0, 1.0};
  const double aB[6] = {fabs(md.b00), 0.0, fabs(md.b10), 0.0, fabs(md.b20), fabs(md.b21)};
  const double qu[2] = {-1.0 * p.R[0] * p.u_des[0], -1.0 * p.R[1] * p.u_des[1]};  // mpc.cpp:226

  Stage s;
  {
    const int kr = (k < N) ? k : (N - 1);  // terminal stage re-uses ref[N-1] (mpc.cpp:228)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double r = act ? rec[11 + 3 * kr + j] : 0.0;
      s.qx[j] = act ? (-1.0 * p.Q[j] * r) : 0.0;
      s.bd[j] = act ? (k == 0 ? -x0[j] : -Cv[j]) : 0.0;
    }
    const double l1a = rec[5], l1b = rec[6], l1c = rec[7], l2a = rec[8], l2b = rec[9], l2c = rec[10];
    if (!act) {
#pragma unroll
      for (int e = 0; e < 6; ++e) s.gm[e] = 0.0;
    } else if (k == 0) {
#pragma unroll
      for (int e = 0; e < 6; ++e) s.gm[e] = 1.0;
    } else {
      s.gm[0] = l1a; s.gm[1] = l1b; s.gm[2] = 0.0;
      s.gm[3] = l2a; s.gm[4] = l2b; s.gm[5] = 0.0;
    }
    s.gl[0] = (p.gap_mode && act) ? -l1c : -OSQP_INFTY;  // mpc.cpp:297 (commented alternative when gap_mode = 1)
    s.gl[1] = (p.gap_mode && act) ? -l2c : -OSQP_INFTY;  // mpc.cpp:298
    s.gu[0] = OSQP_INFTY; s.gu[1] = OSQP_INFTY;          // mpc.cpp:288-290
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      s.bl[j] = actu ? p.u_min[j] : -HUGE_BOUND;
      s.bu[j] = actu ? p.u_max[j] : HUGE_BOUND;
    }
  }

  // ---------------- Ruiz equilibration (OSQP scale_data) -------------------------------------------------
  double dx[3] = {1, 1, 1}, du[2] = {1, 1}, ed[3] = {1, 1, 1}, eg[2] = {1, 1}, eb[2] = {1, 1};
  double c = 1.0;
  for (int it = 0; it < p.scaling; ++it) {
    double edn[3], dxp[3], dup[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) { edn[i] = __shfl_down_sync(FULL, ed[i], 1); edn[i] = actu ? edn[i] : 0.0; }
#pragma unroll
    for (int j = 0; j < 3; ++j) { dxp[j] = __shfl_up_sync(FULL, dx[j], 1); dxp[j] = (act && k > 0) ? dxp[j] : 0.0; }
#pragma unroll
    for (int j = 0; j < 2; ++j) { dup[j] = __shfl_up_sync(FULL, du[j], 1); dup[j] = (act && k > 0) ? dup[j] : 0.0; }
    double tx[3], tu[2], td[3], tg[2], tb[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {  // KKT column of x_k[j]: P, the -1 of dyn row k, A of dyn rows k+1, gap rows k
      double v = c * dx[j] * dx[j] * p.Q[j];
      v = fmax(v, ed[j] * dx[j]);
#pragma unroll
      for (int i = 0; i < 3; ++i) v = fmax(v, edn[i] * aA[3 * i + j] * dx[j]);
#pragma unroll
      for (int r = 0; r < 2; ++r) v = fmax(v, eg[r] * fabs(s.gm[3 * r + j]) * dx[j]);
      tx[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {  // KKT column of u_k[j]
      double v = c * du[j] * du[j] * p.R[j];
#pragma unroll
      for (int i = 0; i < 3; ++i) v = fmax(v, edn[i] * aB[2 * i + j] * du[j]);
      v = fmax(v, eb[j] * du[j]);
      tu[j] = v;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // KKT column (= A row) of dyn row (k, i)
      double v = ed[i] * dx[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) v = fmax(v, ed[i] * aA[3 * i + j] * dxp[j]);
#pragma unroll
      for (int j = 0; j < 2; ++j) v = fmax(v, ed[i] * aB[2 * i + j] * dup[j]);
      td[i] = v;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) v = fmax(v, eg[r] * fabs(s.gm[3 * r + j]) * dx[j]);
      tg[r] = v;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) tb[j] = eb[j] * du[j];
    if (!act) {  // lanes above N: keep D = E = 1
#pragma unroll
      for (int j = 0; j < 3; ++j) { tx[j] = 1.0; td[j] = 1.0; }
#pragma unroll
      for (int j = 0; j < 2; ++j) { tg[j] = 1.0; }
    }
    if (!actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) { tu[j] = 1.0; tb[j] = 1.0; }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) dx[j] *= rsqrt(limit_scaling(tx[j]));
#pragma unroll
    for (int j = 0; j < 2; ++j) du[j] *= rsqrt(limit_scaling(tu[j]));
#pragma unroll
    for (int i = 0; i < 3; ++i) ed[i] *= rsqrt(limit_scaling(td[i]));
#pragma unroll
    for (int r = 0; r < 2; ++r) eg[r] *= rsqrt(limit_scaling(tg[r]));
#pragma unroll
    for (int j = 0; j < 2; ++j) eb[j] *= rsqrt(limit_scaling(tb[j]));
    // cost normalisation: c_temp = 1 / max(mean column norm of P, ||q||_inf)
    double psum = 0.0, qn = 0.0;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { psum += c * dx[j] * dx[j] * p.Q[j]; qn = fmax(qn, fabs(dx[j] * s.qx[j])); }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) { psum += c * du[j] * du[j] * p.R[j]; qn = fmax(qn, fabs(du[j] * qu[j])); }
    }
    const double mean = wsum(psum) / (double)nvar;
    const double qinf = limit_scaling(c * wmax(qn));
    const double ct = limit_scaling(fmax(mean, qinf));
    c *= 1.0 / ct;
  }
  const double cinv = 1.0 / c;
  // park D, E in shared memory: only the rho estimate and the state store read them again
#pragma unroll
  for (int j = 0; j < 3; ++j) sm_de[(0 + j) * 32 + lane] = dx[j];
#pragma unroll
  for (int j = 0; j < 2; ++j) sm_de[(3 + j) * 32 + lane] = du[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) sm_de[(5 + j) * 32 + lane] = ed[j];
#pragma unroll
  for (int j = 0; j < 2; ++j) sm_de[(8 + j) * 32 + lane] = eg[j];
#pragma unroll
  for (int j = 0; j < 2; ++j) sm_de[(10 + j) * 32 + lane] = eb[j];

  // ---------------- row classes (OSQP set_rho_vec, on the SCALED bounds) and constant metric ---------------
  // class: 1 equality (1e3 rho), 0 inequality (rho), -1 loose (RHO_MIN).  Dynamics rows have l = u -> 1.
  int cls_g[2], cls_b[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double lb = eg[r] * s.gl[r], ub = eg[r] * s.gu[r];
    cls_g[r] = (lb < -INF_THRESH && ub > INF_THRESH) ? -1 : ((ub - lb < RHO_TOL) ? 1 : 0);
    const double lbb = eb[r] * s.bl[r], ubb = eb[r] * s.bu[r];
    cls_b[r] = (lbb < -INF_THRESH && ubb > INF_THRESH) ? -1 : ((ubb - lbb < RHO_TOL) ? 1 : 0);
  }
  double wd[3], wg[2], wb[2];  // e_i^2 / c : rho_i = rho_bar_i * w_i
#pragma unroll
  for (int i = 0; i < 3; ++i) wd[i] = ed[i] * ed[i] * cinv;
#pragma unroll
  for (int r = 0; r < 2; ++r) { wg[r] = eg[r] * eg[r] * cinv; wb[r] = eb[r] * eb[r] * cinv; }
#pragma unroll
  for (int j = 0; j < 3; ++j) s.sx[j] = p.sigma * cinv / (dx[j] * dx[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) s.su[j] = p.sigma * cinv / (du[j] * du[j]);
  // constant norms of q (unscaled and scaled)
  double nq, snq;
  {
    double a = 0.0, b = 0.0;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { a = fmax(a, fabs(s.qx[j])); b = fmax(b, fabs(dx[j] * s.qx[j])); }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) { a = fmax(a, fabs(qu[j])); b = fmax(b, fabs(du[j] * qu[j])); }
    }
    nq = wmax(a);
    snq = c * wmax(b);
  }

  // ---------------- iterates: cold start or the slot's stored (scaled) iterates -----------------------------
  double rho_bar = fmin(fmax(p.rho0, RHO_MIN), RHO_MAX);
  const int mcon = 7 * N + 5;
  double* slot = p.state ? p.state + (size_t)qp * state_doubles(N) : nullptr;
  bool warm = false;
  if (slot && p.warm_start) warm = slot[nvar + 2 * mcon + 1] != 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) { s.x[j] = 0; s.zd[j] = 0; s.yd[j] = 0; }
#pragma unroll
  for (int j = 0; j < 2; ++j) { s.u[j] = 0; s.zg[j] = 0; s.zb[j] = 0; s.yg[j] = 0; s.yb[j] = 0; }
  if (warm) {
    // OSQP keeps x, z, y in SCALED coordinates across re-scalings (osqp_update_A rescales the data only):
    // x = D xbar, z = zbar / E, y = E ybar / c with the NEW D, E, c.
    rho_bar = slot[nvar + 2 * mcon];
    const double* sx_ = slot;
    const double* sz_ = slot + nvar;
    const double* sy_ = slot + nvar + mcon;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        s.x[j] = dx[j] * sx_[3 * k + j];
        s.zd[j] = sz_[3 * k + j] / ed[j];
        s.yd[j] = ed[j] * sy_[3 * k + j] * cinv;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        s.zg[r] = sz_[3 * (N + 1) + 2 * k + r] / eg[r];
        s.yg[r] = eg[r] * sy_[3 * (N + 1) + 2 * k + r] * cinv;
      }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        s.u[j] = du[j] * sx_[3 * (N + 1) + 2 * k + j];
        s.zb[j] = sz_[5 * (N + 1) + 2 * k + j] / eb[j];
        s.yb[j] = eb[j] * sy_[5 * (N + 1) + 2 * k + j] * cinv;
      }
    }
  }

  // ---------------- factor step: metric from rho_bar, input elimination, PCR multipliers --------------------
  auto factor = [&]() {
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.rd[i] = RHO_EQ_OVER_RHO_INEQ * rho_bar * wd[i]; s.id[i] = 1.0 / s.rd[i]; }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double rg = cls_g[r] < 0 ? RHO_MIN : (cls_g[r] > 0 ? RHO_EQ_OVER_RHO_INEQ * rho_bar : rho_bar);
      const double rb = cls_b[r] < 0 ? RHO_MIN : (cls_b[r] > 0 ? RHO_EQ_OVER_RHO_INEQ * rho_bar : rho_bar);
      s.rg[r] = rg * wg[r]; s.ig[r] = 1.0 / s.rg[r];
      s.rb[r] = rb * wb[r]; s.ib[r] = 1.0 / s.rb[r];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double t = __shfl_down_sync(FULL, s.rd[i], 1); s.rdn[i] = actu ? t : 0.0; }
    // W_k and its inverse
    double Rn[9];  // R~_{k+1} = diag(rdn) - (rdn.B) W^-1 (rdn.B)'
    {
      const double w00 = p.R[0] + s.su[0] + s.rb[0] + md.b00 * md.b00 * s.rdn[0] + md.b10 * md.b10 * s.rdn[1] + md.b20 * md.b20 * s.rdn[2];
      const double w01 = md.b20 * md.b21 * s.rdn[2];
      const double w11 = p.R[1] + s.su[1] + s.rb[1] + md.b21 * md.b21 * s.rdn[2];
      const double idet = 1.0 / (w00 * w11 - w01 * w01);
      s.wi[0] = actu ? w11 * idet : 0.0;
      s.wi[1] = actu ? -w01 * idet : 0.0;
      s.wi[2] = actu ? w00 * idet : 0.0;
      const double M[6] = {s.rdn[0] * md.b00, 0.0, s.rdn[1] * md.b10, 0.0, s.rdn[2] * md.b20, s.rdn[2] * md.b21};
      double MW[6];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        MW[2 * i] = M[2 * i] * s.wi[0] + M[2 * i + 1] * s.wi[1];
        MW[2 * i + 1] = M[2 * i] * s.wi[1] + M[2 * i + 1] * s.wi[2];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int l = 0; l < 3; ++l) Rn[3 * i + l] = (i == l ? s.rdn[i] : 0.0) - (MW[2 * i] * M[2 * l] + MW[2 * i + 1] * M[2 * l + 1]);
    }
    double Rt[9];  // R~_k: from lane k-1, or diag(rho_d) for the x_0 = x_cur rows
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const double t = __shfl_up_sync(FULL, Rn[e], 1);
      Rt[e] = (k == 0) ? ((e % 4 == 0) ? s.rd[e / 4] : 0.0) : (act ? t : 0.0);
    }
    double Bm[9], Lm[9], Um[9];
    {
      // Hx = diag(Q + sigma_x) + G' diag(rho_g) G
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int l = 0; l < 3; ++l)
          Bm[3 * i + l] = (i == l ? p.Q[i] + s.sx[i] : 0.0) + s.rg[0] * s.gm[i] * s.gm[l] + s.rg[1] * s.gm[3 + i] * s.gm[3 + l] + Rt[3 * i + l];
      // A' Rn (3x3), then + A' Rn A
      double ARn[9];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        ARn[0 + l] = Rn[0 + l];
        ARn[3 + l] = Rn[3 + l];
        ARn[6 + l] = md.a02 * Rn[0 + l] + md.a12 * Rn[3 + l] + Rn[6 + l];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {  // (ARn A)[i][:] = ARn[i][:] A
        const double r0 = ARn[3 * i], r1 = ARn[3 * i + 1], r2 = ARn[3 * i + 2];
        Bm[3 * i + 0] += r0;
        Bm[3 * i + 1] += r1;
        Bm[3 * i + 2] += r0 * md.a02 + r1 * md.a12 + r2;
        Um[3 * i + 0] = -r0; Um[3 * i + 1] = -r1; Um[3 * i + 2] = -r2;
      }
      // L_k = -R~_k A
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double r0 = Rt[3 * i], r1 = Rt[3 * i + 1], r2 = Rt[3 * i + 2];
        Lm[3 * i + 0] = hasp ? -r0 : 0.0;
        Lm[3 * i + 1] = hasp ? -r1 : 0.0;
        Lm[3 * i + 2] = hasp ? -(r0 * md.a02 + r1 * md.a12 + r2) : 0.0;
      }
      if (!act) {
#pragma unroll
        for (int e = 0; e < 9; ++e) { Bm[e] = (e % 4 == 0) ? 1.0 : 0.0; Lm[e] = 0.0; Um[e] = 0.0; }
      }
    }
    // parallel cyclic reduction, log2 levels; multipliers alpha, gamma go to shared memory
    for (int lev = 0; lev < nlev; ++lev) {
      const int h = 1 << lev;
      const bool vlo = act && (k - h >= 0);
      const bool vhi = act && (k + h <= N);
      double Bi[9], XU[9], XL[9];
      inv_spd3(Bm, Bi);
      mm3(Bi, Um, XU);
      mm3(Bi, Lm, XL);
      double nb[9], t1[9], t2[9];
      double al[9], ga[9], Ln[9], Un[9];
      // neighbour k-h
#pragma unroll
      for (int e = 0; e < 9; ++e) nb[e] = __shfl_up_sync(FULL, Bi[e], h);
      mm3(Lm, nb, al);
#pragma unroll
      for (int e = 0; e < 9; ++e) nb[e] = __shfl_up_sync(FULL, XU[e], h);
      mm3(Lm, nb, t1);
#pragma unroll
      for (int e = 0; e < 9; ++e) nb[e] = __shfl_up_sync(FULL, XL[e], h);
      mm3(Lm, nb, Ln);
      // neighbour k+h
#pragma unroll
      for (int e = 0; e < 9; ++e) nb[e] = __shfl_down_sync(FULL, Bi[e], h);
      mm3(Um, nb, ga);
#pragma unroll
      for (int e = 0; e < 9; ++e) nb[e] = __shfl_down_sync(FULL, XL[e], h);
      mm3(Um, nb, t2);
#pragma unroll
      for (int e = 0; e < 9; ++e) nb[e] = __shfl_down_sync(FULL, XU[e], h);
      mm3(Um, nb, Un);
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        al[e] = vlo ? al[e] : 0.0;
        ga[e] = vhi ? ga[e] : 0.0;
        Bm[e] = Bm[e] - (vlo ? t1[e] : 0.0) - (vhi ? t2[e] : 0.0);
        Lm[e] = vlo ? -Ln[e] : 0.0;
        Um[e] = vhi ? -Un[e] : 0.0;
        sm_coef[(lev * 18 + e) * 32 + lane] = al[e];
        sm_coef[(lev * 18 + 9 + e) * 32 + lane] = ga[e];
      }
    }
    {
      double Bi[9];
      inv_spd3(Bm, Bi);
      sm_binv[0 * 32 + lane] = Bi[0]; sm_binv[1 * 32 + lane] = Bi[1]; sm_binv[2 * 32 + lane] = Bi[2];
      sm_binv[3 * 32 + lane] = Bi[4]; sm_binv[4 * 32 + lane] = Bi[5]; sm_binv[5 * 32 + lane] = Bi[8];
    }
    __syncwarp();
  };

  // ---------------- one ADMM iteration (OSQP update_xz_tilde / update_x / update_z / update_y) --------------
  auto iterate = [&]() {
    // s = rho (z - y/rho) = rho z - y per row; right-hand side of the condensed system
    double sd[3], sg[2], sb[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) sd[i] = s.rd[i] * s.zd[i] - s.yd[i];
#pragma unroll
    for (int r = 0; r < 2; ++r) { sg[r] = s.rg[r] * s.zg[r] - s.yg[r]; sb[r] = s.rb[r] * s.zb[r] - s.yb[r]; }
    double sdn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double t = __shfl_down_sync(FULL, sd[i], 1); sdn[i] = actu ? t : 0.0; }
    double gx[3], gu[2], t3[3];
    At_mul(md, sdn, t3);
#pragma unroll
    for (int j = 0; j < 3; ++j)
      gx[j] = s.sx[j] * s.x[j] - s.qx[j] - sd[j] + t3[j] + s.gm[j] * sg[0] + s.gm[3 + j] * sg[1];
    double t2[2];
    Bt_mul(md, sdn, t2);
#pragma unroll
    for (int j = 0; j < 2; ++j) gu[j] = s.su[j] * s.u[j] - qu[j] + t2[j] + sb[j];
    // eliminate u_k: h = W^-1 gu, f = R_{k+1} B h
    double hh[2] = {s.wi[0] * gu[0] + s.wi[1] * gu[1], s.wi[1] * gu[0] + s.wi[2] * gu[1]};
    double f[3];
    B_mul(md, hh, f);
#pragma unroll
    for (int i = 0; i < 3; ++i) f[i] *= s.rdn[i];
    double r[3];
    At_mul(md, f, t3);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double fp = __shfl_up_sync(FULL, f[i], 1);
      r[i] = gx[i] - t3[i] + (hasp ? fp : 0.0);
    }
    // PCR: apply the stored multipliers level by level
    for (int lev = 0; lev < nlev; ++lev) {
      const int h = 1 << lev;
      double lo[3], hi[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) { lo[i] = __shfl_up_sync(FULL, r[i], h); hi[i] = __shfl_down_sync(FULL, r[i], h); }
      const double* cf = sm_coef + (lev * 18) * 32 + lane;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double acc0 = cf[(3 * i + 0) * 32] * lo[0] + cf[(3 * i + 1) * 32] * lo[1];
        double acc1 = cf[(9 + 3 * i + 0) * 32] * hi[0] + cf[(9 + 3 * i + 1) * 32] * hi[1];
        acc0 += cf[(3 * i + 2) * 32] * lo[2];
        acc1 += cf[(9 + 3 * i + 2) * 32] * hi[2];
        r[i] = r[i] - acc0 - acc1;
      }
    }
    double xt[3];
    {
      const double b0 = sm_binv[0 * 32 + lane], b1 = sm_binv[1 * 32 + lane], b2 = sm_binv[2 * 32 + lane];
      const double b4 = sm_binv[3 * 32 + lane], b5 = sm_binv[4 * 32 + lane], b8 = sm_binv[5 * 32 + lane];
      xt[0] = b0 * r[0] + b1 * r[1] + b2 * r[2];
      xt[1] = b1 * r[0] + b4 * r[1] + b5 * r[2];
      xt[2] = b2 * r[0] + b5 * r[1] + b8 * r[2];
    }
    // recover u~_k = h - W^-1 B' R_{k+1} (A x~_k - x~_{k+1})
    double axt[3], v[3];
    A_mul(md, xt, axt);
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double xn = __shfl_down_sync(FULL, xt[i], 1); v[i] = s.rdn[i] * (axt[i] - xn); }
    Bt_mul(md, v, t2);
    double ut[2] = {hh[0] - (s.wi[0] * t2[0] + s.wi[1] * t2[1]), hh[1] - (s.wi[1] * t2[0] + s.wi[2] * t2[1])};
    // z~ = A w~ : dynamics rows need the predecessor's prediction
    double pred[3];
    B_mul(md, ut, pred);
    double ztd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pred[i] += axt[i];
      const double pp = __shfl_up_sync(FULL, pred[i], 1);
      ztd[i] = (hasp ? pp : 0.0) - xt[i];
    }
    const double ztg[2] = {s.gm[0] * xt[0] + s.gm[1] * xt[1] + s.gm[2] * xt[2], s.gm[3] * xt[0] + s.gm[4] * xt[1] + s.gm[5] * xt[2]};
    const double al = p.alpha, oma = 1.0 - p.alpha;
#pragma unroll
    for (int j = 0; j < 3; ++j) s.x[j] = al * xt[j] + oma * s.x[j];
#pragma unroll
    for (int j = 0; j < 2; ++j) s.u[j] = al * ut[j] + oma * s.u[j];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double zr = al * ztd[i] + oma * s.zd[i];
      const double zn = clampd(zr + s.id[i] * s.yd[i], s.bd[i], s.bd[i]);
      s.yd[i] += s.rd[i] * (zr - zn);
      s.zd[i] = zn;
    }
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const double zr = al * ztg[r2] + oma * s.zg[r2];
      const double zn = clampd(zr + s.ig[r2] * s.yg[r2], s.gl[r2], s.gu[r2]);
      s.yg[r2] += s.rg[r2] * (zr - zn);
      s.zg[r2] = zn;
      const double zrb = al * ut[r2] + oma * s.zb[r2];
      const double znb = clampd(zrb + s.ib[r2] * s.yb[r2], s.bl[r2], s.bu[r2]);
      s.yb[r2] += s.rb[r2] * (zrb - znb);
      s.zb[r2] = znb;
    }
  };

  // ---------------- residuals & norms (OSQP update_info + the norms of compute_rho_estimate) -----------------
  double pri_res = 0, dua_res = 0, obj = 0;
  double n_z = 0, n_Ax = 0, n_Aty = 0, n_Px = 0;              // unscaled (termination)
  double s_pri = 0, s_dua = 0, s_z = 0, s_Ax = 0, s_Aty = 0, s_Px = 0;  // scaled (rho estimate)
  auto update_info = [&]() {
    // A x
    double ax[3], pred[3], t3[3];
    A_mul(md, s.x, ax);
    B_mul(md, s.u, pred);
    double Axd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pred[i] += ax[i];
      const double pp = __shfl_up_sync(FULL, pred[i], 1);
      Axd[i] = (hasp ? pp : 0.0) - s.x[i];
    }
    const double Axg[2] = {s.gm[0] * s.x[0] + s.gm[1] * s.x[1] + s.gm[2] * s.x[2], s.gm[3] * s.x[0] + s.gm[4] * s.x[1] + s.gm[5] * s.x[2]};
    // A' y, P x
    double ydn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double t = __shfl_down_sync(FULL, s.yd[i], 1); ydn[i] = actu ? t : 0.0; }
    At_mul(md, ydn, t3);
    double Atx[3], Atu[2], Pxx[3], Pxu[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) { Atx[j] = -s.yd[j] + t3[j] + s.gm[j] * s.yg[0] + s.gm[3 + j] * s.yg[1]; Pxx[j] = p.Q[j] * s.x[j]; }
    double t2[2];
    Bt_mul(md, ydn, t2);
#pragma unroll
    for (int j = 0; j < 2; ++j) { Atu[j] = t2[j] + s.yb[j]; Pxu[j] = p.R[j] * s.u[j]; }
    const double dxv[3] = {sm_de[0 * 32 + lane], sm_de[1 * 32 + lane], sm_de[2 * 32 + lane]};
    const double duv[2] = {sm_de[3 * 32 + lane], sm_de[4 * 32 + lane]};
    const double edv[3] = {sm_de[5 * 32 + lane], sm_de[6 * 32 + lane], sm_de[7 * 32 + lane]};
    const double egv[2] = {sm_de[8 * 32 + lane], sm_de[9 * 32 + lane]};
    const double ebv[2] = {sm_de[10 * 32 + lane], sm_de[11 * 32 + lane]};
    double m_pri = 0, m_z = 0, m_Ax = 0, m_dua = 0, m_Aty = 0, m_Px = 0;
    double q_pri = 0, q_z = 0, q_Ax = 0, q_dua = 0, q_Aty = 0, q_Px = 0, o = 0;
    if (act) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double rr = fabs(Axd[i] - s.zd[i]), zz = fabs(s.zd[i]), aa = fabs(Axd[i]);
        m_pri = fmax(m_pri, rr); m_z = fmax(m_z, zz); m_Ax = fmax(m_Ax, aa);
        q_pri = fmax(q_pri, edv[i] * rr); q_z = fmax(q_z, edv[i] * zz); q_Ax = fmax(q_Ax, edv[i] * aa);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double rr = fabs(Axg[r] - s.zg[r]), zz = fabs(s.zg[r]), aa = fabs(Axg[r]);
        m_pri = fmax(m_pri, rr); m_z = fmax(m_z, zz); m_Ax = fmax(m_Ax, aa);
        q_pri = fmax(q_pri, egv[r] * rr); q_z = fmax(q_z, egv[r] * zz); q_Ax = fmax(q_Ax, egv[r] * aa);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double dr = fabs(Pxx[j] + s.qx[j] + Atx[j]), ay = fabs(Atx[j]), px = fabs(Pxx[j]);
        m_dua = fmax(m_dua, dr); m_Aty = fmax(m_Aty, ay); m_Px = fmax(m_Px, px);
        q_dua = fmax(q_dua, dxv[j] * dr); q_Aty = fmax(q_Aty, dxv[j] * ay); q_Px = fmax(q_Px, dxv[j] * px);
        o += 0.5 * s.x[j] * Pxx[j] + s.qx[j] * s.x[j];
      }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double rr = fabs(s.u[j] - s.zb[j]), zz = fabs(s.zb[j]), aa = fabs(s.u[j]);
        m_pri = fmax(m_pri, rr); m_z = fmax(m_z, zz); m_Ax = fmax(m_Ax, aa);
        q_pri = fmax(q_pri, ebv[j] * rr); q_z = fmax(q_z, ebv[j] * zz); q_Ax = fmax(q_Ax, ebv[j] * aa);
        const double dr = fabs(Pxu[j] + qu[j] + Atu[j]), ay = fabs(Atu[j]), px = fabs(Pxu[j]);
        m_dua = fmax(m_dua, dr); m_Aty = fmax(m_Aty, ay); m_Px = fmax(m_Px, px);
        q_dua = fmax(q_dua, duv[j] * dr); q_Aty = fmax(q_Aty, duv[j] * ay); q_Px = fmax(q_Px, duv[j] * px);
        o += 0.5 * s.u[j] * Pxu[j] + qu[j] * s.u[j];
      }
    }
    pri_res = wmax(m_pri); n_z = wmax(m_z); n_Ax = wmax(m_Ax);
    dua_res = wmax(m_dua); n_Aty = wmax(m_Aty); n_Px = wmax(m_Px);
    s_pri = wmax(q_pri); s_z = wmax(q_z); s_Ax = wmax(q_Ax);
    s_dua = c * wmax(q_dua); s_Aty = c * wmax(q_Aty); s_Px = c * wmax(q_Px);
    obj = wsum(o);
  };

  // ---------------- infeasibility certificates (OSQP is_primal_infeasible / is_dual_infeasible) ----------------
  // delta = last iteration's change; the values before that iteration were parked in shared memory.
  auto primal_infeasible = [&](double eps) -> bool {
    double dyd[3], dyg[2], dyb[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) dyd[i] = s.yd[i] - sm_prev[(5 + i) * 32 + lane];
#pragma unroll
    for (int r = 0; r < 2; ++r) { dyg[r] = s.yg[r] - sm_prev[(8 + r) * 32 + lane]; dyb[r] = s.yb[r] - sm_prev[(10 + r) * 32 + lane]; }
    const double edv[3] = {sm_de[5 * 32 + lane], sm_de[6 * 32 + lane], sm_de[7 * 32 + lane]};
    const double egv[2] = {sm_de[8 * 32 + lane], sm_de[9 * 32 + lane]};
    const double ebv[2] = {sm_de[10 * 32 + lane], sm_de[11 * 32 + lane]};
    // projection on the polar of the recession cone (tests on the SCALED bounds)
    auto proj = [](double dy, double lb, double ub) {
      if (ub > INF_THRESH) { return (lb < -INF_THRESH) ? 0.0 : fmin(dy, 0.0); }
      if (lb < -INF_THRESH) return fmax(dy, 0.0);
      return dy;
    };
#pragma unroll
    for (int i = 0; i < 3; ++i) dyd[i] = proj(dyd[i], edv[i] * s.bd[i], edv[i] * s.bd[i]);
#pragma unroll
    for (int r = 0; r < 2; ++r) { dyg[r] = proj(dyg[r], egv[r] * s.gl[r], egv[r] * s.gu[r]); dyb[r] = proj(dyb[r], ebv[r] * s.bl[r], ebv[r] * s.bu[r]); }
    double mx = 0.0, lhs = 0.0;
    if (act) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { mx = fmax(mx, fabs(dyd[i])); lhs += s.bd[i] * fmax(dyd[i], 0.0) + s.bd[i] * fmin(dyd[i], 0.0); }
#pragma unroll
      for (int r = 0; r < 2; ++r) { mx = fmax(mx, fabs(dyg[r])); lhs += s.gu[r] * fmax(dyg[r], 0.0) + s.gl[r] * fmin(dyg[r], 0.0); }
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) dyd[i] = 0.0;
#pragma unroll
      for (int r = 0; r < 2; ++r) dyg[r] = 0.0;
    }
    if (actu) {
#pragma unroll
      for (int r = 0; r < 2; ++r) { mx = fmax(mx, fabs(dyb[r])); lhs += s.bu[r] * fmax(dyb[r], 0.0) + s.bl[r] * fmin(dyb[r], 0.0); }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) dyb[r] = 0.0;
    }
    const double ndy = wmax(mx);        // unscaled ||dy||; OSQP's scaled-back norm is c * ndy
    if (!(c * ndy > eps)) return false;
    lhs = wsum(lhs);
    if (!(lhs < -eps * ndy)) return false;
    double dn[3], t3[3], t2[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double t = __shfl_down_sync(FULL, dyd[i], 1); dn[i] = actu ? t : 0.0; }
    At_mul(md, dn, t3);
    Bt_mul(md, dn, t2);
    double m2 = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) m2 = fmax(m2, fabs(-dyd[j] + t3[j] + s.gm[j] * dyg[0] + s.gm[3 + j] * dyg[1]));
#pragma unroll
    for (int j = 0; j < 2; ++j) m2 = fmax(m2, fabs(t2[j] + dyb[j]));
    return wmax(m2) < eps * ndy;
  };
  auto dual_infeasible = [&](double eps) -> bool {
    double ddx[3], ddu[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) ddx[j] = s.x[j] - sm_prev[j * 32 + lane];
#pragma unroll
    for (int j = 0; j < 2; ++j) ddu[j] = s.u[j] - sm_prev[(3 + j) * 32 + lane];
    double mx = 0.0, qd = 0.0, mp = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) { mx = fmax(mx, fabs(ddx[j])); qd += s.qx[j] * ddx[j]; mp = fmax(mp, fabs(p.Q[j] * ddx[j])); }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) { mx = fmax(mx, fabs(ddu[j])); qd += qu[j] * ddu[j]; mp = fmax(mp, fabs(p.R[j] * ddu[j])); }
    }
    const double ndx = wmax(mx);
    if (!(ndx > eps)) return false;
    if (!(wsum(qd) < -eps * ndx)) return false;
    if (!(wmax(mp) < eps * ndx)) return false;
    // A dx against the finite sides of [l, u] (finiteness tested on the scaled bounds)
    const double edv[3] = {sm_de[5 * 32 + lane], sm_de[6 * 32 + lane], sm_de[7 * 32 + lane]};
    const double egv[2] = {sm_de[8 * 32 + lane], sm_de[9 * 32 + lane]};
    const double ebv[2] = {sm_de[10 * 32 + lane], sm_de[11 * 32 + lane]};
    double ax[3], pred[3];
    A_mul(md, ddx, ax);
    B_mul(md, ddu, pred);
    bool bad = false;
    const double th = eps * ndx;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pred[i] += ax[i];
      const double pp = __shfl_up_sync(FULL, pred[i], 1);
      const double a = (hasp ? pp : 0.0) - ddx[i];
      const double sb_ = edv[i] * s.bd[i];
      if (act && ((sb_ < INF_THRESH && a > th) || (sb_ > -INF_THRESH && a < -th))) bad = true;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double a = s.gm[3 * r] * ddx[0] + s.gm[3 * r + 1] * ddx[1] + s.gm[3 * r + 2] * ddx[2];
      if (act && ((egv[r] * s.gu[r] < INF_THRESH && a > th) || (egv[r] * s.gl[r] > -INF_THRESH && a < -th))) bad = true;
      const double b = ddu[r];
      if (actu && ((ebv[r] * s.bu[r] < INF_THRESH && b > th) || (ebv[r] * s.bl[r] > -INF_THRESH && b < -th))) bad = true;
    }
    return !__any_sync(FULL, bad);
  };

  int status = ST_UNSOLVED;
  // OSQP check_termination(work, approximate)
  auto check_termination = [&](bool approximate) -> bool {
    double eps_abs = p.eps_abs, eps_rel = p.eps_rel, epi = p.eps_prim_inf, edi = p.eps_dual_inf;
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = ST_NON_CVX; return true; }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, pinf = false, dinf = false;
    const double eps_prim = eps_abs + eps_rel * fmax(n_z, n_Ax);
    if (pri_res < eps_prim) prim_ok = true;
    else pinf = primal_infeasible(epi);
    const double eps_dual = eps_abs + eps_rel * fmax(nq, fmax(n_Aty, n_Px));
    if (dua_res < eps_dual) dual_ok = true;
    else dinf = dual_infeasible(edi);
    if (prim_ok && dual_ok) { status = approximate ? ST_SOLVED_INACC : ST_SOLVED; return true; }
    if (pinf) { status = approximate ? ST_PINF_INACC : ST_PINF; obj = OSQP_INFTY; return true; }
    if (dinf) { status = approximate ? ST_DINF_INACC : ST_DINF; obj = -OSQP_INFTY; return true; }
    return false;
  };

  // ---------------- main loop (OSQP osqp_solve) -----------------------------------------------------------
  factor();
  const int ct = p.check_termination;
  const int ari = p.adaptive_rho ? p.adaptive_rho_interval : 0;
  int iter = 0, n_rho_updates = 0;
  bool can_check = false, done = false;
  for (iter = 1; iter <= p.max_iter; ++iter) {
    const bool chk = ct && (iter % ct == 0);
    const bool adp = ari && (iter % ari == 0);
    if (chk || adp || iter == p.max_iter) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { sm_prev[j * 32 + lane] = s.x[j]; sm_prev[(5 + j) * 32 + lane] = s.yd[j]; }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sm_prev[(3 + j) * 32 + lane] = s.u[j];
        sm_prev[(8 + j) * 32 + lane] = s.yg[j];
        sm_prev[(10 + j) * 32 + lane] = s.yb[j];
      }
    }
    iterate();
#ifdef F110_DEBUG_NAN
    {
      int code = 0;
      for (int j = 0; j < 3; ++j) { if (!isfinite(s.x[j])) code |= 1; if (!isfinite(s.zd[j])) code |= 4; if (!isfinite(s.yd[j])) code |= 32; }
      for (int j = 0; j < 2; ++j) { if (!isfinite(s.u[j])) code |= 2; if (!isfinite(s.zg[j])) code |= 8; if (!isfinite(s.zb[j])) code |= 16; if (!isfinite(s.yg[j])) code |= 64; if (!isfinite(s.yb[j])) code |= 128; }
      unsigned bal = __ballot_sync(FULL, code != 0);
      if (bal) {
        int first = __ffs(bal) - 1;
        int c0 = __shfl_sync(FULL, code, first);
        if (lane == 0 && p.info) { double* io = p.info + 4 * (size_t)qp; io[0] = iter; io[1] = first; io[2] = c0; io[3] = (double)bal; }
        if (p.status && lane == 0) p.status[qp] = -99;
        return;
      }
    }
#endif
    can_check = chk;
    if (chk) {
      update_info();
      if (check_termination(false)) { done = true; break; }
    }
    if (adp) {
      if (!chk) update_info();
      // compute_rho_estimate on the scaled residuals, adapt_rho
      const double pr = s_pri / (fmax(s_z, s_Ax) + 1e-10);
      const double dr = s_dua / (fmax(snq, fmax(s_Aty, s_Px)) + 1e-10);
      double rho_new = rho_bar * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > rho_bar * p.adaptive_rho_tolerance || rho_new < rho_bar / p.adaptive_rho_tolerance) {
        rho_bar = rho_new;
        ++n_rho_updates;
        factor();
      }
    }
  }
  if (!done) iter = p.max_iter;
  if (!can_check) {
    update_info();
    check_termination(false);
  }
  if (status == ST_UNSOLVED) {
    if (!check_termination(true)) status = ST_MAX_ITER;
  }

  // ---------------- store (OSQP store_solution: NaN + cold start when infeasible) ---------------------------
  const bool has_sol = !(status == ST_PINF || status == ST_PINF_INACC || status == ST_DINF || status == ST_DINF_INACC || status == ST_NON_CVX);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (p.x_out) {
    double* xo = p.x_out + (size_t)qp * nvar;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) xo[3 * k + j] = has_sol ? s.x[j] : qnan;
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) xo[3 * (N + 1) + 2 * k + j] = has_sol ? s.u[j] : qnan;
    }
  }
  if (p.y_out) {
    double* yo = p.y_out + (size_t)qp * mcon;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) yo[3 * k + j] = has_sol ? s.yd[j] : qnan;
#pragma unroll
      for (int r = 0; r < 2; ++r) yo[3 * (N + 1) + 2 * k + r] = has_sol ? s.yg[r] : qnan;
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) yo[5 * (N + 1) + 2 * k + j] = has_sol ? s.yb[j] : qnan;
    }
  }
  if (lane == 0) {
    if (p.u0_out) { p.u0_out[2 * (size_t)qp] = has_sol ? s.u[0] : qnan; p.u0_out[2 * (size_t)qp + 1] = has_sol ? s.u[1] : qnan; }
    if (p.status) p.status[qp] = status;
    if (p.iters) p.iters[qp] = iter;
    if (p.rho_updates) p.rho_updates[qp] = n_rho_updates;
    if (p.info) {
      double* io = p.info + 4 * (size_t)qp;
      io[0] = obj; io[1] = pri_res; io[2] = dua_res; io[3] = rho_bar;
    }
  }
  if (slot) {
    // scaled iterates for the next warm start; zeros (cold start) when there is no solution
    const double dxv[3] = {sm_de[0 * 32 + lane], sm_de[1 * 32 + lane], sm_de[2 * 32 + lane]};
    const double duv[2] = {sm_de[3 * 32 + lane], sm_de[4 * 32 + lane]};
    const double edv[3] = {sm_de[5 * 32 + lane], sm_de[6 * 32 + lane], sm_de[7 * 32 + lane]};
    const double egv[2] = {sm_de[8 * 32 + lane], sm_de[9 * 32 + lane]};
    const double ebv[2] = {sm_de[10 * 32 + lane], sm_de[11 * 32 + lane]};
    double* sx_ = slot;
    double* sz_ = slot + nvar;
    double* sy_ = slot + nvar + mcon;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        sx_[3 * k + j] = has_sol ? s.x[j] / dxv[j] : 0.0;
        sz_[3 * k + j] = has_sol ? edv[j] * s.zd[j] : 0.0;
        sy_[3 * k + j] = has_sol ? c * s.yd[j] / edv[j] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        sz_[3 * (N + 1) + 2 * k + r] = has_sol ? egv[r] * s.zg[r] : 0.0;
        sy_[3 * (N + 1) + 2 * k + r] = has_sol ? c * s.yg[r] / egv[r] : 0.0;
      }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sx_[3 * (N + 1) + 2 * k + j] = has_sol ? s.u[j] / duv[j] : 0.0;
        sz_[5 * (N + 1) + 2 * k + j] = has_sol ? ebv[j] * s.zb[j] : 0.0;
        sy_[5 * (N + 1) + 2 * k + j] = has_sol ? c * s.yb[j] / ebv[j] : 0.0;
      }
    }
    if (lane == 0) { slot[nvar + 2 * mcon] = rho_bar; slot[nvar + 2 * mcon + 1] = 1.0; }
  }
}

cudaError_t launch_admm(const KParams& p, cudaStream_t stream, int* launches) {
  constexpr int WARPS = 1;
  const size_t smem = (size_t)WARPS * SM_PER_WARP * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(admm_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = (p.B + WARPS - 1) / WARPS;
  admm_kernel<WARPS><<<grid, 32 * WARPS, smem, stream>>>(p);
  if (launches) *launches = 1;
  return cudaGetLastError();
}

}  // namespace f110
