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
//    on chip (negated, in 16-byte pairs); each iteration only applies them to the right-hand side as FMA
//    chains.  The top level is one-sided (partner = stage k XOR h) and stores a single 3x3 block.
//  * Where the multipliers live: in TENSOR MEMORY for horizons 16..127 of the base row set and 16..63 with
//    steering-rate rows (TM = true: tcgen05.st in the factor step, tcgen05.ld one PCR level ahead of its use in
//    the iteration; each lane reads back exactly what it wrote, so the warp's 32-row strip is private storage
//    with its own datapath), in shared memory otherwise.  The round-1 kernel was bound by the shared-memory /
//    shuffle data pipe; see admm_device.cuh and DESIGN.md section 3.1.
//  * A QP on two or four warps (horizons 32..127, base row set) is solved in partitions: each warp reduces its
//    own 32 stages with shuffles, the blocks that couple neighbouring warps enter as spike corrections from one
//    boundary exchange per partition level and iteration (PART below) — one or two barriers instead of six or
//    seven for the reduction.
//  * The parameter record is staged into shared memory by one bulk asynchronous copy (TMA, cp.async.bulk +
//    mbarrier) when it is 16-byte aligned; the Ruiz passes do not wait for anything but that one transfer.
//  * x/z/y update, projection onto [l,u], residual norms and the termination test are fused in the same
//    kernel; all reductions are warp shuffles / REDUX.  Nothing but the parameter record is read from HBM and
//    nothing but the solution is written (a per-QP scratch line — shared memory for the tensor-memory
//    kernels, L2 otherwise — holds the scaling vectors and the previous iterate, touched once per
//    termination check).
//
//  * RATE = true adds N steering-rate rows  delta_k - delta_{k-1} in [-D, D]  (row 0: delta_0 - steer_prev).  They couple
//    consecutive inputs, so only the speed v_k is eliminated per stage (a scalar pivot) and the steering angle joins
//    the reduced unknown: s_k = (x_k, delta_k), 4x4 blocks, same cyclic reduction and pair layout.
//  * SBOX = true adds 3(N+1) identity rows on x_0..x_N (the state box the reference stores but never stacks,
//    constraints.cpp:14-17, 108-114): rho on the diagonal of T_kk, three more (z, y) pairs per stage.
//
// Lane k owns stage k: x_k(3), u_k(2) (k<N), dynamics rows k (3), gap rows k (2), input-box rows k (2), rate row k (1).
// Lanes above N hold all-zero state and never feed an active lane (every cross-lane read is masked or
// multiplied by a zero multiplier), so no branch in the iteration depends on the lane.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "admm_kernel.cuh"
#include "admm_device.cuh"

namespace f110 {

// One unit of work = one QP on WPQ warps, one horizon stage per thread (stage k = tid) — or QPW short-horizon QPs side by side in one warp.
// NLEV = number of PCR levels = floor(log2(N)) + 1;  LASTFULL = (N + 1 == lanes of the QP): the last thread is an active
// stage, so the "successor" reads of the last stage wrap onto itself and need a mask.
// QPW = QPs per warp (WPQ == 1 only): short horizons (N + 1 <= 16 / 8) put 2 / 4 QPs side by side in one warp, each on its own
// G = 32 / QPW lane segment.  The QPs of a warp iterate in lock step; one that terminates stores its result at once and idles
// (keeps iterating, results discarded) until its neighbours are done too.
// TM = the tensor-memory variant (one-warp QPs, horizons 16..31): the PCR multipliers live in the warp's TMEM strip (`tmb`) instead
// of shared memory, the per-QP scratch line lives in shared memory instead of global memory, and the caller is a persistent
// four-warp CTA whose warps fetch QPs from a work counter (admm_kernel_tm below).  `unit` = blockIdx.x for the one-CTA-per-unit
// kernels, the fetched index for the persistent one; `smem_all` = this unit's shared-memory region; `tid` = thread within the unit.
template <int NLEV, int WPQ, bool LASTFULL, bool RATE, int QPW, bool TM, bool SBOX = false>
__device__ __forceinline__ void solve_unit(const KParams& p, const int unit, double* const smem_all, [[maybe_unused]] const uint32_t tmb,
                                           const int tid, [[maybe_unused]] const uint32_t rec_parity, [[maybe_unused]] const int bar_id = 0) {
  static_assert(QPW == 1 || WPQ == 1, "several QPs per warp only for one-warp horizons");
  static_assert(!TM || QPW == 1, "tensor-memory variant: one QP per unit (measured for 2 / 4 QPs per warp, horizons 1..15: 4-9 % slower than their shared-memory kernels)");
  static_assert(!(TM && RATE) || WPQ <= 2, "steering-rate rows in tensor memory: one or two warps per QP");
  static_assert(!SBOX || (WPQ == 1 && QPW == 1 && !RATE && !TM), "state-box rows: one-warp shared-memory kernel, without steering-rate rows");
  constexpr int T = 32 * WPQ;              // threads per unit = columns of the shared-memory and scratch layouts
  constexpr int G = QPW == 1 ? T : 32 / QPW;   // lanes per QP
  const int k = QPW == 1 ? tid : (tid & (G - 1));                      // stage
  const int qp_raw = QPW == 1 ? unit : (unit * QPW + tid / G);
  // a warp's trailing groups may have no QP: they shadow the batch's last record, use a dummy scratch line and never store
  const bool live = QPW == 1 || qp_raw < p.B;
  const int qp = live ? qp_raw : p.B - 1;
  [[maybe_unused]] bool done = !live;
  // PCR multipliers as double2 pairs, pair-major, stage fastest: 9 pairs (-alpha, -gamma) for each of the first
  // NLEV-1 levels, 5 pairs for the one-sided top level, 3 pairs for the final block inverse
  // Multi-warp QPs keep the multipliers of their top PCR levels in a per-QP global line, read back through L1 / L2 by the lane
  // that wrote them; the shared memory saved buys residency.  Four warps (N 64..127): the two top levels (9 + 5 pairs) -> two QPs
  // per SM instead of one (N = 100: 2.75 -> 2.24 ms per 4096 QPs).  Two warps (N 32..63): only the one-sided top level (5 pairs)
  // -> four QPs per SM instead of three (2-4 % faster; moving both levels loses 5 %).  GL = number of levels kept there.
  constexpr int GL = (RATE || TM) ? 0 : (WPQ == 4 ? 2 : (WPQ == 2 ? ADMM_W2_GLOBAL_LEVELS : 0));
  constexpr bool TOPG = GL > 0;
  constexpr int GLP = GL == 2 ? 14 : 5;            // pairs per stage in the global line
  constexpr int GTOP = GL == 2 ? 9 : 0;            // where the top level starts in it
  // PART: the partitioned solve of a two-warp QP (tensor-memory variant, base row set).  Each warp runs the cyclic reduction on its
  // own 32 stages with shuffles only (NLEVP = 5 levels, no barrier); the coupling block between stage 31 and stage 32 enters as a
  // "spike": with V = T_own^-1 [coupling columns] (three PCR applications per factor step), the true solution is
  //     x_k = y_k + WT_k y_other + WO_k y_own,    y = T_own^-1 r,   y_own / y_other = y at this / the other warp's boundary stage,
  // where the per-stage 3x3 blocks WT = -V_k S, WO = -WT V_other, S = (I - V_other V_own)^-1 are formed once per factor step.
  // One exchange of six doubles (one barrier) per iteration replaces the six barrier-separated exchanges of a 64-lane reduction.
  // A four-warp QP nests the construction: level 1 joins warps (0, 1) and (2, 3) across stages 31 | 32 and 95 | 96, level 2 joins
  // the two pairs across stages 63 | 64, its spikes computed with the level-1 solve — two exchanges per iteration instead of seven.
  constexpr bool PART = TM && (WPQ == 2 || WPQ == 4) && !RATE;
  constexpr int PLEV = WPQ == 4 ? 2 : 1;   // levels of the partition
  constexpr int NLEVP = PART ? 5 : NLEV;   // levels of the reduction that is actually run
  constexpr int SM_PAIRS = PART ? NLEVP * 9 - 1 + 9 * PLEV : NLEV * 9 - 1 - (GL >= 1 ? 5 : 0) - (GL == 2 ? 9 : 0);
  constexpr int FINAL_PAIR = PART ? NLEVP * 9 - 4 : SM_PAIRS - 3;
  constexpr int W_PAIR = NLEVP * 9 - 1;   // (PART) the nine pairs of WT, WO of level 1; level 2's follow
  static_assert(!TM || RATE || 4 * SM_PAIRS <= TM_COLS, "multipliers exceed the warp's tensor-memory strip");
  // Steering-rate rows + tensor memory: the 16 pairs of each of the first TML two-sided levels live in the strip (4 x 16 pairs = all
  // 256 columns; NLEV = 5 has four such levels, NLEV = 6 five); what is left — a fifth two-sided level, the one-sided top level
  // (8 pairs) and the final inverse (5 pairs) — stays in shared memory, re-based to index 0.
  // The scratch line stays in global memory for this variant: its spill traffic needs the L1 the shared memory would take.
  constexpr int TML = (TM && RATE) ? ((NLEV - 1) * 64 <= TM_COLS ? NLEV - 1 : TM_COLS / 64) : 0;   // two-sided levels in tensor memory
  constexpr bool SCR_SM = TM && !RATE;                            // scratch line in shared memory
  constexpr int RT_SM0 = TML * 16;                                 // first steering-rate pair that is NOT in tensor memory
  [[maybe_unused]] double2* sm_pair = reinterpret_cast<double2*>(smem_all) + tid;
  [[maybe_unused]] double2* gl_pair = TOPG ? reinterpret_cast<double2*>(p.mult_global) + (size_t)unit * (GLP * T) + tid : nullptr;
  // steering-rate variant (4x4 blocks), same pair-major layout: 16 pairs per two-sided level, 8 for the one-sided top level,
  // 5 for the symmetric final inverse
  constexpr int SM_DOUBLES = TM ? (RATE ? (NLEV - 1 - TML) * 32 + 26 : 0) : (RATE ? (NLEV - 1) * 32 + 26 : 2 * SM_PAIRS);
  // (tensor-memory variant: the scratch line, when it lives in shared memory, comes first; the exchange buffers of a multi-warp QP follow)
  Comm<WPQ, QPW == 1 ? 32 : G> cm(smem_all + ((TM && !RATE) ? SCR_ROWS_ALLOC * T : 0) + SM_DOUBLES * T, tid, bar_id);
  // scratch line: global memory (L2), or — tensor-memory variant — the shared memory the multipliers no longer occupy
  double* scr;
  if constexpr (SCR_SM) scr = smem_all + k;
  else scr = (live ? p.scratch + (size_t)qp * (SCR_ROWS_ALLOC * T) : p.scratch_dummy + (size_t)(tid / G) * (SCR_ROWS_ALLOC * T)) + k;

  const int N = p.N;
  const bool act = k <= N;         // lane owns a stage
  const bool actu = k < N;         // stage has an input (and box rows, and a successor)
  const bool hasp = act && k > 0;  // stage has a predecessor
  const int nvar = 5 * N + 3;
  const int mcon = 7 * N + 5 + (RATE ? N : 0) + (SBOX ? 3 * (N + 1) : 0);
  [[maybe_unused]] const int row_s0 = 7 * N + 5 + (RATE ? N : 0);    // first state-box row
  const int row_r0 = 7 * N + 5;    // first steering-rate row

  // ---------------- load the parameter record, linearise, stack -----------------------------------------
  const double* rec = p.recs + (size_t)qp * p.stride;
  if (p.rec_bulk_bytes) {
    // one TMA bulk copy of the whole record (host checked 16-byte alignment of base and stride), then every read below
    // is a shared-memory read
    double* rec_sm = smem_all + p.rec_smem_offset;
    if constexpr (TM && (WPQ == 1 || RATE)) {
      // persistent warp / QP slot: the mbarrier was initialised once by the caller, its phase alternates from QP to QP
      uint64_t* bar = reinterpret_cast<uint64_t*>(rec_sm + p.rec_bulk_bytes / 8);
      if (k == 0) bulk_copy_g2s(smem_u32(rec_sm), rec, (uint32_t)p.rec_bulk_bytes, smem_u32(bar));
      mbar_wait(smem_u32(bar), rec_parity);
      rec = rec_sm;
    } else if constexpr (QPW == 1) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(rec_sm + p.rec_bulk_bytes / 8);
      if (k == 0) {
        mbar_init(smem_u32(bar), 1);
        bulk_copy_g2s(smem_u32(rec_sm), rec, (uint32_t)p.rec_bulk_bytes, smem_u32(bar));
      }
      cm.sync();
      mbar_wait(smem_u32(bar), 0);
      rec = rec_sm;
    } else {
      // the warp's records are consecutive rows of the batch: one bulk copy covers them all
      const int first = unit * QPW;
      const int nq = (p.B - first < QPW) ? p.B - first : QPW;
      uint64_t* bar = reinterpret_cast<uint64_t*>(rec_sm + (QPW - 1) * p.stride + p.rec_bulk_bytes / 8);
      if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        bulk_copy_g2s(smem_u32(rec_sm), p.recs + (size_t)first * p.stride, (uint32_t)((nq - 1) * p.stride * 8 + p.rec_bulk_bytes), smem_u32(bar));
      }
      cm.sync();
      mbar_wait(smem_u32(bar), 0);
      rec = rec_sm + (size_t)(qp - first) * p.stride;
    }
  }
  const double x0[3] = {rec[0], rec[1], rec[2]};
  const double vlin = rec[3], slin = rec[4];
  if (!(vlin == vlin)) {
    // NaN linearisation speed marks an empty slot (the planning stage found no valid mini-path for this scene,
    // project.cpp:115-119): nothing to solve, status stays UNSOLVED
    if (k == 0 && !done) {
      const double qn = __longlong_as_double(0x7ff8000000000000LL);
      if (p.u0_out) { p.u0_out[2 * (size_t)qp] = qn; p.u0_out[2 * (size_t)qp + 1] = qn; }
      if (p.status) p.status[qp] = ST_UNSOLVED;
      if (p.iters) p.iters[qp] = 0;
      if (p.rho_updates) p.rho_updates[qp] = 0;
      if (p.packed) { double* po = p.packed + 4 * (size_t)qp; po[0] = qn; po[1] = qn; po[2] = (double)ST_UNSOLVED; po[3] = 0.0; }
    }
    if constexpr (QPW == 1) return;
    done = true;   // the group idles (on NaN data, inside its own lanes) while its neighbours solve
  }
  Model md;
  double Cv[3];
  {
    // Model::Linearize, model.cpp:42-55 (operation order kept)
    const double L = p.wheelbase, dt = p.dt;
    double so, co, ss, cs;
    sincos(x0[2], &so, &co);
    sincos(slin, &ss, &cs);
    const double pw = 1.0 / (cs * cs);   // pow(cos, -2) of model.cpp:51,55 to within 1 ulp
    md.a02 = -1.0 * vlin * so * dt;
    md.a12 = vlin * co * dt;
    md.b00 = co * dt;
    md.b10 = so * dt;
    md.b20 = (ss / cs) * dt / L;         // tan(steer), model.cpp:50, to within 1 ulp
    md.b21 = vlin * pw * dt / L;
    Cv[0] = vlin * x0[2] * so * dt;
    Cv[1] = -1.0 * vlin * x0[2] * co * dt;
    Cv[2] = -1.0 * slin * vlin * pw * dt / L;
  }
  const double* qu = p.qu;  // -R u_des (mpc.cpp:226), precomputed on the host: lives in the constant bank

  StageT<RATE, SBOX> s;
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
    // gap_mode 1: the commented alternative of mpc.cpp:297-298 on every stage (incl. the all-ones stage-0 pair);
    // gap_mode 2: the same on stages k >= 1 only (the stage-0 pair, whose rows are not half-planes, stays loose)
    const bool gap_on = act && (p.gap_mode == 1 || (p.gap_mode == 2 && k > 0));
    s.gl[0] = gap_on ? -l1c : -OSQP_INFTY;  // mpc.cpp:297
    s.gl[1] = gap_on ? -l2c : -OSQP_INFTY;  // mpc.cpp:298; upper bound +INFTY (mpc.cpp:288-290)
    // input box (mpc.cpp:281,290): the same u_min / u_max on every stage, read from the constant bank.  The last stage has no
    // input; its box rows get rho = 0 in the factor step, so whatever its z does never reaches a right-hand side.
    if constexpr (RATE) {
      s.rbase = (k == 0) ? slin : 0.0;   // row 0 is measured from the steering applied last cycle
    }
    if constexpr (SBOX) {   // Constraints::SetXLims (constraints.cpp:108-114): x, y within +-d of the current state, orientation free
      s.slo[0] = x0[0] - p.state_lim; s.shi[0] = x0[0] + p.state_lim;
      s.slo[1] = x0[1] - p.state_lim; s.shi[1] = x0[1] + p.state_lim;
      s.slo[2] = -OSQP_INFTY; s.shi[2] = OSQP_INFTY;
    }
  }

  // ---------------- Ruiz equilibration (OSQP scale_data) -------------------------------------------------
  double c = 1.0, cinv = 1.0;
  // Everything only the factor step or the termination checks need (w_i = e_i^2 / c, row classes, ||q||) is parked in the
  // scratch line instead of registers: the iteration loop runs at the 255-register ceiling.
  {
    double dx[3] = {1, 1, 1}, du[2] = {1, 1}, ed[3] = {1, 1, 1}, eg[2] = {1, 1}, eb[2] = {1, 1};
    double er = 1.0;   // steering-rate row (RATE)
    [[maybe_unused]] double es[3] = {1, 1, 1};   // state-box rows (SBOX)
    const double aA02 = fabs(md.a02), aA12 = fabs(md.a12);
    const double aB[6] = {fabs(md.b00), 0.0, fabs(md.b10), 0.0, fabs(md.b20), fabs(md.b21)};
    double ag[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) ag[e] = fabs(s.gm[e]);
    const double inv_nvar = 1.0 / (double)nvar;   // (one division per QP instead of one per pass: within an ulp of sum / nvar)
    // Two passes per trip of the loop: the tail of a pass (the warp-wide sum and max behind the cost scale c) and the head of the next
    // (neighbour exchange, every maximum that does not contain c) are independent, and only inside one basic block can the
    // scheduler interleave them.  c enters each column norm last for the same reason (a maximum does not care about the order).
#pragma unroll(RATE ? 1 : 2)   // (the steering-rate kernels run at the register cap: the longer body costs them 2 %)
    for (int it = 0; it < p.scaling; ++it) {
      double edn[3], dxp[3], dup[2];
      double ern = 0.0;   // scale of the next stage's rate row (RATE)
      {
        const double snd[5] = {dx[0], dx[1], dx[2], du[0], du[1]};
        double rcv[5];
        if constexpr (RATE) {
          const double se[4] = {ed[0], ed[1], ed[2], er};
          double re[4];
          cm.template dn<4>(se, re, 1);
          edn[0] = re[0]; edn[1] = re[1]; edn[2] = re[2];
          ern = (k + 1 < N) ? re[3] : 0.0;
        } else {
          cm.template dn<3>(ed, edn, 1);
        }
        cm.template up<5>(snd, rcv, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) { edn[i] = actu ? edn[i] : 0.0; dxp[i] = hasp ? rcv[i] : 0.0; }
#pragma unroll
        for (int j = 0; j < 2; ++j) dup[j] = hasp ? rcv[3 + j] : 0.0;
      }
      double tx[3], tu[2], td[3], tg[2], tb[2];
      // Column / row infinity norms of the scaled KKT matrix.  Every entry of a column carries the column's own scale as a positive
      // factor, and rounding is monotone, so  max_i fl(a_i d) = fl((max_i a_i) d):  the factor is applied once, after the max.
      // Structural zeros of B (b01 = b11 = 0) are left out of the maxima.
      const double cQ[3] = {c * p.Q[0], c * p.Q[1], c * p.Q[2]}, cR[2] = {c * p.R[0], c * p.R[1]};
      // KKT column of x_k[j]: P, the -1 of dyn row k, column j of A in dyn rows k+1 (A = I + a02/a12 in col 2), gap rows k
      [[maybe_unused]] double ts[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double v = dmax(ed[j], edn[j]);
        v = dmax(v, dmax(eg[0] * ag[j], eg[1] * ag[3 + j]));
        if (j == 2) v = dmax(v, dmax(edn[0] * aA02, edn[1] * aA12));
        if constexpr (SBOX) {   // identity rows on x_k: one more entry in the column of x_k[j], a one-entry row
          v = dmax(v, es[j]);
          ts[j] = es[j] * dx[j];
        }
        v = dmax(v, cQ[j] * dx[j]);
        tx[j] = v * dx[j];
      }
      {  // KKT columns of u_k = (v_k, delta_k): B has b00, b10, b20 in its first column, b21 in its second
        double v = dmax(dmax(edn[0] * aB[0], edn[1] * aB[2]), dmax(edn[2] * aB[4], eb[0]));
        v = dmax(v, cR[0] * du[0]);
        tu[0] = v * du[0];
        v = dmax(edn[2] * aB[5], eb[1]);
        // steering-rate rows: column of delta_k has +1 in rate row k, -1 in rate row k+1
        if constexpr (RATE) v = dmax(v, dmax(er, ern));
        v = dmax(v, cR[1] * du[1]);
        tu[1] = v * du[1];
      }
      double tr = 0.0;
      if constexpr (RATE) tr = er * dmax(du[1], dup[1]);   // rate row k: +1 on delta_k, -1 on delta_{k-1}
      // KKT column (= A row) of dyn row (k, i): -1 on x_k[i], row i of [A B] on stage k-1
      td[0] = dmax(dmax(dx[0], dxp[0]), dmax(aB[0] * dup[0], aA02 * dxp[2]));
      td[1] = dmax(dmax(dx[1], dxp[1]), dmax(aB[2] * dup[0], aA12 * dxp[2]));
      td[2] = dmax(dmax(dx[2], dxp[2]), dmax(aB[4] * dup[0], aB[5] * dup[1]));
#pragma unroll
      for (int i = 0; i < 3; ++i) td[i] *= ed[i];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) v = dmax(v, ag[3 * r + j] * dx[j]);
        tg[r] = eg[r] * v;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) tb[j] = eb[j] * du[j];
      // (stages above N scale themselves freely: every sum, max and neighbour read masks them out, and each pass moves a
      //  factor by at most 1e2, so nothing overflows)
#pragma unroll
      for (int j = 0; j < 3; ++j) dx[j] *= rsqrt_scaling(limit_scaling(tx[j]));
#pragma unroll
      for (int j = 0; j < 2; ++j) du[j] *= rsqrt_scaling(limit_scaling(tu[j]));
#pragma unroll
      for (int i = 0; i < 3; ++i) ed[i] *= rsqrt_scaling(limit_scaling(td[i]));
#pragma unroll
      for (int r = 0; r < 2; ++r) eg[r] *= rsqrt_scaling(limit_scaling(tg[r]));
#pragma unroll
      for (int j = 0; j < 2; ++j) eb[j] *= rsqrt_scaling(limit_scaling(tb[j]));
      if constexpr (RATE) er *= rsqrt_scaling(limit_scaling(tr));
      if constexpr (SBOX) {
#pragma unroll
        for (int j = 0; j < 3; ++j) es[j] *= rsqrt_scaling(limit_scaling(ts[j]));
      }
      // cost normalisation: c_temp = 1 / max(mean column norm of P, ||q||_inf)
      double psum = 0.0, qn = 0.0;   // (selects, not branches: the pass stays one basic block)
      {
        double ps = 0.0, pu, qs = 0.0, qv = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) { ps = fma(cQ[j] * dx[j], dx[j], ps); qs = dmax(qs, fabs(dx[j] * s.qx[j])); }
        pu = ps;
#pragma unroll
        for (int j = 0; j < 2; ++j) { pu = fma(cR[j] * du[j], du[j], pu); qv = dmax(qv, fabs(du[j] * qu[j])); }
        psum = actu ? pu : (act ? ps : 0.0);
        qn = actu ? dmax(qs, qv) : (act ? qs : 0.0);
      }
      cm.template reduce<1, 1>(&qn, &psum);
      const double mean = psum * inv_nvar;
      const double qinf = limit_scaling(c * qn);
      const double ct = limit_scaling(dmax(mean, qinf));
      c *= rcp_pos(ct);
    }
    cinv = 1.0 / c;
    // park D, E in the scratch line: only the rho estimate, infeasibility tests and the state store read them again
#pragma unroll
    for (int j = 0; j < 3; ++j) { scr[(SCR_DX + j) * T] = dx[j]; scr[(SCR_ED + j) * T] = ed[j]; }
#pragma unroll
    for (int j = 0; j < 2; ++j) { scr[(SCR_DU + j) * T] = du[j]; scr[(SCR_EG + j) * T] = eg[j]; scr[(SCR_EB + j) * T] = eb[j]; }
    // row classes (OSQP set_rho_vec, on the SCALED bounds); dynamics rows have l = u -> equality
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double lb = eg[r] * s.gl[r], ub = eg[r] * OSQP_INFTY;
      scr[(SCR_CG + r) * T] = (lb < -INF_THRESH && ub > INF_THRESH) ? -1.0 : ((ub - lb < RHO_TOL) ? 1.0 : 0.0);
      const double lbb = eb[r] * p.u_min[r], ubb = eb[r] * p.u_max[r];
      scr[(SCR_CB + r) * T] = (lbb < -INF_THRESH && ubb > INF_THRESH) ? -1.0 : ((ubb - lbb < RHO_TOL) ? 1.0 : 0.0);
    }
    if constexpr (RATE) {
      const double lb = er * (s.rbase - p.rate_delta), ub = er * (s.rbase + p.rate_delta);
      scr[SCR_ER * T] = er;
      scr[SCR_WR * T] = er * er * cinv;
      scr[SCR_CR * T] = (lb < -INF_THRESH && ub > INF_THRESH) ? -1.0 : ((ub - lb < RHO_TOL) ? 1.0 : 0.0);
    }
    if constexpr (SBOX) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double lb = es[j] * s.slo[j], ub = es[j] * s.shi[j];
        scr[(SCR_ES + j) * T] = es[j];
        scr[(SCR_WS + j) * T] = es[j] * es[j] * cinv;
        scr[(SCR_CS + j) * T] = (lb < -INF_THRESH && ub > INF_THRESH) ? -1.0 : ((ub - lb < RHO_TOL) ? 1.0 : 0.0);
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) scr[(SCR_WD + i) * T] = ed[i] * ed[i] * cinv;
#pragma unroll
    for (int r = 0; r < 2; ++r) { scr[(SCR_WG + r) * T] = eg[r] * eg[r] * cinv; scr[(SCR_WB + r) * T] = eb[r] * eb[r] * cinv; }
#pragma unroll
    for (int j = 0; j < 3; ++j) s.sx[j] = p.sigma * cinv / (dx[j] * dx[j]);
#pragma unroll
    for (int j = 0; j < 2; ++j) s.su[j] = p.sigma * cinv / (du[j] * du[j]);
    double a = 0.0, b = 0.0;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { a = dmax(a, fabs(s.qx[j])); b = dmax(b, fabs(dx[j] * s.qx[j])); }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) { a = dmax(a, fabs(qu[j])); b = dmax(b, fabs(du[j] * qu[j])); }
    }
    {
      double ab[2] = {a, b};
      cm.template reduce<2, 0>(ab, nullptr);
      scr[SCR_NQ * T] = ab[0];
      scr[SCR_SNQ * T] = c * ab[1];
    }
  }

  // ---------------- iterates: cold start or the slot's stored (scaled) iterates -----------------------------
#pragma unroll
  for (int j = 0; j < 3; ++j) { s.x[j] = 0; s.yd[j] = 0; }
  // z of the dynamics rows is their right-hand side after the first projection (l = u); only a solve's FIRST iteration sees anything
  // else (0 on a cold start, the slot's value on a warm start), and it reads that from the scratch line
  double zd0[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < 2; ++j) { s.u[j] = 0; s.zg[j] = 0; s.zb[j] = 0; s.yg[j] = 0; s.yb[j] = 0; }
  if constexpr (RATE) { s.zr = 0; s.yr = 0; }
  if constexpr (SBOX) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { s.zs[j] = 0; s.ys[j] = 0; }
  }
  double rho_bar = dmin(dmax(p.rho0, RHO_MIN), RHO_MAX);
  double* slot = (p.state && live) ? p.state + (size_t)qp * state_doubles(N, RATE ? 1 : 0, SBOX ? 1 : 0) : nullptr;
  if (slot && p.warm_start && slot[nvar + 2 * mcon + 1] != 0.0) {
    // OSQP keeps x, z, y in SCALED coordinates across re-scalings (osqp_update_A rescales the data only):
    // x = D xbar, z = zbar / E, y = E ybar / c with the NEW D, E, c.
    rho_bar = slot[nvar + 2 * mcon];
    const double* sx_ = slot;
    const double* sz_ = slot + nvar;
    const double* sy_ = slot + nvar + mcon;
    if (act) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double e = scr[(SCR_ED + j) * T];
        s.x[j] = scr[(SCR_DX + j) * T] * sx_[3 * k + j];
        zd0[j] = sz_[3 * k + j] / e;
        s.yd[j] = e * sy_[3 * k + j] * cinv;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double e = scr[(SCR_EG + r) * T];
        s.zg[r] = sz_[3 * (N + 1) + 2 * k + r] / e;
        s.yg[r] = e * sy_[3 * (N + 1) + 2 * k + r] * cinv;
      }
      if constexpr (SBOX) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double e = scr[(SCR_ES + j) * T];
          s.zs[j] = sz_[row_s0 + 3 * k + j] / e;
          s.ys[j] = e * sy_[row_s0 + 3 * k + j] * cinv;
        }
      }
    }
    if (actu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double e = scr[(SCR_EB + j) * T];
        s.u[j] = scr[(SCR_DU + j) * T] * sx_[3 * (N + 1) + 2 * k + j];
        s.zb[j] = sz_[5 * (N + 1) + 2 * k + j] / e;
        s.yb[j] = e * sy_[5 * (N + 1) + 2 * k + j] * cinv;
      }
      if constexpr (RATE) {
        const double e = scr[SCR_ER * T];
        s.zr = sz_[row_r0 + k] / e;
        s.yr = e * sy_[row_r0 + k] * cinv;
      }
    }
  }

#pragma unroll
  for (int j = 0; j < 3; ++j) scr[(SCR_ZD + j) * T] = zd0[j];
  bool first_iter = true;   // uniform over the warp (QPs that share a warp start together)

  // ---------------- main loop (OSQP osqp_solve); factor / update_info / check each have ONE call site ----------
  int status = ST_UNSOLVED;
  int iter = 1, n_rho_updates = 0;
  int ct_left = p.check_termination > 0 ? p.check_termination : -1;  // countdown to the next termination check
  const int ari = p.adaptive_rho ? p.adaptive_rho_interval : 0;
  int ar_left = ari > 0 ? ari : -1;
  bool need_factor = true;
  double pri_res = 0, dua_res = 0, obj = 0;
  const double al = p.alpha, oma = p.one_minus_alpha;

  // ---------------- store (OSQP store_solution: NaN + cold start when infeasible) ---------------------------
  // Runs once per QP: after the loop, or — when several QPs share the warp — the moment this QP terminates.
  auto store = [&]() {
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
        if constexpr (SBOX) {
#pragma unroll
          for (int j = 0; j < 3; ++j) yo[row_s0 + 3 * k + j] = has_sol ? s.ys[j] : qnan;
        }
      }
      if (actu) {
#pragma unroll
        for (int j = 0; j < 2; ++j) yo[5 * (N + 1) + 2 * k + j] = has_sol ? s.yb[j] : qnan;
        if constexpr (RATE) yo[row_r0 + k] = has_sol ? s.yr : qnan;
      }
    }
    if (k == 0) {
      if (p.u0_out) { p.u0_out[2 * (size_t)qp] = has_sol ? s.u[0] : qnan; p.u0_out[2 * (size_t)qp + 1] = has_sol ? s.u[1] : qnan; }
      if (p.status) p.status[qp] = status;
      if (p.iters) p.iters[qp] = iter;
      if (p.rho_updates) p.rho_updates[qp] = n_rho_updates;
      if (p.info) {
        double* io = p.info + 4 * (size_t)qp;
        io[0] = obj; io[1] = pri_res; io[2] = dua_res; io[3] = rho_bar;
      }
      if (p.packed) {
        double* po = p.packed + 4 * (size_t)qp;
        po[0] = has_sol ? s.u[0] : qnan; po[1] = has_sol ? s.u[1] : qnan; po[2] = (double)status; po[3] = (double)iter;
      }
    }
    if (slot) {
      // scaled iterates for the next warm start; zeros (cold start) when there is no solution
      double* sx_ = slot;
      double* sz_ = slot + nvar;
      double* sy_ = slot + nvar + mcon;
      if (act) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double e = scr[(SCR_ED + j) * T];
          sx_[3 * k + j] = has_sol ? s.x[j] / scr[(SCR_DX + j) * T] : 0.0;
          sz_[3 * k + j] = has_sol ? e * s.bd[j] : 0.0;   // z = l = u on the dynamics rows
          sy_[3 * k + j] = has_sol ? c * s.yd[j] / e : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const double e = scr[(SCR_EG + r) * T];
          sz_[3 * (N + 1) + 2 * k + r] = has_sol ? e * s.zg[r] : 0.0;
          sy_[3 * (N + 1) + 2 * k + r] = has_sol ? c * s.yg[r] / e : 0.0;
        }
        if constexpr (SBOX) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double e = scr[(SCR_ES + j) * T];
            sz_[row_s0 + 3 * k + j] = has_sol ? e * s.zs[j] : 0.0;
            sy_[row_s0 + 3 * k + j] = has_sol ? c * s.ys[j] / e : 0.0;
          }
        }
      }
      if (actu) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double e = scr[(SCR_EB + j) * T];
          sx_[3 * (N + 1) + 2 * k + j] = has_sol ? s.u[j] / scr[(SCR_DU + j) * T] : 0.0;
          sz_[5 * (N + 1) + 2 * k + j] = has_sol ? e * s.zb[j] : 0.0;
          sy_[5 * (N + 1) + 2 * k + j] = has_sol ? c * s.yb[j] / e : 0.0;
        }
        if constexpr (RATE) {
          const double e = scr[SCR_ER * T];
          sz_[row_r0 + k] = has_sol ? e * s.zr : 0.0;
          sy_[row_r0 + k] = has_sol ? c * s.yr / e : 0.0;
        }
      }
      if (k == 0) { slot[nvar + 2 * mcon] = rho_bar; slot[nvar + 2 * mcon + 1] = 1.0; }
    }
  };

  // ---------- one ADMM iteration (OSQP update_xz_tilde / update_x / update_z / update_y) ---------------------
  // A solve's FIRST iteration is its own instantiation: it alone sees a z on the dynamics rows that is not their right-hand side
  // (0 on a cold start, the slot's value on a warm start; read from the scratch line).  Every later iteration uses z = b.
  // ---------- (partitioned solve) y = T_own^-1 r by cyclic reduction inside the warp; mc = level 0's multipliers, load already issued ---
  [[maybe_unused]] Comm<1, 32> cw(nullptr, tid & 31);
  [[maybe_unused]] auto pcr_own = [&](double2 (&mc)[9], double (&r)[3], double (&y)[3]) {
    if constexpr (PART) {
      tmem_wait_ld();
      tmem_tie<9>(mc);
#pragma unroll
      for (int lev = 0; lev < NLEVP - 1; ++lev) {
        const int h = 1 << lev;
        double lo[3], hi[3];
        double2 nx[9];
        if (lev + 1 < NLEVP - 1) tmem_ld_pairs<9>(tmb + 36 * (lev + 1), nx);
        else { tmem_ld_pairs<8>(tmb + 36 * (NLEVP - 1), nx); nx[8] = make_double2(0.0, 0.0); }
        cw.template both<3>(r, lo, hi, h);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double2 c0 = mc[3 * i], c1 = mc[3 * i + 1], c2 = mc[3 * i + 2];
          double a = fma(c0.x, lo[0], r[i]);
          a = fma(c1.y, hi[0], a);
          a = fma(c0.y, lo[1], a);
          a = fma(c2.x, hi[1], a);
          a = fma(c1.x, lo[2], a);
          r[i] = fma(c2.y, hi[2], a);
        }
        tmem_wait_ld();
        tmem_tie<9>(nx);
#pragma unroll
        for (int j = 0; j < 9; ++j) mc[j] = nx[j];
      }
      {  // top level: single neighbour lane ^ 16
        constexpr int h = 1 << (NLEVP - 1);
        double nb[3];
        cw.template xr<3>(r, nb, h);
        r[0] = fma(mc[1].x, nb[2], fma(mc[0].y, nb[1], fma(mc[0].x, nb[0], r[0])));
        r[1] = fma(mc[2].y, nb[2], fma(mc[2].x, nb[1], fma(mc[1].y, nb[0], r[1])));
        r[2] = fma(mc[4].x, nb[2], fma(mc[3].y, nb[1], fma(mc[3].x, nb[0], r[2])));
      }
      const double b0 = mc[5].x, b1 = mc[5].y, b2 = mc[6].x, b4 = mc[6].y, b5 = mc[7].x, b8 = mc[7].y;
      y[0] = b0 * r[0] + b1 * r[1] + b2 * r[2];
      y[1] = b1 * r[0] + b4 * r[1] + b5 * r[2];
      y[2] = b2 * r[0] + b5 * r[1] + b8 * r[2];
    }
  };

  // ---------- (partitioned solve) spike correction of one partition level: y += WT y_other + WO y_own, the two boundary values from
  // one exchange through the QP's boundary slots; WT, WO from the strip ---
  [[maybe_unused]] auto spike_correct = [&](auto level_c, double (&y)[3]) {
    if constexpr (PART) {
      constexpr int LEVEL = decltype(level_c)::value;
      double2 wv[9];
      tmem_ld_pairs<9>(tmb + 4 * (W_PAIR + 9 * (LEVEL - 1)), wv);
      double yo[3], yt[3];
      cm.template boundary<LEVEL>(y, yo, yt);
      tmem_wait_ld();
      tmem_tie<9>(wv);
      const double wt[9] = {wv[0].x, wv[0].y, wv[1].x, wv[1].y, wv[2].x, wv[2].y, wv[3].x, wv[3].y, wv[4].x};
      const double wo[9] = {wv[4].y, wv[5].x, wv[5].y, wv[6].x, wv[6].y, wv[7].x, wv[7].y, wv[8].x, wv[8].y};
#pragma unroll
      for (int i = 0; i < 3; ++i)
        y[i] = fma(wo[3 * i + 2], yo[2], fma(wt[3 * i + 2], yt[2], fma(wo[3 * i + 1], yo[1], fma(wt[3 * i + 1], yt[1],
               fma(wo[3 * i], yo[0], fma(wt[3 * i], yt[0], y[i]))))));
    }
  };

  auto iterate = [&](auto first_c) {
    constexpr bool FIRST = decltype(first_c)::value;
      // tensor-memory variant: the multipliers are fetched one PCR level ahead of their use; level 0 flies during the rhs assembly
      [[maybe_unused]] double2 mc[9];
      [[maybe_unused]] double2 mrow[4];   // steering-rate variant: one row (four pairs) of the current level
      if constexpr (TM && RATE) {
        static_assert(!(TM && RATE) || NLEV > 1, "steering-rate rows in tensor memory need a two-sided level");
        tmem_ld_pairs<4>(tmb, mrow);
      } else if constexpr (TM) {
        if constexpr (NLEV > 1) tmem_ld_pairs<9>(tmb, mc);
        else { tmem_ld_pairs<8>(tmb, mc); mc[8] = make_double2(0.0, 0.0); }
      }
      // s = rho (z - y/rho) = rho z - y per row; right-hand side of the condensed system
      double sd[3], sg[2], sb[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) sd[i] = s.rd[i] * (FIRST ? scr[(SCR_ZD + i) * T] : s.bd[i]) - s.yd[i];
#pragma unroll
      for (int r = 0; r < 2; ++r) { sg[r] = s.rg[r] * s.zg[r] - s.yg[r]; sb[r] = s.rb[r] * s.zb[r] - s.yb[r]; }
      double xt[3], ut[2], ztd[3];
      [[maybe_unused]] double ztr = 0.0;
      // 1 where the stage has a predecessor: values shuffled in from stage k-1 enter through an FMA with this factor (exact: the
      // factor is 0 or 1), one select instead of two per masked value
      const double hm = hasp ? 1.0 : 0.0;
      if constexpr (RATE) {
        const double sr = s.rr * s.zr - s.yr;   // 0 on stages without an input (rho = 0, y = 0)
        double sdn[3], srn;
        {
          const double snd[4] = {sd[0], sd[1], sd[2], sr};
          double rcv[4];
          cm.template dn<4>(snd, rcv, 1);
#pragma unroll
          for (int i = 0; i < 3; ++i) sdn[i] = (LASTFULL && !actu) ? 0.0 : rcv[i];
          srn = rcv[3];
        }
        double gx[3], t3[3], t2[2];
        At_mul(md, sdn, t3);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          gx[j] = (s.sx[j] * s.x[j] - s.qx[j] - sd[j] + s.gm[j] * sg[0] + s.gm[3 + j] * sg[1]) + t3[j];
        Bt_mul(md, sdn, t2);
        const double gv = (s.su[0] * s.u[0] - qu[0] + sb[0]) + t2[0];
        double gd = (s.su[1] * s.u[1] - qu[1] + sb[1] + sr - srn) + t2[1];
        gd = actu ? gd : 0.0;
        // eliminate v_k: hv = g_v / w_v, f = R_{k+1} b_v hv
        const double hv = s.wvi * gv;
        const double mv[3] = {s.rdn[0] * md.b00, s.rdn[1] * md.b10, s.rdn[2] * md.b20};   // R_{k+1} b_v (as in the factor step)
        const double f[3] = {mv[0] * hv, mv[1] * hv, mv[2] * hv};
        double r[4], fp[3];
        At_mul(md, f, t3);
        cm.template up<3>(f, fp, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = fma(hm, fp[i], gx[i] - t3[i]);   // + fp from the predecessor, if there is one
        r[3] = gd - md.b21 * f[2];
#pragma unroll
        for (int lev = 0; lev < NLEV - 1; ++lev) {
          const int h = 1 << lev;
          double lo[4], hi[4];
          cm.template both<4>(r, lo, hi, h);
          [[maybe_unused]] const double2* cf = sm_pair + (lev * 16 - RT_SM0) * T;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double2 c0, c1, c2, c3;
            if (TM && lev < TML) {   // (lev is a constant once the level loop is unrolled)
              // the row's four pairs were fetched from tensor memory while the previous row was applied; fetch the next row now
              tmem_wait_ld();
              tmem_tie<4>(mrow);
              c0 = mrow[0]; c1 = mrow[1]; c2 = mrow[2]; c3 = mrow[3];
              if (i < 3) tmem_ld_pairs<4>(tmb + 4 * (lev * 16 + 4 * (i + 1)), mrow);
              else if (lev + 1 < TML) tmem_ld_pairs<4>(tmb + 4 * ((lev + 1) * 16), mrow);
            } else {
              c0 = cf[(4 * i + 0) * T]; c1 = cf[(4 * i + 1) * T]; c2 = cf[(4 * i + 2) * T]; c3 = cf[(4 * i + 3) * T];
            }
            double a = fma(c0.x, lo[0], r[i]);
            a = fma(c2.x, hi[0], a);
            a = fma(c0.y, lo[1], a);
            a = fma(c2.y, hi[1], a);
            a = fma(c1.x, lo[2], a);
            a = fma(c3.x, hi[2], a);
            a = fma(c1.y, lo[3], a);
            r[i] = fma(c3.y, hi[3], a);
          }
        }
        {  // top level: single neighbour k ^ h
          constexpr int h = 1 << (NLEV - 1);
          double nb[4];
          cm.template xr<4>(r, nb, h);
          const double2* cf = sm_pair + ((NLEV - 1) * 16 - RT_SM0) * T;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 c0 = cf[(2 * i) * T], c1 = cf[(2 * i + 1) * T];
            r[i] = fma(c1.y, nb[3], fma(c1.x, nb[2], fma(c0.y, nb[1], fma(c0.x, nb[0], r[i]))));
          }
        }
        double st[4];
        {
          const double2* cf = sm_pair + ((NLEV - 1) * 16 + 8 - RT_SM0) * T;
          const double2 q0 = cf[0 * T], q1 = cf[1 * T], q2 = cf[2 * T], q3 = cf[3 * T], q4 = cf[4 * T];
          const double b00 = q0.x, b01 = q0.y, b02 = q1.x, b03 = q1.y, b11 = q2.x, b12 = q2.y, b13 = q3.x, b22 = q3.y, b23 = q4.x, b33 = q4.y;
          st[0] = fma(b03, r[3], fma(b02, r[2], fma(b01, r[1], b00 * r[0])));
          st[1] = fma(b13, r[3], fma(b12, r[2], fma(b11, r[1], b01 * r[0])));
          st[2] = fma(b23, r[3], fma(b22, r[2], fma(b12, r[1], b02 * r[0])));
          st[3] = fma(b33, r[3], fma(b23, r[2], fma(b13, r[1], b03 * r[0])));
        }
        xt[0] = st[0]; xt[1] = st[1]; xt[2] = st[2];
        // recover v~_k = hv - b_v' R_{k+1} (C s~_k - x~_{k+1}) / w_v   (mv = 0 on the last stage)
        double axt[3], xn[3];
        A_mul(md, xt, axt);
        cm.template dn<3>(xt, xn, 1);
        const double yv[3] = {axt[0] - xn[0], axt[1] - xn[1], axt[2] + md.b21 * st[3] - xn[2]};
        ut[0] = hv - s.wvi * (mv[0] * yv[0] + mv[1] * yv[1] + mv[2] * yv[2]);
        ut[1] = st[3];
        // z~ = A w~
        double pred[3];
        B_mul(md, ut, pred);
        const double snd[4] = {pred[0] + axt[0], pred[1] + axt[1], pred[2] + axt[2], ut[1]};
        double rcv[4];
        cm.template up<4>(snd, rcv, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) ztd[i] = fma(hm, rcv[i], -xt[i]);
        ztr = fma(-hm, rcv[3], ut[1]);
      } else {
        double sdn[3];
        cm.template dn<3>(sd, sdn, 1);  // stages above N hold zeros, so only a full last warp needs the mask
        if (LASTFULL) {
#pragma unroll
          for (int i = 0; i < 3; ++i) sdn[i] = actu ? sdn[i] : 0.0;
        }
        double gx[3], gu[2], t3[3], t2[2];
        At_mul(md, sdn, t3);
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // the neighbour-dependent term (t3, from sdn) is added last: the rest is ready while the shuffle flies
          double own = s.sx[j] * s.x[j] - s.qx[j] - sd[j] + s.gm[j] * sg[0] + s.gm[3 + j] * sg[1];
          if constexpr (SBOX) own += s.rs[j] * s.zs[j] - s.ys[j];
          gx[j] = own + t3[j];
        }
        Bt_mul(md, sdn, t2);
#pragma unroll
        for (int j = 0; j < 2; ++j) gu[j] = (s.su[j] * s.u[j] - qu[j] + sb[j]) + t2[j];
        // eliminate u_k: h = W^-1 gu, f = R_{k+1} B h
        const double hh[2] = {s.wi[0] * gu[0] + s.wi[1] * gu[1], s.wi[1] * gu[0] + s.wi[2] * gu[1]};
        const double f[3] = {s.rbm[0] * hh[0], s.rbm[1] * hh[0], s.rbm[2] * hh[0] + s.rbm[3] * hh[1]};
        double r[3], fp[3];
        At_mul(md, f, t3);
        cm.template up<3>(f, fp, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = fma(hm, fp[i], gx[i] - t3[i]);   // + fp from the predecessor, if there is one
        if constexpr (PART) {
          // own-block solve, then the spike correction from the two boundary values (posted in the QP's exchange slots)
          double y[3];
          pcr_own(mc, r, y);
          spike_correct(std::integral_constant<int, 1>{}, y);
          if constexpr (PLEV == 2) spike_correct(std::integral_constant<int, 2>{}, y);
#pragma unroll
          for (int i = 0; i < 3; ++i) xt[i] = y[i];
        } else {
        // PCR: apply the stored multipliers level by level (fully unrolled, constant offsets)
        if constexpr (TM) { tmem_wait_ld(); tmem_tie<9>(mc); }
#pragma unroll
        for (int lev = 0; lev < NLEV - 1; ++lev) {
          const int h = 1 << lev;
          double lo[3], hi[3];
          [[maybe_unused]] double2 nx[9];
          if constexpr (TM) {   // next level (or: one-sided top level + final inverse, 8 pairs in a row) while this one is applied
            if (lev + 1 < NLEV - 1) tmem_ld_pairs<9>(tmb + 36 * (lev + 1), nx);
            else { tmem_ld_pairs<8>(tmb + 36 * (NLEV - 1), nx); nx[8] = make_double2(0.0, 0.0); }
          }
          cm.template both<3>(r, lo, hi, h);
          [[maybe_unused]] const double2* cf = (GL == 2 && lev == NLEV - 2) ? gl_pair : sm_pair + (lev * 9) * T;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double2 c0, c1, c2;
            if constexpr (TM) { c0 = mc[3 * i]; c1 = mc[3 * i + 1]; c2 = mc[3 * i + 2]; }
            else { c0 = cf[(3 * i + 0) * T]; c1 = cf[(3 * i + 1) * T]; c2 = cf[(3 * i + 2) * T]; }
            double a = fma(c0.x, lo[0], r[i]);
            a = fma(c1.y, hi[0], a);
            a = fma(c0.y, lo[1], a);
            a = fma(c2.x, hi[1], a);
            a = fma(c1.x, lo[2], a);
            r[i] = fma(c2.y, hi[2], a);
          }
          if constexpr (TM) {
            tmem_wait_ld();
            tmem_tie<9>(nx);
#pragma unroll
            for (int j = 0; j < 9; ++j) mc[j] = nx[j];
          }
        }
        {  // top level: single neighbour k ^ h
          constexpr int h = 1 << (NLEV - 1);
          double nb[3];
          cm.template xr<3>(r, nb, h);
          double2 c0, c1, c2, c3, c4;
          if constexpr (TM) { c0 = mc[0]; c1 = mc[1]; c2 = mc[2]; c3 = mc[3]; c4 = mc[4]; }
          else {
            const double2* cf = TOPG ? gl_pair + GTOP * T : sm_pair + ((NLEV - 1) * 9) * T;
            c0 = cf[0 * T]; c1 = cf[1 * T]; c2 = cf[2 * T]; c3 = cf[3 * T]; c4 = cf[4 * T];
          }
          r[0] = fma(c1.x, nb[2], fma(c0.y, nb[1], fma(c0.x, nb[0], r[0])));
          r[1] = fma(c2.y, nb[2], fma(c2.x, nb[1], fma(c1.y, nb[0], r[1])));
          r[2] = fma(c4.x, nb[2], fma(c3.y, nb[1], fma(c3.x, nb[0], r[2])));
        }
        {
          double2 q0, q1, q2;
          if constexpr (TM) { q0 = mc[5]; q1 = mc[6]; q2 = mc[7]; }
          else { q0 = sm_pair[(FINAL_PAIR + 0) * T]; q1 = sm_pair[(FINAL_PAIR + 1) * T]; q2 = sm_pair[(FINAL_PAIR + 2) * T]; }
          const double b0 = q0.x, b1 = q0.y, b2 = q1.x, b4 = q1.y, b5 = q2.x, b8 = q2.y;
          xt[0] = b0 * r[0] + b1 * r[1] + b2 * r[2];
          xt[1] = b1 * r[0] + b4 * r[1] + b5 * r[2];
          xt[2] = b2 * r[0] + b5 * r[1] + b8 * r[2];
        }
        }
        // recover u~_k = h - W^-1 B' R_{k+1} (A x~_k - x~_{k+1})   (rdn = 0 on the last stage)
        double axt[3], v[3], xn[3];
        A_mul(md, xt, axt);
        cm.template dn<3>(xt, xn, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = axt[i] - xn[i];
        t2[0] = s.rbm[0] * v[0] + s.rbm[1] * v[1] + s.rbm[2] * v[2];   // (R_{k+1} B)' v
        t2[1] = s.rbm[3] * v[2];
        ut[0] = hh[0] - (s.wi[0] * t2[0] + s.wi[1] * t2[1]);
        ut[1] = hh[1] - (s.wi[1] * t2[0] + s.wi[2] * t2[1]);
        // z~ = A w~ : dynamics rows need the predecessor's prediction
        double pred[3], pp[3];
        B_mul(md, ut, pred);
#pragma unroll
        for (int i = 0; i < 3; ++i) pred[i] += axt[i];
        cm.template up<3>(pred, pp, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) ztd[i] = fma(hm, pp[i], -xt[i]);
      }
      const double ztg[2] = {s.gm[0] * xt[0] + s.gm[1] * xt[1] + s.gm[2] * xt[2], s.gm[3] * xt[0] + s.gm[4] * xt[1] + s.gm[5] * xt[2]};
#pragma unroll
      for (int j = 0; j < 3; ++j) s.x[j] = al * xt[j] + oma * s.x[j];
#pragma unroll
      for (int j = 0; j < 2; ++j) s.u[j] = al * ut[j] + oma * s.u[j];
      if constexpr (FIRST) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {  // equality rows: the projection onto [l, u] = {b} is b itself
          const double zr = al * ztd[i] + oma * scr[(SCR_ZD + i) * T];
          s.yd[i] += s.rd[i] * (zr - s.bd[i]);
        }
      } else {
        // z = b from the second iteration on:  alpha z~ + (1 - alpha) b - b = alpha (z~ - b)
#pragma unroll
        for (int i = 0; i < 3; ++i) s.yd[i] = fma(s.rda[i], ztd[i] - s.bd[i], s.yd[i]);
      }
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        const double zr = al * ztg[r2] + oma * s.zg[r2];
        const double zn = dmax(zr + s.ig[r2] * s.yg[r2], s.gl[r2]);  // upper bound is +INFTY: the projection is a max
        s.yg[r2] += s.rg[r2] * (zr - zn);
        s.zg[r2] = zn;
        const double zrb = al * ut[r2] + oma * s.zb[r2];
        const double znb = clampd(zrb + s.ib[r2] * s.yb[r2], p.u_min[r2], p.u_max[r2]);
        s.yb[r2] += s.rb[r2] * (zrb - znb);
        s.zb[r2] = znb;
      }
      if constexpr (SBOX) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // z~ = x~_k[j]
          const double zr = al * xt[j] + oma * s.zs[j];
          const double zn = clampd(zr + s.is[j] * s.ys[j], s.slo[j], s.shi[j]);
          s.ys[j] += s.rs[j] * (zr - zn);
          s.zs[j] = zn;
        }
      }
      if constexpr (RATE) {
        const double zr = al * ztr + oma * s.zr;
        const double zn = clampd(zr + s.ir * s.yr, s.rbase - p.rate_delta, s.rbase + p.rate_delta);
        s.yr += s.rr * (zr - zn);
        s.zr = zn;
      }
      };

  for (;;) {
    if (QPW == 1 ? need_factor : __any_sync(FULL, need_factor)) {   // (re-factoring an unchanged rho is idempotent)
      // ---------- factor step: metric from rho_bar, input elimination, PCR multipliers -------------------------
      need_factor = false;
#pragma unroll
      for (int i = 0; i < 3; ++i) { s.rd[i] = RHO_EQ_OVER_RHO_INEQ * rho_bar * scr[(SCR_WD + i) * T]; s.rda[i] = s.rd[i] * al; }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double cg = scr[(SCR_CG + r) * T], cb = scr[(SCR_CB + r) * T];
        const double rg = cg < 0 ? RHO_MIN : (cg > 0 ? RHO_EQ_OVER_RHO_INEQ * rho_bar : rho_bar);
        const double rb = cb < 0 ? RHO_MIN : (cb > 0 ? RHO_EQ_OVER_RHO_INEQ * rho_bar : rho_bar);
        s.rg[r] = rg * scr[(SCR_WG + r) * T]; s.ig[r] = rcp_pos(s.rg[r]);
        s.rb[r] = actu ? rb * scr[(SCR_WB + r) * T] : 0.0;
        s.ib[r] = actu ? rcp_pos(s.rb[r]) : 0.0;
      }
      if constexpr (SBOX) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double cs = scr[(SCR_CS + j) * T];
          const double rsb = cs < 0 ? RHO_MIN : (cs > 0 ? RHO_EQ_OVER_RHO_INEQ * rho_bar : rho_bar);
          s.rs[j] = act ? rsb * scr[(SCR_WS + j) * T] : 0.0;
          s.is[j] = act ? rcp_pos(s.rs[j]) : 0.0;
        }
      }
      if constexpr (RATE) {
        // ---- steering-rate variant: eliminate the speed v_k only; reduced unknown s_k = (x_k, delta_k), 4x4 blocks ----
        {
          const double cr = scr[SCR_CR * T];
          const double rrb = cr < 0 ? RHO_MIN : (cr > 0 ? RHO_EQ_OVER_RHO_INEQ * rho_bar : rho_bar);
          s.rr = actu ? rrb * scr[SCR_WR * T] : 0.0;
          s.ir = actu ? rcp_pos(s.rr) : 0.0;
          const double snd[4] = {s.rd[0], s.rd[1], s.rd[2], s.rr};
          double rcv[4];
          cm.template dn<4>(snd, rcv, 1);
#pragma unroll
          for (int i = 0; i < 3; ++i) s.rdn[i] = actu ? rcv[i] : 0.0;
          s.rrn = (k + 1 < N) ? rcv[3] : 0.0;
        }
        const double bv[3] = {md.b00, md.b10, md.b20};   // column of B that multiplies the speed
        double wv = p.R[0] + s.su[0] + s.rb[0];
        double mv[3];   // R_{k+1} b_v
#pragma unroll
        for (int i = 0; i < 3; ++i) { mv[i] = s.rdn[i] * bv[i]; wv += mv[i] * bv[i]; }
        s.wvi = actu ? rcp_pos(wv) : 0.0;
        double Rn[9];  // R~_{k+1} = diag(rdn) - (rdn.b_v)(rdn.b_v)' / w_v
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int l = 0; l < 3; ++l) Rn[3 * i + l] = (i == l ? s.rdn[i] : 0.0) - mv[i] * mv[l] * s.wvi;
        double Rt[9];  // R~_k: from stage k-1, or diag(rho_d) for the x_0 = x_cur rows
        cm.template up<9>(Rn, Rt, 1);
#pragma unroll
        for (int e = 0; e < 9; ++e) Rt[e] = (k == 0) ? ((e % 4 == 0) ? s.rd[e / 4] : 0.0) : (act ? Rt[e] : 0.0);
        // C = [A | b_delta] (3x4): x_{k+1} = C s_k + b_v v_k.   RnC = R~_{k+1} C,  RtC = R~_k C
        double RnC[12], RtC[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          RnC[4 * i + 0] = Rn[3 * i]; RnC[4 * i + 1] = Rn[3 * i + 1];
          RnC[4 * i + 2] = Rn[3 * i] * md.a02 + Rn[3 * i + 1] * md.a12 + Rn[3 * i + 2];
          RnC[4 * i + 3] = Rn[3 * i + 2] * md.b21;
          RtC[4 * i + 0] = Rt[3 * i]; RtC[4 * i + 1] = Rt[3 * i + 1];
          RtC[4 * i + 2] = Rt[3 * i] * md.a02 + Rt[3 * i + 1] * md.a12 + Rt[3 * i + 2];
          RtC[4 * i + 3] = Rt[3 * i + 2] * md.b21;
        }
        double Bm[16], Lm[16], Um[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // C' (R~ C)
          const double c0 = RnC[j], c1 = RnC[4 + j], c2 = RnC[8 + j];
          Bm[j] = c0; Bm[4 + j] = c1; Bm[8 + j] = md.a02 * c0 + md.a12 * c1 + c2; Bm[12 + j] = md.b21 * c2;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int l = 0; l < 3; ++l)
            Bm[4 * i + l] += (i == l ? p.Q[i] + s.sx[i] : 0.0) + s.rg[0] * s.gm[i] * s.gm[l] + s.rg[1] * s.gm[3 + i] * s.gm[3 + l] + Rt[3 * i + l];
        Bm[15] += p.R[1] + s.su[1] + s.rb[1] + s.rr + s.rrn;
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // U_k = -C' R~_{k+1} [I 0] - rho_r(k+1) e4 e4'
          Um[j] = -Rn[j]; Um[4 + j] = -Rn[3 + j];
          Um[8 + j] = -(md.a02 * Rn[j] + md.a12 * Rn[3 + j] + Rn[6 + j]);
          Um[12 + j] = -md.b21 * Rn[6 + j];
        }
        Um[3] = 0.0; Um[7] = 0.0; Um[11] = 0.0; Um[15] = -s.rrn;
#pragma unroll
        for (int e = 0; e < 12; ++e) Lm[e] = hasp ? -RtC[e] : 0.0;   // L_k = U_{k-1}'
        Lm[12] = 0.0; Lm[13] = 0.0; Lm[14] = 0.0; Lm[15] = hasp ? -s.rr : 0.0;
        if (!act) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { Bm[e] = (e % 5 == 0) ? 1.0 : 0.0; Lm[e] = 0.0; Um[e] = 0.0; }
        }
#pragma unroll 1
        for (int lev = 0; lev < NLEV; ++lev) {
          const int h = 1 << lev;
          const bool vlo = act && (k - h >= 0);
          const bool vhi = act && (k + h <= N);
          // six products per level, as in the 3x3 variant: alpha, gamma, then the reduced blocks from the neighbours' U, L
          // (ordered so that few 4x4 temporaries are alive at once: the register peak of this rare step decides what the hot loop spills)
          double alp[16], gam[16], nlo[16], nhi[16];
          {
            double Bi[16];
            inv_spdD<4>(Bm, Bi);
            cm.template both<8>(Bi, nlo, nhi, h);
            cm.template both<8>(Bi + 8, nlo + 8, nhi + 8, h);
          }
          mmD<4>(Lm, nlo, alp);
          mmD<4>(Um, nhi, gam);
#pragma unroll
          for (int e = 0; e < 16; ++e) {   // a missing neighbour contributes nothing: masking alpha, gamma masks every product below
            alp[e] = vlo ? alp[e] : 0.0;
            gam[e] = vhi ? gam[e] : 0.0;
          }
          cm.template both<8>(Um, nlo, nhi, h);
          cm.template both<8>(Um + 8, nlo + 8, nhi + 8, h);
          {
            double t1[16];
            mmD<4>(alp, nlo, t1);
#pragma unroll
            for (int e = 0; e < 16; ++e) Bm[e] -= t1[e];
          }
          double Un[16];
          mmD<4>(gam, nhi, Un);
          cm.template both<8>(Lm, nlo, nhi, h);
          cm.template both<8>(Lm + 8, nlo + 8, nhi + 8, h);
          {
            double t2[16];
            mmD<4>(gam, nhi, t2);
#pragma unroll
            for (int e = 0; e < 16; ++e) Bm[e] -= t2[e];
          }
          {
            double Ln[16];
            mmD<4>(alp, nlo, Ln);
#pragma unroll
            for (int e = 0; e < 16; ++e) { Lm[e] = -Ln[e]; Um[e] = -Un[e]; }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            alp[e] = -alp[e];        // stored negated: the solve is r += coef * neighbour
            gam[e] = -gam[e];
          }
          if (lev < NLEV - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (TM && lev < TML) {
                const uint32_t dst = tmb + 4 * (lev * 16 + 4 * i);
                tmem_st_pair(dst, alp[4 * i], alp[4 * i + 1]);
                tmem_st_pair(dst + 4, alp[4 * i + 2], alp[4 * i + 3]);
                tmem_st_pair(dst + 8, gam[4 * i], gam[4 * i + 1]);
                tmem_st_pair(dst + 12, gam[4 * i + 2], gam[4 * i + 3]);
              } else if constexpr (!TM || TML < NLEV - 1) {
                sm_pair[(lev * 16 + 4 * i + 0 - RT_SM0) * T] = make_double2(alp[4 * i], alp[4 * i + 1]);
                sm_pair[(lev * 16 + 4 * i + 1 - RT_SM0) * T] = make_double2(alp[4 * i + 2], alp[4 * i + 3]);
                sm_pair[(lev * 16 + 4 * i + 2 - RT_SM0) * T] = make_double2(gam[4 * i], gam[4 * i + 1]);
                sm_pair[(lev * 16 + 4 * i + 3 - RT_SM0) * T] = make_double2(gam[4 * i + 2], gam[4 * i + 3]);
              }
            }
          } else {
            // top level: the stage has its k-h or its k+h neighbour, never both (stage k ^ h): one 4x4 block
#pragma unroll
            for (int q = 0; q < 8; ++q)
              sm_pair[(lev * 16 + q - RT_SM0) * T] = make_double2(alp[2 * q] + gam[2 * q], alp[2 * q + 1] + gam[2 * q + 1]);
          }
        }
        {
          double Bi[16];
          inv_spdD<4>(Bm, Bi);
          constexpr int FB = (NLEV - 1) * 16 + 8 - RT_SM0;   // symmetric: 10 distinct entries
          sm_pair[(FB + 0) * T] = make_double2(Bi[0], Bi[1]);
          sm_pair[(FB + 1) * T] = make_double2(Bi[2], Bi[3]);
          sm_pair[(FB + 2) * T] = make_double2(Bi[5], Bi[6]);
          sm_pair[(FB + 3) * T] = make_double2(Bi[7], Bi[10]);
          sm_pair[(FB + 4) * T] = make_double2(Bi[11], Bi[15]);
        }
      } else {
        cm.template dn<3>(s.rd, s.rdn, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) s.rdn[i] = actu ? s.rdn[i] : 0.0;
        double Rn[9];  // R~_{k+1} = diag(rdn) - (rdn.B) W^-1 (rdn.B)'
        {
          const double w00 = p.R[0] + s.su[0] + s.rb[0] + md.b00 * md.b00 * s.rdn[0] + md.b10 * md.b10 * s.rdn[1] + md.b20 * md.b20 * s.rdn[2];
          const double w01 = md.b20 * md.b21 * s.rdn[2];
          const double w11 = p.R[1] + s.su[1] + s.rb[1] + md.b21 * md.b21 * s.rdn[2];
          const double idet = rcp_pos(w00 * w11 - w01 * w01);
          s.wi[0] = actu ? w11 * idet : 0.0;
          s.wi[1] = actu ? -w01 * idet : 0.0;
          s.wi[2] = actu ? w00 * idet : 0.0;
          const double M[6] = {s.rdn[0] * md.b00, 0.0, s.rdn[1] * md.b10, 0.0, s.rdn[2] * md.b20, s.rdn[2] * md.b21};
          s.rbm[0] = M[0]; s.rbm[1] = M[2]; s.rbm[2] = M[4]; s.rbm[3] = M[5];   // R_{k+1} B, used by every iteration
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
        double Rt[9];  // R~_k: from stage k-1, or diag(rho_d) for the x_0 = x_cur rows
        cm.template up<9>(Rn, Rt, 1);
#pragma unroll
        for (int e = 0; e < 9; ++e) Rt[e] = (k == 0) ? ((e % 4 == 0) ? s.rd[e / 4] : 0.0) : (act ? Rt[e] : 0.0);
        double Bm[9], Lm[9], Um[9];
        {
          // Hx = diag(Q + sigma_x) + G' diag(rho_g) G  (+ R~_k)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int l = 0; l < 3; ++l)
              Bm[3 * i + l] = (i == l ? p.Q[i] + s.sx[i] : 0.0) + s.rg[0] * s.gm[i] * s.gm[l] + s.rg[1] * s.gm[3 + i] * s.gm[3 + l] + Rt[3 * i + l];
          if constexpr (SBOX) {   // identity rows: rho on the diagonal
#pragma unroll
            for (int i = 0; i < 3; ++i) Bm[4 * i] += s.rs[i];
          }
          double ARn[9];  // A' Rn
#pragma unroll
          for (int l = 0; l < 3; ++l) {
            ARn[0 + l] = Rn[0 + l];
            ARn[3 + l] = Rn[3 + l];
            ARn[6 + l] = md.a02 * Rn[0 + l] + md.a12 * Rn[3 + l] + Rn[6 + l];
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {  // + A' Rn A ; U_k = -A' Rn
            const double r0 = ARn[3 * i], r1 = ARn[3 * i + 1], r2 = ARn[3 * i + 2];
            Bm[3 * i + 0] += r0;
            Bm[3 * i + 1] += r1;
            Bm[3 * i + 2] += r0 * md.a02 + r1 * md.a12 + r2;
            Um[3 * i + 0] = -r0; Um[3 * i + 1] = -r1; Um[3 * i + 2] = -r2;
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {  // L_k = -R~_k A
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
        // (partitioned solve) the block that couples the two warps leaves the reduction and becomes the spike's right-hand side:
        // U_31 on stage 31, L_32 on stage 32
        // (four-warp QP: level 1 couples stages 31 | 32 and 95 | 96, level 2 stages 63 | 64)
        [[maybe_unused]] double cpl[9], cpl2[9];
        if constexpr (PART) {
          const int t64 = tid & 63;
#pragma unroll
          for (int e = 0; e < 9; ++e) {
            cpl[e] = (t64 == 31) ? Um[e] : ((t64 == 32) ? Lm[e] : 0.0);
            cpl2[e] = (PLEV == 2 && tid == 63) ? Um[e] : ((PLEV == 2 && tid == 64) ? Lm[e] : 0.0);
            Um[e] = ((tid & 31) == 31) ? 0.0 : Um[e];
            Lm[e] = ((tid & 31) == 0) ? 0.0 : Lm[e];
          }
        }
        const int kq = PART ? (tid & 31) : k;   // stage index inside the system that is reduced
        auto both9 = [&](const double* v, double* lo, double* hi, int h) {
          if constexpr (PART) cw.template both<9>(v, lo, hi, h);
          else cm.template both<9>(v, lo, hi, h);
        };
        // parallel cyclic reduction; multipliers alpha, gamma go to tensor memory / shared memory
#pragma unroll 1   // (two levels per trip measured: 0.223 vs 0.221 ms per 4096 N=30 QPs)
        for (int lev = 0; lev < NLEVP; ++lev) {
          const int h = 1 << lev;
          const bool vlo = act && (kq - h >= 0);
          const bool vhi = act && (k + h <= N) && (!PART || kq + h <= 31);
          // alpha = L_k B_{k-h}^-1, gamma = U_k B_{k+h}^-1; the reduced blocks follow from them and the neighbours' L, U:
          //   B_k -= alpha U_{k-h} + gamma L_{k+h},   L_k <- -alpha L_{k-h},   U_k <- -gamma U_{k+h}      (6 products per level)
          // (ordered so that few temporaries are alive at once: the register peak of this rare step decides what the hot loop spills)
          double alp[9], gam[9], nlo[9], nhi[9];
          {
            double Bi[9];
            inv_spd3(Bm, Bi);
            both9(Bi, nlo, nhi, h);
          }
          mm3(Lm, nlo, alp);
          mm3(Um, nhi, gam);
#pragma unroll
          for (int e = 0; e < 9; ++e) {   // a missing neighbour contributes nothing: masking alpha, gamma masks every product below
            alp[e] = vlo ? alp[e] : 0.0;
            gam[e] = vhi ? gam[e] : 0.0;
          }
          both9(Um, nlo, nhi, h);
          {
            double t1[9];
            mm3(alp, nlo, t1);
#pragma unroll
            for (int e = 0; e < 9; ++e) Bm[e] -= t1[e];
          }
          double Un[9];
          mm3(gam, nhi, Un);
          both9(Lm, nlo, nhi, h);
          {
            double t2[9];
            mm3(gam, nhi, t2);
#pragma unroll
            for (int e = 0; e < 9; ++e) Bm[e] -= t2[e];
          }
          {
            double Ln[9];
            mm3(alp, nlo, Ln);
#pragma unroll
            for (int e = 0; e < 9; ++e) { Lm[e] = -Ln[e]; Um[e] = -Un[e]; }
          }
#pragma unroll
          for (int e = 0; e < 9; ++e) {
            alp[e] = -alp[e];   // stored negated: the solve is r += coef * neighbour
            gam[e] = -gam[e];
          }
          if (lev < NLEVP - 1) {
            // 9 pairs per level, each pair one 16-byte word per stage: (a0,a1) (a2,g0) (g1,g2) per row
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              if constexpr (TM) {
                const uint32_t dst = tmb + 4 * (lev * 9 + 3 * i);
                tmem_st_pair(dst, alp[3 * i], alp[3 * i + 1]);
                tmem_st_pair(dst + 4, alp[3 * i + 2], gam[3 * i]);
                tmem_st_pair(dst + 8, gam[3 * i + 1], gam[3 * i + 2]);
              } else {
                double2* dst = (GL == 2 && lev == NLEV - 2) ? gl_pair : sm_pair + (lev * 9) * T;
                dst[(3 * i + 0) * T] = make_double2(alp[3 * i], alp[3 * i + 1]);
                dst[(3 * i + 1) * T] = make_double2(alp[3 * i + 2], gam[3 * i]);
                dst[(3 * i + 2) * T] = make_double2(gam[3 * i + 1], gam[3 * i + 2]);
              }
            }
          } else {
            // top level: h = 2^(NLEV-1) > N/2, so a stage has its k-h or its k+h neighbour, never both, and that
            // neighbour is stage k ^ h.  One 3x3 block (the non-zero one) + padding: 5 pairs.
            double one[10];
#pragma unroll
            for (int e = 0; e < 9; ++e) one[e] = alp[e] + gam[e];   // exactly one of them is non-zero
            one[9] = 0.0;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              if constexpr (TM) tmem_st_pair(tmb + 4 * (lev * 9 + q), one[2 * q], one[2 * q + 1]);
              else (TOPG ? gl_pair + GTOP * T : sm_pair + (lev * 9) * T)[q * T] = make_double2(one[2 * q], one[2 * q + 1]);
            }
          }
        }
        {
          double Bi[9];
          inv_spd3(Bm, Bi);
          if constexpr (TM) {
            tmem_st_pair(tmb + 4 * (FINAL_PAIR + 0), Bi[0], Bi[1]);
            tmem_st_pair(tmb + 4 * (FINAL_PAIR + 1), Bi[2], Bi[4]);
            tmem_st_pair(tmb + 4 * (FINAL_PAIR + 2), Bi[5], Bi[8]);
          } else {
            sm_pair[(FINAL_PAIR + 0) * T] = make_double2(Bi[0], Bi[1]);
            sm_pair[(FINAL_PAIR + 1) * T] = make_double2(Bi[2], Bi[4]);
            sm_pair[(FINAL_PAIR + 2) * T] = make_double2(Bi[5], Bi[8]);
          }
        }
        if constexpr (PART) {
          // ---- spikes of one level: V = (solve below this level) [coupling block's columns], one application per column; the
          // boundary stages' V (own: Vo, the other side's: Vt) reach every lane through the QP's boundary slots
          auto spikes = [&](auto level_c, const double* cp) {
            constexpr int LEVEL = decltype(level_c)::value;
            tmem_wait_st();
            double V[9], Vo[9], Vt[9];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              double2 mc[9];
              tmem_ld_pairs<9>(tmb, mc);
              double rr[3] = {cp[c], cp[3 + c], cp[6 + c]}, y[3], yo[3], yt[3];
              pcr_own(mc, rr, y);
              if constexpr (LEVEL == 2) spike_correct(std::integral_constant<int, 1>{}, y);
              cm.template boundary<LEVEL>(y, yo, yt);
#pragma unroll
              for (int i = 0; i < 3; ++i) { V[3 * i + c] = y[i]; Vo[3 * i + c] = yo[i]; Vt[3 * i + c] = yt[i]; }
            }
            // x_other_boundary = S (y_other - Vt y_own),  S = (I - Vt Vo)^-1;   x_k = y_k - V_k x_other_boundary
            double M[9], S[9];
            mm3(Vt, Vo, M);
#pragma unroll
            for (int e = 0; e < 9; ++e) M[e] = (e % 4 == 0 ? 1.0 : 0.0) - M[e];
            {
              const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
              const double idet = 1.0 / (M[0] * c00 + M[1] * c01 + M[2] * c02);
              S[0] = c00 * idet; S[1] = (M[2] * M[7] - M[1] * M[8]) * idet; S[2] = (M[1] * M[5] - M[2] * M[4]) * idet;
              S[3] = c01 * idet; S[4] = (M[0] * M[8] - M[2] * M[6]) * idet; S[5] = (M[2] * M[3] - M[0] * M[5]) * idet;
              S[6] = c02 * idet; S[7] = (M[1] * M[6] - M[0] * M[7]) * idet; S[8] = (M[0] * M[4] - M[1] * M[3]) * idet;
            }
            double wt[9], wo[9];
            mm3(V, S, wt);
#pragma unroll
            for (int e = 0; e < 9; ++e) wt[e] = -wt[e];   // WT = -V S      (multiplies y_other)
            mm3(wt, Vt, wo);
#pragma unroll
            for (int e = 0; e < 9; ++e) wo[e] = -wo[e];   // WO = -WT Vt    (multiplies y_own)
            const double w18[18] = {wt[0], wt[1], wt[2], wt[3], wt[4], wt[5], wt[6], wt[7], wt[8],
                                    wo[0], wo[1], wo[2], wo[3], wo[4], wo[5], wo[6], wo[7], wo[8]};
#pragma unroll
            for (int q = 0; q < 9; ++q) tmem_st_pair(tmb + 4 * (W_PAIR + 9 * (LEVEL - 1) + q), w18[2 * q], w18[2 * q + 1]);
          };
          spikes(std::integral_constant<int, 1>{}, cpl);
          if constexpr (PLEV == 2) spikes(std::integral_constant<int, 2>{}, cpl2);   // (its solves use level 1's WT, WO: stored and waited for)
        }
      }
      if constexpr (TM) tmem_wait_st();   // the iteration below reads the strip back
      cm.sync();
    }

    // Plain iterations run in their own counted inner loop (its back edge touches none of the rare blocks around it and carries one
    // counter), up to and including the next iteration that needs residuals: a termination check, a rho adaptation or the last one.
    int togo = p.max_iter - iter + 1;
    if (ct_left > 0 && ct_left < togo) togo = ct_left;
    if (ar_left > 0 && ar_left < togo) togo = ar_left;
    iter += togo - 1;
    ct_left -= togo;   // (a disabled countdown is negative and stays negative)
    ar_left -= togo;
    const bool last = (iter == p.max_iter), chk = (ct_left == 0), adp = (ar_left == 0);
    if (chk) ct_left = p.check_termination;
    if (adp) ar_left = ari;
    auto keep_previous = [&]() {   // the infeasibility tests need delta x, delta y of the trip's last iteration
#pragma unroll
      for (int j = 0; j < 3; ++j) { scr[(SCR_PX + j) * T] = s.x[j]; scr[(SCR_PYD + j) * T] = s.yd[j]; }
#pragma unroll
      for (int j = 0; j < 2; ++j) { scr[(SCR_PU + j) * T] = s.u[j]; scr[(SCR_PYG + j) * T] = s.yg[j]; scr[(SCR_PYB + j) * T] = s.yb[j]; }
      if constexpr (RATE) scr[SCR_PYR * T] = s.yr;
      if constexpr (SBOX) {
#pragma unroll
        for (int j = 0; j < 3; ++j) scr[(SCR_PYS + j) * T] = s.ys[j];
      }
    };
    // (the solve's first iteration is peeled off: the loop below holds one instantiation of the iteration and nothing else)
    if (first_iter) {
      first_iter = false;
      --togo;
      if (togo == 0) keep_previous();
      iterate(std::true_type{});
    }
    if (togo != 0) {   // (the trip's last iteration stands after the loop: nothing but the iteration inside it)
      while (--togo != 0) iterate(std::false_type{});
      keep_previous();
      iterate(std::false_type{});
    }

    // ---------- residuals & norms (OSQP update_info + the norms of compute_rho_estimate) -----------------------
    double n_z, n_Ax, n_Aty, n_Px;                      // unscaled (termination)
    double s_pri, s_dua, s_z, s_Ax, s_Aty, s_Px;        // scaled (rho estimate)
    const double edv[3] = {scr[(SCR_ED + 0) * T], scr[(SCR_ED + 1) * T], scr[(SCR_ED + 2) * T]};
    const double egv[2] = {scr[(SCR_EG + 0) * T], scr[(SCR_EG + 1) * T]};
    const double ebv[2] = {scr[(SCR_EB + 0) * T], scr[(SCR_EB + 1) * T]};
    [[maybe_unused]] double erv = 0.0;
    if constexpr (RATE) erv = scr[SCR_ER * T];
    [[maybe_unused]] double esv[3] = {0.0, 0.0, 0.0};
    if constexpr (SBOX) {
#pragma unroll
      for (int j = 0; j < 3; ++j) esv[j] = scr[(SCR_ES + j) * T];
    }
    {
      double ax[3], pred[3], t3[3], t2[2];
      A_mul(md, s.x, ax);
      B_mul(md, s.u, pred);
      double Axd[3], pp[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) pred[i] += ax[i];
      [[maybe_unused]] double Axr = 0.0, yrn = 0.0;
      if constexpr (RATE) {
        const double snd[4] = {pred[0], pred[1], pred[2], s.u[1]};
        double rcv[4];
        cm.template up<4>(snd, rcv, 1);
        pp[0] = rcv[0]; pp[1] = rcv[1]; pp[2] = rcv[2];
        Axr = s.u[1] - (hasp ? rcv[3] : 0.0);
      } else {
        cm.template up<3>(pred, pp, 1);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) Axd[i] = (hasp ? pp[i] : 0.0) - s.x[i];
      const double Axg[2] = {s.gm[0] * s.x[0] + s.gm[1] * s.x[1] + s.gm[2] * s.x[2], s.gm[3] * s.x[0] + s.gm[4] * s.x[1] + s.gm[5] * s.x[2]};
      double ydn[3];
      if constexpr (RATE) {
        const double snd[4] = {s.yd[0], s.yd[1], s.yd[2], s.yr};
        double rcv[4];
        cm.template dn<4>(snd, rcv, 1);
        ydn[0] = rcv[0]; ydn[1] = rcv[1]; ydn[2] = rcv[2];
        yrn = (k + 1 < N) ? rcv[3] : 0.0;
      } else {
        cm.template dn<3>(s.yd, ydn, 1);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) ydn[i] = actu ? ydn[i] : 0.0;
      At_mul(md, ydn, t3);
      Bt_mul(md, ydn, t2);
      if constexpr (RATE) t2[1] += s.yr - yrn;   // A' y on delta_k: +1 in rate row k, -1 in rate row k+1
      const double dxv[3] = {scr[(SCR_DX + 0) * T], scr[(SCR_DX + 1) * T], scr[(SCR_DX + 2) * T]};
      const double duv[2] = {scr[(SCR_DU + 0) * T], scr[(SCR_DU + 1) * T]};
      double m_pri = 0, m_z = 0, m_Ax = 0, m_dua = 0, m_Aty = 0, m_Px = 0;
      double q_pri = 0, q_z = 0, q_Ax = 0, q_dua = 0, q_Aty = 0, q_Px = 0, o = 0;
      bool poisoned = false;  // NaN must not hide inside a compare-select max
      if (act) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double rr = fabs(Axd[i] - s.bd[i]), zz = fabs(s.bd[i]), aa = fabs(Axd[i]);   // z = b (at least one iteration has run)
          poisoned |= !(rr == rr);
          m_pri = dmax(m_pri, rr); m_z = dmax(m_z, zz); m_Ax = dmax(m_Ax, aa);
          q_pri = dmax(q_pri, edv[i] * rr); q_z = dmax(q_z, edv[i] * zz); q_Ax = dmax(q_Ax, edv[i] * aa);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const double rr = fabs(Axg[r] - s.zg[r]), zz = fabs(s.zg[r]), aa = fabs(Axg[r]);
          poisoned |= !(rr == rr);
          m_pri = dmax(m_pri, rr); m_z = dmax(m_z, zz); m_Ax = dmax(m_Ax, aa);
          q_pri = dmax(q_pri, egv[r] * rr); q_z = dmax(q_z, egv[r] * zz); q_Ax = dmax(q_Ax, egv[r] * aa);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double Atx = -s.yd[j] + t3[j] + s.gm[j] * s.yg[0] + s.gm[3 + j] * s.yg[1];
          if constexpr (SBOX) {   // identity row on x_k[j]
            Atx += s.ys[j];
            const double rr = fabs(s.x[j] - s.zs[j]), zz = fabs(s.zs[j]), aa = fabs(s.x[j]);
            poisoned |= !(rr == rr);
            m_pri = dmax(m_pri, rr); m_z = dmax(m_z, zz); m_Ax = dmax(m_Ax, aa);
            q_pri = dmax(q_pri, esv[j] * rr); q_z = dmax(q_z, esv[j] * zz); q_Ax = dmax(q_Ax, esv[j] * aa);
          }
          const double Pxx = p.Q[j] * s.x[j];
          const double dr = fabs(Pxx + s.qx[j] + Atx), ay = fabs(Atx), px = fabs(Pxx);
          poisoned |= !(dr == dr);
          m_dua = dmax(m_dua, dr); m_Aty = dmax(m_Aty, ay); m_Px = dmax(m_Px, px);
          q_dua = dmax(q_dua, dxv[j] * dr); q_Aty = dmax(q_Aty, dxv[j] * ay); q_Px = dmax(q_Px, dxv[j] * px);
          o += 0.5 * s.x[j] * Pxx + s.qx[j] * s.x[j];
        }
      }
      if (actu) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double rr = fabs(s.u[j] - s.zb[j]), zz = fabs(s.zb[j]), aa = fabs(s.u[j]);
          m_pri = dmax(m_pri, rr); m_z = dmax(m_z, zz); m_Ax = dmax(m_Ax, aa);
          q_pri = dmax(q_pri, ebv[j] * rr); q_z = dmax(q_z, ebv[j] * zz); q_Ax = dmax(q_Ax, ebv[j] * aa);
          const double Atu = t2[j] + s.yb[j];
          const double Pxu = p.R[j] * s.u[j];
          const double dr = fabs(Pxu + qu[j] + Atu), ay = fabs(Atu), px = fabs(Pxu);
          poisoned |= !(dr == dr) || !(rr == rr);
          m_dua = dmax(m_dua, dr); m_Aty = dmax(m_Aty, ay); m_Px = dmax(m_Px, px);
          q_dua = dmax(q_dua, duv[j] * dr); q_Aty = dmax(q_Aty, duv[j] * ay); q_Px = dmax(q_Px, duv[j] * px);
          o += 0.5 * s.u[j] * Pxu + qu[j] * s.u[j];
        }
        if constexpr (RATE) {
          const double rr = fabs(Axr - s.zr), zz = fabs(s.zr), aa = fabs(Axr);
          poisoned |= !(rr == rr);
          m_pri = dmax(m_pri, rr); m_z = dmax(m_z, zz); m_Ax = dmax(m_Ax, aa);
          q_pri = dmax(q_pri, erv * rr); q_z = dmax(q_z, erv * zz); q_Ax = dmax(q_Ax, erv * aa);
        }
      }
      if constexpr (WPQ == 1) {
        poisoned = cm.any(poisoned);
        pri_res = cm.rmax(m_pri); n_z = cm.rmax(m_z); n_Ax = cm.rmax(m_Ax);
        if (poisoned) pri_res = 2.0 * OSQP_INFTY;
        dua_res = cm.rmax(m_dua); n_Aty = cm.rmax(m_Aty); n_Px = cm.rmax(m_Px);
        s_pri = cm.rmax(q_pri); s_z = cm.rmax(q_z); s_Ax = cm.rmax(q_Ax);
        s_dua = c * cm.rmax(q_dua); s_Aty = c * cm.rmax(q_Aty); s_Px = c * cm.rmax(q_Px);
        obj = cm.rsum(o);
      } else {   // a QP on several warps: all thirteen maxima (the NaN flag among them) and the sum in one exchange, one barrier
        double mv[13] = {m_pri, m_z, m_Ax, m_dua, m_Aty, m_Px, q_pri, q_z, q_Ax, q_dua, q_Aty, q_Px, poisoned ? 1.0 : 0.0};
        cm.template reduce<13, 1>(mv, &o);
        pri_res = mv[12] != 0.0 ? 2.0 * OSQP_INFTY : mv[0]; n_z = mv[1]; n_Ax = mv[2];
        dua_res = mv[3]; n_Aty = mv[4]; n_Px = mv[5];
        s_pri = mv[6]; s_z = mv[7]; s_Ax = mv[8];
        s_dua = c * mv[9]; s_Aty = c * mv[10]; s_Px = c * mv[11];
        obj = o;
      }
    }

    // ---------- termination (OSQP check_termination, exact then — on the last iteration — approximate) ---------
    bool finished = false, exact_hit = false;
    if (chk || last) {
      const int passes = last ? 2 : 1;
      for (int pass = 0; pass < passes && !finished; ++pass) {
        const bool approximate = pass == 1;
        double eps_abs = p.eps_abs, eps_rel = p.eps_rel, epi = p.eps_prim_inf, edi = p.eps_dual_inf;
        if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = ST_NON_CVX; finished = true; exact_hit = !approximate; break; }
        if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
        const bool prim_ok = pri_res < eps_abs + eps_rel * dmax(n_z, n_Ax);
        const bool dual_ok = dua_res < eps_abs + eps_rel * dmax(scr[SCR_NQ * T], dmax(n_Aty, n_Px));
        bool pinf = false, dinf = false;
        if (!prim_ok) {
          // is_primal_infeasible: delta_y projected on the polar of the recession cone of [l, u]
          double dyd[3], dyg[2], dyb[2];
#pragma unroll
          for (int i = 0; i < 3; ++i) dyd[i] = s.yd[i] - scr[(SCR_PYD + i) * T];   // finite l = u: no projection
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            dyg[r] = s.yg[r] - scr[(SCR_PYG + r) * T];
            dyb[r] = s.yb[r] - scr[(SCR_PYB + r) * T];
            // upper bound infinite (scaled test): keep the non-positive part, or nothing if the lower is infinite too
            dyg[r] = (egv[r] * s.gl[r] < -INF_THRESH) ? 0.0 : dmin(dyg[r], 0.0);
            const double lbb = ebv[r] * p.u_min[r], ubb = ebv[r] * p.u_max[r];
            if (ubb > INF_THRESH) dyb[r] = (lbb < -INF_THRESH) ? 0.0 : dmin(dyb[r], 0.0);
            else if (lbb < -INF_THRESH) dyb[r] = dmax(dyb[r], 0.0);
          }
          double mx = 0.0, lhs = 0.0;
          [[maybe_unused]] double dyr = 0.0;
          if constexpr (RATE) {
            dyr = s.yr - scr[SCR_PYR * T];
            const double rlo = s.rbase - p.rate_delta, rhi = s.rbase + p.rate_delta;
            const double lbr = erv * rlo, ubr = erv * rhi;
            if (ubr > INF_THRESH) dyr = (lbr < -INF_THRESH) ? 0.0 : dmin(dyr, 0.0);
            else if (lbr < -INF_THRESH) dyr = dmax(dyr, 0.0);
            dyr = actu ? dyr : 0.0;
            mx = fabs(dyr);
            lhs = rhi * dmax(dyr, 0.0) + rlo * dmin(dyr, 0.0);
          }
          [[maybe_unused]] double dys[3] = {0.0, 0.0, 0.0};
          if constexpr (SBOX) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double d = s.ys[j] - scr[(SCR_PYS + j) * T];
              const double lbs = esv[j] * s.slo[j], ubs = esv[j] * s.shi[j];
              if (ubs > INF_THRESH) d = (lbs < -INF_THRESH) ? 0.0 : dmin(d, 0.0);
              else if (lbs < -INF_THRESH) d = dmax(d, 0.0);
              d = act ? d : 0.0;
              dys[j] = d;
              mx = dmax(mx, fabs(d));
              // (an infinite bound only ever meets a zero here: its side of delta_y was projected away above)
              lhs += (d > 0.0 ? s.shi[j] * d : 0.0) + (d < 0.0 ? s.slo[j] * d : 0.0);
            }
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) { dyd[i] = act ? dyd[i] : 0.0; mx = dmax(mx, fabs(dyd[i])); lhs += s.bd[i] * dyd[i]; }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            dyg[r] = act ? dyg[r] : 0.0; dyb[r] = actu ? dyb[r] : 0.0;
            mx = dmax(mx, dmax(fabs(dyg[r]), fabs(dyb[r])));
            lhs += s.gl[r] * dmin(dyg[r], 0.0) + p.u_max[r] * dmax(dyb[r], 0.0) + p.u_min[r] * dmin(dyb[r], 0.0);
          }
          cm.template reduce<1, 1>(&mx, &lhs);
          const double ndy = mx;  // unscaled ||dy||; OSQP's scaled-back norm is c * ndy
          const double lhs_all = lhs;
          if (c * ndy > epi && lhs_all < -epi * ndy) {
            double dn[3], t3[3], t2[2];
            [[maybe_unused]] double dyrn = 0.0;
            if constexpr (RATE) {
              const double snd[4] = {dyd[0], dyd[1], dyd[2], dyr};
              double rcv[4];
              cm.template dn<4>(snd, rcv, 1);
              dn[0] = rcv[0]; dn[1] = rcv[1]; dn[2] = rcv[2];
              dyrn = (k + 1 < N) ? rcv[3] : 0.0;
            } else {
              cm.template dn<3>(dyd, dn, 1);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) dn[i] = actu ? dn[i] : 0.0;
            At_mul(md, dn, t3);
            Bt_mul(md, dn, t2);
            if constexpr (RATE) t2[1] += dyr - dyrn;
            double m2 = 0.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) m2 = dmax(m2, fabs(-dyd[j] + t3[j] + s.gm[j] * dyg[0] + s.gm[3 + j] * dyg[1] + (SBOX ? dys[j] : 0.0)));
#pragma unroll
            for (int j = 0; j < 2; ++j) m2 = dmax(m2, fabs(t2[j] + dyb[j]));
            pinf = cm.rmax(m2) < epi * ndy;
          }
        }
        if (!dual_ok) {
          // is_dual_infeasible
          double ddx[3], ddu[2];
#pragma unroll
          for (int j = 0; j < 3; ++j) ddx[j] = s.x[j] - scr[(SCR_PX + j) * T];
#pragma unroll
          for (int j = 0; j < 2; ++j) ddu[j] = s.u[j] - scr[(SCR_PU + j) * T];
          double mx = 0.0, qd = 0.0, mp = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j) { mx = dmax(mx, fabs(ddx[j])); qd += s.qx[j] * ddx[j]; mp = dmax(mp, fabs(p.Q[j] * ddx[j])); }
          if (actu) {
#pragma unroll
            for (int j = 0; j < 2; ++j) { mx = dmax(mx, fabs(ddu[j])); qd += qu[j] * ddu[j]; mp = dmax(mp, fabs(p.R[j] * ddu[j])); }
          }
          double nm[2] = {mx, mp};
          cm.template reduce<2, 1>(nm, &qd);
          const double ndx = nm[0];
          const double qd_all = qd, mp_all = nm[1];
          if (ndx > edi && qd_all < -edi * ndx && mp_all < edi * ndx) {
            double ax[3], pred[3];
            A_mul(md, ddx, ax);
            B_mul(md, ddu, pred);
            bool bad = false;
            const double th = edi * ndx;
#pragma unroll
            for (int i = 0; i < 3; ++i) pred[i] += ax[i];
            double pp[3];
            if constexpr (RATE) {
              const double snd[4] = {pred[0], pred[1], pred[2], ddu[1]};
              double rcv[4];
              cm.template up<4>(snd, rcv, 1);
              pp[0] = rcv[0]; pp[1] = rcv[1]; pp[2] = rcv[2];
              const double a = ddu[1] - (hasp ? rcv[3] : 0.0);
              if (actu && ((erv * (s.rbase + p.rate_delta) < INF_THRESH && a > th) || (erv * (s.rbase - p.rate_delta) > -INF_THRESH && a < -th))) bad = true;
            } else {
              cm.template up<3>(pred, pp, 1);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {  // equality rows: both sides finite
              const double a = (hasp ? pp[i] : 0.0) - ddx[i];
              if (act && (a > th || a < -th)) bad = true;
            }
            if constexpr (SBOX) {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const double a = ddx[j];
                if (act && ((esv[j] * s.shi[j] < INF_THRESH && a > th) || (esv[j] * s.slo[j] > -INF_THRESH && a < -th))) bad = true;
              }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const double a = s.gm[3 * r] * ddx[0] + s.gm[3 * r + 1] * ddx[1] + s.gm[3 * r + 2] * ddx[2];
              if (act && (egv[r] * s.gl[r] > -INF_THRESH && a < -th)) bad = true;   // upper side is infinite
              const double b = ddu[r];
              if (actu && ((ebv[r] * p.u_max[r] < INF_THRESH && b > th) || (ebv[r] * p.u_min[r] > -INF_THRESH && b < -th))) bad = true;
            }
            dinf = !cm.any(bad);
          }
        }
        if (prim_ok && dual_ok) { status = approximate ? ST_SOLVED_INACC : ST_SOLVED; finished = true; }
        else if (pinf) { status = approximate ? ST_PINF_INACC : ST_PINF; obj = OSQP_INFTY; finished = true; }
        else if (dinf) { status = approximate ? ST_DINF_INACC : ST_DINF; obj = -OSQP_INFTY; finished = true; }
        exact_hit = finished && !approximate;
      }
      if (!finished && last) { status = ST_MAX_ITER; finished = true; }
      // OSQP leaves the loop BEFORE adapt_rho only when the in-loop exact check fires; the after-loop checks
      // (iteration max_iter) come after that iteration's adapt_rho
      if constexpr (QPW == 1) {
        if (chk && exact_hit) break;
      }
    }
    // (several QPs per warp: a QP whose in-loop check fired skips adapt_rho like OSQP's break does, and is stored below)
    if (adp && (QPW == 1 || !(chk && exact_hit))) {
      // compute_rho_estimate on the scaled residuals, adapt_rho
      const double pr = s_pri / (dmax(s_z, s_Ax) + 1e-10);
      const double dr = s_dua / (dmax(scr[SCR_SNQ * T], dmax(s_Aty, s_Px)) + 1e-10);
      double rho_new = rho_bar * sqrt(pr / (dr + 1e-10));
      rho_new = dmin(dmax(rho_new, RHO_MIN), RHO_MAX);
      if (rho_new > rho_bar * p.adaptive_rho_tolerance || rho_new < rho_bar / p.adaptive_rho_tolerance) {
        rho_bar = rho_new;
        ++n_rho_updates;
        need_factor = !finished && !done;
      }
    }
    if constexpr (QPW == 1) {
      if (finished) break;
    } else {
      if (finished && !done) { store(); done = true; }
      if (__all_sync(FULL, done)) break;
    }
    ++iter;
  }

  if constexpr (QPW == 1) store();
}

// One CTA per unit: a QP on WPQ warps (horizons 32..127, steering-rate rows), or QPW short-horizon QPs in one warp.
// (255 registers / 8 warps per SM is the measured optimum: capping at 224 for 9 warps costs 17 % in spills)
template <int NLEV, int WPQ, bool LASTFULL, bool RATE, int QPW, bool SBOX = false>
__global__ void __launch_bounds__(32 * WPQ, (WPQ == 1 && !RATE && !SBOX) ? ADMM_MIN_BLOCKS : 1) admm_kernel(const KParams p) {
  extern __shared__ __align__(16) double smem_all[];
  solve_unit<NLEV, WPQ, LASTFULL, RATE, QPW, false, SBOX>(p, (int)blockIdx.x, smem_all, 0u, (int)threadIdx.x, 0u);
}

// Tensor-memory variant (horizons 16..31, base row set): persistent CTAs of four warps, two per SM.  Each CTA allocates TM_COLS
// columns of tensor memory once; each warp owns a 32-row strip of it for its multipliers and fetches QPs from a work counter until
// the batch is exhausted, so warps whose QPs stop after 25 iterations take up the slack of those that need 75.  Per-warp shared
// memory: the QP's scratch line and the TMA landing zone of its parameter record.
// RATE = true: the steering-rate variant (4x4 blocks).  Its two-sided levels fill the strip, the rest of its multipliers (13 pairs) sit
// in the warp's shared memory where the base variant keeps its scratch line; its scratch line stays in global memory.
template <int RATE>
__host__ __device__ constexpr int tm_warp_head() { return RATE ? 26 * 32 : SCR_ROWS_ALLOC * 32; }   // doubles at the head of a warp's shared-memory region
template <int NLEV, bool LASTFULL, bool RATE = false>
__global__ void __launch_bounds__(128, 2) admm_kernel_tm(const KParams p) {
  extern __shared__ __align__(16) double smem_all[];
  __shared__ uint32_t tmem_base;
  const int w = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  if (w == 0) tmem_alloc(&tmem_base, TM_COLS);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tmb = tmem_base + ((uint32_t)(32 * w) << 16);
  double* const smem_w = smem_all + (size_t)w * (tm_warp_head<RATE>() + p.rec_bulk_bytes / 8 + 2);
  if (p.rec_bulk_bytes && lane == 0) mbar_init(smem_u32(smem_w + p.rec_smem_offset + p.rec_bulk_bytes / 8), 1);
  __syncwarp();
  // A batch that fits the resident warps needs no work queue: warp i of the grid solves QP i, without the atomic round trip in front
  // of the solve (about 2 us of a lone QP's 50), and the counter pair stays zero.  Larger batches: QPs are claimed one at a time,
  // when the warp is free to solve them.  (Claiming the next QP while the current one is being solved — to have its record in flight
  // early — was measured: the one-QP lookahead costs more in load balance at the batch's tail than the hidden atomic + record
  // latency gains, 0.230 vs 0.227 ms per 4096 QPs, and far more on batches under two QPs per warp.)
  const bool queue = p.B > (int)gridDim.x * 4;
  int unit = (int)blockIdx.x * 4 + w;
  uint32_t parity = 0;
  for (;;) {
    if (queue) {
      if (lane == 0) unit = atomicAdd(p.work, 1);
      unit = __shfl_sync(FULL, unit, 0);
    }
    if (unit >= p.B) break;
    solve_unit<NLEV, 1, LASTFULL, RATE, 1, true>(p, unit, smem_w, tmb, lane, parity);
    if (!queue) {
      if (p.done_flag) {   // (single-QP latency path, B = 1: this warp's results are the call's results)
        __threadfence_system();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int32_t*>(p.done_flag) = p.done_seq;
      }
      break;
    }
    if (p.rec_bulk_bytes) parity ^= 1u;
    __syncwarp();   // every lane is done with the record and the scratch line before the next QP overwrites them
  }
  // the last warp to run dry re-arms the counter pair for the next launch that uses it
  if (queue && lane == 0) {
    const int finished = atomicAdd(p.work + 1, 1);
    if (finished == (int)gridDim.x * 4 - 1) { p.work[0] = 0; p.work[1] = 0; }
  }
  tmem_fence_before_sync();
  __syncthreads();
  if (w == 0) tmem_free(tmem_base, TM_COLS);
}

// Tensor-memory variant for multi-warp QPs (horizons 32..127 of the base row set, 32..63 with steering-rate rows): a CTA of four
// warps holds 4 / WPQ QPs (two two-warp QPs side by side, or one four-warp QP), allocates 256 tensor-memory columns and keeps the PCR
// multipliers of its QPs there (all of them for the base row set, four of five two-sided levels with steering-rate rows) — the
// shared-memory kernels of these horizons were bound by the multiplier loads (LSU data pipe 73-78 % busy).  Two QPs of one CTA
// synchronise on their own hardware barriers (bar.sync 1 + q, 64).  The scratch line lives in shared memory (base row set) or in
// global memory (steering-rate rows).
template <int NLEV, int WPQ, bool LASTFULL, bool RATE = false>
__global__ void __launch_bounds__(128, 2) admm_kernel_tmw(const KParams p) {
  extern __shared__ __align__(16) double smem_all[];
  __shared__ uint32_t tmem_base;
  constexpr int T = 32 * WPQ, QPC = 4 / WPQ;
  const int w = (int)threadIdx.x >> 5;
  if (w == 0) tmem_alloc(&tmem_base, TM_COLS);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tmb = tmem_base + ((uint32_t)(32 * w) << 16);
  const int q = (int)threadIdx.x / T, tq = (int)threadIdx.x % T;
  double* const smem_q = smem_all + (size_t)q * p.tm_unit_doubles;
  if constexpr (!RATE) {
    // one QP per slot and launch: with the base row set nearly every QP of a batch runs the same number of iterations, and the
    // work queue below measured 4 % slower than plain CTAs (N=50: 0.546 vs 0.522 ms per 4096 QPs)
    const int unit = (int)blockIdx.x * QPC + q;
    if (unit < p.B) solve_unit<NLEV, WPQ, LASTFULL, RATE, 1, true>(p, unit, smem_q, tmb, tq, 0u, QPC == 1 ? 0 : 1 + q);
  } else {
    // Steering-rate rows: iteration counts spread from 25 to several hundred, so each QP slot of the CTA (its T threads, on their
    // own hardware barrier) fetches QPs from the work counter until the batch is exhausted — a slot whose QP stops early does not
    // idle next to a neighbour that runs on (N=50: 3.79 -> 3.21 ms per 4096 QPs).
    volatile int* const next = reinterpret_cast<volatile int*>(smem_q + p.tm_unit_doubles - 2);   // the slot's current QP index
    auto slot_barrier = [&]() {
      if constexpr (QPC == 1) asm volatile("bar.sync 0, %0;" ::"n"(T) : "memory");
      else if (q == 0) asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
      else asm volatile("bar.sync 2, %0;" ::"n"(T) : "memory");
    };
    if (p.rec_bulk_bytes && tq == 0) mbar_init(smem_u32(smem_q + p.rec_smem_offset + p.rec_bulk_bytes / 8), 1);
    uint32_t parity = 0;
    for (;;) {
      if (tq == 0) *next = atomicAdd(p.work, 1);
      slot_barrier();   // (also orders the barrier's initialisation before its first use)
      const int unit = *next;
      if (unit >= p.B) break;
      solve_unit<NLEV, WPQ, LASTFULL, RATE, 1, true>(p, unit, smem_q, tmb, tq, parity, QPC == 1 ? 0 : 1 + q);
      if (p.rec_bulk_bytes) parity ^= 1u;
      slot_barrier();   // every thread is done with the record, the scratch line and the slot index before the next QP overwrites them
    }
    // the last slot to run dry re-arms the counter pair for the next launch that uses it
    if (tq == 0) {
      const int finished = atomicAdd(p.work + 1, 1);
      if (finished == (int)gridDim.x * QPC - 1) { p.work[0] = 0; p.work[1] = 0; }
    }
  }
  tmem_fence_before_sync();
  __syncthreads();
  if (w == 0) tmem_free(tmem_base, TM_COLS);
}

template <int NLEV, int WPQ, bool LASTFULL, bool RATE = false>
static cudaError_t launch_tmw(const KParams& pin, cudaStream_t stream) {
  constexpr int T = 32 * WPQ, QPC = 4 / WPQ;
  KParams p = pin;
  if (!p.work) return cudaErrorInvalidValue;
  static int sms_of[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!sms_of[dev]) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    sms_of[dev] = n;
  }
  // per QP slot: the scratch line — or, with steering-rate rows (scratch line in global memory), the multipliers that do not fit
  // the tensor-memory strip: the two-sided levels beyond the fourth, the top level and the final inverse — then the exchange
  // buffers, the record's landing zone and the slot's QP index
  constexpr int RATE_SM_DOUBLES = ((NLEV - 1) * 64 <= TM_COLS ? 0 : NLEV - 1 - TM_COLS / 64) * 32 + 26;
  size_t unit = (size_t)(RATE ? RATE_SM_DOUBLES : SCR_ROWS_ALLOC) * T + Comm<WPQ, 32>::doubles();
  const int rec_even = (11 + 3 * p.N + 1) & ~1;
  p.rec_bulk_bytes = 0;
  unit = (unit + 1) & ~(size_t)1;
  if (reinterpret_cast<uintptr_t>(p.recs) % 16 == 0 && p.stride % 2 == 0 && p.stride >= rec_even) {
    p.rec_smem_offset = (int)unit;
    p.rec_bulk_bytes = rec_even * (int)sizeof(double);
    unit += (size_t)rec_even + 2;   // the record + its mbarrier
  }
  unit += 2;   // the slot's QP index
  p.tm_unit_doubles = (int)unit;
  const size_t smem = unit * QPC * sizeof(double);
  static size_t attr_smem[64] = {0};   // (one array per instantiation of this template)
  if (smem > attr_smem[dev]) {
    e = cudaFuncSetAttribute(admm_kernel_tmw<NLEV, WPQ, LASTFULL, RATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_smem[dev] = smem;
  }
  int grid = (p.B + QPC - 1) / QPC;
  if (RATE && grid > 2 * sms_of[dev]) grid = 2 * sms_of[dev];   // (steering-rate rows: persistent CTAs, two per SM)
  admm_kernel_tmw<NLEV, WPQ, LASTFULL, RATE><<<grid, 128, smem, stream>>>(p);
  return cudaGetLastError();
}

template <int NLEV, int WPQ, bool LASTFULL, bool RATE = false, int QPW = 1, bool SBOX = false>
static cudaError_t launch_one(const KParams& pin, cudaStream_t stream) {
  constexpr int T = 32 * WPQ;
  KParams p = pin;
  constexpr int GL = RATE ? 0 : (WPQ == 4 ? 2 : (WPQ == 2 ? ADMM_W2_GLOBAL_LEVELS : 0));   // top levels' multipliers in global memory (see the kernel)
  constexpr bool TOPG = GL > 0;
  size_t smem = (size_t)(RATE ? (NLEV - 1) * 32 + 26 : 2 * (NLEV * 9 - 1 - (GL >= 1 ? 5 : 0) - (GL == 2 ? 9 : 0))) * T * sizeof(double);
  if constexpr (WPQ > 1) smem += (size_t)Comm<WPQ, 32>::doubles() * sizeof(double);
  if (TOPG && !p.mult_global) return cudaErrorInvalidValue;
  // TMA staging of the record: base and stride 16-byte aligned, record rounded up to 16 bytes fits inside the stride
  const int rec_even = (11 + 3 * p.N + 1) & ~1;
  p.rec_bulk_bytes = 0;
  if (reinterpret_cast<uintptr_t>(p.recs) % 16 == 0 && p.stride % 2 == 0 && p.stride >= rec_even) {
    smem = (smem + 15) / 16 * 16;
    p.rec_smem_offset = (int)(smem / sizeof(double));
    p.rec_bulk_bytes = rec_even * (int)sizeof(double);
    smem += (size_t)(QPW - 1) * p.stride * sizeof(double) + (size_t)p.rec_bulk_bytes + 16;  // QPW records + the mbarrier
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(admm_kernel<NLEV, WPQ, LASTFULL, RATE, QPW, SBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (RATE && WPQ == 1) {
    // The 4x4 variant runs at the register cap with a few hundred bytes of spill per thread.  With the largest shared-memory
    // carve-out (5 CTAs per SM) only 28 KB of L1 is left and three quarters of the spill reloads miss it; 4 CTAs inside a
    // 164 KB carve-out leave 92 KB of L1, which holds them (ncu: profiles/r1f_rate_kernel_ncu_summary.txt).
    int carve = 72;
    if (const char* ev = std::getenv("F110_RATE_CARVEOUT")) { const int v = std::atoi(ev); if (v >= 0 && v <= 100) carve = v; }   // tuning override (percent)
    cudaError_t e = cudaFuncSetAttribute(admm_kernel<NLEV, WPQ, LASTFULL, RATE, QPW, SBOX>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (e != cudaSuccess) return e;
  }
  admm_kernel<NLEV, WPQ, LASTFULL, RATE, QPW, SBOX><<<(p.B + QPW - 1) / QPW, T, smem, stream>>>(p);
  return cudaGetLastError();
}

// Launch of the tensor-memory variant: at most two CTAs per SM (they own the SM's 512 tensor-memory columns between them),
// fewer when the batch has fewer than 8 QPs per SM.
template <int NLEV, bool LASTFULL, bool RATE = false>
static cudaError_t launch_tm(const KParams& pin, cudaStream_t stream) {
  KParams p = pin;
  if (!p.work) return cudaErrorInvalidValue;
  static int sms_of[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!sms_of[dev]) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    sms_of[dev] = n;
  }
  const int rec_even = (11 + 3 * p.N + 1) & ~1;
  p.rec_bulk_bytes = 0;
  p.rec_smem_offset = tm_warp_head<RATE>();
  if (reinterpret_cast<uintptr_t>(p.recs) % 16 == 0 && p.stride % 2 == 0 && p.stride >= rec_even) p.rec_bulk_bytes = rec_even * (int)sizeof(double);
  const size_t smem = 4 * (size_t)(tm_warp_head<RATE>() + p.rec_bulk_bytes / 8 + 2) * sizeof(double);
  static bool attr_set[64] = {false};   // (one flag array per instantiation of this template)
  if (!attr_set[dev]) {
    e = cudaFuncSetAttribute(admm_kernel_tm<NLEV, LASTFULL, RATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  int grid = (p.B + 3) / 4;
  if (grid > 2 * sms_of[dev]) grid = 2 * sms_of[dev];
  admm_kernel_tm<NLEV, LASTFULL, RATE><<<grid, 128, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace f110
