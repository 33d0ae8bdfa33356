"""Exact solutions of the f110-mpc tracking QP by a dense active-set method with KKT certification — test infrastructure.

Purpose: pin the FIXED POINT the OSQP iteration converges to by something that shares no algorithm and no code with the OSQP
restatements (oracle/osqp_restated.hpp, tests/osqp_numpy.py).  Two independent pieces:

  * assemble(): the QP (P, q, A, l, u) written straight from the reference's builders, in numpy —
      CreateHessianMatrix (mpc.cpp:208-219), CreateGradientVector (mpc.cpp:221-229), CreateLinearConstraintMatrix /
      UpdateLinearConstraintMatrix (mpc.cpp:231-273), Create/Update{Lower,Upper}Bound (mpc.cpp:275-306), Model::Linearize
      (model.cpp:30-59) — variables [x_0..x_N | u_0..u_{N-1}], rows [dynamics | gap pairs | input box] (mpc.cpp:26-29).
  * solve_exact(): primal active-set iterations on the dense KKT system (numpy.linalg), accepted only when the KKT conditions
      hold to 1e-9: stationarity  P x + q + A' y = 0,  primal feasibility  l <= A x <= u,  dual signs  y_i >= 0 only on rows at
      their upper bound, y_i <= 0 only on rows at their lower bound, y_i = 0 elsewhere (OSQP's sign convention).
    The QP is strictly convex on the dynamics' null space (R > 0 on every input, the states follow from the inputs), so the
    certified point is THE solution, however the active set was found.
"""
import numpy as np

INFTY = 1e30                      # OsqpEigen::INFTY
DT = float(np.float32(0.01))      # float dt_ widened (mpc.h:48, mpc.cpp:73)
WHEELBASE = float(np.float32(0.3302))   # model.cpp:32
Q = np.array([10.0, 10.0, 0.0])   # params.yaml:1-3
R = np.array([0.10, 5.0])         # params.yaml:5-6
U_DES = np.array([4.5, 0.0])      # params.yaml:42-43
U_MIN = np.array([float(np.float32(3.0)), float(np.float32(-0.43))])   # constraints.cpp:20-21
U_MAX = np.array([float(np.float32(4.5)), float(np.float32(0.43))])    # constraints.cpp:18-19


def linearize(theta, v, delta, dt=DT, L=WHEELBASE):
    """Model::Linearize (model.cpp:42-55)."""
    A = np.eye(3)
    A[0, 2] = -v * np.sin(theta) * dt
    A[1, 2] = v * np.cos(theta) * dt
    B = np.zeros((3, 2))
    B[0, 0] = np.cos(theta) * dt
    B[1, 0] = np.sin(theta) * dt
    B[2, 0] = np.tan(delta) * dt / L
    B[2, 1] = v * np.cos(delta) ** -2 * dt / L
    C = np.array([v * theta * np.sin(theta) * dt, -v * theta * np.cos(theta) * dt, -delta * v * np.cos(delta) ** -2 * dt / L])
    return A, B, C


def assemble(rec, N, gap_mode=0, state_lim=None):
    """record (x0[3] | u_lin[2] | l1[3] | l2[3] | ref[3N]) -> dense P, q, A, l, u.
    state_lim = d: append the state box the reference stores but never stacks (constraints.cpp:14-17, SetXLims :108-114):
    3(N+1) identity rows on x_0..x_N, x and y within +-d of the current state, the orientation rows at +-INFTY."""
    rec = np.asarray(rec, dtype=np.float64)
    x0, ulin, l1, l2, ref = rec[0:3], rec[3:5], rec[5:8], rec[8:11], rec[11:11 + 3 * N].reshape(N, 3)
    nx, nu = 3 * (N + 1), 2 * N
    n, m = nx + nu, nx + 2 * (N + 1) + nu
    Am, Bm, Cv = linearize(x0[2], ulin[0], ulin[1])
    P = np.zeros((n, n))
    q = np.zeros(n)
    for k in range(N + 1):          # terminal weight = Q, terminal reference = ref[N-1] (mpc.cpp:214, 228)
        P[3 * k:3 * k + 3, 3 * k:3 * k + 3] = np.diag(Q)
        q[3 * k:3 * k + 3] = -Q * ref[min(k, N - 1)]
    for k in range(N):
        P[nx + 2 * k:nx + 2 * k + 2, nx + 2 * k:nx + 2 * k + 2] = np.diag(R)
        q[nx + 2 * k:nx + 2 * k + 2] = -R * U_DES
    A = np.zeros((m, n))
    l = np.zeros(m)
    u = np.zeros(m)
    A[0:3, 0:3] = -np.eye(3)        # -x_0 = -x_cur (mpc.cpp:233-235, 299)
    l[0:3] = u[0:3] = -x0
    for k in range(1, N + 1):       # A x_{k-1} + B u_{k-1} - x_k = -C (mpc.cpp:243-249, 256-265, 305)
        r = slice(3 * k, 3 * k + 3)
        A[r, 3 * (k - 1):3 * k] = Am
        A[r, nx + 2 * (k - 1):nx + 2 * k] = Bm
        A[r, 3 * k:3 * k + 3] = -np.eye(3)
        l[r] = u[r] = -Cv
    for k in range(N + 1):          # gap pair: all-ones at k = 0 (never overwritten), the two lines after (mpc.cpp:237-241, 267-272)
        r = nx + 2 * k
        if k == 0:
            A[r:r + 2, 0:3] = 1.0
        else:
            A[r, 3 * k:3 * k + 2] = l1[0:2]
            A[r + 1, 3 * k:3 * k + 2] = l2[0:2]
        on = gap_mode == 1 or (gap_mode == 2 and k > 0)
        l[r] = -l1[2] if on else -INFTY          # mpc.cpp:297-298 (the commented alternative) / as shipped
        l[r + 1] = -l2[2] if on else -INFTY
        u[r] = u[r + 1] = INFTY
    for k in range(N):              # input box (mpc.cpp:251-253, 281, 290)
        r = nx + 2 * (N + 1) + 2 * k
        A[r, nx + 2 * k] = 1.0
        A[r + 1, nx + 2 * k + 1] = 1.0
        l[r:r + 2] = U_MIN
        u[r:r + 2] = U_MAX
    if state_lim is not None:
        As = np.zeros((nx, n))
        As[:, :nx] = np.eye(nx)
        ls, us = np.full(nx, -INFTY), np.full(nx, INFTY)
        for k in range(N + 1):
            ls[3 * k:3 * k + 2] = x0[0:2] - state_lim
            us[3 * k:3 * k + 2] = x0[0:2] + state_lim
        A, l, u = np.vstack([A, As]), np.concatenate([l, ls]), np.concatenate([u, us])
    return P, q, A, l, u


def kkt_residuals(P, q, A, l, u, x, y):
    Ax = A @ x
    stat = np.abs(P @ x + q + A.T @ y).max()
    feas = max(0.0, float((l - Ax).max()), float((Ax - u).max()))
    at_u = np.abs(Ax - u) <= 1e-9 * (1 + np.abs(u))
    at_l = np.abs(Ax - l) <= 1e-9 * (1 + np.abs(l))
    sign = max(0.0, float(np.where(at_u, 0.0, np.maximum(y, 0.0)).max()), float(np.where(at_l, 0.0, np.maximum(-y, 0.0)).max()))
    return stat, feas, sign


def solve_exact(P, q, A, l, u, max_changes=400):
    """Returns (x, y, info).  Raises if the QP is infeasible or the iteration does not certify a KKT point."""
    n, m = P.shape[0], A.shape[0]
    eq = np.abs(u - l) < 1e-12
    side = np.zeros(m, dtype=int)        # 0 inactive, +1 at upper, -1 at lower; equality rows are always in
    # start from the unconstrained-in-the-inequalities solution and add the most violated rows
    for change in range(max_changes):
        act = eq | (side != 0)
        idx = np.nonzero(act)[0]
        b = np.where(eq, l, np.where(side > 0, u, l))[idx]
        Aa = A[idx]
        K = np.block([[P, Aa.T], [Aa, np.zeros((len(idx), len(idx)))]])
        rhs = np.concatenate([-q, b])
        try:
            sol = np.linalg.solve(K, rhs)
        except np.linalg.LinAlgError:
            sol = np.linalg.lstsq(K, rhs, rcond=None)[0]
        x = sol[:n]
        y = np.zeros(m)
        y[idx] = sol[n:]
        Ax = A @ x
        # dual feasibility of the working set: drop the row with the most wrong-signed multiplier
        wrong = np.where(eq, 0.0, np.where(side > 0, -y, np.where(side < 0, y, 0.0)))
        viol = np.maximum(np.maximum(Ax - u, l - Ax), 0.0)
        viol[act] = 0.0
        if viol.max() > 1e-10:
            i = int(np.argmax(viol))
            side[i] = 1 if Ax[i] > u[i] else -1
            continue
        if wrong.max() > 1e-10:
            side[int(np.argmax(wrong))] = 0
            continue
        stat, feas, sign = kkt_residuals(P, q, A, l, u, x, y)
        if stat > 1e-8 or feas > 1e-9 or sign > 1e-9:
            raise RuntimeError("active-set point fails the KKT check: stationarity %.2e feasibility %.2e sign %.2e" % (stat, feas, sign))
        return x, y, {"changes": change, "active": int((side != 0).sum()), "stationarity": stat, "feasibility": feas}
    raise RuntimeError("active set did not settle (infeasible or cycling)")
