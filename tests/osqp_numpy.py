"""Independent dense numpy statement of the OSQP algorithm (Stellato et al. 2020, v0.6 defaults) used ONLY to cross-check
the C++ restatement in oracle/osqp_restated.hpp at the iterate level: dense matrices, numpy.linalg for the KKT system,
no code shared with the C++ version.  Test infrastructure."""
import numpy as np

INFTY, RHO_MIN, RHO_MAX, RHO_EQ, RHO_TOL, MIN_S, MAX_S = 1e30, 1e-6, 1e6, 1e3, 1e-4, 1e-4, 1e4


def _limit(v):
    v = np.where(v < MIN_S, 1.0, v)
    return np.where(v > MAX_S, MAX_S, v)


def ruiz(P, q, A, l, u, passes=10):
    n, m = P.shape[0], A.shape[0]
    P, q, A = P.copy(), q.copy(), A.copy()
    D, E, c = np.ones(n), np.ones(m), 1.0
    for _ in range(passes):
        dcol = np.maximum(np.abs(P).max(axis=0), np.abs(A).max(axis=0) if m else 0.0)
        erow = np.abs(A).max(axis=1)
        dt = 1.0 / np.sqrt(_limit(dcol))
        et = 1.0 / np.sqrt(_limit(erow))
        P = dt[:, None] * P * dt[None, :]
        A = et[:, None] * A * dt[None, :]
        q = dt * q
        D, E = D * dt, E * et
        ct = max(np.abs(P).max(axis=0).mean(), float(_limit(np.array([np.abs(q).max()]))[0]))
        ct = 1.0 / float(_limit(np.array([ct]))[0])
        P, q, c = P * ct, q * ct, c * ct
    return P, q, A, E * l, E * u, D, E, c


def solve(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3, rho=0.1, sigma=1e-6, alpha=1.6, max_iter=4000, check=25, interval=25, tol=5.0):
    n, m = P.shape[0], A.shape[0]
    Ps, qs, As, ls, us, D, E, c = ruiz(P, q, A, l, u)
    loose = (ls < -INFTY * MIN_S) & (us > INFTY * MIN_S)
    eq = (~loose) & (us - ls < RHO_TOL)

    def rho_vec(r):
        return np.where(loose, RHO_MIN, np.where(eq, RHO_EQ * r, r))

    def kkt(rv):
        return np.block([[Ps + sigma * np.eye(n), As.T], [As, -np.diag(1.0 / rv)]])
    rv = rho_vec(rho)
    K = kkt(rv)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    n_updates = 0
    for it in range(1, max_iter + 1):
        rhs = np.concatenate([sigma * x - qs, z - y / rv])
        sol = np.linalg.solve(K, rhs)
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rv
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        z = np.clip(zr + y / rv, ls, us)
        y = y + rv * (zr - z)
        if it % check == 0 or it % interval == 0:
            Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
            rp, rd = Ax - z, Px + qs + Aty
            pri = np.abs(rp / E).max()
            dua = np.abs(rd / D).max() / c
            if it % check == 0:
                eps_p = eps_abs + eps_rel * max(np.abs(z / E).max(), np.abs(Ax / E).max())
                eps_d = eps_abs + eps_rel * max(np.abs(qs / D).max(), np.abs(Aty / D).max(), np.abs(Px / D).max()) / c
                if pri < eps_p and dua < eps_d:
                    return dict(x=D * x, y=E * y / c, iters=it, rho_updates=n_updates, rho=rho, status=1)
            if it % interval == 0:
                pn = np.abs(rp).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
                dn = np.abs(rd).max() / (max(np.abs(qs).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
                rn = float(np.clip(rho * np.sqrt(pn / (dn + 1e-10)), RHO_MIN, RHO_MAX))
                if rn > rho * tol or rn < rho / tol:
                    rho = rn
                    rv = rho_vec(rho)
                    K = kkt(rv)
                    n_updates += 1
    return dict(x=D * x, y=E * y / c, iters=max_iter, rho_updates=n_updates, rho=rho, status=-2)
