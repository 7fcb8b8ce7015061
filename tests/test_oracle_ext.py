"""The extended-mode checker (oracle/cbfnmpc_oracle_ext.c; SURVEY.md A.6 -- no reference implementation exists): reduces bit for bit
to the reference-faithful oracle, returns KKT points (independent numpy residual), agrees with scipy SLSQP, and its team mode ends
at a fixed point of the members' best responses."""
import numpy as np
import oracle as O
import helpers as H
from srb_cbf_nmpc_b200 import scenarios as S


def test_reduces_to_the_faithful_oracle_bit_for_bit():
    insts = S.cfg3_instances(60)
    for (x0, xdes, feet, contact, obs) in insts:
        r = O.solve_instance(x0, xdes, feet, contact, obs, mode=1)
        e = O.ext_solve(x0, xdes, np.asarray(feet)[None], contact, r["obs"].reshape(2, 1), N=4)
        assert e["status"] == r["status"] and e["nlp_iters"] == r["nlp_iters"] and e["qp_iters"] == r["qp_iters"]
        if r["status"] == 0:
            assert np.array_equal(e["x"], r["x"][:e["nv"]])


def test_pack_dense_for_eight_grid_points_extends_the_reference_pattern():
    x0, xdes, feet, contact, obs = S.cfg5_instances(1, team=False)[0]
    F = np.stack([O.pack_dense.__globals__["np"].asarray(f)[:, :] for f in feet])
    if sum(contact) == 2:
        F = F[:, :, [1, 2]] if contact[0] == 0 else F[:, :, [0, 3]]
    Q, f, A, b, G, h = O.ext_pack_dense(8, x0, xdes, F)
    c = F.shape[2]
    assert Q.shape == (8 * (6 + c) + 1,) * 2 and A.shape[0] == 56 and G.shape[0] == 16 * 8 - 4 + 2 * c * 8
    Q4, f4, A4, b4, G4, h4 = O.ext_pack_dense(4, x0, xdes[:16], F[:1])
    Qr, fr, Ar, br, Gr, hr = O.pack_dense(x0, xdes[:16], F[0])
    for m1, m2 in ((Q4, Qr), (f4, fr), (A4, Ar), (b4, br), (G4, Gr), (h4, hr)):
        assert np.array_equal(m1, m2)
    assert np.abs(np.diag(Q)[28:32] - 2000.0).max() == 0 and np.abs(np.diag(Q)[:28] - 300.0).max() == 0   # terminal weight on the last grid point
    assert (A[32:56, 48:48 + 8 * c] != 0).sum() == 2 * c * 8 + c * 8                                       # u = F lambda and 1'lambda = 1 rows


def _kkt(insts_b, r, N, nbr=None, eps_nbr=S.EPS_NBR):
    """independent numpy KKT residual of an extended-mode answer (stationarity in X, U, delta; equalities; complementarity of the rows)"""
    x0, xdes, feet, contact, obs = insts_b
    F = np.asarray(feet)
    if sum(contact) == 2:
        F = F[:, :, [1, 2]] if contact[0] == 0 else F[:, :, [0, 3]]
    Q, f, A, b, G, h = O.ext_pack_dense(N, x0, xdes, F)
    x = r["x"]; nv = r["nv"]
    g = O.default_gains()
    centres = [np.tile(obs[:, i], (N, 1)) for i in range(obs.shape[1])] + ([] if nbr is None else [nbr[j] for j in range(nbr.shape[0])])
    eps = [g.eps_dist] * obs.shape[1] + ([] if nbr is None else [eps_nbr] * nbr.shape[0])
    cmin = 1e300; comp = 0.0
    grad = Q @ x + f
    for i, (ctr, e) in enumerate(zip(centres, eps)):
        for k in range(N):
            d = np.array([x[4 * k] - ctr[k, 0], x[4 * k + 2] - ctr[k, 1]])
            cv = d @ d + x[nv - 1] - e
            z = r["zd"][i, k]
            cmin = min(cmin, cv); comp = max(comp, abs(cv * z))
            grad[4 * k] -= 2 * d[0] * z; grad[4 * k + 2] -= 2 * d[1] * z; grad[nv - 1] -= z
    return np.abs(A @ x - b).max(), cmin, comp, abs(grad[nv - 1])


def test_single_agents_with_many_rows_and_eight_grid_points_are_kkt_points():
    for N, K in ((4, 16), (8, 16)):
        insts = S.cfg5_instances(24, N=N, K_obs=K, team=False)
        res, _ = O.ext_solve_teams(*H.ext_arrays(insts), T=1, N=N, nthreads=8)
        nconv = 0
        for b, r in enumerate(res):
            if r["status"] != 0: continue
            nconv += 1
            eq, cmin, comp, gd = _kkt(insts[b], r, N)
            assert eq < 1e-9 and cmin > 0 and comp < 1e-8 and gd < 1e-8 and r["kkt"] < 1e-8
        assert nconv >= 14


def test_team_mode_ends_at_a_fixed_point_of_the_best_responses():
    """BASELINE config 4 teams: every member's part of the team answer equals its own solve against the OTHER members' final
    trajectories held fixed (from its QP start) -- and those are KKT points with the neighbour rows included."""
    N = 4
    insts = S.cfg4_teams(10)
    res, ti = O.ext_solve_teams(*H.ext_arrays(insts), T=4, N=N, eps_nbr=S.EPS_NBR, nthreads=8)
    nchk = 0
    for t0 in range(0, len(insts), 4):
        if not all(res[t0 + k]["status"] == 0 for k in range(4)): continue
        for k in range(4):
            nb = np.array([np.stack([res[t0 + j]["X"][0::4], res[t0 + j]["X"][2::4]], 1) for j in range(4) if j != k])
            x0, xdes, feet, contact, obs = insts[t0 + k]
            r = O.ext_solve(x0, xdes, feet, contact, obs, nbr=nb, N=N, eps_nbr=S.EPS_NBR)
            assert r["status"] == 0
            assert np.abs(r["X"] - res[t0 + k]["X"]).max() < 1e-6 and abs(r["slack"] - res[t0 + k]["slack"]) < 1e-6
            eq, cmin, comp, gd = _kkt(insts[t0 + k], res[t0 + k], N, nbr=nb)
            assert eq < 1e-9 and cmin > -1e-9 and comp < 2e-8 and gd < 2e-8   # slacks are measured against the PUBLISHED trajectories (within pub_tol)
            nchk += 1
    assert nchk >= 24 and ti.max() < 200


def test_scipy_slsqp_reaches_the_same_extended_solution():
    from scipy.optimize import minimize
    N, K = 8, 3
    insts = [i for i in S.cfg5_instances(12, N=N, K_obs=K, team=False) if sum(i[3]) == 4][:3]
    g = O.default_gains()
    for (x0, xdes, feet, contact, obs) in insts:
        r = O.ext_solve(x0, xdes, feet, contact, obs, N=N)
        if r["status"] != 0: continue
        Q, f, A, b, G, h = O.ext_pack_dense(N, x0, xdes, np.asarray(feet))
        nv = r["nv"]
        idx = [4 * k + 1 for k in range(N)] + [4 * k + 3 for k in range(N)]

        def dist(v):
            return np.array([(v[4 * k] - obs[0, i]) ** 2 + (v[4 * k + 2] - obs[1, i]) ** 2 + v[nv - 1] - g.eps_dist for i in range(K) for k in range(N)])
        cons = [{"type": "eq", "fun": lambda v: A @ v - b, "jac": lambda v: A}, {"type": "ineq", "fun": lambda v: h - G @ v, "jac": lambda v: -G},
                {"type": "ineq", "fun": dist}, {"type": "ineq", "fun": lambda v: np.concatenate([g.v_max - v[idx], g.v_max + v[idx]])}]
        s = minimize(lambda v: 0.5 * v @ Q @ v + f @ v, r["x_qp"], jac=lambda v: Q @ v + f, constraints=cons, method="SLSQP", options=dict(maxiter=800, ftol=1e-13))
        fo = 0.5 * r["x"] @ Q @ r["x"] + f @ r["x"]
        assert abs(s.fun - fo) < 1e-5 * max(1.0, abs(fo))
        assert np.abs(s.x[:4 * N] - r["X"]).max() < 2e-4 and abs(s.x[nv - 1] - r["slack"]) < 1e-5
