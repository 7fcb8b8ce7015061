"""NLP stage of the oracle ("parity unpinned": no reference implementation is runnable here).  Checks that the restated
NLP has the structure of the reference's SNOPT print file, that the oracle's answers are KKT points verified by an
independent residual computation, and that scipy SLSQP reaches the same local solution from the same start."""
import os
import numpy as np
import oracle as O
import proto_numpy as P
from srb_cbf_nmpc_b200 import scenarios as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_structure_matches_snopt_print_file():
    """print_file.out:15,94-226: 133 rows = 1 objective + 28 equalities + 92 linear + 4 distance + 8 velocity at 4 stance feet;
    friction bound 0.14354, distance lower bound 1.90000, velocity bounds +-0.35000; 4981 declared derivatives = 121*41+20."""
    x0, xdes, feet, contact, obs = S.cfg1_instance(4)
    Q, f, A, b, G, h = O.pack_dense(x0, xdes, P.stance_matrix(feet, contact))
    assert Q.shape == (41, 41) and A.shape == (28, 41) and G.shape == (92, 41)
    assert 1 + A.shape[0] + G.shape[0] + 4 + 8 == 133
    assert (A.shape[0] + 1 + G.shape[0]) * 41 + 4 * 3 + 8 == 4981
    g = O.default_gains()
    assert abs(h[0] - 0.14354) < 5e-6 and abs(g.eps_dist - 1.9) < 1e-6 and abs(g.v_max - 0.35) < 1e-7
    assert g.eps_dist == float(np.float32(1.9)) and g.v_max == float(np.float32(0.35))   # float literals promoted to double
    Q2, *_ = O.pack_dense(x0, xdes, P.stance_matrix(feet, [0, 1, 1, 0]))
    assert Q2.shape == (33, 33)


def kkt_numpy(r, inst):
    x0, xdes, feet, contact, obs = inst
    Q, f, A, b, G, h = P.pack_dense(x0, xdes, P.stance_matrix(feet, contact))
    prob = P.NlpProblem(Q, f, A, b, G, h, r["obs"])
    x, y, z = r["x"], r["y"], r["z"]
    c = prob.cons(x)
    rd = prob.grad(x) + A.T @ y - prob.jac(x).T @ z
    return max(np.abs(rd).max(), np.abs(A @ x - b).max(), np.abs(c * z).max()), c.min(), z.min()


def test_oracle_returns_kkt_points():
    insts = [S.cfg1_instance(4), S.cfg1_instance(2)] + S.cfg3_instances(40)
    nconv = 0
    for inst in insts:
        obs = inst[4]
        r = O.solve_instance(*inst[:4], obs)
        if r["status"] != 0:
            assert r["status"] == 2
            continue
        nconv += 1
        e, cmin, zmin = kkt_numpy(r, inst)
        assert e < 1e-8 and cmin > 0 and zmin >= 0
    assert nconv >= 36


def test_cfg1_matches_survey_probe():
    """objective 1093.0396, distance rows 1 and 4 active at 1.9, slack ~0.8527 (scipy probe)."""
    inst = S.cfg1_instance(4)
    r = O.solve_instance(*inst)
    assert r["status"] == 0 and r["qp_iters"] == 5
    assert np.abs(r["x_qp"][:2] - [0.0032120922582, 0.100424967798]).max() < 1e-9
    x = r["x"]
    d = [(x[4 * k] - 1.0) ** 2 + (x[4 * k + 2] - 0.2) ** 2 + x[-1] for k in range(4)]
    assert abs(d[0] - 1.9) < 1e-6 and abs(d[3] - 1.9) < 1e-6 and d[1] > 1.9 + 1e-3 and d[2] > 1.9 + 1e-3
    assert abs(x[-1] - 0.8527) < 1e-3


def test_golden_scipy_agreement():
    """Independent solver (scipy SLSQP, run when the fixture was minted) reaches the same local solution."""
    g = np.load(os.path.join(GOLD, "mpc_golden.npz"))
    for i in range(len(g["nv"])):
        nv = int(g["nv"][i])
        r = O.solve_instance(g["x0"][i], g["xdes"][i], g["feet"][i], list(g["contact"][i]), g["obs"][i])
        assert r["status"] == int(g["status"][i])
        if r["status"] != 0:
            continue
        assert np.abs(r["x"][:24] - g["x_oracle"][i][:24]).max() < 1e-9
        xs = g["x_scipy"][i][:nv]
        assert np.abs(r["x"][:24] - xs[:24]).max() < 2e-5 and abs(r["x"][-1] - xs[-1]) < 1e-6
        assert abs(g["obj_oracle"][i] - g["obj_scipy"][i]) < 1e-3 * max(1.0, abs(g["obj_oracle"][i]) * 1e-4)


def test_live_scipy_agreement_on_fresh_instances():
    """The same cross-check run live on instances outside the fixture: scipy SLSQP (a sequential QP method, like the
    reference's SNOPT) started at the QP optimum lands on the oracle's local solution on every instance: X, U to 1e-4 and the
    slack variable to 1e-6 (SLSQP stops where its line search stalls at ftol = 1e-15, up to 1e-7 outside the feasible set),
    objective to 2e-6 relative."""
    from scipy.optimize import minimize
    insts = S.cfg3_instances(40, seed=20251018 + 9000)
    checked = 0
    for (x0, xdes, feet, contact, obs) in insts:
        r = O.solve_instance(x0, xdes, feet, contact, obs)
        if r["status"] != 0:
            continue
        F = P.stance_matrix(feet, contact)
        Q, f, A, b, G, h = O.pack_dense(x0, xdes, F)
        prob = P.NlpProblem(Q, f, A, b, G, h, r["obs"])
        cons = [dict(type="eq", fun=lambda v: A @ v - b, jac=lambda v: A), dict(type="ineq", fun=prob.cons, jac=prob.jac)]
        res = minimize(prob.obj, r["x_qp"], jac=prob.grad, constraints=cons, method="SLSQP", options=dict(ftol=1e-15, maxiter=500))
        fo = prob.obj(r["x"])
        assert prob.cons(res.x).min() > -1e-6
        assert abs(res.fun - fo) <= 2e-6 * max(1.0, abs(fo)), (res.fun, fo)
        assert np.abs(res.x[:24] - r["x"][:24]).max() < 1e-4 and abs(res.x[-1] - r["x"][-1]) < 1e-6
        checked += 1
    assert checked >= 30


def test_numpy_prototype_follows_c_oracle():
    for inst in [S.cfg1_instance(4)] + S.cfg3_instances(6):
        rc = O.solve_instance(*inst)
        rn = P.solve_instance(*inst)
        assert (rc["status"] == 0) == (rn["info"]["status"] == 0)
        if rc["status"] == 0:
            assert np.abs(rc["x"][:24] - rn["x"][:24]).max() < 1e-9


def _ext_problem(inst, n_obs=3, n_nbr=2, seed=0):
    """Extended-mode rows for one instance: the n_obs closest obstacles (one centre for all stages, eps 1.9) and n_nbr
    neighbours walking beside the robot (a different centre per stage, eps 2.2 = the reference's commented robot-to-robot
    threshold, dec_vars_constr_cost.h:402)."""
    x0, xdes, feet, contact, obs = inst
    rng = np.random.default_rng(seed)
    d = np.hypot(obs[0] - x0[0], obs[1] - x0[2])
    centres = [np.tile(obs[:, k], (4, 1)) for k in np.argsort(d)[:n_obs]]
    for j in range(n_nbr):
        p = np.array([x0[0], x0[2]]) + rng.uniform(0.6, 1.0) * np.array([np.cos(2.1 * j + 1.0), np.sin(2.1 * j + 1.0)])
        v = rng.uniform(-0.2, 0.2, 2)
        centres.append(np.stack([p + v * 0.043 * (k + 1) for k in range(4)]))
    eps = np.array([1.9] * n_obs + [2.2] * n_nbr)
    return np.array(centres), eps


def test_extended_rows_reduce_to_the_reference_form():
    """orc_nlp_ipm_rows with the reference's 4 rows (one centre, one per stage) is orc_nlp_ipm, bit for bit."""
    for inst in S.cfg3_instances(12):
        x0, xdes, feet, contact, obs = inst
        r = O.solve_instance(x0, xdes, feet, contact, obs)
        e = O.solve_instance_ext(x0, xdes, feet, contact, np.tile(r["obs"], (1, 4, 1)), [float(np.float32(1.9))])
        assert e["status"] == r["status"] and e["nlp_iters"] == r["nlp_iters"]
        assert np.array_equal(e["x"], r["x"]) and np.array_equal(e["z"], r["z"])


def test_extended_rows_give_kkt_points_that_scipy_confirms():
    """Extended mode (K obstacle rows and neighbour rows per stage; no reference implementation): the oracle's
    answers satisfy the KKT conditions of the enlarged NLP (independent numpy residual) and scipy SLSQP agrees."""
    from scipy.optimize import minimize
    n_ok = 0
    for k, inst in enumerate(S.cfg3_instances(16, seed=20251018 + 9500)):
        x0, xdes, feet, contact, obs = inst
        centres, eps = _ext_problem(inst, seed=k)
        r = O.solve_instance_ext(x0, xdes, feet, contact, centres, eps)
        if r["status"] != 0:
            continue
        F = P.stance_matrix(feet, contact)
        Q, f, A, b, G, h = O.pack_dense(x0, xdes, F)
        nv, mi, n_c = Q.shape[0], G.shape[0], len(eps)
        x, y, z = r["x"], r["y"], r["z"]
        vel = [1, 5, 9, 13, 3, 7, 11, 15]

        def cons(v):
            dist = [(v[4 * s] - centres[i, s, 0]) ** 2 + (v[4 * s + 2] - centres[i, s, 1]) ** 2 + v[-1] - eps[i] for i in range(n_c) for s in range(4)]
            vm = float(np.float32(0.35))
            return np.concatenate([h - G @ v, dist, vm - v[vel], vm + v[vel]])

        def jac(v):
            J = np.zeros((mi + 4 * n_c + 16, nv)); J[:mi] = -G
            for i in range(n_c):
                for s in range(4):
                    J[mi + 4 * i + s, 4 * s] = 2 * (v[4 * s] - centres[i, s, 0]); J[mi + 4 * i + s, 4 * s + 2] = 2 * (v[4 * s + 2] - centres[i, s, 1])
                    J[mi + 4 * i + s, -1] = 1.0
            for q, j in enumerate(vel):
                J[mi + 4 * n_c + q, j] = -1.0; J[mi + 4 * n_c + 8 + q, j] = 1.0
            return J
        c = cons(x)
        assert len(z) == len(c) and c.min() > -1e-9 and z.min() >= 0
        assert np.abs(Q @ x + f + A.T @ y - jac(x).T @ z).max() < 1e-8          # stationarity
        assert np.abs(A @ x - b).max() < 1e-8 and np.abs(c * z).max() < 1e-8    # feasibility, complementarity
        res = minimize(lambda v: 0.5 * v @ Q @ v + f @ v, r["x_qp"], jac=lambda v: Q @ v + f, method="SLSQP",
                       constraints=[dict(type="eq", fun=lambda v: A @ v - b, jac=lambda v: A), dict(type="ineq", fun=cons, jac=jac)],
                       options=dict(ftol=1e-15, maxiter=500))
        fo = 0.5 * x @ Q @ x + f @ x
        if cons(res.x).min() > -1e-6 and abs(res.fun - fo) <= 2e-6 * max(1.0, abs(fo)):
            assert np.abs(res.x[:24] - x[:24]).max() < 1e-3   # SLSQP's accuracy on these flatter problems; the objectives agree to 2e-6
            n_ok += 1
        else:
            assert res.fun > fo - 1e-6 * max(1.0, abs(fo)) or cons(res.x).min() < -1e-6   # scipy never finds a better feasible point
    assert n_ok >= 8


def test_weak_pin_against_the_reference_snopt_print_file():
    """The only NLP output the reference tree holds (print_file.out): see helpers.snopt_pin_checks for what is asserted and why the
    pin is weak.  Also: scipy SLSQP started AT SNOPT's printed point moves on to the oracle's answer, i.e. the printed point is an
    unfinished iterate of the same NLP, not another local solution."""
    import helpers as H
    from scipy.optimize import minimize
    fix = np.load(os.path.join(GOLD, "snopt_print_file.npz"))
    obs = fix["obs"]
    r = O.solve_instance(fix["x0"], fix["xdes"], fix["feet"], fix["contact"], obs.reshape(2, 1), mode=1)
    mG = 92
    H.snopt_pin_checks(r["x"], r["x_qp"], r["status"], r["kkt"], r["z"][mG:mG + 4] > r["s"][mG:mG + 4], fix)
    assert r["qp_iters"] <= 25 and r["qp_flag"] == 0
    Q, f, A, b, G, h = O.pack_dense(fix["x0"], fix["xdes"], fix["feet"])
    g = O.default_gains()
    idx = [1, 5, 9, 13, 3, 7, 11, 15]

    def dist(v): return np.array([(v[4 * k] - obs[0]) ** 2 + (v[4 * k + 2] - obs[1]) ** 2 + v[40] - g.eps_dist for k in range(4)])

    def djac(v):
        J = np.zeros((4, 41))
        for k in range(4):
            J[k, 4 * k] = 2 * (v[4 * k] - obs[0]); J[k, 4 * k + 2] = 2 * (v[4 * k + 2] - obs[1]); J[k, 40] = 1
        return J
    cons = [{"type": "eq", "fun": lambda v: A @ v - b, "jac": lambda v: A}, {"type": "ineq", "fun": lambda v: h - G @ v, "jac": lambda v: -G},
            {"type": "ineq", "fun": dist, "jac": djac}, {"type": "ineq", "fun": lambda v: np.concatenate([g.v_max - v[idx], g.v_max + v[idx]])}]
    xf = fix["x_final"]
    assert np.abs(A @ xf - b).max() < 2e-5 and dist(xf).min() > -3.2e-4     # SNOPT's point: feasible to its printed 3.1e-4
    res = minimize(lambda v: 0.5 * v @ Q @ v + f @ v, xf, jac=lambda v: Q @ v + f, constraints=cons, method="SLSQP",
                   options=dict(maxiter=500, ftol=1e-14))
    assert res.fun < float(fix["objective"]) - 25.0                          # 803.70 -> 773.24: SNOPT's point was not a minimum
    assert np.abs(res.x[:16] - r["x"][:16]).max() < 1e-3 and abs(res.x[40] - r["x"][40]) < 1e-3
