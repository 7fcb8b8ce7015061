"""Pack / model / epilogue restatements."""
import numpy as np
import oracle as O
import proto_numpy as P
from srb_cbf_nmpc_b200 import scenarios as S


def test_lip_constants():
    Ad, Bd = O.lip_discrete()
    a = np.array([[1.031273603448276, 0.04344825498275862], [1.4697495909684901, 1.031273603448276]])
    assert np.abs(Ad[:2, :2] - a).max() < 1e-15 and np.abs(Ad[2:, 2:] - a).max() < 1e-15
    assert abs(Bd[0, 0] + 0.0312736034482759) < 1e-15 and abs(Bd[1, 0] + 1.4697495909684901) < 1e-14
    assert Ad[0, 2] == 0 and Bd[0, 1] == 0


def test_pack_c_vs_numpy():
    for c in (4, 2):
        x0, xdes, feet, contact, obs = S.cfg1_instance(c)
        F = P.stance_matrix(feet, contact)
        for a, b in zip(P.pack_dense(x0, xdes, F), O.pack_dense(x0, xdes, F)):
            assert np.abs(a - b).max() < 1e-15


def test_bezier_interpolates():
    rng = np.random.default_rng(0)
    X = rng.normal(size=16); buf = rng.normal(size=4)
    r = O.Result()
    al = np.zeros(20); nb = np.zeros(4)
    O.lib().orc_bezier_fit(buf, X, al, nb)
    al = al.reshape(4, 5)
    M = P.bernstein4()
    Y = np.column_stack([buf, X.reshape(4, 4).T])
    assert np.abs(al @ M.T - Y).max() < 1e-12          # exact interpolation at t = 0, 1/4, ..., 1
    assert np.abs(nb - X[12:16]).max() == 0
    a2, nb2 = P.bezier_fit(buf, X)
    assert np.abs(a2 - al).max() < 1e-12


def test_closest_obstacle_first_minimum():
    pos = np.array([0.0, 0.0])
    obs = np.array([[3.0, 1.0, -1.0, 5.0], [0.0, 0.0, 0.0, 0.0]])
    out = np.zeros(2)
    i = O.lib().orc_closest_obstacle(pos, np.ascontiguousarray(obs.ravel()), 4, out)
    assert i == 1 and out[0] == 1.0                    # tie between columns 1 and 2: the first one wins (strict <)
    far = np.array([[2000.0], [0.0]])
    i = O.lib().orc_closest_obstacle(pos, np.ascontiguousarray(far.ravel()), 1, out)
    assert i == 0 and out[0] == 2000.0                 # min_dist starts at 1000: nothing closer -> index 0


def test_scenarios_are_seeded():
    a = S.cfg3_instances(5); b = S.cfg3_instances(5); c = P.cfg3_instances(5)
    for u, v, w in zip(a, b, c):
        assert all(np.array_equal(np.asarray(p), np.asarray(q)) for p, q in zip(u, v))
        assert all(np.array_equal(np.asarray(p), np.asarray(q)) for p, q in zip(u, w))


class _NoSolver:  # the mirror's planners are pure host logic; no GPU needed for them
    pass


def _random_team(rng, n_agents=4, ref_cols=60, n_obs=20):
    Pr = rng.normal(size=(8, ref_cols)); Prd = rng.normal(size=(8, ref_cols))
    pstart = rng.normal(size=2 * n_agents)
    obs = rng.uniform(-3, 9, size=(2, n_obs))
    return Pr, Prd, pstart, obs


def test_planner_restatement_matches_the_mirror():
    """oracle/: copPlanner_eventbase + footholdsPlanner (MPC_dist.cpp:702-782, 1204-1216) against the host mirror class."""
    from srb_cbf_nmpc_b200.mpc_dist import MPC_dist
    rng = np.random.default_rng(5)
    Pr, Prd, pstart, obs = _random_team(rng)
    for agent in range(4):
        for g in (0, 1, 2, 7, 14):
            m = MPC_dist(solver=_NoSolver())
            m.setAgentID(agent); m.setPstart(pstart); m.setPobs_real(obs); m.setReferenceTrajectory(Pr, Prd)
            m.gaitDomain_ = g
            toe = rng.normal(size=(3, 4))
            m.updateState(rng.normal(size=18), rng.normal(size=18), [1, 0, 0, 1], toe, np.zeros(4))
            x0, xdes, feet, contact, o = m._instance()
            assert np.array_equal(O.cop_planner(Pr, Prd, agent, g), xdes)
            assert np.array_equal(O.footholds(toe, g, pstart[2 * agent:2 * agent + 2]), np.asarray(feet).reshape(-1))
    # columns past the end are clamped to the last one
    assert np.array_equal(O.cop_planner(Pr, Prd, 0, 1000)[:4], [Pr[0, -1], Prd[0, -1], Pr[1, -1], Prd[1, -1]])


def test_oracle_closed_loop_small():
    """oracle/: receding-horizon loop on the template: the state follows the straight-line reference, the
    obstacle-blind QP planner fails in a world whose reference line passes an obstacle, the failing distance is the distance
    from the origin at the first grid point within 0.5 m (src/MPC_dist.cpp:21-40)."""
    w = S.mc_worlds(6, 8)
    w["Pobs_real"][:] = 100.0                      # all obstacles far away ...
    w["Pobs_real"][0, 0] = [1.0 * 0.6, 0.0]        # ... except one on world 0's path, 0.6 m from the origin line start
    w["q"][0, 0] = 0.0; w["q"][1, 0] = 0.0; w["agent_initial"][:, 0] = 0.0; w["alpha_buf"][:, 0] = 0.0
    w["Pr_refined"][0, :, 0] = 0.2 * (np.arange(w["ref_cols"]) + 1) * S.TS; w["Pr_refined"][0, :, 1] = 0.0
    w["Prd_refined"][0, :, 0] = 0.2; w["Prd_refined"][0, :, 1] = 0.0
    for i in range(4):
        w["toe_pos"][3 * i, 0] = S.NOMINAL_FEET[0, i]; w["toe_pos"][3 * i + 1, 0] = S.NOMINAL_FEET[1, i]
    traj = O.closed_loop(w, 8, mode=0)
    assert (w["gait_domain"] == 8).all() and (w["n_not_converged"] == 0).all()
    assert w["success"][0] == 0 and (w["success"][1:] == 1).all()
    assert 0.09 < w["distance_to_fail"][0] < 0.6 and (w["distance_to_fail"][1:] == 10.0).all()
    t_end = 8 * 4 * S.TS
    assert np.abs(traj[-1, 0, 1:] - (w["agent_initial"][0, 1:] + w["Prd_refined"][1:, 0, 0] * t_end)).max() < 0.05   # tracks the reference


def test_reference_generator_restatement():
    """oracle/: generateReferenceTrajectory (src/MPC_dist.cpp:930-1062) on the 4-agent scenario of src/A1_Sim.cpp:1013 /
    A1_Sim_4_agents.cpp:872-873: shapes, start, goal, the decimation quirk of the last two columns, determinism."""
    Pstart = [0.0, 0.0, 0.0, -0.9, -1, 0, -1, -0.9]
    obs = np.array([[2.2, 1, 0.85, 1, 1, 3, 3.0, 3, -100], [0.9, 1.2, -0.65, 2, -1.75, 0.2, -0.70, -2.2, -0.5 + 100]])
    Pr, Prd = O.generate_reference(Pstart, obs, loop=20000)
    assert Pr.shape == (8, 500) and Prd.shape == (8, 500)
    assert np.abs(Pr[:, 0] - np.array(Pstart)).max() < 0.05            # column 0 = step 80: still next to the start (force ramp)
    assert np.all(np.diff(Pr[0, :498]) > -1e-3)                        # agent 0 advances toward the goal (10, 0)
    assert np.abs(Pr[:, 498] - Pr[:, 11]).max() < 0.02                 # quirk: the last two columns are undecimated steps 500, 501
    Pr2, _ = O.generate_reference(Pstart, obs, loop=20000)
    assert np.array_equal(Pr, Pr2)
    # no obstacles, one agent far from the others: pure attraction, speed settles at b3 alpha / (1 - a33) per step / a13-scaled
    Pr3, Prd3 = O.generate_reference([0, 0, 0, -50, 0, 50, 0, 100], np.array([[1e3], [1e3]]), loop=20000)
    v_inf = 0.000128132654983983 * 150.0 / (1 - 0.929527039758809)
    assert abs(Prd3[0, 300] - v_inf) < 1e-3 * v_inf
