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
