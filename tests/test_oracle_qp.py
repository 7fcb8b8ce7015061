"""QP stage: the oracle's restatement of iSWIFT against (a) the reference's own C core compiled from /root/reference
(oracle/_ref/libiswift_ref.so) and (b) the committed golden vectors that core produced."""
import os
import numpy as np
import pytest
import oracle as O
import proto_numpy as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_canned_qp_known_answer():
    """optimization/iSWIFT/include/Matrices_small.h: the reference solves it in 7 iterations."""
    g = np.load(os.path.join(GOLD, "iswift_small.npz"))
    assert int(g["iters"]) == 7 and int(g["flag"]) == 0
    x, y, z, s, it, flag = O.qp_mehrotra(g["P"], g["c"], g["A"], g["b"], g["G"], g["h"])
    assert flag == 0 and it == 7
    assert np.abs(x - g["x"]).max() < 1e-9 * max(1.0, np.abs(g["x"]).max())
    assert np.abs(z - g["z"]).max() < 1e-8 and np.abs(s - g["s"]).max() < 1e-8


def test_canned_qp_against_live_reference_core():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    g = np.load(os.path.join(GOLD, "iswift_small.npz"))
    xr, yr, zr, sr, itr, flr, _, _ = O.iswift_ref_dense(g["P"], g["c"], g["A"], g["b"], g["G"], g["h"])
    assert flr == 0 and itr == 7
    assert np.abs(xr - g["x"]).max() < 1e-9 * np.abs(g["x"]).max()   # ordering-independent to round-off


def test_mpc_qp_golden():
    g = np.load(os.path.join(GOLD, "mpc_golden.npz"))
    for i in range(len(g["nv"])):
        nv = int(g["nv"][i])
        F = P.stance_matrix(g["feet"][i], list(g["contact"][i]))
        Q, f, A, b, G, h = O.pack_dense(g["x0"][i], g["xdes"][i], F)
        x, y, z, s, it, flag = O.qp_mehrotra(Q, f, A, b, G, h)
        assert flag == 0 and it == int(g["qp_iters_ref"][i])
        ref = g["xqp_ref"][i][:nv]
        assert np.abs(x[:24] - ref[:24]).max() < 1e-10          # X, U: same iterates as the reference core
        assert abs(x[-1] - ref[-1]) < 1e-12                     # slack variable


def test_mpc_qp_against_live_reference_core():
    if O.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    from srb_cbf_nmpc_b200 import scenarios as S
    for inst in [S.cfg1_instance(4), S.cfg1_instance(2)] + S.cfg3_instances(24):
        x0, xdes, feet, contact, obs = inst
        F = P.stance_matrix(feet, contact)
        Q, f, A, b, G, h = O.pack_dense(x0, xdes, F)
        xo = O.qp_mehrotra(Q, f, A, b, G, h)
        xr = O.iswift_ref_dense(Q, f, A, b, G, h)
        assert xo[4] == xr[4] and xo[5] == xr[5]                # same iteration count and exit flag
        assert np.abs(xo[0][:24] - xr[0][:24]).max() < 1e-9
        # numpy prototype follows the same path
        xn = P.mehrotra_qp(Q, f, A, b, G, h)
        assert xn[4] == xo[4] and np.abs(xn[0] - xo[0]).max() < 1e-9


def test_lowlevel_size_qps_against_the_reference_core():
    """Dense QPs with the low-level controller's dimensions (tests/qp_cases.py): the oracle's Mehrotra restatement against
    the answers of the reference's own iSWIFT core (tests/golden/lowlevel_qp_golden.npz): same iteration counts, same x."""
    import os
    import qp_cases
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lowlevel_qp_golden.npz"))
    qps = qp_cases.lowlevel_like_qps(12)
    for k, (P_, c_, A_, b_, G_, h_) in enumerate(qps):
        x, y, z, s, it, flag = O.qp_mehrotra(P_, c_, A_, b_, G_, h_)
        assert flag == 0 and g["flag"][k] == 0
        assert it == g["iters"][k]
        assert np.abs(x - g["x"][k]).max() < 1e-8 * max(1.0, np.abs(g["x"][k]).max())
