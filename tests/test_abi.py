"""The C-ABI library loads on a CPU-only machine and exports every symbol include/cbfnmpc.h declares; no compute calls."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(built):
    import srb_cbf_nmpc_b200 as m
    hdr = open(os.path.join(ROOT, "include", "cbfnmpc.h")).read()
    declared = set(re.findall(r"\b(cbfnmpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(m.EXPORTS), declared ^ set(m.EXPORTS)
    L = m.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.cbfnmpc_version() >= 100


def test_struct_layouts_match(built):
    import srb_cbf_nmpc_b200 as m
    g = m.default_gains(); o = m.default_options()
    assert g.Qw == 300.0 and g.Pw == 2000.0 and g.slack_w == 3000.0 and g.box == 1000.0
    assert o.qp_maxit == 25 and o.qp_reltol == 1e-6 and o.max_iter == 200 and o.max_refine == 2 and o.sigma_cap == 1e8


def test_no_cpu_fallback(built):
    import torch
    import srb_cbf_nmpc_b200 as m
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(m.CbfNmpcError, match="no CUDA device"):
        m.BatchSolver()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "srb-cbf-nmpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "api.py" or "import oracle" not in txt
                assert "import oracle" not in txt and "liborc" not in txt and "emu_warp" not in txt


def test_mpc_dist_host_logic_without_gpu():
    """copPlanner_eventbase column selection (MPC_dist.cpp:716-719) and the nominal-stance rule (:1211-1216)."""
    from srb_cbf_nmpc_b200.mpc_dist import MPC_dist, INIT_FOOTPRINT
    m = MPC_dist.__new__(MPC_dist)
    MPC_dist.__init__(m, solver=object())
    m.setAgentID(1)
    m.setPstart([0, 0, 0, -0.9, -1, 0, -1, -0.9])
    assert np.allclose(m.mpc_state_alpha_buffer_, [0, 0, -0.9, 0])
    Pr = np.arange(8 * 2500, dtype=float).reshape(8, 2500); Prd = -Pr
    m.setReferenceTrajectory(Pr, Prd)
    m.gaitDomain_ = 3
    xd = m.copPlanner_eventbase()
    assert xd.shape == (16,)
    assert np.allclose(xd[0::4], Pr[2, 12:16]) and np.allclose(xd[1::4], Prd[2, 12:16]) and np.allclose(xd[2::4], Pr[3, 12:16])
    m.gaitDomain_ = 1
    m.updateState(np.zeros(18), np.zeros(18), [1, 1, 1, 1], np.ones((3, 4)), np.zeros(4))
    feet = m.footholdsPlanner()
    assert np.allclose(feet[0], INIT_FOOTPRINT[:, 0] + 0.0) and np.allclose(feet[1], INIT_FOOTPRINT[:, 1] - 0.9)
    m.gaitDomain_ = 2
    m.updateState(np.zeros(18), np.zeros(18), [1, 0, 0, 1], np.full((3, 4), 7.0), np.zeros(4))
    assert np.allclose(m.footholdsPlanner(), 7.0)
    assert np.allclose(m.get_lastState(), 0) and m.getDomain() == 0 and m.use_snopt is False
