"""The extended-mode kernel SOURCE (srb-cbf-nmpc_b200/csrc/cbfnmpc_ext.h) compiled for the host (tests/emu/emu_ext.cpp: one warp on
32 cooperative fibers, shuffles and votes as mask-based rendezvous) against the checker -- the CPU-side check of the solver logic,
including the team exchange.  Dev/test tool; never a product path."""
import numpy as np
import oracle as O
import helpers as H
import emu_ext
from srb_cbf_nmpc_b200 import scenarios as S


def test_emulated_ext_kernel_reduces_to_the_faithful_oracle():
    insts = S.cfg3_instances(24)
    refs = [O.solve_instance(*i, mode=1) for i in insts]
    ext = [(x0, xdes, np.asarray(feet)[None], contact, refs[b]["obs"].reshape(2, 1)) for b, (x0, xdes, feet, contact, obs) in enumerate(insts)]
    out = emu_ext.solve(ext, N=4, T=1)
    rs = np.array([r["status"] for r in refs])
    assert (out["status"][0] == rs).all()
    for b in np.nonzero(rs == 0)[0]:
        assert np.abs(out["X"][:, b] - refs[b]["x"][:16]).max() < 1e-8 and abs(out["slack"][0, b] - refs[b]["x"][refs[b]["nv"] - 1]) < 1e-8
        assert out["qp_iters"][0, b] == refs[b]["qp_iters"] and abs(out["nlp_iters"][0, b] - refs[b]["nlp_iters"]) <= 1


def test_emulated_ext_kernel_teams_and_long_horizon():
    for N, K, insts in ((4, 1, S.cfg4_teams(6)), (8, 16, S.cfg5_instances(8))):
        res, ti = O.ext_solve_teams(*H.ext_arrays(insts), T=4, N=N, eps_nbr=S.EPS_NBR, nthreads=8)
        out = emu_ext.solve(insts, N=N, T=4, eps_nbr=S.EPS_NBR)
        c = H.ext_compare(out, res, T=4)
        assert c["status_equal"].all() and c["compared"].sum() >= 4 and c["worst"] < 1e-7
        okt = c["compared"].reshape(-1, 4).all(1)
        assert (np.abs(out["team_iters"][0][::4][okt] - ti[okt]) <= 1).all()
