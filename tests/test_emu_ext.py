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


def test_emulated_fused_exchange_equals_the_team_in_one_warp():
    """The fused exchange of cbfnmpc_ext_solve_fused (xmode 4: members of a team in different groups, published trajectories and
    sequence flags in a mailbox, polling members) run by the host emulation of the kernel source -- four-robot teams, N = 4 (two teams =
    the eight groups of one warp) and N = 8 (one team), and a team with a one-footed member, which must not stall its team mates --
    against the same teams exchanged by shuffles (xmode 1): same statuses, same team iteration counts, same answers."""
    for N, K, n_teams in ((4, 1, 2), (8, 16, 1)):
        pool = S.cfg4_teams(12, N=N, K_obs=K)
        by_cls = {}
        for t in range(12):
            by_cls.setdefault(sum(pool[4 * t][3]), []).append(t)
        teams = max(by_cls.values(), key=len)[:n_teams]                   # teams of ONE stance class: they share the emulated warp
        insts = [pool[4 * t + m] for t in teams for m in range(4)]         # team-major
        mm = [pool[4 * t + m] for m in range(4) for t in teams]            # member-major (the fused path's layout)
        ref = emu_ext.solve(insts, N=N, T=4, eps_nbr=S.EPS_NBR)
        out = emu_ext.solve(mm, N=N, T=1, eps_nbr=S.EPS_NBR, fused_members=4)
        perm = np.array([m * n_teams + t for t in range(n_teams) for m in range(4)])   # team-major index -> member-major index
        assert (out["status"][0][perm] == ref["status"][0]).all()
        assert (out["team_iters"][0][perm] == ref["team_iters"][0]).all()
        ok = ref["status"][0] == 0
        assert ok.any()
        assert np.abs(out["X"][:, perm] - ref["X"])[:, ok].max() < 1e-9 and np.abs(out["slack"][0][perm] - ref["slack"][0])[ok].max() < 1e-9
    # a one-footed member (status 4, no lane runs it): published once and for all, its team mates finish
    pool = S.cfg4_teams(4, N=4, K_obs=1)
    mm = [pool[m] for m in range(4)]
    a, b_, c_, d_, e = mm[2]; mm[2] = (a, b_, c_, [1, 0, 0, 0], e)
    out = emu_ext.solve(mm, N=4, T=1, eps_nbr=S.EPS_NBR, fused_members=4)
    assert out["status"][0][2] == 4 and (out["status"][0][[0, 1, 3]] != 4).all()
