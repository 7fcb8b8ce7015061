"""-m gpu: the EXTENDED mode (BASELINE configs 4 and 5) through the C ABI against its CPU checker
oracle/cbfnmpc_oracle_ext.c.  No reference implementation of this mode exists; results are labelled accordingly."""
import numpy as np
import pytest
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(built):
    import srb_cbf_nmpc_b200 as m
    return m


@pytest.fixture(scope="module")
def solver(pkg):
    return pkg.BatchSolver(device=0, max_batch=64, max_obs=20)


def test_ext_reduces_to_the_faithful_problem(pkg, solver):
    """N = 4, one obstacle row (the closest obstacle, chosen by the caller), no neighbours: the answers of the reference-faithful
    path -- same statuses, same iteration counts, X / U / slack to 1e-7 -- although the linear algebra differs (fifth state)."""
    import oracle as O
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    insts = S.cfg3_instances(1024)
    ref = H.oracle_batch(insts, mode=1, nthreads=16)
    ext = [(x0, xdes, np.asarray(feet)[None], contact, ref["obs"][b].reshape(2, 1)) for b, (x0, xdes, feet, contact, obs) in enumerate(insts)]
    out = solver.ext_solve_host(capi.ext_soa_from_instances(ext, 4), N=4, team_size=1)
    st = out["status"][0]
    assert (st == ref["status"]).mean() > 0.998, np.nonzero(st != ref["status"])[0][:10]
    both = (st == 0) & (ref["status"] == 0)
    assert both.sum() > 950
    assert np.abs(out["X"].T - ref["X"])[both].max() < 1e-7 and np.abs(out["U"].T - ref["U"])[both].max() < 1e-7
    assert np.abs(out["slack"][0] - ref["slack"])[both].max() < 1e-7
    assert (out["qp_iters"][0] == ref["qp_iters"]).mean() > 0.999
    assert (out["nlp_iters"][0] == ref["nlp_iters"])[both].mean() > 0.97
    assert out["kkt"][0][both].max() < 1e-8
    assert np.abs(out["alpha"].T - ref["alpha"])[both].max() < 1e-6       # Bezier epilogue of the first domain
    # the same instances through the faithful CUDA path
    f = solver.__class__(device=0, max_batch=1024, max_obs=20)
    fo = f.solve_host(capi.soa_from_instances(insts), mode=capi.MODE_QP_NLP)
    bb = both & (fo["status"][0] == 0)
    assert np.abs(out["X"][:16] - fo["X"])[:, bb].max() < 1e-7
    f.close()


@pytest.mark.parametrize("N,K,B", [(4, 16, 192), (8, 16, 96), (8, 1, 96)])
def test_ext_single_agents_against_the_checker(pkg, solver, N, K, B):
    """K obstacle rows per grid point, N = 4 / 8 grid points, no neighbours."""
    import oracle as O
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    insts = S.cfg5_instances(B, N=N, K_obs=K, team=False)
    res, _ = O.ext_solve_teams(*H.ext_arrays(insts), T=1, N=N, nthreads=16)
    out = solver.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=1)
    c = H.ext_compare(out, res)
    assert c["status_equal"].mean() >= 0.98, (np.bincount(out["status"][0], minlength=5), np.bincount([r["status"] for r in res], minlength=5))
    assert c["compared"].sum() >= 0.6 * B and c["worst"] < 1e-6, c["worst"]
    assert c["iters_equal"][c["compared"]].mean() > 0.9
    assert out["kkt"][0][c["compared"]].max() < 1e-8


@pytest.mark.parametrize("N,K,n_teams", [(4, 1, 96), (8, 16, 24)])
def test_ext_teams_in_one_warp_against_the_checker(pkg, solver, N, K, n_teams):
    """BASELINE config 4 (N = 4, 1 obstacle + 3 neighbour rows) and config 5 (N = 8, 16 obstacle + 3 neighbour rows): four-robot
    teams, under-relaxed Jacobi over the iterates, exchange by warp shuffles.  Teams in which every member converges reproduce the
    checker's answers and its number of team iterations."""
    import oracle as O
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    insts = S.cfg4_teams(n_teams, N=N, K_obs=K)
    res, ti = O.ext_solve_teams(*H.ext_arrays(insts), T=4, N=N, eps_nbr=S.EPS_NBR, nthreads=16)
    out = solver.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=4, eps_nbr=S.EPS_NBR)
    c = H.ext_compare(out, res, T=4)
    assert c["status_equal"].mean() >= 0.97
    assert c["compared"].sum() >= 0.45 * 4 * n_teams and c["worst"] < 1e-6, (c["compared"].sum(), c["worst"])
    okt = c["compared"].reshape(-1, 4).all(1)
    # same number of team iterations on most teams; on the others the hold / resume decisions of the last iterations (KKT error at
    # the tolerance, neighbours still settling) fall differently with the GPU's reciprocals -- same answers, a few iterations apart
    tg, to = out["team_iters"][0][::4][okt].astype(float), ti[okt].astype(float)
    assert (tg == to).mean() > 0.6 and 0.8 < tg.sum() / to.sum() < 1.25, (np.bincount(np.abs(tg - to).astype(int)), tg.sum(), to.sum())
    assert out["kkt"][0][c["compared"]].max() < 1e-8


def test_ext_fixed_neighbour_trajectories(pkg, solver):
    """Neighbour rows fed by FIXED trajectories (cbfnmpc_ext_in.nbr_traj): each member's best response to the team's final
    trajectories is its own part of the team solution (the fixed-point property of the Jacobi iteration), and equals the checker."""
    import oracle as O
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    N, n_teams = 4, 24
    insts = S.cfg4_teams(n_teams, N=N, K_obs=1)
    team = solver.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=4, eps_nbr=S.EPS_NBR)
    B = 4 * n_teams
    nbr = np.zeros((3, 2 * N, B))
    for b in range(B):
        t0 = 4 * (b // 4)
        others = [t0 + k for k in range(4) if t0 + k != b]
        for i, o in enumerate(others):
            nbr[i, 0::2, b] = team["X"][0::4, o]; nbr[i, 1::2, b] = team["X"][2::4, o]
    out = solver.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=1, nbr_traj=nbr, eps_nbr=S.EPS_NBR)
    okt = np.repeat((team["status"][0] == 0).reshape(-1, 4).all(1), 4)
    ok = okt & (out["status"][0] == 0)
    assert ok.sum() >= 0.7 * B
    assert np.abs(out["X"] - team["X"])[:, ok].max() < 1e-6 and np.abs(out["slack"] - team["slack"])[:, ok].max() < 1e-6
    for b in np.nonzero(ok)[0][:12]:
        x0, xdes, feet, contact, obs = insts[b]
        nb = np.stack([nbr[:, 0::2, b], nbr[:, 1::2, b]], 2)
        r = O.ext_solve(x0, xdes, feet, contact, obs, nbr=nb, N=N, eps_nbr=S.EPS_NBR)
        assert r["status"] == 0 and np.abs(r["X"] - out["X"][:, b]).max() < 1e-6


def test_ext_sharded_path_on_one_gpu_equals_the_warp_path(pkg):
    """cbfnmpc_ext_solve_sharded without a communicator: one launch per iteration, the members of a team in different member
    blocks, published trajectories exchanged through global memory (where several GPUs all-gather them).  Same answers and the same
    number of team iterations as the team-in-one-warp path."""
    import torch
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    dev = torch.device("cuda", 0)
    for N, K, n_teams in ((4, 1, 64), (8, 16, 16)):
        insts = S.cfg4_teams(n_teams, N=N, K_obs=K)
        B = 4 * n_teams
        s = pkg.BatchSolver(device=0, max_batch=B, max_obs=20)
        team = s.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=4, eps_nbr=S.EPS_NBR)
        mm = [insts[4 * t + m] for m in range(4) for t in range(n_teams)]           # member-major
        d = capi.ext_soa_from_instances(mm, N)
        tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
        tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items() if k != "pub"}
        tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
        stats = s.ext_solve_sharded(n_teams, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}, d["n_obs"],
                                    N=N, team_size=4, eps_nbr=S.EPS_NBR)
        torch.cuda.synchronize()
        perm = np.array([m * n_teams + t for t in range(n_teams) for m in range(4)])   # team-major index -> member-major index
        st = tout["status"][0].cpu().numpy()[perm]
        assert (st == team["status"][0]).all()
        ok = np.repeat((st == 0).reshape(-1, 4).all(1), 4)
        X = tout["X"].cpu().numpy()[:, perm]; sl = tout["slack"][0].cpu().numpy()[perm]
        assert np.abs(X - team["X"])[:, ok].max() < 1e-9 and np.abs(sl - team["slack"][0])[ok].max() < 1e-9
        ti = tout["team_iters"][0].cpu().numpy()[perm]
        assert (ti == team["team_iters"][0])[ok].all()
        assert stats["iterations"] >= ti.max() and stats["bytes_per_exchange"] == 8 * (2 * N + 2) * n_teams * 4 and stats["nranks"] == 1
        s.close()


def test_ext_fused_path_on_one_gpu_equals_the_sharded_path(pkg):
    """cbfnmpc_ext_solve_fused without a communicator: the whole solve in ONE launch, the members of a team in different CTAs, their
    published trajectories exchanged through the mailbox and its sequence flags (on several GPUs: through each other's memory).  Bit for
    bit the answers, statuses and iteration counts of the per-iteration path; a member with a bad contact pattern does not stall its
    team mates; a batch that cannot be resident at once runs as launches over slices of the teams."""
    import torch
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    dev = torch.device("cuda", 0)
    for N, K, n_teams, bad in ((4, 1, 64, False), (8, 16, 16, False), (4, 1, 96, True), (4, 1, 2048, False)):
        insts = S.cfg4_teams(n_teams, N=N, K_obs=K)
        if bad:
            a, b_, c_, d_, e = insts[4 * 5 + 2]; insts[4 * 5 + 2] = (a, b_, c_, [1, 0, 0, 0], e)   # member 2 of team 5: one foot
        B = 4 * n_teams
        s = pkg.BatchSolver(device=0, max_batch=B, max_obs=20)
        mm = [insts[4 * t + m] for m in range(4) for t in range(n_teams)]           # member-major
        d = capi.ext_soa_from_instances(mm, N)
        tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
        res = []
        for fn in (s.ext_solve_sharded, s.ext_solve_fused, s.ext_solve_fused):
            tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items() if k != "pub"}
            tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
            stats = fn(n_teams, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}, d["n_obs"],
                       N=N, team_size=4, eps_nbr=S.EPS_NBR)
            torch.cuda.synchronize()
            res.append((stats, {k: v.cpu().numpy() for k, v in tout.items()}))
        (ss, so), (fs, fo), (fs2, fo2) = res
        assert fs["fused"] == 1 and ss["fused"] == 0
        keep = np.ones(B, bool)
        if bad:
            keep[[m * n_teams + 5 for m in range(4)]] = False      # the per-iteration path reads an unwritten block for that member
            st = fo["status"][0]
            assert st[2 * n_teams + 5] == 4 and (st[[m * n_teams + 5 for m in (0, 1, 3)]] != 4).all()
        for k in so:
            assert np.array_equal(so[k][..., keep], fo[k][..., keep]), (N, n_teams, k)
            assert np.array_equal(fo[k], fo2[k]), (N, n_teams, k, "second call")
        if fs["fused"]:
            assert fs["launches"] == (4 if n_teams == 2048 else 2) and fs["iterations"] == fo["team_iters"][0].max() + 1
        s.close()


def test_ext_edge_inputs(pkg, solver):
    """Mixed stance classes inside a team and bad contact patterns (status 4 for the whole team), non-finite data of one member
    (that member status 4, the others still solve against its measured position; a non-finite POSITION would reach them as a NaN row), argument checks."""
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    N = 4
    insts = S.cfg4_teams(8, N=N, K_obs=1)
    bad = list(insts)
    a, b_, c_, d_, e = bad[1]; bad[1] = (a, b_, c_, [1, 1, 1, 1] if sum(d_) == 2 else [0, 1, 1, 0], e)       # team 0: mixed classes
    a, b_, c_, d_, e = bad[6]; bad[6] = (a, b_, c_, [1, 0, 0, 0], e)                                          # team 1: one foot
    a, b_, c_, d_, e = bad[9]; b_ = np.array(b_, float); b_[5] = np.nan; bad[9] = (a, b_, c_, d_, e)          # team 2: NaN reference of member 1
    out = solver.ext_solve_host(capi.ext_soa_from_instances(bad, N), N=N, team_size=4, eps_nbr=S.EPS_NBR)
    st = out["status"][0]
    assert (st[0:8] == 4).all() and st[9] == 4 and (st[[8, 10, 11]] != 4).all()
    good = solver.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=4, eps_nbr=S.EPS_NBR)
    assert (st[12:] == good["status"][0][12:]).all() and np.abs(out["X"] - good["X"])[:, 12:].max() < 1e-9   # other teams unaffected
    with pytest.raises(pkg.CbfNmpcError):
        solver.ext_solve_host(capi.ext_soa_from_instances(insts[:6], N), N=N, team_size=4)                    # batch not a multiple of the team size
    with pytest.raises(pkg.CbfNmpcError):
        solver.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=5, team_size=1)
