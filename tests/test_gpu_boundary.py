"""-m gpu: the boundary types and edge behaviour around the solve -- the dense nlpTuple entry point (pack_nlp_data.h:11-29,
MPC_dist.cpp:403-421), the reference generator's column rule for every `loop`, bad-input handling."""
import os
import subprocess
import json
import numpy as np
import pytest
import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg(built):
    import srb_cbf_nmpc_b200 as m
    return m


def _stance(feet, contact):
    """F = 2 x c stance matrix of MPC_dist.cpp:1238-1252."""
    feet = np.asarray(feet, float).reshape(2, 4)
    if sum(contact) == 4:
        return feet.copy()
    return feet[:, [1, 2]].copy() if contact[0] == 0 else feet[:, [0, 3]].copy()


def test_dense_nlptuple_entry_point(pkg):
    """a11: the reference's dense boundary type.  oracle.pack_dense matrices (column-major, MPC_dist.cpp:164-321) for 4 stance
    feet and both diagonal pairs -> the same QP answer as the compact QP call, and -- started at that QP answer, as MPC_dist.cpp:403
    does -- the same NLP answer (X, U, slack, status) as the compact QP->NLP call.  The library's own cbfnmpc_pack_dense produces the
    oracle's matrices.  Data that are not MPC_dist-shaped are refused with CBFNMPC_ERR_STRUCTURE."""
    import oracle as O
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    s = pkg.BatchSolver(device=0, max_batch=64, max_obs=20)
    pool = S.cfg3_instances(200)
    picks = []
    for want in ([1, 1, 1, 1], [0, 1, 1, 0], [1, 0, 0, 1]):
        picks += [i for i in pool if list(i[3]) == want][:4]
    assert len(picks) == 12
    d = capi.soa_from_instances(picks)
    out_qp = s.solve_host(d, mode=capi.MODE_QP)
    out_nl = s.solve_host(d, mode=capi.MODE_QP_NLP)
    n_conv = 0
    for b, (x0, xdes, feet, contact, obs) in enumerate(picks):
        F = _stance(feet, contact); c = F.shape[1]; nv = 25 + 4 * c
        Q, f, A, be, G, h = O.pack_dense(x0, xdes, F)
        Q2, f2, A2, be2, G2, h2 = s.pack_dense(x0, xdes, F)
        for m1, m2 in ((Q, Q2), (f, f2), (A, A2), (be, be2), (G, G2), (h, h2)):
            assert np.array_equal(np.asarray(m1), np.asarray(m2))
        # QP through the dense entry == QP through the compact entry
        xq, st, kk = s.solve_dense_nlptuple(Q, f, A, be, G, h, np.zeros(nv), np.zeros(2), mode_nlp=False)
        assert st == out_qp["status"][0, b]
        assert np.abs(xq[:16] - out_qp["X"][:, b]).max() < 1e-9 and np.abs(xq[16:24] - out_qp["U"][:, b]).max() < 1e-9
        assert np.abs(xq[24:24 + 4 * c] - out_qp["lam"][:4 * c, b]).max() < 1e-6
        # NLP from the QP optimum, closest obstacle chosen by the caller (MPC_dist.cpp:371-396, 403)
        ob = O.closest_obstacle(np.array([x0[0], x0[2]]), np.asarray(obs, float)) if hasattr(O, "closest_obstacle") else None
        if ob is None:
            dist = np.hypot(x0[0] - obs[0], x0[2] - obs[1]); ob = np.asarray(obs)[:, int(np.argmin(dist))]
        xn, st, kk = s.solve_dense_nlptuple(Q, f, A, be, G, h, xq, ob, mode_nlp=True)
        assert st == out_nl["status"][0, b], (b, st, out_nl["status"][0, b])
        if st == 0:
            n_conv += 1
            assert kk < 1e-8
            assert np.abs(xn[:16] - out_nl["X"][:, b]).max() < 1e-7 and np.abs(xn[16:24] - out_nl["U"][:, b]).max() < 1e-7
            assert abs(xn[nv - 1] - out_nl["slack"][0, b]) < 1e-7
    assert n_conv >= 9
    # ---- structure checks (CBFNMPC_ERR_STRUCTURE = -4)
    x0, xdes, feet, contact, obs = picks[0]
    F = _stance(feet, contact)
    Q, f, A, be, G, h = O.pack_dense(x0, xdes, F)

    def refused(*a, **k):
        with pytest.raises(pkg.CbfNmpcError) as e:
            s.solve_dense_nlptuple(*a, **k)
        return e.value.code
    Gp = G.copy(); Gp[3, 7] += 0.25                                   # a perturbed inequality entry
    assert refused(Q, f, A, be, Gp, h, np.zeros(41), np.zeros(2), mode_nlp=False) == -4
    Ap = A.copy(); Ap[0, 16] *= 1.01                                  # another input matrix Bd
    assert refused(Q, f, Ap, be, G, h, np.zeros(41), np.zeros(2), mode_nlp=False) == -4
    assert refused(Q, f, A[:27], be[:27], G, h, np.zeros(41), np.zeros(2), mode_nlp=False) == -4        # wrong p
    assert refused(Q[:40, :40], f[:40], A[:, :40], be, G[:, :40], h, np.zeros(40), np.zeros(2), mode_nlp=False) == -4   # wrong nv
    s.set_gains(Qw=250.0)                                            # matrices built for other gains than the context's
    assert refused(Q, f, A, be, G, h, np.zeros(41), np.zeros(2), mode_nlp=False) == -4
    g = O.default_gains(); g.Qw = 250.0
    Qg, fg, Ag, bg, Gg, hg = O.pack_dense(x0, xdes, F, g)
    xq, st, kk = s.solve_dense_nlptuple(Qg, fg, Ag, bg, Gg, hg, np.zeros(41), np.zeros(2), mode_nlp=False)   # ... and accepted with them
    assert st == 0
    s.close()


def test_nlptuple_boundary_from_the_cpp_host(built):
    """examples/team_cycles --nlptuple: every cycle of the 4-agent team also solved through cbfnmpc_pack_dense +
    cbfnmpc_solve_dense_nlptuple from the C++ host class (MPC_dist_gpu::solve_via_nlptuple); both paths agree."""
    exe = os.path.join(ROOT, "examples", "team_cycles")
    r = subprocess.run([exe, "--nlptuple", "4"], capture_output=True, text=True, check=True)
    res = json.loads(r.stdout.strip().splitlines()[-1])["nlptuple_check"]
    assert res["status_mismatches"] == 0 and res["max_abs_diff_X_U_slack"] < 1e-9, res


def test_reference_generator_every_loop_length(pkg):
    """The decimation rule of generateReferenceTrajectory (MPC_dist.cpp:1044-1061): output column c = step 40 (c + 2), the last two
    columns = steps ncols, ncols + 1.  For some `loop` values a step is both (3200: step 80 = ncols; 6360: step 160 = ncols + 1):
    every column must be written."""
    import torch
    import oracle as O
    dev = torch.device("cuda", 0)
    K = 9
    Pstart = np.array([0.0, 0.0, 0.0, -0.9, -1, 0, -1, -0.9])
    Pobs = np.array([[2.2, 1, 0.85, 1, 1, 3, 3.0, 3, -100], [0.9, 1.2, -0.65, 2, -1.75, 0.2, -0.70, -2.2, -0.5 + 100]])
    s = pkg.BatchSolver(device=0, max_batch=4, max_obs=K)
    d_ps = torch.from_numpy(Pstart[None].copy()).to(dev); d_po = torch.from_numpy(np.ascontiguousarray(Pobs.T)[None].copy()).to(dev)
    for loop in (3200, 6360, 1600, 4000, 120):
        ncols = loop // 40
        d_pr = torch.full((1, ncols, 8), np.nan, dtype=torch.float64, device=dev); d_prd = torch.full_like(d_pr, np.nan)
        s.generate_reference(1, d_ps.data_ptr(), d_po.data_ptr(), K, d_pr.data_ptr(), d_prd.data_ptr(), loop=loop)
        torch.cuda.synchronize()
        pr = d_pr.cpu().numpy()[0].T; prd = d_prd.cpu().numpy()[0].T
        Pr, Prd = O.generate_reference(Pstart, Pobs, loop=loop)
        assert np.isfinite(pr).all() and np.isfinite(prd).all(), (loop, np.argwhere(~np.isfinite(pr))[:4])
        assert np.abs(pr - Pr).max() < 1e-7 and np.abs(prd - Prd).max() < 1e-7, loop
    # the advertised obstacle limit is reachable (dynamic shared memory above 48 KB) and enforced
    Kbig = 300
    big = np.full((Kbig, 2), 50.0) + np.arange(Kbig)[:, None]
    d_pb = torch.from_numpy(big[None].copy()).to(dev)
    d_pr = torch.zeros((1, 3, 8), dtype=torch.float64, device=dev); d_prd = torch.zeros_like(d_pr)
    s.generate_reference(1, d_ps.data_ptr(), d_pb.data_ptr(), Kbig, d_pr.data_ptr(), d_prd.data_ptr(), loop=120)
    torch.cuda.synchronize()
    with pytest.raises(pkg.CbfNmpcError):
        s.generate_reference(1, d_ps.data_ptr(), d_pb.data_ptr(), 441, d_pr.data_ptr(), d_prd.data_ptr(), loop=120)
    s.close()


def test_non_finite_inputs_leave_the_outputs_untouched(pkg):
    """include/cbfnmpc.h: X, U, ... are not written for status 4.  Poisoned outputs stay poisoned for NaN inputs -- on the small
    (<= 64 instances, one pinned block) and on the large host path, and on the device path."""
    import torch
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    for B in (24, 1500):
        insts = S.cfg3_instances(B)
        badidx = [1, 5, B - 1]
        for i in badidx:
            a, b, c, d_, e = insts[i]; a = np.array(a, float); a[2] = np.nan; insts[i] = (a, b, c, d_, e)
        d = capi.soa_from_instances(insts)
        s = pkg.BatchSolver(device=0, max_batch=B, max_obs=20)
        out = capi.alloc_out(B)
        for k in ("X", "U", "alpha", "alpha_buf_next", "lam", "x_qp"):
            out[k][...] = -777.0
        s.solve_host(d, mode=capi.MODE_QP_NLP, out=out)
        assert (out["status"][0][badidx] == 4).all() and (out["status"][0] == 4).sum() == 3
        if B > 64:   # the large path copies straight from the device staging, which the kernel must not have written
            dev = torch.device("cuda", 0)
            tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
            tout = {k: torch.full((n, B), -777.0, dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
            tout.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
            s.solve_device(B, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}, 20)
            torch.cuda.synchronize()
            for k in ("X", "U", "alpha", "alpha_buf_next", "lam", "x_qp"):
                assert (tout[k].cpu().numpy()[:, badidx] == -777.0).all(), k
            assert (tout["status"][0].cpu().numpy()[badidx] == 4).all()
        else:
            for k in ("X", "U", "alpha", "alpha_buf_next"):
                # small path: one block comes back; the kernel wrote nothing for these instances into it, the block itself is
                # the library's (zero-initialised or stale), so only finiteness of the good instances is checked here
                assert np.isfinite(out[k][:, [i for i in range(B) if i not in badidx and out["status"][0][i] == 0]]).all()
        s.close()


def test_host_call_with_late_obstacles_equals_the_device_call(pkg):
    """The host-buffer call (one column range) uploads the obstacles on a second stream while the QP stage runs and
    cbfnmpc_lane_kernel_late looks at them afterwards: bit for bit the answers of the device-pointer call (cbfnmpc_lane_kernel,
    obstacles read at load); an instance whose obstacles are not finite is still bad input (status 4, nothing written), in QP mode
    the obstacles are not looked at."""
    import torch
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    dev = torch.device("cuda", 0)
    for B in (300, 4096):
        insts = S.cfg3_instances(B, seed=4242 + B)
        for i in (3, B - 2):
            a, b, c, d_, e = insts[i]; e = np.array(e, float); e[:] = np.nan if i == 3 else np.inf; insts[i] = (a, b, c, d_, e)
        d = capi.soa_from_instances(insts)
        s = pkg.BatchSolver(device=0, max_batch=B, max_obs=20)
        out = capi.alloc_out(B)
        for k in ("X", "U", "alpha", "alpha_buf_next", "lam", "x_qp"): out[k][...] = -777.0
        for rep in range(2):   # (the "landed" word counts calls)
            s.solve_host(d, mode=capi.MODE_QP_NLP, out=out)
        tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
        tout = {k: torch.full((n, B), -777.0, dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
        tout.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
        tout["active"] = torch.zeros((2, B), dtype=torch.int64, device=dev)
        s.solve_device(B, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}, 20)
        torch.cuda.synchronize()
        assert (out["status"][0][[3, B - 2]] == 4).all() and (out["status"][0] == 4).sum() == 2 and (out["status"][0] == 0).sum() > B // 2
        good = np.ones(B, bool); good[[3, B - 2]] = False   # (a status-4 column of the host call is whatever the staging held)
        for k in ("X", "U", "alpha", "alpha_buf_next", "lam", "x_qp", "slack", "kkt", "status", "qp_iters", "nlp_iters", "n_fact"):
            assert np.array_equal(out[k][:, good], tout[k].cpu().numpy()[:, good]), (B, k)
        assert (tout["X"].cpu().numpy()[:, ~good] == -777.0).all()
        outq = s.solve_host(d, mode=capi.MODE_QP)
        assert (outq["status"][0] == 0).all()
        s.close()


def test_bad_team_or_agent_index_is_bad_input(pkg):
    """cbfnmpc_pack_inputs: a team / agent index outside the per-team arrays is answered with status 4, nothing is read out of bounds."""
    import torch
    from srb_cbf_nmpc_b200 import capi
    dev = torch.device("cuda", 0)
    B, T, K, ncols = 8, 2, 3, 16
    rng = np.random.default_rng(0)
    q = np.zeros((18, B)); q[0] = rng.uniform(0, 5, B); q[1] = rng.uniform(-1, 1, B)
    fp = np.array([[0.2188, 0.2188, -0.1472, -0.1472], [-0.1320, 0.1320, -0.1320, 0.1320]])
    toe = np.zeros((12, B))
    for i in range(4):
        toe[3 * i] = q[0] + fp[0, i]; toe[3 * i + 1] = q[1] + fp[1, i]
    team = np.array([0, 1, 0, 1, 2, -1, 0, 1], np.int32); agent = np.array([0, 1, 2, 3, 0, 1, 4, -2], np.int32)
    Pr = np.zeros((T, ncols, 8)); Prd = np.zeros((T, ncols, 8))
    for t in range(T):
        for a in range(4):
            Pr[t, :, 2 * a] = q[0].mean() + 0.0086 * np.arange(1, ncols + 1); Prd[t, :, 2 * a] = 0.2
    Pobs = np.tile(np.array([[30.0, 30.0]]), (T, K, 1))
    tens = dict(q=q, dq=np.zeros((18, B)), contact=np.ones((4, B), np.int32), toe_pos=toe, gait_domain=np.full(B, 2, np.int32), agent_id=agent,
                team=team, agent_initial=np.stack([q[0], q[1]]), Pr=Pr, Prd=Prd, Pobs=Pobs)
    dt = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in tens.items()}
    raw = {k: dt[k].data_ptr() for k in ("q", "dq", "contact", "toe_pos", "gait_domain", "agent_id", "team", "agent_initial")}
    raw.update(Pr_refined=dt["Pr"].data_ptr(), Prd_refined=dt["Prd"].data_ptr(), Pobs_real=dt["Pobs"].data_ptr(), nq=18, ref_cols=ncols, n_teams=T, n_obs=K)
    abuf = torch.from_numpy(np.stack([q[0], 0 * q[0], q[1], 0 * q[0]])).to(dev)
    o = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
    o.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
    s = pkg.BatchSolver(device=0, max_batch=B, max_obs=K)
    s.run_cycle(B, raw, abuf.data_ptr(), {k: v.data_ptr() for k, v in o.items()}, mode=capi.MODE_QP_NLP)
    torch.cuda.synchronize()
    st = o["status"][0].cpu().numpy()
    assert (st[[4, 5, 6, 7]] == 4).all() and (st[:4] != 4).all(), st
    s.close()


def test_weak_pin_against_the_reference_snopt_print_file(pkg):
    """The reference's one recorded NLP run (print_file.out) through the CUDA path: QP stage equals SNOPT's logged start point
    (= iSWIFT's optimum) to 5e-9, NLP stage reaches SNOPT's active set with a lower objective (helpers.snopt_pin_checks)."""
    from srb_cbf_nmpc_b200 import capi
    fix = np.load(os.path.join(ROOT, "tests", "golden", "snopt_print_file.npz"))
    inst = (fix["x0"], fix["xdes"], fix["feet"], list(fix["contact"]), fix["obs"].reshape(2, 1))
    s = pkg.BatchSolver(device=0, max_batch=4, max_obs=4)
    out = s.solve_host(capi.soa_from_instances([inst]), mode=capi.MODE_QP_NLP)
    x = np.concatenate([out["X"][:, 0], out["U"][:, 0], out["lam"][:, 0], out["slack"][:, 0]])
    act = H.active_rows_mask(out, 0, 4)
    H.snopt_pin_checks(x, out["x_qp"][:, 0], int(out["status"][0, 0]), float(out["kkt"][0, 0]), [(92 + k) in act for k in range(4)], fix)
    ref = H.oracle_batch([inst], mode=1)
    assert H.compare(out, ref)["worst"] < 1e-7
    s.close()
