"""CPU-side check of the *kernel source*: srb-cbf-nmpc_b200/csrc/cbfnmpc_lane.h compiled for the host, each 4-lane group on 4
cooperative fibers (tests/emu).  This is a logic test of the product kernel, not a fallback: the C ABI never reaches the
emulator."""
import numpy as np
import helpers as H
import emu
from srb_cbf_nmpc_b200 import scenarios as S


def test_emulated_kernel_qp_mode():
    insts = H.pad_obs([S.cfg1_instance(4), S.cfg1_instance(2)]) + S.cfg3_instances(10)
    out = emu.solve(insts, mode=0)
    ref = H.oracle_batch(insts, mode=0)
    assert (out["qp_iters"][0] == ref["qp_iters"]).all()
    assert np.abs(out["X"].T - ref["X"]).max() < 1e-9 and np.abs(out["U"].T - ref["U"]).max() < 1e-9


def test_emulated_kernel_nlp_mode():
    insts = H.pad_obs([S.cfg1_instance(4), S.cfg1_instance(2)]) + S.cfg3_instances(22)
    out = emu.solve(insts, mode=1)
    ref = H.oracle_batch(insts, mode=1)
    c = H.compare(out, ref)
    assert c["status_equal"].all()
    assert c["both"].sum() >= 20 and c["worst"] < 1e-8
    assert out["kkt"][0][c["both"]].max() < 1e-8
    for i in np.nonzero(c["both"])[0]:
        cc = int(ref["c"][i])
        assert H.active_rows_oracle(ref, i, cc) == H.active_rows_mask(out, i, cc)
    assert np.abs(out["alpha"].T - ref["alpha"])[c["both"]].max() < 1e-8


def test_bad_contact_pattern_is_flagged():
    inst = list(S.cfg1_instance(4)); inst[3] = [1, 1, 1, 0]
    out = emu.solve(H.pad_obs([tuple(inst)]), mode=1)
    assert out["status"][0, 0] == 4


def test_group_lanes_meet_at_the_same_shuffle_sites():
    """-O0 build: one call address per source-level shuffle; the harness fails if the 4 lanes of a group ever wait at
    different sites (which __shfl_sync with the group mask would not tolerate on the device)."""
    insts = H.pad_obs([S.cfg1_instance(4), S.cfg1_instance(2)]) + S.cfg3_instances(14)
    for mode in (0, 1):
        out = emu.solve(insts, mode=mode, sitecheck=True)
        ref = H.oracle_batch(insts, mode=mode)
        assert (out["status"][0] == ref["status"]).all()


def test_degenerate_support_weights_qp():
    """Instances whose lambdas end strictly inside (0, 1) at every stage: the reduced Hessian loses one pivot to rounding
    (cbfnmpc_lane.h, l_chol pivot rule); the QP stage must still reach the iSWIFT answer."""
    allinst = S.cfg3_instances(2400)
    insts = [allinst[i] for i in (159, 1064, 1356, 1631, 2229, 2308)]
    out = emu.solve(insts, mode=0)
    ref = H.oracle_batch(insts, mode=0)
    assert (out["status"][0] == 0).all() and (ref["status"] == 0).all()
    assert np.abs(out["X"].T - ref["X"]).max() < 1e-8 and np.abs(out["U"].T - ref["U"]).max() < 1e-8


def test_non_finite_and_absurd_inputs_come_back_flagged():
    """NaN / Inf / absurd magnitudes in the inputs: every solve terminates; non-finite data are status 4 in the oracle and in
    the kernel source, finite nonsense is 'not converged' in both (the reference would hand them to its solvers unchecked)."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("scen", os.path.join(os.path.dirname(__file__), "..", "srb-cbf-nmpc_b200", "scenarios.py"))
    S = importlib.util.module_from_spec(spec); spec.loader.exec_module(S)
    cases = H.degenerate_instances(S.cfg3_instances(6))
    insts = [c[2] for c in cases]
    for mode in (0, 1):
        out = emu.solve(insts, mode=mode)
        ref = H.oracle_batch(insts, mode=mode)
        for k, (label, bad, _) in enumerate(cases):
            if bad and (mode == 1 or not label.startswith("obs")): assert out["status"][0][k] == 4 and ref["status"][k] == 4, (label, mode, out["status"][0][k], ref["status"][k])
            elif mode == 1: assert out["status"][0][k] != 0 and ref["status"][k] != 0, (label, out["status"][0][k], ref["status"][k])


def test_wide_fuzz_against_the_oracle():
    """The kernel source against the oracle far outside the cfg-3 distribution (helpers.wide_fuzz_instances): same statuses and the
    same answers in QP -> NLP mode; in QP mode the same answers, statuses and iteration counts (the rare 4-foot instance whose last
    support weight sits on its bound, which stalled in round 1, is what LaneSolver::repivot is for)."""
    insts = H.wide_fuzz_instances(300)
    out = emu.solve(insts, mode=1); ref = H.oracle_batch(insts, mode=1)
    c = H.compare(out, ref)
    assert c["status_equal"].mean() >= 0.995 and c["worst"] < 1e-6
    assert (out["status"][0] == 0).sum() > 150 and (out["status"][0] == 2).sum() > 20      # both outcomes are exercised
    outq = emu.solve(insts, mode=0); refq = H.oracle_batch(insts, mode=0)
    both = (outq["status"][0] == 0) & (refq["status"] == 0)
    assert (outq["status"][0] == refq["status"]).all() and both.sum() > 250
    assert np.abs(outq["X"].T - refq["X"])[both].max() < 1e-8
    assert (outq["qp_iters"][0][both] == refq["qp_iters"][both]).all()


def test_qp_mode_against_the_reference_iswift_core():
    """QP mode (use_snopt = false, the reference's default) of the kernel source against the REFERENCE's own iSWIFT core
    (oracle/_ref, compiled from /root/reference/optimization/iSWIFT): every status equal on a wide fuzz, iteration counts equal on
    all but the odd instance where the core's LDL' dynamic regularisation (ldl.c:273-275) costs it extra iterations that the dense
    restatement does not need either."""
    import oracle as O
    if O.ref_lib() is None:
        import pytest
        pytest.skip("oracle/_ref not built (the reference tree is absent)")
    insts = H.wide_fuzz_instances(600)
    outq = emu.solve(insts, mode=0)
    O.use_reference_qp(True)
    try:
        refq = H.oracle_batch(insts, mode=0)
    finally:
        O.use_reference_qp(False)
    st, rs = outq["status"][0], refq["status"]
    assert (st == rs).all(), np.nonzero(st != rs)[0]
    both = (st == 0) & (rs == 0)
    assert both.sum() > 500
    assert (outq["qp_iters"][0][both] == refq["qp_iters"][both]).mean() > 0.995
    assert np.abs(outq["X"].T - refq["X"])[both].max() < 1e-6


def test_support_weights_come_back_in_the_order_of_the_feet():
    """LaneSolver::repivot relabels the feet of a lane while it solves; the outputs must not show it: u = F lambda and 1'lambda = 1
    hold with the reported lambda in the caller's foot order, and where lambda is unique (at least two feet of a stage unloaded --
    which is also when the last foot's block becomes stiff and the relabelling happens) it equals the oracle's."""
    insts = [i for i in H.wide_fuzz_instances(500) if sum(i[3]) == 4]
    for mode in (0, 1):
        out = emu.solve(insts, mode=mode); ref = H.oracle_batch(insts, mode=mode)
        ok = (out["status"][0] == 0) & (ref["status"] == 0)
        lam = out["lam"].T[ok].reshape(-1, 4, 4); rl = ref["lam"][ok].reshape(-1, 4, 4)      # [instance, stage, foot]
        U = out["U"].T[ok].reshape(-1, 4, 2)
        F = np.array([np.asarray(insts[i][2]) for i in np.nonzero(ok)[0]])                    # [instance, 2, 4]
        assert np.abs(U - np.einsum("nkf,nsf->nsk", F, lam)).max() < 1e-12 and np.abs(1 - lam.sum(2)).max() < 1e-12
        uniq = (rl < 1e-7).sum(2) >= 2
        assert uniq.sum() > 100 and (rl[:, :, 3] < 1e-7).sum() > 20          # the last foot is unloaded often enough
        assert np.abs(lam - rl)[uniq].max() < 1e-9


def test_late_obstacle_instantiation_gives_the_same_answers():
    """lane_run_group<C, true> (cbfnmpc_lane_kernel_late: the host-buffer call uploads the obstacles while the QP stage runs, the kernel
    looks at them afterwards) against lane_run_group<C, false>: bit for bit the same answers; non-finite obstacles are still bad input
    (status 4), in QP mode they are not looked at."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("scen", os.path.join(os.path.dirname(__file__), "..", "srb-cbf-nmpc_b200", "scenarios.py"))
    S = importlib.util.module_from_spec(spec); spec.loader.exec_module(S)
    insts = S.cfg3_instances(40, seed=31)
    a, b, c, d, e = insts[5]; e = np.array(e, float); e[:] = np.nan; insts[5] = (a, b, c, d, e)
    ref = emu.solve(insts, mode=1); out = emu.solve(insts, mode=1, late=True)
    assert out["status"][0][5] == 4 and ref["status"][0][5] == 4 and (out["status"][0] == 0).sum() > 25
    for k in ("X", "U", "slack", "status", "qp_iters", "nlp_iters", "kkt", "alpha"):
        assert np.array_equal(out[k], ref[k]), k
    outq = emu.solve(insts, mode=0, late=True)
    assert (outq["status"][0] == 0).all()

