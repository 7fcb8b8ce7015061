"""Shared helpers for the parity tests: oracle batch call from a list of instances."""
import numpy as np
import oracle as O


def pad_obs(insts, K=20):
    out = []
    for a, b, c, d, e in insts:
        e = np.asarray(e, float)
        if e.shape[1] < K:
            e = np.hstack([e, np.repeat(e[:, -1:], K - e.shape[1], 1)])
        out.append((a, b, c, d, e))
    return out


def oracle_batch(insts, mode=1, nthreads=8, alpha_buf=None, options=None):
    x0 = np.array([i[0] for i in insts]); xd = np.array([i[1] for i in insts])
    ft = np.array([np.asarray(i[2]).reshape(2, 4) for i in insts]); ct = np.array([i[3] for i in insts], np.int32)
    ob = np.array([i[4] for i in insts])
    ab = np.stack([x0[:, 0], np.zeros(len(x0)), x0[:, 2], np.zeros(len(x0))], 1) if alpha_buf is None else np.asarray(alpha_buf, float)
    o = O.default_options()
    for k, v in (options or {}).items():
        setattr(o, k, v)
    return O.solve_batch(x0, xd, ft, ct, ob, ab, mode=mode, nthreads=nthreads, o=o)


def compare(out, ref, tol=1e-6):
    """Parity bar of BASELINE.json: X, U, slack to `tol` relative (scale = max(1, |value|_inf per instance))."""
    st = out["status"][0]
    both = (st == 0) & (ref["status"] == 0)
    sx = np.maximum(1.0, np.abs(ref["X"]).max(1)); su = np.maximum(1.0, np.abs(ref["U"]).max(1))
    ex = np.abs(out["X"].T - ref["X"]).max(1) / sx
    eu = np.abs(out["U"].T - ref["U"]).max(1) / su
    es = np.abs(out["slack"][0] - ref["slack"]) / np.maximum(1.0, np.abs(ref["slack"]))
    return dict(status_equal=(st == ref["status"]), both=both, ex=ex, eu=eu, es=es,
                worst=float(max(ex[both].max(), eu[both].max(), es[both].max())) if both.any() else 0.0)


# reference row numbers (0-based inside NeqConstraints, DV:245-324) used for active-set comparison
def active_rows_oracle(ref, i, c):
    """Set of active rows (z > s) among friction, distance and velocity rows (lambda / box rows excluded)."""
    mG = 60 + 8 * c
    rows = list(range(12)) + list(range(mG, mG + 20))
    z, s = ref["z"][i], ref["s"][i]
    return {r for r in rows if z[r] > s[r]}


def active_rows_mask(out, i, c):
    mG = 60 + 8 * c
    rows = list(range(12)) + list(range(mG, mG + 20))
    lo, hi = int(out["active"][0, i]), int(out["active"][1, i])
    m = lo | (hi << 64)
    return {r for r in rows if (m >> r) & 1}


def degenerate_instances(base):
    """Non-finite and absurd data on top of 6 regular instances: (label, must_be_status_4, instance).  Non-finite measured state,
    reference column, stance-foot coordinate or closest obstacle -> status 4 (bad input) in the oracle and in the kernel;
    finite nonsense only has to come back, flagged as not converged, in both."""
    out = []
    spec = [("x0", np.nan, True), ("x0", np.inf, True), ("xdes", np.nan, True), ("xdes", -np.inf, True), ("feet", np.nan, True),
            ("obs", np.nan, True), ("obs", np.inf, True), ("x0", 1e8, False), ("xdes", 1e12, False), ("feet0", 0.0, False), ("x0", 1e300, False)]
    for k, (what, val, bad) in enumerate(spec):
        a, b, c, d, e = base[k % len(base)]
        a = np.array(a, float); b = np.array(b, float); c = np.array(c, float); e = np.array(e, float); d = [1, 1, 1, 1] if what == "feet" else d
        if what == "x0": a[0] = val
        if what == "xdes": b.flat[5] = val
        if what == "feet": c.flat[2] = val
        if what == "feet0": c[:] = 0.0
        if what == "obs": e[:] = val
        out.append(("%s=%g" % (what, val), bad, (a, b, c, d, e)))
    return out


def wide_fuzz_instances(n, seed=2025):
    """Instances well outside the cfg-3 distribution: feet scaled 0.5-1.5 and perturbed by 8 cm, start velocities sometimes
    outside the velocity box (the NLP then has no interior), reference speeds up to 0.34 m/s in any direction with noisy
    positions, 1-20 obstacles at 0.2-6 m, sometimes one almost on top of the robot."""
    rng = np.random.default_rng(seed)
    nominal = np.array([[0.183, 0.183, -0.183, -0.183], [-0.132, 0.132, -0.132, 0.132]])
    out = []
    for _ in range(n):
        p0 = np.array([rng.uniform(-3, 12), rng.uniform(-4, 4)])
        v0 = rng.uniform(-0.349, 0.349, 2) if rng.random() < 0.7 else rng.uniform(-0.6, 0.6, 2)
        x0 = np.array([p0[0], v0[0], p0[1], v0[1]])
        t = rng.uniform(0, 2 * np.pi)
        vref = rng.uniform(0.0, 0.34) * np.array([np.cos(t), np.sin(t)])
        xdes = np.zeros(16)
        for k in range(4):
            xdes[4 * k:4 * k + 4] = [p0[0] + vref[0] * 0.043 * (k + 1) + rng.normal(0, 0.02), vref[0],
                                     p0[1] + vref[1] * 0.043 * (k + 1) + rng.normal(0, 0.02), vref[1]]
        u = rng.random()
        contact = [1, 1, 1, 1] if u < 0.4 else ([0, 1, 1, 0] if u < 0.7 else [1, 0, 0, 1])
        feet = p0[:, None] + nominal * rng.uniform(0.5, 1.5) + rng.uniform(-0.08, 0.08, (2, 4))
        K = int(rng.integers(1, 21))
        r = rng.uniform(0.2, 6.0, K); a = rng.uniform(0, 2 * np.pi, K)
        obs = p0[:, None] + np.stack([r * np.cos(a), r * np.sin(a)])
        if rng.random() < 0.3:
            obs[:, 0] = p0 + rng.uniform(-0.3, 0.3, 2)
        obs = np.hstack([obs, np.repeat(obs[:, -1:], 20 - K, 1)])
        out.append((x0, xdes, feet, contact, obs))
    return out


def snopt_pin_checks(x, x_qp, status, kkt, active_dist, fix):
    """The asserts both the oracle (CPU) and the CUDA path (GPU) must pass on the reference's one recorded NLP run
    (tests/golden/snopt_print_file.npz, minted by tests/golden/make_snopt_pin.py from /root/reference/print_file.out).
    x = [X16 | U8 | Lambda16 | delta].  A WEAK pin: the inputs are reconstructed from the print file (x_des to ~1e-9, x_0 to ~1e-7,
    obstacle = (0.85, -0.65) of src/A1_Sim_4_agents.cpp:872-873 recovered by a fit to 2e-4) and SNOPT's printed point has 5 digits and
    is not converged (its own exit: "requested accuracy could not be achieved", primal infeasibility 3.1e-4, dual 1.6e3)."""
    xf = fix["x_final"]
    # (1) QP stage: SNOPT's derivative check logs its start point = iSWIFT's QP optimum (MPC_dist.cpp:356-361, 403) to 9 digits
    assert np.abs(x_qp[:24] - fix["x_start"]).max() < 5e-9
    # (2) NLP stage: converged to the BASELINE bar
    assert status == 0 and kkt < 1e-8
    # (3) same active set on the distance rows as SNOPT's final basis: rows 1 and 4 at their bound with positive duals, 2 and 3 free
    assert list(active_dist) == [True, False, False, True]
    assert (fix["dist_duals"] > 0).tolist() == [True, False, False, True]
    # (4) slack within 2e-2 of SNOPT's 0.76247 (measured: 7.6e-3)
    assert abs(x[40] - xf[40]) < 2e-2
    # (5) a feasible point whose objective is not above SNOPT's 803.70 (measured: 773.25 -- SNOPT stopped early; scipy SLSQP
    #     continues from SNOPT's printed point to our answer, tests/test_oracle_nlp.py)
    w = np.array([300.0] * 12 + [2000.0] * 4 + [0.1] * 8 + [0.0] * 16 + [3000.0])
    f = np.zeros(41); f[:16] = -w[:16] * fix["xdes"]
    obj = lambda v: 0.5 * (w * v * v).sum() + f @ v
    assert abs(obj(xf) - float(fix["objective"])) < 2e-2            # the reconstruction reproduces SNOPT's printed objective
    assert obj(x) <= float(fix["objective"]) * (1 + 1e-3)
    # (6) planned positions within 1e-2 m of SNOPT's (measured 6e-3; velocities and inputs differ by up to 0.15: they are what
    #     SNOPT had not finished optimising)
    pos = [0, 2, 4, 6, 8, 10, 12, 14]
    assert np.abs(x[pos] - xf[pos]).max() < 1e-2


# ---------------------------------------------------------------------------------------------------------------- extended mode
def ext_arrays(insts):
    return (np.array([i[0] for i in insts]), np.array([i[1] for i in insts]), np.array([i[2] for i in insts]),
            np.array([i[3] for i in insts], np.int32), np.array([i[4] for i in insts]))


def ext_compare(out, res, T=1):
    """CUDA extended-mode outputs against the list of oracle result dicts.  Teams are compared only when every member converged in
    both (the answer of a team with a failed member depends on where that member froze)."""
    B = len(res)
    st = out["status"][0]; rs = np.array([r["status"] for r in res])
    okm = (st == 0) & (rs == 0)
    if T > 1:
        okt = okm.reshape(-1, T).all(1); okm = np.repeat(okt, T)
    worst = 0.0
    for b in np.nonzero(okm)[0]:
        r = res[b]
        sx = max(1.0, np.abs(r["X"]).max()); su = max(1.0, np.abs(r["U"]).max())
        worst = max(worst, np.abs(out["X"][:, b] - r["X"]).max() / sx, np.abs(out["U"][:, b] - r["U"]).max() / su,
                    abs(out["slack"][0, b] - r["slack"]) / max(1.0, abs(r["slack"])))
    return dict(status_equal=(st == rs), compared=okm, worst=worst, iters_equal=(out["nlp_iters"][0] == np.array([r["nlp_iters"] for r in res])))
