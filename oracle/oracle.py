"""TEST INFRASTRUCTURE -- ctypes bindings for the C oracle (liborc.so) and the compiled reference iSWIFT
core (_ref/libiswift_ref.so).  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NV_MAX, NEQ, MROWS_MAX = 41, 28, 184


class Options(C.Structure):
    _fields_ = [("qp_reltol", C.c_double), ("qp_abstol", C.c_double), ("qp_maxit", C.c_int),
                ("tol", C.c_double), ("mu_init", C.c_double), ("kappa_eps", C.c_double), ("kappa_mu", C.c_double),
                ("theta_mu", C.c_double), ("tau_min", C.c_double), ("push", C.c_double),
                ("gamma_theta", C.c_double), ("gamma_phi", C.c_double), ("delta", C.c_double),
                ("s_theta", C.c_double), ("s_phi", C.c_double), ("eta_phi", C.c_double),
                ("kappa_sigma", C.c_double), ("alpha_min", C.c_double),
                ("dw_first", C.c_double), ("dw_min", C.c_double), ("kw_minus", C.c_double), ("kw_plus", C.c_double),
                ("kw_plus_first", C.c_double), ("dw_max", C.c_double), ("rho_lambda", C.c_double), ("z_div", C.c_double),
                ("max_iter", C.c_int), ("filter_cap", C.c_int), ("acc_tol", C.c_double)]


class Gains(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("Qw", "Rw", "Pw", "slack_w", "mu_fric", "h_com", "g_acc", "Ts", "eps_dist",
                                          "v_max", "box")]


class Result(C.Structure):
    _fields_ = [("x", C.c_double * NV_MAX), ("x_qp", C.c_double * NV_MAX), ("y", C.c_double * NEQ),
                ("z", C.c_double * MROWS_MAX), ("s", C.c_double * MROWS_MAX), ("alpha", C.c_double * 20),
                ("alpha_buf_next", C.c_double * 4), ("obs", C.c_double * 2), ("kkt", C.c_double), ("mu", C.c_double),
                ("nv", C.c_int), ("c", C.c_int), ("m", C.c_int), ("status", C.c_int), ("qp_flag", C.c_int),
                ("qp_iters", C.c_int), ("nlp_iters", C.c_int), ("n_fact", C.c_int), ("n_ls", C.c_int)]


def build(force=False):
    """Compile liborc.so (always possible) and _ref/libiswift_ref.so (only where /root/reference exists)."""
    lib = os.path.join(HERE, "liborc.so")
    src = [os.path.join(HERE, f) for f in ("cbfnmpc_oracle.c", "cbfnmpc_oracle.h", "cbfnmpc_oracle_ext.c", "cbfnmpc_oracle_ext.h")]
    stale = (not os.path.exists(lib)) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in src)
    ref = os.path.join(HERE, "_ref", "libiswift_ref.so")
    need_ref = os.path.isdir("/root/reference/optimization/iSWIFT/src") and (
        force or not os.path.exists(ref) or not os.path.exists(ref.replace(".so", "_v3.so"))
        or os.path.getmtime(os.path.join(HERE, "ref_driver.c")) > os.path.getmtime(ref))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "liborc.so"] + (["-B"] if force else []))
    if need_ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def build_native():
    """CPU-baseline build of the oracle for THIS machine: gcc -O3 -march=native, compiled where it runs (the
    GPU box's host differs from the build container, so this file never travels: _native/ is git- and gpurun-ignored).
    The parity tests keep liborc.so (-march=x86-64-v2: no FMA contraction, the same bits everywhere).  Returns the path or None."""
    import hashlib
    flags = " ".join(sorted(_cpu_flags()))
    tag = hashlib.sha1(flags.encode()).hexdigest()[:10]
    out = os.path.join(HERE, "_native", "liborc_native_%s.so" % tag)
    src = os.path.join(HERE, "cbfnmpc_oracle.c"); src2 = os.path.join(HERE, "cbfnmpc_oracle_ext.c")
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(src2)):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    try:
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=gnu99", "-w", "-shared", "-o", out, src, src2, "-lm", "-lpthread"])
        return out
    except Exception:
        return None


_lib = None
_ref = None
_native = False
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        path = os.path.join(HERE, "liborc.so")
        if _native:
            path = build_native() or path
        _lib = C.CDLL(path)
        _lib.build_path = path
        _lib.orc_default_options.argtypes = [C.POINTER(Options)]
        _lib.orc_default_gains.argtypes = [C.POINTER(Gains)]
        _lib.orc_lip_discrete.argtypes = [C.POINTER(Gains), dp, dp]
        _lib.orc_stance_matrix.argtypes = [dp, ip, dp]
        _lib.orc_pack_dense.argtypes = [C.POINTER(Gains), dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp]
        _lib.orc_closest_obstacle.argtypes = [dp, dp, C.c_int, dp]
        _lib.orc_generate_reference.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, dp, dp]
        _lib.orc_cop_planner.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, dp]
        _lib.orc_footholds.argtypes = [dp, C.c_int, dp, dp]
        _lib.orc_bezier_fit.argtypes = [dp, dp, dp, dp]
        _lib.orc_qp_mehrotra.argtypes = [C.c_int] * 3 + [dp] * 6 + [C.c_double, C.c_double, C.c_int, C.c_double] + [dp] * 4 + [C.POINTER(C.c_int)]
        _lib.orc_nlp_ipm.argtypes = [C.c_int, C.c_int] + [dp] * 6 + [dp, dp, C.c_int, C.POINTER(Gains), C.POINTER(Options), C.POINTER(Result)]
        _lib.orc_solve_instance.argtypes = [C.POINTER(Gains), C.POINTER(Options), dp, dp, dp, ip, dp, C.c_int, dp, C.c_int, C.POINTER(Result)]
        _lib.orc_solve_batch.argtypes = [C.POINTER(Gains), C.POINTER(Options), C.c_int, dp, dp, dp, ip, dp, C.c_int, dp, C.c_int, C.c_int, C.POINTER(Result)]
    return _lib


def ref_lib():
    """The reference's own iSWIFT C core (None when it was never built, e.g. a checkout without _ref/)."""
    global _ref
    if _ref is None:
        build()
        path = os.path.join(HERE, "_ref", "libiswift_ref.so")
        v3 = os.path.join(HERE, "_ref", "libiswift_ref_v3.so")
        if _native and os.path.exists(v3) and {"avx2", "fma", "bmi2"} <= _cpu_flags():
            path = v3   # the reference core built -O3 -march=x86-64-v3 (its sources do not travel, so it cannot be -march=native on the box)
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.iswift_ref_solve_dense.argtypes = [C.c_int] * 3 + [dp] * 6 + [dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _ref.iswift_ref_solve_ccs.argtypes = [C.c_int] * 3 + [ip, ip, dp] * 3 + [dp, dp, dp, C.c_double, ip, dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return _ref


def use_native_build(on=True):
    """bench.py's CPU baseline only: switch this module to the -O3 -march=native build of the oracle (and the -march=x86-64-v3
    build of the reference core).  Drops the loaded handles; returns a description of what is now in use."""
    global _lib, _ref, _native
    _native = bool(on); _lib = None; _ref = None
    L = lib(); r = ref_lib()
    return {"oracle": os.path.relpath(L.build_path, HERE), "iswift": os.path.basename(r._name) if r is not None else None}


def use_reference_qp(on=True):
    """QP stage of solve_instance / solve_batch by the reference's own iSWIFT C core (oracle/_ref) instead of the restatement.
    Returns True when the core is available and now in use.  Used by bench.py's CPU baseline only; the parity tests keep the
    restatement (which they pin against that core separately)."""
    L = lib()
    L.orc_set_qp_backend.argtypes = [C.c_void_p]
    r = ref_lib() if on else None
    if r is None:
        L.orc_set_qp_backend(None)
        return False
    L.orc_set_qp_backend(C.cast(r.iswift_ref_solve_dense, C.c_void_p))
    return True


def default_options():
    o = Options(); lib().orc_default_options(C.byref(o)); return o


def default_gains():
    g = Gains(); lib().orc_default_gains(C.byref(g)); return g


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def generate_reference(Pstart, Pobs, goal=(10.0, 0.0), loop=100000):
    """generateReferenceTrajectory (src/MPC_dist.cpp:930-1062).  Pstart (8), Pobs 2 x K.  Returns Pr, Prd (8 x loop/40)."""
    Pobs = np.asarray(Pobs, float)
    ncols = loop // 40
    Pr = np.zeros(8 * ncols); Prd = np.zeros(8 * ncols)
    lib().orc_generate_reference(_f(Pstart), np.ascontiguousarray(Pobs.T).reshape(-1), Pobs.shape[1], goal[0], goal[1], loop, Pr, Prd)
    return Pr.reshape(ncols, 8).T.copy(), Prd.reshape(ncols, 8).T.copy()


def cop_planner(Pr, Prd, agent_id, gait_domain):
    """Pr, Prd: 8 x ref_cols (any layout; passed column-major like Eigen).  Returns xdes (16)."""
    Prc = np.ascontiguousarray(np.asarray(Pr, float).T); Prdc = np.ascontiguousarray(np.asarray(Prd, float).T)
    out = np.zeros(16)
    lib().orc_cop_planner(Prc.reshape(-1), Prdc.reshape(-1), Prc.shape[0], int(agent_id), int(gait_domain), out)
    return out


def footholds(toe_pos, gait_domain, agent_initial):
    """toe_pos 3 x 4.  Returns feet_xy (8): x of FR,FL,RR,RL then y."""
    tp = np.ascontiguousarray(np.asarray(toe_pos, float).T).reshape(-1)  # column-major 3x4
    out = np.zeros(8)
    lib().orc_footholds(tp, int(gait_domain), _f(agent_initial), out)
    return out


def lip_discrete(g=None):
    g = g or default_gains()
    Ad = np.zeros(16); Bd = np.zeros(8)
    lib().orc_lip_discrete(C.byref(g), Ad, Bd)
    return Ad.reshape(4, 4), Bd.reshape(4, 2)


def pack_dense(x0, xdes, F, g=None):
    g = g or default_gains()
    c = F.shape[1]
    nv, p, mi = 25 + 4 * c, 28, 60 + 8 * c
    Q = np.zeros(nv * nv); f = np.zeros(nv); A = np.zeros(p * nv); b = np.zeros(p); G = np.zeros(mi * nv); h = np.zeros(mi)
    lib().orc_pack_dense(C.byref(g), _f(x0), _f(xdes), _f(F).ravel(), c, Q, f, A, b, G, h)
    return (Q.reshape(nv, nv, order="F"), f, A.reshape(p, nv, order="F"), b, G.reshape(mi, nv, order="F"), h)


def qp_mehrotra(P, c, A, b, G, h, reltol=1e-6, abstol=1e-6, maxit=25, sigma_d=0.0):
    n, m, p = P.shape[0], G.shape[0], A.shape[0]
    x = np.zeros(n); y = np.zeros(max(p, 1)); z = np.zeros(m); s = np.zeros(m); it = C.c_int(0)
    flag = lib().orc_qp_mehrotra(n, m, p, _f(P.ravel(order="F")), _f(c), _f(A.ravel(order="F")), _f(b),
                                 _f(G.ravel(order="F")), _f(h), reltol, abstol, maxit, sigma_d, x, y, z, s, C.byref(it))
    return x, y[:p], z, s, it.value, flag


def iswift_ref_dense(P, c, A, b, G, h):
    r = ref_lib()
    n, m, p = P.shape[0], G.shape[0], A.shape[0]
    x = np.zeros(n); y = np.zeros(max(p, 1)); z = np.zeros(m); s = np.zeros(m); it = C.c_int(0)
    ts, tv = C.c_double(0), C.c_double(0)
    flag = r.iswift_ref_solve_dense(n, m, p, _f(P.ravel(order="F")), _f(c), _f(A.ravel(order="F")), _f(b),
                                    _f(G.ravel(order="F")), _f(h), x, y, z, s, C.byref(it), C.byref(ts), C.byref(tv))
    return x, y[:p], z, s, it.value, flag, ts.value, tv.value


def res_to_dict(r):
    nv, m = r.nv, r.m
    return dict(x=np.array(r.x[:nv]), x_qp=np.array(r.x_qp[:nv]), y=np.array(r.y[:]), z=np.array(r.z[:m]), s=np.array(r.s[:m]),
                alpha=np.array(r.alpha[:]).reshape(4, 5), alpha_buf_next=np.array(r.alpha_buf_next[:]), obs=np.array(r.obs[:]),
                kkt=r.kkt, mu=r.mu, nv=nv, c=r.c, m=m, status=r.status, qp_flag=r.qp_flag, qp_iters=r.qp_iters,
                nlp_iters=r.nlp_iters, n_fact=r.n_fact, n_ls=r.n_ls)


def nlp_ipm(Q, f, A, b, G, h, obs, x_start, with_nl=True, g=None, o=None):
    g = g or default_gains(); o = o or default_options()
    nv, mi = Q.shape[0], G.shape[0]
    r = Result()
    lib().orc_nlp_ipm(nv, mi, _f(Q.ravel(order="F")), _f(f), _f(A.ravel(order="F")), _f(b), _f(G.ravel(order="F")), _f(h),
                      _f(obs), _f(x_start), int(with_nl), C.byref(g), C.byref(o), C.byref(r))
    return res_to_dict(r)


def nlp_ipm_rows(Q, f, A, b, G, h, dstage, dctr, deps, x_start, with_vel=True, g=None, o=None):
    """The NLP with a general set of distance rows |p_{dstage[r]+1} - dctr[r]|^2 + delta >= deps[r] (extended mode)."""
    g = g or default_gains(); o = o or default_options()
    nv, mi = Q.shape[0], G.shape[0]
    ds = np.ascontiguousarray(dstage, dtype=np.int32); nd = len(ds)
    r = Result()
    L = lib()
    L.orc_nlp_ipm_rows.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, ip, dp, dp, dp, C.c_int, C.POINTER(Gains),
                                   C.POINTER(Options), C.POINTER(Result)]
    L.orc_nlp_ipm_rows(nv, mi, _f(Q.ravel(order="F")), _f(f), _f(A.ravel(order="F")), _f(b), _f(G.ravel(order="F")), _f(h), nd, ds,
                       _f(np.asarray(dctr, float).reshape(-1)), _f(deps), _f(x_start), int(with_vel), C.byref(g), C.byref(o), C.byref(r))
    return res_to_dict(r)


def solve_instance_ext(x0, xdes, feet, contact, centres, eps, g=None, o=None):
    """Extended-mode instance: QP stage as in the reference, then the NLP with n_c distance rows per stage.  centres
    (n_c, 4, 2): centre of row i at stage k (an obstacle: the same point at every stage; a neighbour: its predicted position at
    that stage); eps (n_c,).  Returns the dict of nlp_ipm_rows plus x_qp."""
    g = g or default_gains(); o = o or default_options()
    F = np.zeros(8); c = lib().orc_stance_matrix(_f(np.asarray(feet, float).reshape(-1)), np.ascontiguousarray(contact, dtype=np.int32), F)
    if c < 0:
        return dict(status=4)
    Fm = F[:2 * c].reshape(2, c)
    Q, f, A, b, G, h = pack_dense(np.asarray(x0, float), np.asarray(xdes, float), Fm, g)
    xq = qp_mehrotra(Q, f, A, b, G, h, o.qp_reltol, o.qp_abstol, o.qp_maxit)[0]
    centres = np.asarray(centres, float); n_c = centres.shape[0]
    dstage = np.tile(np.arange(4), n_c); dctr = centres.reshape(-1, 2); deps = np.repeat(np.asarray(eps, float), 4)
    r = nlp_ipm_rows(Q, f, A, b, G, h, dstage, dctr, deps, xq, True, g, o)
    r["x_qp"] = xq
    return r


def solve_instance(x0, xdes, feet, contact, obs, alpha_buf=None, mode=1, g=None, o=None):
    g = g or default_gains(); o = o or default_options()
    r = Result()
    obs = _f(obs)
    ab = _f(alpha_buf if alpha_buf is not None else [x0[0], 0, x0[2], 0])
    lib().orc_solve_instance(C.byref(g), C.byref(o), _f(x0), _f(xdes), _f(feet).ravel(), np.ascontiguousarray(contact, np.int32),
                             obs.ravel(), obs.shape[1], ab, mode, C.byref(r))
    return res_to_dict(r)


def solve_batch(x0, xdes, feet, contact, obs, alpha_buf, mode=1, nthreads=1, g=None, o=None, raw=False):
    """AoS batch: x0 (B,4) xdes (B,16) feet (B,2,4) contact (B,4) obs (B,2,K) alpha_buf (B,4)."""
    g = g or default_gains(); o = o or default_options()
    B = x0.shape[0]; K = obs.shape[2]
    res = (Result * B)()
    lib().orc_solve_batch(C.byref(g), C.byref(o), B, _f(x0).ravel(), _f(xdes).ravel(), _f(feet).ravel(),
                          np.ascontiguousarray(contact, np.int32).ravel(), _f(obs).ravel(), K, _f(alpha_buf).ravel(), mode,
                          nthreads, res)
    if raw:
        return res
    out = dict(X=np.zeros((B, 16)), U=np.zeros((B, 8)), lam=np.zeros((B, 16)), slack=np.zeros(B), x_qp=np.zeros((B, NV_MAX)),
               alpha=np.zeros((B, 20)), alpha_buf_next=np.zeros((B, 4)), status=np.zeros(B, np.int32), kkt=np.zeros(B),
               qp_iters=np.zeros(B, np.int32), qp_flag=np.zeros(B, np.int32), nlp_iters=np.zeros(B, np.int32),
               n_fact=np.zeros(B, np.int32), n_ls=np.zeros(B, np.int32), c=np.zeros(B, np.int32), obs=np.zeros((B, 2)),
               z=np.zeros((B, MROWS_MAX)), s=np.zeros((B, MROWS_MAX)), y=np.zeros((B, NEQ)))
    for i in range(B):
        r = res[i]; nv = r.nv; c = r.c
        xx = np.array(r.x[:NV_MAX])
        out["X"][i] = xx[:16]; out["U"][i] = xx[16:24]
        if nv > 0:
            out["lam"][i, :4 * c] = xx[24:24 + 4 * c]; out["slack"][i] = xx[nv - 1]
        out["x_qp"][i] = r.x_qp[:]; out["alpha"][i] = r.alpha[:]; out["alpha_buf_next"][i] = r.alpha_buf_next[:]
        out["status"][i] = r.status; out["kkt"][i] = r.kkt; out["qp_iters"][i] = r.qp_iters; out["qp_flag"][i] = r.qp_flag
        out["nlp_iters"][i] = r.nlp_iters; out["n_fact"][i] = r.n_fact; out["n_ls"][i] = r.n_ls; out["c"][i] = c
        out["obs"][i] = r.obs[:]; out["z"][i] = r.z[:]; out["s"][i] = r.s[:]; out["y"][i] = r.y[:]
    return out


def closed_loop(world, n_cycles, mode=1, fail_radius=0.5, nthreads=8):
    """CPU restatement of cbfnmpc_closed_loop: per cycle the planners (orc_cop_planner, orc_footholds), one
    solve per world (orc_solve_batch), MPC_dist::updateDistance_to_fail (src/MPC_dist.cpp:21-40) at the 4 grid points, and the
    world update (state <- x_4, gaitDomain_ += 1, alpha buffer, nominal footprint about the new position, trot contacts).
    `world`: host arrays in the layout of scenarios.mc_worlds; updated in place; returns the trajectory [n_cycles][4][B]."""
    w = world
    B = w["q"].shape[1]; K = w["n_obs"]
    fp = np.array([[0.2188, 0.2188, -0.1472, -0.1472], [-0.1320, 0.1320, -0.1320, 0.1320]])   # src/MPC_dist.cpp:1206-1209
    traj = np.zeros((n_cycles, 4, B))
    for cyc in range(n_cycles):
        x0 = np.stack([w["q"][0], w["dq"][0], w["q"][1], w["dq"][1]], 1)
        xdes = np.zeros((B, 16)); feet = np.zeros((B, 2, 4)); obs = np.zeros((B, 2, K))
        for b in range(B):
            t = int(w["team"][b])
            xdes[b] = cop_planner(w["Pr_refined"][t].T, w["Prd_refined"][t].T, int(w["agent_id"][b]), int(w["gait_domain"][b]))
            feet[b] = footholds(w["toe_pos"][:, b].reshape(4, 3).T, int(w["gait_domain"][b]), w["agent_initial"][:, b]).reshape(2, 4)
            obs[b] = w["Pobs_real"][t].T
        r = solve_batch(x0, xdes, feet, np.ascontiguousarray(w["contact"].T), obs, np.ascontiguousarray(w["alpha_buf"].T), mode=mode,
                        nthreads=nthreads)
        X = r["X"]
        w["n_not_converged"] += (r["status"] != 0).astype(np.int32)
        for b in range(B):
            if w["success"][b]:
                done = False
                for k in range(4):
                    px, py = X[b, 4 * k], X[b, 4 * k + 2]
                    for i in range(K):
                        dx, dy = px - obs[b, 0, i], py - obs[b, 1, i]
                        if np.sqrt(dx * dx + dy * dy) < fail_radius:
                            w["success"][b] = 0; w["distance_to_fail"][b] = np.sqrt(px * px + py * py); done = True
                            break
                    if done: break
        w["q"][0] = X[:, 12]; w["dq"][0] = X[:, 13]; w["q"][1] = X[:, 14]; w["dq"][1] = X[:, 15]
        traj[cyc] = X[:, 12:16].T
        w["alpha_buf"][:] = r["alpha_buf_next"].T
        w["gait_domain"] += 1
        g = w["gait_domain"]
        for i in range(4):
            w["toe_pos"][3 * i] = w["q"][0] + fp[0, i]; w["toe_pos"][3 * i + 1] = w["q"][1] + fp[1, i]; w["toe_pos"][3 * i + 2] = 0.0
            diag = 1 if i in (0, 3) else 0
            w["contact"][i] = np.where(g <= 1, 1, np.where(g & 1, diag, 1 - diag))
    return traj


# ------------------------------------------------------------------------------------------------ extended mode
EXT_NMAX, EXT_KMAX = 8, 19
EXT_NV_MAX = EXT_NMAX * 10 + 1


class ExtResult(C.Structure):
    _fields_ = [("x", C.c_double * EXT_NV_MAX), ("x_qp", C.c_double * EXT_NV_MAX), ("kkt", C.c_double), ("mu", C.c_double),
                ("N", C.c_int), ("c", C.c_int), ("nv", C.c_int), ("n_rows", C.c_int), ("status", C.c_int), ("qp_flag", C.c_int),
                ("qp_iters", C.c_int), ("nlp_iters", C.c_int), ("n_fact", C.c_int), ("n_ls", C.c_int), ("n_bump", C.c_int),
                ("zd", C.c_double * (EXT_NMAX * EXT_KMAX)), ("sd", C.c_double * (EXT_NMAX * EXT_KMAX))]


def _ext_lib():
    L = lib()
    if not getattr(L, "_ext_ready", False):
        vp = C.c_void_p
        L.orc_ext_pack_dense.argtypes = [C.POINTER(Gains), C.c_int, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp]
        L.orc_ext_solve.argtypes = [C.POINTER(Gains), C.POINTER(Options), C.c_int, dp, dp, dp, ip, C.c_int, dp, C.c_double, C.c_int,
                                    C.c_double, vp, C.POINTER(ExtResult)]
        L.orc_ext_solve_teams.argtypes = [C.POINTER(Gains), C.POINTER(Options), C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, C.c_int, dp,
                                          C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(ExtResult), ip]
        L._ext_ready = True
    return L


def ext_pack_dense(N, x0, xdes, F, g=None):
    """F: (ND, 2, c) stance matrices, one per gait domain."""
    g = g or default_gains()
    F = np.asarray(F, float); c = F.shape[2]
    nv, p, mi = N * (6 + c) + 1, 7 * N, 16 * N - 4 + 2 * c * N
    Q = np.zeros(nv * nv); f = np.zeros(nv); A = np.zeros(p * nv); b = np.zeros(p); G = np.zeros(mi * nv); h = np.zeros(mi)
    _ext_lib().orc_ext_pack_dense(C.byref(g), N, _f(x0), _f(xdes), _f(F).ravel(), c, Q, f, A, b, G, h)
    return (Q.reshape(nv, nv, order="F"), f, A.reshape(p, nv, order="F"), b, G.reshape(mi, nv, order="F"), h)


def ext_res_to_dict(r):
    N, c, nv, K = r.N, r.c, r.nv, r.n_rows
    x = np.array(r.x[:max(nv, 1)])
    d = dict(status=r.status, N=N, c=c, nv=nv, kkt=r.kkt, mu=r.mu, qp_flag=r.qp_flag, qp_iters=r.qp_iters, nlp_iters=r.nlp_iters,
             n_fact=r.n_fact, n_ls=r.n_ls, n_bump=r.n_bump, x=x, x_qp=np.array(r.x_qp[:max(nv, 1)]))
    if nv > 0:
        d.update(X=x[:4 * N], U=x[4 * N:6 * N], lam=x[6 * N:6 * N + c * N], slack=x[nv - 1],
                 zd=np.array(r.zd[:K * N]).reshape(K, N), sd=np.array(r.sd[:K * N]).reshape(K, N))
    return d


def ext_solve(x0, xdes, feet, contact, obs, nbr=None, N=4, eps_obs=None, eps_nbr=2.2, g=None, o=None):
    """One extended-mode agent with FIXED neighbour trajectories.  feet (ND, 2, 4); obs (2, K_obs); nbr (n_nbr, N, 2) or None."""
    g = g or default_gains(); o = o or default_options()
    obs = _f(obs); K = obs.shape[1] if obs.size else 0
    nn = 0 if nbr is None else np.asarray(nbr).shape[0]
    nb = _f(nbr) if nn else None
    r = ExtResult()
    _ext_lib().orc_ext_solve(C.byref(g), C.byref(o), N, _f(x0), _f(xdes), _f(feet).ravel(), np.ascontiguousarray(contact, np.int32), K,
                             obs.ravel(), g.eps_dist if eps_obs is None else eps_obs, nn, eps_nbr, nb.ctypes.data if nn else None, C.byref(r))
    return ext_res_to_dict(r)


EXT_RELAX, EXT_PUB_TOL, EXT_MAX_TEAM_ITERS, EXT_MU_STALL = 0.7, 1e-10, 64, 6


def ext_solve_teams(x0, xdes, feet, contact, obs, T, N=4, eps_obs=None, eps_nbr=2.2, relax=EXT_RELAX, pub_tol=EXT_PUB_TOL, max_team_iters=EXT_MAX_TEAM_ITERS, mu_stall=EXT_MU_STALL, nthreads=8, g=None, o=None):
    """nteams teams of T agents (agent-major arrays with leading dimension nteams * T): x0 (B,4) xdes (B,4N) feet (B,ND,2,4)
    contact (B,4) obs (B,2,K_obs).  T = 1: independent agents.  Returns (list of result dicts, team iteration counts)."""
    g = g or default_gains(); o = o or default_options()
    B = x0.shape[0]; nteams = B // T; K = obs.shape[2]
    res = (ExtResult * B)(); ti = np.zeros(nteams, np.int32)
    _ext_lib().orc_ext_solve_teams(C.byref(g), C.byref(o), nteams, T, N, _f(x0).ravel(), _f(xdes).ravel(), _f(feet).ravel(),
                                   np.ascontiguousarray(contact, np.int32).ravel(), K, _f(obs).ravel(), g.eps_dist if eps_obs is None else eps_obs,
                                   eps_nbr, relax, pub_tol, max_team_iters, mu_stall, nthreads, res, ti)
    return [ext_res_to_dict(r) for r in res], ti
