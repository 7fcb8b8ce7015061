"""ctypes view of include/cbfnmpc.h (structs + prototypes).  Pure declarations: used by the product binding
(api.py, which loads libcbfnmpc.so -- CUDA only) and by tests/emu (which loads the host emulator of the kernel
source for CPU-side logic tests)."""
import ctypes as C
import numpy as np


class Gains(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("Qw", "Rw", "Pw", "slack_w", "mu_fric", "h_com", "g_acc", "Ts", "eps_dist", "v_max", "box")]


class Options(C.Structure):
    _fields_ = [("qp_reltol", C.c_double), ("qp_abstol", C.c_double), ("qp_maxit", C.c_int),
                ("tol", C.c_double), ("mu_init", C.c_double), ("kappa_eps", C.c_double), ("kappa_mu", C.c_double),
                ("tau_min", C.c_double), ("push", C.c_double), ("gamma_theta", C.c_double), ("gamma_phi", C.c_double),
                ("eta_phi", C.c_double), ("kappa_sigma", C.c_double), ("alpha_min", C.c_double),
                ("dw_first", C.c_double), ("dw_min", C.c_double), ("kw_minus", C.c_double), ("kw_plus", C.c_double),
                ("kw_plus_first", C.c_double), ("dw_max", C.c_double), ("rho_lambda", C.c_double), ("z_div", C.c_double),
                ("sigma_cap", C.c_double), ("refine_rel", C.c_double), ("refine_abs", C.c_double),
                ("max_iter", C.c_int), ("filter_cap", C.c_int), ("max_refine", C.c_int), ("acc_tol", C.c_double)]


class In(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("xdes", C.c_void_p), ("feet_xy", C.c_void_p), ("contact", C.c_void_p),
                ("obs_xy", C.c_void_p), ("n_obs", C.c_int), ("alpha_buf", C.c_void_p)]


class Out(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("X", "U", "lam", "slack", "x_qp", "alpha", "alpha_buf_next", "kkt",
                                          "status", "qp_iters", "nlp_iters", "n_fact", "active")]


class RawIn(C.Structure):
    """cbfnmpc_raw_in (include/cbfnmpc.h): device pointers."""
    _fields_ = [("q", C.c_void_p), ("dq", C.c_void_p), ("nq", C.c_int), ("contact", C.c_void_p), ("toe_pos", C.c_void_p),
                ("gait_domain", C.c_void_p), ("agent_id", C.c_void_p), ("team", C.c_void_p), ("agent_initial", C.c_void_p),
                ("Pr_refined", C.c_void_p), ("Prd_refined", C.c_void_p), ("ref_cols", C.c_int), ("n_teams", C.c_int),
                ("Pobs_real", C.c_void_p), ("n_obs", C.c_int), ("state_other", C.c_void_p), ("n_other", C.c_int)]


class WorldState(C.Structure):
    """cbfnmpc_world_state (include/cbfnmpc.h): device pointers."""
    _fields_ = [(k, C.c_void_p) for k in ("q", "dq", "contact", "toe_pos", "gait_domain", "alpha_buf", "success", "distance_to_fail",
                                          "n_not_converged", "traj")]


OUT_F64 = dict(X=16, U=8, lam=16, slack=1, x_qp=24, alpha=20, alpha_buf_next=4, kkt=1)
OUT_I32 = dict(status=1, qp_iters=1, nlp_iters=1, n_fact=1)
IN_F64 = dict(x0=4, xdes=16, feet_xy=8, alpha_buf=4)  # + obs_xy 2K, contact (int32) 4

MODE_QP, MODE_QP_NLP = 0, 1
ST_CONVERGED, ST_MAXITER, ST_INFEASIBLE, ST_NUMERICAL, ST_BADINPUT = range(5)


def soa_from_instances(insts, alpha_buf=None):
    """[(x0, xdes, feet(2,4), contact, obs(2,K)), ...] -> dict of SoA numpy arrays ([field][B], C-contiguous)."""
    B = len(insts)
    K = insts[0][4].shape[1]
    d = dict(x0=np.zeros((4, B)), xdes=np.zeros((16, B)), feet_xy=np.zeros((8, B)), contact=np.zeros((4, B), np.int32),
             obs_xy=np.zeros((2 * K, B)), alpha_buf=np.zeros((4, B)))
    for b, (x0, xdes, feet, contact, obs) in enumerate(insts):
        d["x0"][:, b] = x0; d["xdes"][:, b] = xdes; d["feet_xy"][:, b] = np.asarray(feet).reshape(8)
        d["contact"][:, b] = contact; d["obs_xy"][:, b] = np.asarray(obs).reshape(2 * K)
        d["alpha_buf"][:, b] = (x0[0], 0.0, x0[2], 0.0) if alpha_buf is None else alpha_buf[b]
    d["n_obs"] = K
    return d


def alloc_out(B):
    o = {k: np.zeros((n, B)) for k, n in OUT_F64.items()}
    o.update({k: np.zeros((n, B), np.int32) for k, n in OUT_I32.items()})
    o["active"] = np.zeros((2, B), np.uint64)
    return o


def _check(name, a, dtype, rows, B):
    """The C ABI reads raw pointers: a wrong dtype, a transposed view or a short array would be read as garbage."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("%s must be a C-contiguous numpy array of %s" % (name, np.dtype(dtype).name))
    if a.size != rows * B:
        raise ValueError("%s must hold %d x %d elements ([field][B]), has %d" % (name, rows, B, a.size))


def in_struct(d):
    s = In()
    B = d["x0"].shape[1]; K = int(d["n_obs"])
    rows = dict(IN_F64, obs_xy=2 * K)
    for k in ("x0", "xdes", "feet_xy", "obs_xy", "alpha_buf"):
        _check(k, d[k], np.float64, rows[k], B)
        setattr(s, k, d[k].ctypes.data)
    _check("contact", d["contact"], np.int32, 4, B)
    s.contact = d["contact"].ctypes.data
    s.n_obs = K
    return s


def out_struct(o, B=None):
    """Arrays that are None (or absent) are passed as NULL: the library then skips them (and their download)."""
    s = Out()
    for k, v in o.items():
        if v is None:
            continue
        dt, rows = (np.float64, OUT_F64[k]) if k in OUT_F64 else ((np.int32, 1) if k in OUT_I32 else (np.uint64, 2))
        _check(k, v, dt, rows, v.size // rows if B is None else B)
        setattr(s, k, v.ctypes.data)
    return s


# ---------------------------------------------------------------------------------------------------------------- extended mode
class ExtCfg(C.Structure):
    _fields_ = [("N", C.c_int), ("team_size", C.c_int), ("eps_obs", C.c_double), ("eps_nbr", C.c_double), ("relax", C.c_double), ("pub_tol", C.c_double), ("max_team_iters", C.c_int), ("mu_stall_iters", C.c_int)]


class ExtIn(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("xdes", C.c_void_p), ("feet_xy", C.c_void_p), ("contact", C.c_void_p), ("obs_xy", C.c_void_p), ("n_obs", C.c_int),
                ("alpha_buf", C.c_void_p), ("nbr_traj", C.c_void_p), ("n_nbr", C.c_int)]


class ExtOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("X", "U", "lam", "slack", "x_qp", "alpha", "alpha_buf_next", "kkt", "pub", "status", "qp_iters", "nlp_iters", "team_iters")]


class ExtStats(C.Structure):
    _fields_ = [("launches", C.c_int), ("iterations", C.c_int), ("exchanges", C.c_int), ("nranks", C.c_int), ("bytes_per_exchange", C.c_longlong),
                ("total_ms", C.c_float), ("fused", C.c_int)]


def ext_out_rows(N):
    return dict(X=4 * N, U=2 * N, lam=4 * N, slack=1, x_qp=6 * N, alpha=20, alpha_buf_next=4, kkt=1, pub=2 * N)


EXT_OUT_I32 = ("status", "qp_iters", "nlp_iters", "team_iters")


def ext_soa_from_instances(insts, N):
    """[(x0, xdes[4N], feet[ND,2,4], contact, obs[2,K]), ...] -> SoA arrays of cbfnmpc_ext_in."""
    B = len(insts); K = insts[0][4].shape[1]; ND = N // 4
    d = dict(x0=np.zeros((4, B)), xdes=np.zeros((4 * N, B)), feet_xy=np.zeros((8 * ND, B)), contact=np.zeros((4, B), np.int32),
             obs_xy=np.zeros((2 * max(K, 1), B)), alpha_buf=np.zeros((4, B)))
    for b, (x0, xdes, feet, contact, obs) in enumerate(insts):
        d["x0"][:, b] = x0; d["xdes"][:, b] = xdes; d["feet_xy"][:, b] = np.asarray(feet, float).reshape(8 * ND); d["contact"][:, b] = contact
        if K:
            d["obs_xy"][:2 * K, b] = np.asarray(obs).reshape(2 * K)
        d["alpha_buf"][:, b] = (x0[0], 0.0, x0[2], 0.0)
    d["n_obs"] = K
    return d


def ext_alloc_out(B, N):
    o = {k: np.zeros((n, B)) for k, n in ext_out_rows(N).items()}
    o.update({k: np.zeros((1, B), np.int32) for k in EXT_OUT_I32})
    return o
