"""DEV/TEST TOOL: host emulation of the extended-mode kernel source (tests/emu/emu_ext.cpp: srb-cbf-nmpc_b200/csrc/cbfnmpc_ext.h
compiled for the host, one warp on 32 cooperative fibers).  Never imported by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np
import emu as _emu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
capi = _emu.capi
_lib = None


class Args(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "N", "T", "K_obs", "n_nbr", "xmode")] + [(k, C.c_double) for k in ("eps_obs", "eps_nbr", "relax", "pub_tol")] + [("max_team_iters", C.c_int), ("mu_stall", C.c_int)] + \
               [(k, C.c_void_p) for k in ("x0", "xdes", "feet", "obs", "abuf", "nbr", "contact", "X", "U", "lam", "slack", "xqp", "alpha", "abuf_next",
                                          "kkt", "pub", "status", "qp_iters", "nlp_iters", "team_iters")]


def build(force=False, trace=False):
    lib = os.path.join(HERE, "libcbfnmpc_ext_emu%s.so" % ("_trace" if trace else ""))
    csrc = os.path.join(ROOT, "srb-cbf-nmpc_b200", "csrc")
    srcs = [os.path.join(HERE, "emu_ext.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas"] + (["-DCBF_TRACE"] if trace else []) +
                              ["-o", lib, os.path.join(HERE, "emu_ext.cpp")])
    return lib


def ext_soa(insts, N):
    """[(x0, xdes[4N], feet[ND,2,4], contact, obs[2,K]), ...] -> SoA arrays of the extended ABI."""
    B = len(insts); K = insts[0][4].shape[1]; ND = N // 4
    d = dict(x0=np.zeros((4, B)), xdes=np.zeros((4 * N, B)), feet_xy=np.zeros((8 * ND, B)), contact=np.zeros((4, B), np.int32),
             obs_xy=np.zeros((2 * max(K, 1), B)), alpha_buf=np.zeros((4, B)))
    for b, (x0, xdes, feet, contact, obs) in enumerate(insts):
        d["x0"][:, b] = x0; d["xdes"][:, b] = xdes; d["feet_xy"][:, b] = np.asarray(feet, float).reshape(8 * ND); d["contact"][:, b] = contact
        if K: d["obs_xy"][:2 * K, b] = np.asarray(obs).reshape(2 * K)
        d["alpha_buf"][:, b] = (x0[0], 0.0, x0[2], 0.0)
    d["n_obs"] = K
    return d


def ext_alloc_out(B, N):
    o = dict(X=np.zeros((4 * N, B)), U=np.zeros((2 * N, B)), lam=np.zeros((4 * N, B)), slack=np.zeros((1, B)), x_qp=np.zeros((6 * N, B)),
             alpha=np.zeros((20, B)), alpha_buf_next=np.zeros((4, B)), kkt=np.zeros((1, B)), pub=np.zeros((2 * N + 2, B)))
    o.update({k: np.zeros((1, B), np.int32) for k in ("status", "qp_iters", "nlp_iters", "team_iters")})
    return o


def solve(insts, N=4, T=1, nbr=None, eps_obs=None, eps_nbr=2.2, relax=0.7, pub_tol=1e-10, max_team_iters=64, mu_stall=6, gains=None, options=None, trace=False,
          fused_members=0):
    """nbr: (B, n_nbr, N, 2) fixed neighbour trajectories (T == 1) or None; T == 4: teams exchange through the emulated warp.
    fused_members = m > 1 (with T == 1): the batch is MEMBER-MAJOR (instance b = member * n_teams + team, m members per team) and the
    members exchange through the mailbox and sequence flags of the fused path (xmode 4); the whole batch must fit one warp."""
    global _lib
    L = C.CDLL(build(trace=trace))
    L.emu_ext_solve.argtypes = [C.POINTER(capi.Gains), C.POINTER(capi.Options), C.POINTER(Args)]
    E = _emu.lib()
    g = capi.Gains(); E.cbfnmpc_default_gains(C.byref(g))
    o = capi.Options(); E.cbfnmpc_default_options(C.byref(o))
    for k, v in (gains or {}).items(): setattr(g, k, v)
    for k, v in (options or {}).items(): setattr(o, k, v)
    B = len(insts)
    d = ext_soa(insts, N); out = ext_alloc_out(B, N)
    a = Args()
    a.B, a.N, a.T, a.K_obs = B, N, T, d["n_obs"]
    nb = None
    if fused_members > 1:
        assert T == 1 and B % fused_members == 0 and B * N <= 32
        a.n_nbr, a.xmode = fused_members - 1, 4
    elif T > 1:
        a.n_nbr, a.xmode = T - 1, 1
    elif nbr is not None:
        nbr = np.asarray(nbr, float); nn = nbr.shape[1]
        nb = np.zeros((nn, 2 * N + 2, B))
        for i in range(nn):
            for k in range(N):
                nb[i, 2 * k] = nbr[:, i, k, 0]; nb[i, 2 * k + 1] = nbr[:, i, k, 1]
        a.n_nbr, a.xmode, a.nbr = nn, 2, nb.ctypes.data
    else:
        a.n_nbr, a.xmode = 0, 0
    a.eps_obs = g.eps_dist if eps_obs is None else eps_obs; a.eps_nbr = eps_nbr; a.relax = relax; a.pub_tol = pub_tol; a.max_team_iters = max_team_iters; a.mu_stall = mu_stall
    for k, kk in (("x0", "x0"), ("xdes", "xdes"), ("feet", "feet_xy"), ("obs", "obs_xy"), ("abuf", "alpha_buf"), ("contact", "contact")):
        setattr(a, k, d[kk].ctypes.data)
    for k, kk in (("X", "X"), ("U", "U"), ("lam", "lam"), ("slack", "slack"), ("xqp", "x_qp"), ("alpha", "alpha"), ("abuf_next", "alpha_buf_next"),
                  ("kkt", "kkt"), ("pub", "pub"), ("status", "status"), ("qp_iters", "qp_iters"), ("nlp_iters", "nlp_iters"), ("team_iters", "team_iters")):
        setattr(a, k, out[kk].ctypes.data)
    rc = L.emu_ext_solve(C.byref(g), C.byref(o), C.byref(a))
    if rc != 0:
        raise RuntimeError("emu_ext failed rc=%d" % rc)
    return out
