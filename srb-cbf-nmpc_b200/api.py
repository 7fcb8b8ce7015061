"""Product binding: loads srb-cbf-nmpc_b200/lib/libcbfnmpc.so (hand-written CUDA for sm_100a) through ctypes.

There is no CPU path.  Importing works anywhere (so that host-side logic can be tested), but creating a solver
raises unless the CUDA library is built and a CUDA device is visible."""
import ctypes as C
import os
import numpy as np
from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CBFNMPC_LIB", os.path.join(HERE, "lib", "libcbfnmpc.so"))  # override: tuning experiments only
INCLUDE = os.path.join(os.path.dirname(HERE), "include", "cbfnmpc.h")
_lib = None


class CbfNmpcError(RuntimeError):
    pass


def load():
    """dlopen the CUDA library and declare the prototypes of include/cbfnmpc.h.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CbfNmpcError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                           "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.cbfnmpc_version.restype = C.c_int
    L.cbfnmpc_last_error.restype = C.c_char_p
    L.cbfnmpc_default_gains.argtypes = [C.POINTER(capi.Gains)]
    L.cbfnmpc_default_options.argtypes = [C.POINTER(capi.Options)]
    L.cbfnmpc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
    L.cbfnmpc_destroy.argtypes = [C.c_void_p]
    L.cbfnmpc_set_gains.argtypes = [C.c_void_p, C.POINTER(capi.Gains)]
    L.cbfnmpc_set_options.argtypes = [C.c_void_p, C.POINTER(capi.Options)]
    L.cbfnmpc_solve_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(capi.In), C.POINTER(capi.Out), C.c_void_p]
    L.cbfnmpc_solve_batch_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(capi.In), C.POINTER(capi.Out)]
    L.cbfnmpc_pack_inputs.argtypes = [C.c_void_p, C.c_int, C.POINTER(capi.RawIn), C.POINTER(capi.In), C.c_void_p]
    L.cbfnmpc_team_states.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cbfnmpc_run_cycle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(capi.RawIn), C.c_void_p, C.POINTER(capi.Out), C.c_void_p]
    L.cbfnmpc_generate_reference.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    L.cbfnmpc_qp_dense_batch.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 11 + [C.c_void_p]
    L.cbfnmpc_closed_loop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(capi.RawIn), C.POINTER(capi.WorldState), C.c_double, C.c_void_p]
    L.cbfnmpc_launch_count.argtypes = [C.c_void_p]
    L.cbfnmpc_launch_count.restype = C.c_longlong
    L.cbfnmpc_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.cbfnmpc_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    L.cbfnmpc_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.cbfnmpc_solve_dense_nlptuple.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.cbfnmpc_ext_default_cfg.argtypes = [C.POINTER(capi.ExtCfg)]
    L.cbfnmpc_ext_solve_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(capi.ExtCfg), C.POINTER(capi.ExtIn), C.POINTER(capi.ExtOut), C.c_void_p]
    L.cbfnmpc_ext_solve_batch_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(capi.ExtCfg), C.POINTER(capi.ExtIn), C.POINTER(capi.ExtOut)]
    L.cbfnmpc_comm_unique_id.argtypes = [C.c_void_p]
    L.cbfnmpc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.cbfnmpc_comm_destroy.argtypes = [C.c_void_p]
    L.cbfnmpc_ext_solve_sharded.argtypes = [C.c_void_p, C.c_int, C.POINTER(capi.ExtCfg), C.POINTER(capi.ExtIn), C.POINTER(capi.ExtOut),
                                            C.POINTER(capi.ExtStats), C.c_void_p]
    L.cbfnmpc_ext_solve_fused.argtypes = [C.c_void_p, C.c_int, C.POINTER(capi.ExtCfg), C.POINTER(capi.ExtIn), C.POINTER(capi.ExtOut),
                                          C.POINTER(capi.ExtStats), C.c_void_p]
    L.cbfnmpc_pack_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6
    _lib = L
    return L


EXPORTS = ["cbfnmpc_version", "cbfnmpc_last_error", "cbfnmpc_default_gains", "cbfnmpc_default_options", "cbfnmpc_create",
           "cbfnmpc_destroy", "cbfnmpc_set_gains", "cbfnmpc_set_options", "cbfnmpc_solve_batch", "cbfnmpc_solve_batch_host",
           "cbfnmpc_launch_count", "cbfnmpc_last_kernel_ms", "cbfnmpc_set_kernel_timing", "cbfnmpc_solve_dense_nlptuple", "cbfnmpc_measure_fp64_peak",
           "cbfnmpc_pack_inputs", "cbfnmpc_run_cycle", "cbfnmpc_closed_loop", "cbfnmpc_generate_reference", "cbfnmpc_qp_dense_batch",
           "cbfnmpc_team_states", "cbfnmpc_pack_dense", "cbfnmpc_ext_default_cfg", "cbfnmpc_ext_solve_batch", "cbfnmpc_ext_solve_batch_host",
           "cbfnmpc_comm_unique_id", "cbfnmpc_comm_init", "cbfnmpc_comm_destroy", "cbfnmpc_ext_solve_sharded", "cbfnmpc_ext_solve_fused"]


def default_gains():
    g = capi.Gains(); load().cbfnmpc_default_gains(C.byref(g)); return g


def default_options():
    o = capi.Options(); load().cbfnmpc_default_options(C.byref(o)); return o


class BatchSolver:
    """Owns one cbfnmpc_ctx (device buffers + stream) on one GPU."""

    def __init__(self, device=0, max_batch=4096, max_obs=20, gains=None, options=None):
        self.L = load()
        self.ctx = C.c_void_p()
        rc = self.L.cbfnmpc_create(C.byref(self.ctx), device, max_batch, max_obs)
        if rc != 0:
            raise CbfNmpcError("cbfnmpc_create failed (%d): %s" % (rc, self.L.cbfnmpc_last_error().decode()))
        self.device, self.max_batch, self.max_obs = device, max_batch, max_obs
        self.gains, self.options = default_gains(), default_options()
        if gains: self.set_gains(**gains)
        if options: self.set_options(**options)

    def _ck(self, rc):
        if rc != 0:
            raise CbfNmpcError("cbfnmpc call failed (%d): %s" % (rc, self.L.cbfnmpc_last_error().decode()))

    def set_gains(self, **kw):
        for k, v in kw.items(): setattr(self.gains, k, v)
        self._ck(self.L.cbfnmpc_set_gains(self.ctx, C.byref(self.gains)))

    def set_options(self, **kw):
        for k, v in kw.items(): setattr(self.options, k, v)
        self._ck(self.L.cbfnmpc_set_options(self.ctx, C.byref(self.options)))

    def close(self):
        if self.ctx:
            self.L.cbfnmpc_destroy(self.ctx); self.ctx = C.c_void_p()

    def __del__(self):
        try: self.close()
        except Exception: pass

    # ---- host buffers (numpy, ideally views of one pinned block): copies inside the call
    def solve_host(self, d, mode=capi.MODE_QP_NLP, out=None):
        B = d["x0"].shape[1]
        out = out if out is not None else capi.alloc_out(B)
        si, so = capi.in_struct(d), capi.out_struct(out)
        self._ck(self.L.cbfnmpc_solve_batch_host(self.ctx, B, mode, C.byref(si), C.byref(so)))
        return out

    # ---- device buffers: dicts of raw device pointers (ints), e.g. torch tensors' .data_ptr()
    def solve_device(self, B, in_ptrs, out_ptrs, n_obs, mode=capi.MODE_QP_NLP, stream=None):
        si = capi.In(); so = capi.Out()
        for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf"):
            setattr(si, k, in_ptrs.get(k))
        si.n_obs = n_obs
        for k, _ in capi.Out._fields_:
            setattr(so, k, out_ptrs.get(k))
        self._ck(self.L.cbfnmpc_solve_batch(self.ctx, B, mode, C.byref(si), C.byref(so), stream))

    # ---- raw per-cycle inputs (what MPC_dist::updateState receives), device pointers
    @staticmethod
    def _raw_struct(raw):
        r = capi.RawIn()
        for k, _ in capi.RawIn._fields_:
            if k in ("state_other", "n_other"):          # optional: the reference's disabled inter-agent rule
                setattr(r, k, raw.get(k) or 0)
            else:
                setattr(r, k, raw[k])
        return r

    def pack_inputs(self, B, raw, packed_ptrs, stream=None):
        """updateState + copPlanner_eventbase + footholdsPlanner for B agents on the device (cbfnmpc_pack_inputs)."""
        si = capi.In()
        for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy"):
            setattr(si, k, packed_ptrs[k])
        si.n_obs = raw["n_obs"] + (raw.get("n_other", 0) if raw.get("state_other") else 0)
        r = self._raw_struct(raw)
        self._ck(self.L.cbfnmpc_pack_inputs(self.ctx, B, C.byref(r), C.byref(si), stream))

    def team_states(self, B, team_size, q_ptr, dq_ptr, state_other_ptr, stream=None):
        """state_other [4*(team_size-1)][B] of agents stored team by team (cbfnmpc_team_states)."""
        self._ck(self.L.cbfnmpc_team_states(self.ctx, B, team_size, q_ptr, dq_ptr, state_other_ptr, stream))

    def run_cycle(self, B, raw, alpha_buf_ptr, out_ptrs, mode=capi.MODE_QP_NLP, stream=None):
        """updateState + run_NMPC from the raw inputs: pack kernel, class sort, solve kernel (cbfnmpc_run_cycle)."""
        so = capi.Out()
        for k, _ in capi.Out._fields_:
            setattr(so, k, out_ptrs.get(k))
        r = self._raw_struct(raw)
        self._ck(self.L.cbfnmpc_run_cycle(self.ctx, B, mode, C.byref(r), alpha_buf_ptr, C.byref(so), stream))

    def generate_reference(self, n_teams, pstart_ptr, pobs_ptr, n_obs, pr_ptr, prd_ptr, goal=(10.0, 0.0), loop=100000, stream=None):
        """generateReferenceTrajectory for n_teams teams on the device (cbfnmpc_generate_reference)."""
        self._ck(self.L.cbfnmpc_generate_reference(self.ctx, n_teams, pstart_ptr, pobs_ptr, n_obs, goal[0], goal[1], loop, pr_ptr, prd_ptr, stream))

    def qp_dense_batch(self, B, n, p, m, ptrs, stream=None):
        """B dense QPs (iSWIFT's algorithm, one warp each; cbfnmpc_qp_dense_batch).  ptrs: device pointers P, c, A, b, G, h, x, y, z,
        status, iters (A, b, y, z may be None)."""
        a = [ptrs.get(k) for k in ("P", "c", "A", "b", "G", "h", "x", "y", "z", "status", "iters")]
        self._ck(self.L.cbfnmpc_qp_dense_batch(self.ctx, B, n, p, m, *a, stream))

    def closed_loop(self, B, n_cycles, raw, world, mode=capi.MODE_QP_NLP, fail_radius=0.5, stream=None):
        """n_cycles receding-horizon cycles of B worlds, entirely on the device (cbfnmpc_closed_loop).  `world`: dict of
        device pointers (q, dq, contact, toe_pos, gait_domain must be the arrays of `raw`)."""
        w = capi.WorldState()
        for k, _ in capi.WorldState._fields_:
            setattr(w, k, world.get(k))
        r = self._raw_struct(raw)
        self._ck(self.L.cbfnmpc_closed_loop(self.ctx, B, n_cycles, mode, C.byref(r), C.byref(w), fail_radius, stream))

    # ---- extended mode (include/cbfnmpc.h "EXTENDED MODE"; no reference implementation exists)
    def ext_cfg(self, N=4, team_size=1, **kw):
        cfg = capi.ExtCfg(); self.L.cbfnmpc_ext_default_cfg(C.byref(cfg))
        cfg.N, cfg.team_size = N, team_size
        for k, v in kw.items(): setattr(cfg, k, v)
        return cfg

    def ext_solve_host(self, d, N=4, team_size=1, nbr_traj=None, out=None, **cfgkw):
        """d: capi.ext_soa_from_instances(...); nbr_traj: optional (n_nbr, 2N, B) fixed neighbour trajectories (team_size 1)."""
        B = d["x0"].shape[1]
        out = out if out is not None else capi.ext_alloc_out(B, N)
        cfg = self.ext_cfg(N, team_size, **cfgkw)
        si = capi.ExtIn()
        for k in ("x0", "xdes", "feet_xy", "obs_xy", "alpha_buf"):
            capi._check(k, d[k], np.float64, d[k].shape[0], B); setattr(si, k, d[k].ctypes.data)
        capi._check("contact", d["contact"], np.int32, 4, B); si.contact = d["contact"].ctypes.data
        si.n_obs = int(d["n_obs"])
        if nbr_traj is not None:
            nbr_traj = np.ascontiguousarray(nbr_traj, dtype=np.float64)
            capi._check("nbr_traj", nbr_traj, np.float64, nbr_traj.shape[0] * 2 * N, B)
            si.nbr_traj = nbr_traj.ctypes.data; si.n_nbr = nbr_traj.shape[0]
        so = capi.ExtOut()
        rows = capi.ext_out_rows(N)
        for k, v in out.items():
            if v is None: continue
            capi._check(k, v, np.int32 if k in capi.EXT_OUT_I32 else np.float64, 1 if k in capi.EXT_OUT_I32 else rows[k], B)
            setattr(so, k, v.ctypes.data)
        self._ck(self.L.cbfnmpc_ext_solve_batch_host(self.ctx, B, C.byref(cfg), C.byref(si), C.byref(so)))
        return out

    def _ext_structs(self, in_ptrs, out_ptrs, n_obs, n_nbr=0):
        si = capi.ExtIn(); so = capi.ExtOut()
        for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf", "nbr_traj"):
            setattr(si, k, in_ptrs.get(k))
        si.n_obs, si.n_nbr = n_obs, n_nbr
        for k, _ in capi.ExtOut._fields_:
            setattr(so, k, out_ptrs.get(k))
        return si, so

    def ext_solve_device(self, B, in_ptrs, out_ptrs, n_obs, N=4, team_size=1, n_nbr=0, stream=None, **cfgkw):
        cfg = self.ext_cfg(N, team_size, **cfgkw)
        si, so = self._ext_structs(in_ptrs, out_ptrs, n_obs, n_nbr)
        self._ck(self.L.cbfnmpc_ext_solve_batch(self.ctx, B, C.byref(cfg), C.byref(si), C.byref(so), stream))

    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = load().cbfnmpc_comm_unique_id(buf)
        if rc != 0:
            raise CbfNmpcError("cbfnmpc_comm_unique_id failed (%d): %s" % (rc, load().cbfnmpc_last_error().decode()))
        return bytes(buf)

    def comm_init(self, nranks, rank, uid):
        self._ck(self.L.cbfnmpc_comm_init(self.ctx, nranks, rank, C.c_char_p(uid)))

    def comm_destroy(self):
        self._ck(self.L.cbfnmpc_comm_destroy(self.ctx))

    def ext_solve_sharded(self, n_teams, in_ptrs, out_ptrs, n_obs, N=4, team_size=4, stream=None, **cfgkw):
        """Agent-sharded teams (member-major layout, device pointers); returns the cbfnmpc_ext_stats of the call."""
        cfg = self.ext_cfg(N, team_size, **cfgkw)
        si, so = self._ext_structs(in_ptrs, out_ptrs, n_obs)
        st = capi.ExtStats()
        self._ck(self.L.cbfnmpc_ext_solve_sharded(self.ctx, n_teams, C.byref(cfg), C.byref(si), C.byref(so), C.byref(st), stream))
        return dict(launches=st.launches, iterations=st.iterations, exchanges=st.exchanges, nranks=st.nranks,
                    bytes_per_exchange=st.bytes_per_exchange, total_ms=st.total_ms, fused=st.fused)

    def ext_solve_fused(self, n_teams, in_ptrs, out_ptrs, n_obs, N=4, team_size=4, stream=None, **cfgkw):
        """Agent-sharded teams with the exchange inside ONE launch (peer-memory mailboxes); same layout and answers as ext_solve_sharded,
        which it falls back to where the launch cannot be resident at once or the ranks cannot map each other's memory (stats['fused'])."""
        cfg = self.ext_cfg(N, team_size, **cfgkw)
        si, so = self._ext_structs(in_ptrs, out_ptrs, n_obs)
        st = capi.ExtStats()
        self._ck(self.L.cbfnmpc_ext_solve_fused(self.ctx, n_teams, C.byref(cfg), C.byref(si), C.byref(so), C.byref(st), stream))
        return dict(launches=st.launches, iterations=st.iterations, exchanges=st.exchanges, nranks=st.nranks,
                    bytes_per_exchange=st.bytes_per_exchange, total_ms=st.total_ms, fused=st.fused)

    def launch_count(self):
        return int(self.L.cbfnmpc_launch_count(self.ctx))

    def set_kernel_timing(self, on):
        """Event pair around the solve kernel on / off (off: the solve kernel becomes a programmatic dependent of the classify pre-pass)."""
        self._ck(self.L.cbfnmpc_set_kernel_timing(self.ctx, 1 if on else 0))

    def last_kernel_ms(self):
        ms = C.c_float(0); self._ck(self.L.cbfnmpc_last_kernel_ms(self.ctx, C.byref(ms))); return ms.value

    def fp64_peak_tflops(self):
        t = C.c_double(0); self._ck(self.L.cbfnmpc_measure_fp64_peak(self.ctx, C.byref(t))); return t.value

    def pack_dense(self, x0, xdes, F):
        """The dense nlpTuple matrices of MPC_dist.cpp:164-321 for one agent (host code; F = 2 x c stance matrix)."""
        F = np.asarray(F, float); c = F.shape[1]
        nv, p, mi = 25 + 4 * c, 28, 60 + 8 * c
        F4 = np.zeros((2, 4)); F4[:, :c] = F
        x0 = np.ascontiguousarray(x0, dtype=np.float64); xdes = np.ascontiguousarray(xdes, dtype=np.float64)
        Q = np.zeros((nv, nv), order="F"); f = np.zeros(nv); A = np.zeros((p, nv), order="F"); b = np.zeros(p)
        G = np.zeros((mi, nv), order="F"); h = np.zeros(mi)
        self._ck(self.L.cbfnmpc_pack_dense(self.ctx, x0.ctypes.data, xdes.ctypes.data, F4.ctypes.data, c, Q.ctypes.data, f.ctypes.data,
                                           A.ctypes.data, b.ctypes.data, G.ctypes.data, h.ctypes.data))
        return Q, f, A, b, G, h

    def solve_dense_nlptuple(self, Q, f, Aeq, beq, G, h, x0_start, obs, mode_nlp=True):
        """Reference layout: Eigen column-major matrices of pack_nlp_data.h:11-29."""
        nv, p, mi = Q.shape[0], Aeq.shape[0], G.shape[0]
        cm = lambda a: np.asfortranarray(a, dtype=np.float64)
        Qc, Ac, Gc = cm(Q), cm(Aeq), cm(G)
        fv, bv, hv = (np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (f, beq, h))
        xs = np.ascontiguousarray(x0_start, dtype=np.float64); ob = np.ascontiguousarray(obs, dtype=np.float64)
        xo = np.zeros(nv); st = C.c_int(0); kk = C.c_double(0)
        rc = self.L.cbfnmpc_solve_dense_nlptuple(self.ctx, nv, p, mi, Qc.ctypes.data, fv.ctypes.data, Ac.ctypes.data, bv.ctypes.data,
                                                 Gc.ctypes.data, hv.ctypes.data, xs.ctypes.data, ob.ctypes.data, int(mode_nlp),
                                                 xo.ctypes.data, C.byref(st), C.byref(kk))
        if rc != 0:
            e = CbfNmpcError("cbfnmpc_solve_dense_nlptuple failed (%d): %s" % (rc, self.L.cbfnmpc_last_error().decode()))
            e.code = rc
            raise e
        return xo, st.value, kk.value
