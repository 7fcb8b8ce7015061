"""DEV/TEST TOOL: loads the host emulation of the kernel source (tests/emu/libcbfnmpc_lane_emu.so: the product's
cbfnmpc_lane.h compiled for the host, each 4-lane group on 4 cooperative fibers).  Never imported by the product package.

Two builds: the default -O2 one, and a "sitecheck" build (-O0 -DCBF_EMU_SITECHECK) in which every source-level shuffle site
has its own call address and the harness verifies that the 4 lanes of a group always meet at the SAME site -- the
convergence requirement of __shfl_sync with the group mask on the device."""
import ctypes as C
import importlib.util
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_spec = importlib.util.spec_from_file_location("cbf_capi", os.path.join(ROOT, "srb-cbf-nmpc_b200", "capi.py"))
capi = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(capi)

_libs = {}


def build(force=False, sitecheck=False):
    lib = os.path.join(HERE, "libcbfnmpc_lane_emu_sitecheck.so" if sitecheck else "libcbfnmpc_lane_emu.so")
    csrc = os.path.join(ROOT, "srb-cbf-nmpc_b200", "csrc")
    srcs = [os.path.join(HERE, "emu_lane.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    srcs.append(os.path.join(ROOT, "include", "cbfnmpc.h"))
    if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        flags = ["-O0", "-DCBF_EMU_SITECHECK"] if sitecheck else ["-O2"]
        subprocess.check_call(["g++"] + flags + ["-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", lib,
                               os.path.join(HERE, "emu_lane.cpp"), "-ldl"])
    return lib


def lib(sitecheck=False):
    if sitecheck not in _libs:
        L = C.CDLL(build(sitecheck=sitecheck))
        L.emu_solve_batch_host.argtypes = [C.POINTER(capi.Gains), C.POINTER(capi.Options), C.c_int, C.c_int,
                                           C.POINTER(capi.In), C.POINTER(capi.Out), C.c_void_p]
        _libs[sitecheck] = L
    return _libs[sitecheck]


def solve(insts, mode=1, gains=None, options=None, xstart=None, sitecheck=False, late=False):
    """late: the instantiation of cbfnmpc_lane_kernel_late (obstacles looked at after the QP stage: the host-buffer call's kernel)"""
    L = lib(sitecheck)
    L.emu_set_late_obstacles(1 if late else 0)
    g = capi.Gains(); L.cbfnmpc_default_gains(C.byref(g))
    o = capi.Options(); L.cbfnmpc_default_options(C.byref(o))
    for k, v in (gains or {}).items(): setattr(g, k, v)
    for k, v in (options or {}).items(): setattr(o, k, v)
    d = capi.soa_from_instances(insts)
    out = capi.alloc_out(len(insts))
    si, so = capi.in_struct(d), capi.out_struct(out)
    xs = None
    if xstart is not None:
        xs = np.ascontiguousarray(xstart, dtype=np.float64)
    rc = L.emu_solve_batch_host(C.byref(g), C.byref(o), len(insts), mode, C.byref(si), C.byref(so), xs.ctypes.data if xs is not None else None)
    if rc != 0:
        raise RuntimeError("emu failed rc=%d" % rc)
    return out
