"""Dev tool: solve one cfg-3 instance with the CBF_DBGBUF build and dump the per-iteration log."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
i = int(sys.argv[1])
inst = S.cfg3_instances(32768)[i]
s = m.BatchSolver(max_batch=8)
L = m.load()
buf = np.zeros(8192)
L.cbfnmpc_debug_read.argtypes = [C.c_void_p, C.c_int]
L.cbfnmpc_debug_read(buf.ctypes.data, 1)
o = s.solve_host(capi.soa_from_instances([inst]), mode=1)
n = L.cbfnmpc_debug_read(buf.ctypes.data, 1)
print("status", o["status"][0], "nlp", o["nlp_iters"][0], "log entries", n)
np.save("gpurun_out/dbg_%d.npy" % i, buf[:n])
