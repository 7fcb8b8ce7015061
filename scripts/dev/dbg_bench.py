"""Dev tool: one instance of bench.py's global workload solved alone with the CBF_DBGBUF build (CBFNMPC_LIB=build/libcbfnmpc_dbg.so):
per-iteration log of group 0 (it, mu, E0, e_d, e_p, e_c, theta | dw, alpha_p max, alpha_d max | alpha, phi, dphi)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
i = int(sys.argv[1])
inst = S.cfg3_instances(4096 * 8, seed=20251018 + 3000)[i]
s = m.BatchSolver(max_batch=8)
L = m.load()
buf = np.zeros(8192)
L.cbfnmpc_debug_read.argtypes = [C.c_void_p, C.c_int]
L.cbfnmpc_debug_read(buf.ctypes.data, 1)
o = s.solve_host(capi.soa_from_instances([inst]), mode=1)
n = L.cbfnmpc_debug_read(buf.ctypes.data, 1)
print("status", o["status"][0], "nlp", o["nlp_iters"][0], "kkt", o["kkt"][0], "log entries", n)
v = buf[:n]; k = 0
while k < n:
    if v[k] >= 1000.0:
        print("it %2d mu %.2e E0 %.3e ed %.2e ep %.2e ec %.2e th %.1e" % ((int(v[k] - 1000),) + tuple(v[k + 1:k + 7]))); k += 7
        while k < n and v[k] <= -3000.0:
            print("     lane %d: e_d parts X %.2e U %.2e lambda %.2e slack %.2e" % ((int(-v[k] - 3000),) + tuple(v[k + 1:k + 5]))); k += 5
        if k + 3 <= n and v[k] < 1000.0:
            print("     dw %.1e apmax %.3e admax %.3e" % tuple(v[k:k + 3])); k += 3
            while k < n and -3000.0 < v[k] <= -2000.0:
                print("     lane %d: max -dz/z %.3e at slot %d: de %.3e slo %.3e shi %.3e zlo %.3e zhi %.3e" % ((int(-v[k] - 2000), v[k + 1], int(v[k + 2])) + tuple(v[k + 3:k + 8]))); k += 8
            if k + 3 <= n and v[k] < 1000.0:
                print("     alpha %.3e ph %.10e dphi %.2e" % tuple(v[k:k + 3])); k += 3
    else:
        k += 1
