import sys, time, numpy as np
sys.path.insert(0, '.')  # run from the repo root
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
insts = S.cfg3_instances(B)
d = capi.soa_from_instances(insts)
s = m.BatchSolver(max_batch=B)
for mode in (0, 1):
    for it in range(4):
        t = time.time(); out = s.solve_host(d, mode=mode); dt = time.time() - t
        print("mode", mode, "B", B, "host call %.3f ms kernel %.3f ms" % (dt * 1e3, s.last_kernel_ms()), "status", np.bincount(out['status'][0], minlength=5),
              "iters qp %.2f nlp %.2f nfact %.2f" % (out['qp_iters'].mean(), out['nlp_iters'].mean(), out['n_fact'].mean()))
print("fp64 peak TFLOP/s", s.fp64_peak_tflops())
