"""Dev tool: one instance of bench.py's global workload (index into the 8 x 4096 set) solved alone; with the tracing build of the
library (CBFNMPC_LIB=build/libcbfnmpc_trace.so, nvcc -DCBF_TRACE) group 0 narrates its iterations."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
i = int(sys.argv[1])
inst = S.cfg3_instances(4096 * 8, seed=20251018 + 3000)[i]
s = m.BatchSolver(max_batch=8)
o = s.solve_host(capi.soa_from_instances([inst]), mode=1)
import torch; torch.cuda.synchronize()
print("status", o["status"][0], "qp", o["qp_iters"][0], "nlp", o["nlp_iters"][0], "kkt", o["kkt"][0])
