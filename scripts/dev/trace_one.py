"""Dev tool: solve one cfg-3 instance (B = 1) with the tracing build of the library (CBFNMPC_LIB=.../libcbfnmpc_trace.so)."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
i = int(sys.argv[1])
inst = S.cfg3_instances(4096)[i]
s = m.BatchSolver(max_batch=8)
o = s.solve_host(capi.soa_from_instances([inst]), mode=1)
import torch; torch.cuda.synchronize()
print("status", o["status"][0], "nlp", o["nlp_iters"][0])
