"""Dev tool: one instance of tests/helpers.wide_fuzz_instances(2048, seed) (B = 1), e.g. with the tracing build (CBFNMPC_LIB=...)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi
import helpers as H
i = int(sys.argv[1]); seed = int(sys.argv[2]) if len(sys.argv) > 2 else 77
inst = H.wide_fuzz_instances(2048, seed=seed)[i]
s = m.BatchSolver(max_batch=8)
o = s.solve_host(capi.soa_from_instances([inst]), mode=1)
import torch; torch.cuda.synchronize()
print("status", o["status"][0], "qp", o["qp_iters"][0], "nlp", o["nlp_iters"][0], "kkt", o["kkt"][0])
