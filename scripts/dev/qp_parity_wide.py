"""QP-mode parity of the CUDA path against the oracle AND the reference iSWIFT core on a wide fuzz (run on the GPU box)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import helpers as H, oracle as O
import srb_cbf_nmpc_b200 as pkg
from srb_cbf_nmpc_b200 import capi
N = 8192
insts = H.wide_fuzz_instances(N, seed=4242)
s = pkg.BatchSolver(device=0, max_batch=N, max_obs=20)
outq = s.solve_host(capi.soa_from_instances(insts), mode=capi.MODE_QP)
refq = H.oracle_batch(insts, mode=0, nthreads=16)
st, rs = outq['status'][0], refq['status']
print('wide fuzz', N, 'QP mode: status mismatches vs oracle', int((st != rs).sum()), np.nonzero(st != rs)[0][:10])
both = (st == 0) & (rs == 0)
qi, rq = outq['qp_iters'][0], refq['qp_iters']
print('both converged', int(both.sum()), 'iteration-count mismatches among them', int((qi[both] != rq[both]).sum()), 'max |dX|', np.abs(outq['X'].T - refq['X'])[both].max())
# cfg-3 distribution, 32768 instances
S = __import__('srb_cbf_nmpc_b200.scenarios', fromlist=['x'])
insts = S.cfg3_instances(32768, seed=11)
s2 = pkg.BatchSolver(device=0, max_batch=32768, max_obs=20)
outq = s2.solve_host(capi.soa_from_instances(insts), mode=capi.MODE_QP)
refq = H.oracle_batch(insts, mode=0, nthreads=16)
st, rs = outq['status'][0], refq['status']
both = (st == 0) & (rs == 0)
print('cfg3 32768 QP mode: status mismatches', int((st != rs).sum()), 'iteration-count mismatches', int((outq['qp_iters'][0][both] != refq['qp_iters'][both]).sum()), 'max |dX|', np.abs(outq['X'].T - refq['X'])[both].max())
