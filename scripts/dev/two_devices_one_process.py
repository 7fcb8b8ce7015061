"""Dev tool (needs 2 GPUs): one process, one context per device; both must give the same bits, dense QP included."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
import qp_cases
insts = S.cfg3_instances(300)
d = capi.soa_from_instances(insts)
outs = []
for dev in (0, 1):
    s = m.BatchSolver(device=dev, max_batch=512, max_obs=20)
    outs.append(s.solve_host(d, mode=capi.MODE_QP_NLP))
    # dense QP on this device
    td = torch.device("cuda", dev)
    qps = qp_cases.lowlevel_like_qps(64, seed=3)
    cm = lambda k: torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(q[k], float).T for q in qps]))).to(td)
    vec = lambda k: torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(q[k], float) for q in qps]))).to(td)
    t = dict(P=cm(0), c=vec(1), A=cm(2), b=vec(3), G=cm(4), h=vec(5), x=torch.zeros((64, 31), dtype=torch.float64, device=td),
             status=torch.zeros(64, dtype=torch.int32, device=td), iters=torch.zeros(64, dtype=torch.int32, device=td))
    s.qp_dense_batch(64, 31, 18, 45, {k: v.data_ptr() for k, v in t.items()})
    torch.cuda.synchronize(td)
    outs[-1]["qpx"] = t["x"].cpu().numpy(); outs[-1]["qpst"] = t["status"].cpu().numpy()
    s.close()
for k in outs[0]:
    assert np.array_equal(outs[0][k], outs[1][k]), k
print("two devices, one process: identical answers on cuda:0 and cuda:1 (%d MPC instances, 64 dense QPs, all QP statuses %s)" % (len(insts), set(outs[0]["qpst"].tolist())))
