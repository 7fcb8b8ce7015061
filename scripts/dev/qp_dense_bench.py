"""Dev tool: throughput of cbfnmpc_qp_dense_batch on low-level-controller-size QPs (31 x 18 x 45)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import srb_cbf_nmpc_b200 as m
import qp_cases
dev = torch.device("cuda", 0)
base = qp_cases.lowlevel_like_qps(256, seed=5)
for B in (256, 4096, 32768):
    qps = [base[i % 256] for i in range(B)]
    cm = lambda k: torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(q[k], float).T for q in qps]))).to(dev)
    vec = lambda k: torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(q[k], float) for q in qps]))).to(dev)
    n, p, mm = 31, 18, 45
    t = dict(P=cm(0), c=vec(1), A=cm(2), b=vec(3), G=cm(4), h=vec(5), x=torch.zeros((B, n), dtype=torch.float64, device=dev),
             status=torch.zeros(B, dtype=torch.int32, device=dev), iters=torch.zeros(B, dtype=torch.int32, device=dev))
    s = m.BatchSolver(max_batch=8)
    ptr = {k: v.data_ptr() for k, v in t.items()}
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.qp_dense_batch(B, n, p, mm, ptr); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("B %6d: %.3f ms, %.2f M QP/s, converged %d, mean iterations %.2f" % (B, dt * 1e3, B / dt / 1e6, int((t["status"] == 0).sum()), t["iters"].double().mean().item()))
    s.close()
