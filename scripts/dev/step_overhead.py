"""Dev tool: where does a step's time go beyond the solve kernel?"""
import sys, numpy as np, torch, statistics
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = 4096
d = capi.soa_from_instances(S.cfg3_instances(B)); K = int(d["n_obs"])
dev = torch.device("cuda", 0)
s = m.BatchSolver(max_batch=B, max_obs=K)
tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
tout.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
tout["active"] = torch.zeros((2, B), dtype=torch.int64, device=dev)
ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
st = torch.cuda.Stream(dev); flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for do_flush in (True, False):
    ev = []; km = []
    with torch.cuda.stream(st):
        for i in range(25):
            if do_flush: flush.fill_(1)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(st); s.solve_device(B, ip, op, K, mode=1, stream=st.cuda_stream); b.record(st); ev.append((a, b))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev][5:]
    print("flush", do_flush, "step ms mean %.4f min %.4f" % (statistics.mean(t), min(t)), "last kernel ms %.4f" % s.last_kernel_ms())
