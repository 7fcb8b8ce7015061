"""Dev tool: kernel time of one cfg-3 batch of B instances (device-resident inputs, one launch per call) and the host call."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')  # run from the repo root
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
MODES = [int(a) for a in sys.argv[2:]] or [0, 1]  # 0 = QP only, 1 = QP -> NLP
insts = S.cfg3_instances(B)
d = capi.soa_from_instances(insts); K = int(d["n_obs"])
s = m.BatchSolver(max_batch=B)
dev = torch.device("cuda", 0)
tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
tout.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
tout["active"] = torch.zeros((2, B), dtype=torch.int64, device=dev)
ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
for mode in MODES:
    for it in range(4):
        s.solve_device(B, ip, op, K, mode=mode); torch.cuda.synchronize()
        km = s.last_kernel_ms()
        t = time.time(); out = s.solve_host(d, mode=mode); dt = time.time() - t
        st = tout["status"][0].cpu().numpy()
        print("mode", mode, "B", B, "host call %.3f ms kernel %.3f ms" % (dt * 1e3, km), "status", np.bincount(st, minlength=5),
              "iters qp %.2f nlp %.2f nfact %.2f" % (tout['qp_iters'].double().mean().item(), tout['nlp_iters'].double().mean().item(), tout['n_fact'].double().mean().item()))
print("fp64 peak TFLOP/s", s.fp64_peak_tflops())
