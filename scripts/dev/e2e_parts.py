"""Dev tool: what the host-buffer call of 4096 cfg-3 instances is made of -- the call itself for 1 / 2 / 4 column ranges (CBFNMPC_CHUNKS),
the bare copies of the same pinned buffers, and the device-resident step."""
import os, sys, time, statistics, subprocess, numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":   # usage: e2e_parts.py [B ...]
    import torch
    sys.path.insert(0, '.')
    import srb_cbf_nmpc_b200 as m
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    B = int(os.environ.get("E2E_B", "4096"))
    d = capi.soa_from_instances(S.cfg3_instances(B)); K = int(d["n_obs"])
    s = m.BatchSolver(max_batch=B); s.set_kernel_timing(False)
    n_in = 32 + 2 * K
    hin = torch.empty((n_in, B), dtype=torch.float64).pin_memory(); hct = torch.empty((4, B), dtype=torch.int32).pin_memory()
    hout = torch.empty((50, B), dtype=torch.float64).pin_memory(); houti = torch.empty((3, B), dtype=torch.int32).pin_memory()
    hn, ho, hi = hin.numpy(), hout.numpy(), houti.numpy()
    dh = {"x0": hn[0:4], "xdes": hn[4:20], "feet_xy": hn[20:28], "alpha_buf": hn[28:32], "obs_xy": hn[32:32 + 2 * K], "contact": hct.numpy(), "n_obs": K}
    for k in ("x0", "xdes", "feet_xy", "alpha_buf", "obs_xy", "contact"): dh[k][...] = d[k]
    oh = {"X": ho[0:16], "U": ho[16:24], "slack": ho[24:25], "alpha": ho[25:45], "alpha_buf_next": ho[45:49], "kkt": ho[49:50],
          "status": hi[0:1], "qp_iters": hi[1:2], "nlp_iters": hi[2:3], "lam": None, "x_qp": None, "n_fact": None, "active": None}
    for _ in range(5): s.solve_host(dh, mode=1, out=oh)
    ts = []
    for _ in range(40):
        t = time.perf_counter(); s.solve_host(dh, mode=1, out=oh); ts.append(time.perf_counter() - t)
    print("B %d chunks %s: host call median %.4f ms min %.4f" % (B, os.environ.get("CBFNMPC_CHUNKS", "default"), 1e3 * statistics.median(ts), 1e3 * min(ts)))
    if os.environ.get("CBFNMPC_CHUNKS") == "1":
        dev = torch.device("cuda", 0)
        din = torch.empty((n_in, B), dtype=torch.float64, device=dev); dct = torch.empty((4, B), dtype=torch.int32, device=dev)
        dout = torch.empty((50, B), dtype=torch.float64, device=dev); douti = torch.empty((3, B), dtype=torch.int32, device=dev)
        for name, fn in (("H2D 2.4 MB", lambda: (din.copy_(hin, non_blocking=True), dct.copy_(hct, non_blocking=True))),
                         ("D2H 1.7 MB", lambda: (hout.copy_(dout, non_blocking=True), houti.copy_(douti, non_blocking=True)))):
            for _ in range(5): fn(); torch.cuda.synchronize()
            tt = []
            for _ in range(40):
                t = time.perf_counter(); fn(); torch.cuda.synchronize(); tt.append(time.perf_counter() - t)
            print("  bare %s: median %.4f ms" % (name, 1e3 * statistics.median(tt)))
    sys.exit(0)
for B in (sys.argv[1:] or ["4096"]):
    for ch in ("1", "2", "4"):
        env = dict(os.environ); env["CBFNMPC_CHUNKS"] = ch; env["E2E_B"] = B
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
