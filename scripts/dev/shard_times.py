"""Dev tool: the 8 shards of bench.py's 8-GPU workload, one after the other on ONE GPU: step / lane-kernel time and the iteration
tails of every shard (what the max over ranks of an 8-GPU run is made of); per-instance counts to gpurun_out/shard_iters.npz."""
import sys, statistics, numpy as np, torch
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = 4096; W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
allinst = S.cfg3_instances(B * W, seed=20251018 + 3000)
dev = torch.device("cuda", 0)
s = m.BatchSolver(max_batch=B)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(dev)
dump = {}
for r in range(W):
    d = capi.soa_from_instances(allinst[r * B:(r + 1) * B]); K = int(d["n_obs"])
    tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
    tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
    tout.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
    tout["active"] = torch.zeros((2, B), dtype=torch.int64, device=dev)
    ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
    sm, km = [], []
    with torch.cuda.stream(stream):
        for i in range(13):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); s.solve_device(B, ip, op, K, mode=1, stream=stream.cuda_stream); e1.record(stream)
            torch.cuda.synchronize()
            if i >= 3: sm.append(e0.elapsed_time(e1)); km.append(s.last_kernel_ms())
    st = tout["status"][0].cpu().numpy(); qi = tout["qp_iters"][0].cpu().numpy(); ni = tout["nlp_iters"][0].cpu().numpy()
    dump["st%d" % r] = st; dump["qi%d" % r] = qi; dump["ni%d" % r] = ni
    print("shard %d step %.4f (min %.4f max %.4f) kernel %.4f | conv %.4f qp max %d nlp max %d top4 %s" % (
        r, statistics.mean(sm), min(sm), max(sm), statistics.mean(km), (st == 0).mean(), qi.max(), ni.max(), np.sort(ni)[-4:]))
np.savez_compressed("gpurun_out/shard_iters.npz", **dump)
