"""Dev tool: per-warp durations of one cbfnmpc_lane_kernel launch (library built with -DCBF_CTATIME, load it with CBFNMPC_LIB=...):
which CTA sets the time of the launch, and what its 16 instances look like.  usage: cta_times.py <shard> [n_shards]"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = 4096; r = int(sys.argv[1]); W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
insts = S.cfg3_instances(B * W, seed=20251018 + 3000)[r * B:(r + 1) * B]
d = capi.soa_from_instances(insts)
s = m.BatchSolver(max_batch=B)
L = m.load()
L.cbfnmpc_debug_ctatime.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda", 0)
tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
tout.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
tout["active"] = torch.zeros((2, B), dtype=torch.int64, device=dev)
ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
FLUSH = len(sys.argv) > 3 and sys.argv[3] == "flush"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(dev)
for rep in range(4):   # ONE launch over the whole shard (the host-buffer call would cut it into 4 column ranges)
    with torch.cuda.stream(stream):   # flush and solve on ONE stream (the library's own stream does not wait for torch's)
        if FLUSH: flush.fill_(1)
        s.solve_device(B, ip, op, int(d["n_obs"]), mode=1, stream=stream.cuda_stream)
    torch.cuda.synchronize()
out = {k: v.cpu().numpy() for k, v in tout.items()}
n_cta = B // 16 + 1
buf = np.zeros(3 * 2 * n_cta, np.int64)
L.cbfnmpc_debug_ctatime(buf.ctypes.data, buf.size)
t = buf.reshape(n_cta, 2, 3)
dur = t[:, :, 0].max(1) * 1e-3  # us
item = t[:, 0, 1]; sm = t[:, 0, 2]
c = d["contact"].sum(0)
order = np.concatenate([np.nonzero(c == 4)[0], np.nonzero(c == 2)[0]])   # class lists are stable
n4 = int((c == 4).sum()); c4 = (n4 + 15) // 16
st = out["status"][0]; qi = out["qp_iters"][0]; ni = out["nlp_iters"][0]; nf = out["n_fact"][0]
print("kernel %.4f ms | CTA durations us: mean %.1f p50 %.1f p90 %.1f max %.1f" % (s.last_kernel_ms(), dur.mean(), np.median(dur), np.percentile(dur, 90), dur.max()))
for k in np.argsort(-dur)[:6]:
    it = int(item[k])
    if it < c4: ids = order[it * 16: min(it * 16 + 16, n4)]
    else: ids = order[n4 + (it - c4) * 16: n4 + (it - c4) * 16 + 16]
    print("cta %d item %d sm %d dur %.1f us warps %s | inst %s\n   st %s\n   qp %s\n   nlp %s\n   nfact %s" % (k, it, sm[k], dur[k], t[k, :, 0] * 1e-3, ids, st[ids], qi[ids], ni[ids], nf[ids]))
np.savez_compressed("gpurun_out/cta_times_%d%s.npz" % (r, "f" if FLUSH else ""), t=t, st=st, qi=qi, ni=ni, nf=nf)
