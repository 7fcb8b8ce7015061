"""Dev tool: (a) bare pitched copies of column ranges (what a chunked host-buffer call issues), (b) four quarter-batch solves on four streams."""
import sys, time, statistics, numpy as np, torch
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = 4096; dev = torch.device("cuda", 0)
hin = torch.empty((72, B), dtype=torch.float64).pin_memory(); din = torch.empty((72, B), dtype=torch.float64, device=dev)
hout = torch.empty((50, B), dtype=torch.float64).pin_memory(); dout = torch.empty((50, B), dtype=torch.float64, device=dev)
def timeit(fn, n=40):
    for _ in range(5): fn(); torch.cuda.synchronize()
    tt = []
    for _ in range(n):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); tt.append(time.perf_counter() - t)
    return 1e3 * statistics.median(tt)
streams = [torch.cuda.Stream(dev) for _ in range(4)]
for nch in (1, 2, 4, 8):
    w = B // nch
    def up():
        for k in range(nch):
            with torch.cuda.stream(streams[k % 4]): din[:, k * w:(k + 1) * w].copy_(hin[:, k * w:(k + 1) * w], non_blocking=True)
    def down():
        for k in range(nch):
            with torch.cuda.stream(streams[k % 4]): hout[:, k * w:(k + 1) * w].copy_(dout[:, k * w:(k + 1) * w], non_blocking=True)
    print("column ranges %d: H2D 2.36 MB %.4f ms, D2H 1.64 MB %.4f ms" % (nch, timeit(up), timeit(down)))
# four quarter batches side by side, device-resident
d = capi.soa_from_instances(S.cfg3_instances(B)); K = int(d["n_obs"])
sol = [m.BatchSolver(max_batch=B) for _ in range(4)]
for s in sol: s.set_kernel_timing(False)
def mk(lo, hi):
    tin = {k: torch.from_numpy(np.ascontiguousarray(d[k][:, lo:hi])).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
    n = hi - lo
    tout = {k: torch.zeros((r, n), dtype=torch.float64, device=dev) for k, r in capi.OUT_F64.items()}
    tout.update({k: torch.zeros((r, n), dtype=torch.int32, device=dev) for k, r in capi.OUT_I32.items()})
    tout["active"] = torch.zeros((2, n), dtype=torch.int64, device=dev)
    return tin, tout, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}
full = mk(0, B); q = [mk(k * 1024, (k + 1) * 1024) for k in range(4)]
print("one launch of 4096: %.4f ms" % timeit(lambda: sol[0].solve_device(B, full[2], full[3], K, mode=1, stream=streams[0].cuda_stream)))
print("one launch of 1024: %.4f ms" % timeit(lambda: sol[0].solve_device(1024, q[0][2], q[0][3], K, mode=1, stream=streams[0].cuda_stream)))
def four():
    for k in range(4): sol[k].solve_device(1024, q[k][2], q[k][3], K, mode=1, stream=streams[k].cuda_stream)
print("four launches of 1024 on four streams: %.4f ms" % timeit(four))
