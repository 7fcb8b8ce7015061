"""Dev tool: time the data-movement kernels around the solve (pack, team states, classify + solve split, world update) on a
large batch and print achieved GB/s against their algorithmic bytes."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_cycles = 4
dev = torch.device("cuda", 0)
w = S.mc_worlds(B, n_cycles)
keys = ("q", "dq", "contact", "toe_pos", "gait_domain", "agent_id", "team", "agent_initial", "Pr_refined", "Prd_refined", "Pobs_real",
        "alpha_buf", "success", "distance_to_fail", "n_not_converged")
t = {k: torch.from_numpy(np.ascontiguousarray(w[k])).to(dev) for k in keys}
raw = {k: t[k].data_ptr() for k in keys[:11]}
raw.update(nq=w["nq"], ref_cols=w["ref_cols"], n_teams=w["n_teams"], n_obs=w["n_obs"])
K = w["n_obs"]
s = m.BatchSolver(device=0, max_batch=B, max_obs=K + 3)
packed = {"x0": torch.zeros((4, B), dtype=torch.float64, device=dev), "xdes": torch.zeros((16, B), dtype=torch.float64, device=dev),
          "feet_xy": torch.zeros((8, B), dtype=torch.float64, device=dev), "contact": torch.zeros((4, B), dtype=torch.int32, device=dev),
          "obs_xy": torch.zeros((2 * K, B), dtype=torch.float64, device=dev)}
pp = {k: v.data_ptr() for k, v in packed.items()}
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
ms = timeit(lambda: s.pack_inputs(B, raw, pp, stream=st.cuda_stream))
by = B * (8 * (4 + 12 + 2 + 32 + 2 * K) + 4 * 7 + 8 * (28 + 2 * K) + 16)
print("pack_kernel      B %d: %.3f ms, %.1f KB per agent algorithmic, %.0f GB/s" % (B, ms, by / B / 1e3, by / ms / 1e6))
so = torch.zeros((12, B), dtype=torch.float64, device=dev)
ms = timeit(lambda: s.team_states(B, 4, t["q"].data_ptr(), t["dq"].data_ptr(), so.data_ptr(), stream=st.cuda_stream))
by = B * 8 * (12 + 12)
print("team_states      B %d: %.3f ms, %d B per agent algorithmic, %.0f GB/s" % (B, ms, by // B, by / ms / 1e6))
s.close()
