"""Dev tool (CPU only): what a second-order (Mehrotra-type) correction of the complementarity rows would do to the NLP iteration tail.
Builds a PATCHED COPY of the oracle in a temporary directory (scripts/dev/soc_experiment.patch: after the Newton direction, the same
factor is used once more with mu - clip(ds dz, +-cap max(mu, s z)) in place of mu; the corrected direction is taken only when it does
not shorten the step lengths by more than 10 %) and runs it on the first two shards of the bench workload.  The oracle itself is
untouched.  usage: python scripts/dev/soc_experiment.py"""
import os, shutil, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tmp = tempfile.mkdtemp(prefix="soc_")
for f in ("cbfnmpc_oracle.c", "cbfnmpc_oracle.h", "cbfnmpc_oracle_ext.c", "cbfnmpc_oracle_ext.h"):
    shutil.copy(os.path.join(ROOT, "oracle", f), tmp)
subprocess.check_call(["patch", "-s", os.path.join(tmp, "cbfnmpc_oracle.c"), os.path.join(ROOT, "scripts", "dev", "soc_experiment.patch")])
subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu99", "-shared", "-o", os.path.join(tmp, "liborc.so"), os.path.join(tmp, "cbfnmpc_oracle.c"),
                       os.path.join(tmp, "cbfnmpc_oracle_ext.c"), "-lm", "-lpthread"])
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import oracle as O
O.build = lambda force=False: None; O.HERE = tmp; O._lib = None
from srb_cbf_nmpc_b200 import scenarios as S
insts = S.cfg3_instances(4096 * 2, seed=20251018 + 3000)
x0 = np.array([i[0] for i in insts]); xd = np.array([i[1] for i in insts]); ft = np.array([i[2] for i in insts])
ct = np.array([i[3] for i in insts], np.int32); ob = np.array([i[4] for i in insts])
ab = np.stack([x0[:, 0], 0 * x0[:, 0], x0[:, 2], 0 * x0[:, 0]], 1)
base = None
for soc in (None, "0.5", "1", "3"):
    if soc is None: os.environ.pop("ORC_SOC", None)
    else: os.environ["ORC_SOC"] = soc
    res = O.solve_batch(x0, xd, ft, ct, ob, ab, mode=1, nthreads=os.cpu_count() or 1, raw=True)
    st = np.array([r.status for r in res]); qi = np.array([r.qp_iters for r in res]); ni = np.array([r.nlp_iters for r in res]); kkt = np.array([r.kkt for r in res])
    if base is None: base = st.copy()
    cv = st == 0
    print("corrector cap %-4s converged %.4f (statuses as without: %.4f)  NLP iterations mean %.2f p99 %d max %d  histogram >= 13: %s  worst path 0.75 qp + nlp = %.1f  KKT max %.1e"
          % (soc, cv.mean(), (st == base).mean(), ni[cv].mean(), np.percentile(ni[cv], 99), ni[cv].max(), np.bincount(ni[cv])[13:], (0.75 * qi + ni)[cv].max(), kkt[cv].max()))
shutil.rmtree(tmp)
