"""Dev tool: compare the dumps of scripts/dev/ab_compare.sh bit for bit.  python scripts/dev/ab_diff.py tagA tagB"""
import sys, numpy as np
a, b = sys.argv[1], sys.argv[2]
for n in (4096, 32768):
    A = np.load("gpurun_out/status_%s_%d.npz" % (a, n)); B = np.load("gpurun_out/status_%s_%d.npz" % (b, n))
    bad = [k for k in A.files if not np.array_equal(A[k], B[k], equal_nan=True)]
    print(n, "arrays that differ:", bad or "none (bit-identical)")
    for k in bad:
        d = np.abs(A[k].astype(float) - B[k].astype(float)); print("   ", k, "max |diff|", np.nanmax(d), "instances", int((d.reshape(d.shape[0], -1) > 0).any(0).sum()))
