"""The C++ host side (include/mpc_dist_gpu.hpp + examples/team_cycles.cpp) above the C ABI."""
import os
import subprocess
import numpy as np
import pytest
import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "team_cycles")
TS = 0.043
PSTART = np.array([[0.0, 0.0], [0.0, -0.9], [-1.0, 0.0], [-1.0, -0.9]])
OBS = np.array([[2.2, 1, 0.85, 1, 1, 3, 3.0, 3, -100], [0.9, 1.2, -0.65, 2, -1.75, 0.2, -0.70, -2.2, 99.5]])
FX = np.array([0.2188, 0.2188, -0.1472, -0.1472]); FY = np.array([-0.1320, 0.1320, -0.1320, 0.1320])


def test_host_example_builds_and_fails_loudly_without_gpu(built):
    import torch
    assert os.path.exists(EXE)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE, "1", "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_team_cycles_cpp_matches_oracle(built):
    """5 control cycles of the 4-agent team through the C++ MPC_dist_gpu shim: every cycle's X / alpha equals the oracle's
    answer for the same hidden state (gaitDomain_, alpha buffer, nominal-stance rule, trot contact pattern)."""
    cycles = 5
    r = subprocess.run([EXE, str(cycles), "1"], capture_output=True, text=True, check=True)
    rows = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
    assert rows.shape == (4 * cycles, 5 + 16 + 20)
    goal = np.array([10.0, 0.0])
    q = PSTART.copy(); dq = np.zeros((4, 2))
    abuf = np.stack([PSTART[:, 0], 0 * PSTART[:, 0], PSTART[:, 1], 0 * PSTART[:, 0]], 1)
    for c in range(cycles):
        insts = []
        for a in range(4):
            v = 0.2 * (goal - PSTART[a]) / np.linalg.norm(goal - PSTART[a])
            xdes = np.zeros(16)
            for k in range(4):
                j = 4 * c + k
                xdes[4 * k:4 * k + 4] = [PSTART[a, 0] + v[0] * (j + 1) * TS, v[0], PSTART[a, 1] + v[1] * (j + 1) * TS, v[1]]
            base = PSTART[a] if c <= 1 else q[a]
            feet = np.vstack([base[0] + FX, base[1] + FY])
            ind = [1, 1, 1, 1] if c < 2 else ([1, 0, 0, 1] if c % 2 == 0 else [0, 1, 1, 0])
            insts.append((np.array([q[a, 0], dq[a, 0], q[a, 1], dq[a, 1]]), xdes, feet, ind, OBS.copy()))
        ref = H.oracle_batch(insts, mode=1, alpha_buf=abuf)
        for a in range(4):
            row = rows[4 * c + a]
            assert int(row[0]) == c and int(row[1]) == a
            assert int(row[2]) == ref["status"][a]
            if ref["status"][a] == 0:
                assert row[3] < 1e-8
                assert abs(row[4] - ref["slack"][a]) < 1e-6 * max(1.0, abs(ref["slack"][a]))
                assert np.abs(row[5:21] - ref["X"][a]).max() < 1e-6 * max(1.0, np.abs(ref["X"][a]).max())
                al = row[21:41].reshape(5, 4).T          # column-major 4x5 from get_alphaCOM
                assert np.abs(al - ref["alpha"][a].reshape(4, 5)).max() < 1e-6
            # roll forward exactly like the driver (with ITS numbers, so that later cycles see identical inputs)
            X = row[5:21]
            q[a] = [X[12], X[14]]; dq[a] = [X[13], X[15]]
            abuf[a] = X[12:16]


@pytest.mark.gpu
def test_team_cycles_cpp_extended_mode_matches_the_checker(built):
    """BASELINE config 2 in extended form (no reference implementation exists): 4 receding-horizon cycles of the 4-agent team of
    A1_Sim_4_agents with inter-agent rows, through the C++ host class (MPC_dist_gpu_team::run_NMPC_extended); every cycle equals the
    checker's team solve (oracle/cbfnmpc_oracle_ext.c) for the same hidden state."""
    import oracle as O
    cycles = 4
    r = subprocess.run([EXE, "--extended", str(cycles), "1"], capture_output=True, text=True, check=True)
    rows = np.array([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
    assert rows.shape == (4 * cycles, 5 + 16 + 20)
    goal = np.array([10.0, 0.0])
    q = PSTART.copy(); dq = np.zeros((4, 2))
    eps_nbr = float(np.float32(2.2))
    for c in range(cycles):
        insts = []
        for a in range(4):
            v = 0.2 * (goal - PSTART[a]) / np.linalg.norm(goal - PSTART[a])
            xdes = np.zeros(16)
            for k in range(4):
                j = 4 * c + k
                xdes[4 * k:4 * k + 4] = [PSTART[a, 0] + v[0] * (j + 1) * TS, v[0], PSTART[a, 1] + v[1] * (j + 1) * TS, v[1]]
            base = PSTART[a] if c <= 1 else q[a]
            feet = np.vstack([base[0] + FX, base[1] + FY])
            ind = [1, 1, 1, 1] if c < 2 else ([1, 0, 0, 1] if c % 2 == 0 else [0, 1, 1, 0])
            near = np.argsort(np.hypot(OBS[0] - q[a, 0], OBS[1] - q[a, 1]), kind="stable")[:1]
            insts.append((np.array([q[a, 0], dq[a, 0], q[a, 1], dq[a, 1]]), xdes, feet[None], ind, OBS[:, near].copy()))
        res, ti = O.ext_solve_teams(*H.ext_arrays(insts), T=4, N=4, eps_nbr=eps_nbr, nthreads=1)
        for a in range(4):
            row = rows[4 * c + a]
            assert int(row[2]) == res[a]["status"], (c, a, row[2], res[a]["status"])
            if res[a]["status"] == 0 and all(r_["status"] == 0 for r_ in res):
                assert row[3] < 1e-8
                assert abs(row[4] - res[a]["slack"]) < 1e-6 * max(1.0, abs(res[a]["slack"]))
                assert np.abs(row[5:21] - res[a]["X"]).max() < 1e-6 * max(1.0, np.abs(res[a]["X"]).max())
            X = row[5:21]
            q[a] = [X[12], X[14]]; dq[a] = [X[13], X[15]]
    assert rows[:, 2].max() == 0      # the reference's own 4-agent scenario converges in every cycle, inter-agent rows included
