"""CPU-side check of the *kernel source*: srb-cbf-nmpc_b200/csrc compiled for the host with 32 cooperative fibers per warp
(tests/emu).  This is a logic test of the product kernels, not a fallback: the C ABI never reaches the emulator."""
import numpy as np
import helpers as H
import emu
from srb_cbf_nmpc_b200 import scenarios as S


def test_emulated_kernel_qp_mode():
    insts = H.pad_obs([S.cfg1_instance(4), S.cfg1_instance(2)]) + S.cfg3_instances(10)
    out = emu.solve(insts, mode=0)
    ref = H.oracle_batch(insts, mode=0)
    assert (out["qp_iters"][0] == ref["qp_iters"]).all()
    assert np.abs(out["X"].T - ref["X"]).max() < 1e-9 and np.abs(out["U"].T - ref["U"]).max() < 1e-9


def test_emulated_kernel_nlp_mode():
    insts = H.pad_obs([S.cfg1_instance(4), S.cfg1_instance(2)]) + S.cfg3_instances(22)
    out = emu.solve(insts, mode=1)
    ref = H.oracle_batch(insts, mode=1)
    c = H.compare(out, ref)
    assert c["status_equal"].all()
    assert c["both"].sum() >= 20 and c["worst"] < 1e-8
    assert out["kkt"][0][c["both"]].max() < 1e-8
    for i in np.nonzero(c["both"])[0]:
        cc = int(ref["c"][i])
        assert H.active_rows_oracle(ref, i, cc) == H.active_rows_mask(out, i, cc)
    assert np.abs(out["alpha"].T - ref["alpha"])[c["both"]].max() < 1e-8


def test_bad_contact_pattern_is_flagged():
    inst = list(S.cfg1_instance(4)); inst[3] = [1, 1, 1, 0]
    out = emu.solve(H.pad_obs([tuple(inst)]), mode=1)
    assert out["status"][0, 0] == 4
