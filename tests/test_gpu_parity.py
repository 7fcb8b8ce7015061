"""-m gpu: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(built):
    import srb_cbf_nmpc_b200 as m
    return m


@pytest.fixture(scope="module")
def solver(pkg):
    return pkg.BatchSolver(device=0, max_batch=8192, max_obs=20)


def test_library_is_cuda_and_loaded(pkg, solver):
    assert pkg.load().cbfnmpc_version() >= 100
    maps = open("/proc/self/maps").read()
    assert "libcbfnmpc.so" in maps
    assert solver.fp64_peak_tflops() > 1.0


def test_cfg1_single_solves(pkg, solver):
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    insts = H.pad_obs([S.cfg1_instance(4), S.cfg1_instance(2)])
    for mode in (capi.MODE_QP, capi.MODE_QP_NLP):
        out = solver.solve_host(capi.soa_from_instances(insts), mode=mode)
        ref = H.oracle_batch(insts, mode=mode)
        c = H.compare(out, ref)
        assert c["status_equal"].all() and c["both"].all()
        assert c["worst"] < 1e-9, c
        assert (out["qp_iters"][0] == ref["qp_iters"]).all()
    assert out["kkt"][0].max() < 1e-8


def test_cfg3_1024_parity(pkg, solver):
    from srb_cbf_nmpc_b200 import capi, scenarios as S
    insts = S.cfg3_instances(1024)
    out = solver.solve_host(capi.soa_from_instances(insts), mode=capi.MODE_QP_NLP)
    ref = H.oracle_batch(insts, mode=1)
    c = H.compare(out, ref, 1e-6)
    assert c["status_equal"].mean() > 0.995, np.nonzero(~c["status_equal"])[0][:20]
    both = c["both"]
    assert both.sum() > 900
    bad = np.nonzero(both & ((c["ex"] > 1e-6) | (c["eu"] > 1e-6) | (c["es"] > 1e-6)))[0]
    assert len(bad) == 0, (bad[:10], c["worst"])
    assert out["kkt"][0][both].max() < 1e-8
    # same active set on friction / distance / velocity rows
    mism = 0
    for i in np.nonzero(both)[0][:300]:
        cc = int(ref["c"][i])
        if H.active_rows_oracle(ref, i, cc) != H.active_rows_mask(out, i, cc):
            mism += 1
    assert mism == 0, mism
    # Bezier epilogue and QP stage
    assert np.abs(out["alpha"].T - ref["alpha"])[both].max() < 1e-6
    assert np.abs(out["x_qp"].T - ref["x_qp"][:, :24]).max() < 1e-8
