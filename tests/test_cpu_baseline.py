"""The threaded C/OpenMP CPU baseline (oracle/cpu_baseline.py + oracle/c/ref_kernels.c) against the NumPy oracle.

Both are test/measurement infrastructure; the baseline is what bench.py times on the host cores, so it must be the
same algorithm: identical iteration counts, CG iteration counts and iterates up to reduction-order rounding."""
import copy

import numpy as np
import pytest

import problems as pr

orc = pr.OracleAPI()


def _run_both(spec, tweak=None):
    from oracle import cpu_baseline as cb
    outs = []
    for fast in (False, True):
        opt = orc.PARSDMM_options()
        if tweak:
            tweak(opt)
        b = pr.build(orc, copy.deepcopy(spec), opt)
        m = spec["m"].copy()
        if fast:
            outs.append(cb.PARSDMM(m, b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"], constraint=b["cons"]))
        else:
            outs.append(orc.PARSDMM(m, b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"]))
    return outs


@pytest.mark.parametrize("TF,tol", [(np.float64, 1e-9), (np.float32, 2e-5)])
def test_threaded_baseline_matches_oracle_config1(TF, tol):
    (xo, lo, l_o, y_o), (xf, lf, l_f, y_f) = _run_both(pr.spec_config1((48, 40), TF))
    assert len(lf.obj) == len(lo.obj)
    assert np.array_equal(lf.cg_it, lo.cg_it)
    assert np.linalg.norm(xf.astype(np.float64) - xo) <= tol * np.linalg.norm(xo.astype(np.float64))
    assert np.allclose(lf.set_feasibility, lo.set_feasibility, rtol=1e-3, atol=1e-9)
    assert np.allclose(lf.rho, lo.rho, rtol=1e-4) and np.allclose(lf.gamma, lo.gamma, rtol=1e-4)
    for a, b in zip(y_f, y_o):
        assert np.linalg.norm(a.astype(np.float64) - b) <= 50 * tol * max(np.linalg.norm(b.astype(np.float64)), 1.0)


def test_threaded_baseline_matches_oracle_config2_3d():
    (xo, lo, _, _), (xf, lf, _, _) = _run_both(pr.spec_config2((20, 16, 12), np.float32))
    assert len(lf.obj) == len(lo.obj) and np.array_equal(lf.cg_it, lo.cg_it)
    assert np.linalg.norm(xf.astype(np.float64) - xo) <= 2e-5 * np.linalg.norm(xo.astype(np.float64))
    assert np.allclose(lf.obj, lo.obj, rtol=1e-4)


def test_threaded_baseline_kernels():
    """C kernels against NumPy on random data: CSR/CSC products, CDS product, sort-based l1 projection."""
    from oracle import cpu_baseline as cb
    from oracle import operators as oops, projectors as oproj
    for TF in (np.float32, np.float64):
        k = cb._K(TF)
        A, *_ = orc.get_TD_operator(orc.compgrid((2.0, 3.0, 1.5), (9, 7, 5)), "TV", TF)
        op = cb._Op(A, TF)
        x = pr.splitmix_uniform(1, A.shape[1]).astype(TF)
        v = pr.splitmix_uniform(2, A.shape[0]).astype(TF)
        s = np.empty(A.shape[0], dtype=TF)
        k.csr_matvec(A.shape[0], op.rp.ctypes.data, op.ci.ctypes.data, op.va.ctypes.data, x.ctypes.data, s.ctypes.data)
        assert np.array_equal(s, oops.spmv(A, x))
        t = np.empty(A.shape[1], dtype=TF)
        k.csc_rmatvec(A.shape[1], op.cp.ctypes.data, op.ri.ctypes.data, op.vt.ctypes.data, TF(1), v.ctypes.data, None, t.ctypes.data, 0)
        assert np.array_equal(t, oops.spmv_t(A, v))
        R, off = oops.mat2CDS(oops.AtA_sparse(A))
        R = np.asfortranarray(R)
        yy = np.empty(A.shape[1], dtype=TF)
        k.cds_mvp(A.shape[1], off.size, R.ctypes.data, off.astype(np.int64).ctypes.data, x.ctypes.data, yy.ctypes.data)
        assert np.array_equal(yy, oops.Ax_CDS(x, R, off))
        w = (pr.splitmix_uniform(3, 5000) * 3).astype(TF)
        w[::7] = 0
        ref = oproj.project_l1_Duchi(w.copy(), TF(0.2 * np.abs(w).sum()))
        key = np.uint32 if TF == np.float32 else np.uint64
        work = np.empty(2 * w.size, dtype=key)
        got = w.copy()
        assert k.project_l1(w.size, got.ctypes.data, TF(0.2 * np.abs(w).sum()), work.ctypes.data) == 1
        assert np.allclose(got, ref, rtol=1e-5 if TF == np.float32 else 1e-12, atol=1e-6 if TF == np.float32 else 1e-13)
        assert abs(np.abs(got).sum() / (0.2 * np.abs(w).sum()) - 1) < 1e-4
        inside = w.copy()
        assert k.project_l1(w.size, inside.ctypes.data, TF(10 * np.abs(w).sum()), work.ctypes.data) == 0 and np.array_equal(inside, w)
    assert cb.threads() >= 1
