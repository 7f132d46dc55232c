"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): final x relative L2 error <= 1e-5 in Float64 and <= 1e-3 in
Float32, with matching iteration counts; bit-exact for index work (offsets, support sets)."""
import copy

import numpy as np
import pytest

import problems as pr

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-3, np.float64: 1e-5}


def relerr(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


@pytest.fixture(scope="module")
def orc():
    return pr.OracleAPI()


# ---------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n,kind", [((9, 6), "TV"), ((9, 6), "D_x"), ((9, 6), "D_z"), ((9, 6), "D_xz"),
                                    ((9, 6), "identity"), ((4, 6, 5), "TV"), ((4, 6, 5), "D_x"),
                                    ((4, 6, 5), "D_y"), ((4, 6, 5), "D_z"), ((33, 17, 9), "TV")])
def test_operator_forward_adjoint_bit_exact(sip, orc, TF, n, kind):
    d = (0.99, 1.123, 1.7)[: len(n)]
    cg = orc.compgrid(d, n)
    A = orc.get_TD_operator(cg, kind, TF)[0]
    op = sip.get_TD_operator(sip.compgrid(d, n), kind, TF)[0]
    rng = np.random.default_rng(1)
    x = rng.standard_normal(A.shape[1]).astype(TF)
    v = rng.standard_normal(A.shape[0]).astype(TF)
    assert np.array_equal(op @ x, orc.ops.spmv(A, x))
    assert np.array_equal(op.T @ v, orc.ops.spmv_t(A, v))


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_operator_minkowski_blocks(sip, orc, TF):
    import scipy.sparse as sp
    n, d = (7, 5), (2.0, 3.0)
    A = orc.get_TD_operator(orc.compgrid(d, n), "TV", TF)[0]
    op = sip.get_TD_operator(sip.compgrid(d, n), "TV", TF)[0]
    rng = np.random.default_rng(2)
    x = rng.standard_normal(2 * A.shape[1]).astype(TF)
    v = rng.standard_normal(A.shape[0]).astype(TF)
    Z = sp.csc_matrix(A.shape, dtype=TF)
    for mode, blocks in ((1, [A, Z]), (2, [Z, A]), (3, [A, A])):
        B = sp.hstack(blocks, format="csc", dtype=TF)
        B.sort_indices()
        o = op.with_block(mode)
        assert np.array_equal(o @ x, orc.ops.spmv(B, x))
        assert np.array_equal(o.T @ v, orc.ops.spmv_t(B, v))


@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n", [(30, 20), (13, 11, 7), (64, 50, 3)])
def test_cds_spmv_bit_exact(sip, orc, TF, n):
    """test_CDS_Mvp.jl:13-22 on TV'TV, plus bit-exactness w.r.t. the per-diagonal accumulation order."""
    d = (25.0,) * len(n)
    A = orc.get_TD_operator(orc.compgrid(d, n), "TV", TF)[0]
    R, off = orc.mat2CDS(orc.ops.AtA_sparse(A))
    # reorder like Q_offsets (0 first) to exercise a non-sorted accumulation order
    order = np.argsort(np.where(off == 0, -10**12, off), kind="stable")
    R, off = np.asfortranarray(R[:, order]), off[order]
    N = R.shape[0]
    x = np.random.default_rng(3).standard_normal(N).astype(TF)
    want = orc.ops.CDS_MVp(N, R.shape[1], R, off, x, np.zeros(N, dtype=TF))
    got = sip.CDS_MVp(N, R.shape[1], R, off, x, np.zeros(N, dtype=TF))
    assert np.array_equal(got, want)
    native = np.asarray(orc.ops.AtA_sparse(A) @ x).ravel()
    assert np.allclose(got, native, rtol=0, atol=10 * np.finfo(TF).eps * np.abs(native).max())


@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n,order", [((8, 6, 5), "q"), ((24, 20, 17), "q"), ((40, 28, 20), "z"), ((64, 50, 9), "sorted"),
                                     ((200, 37, 12), "q"), ((512, 9, 4), "z"), ((4, 3, 2), "q"), ((256, 5, 3), "q")])
def test_cds_spmv_tiled_bit_exact(sip, orc, TF, n, order):
    """The tiled, TMA-staged plane-sweep SpMV (spmv_tile.cuh) against CDS_MVp (CDS_MVp.jl:9-28): random values on
    all seven stencil diagonals — INCLUDING the entries that wrap around line and plane ends, which a general CDS
    matrix may hold — in three accumulation orders; results must be bit-identical."""
    import ctypes as C
    L = sip._lib
    n0, n1, n2 = n
    N, P = n0 * n1 * n2, n0 * n1
    off = {"q": [0, -P, -n0, -1, 1, n0, P], "z": [0, -P, P, -1, 1, -n0, n0], "sorted": [-P, -n0, -1, 0, 1, n0, P]}[order]
    off = np.array(off, dtype=np.int64)
    rng = np.random.default_rng(11)
    R = np.asfortranarray(rng.standard_normal((N, off.size)).astype(TF))
    x = rng.standard_normal(N).astype(TF)
    want = orc.ops.CDS_MVp(N, off.size, R, off, x, np.zeros(N, dtype=TF))
    got = np.empty(N, dtype=TF)
    used = C.c_int(0)
    n3 = (C.c_int64 * 3)(*n)
    L.check(L.load().sipb_cds_spmv_grid(L.ctx(), L.dtype_code(TF), n3, off.size, R.ctypes.data,
                                        off.ctypes.data_as(C.POINTER(C.c_int64)), x.ctypes.data, got.ctypes.data, C.byref(used)))
    assert used.value == 1            # these grids are inside the tiled kernel's domain
    assert np.array_equal(got, want)
    # a subset of the diagonals (a D_x-only Q) is outside the tiled kernel's domain (full 7-point stencils): generic kernel
    sub = np.array([0, -1, 1], dtype=np.int64)
    Rs = np.asfortranarray(R[:, [0, 3, 4]]) if order == "q" else np.asfortranarray(R[:, :3])
    want = orc.ops.CDS_MVp(N, 3, Rs, sub, x, np.zeros(N, dtype=TF))
    L.check(L.load().sipb_cds_spmv_grid(L.ctx(), L.dtype_code(TF), n3, 3, Rs.ctypes.data,
                                        sub.ctypes.data_as(C.POINTER(C.c_int64)), x.ctypes.data, got.ctypes.data, C.byref(used)))
    assert used.value == 0 and np.array_equal(got, want)


def test_cds_spmv_grid_falls_back_outside_tile_domain(sip, orc):
    import ctypes as C
    L = sip._lib
    n = (13, 11, 7)
    N, P = 13 * 11 * 7, 13 * 11
    off = np.array([0, -P, -13, -1, 1, 13, P], dtype=np.int64)
    rng = np.random.default_rng(12)
    R = np.asfortranarray(rng.standard_normal((N, 7)).astype(np.float32))
    x = rng.standard_normal(N).astype(np.float32)
    got = np.empty(N, dtype=np.float32)
    used = C.c_int(1)
    L.check(L.load().sipb_cds_spmv_grid(L.ctx(), 0, (C.c_int64 * 3)(*n), 7, R.ctypes.data,
                                        off.ctypes.data_as(C.POINTER(C.c_int64)), x.ctypes.data, got.ctypes.data, C.byref(used)))
    assert used.value == 0
    assert np.array_equal(got, orc.ops.CDS_MVp(N, 7, R, off, x, np.zeros(N, dtype=np.float32)))


def test_cds_spmv_random_offsets(sip, orc):
    """random banded matrix (test_CDS_Mvp.jl:13-22 second half), Float64."""
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    N = 1000
    A = sp.random(N, N, density=0.01, random_state=5, format="csc", dtype=np.float64)
    R, off = orc.mat2CDS(A)
    if R.shape[1] > 32:       # device limit on diagonals: keep the 32 first
        R, off = np.asfortranarray(R[:, :32]), off[:32]
    x = rng.standard_normal(N)
    want = orc.ops.CDS_MVp(N, R.shape[1], R, off, x, np.zeros(N))
    got = sip.CDS_MVp(N, R.shape[1], R, off, x, np.zeros(N))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_cds_scaled_add(sip, orc, TF):
    """test_CDS_scaled_add.jl:24-39: CDS(A) + CDS(B) == CDS(A+B) exactly; missing diagonal raises."""
    n, d = (30, 20), (25.0, 25.0)
    TV = orc.get_TD_operator(orc.compgrid(d, n), "TV", TF)[0]
    Dz = orc.get_TD_operator(orc.compgrid(d, n), "D_z", TF)[0]
    A, B = orc.ops.AtA_sparse(TV), orc.ops.AtA_sparse(Dz)
    RA, oA = orc.mat2CDS(A)
    RB, oB = orc.mat2CDS(B)
    RC, oC = orc.mat2CDS((A + B).astype(TF))
    got = sip.CDS_scaled_add(RA.copy(order="F"), RB, oA, oB, 1.0)
    assert np.array_equal(got, RC) and np.array_equal(oA, oC)
    with pytest.raises(Exception):
        sip.CDS_scaled_add(RB.copy(order="F"), RA, oB, oA, 1.0)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_cg_matches_oracle(sip, orc, TF):
    """cg.jl semantics: residual <= tol, exact start => iter == 1 and x untouched (test_cg.jl:25-29),
    zero rhs => flag -9, iterates close to the oracle's."""
    n, d = (24, 20), (1.0, 1.0)
    TV = orc.get_TD_operator(orc.compgrid(d, n), "TV", TF)[0]
    R, off = orc.mat2CDS(orc.ops.AtA_sparse(TV))
    R[:, list(off).index(0)] += TF(1.0)        # Q = TV'TV + I  (SPD)
    N = R.shape[0]
    rng = np.random.default_rng(6)
    xt = rng.standard_normal(N).astype(TF)
    b = orc.ops.Ax_CDS(xt, R, off)
    tol = 1e-5 if TF == np.float32 else 1e-10
    x, flag, relres, it = sip.cg(R, off, b, tol=tol, maxIter=1000)
    assert flag == 0 and np.linalg.norm(orc.ops.Ax_CDS(x, R, off) - b) / np.linalg.norm(b) <= 2 * tol
    xo, fo, ro, ito = orc.parsdmm.cg(lambda v: orc.ops.Ax_CDS(v, R, off), b.copy(), TF(tol), 1000, np.zeros(N, dtype=TF))
    assert it == ito and relerr(x, xo) < TOL[TF]
    assert abs(float(relres) - float(ro)) <= 1e-2 * float(ro) + 1e-30
    x2, flag2, relres2, it2 = sip.cg(R, off, b, tol=tol, maxIter=1000, x=xt.copy())
    assert it2 == 1 and flag2 == 0 and np.array_equal(x2, xt) and relres2 == 0
    x3, flag3, relres3, it3 = sip.cg(R, off, np.zeros(N, dtype=TF), tol=tol, maxIter=10, x=xt.copy())
    assert flag3 == -9 and it3 == 0 and not x3.any()


# ---------------------------------------------------------------------------------------------
# projectors
# ---------------------------------------------------------------------------------------------
def _P(sip, st, lo, hi, TF, n=(50, 2)):
    cons = [sip.set_definitions(st, "identity", lo, hi, ("matrix", ""))]
    return sip.setup_constraints(cons, sip.compgrid((1.0, 1.0), n), TF)[0][0]


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_projectors_match_oracle(sip, orc, TF):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(100).astype(TF)
    # bounds (bit exact)
    assert np.array_equal(_P(sip, "bounds", -0.11, 0.01, TF)(x.copy()), orc.proj.project_bounds(x.copy(), TF(-0.11), TF(0.01)))
    lo = (rng.standard_normal(100) - 1).astype(TF)
    hi = (rng.standard_normal(100) + 1).astype(TF)
    assert np.array_equal(_P(sip, "bounds", lo, hi, TF)(x.copy()), orc.proj.project_bounds(x.copy(), lo, hi))
    # l1: untouched inside the ball, ||x||_1 == tau outside (test_projectors.jl:22-35), close to Duchi
    tau = TF(np.abs(x).sum() * 2)
    assert np.array_equal(_P(sip, "l1", 0.0, tau, TF)(x.copy()), x)
    tau = TF(np.abs(x).sum() * 0.234)
    got = _P(sip, "l1", 0.0, tau, TF)(x.copy())
    assert abs(np.abs(got.astype(np.float64)).sum() - float(tau)) <= 20 * np.finfo(TF).eps * float(tau)
    assert relerr(got, orc.proj.project_l1_Duchi(x.copy(), tau)) < 20 * np.finfo(TF).eps
    # l2 / annulus
    # (the 2-norm is a reduction: summation order differs, so agreement is to a few ulps)
    ulps = 8 * np.finfo(TF).eps
    got = _P(sip, "l2", 0.0, 0.123, TF)(x.copy())
    assert relerr(got, orc.proj.project_l2(x.copy(), TF(0.123))) < ulps
    assert abs(np.linalg.norm(got.astype(np.float64)) - 0.123) < 10 * np.finfo(TF).eps
    assert np.array_equal(_P(sip, "l2", 0.0, 1e3, TF)(x.copy()), x)
    for lo_, hi_ in ((20.0, 30.0), (0.1, 0.2), (1.0, 100.0)):
        assert relerr(_P(sip, "annulus", lo_, hi_, TF)(x.copy()), orc.proj.project_annulus(x.copy(), TF(lo_), TF(hi_))) < ulps
    z = np.zeros(100, dtype=TF)
    assert np.array_equal(_P(sip, "annulus", 2.0, 3.0, TF)(z.copy()), orc.proj.project_annulus(z.copy(), TF(2.0), TF(3.0)))
    # cardinality: literals of test_projectors.jl:49-56 and support equality with stable ties
    for lit, k, want in (([0, 0, 1, 2, 3], 2, [0, 0, 0, 2, 3]), ([0, 0, -1, 2, -3], 2, [0, 0, 0, 2, -3])):
        got = _P(sip, "cardinality", 0, k, TF, n=(5, 2))(np.array(lit, dtype=TF))
        assert np.array_equal(got, np.array(want, dtype=TF))
    got = _P(sip, "cardinality", 0, 5, TF)(x.copy())
    assert np.count_nonzero(got) == 5 and np.array_equal(got, orc.proj.project_cardinality(x.copy(), 5))
    t = np.round(rng.standard_normal(4000) * 3).astype(TF)      # many exact ties
    for k in (0, 1, 17, 500, 1999, 3999, 4000, 5000):
        got = _P(sip, "cardinality", 0, k, TF, n=(2000, 2))(t.copy())
        assert np.array_equal(got, orc.proj.project_cardinality(t.copy(), k)), k
    # prox_l1
    assert np.array_equal(_P(sip, "prox_l1", 0.0, 3.0, TF)(x.copy()), orc.proj.prox_l1(x.copy(), TF(3.0)))


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_l1_all_entries_active_cap(sip, orc, TF):
    """project_l1_Duchi!.jl:42-46: the scan stops at rho = lv-1, so when EVERY entry stays above the threshold the
    reference uses theta = max(0, (sv[lv-1]-b)/(lv-1)) instead of the exact root — including theta = 0, i.e. the
    vector comes back unprojected.  The device reproduces this (k_l1_cap)."""
    v = np.array([3.0, -2.5, 2.8, 3.1], dtype=TF)
    for tau in (10.0, 11.0, 5.0, 1.0):                    # 10, 11: theta = 0 (unchanged); 5, 1: capped but positive
        want = orc.proj.project_l1_Duchi(v.copy(), TF(tau))
        got = _P(sip, "l1", 0.0, tau, TF, n=(2, 2))(v.copy())
        assert relerr(got, want) < 20 * np.finfo(TF).eps, (tau, got, want)
    assert np.array_equal(_P(sip, "l1", 0.0, 10.0, TF, n=(2, 2))(v.copy()), v)
    rng = np.random.default_rng(3)
    w = (5.0 + rng.random(1000)).astype(TF) * np.where(rng.random(1000) < 0.5, -1, 1).astype(TF)
    tau = TF(np.abs(w).sum() * 0.97)                       # exact root 0.03*mean < min|w|: all 1000 entries active
    want = orc.proj.project_l1_Duchi(w.copy(), tau)
    got = _P(sip, "l1", 0.0, tau, TF, n=(500, 2))(w.copy())
    assert relerr(got, want) < 20 * np.finfo(TF).eps
    # a vector with ONE inactive entry takes the exact root (no cap)
    w[7] = TF(1e-3)
    tau = TF(np.abs(w).sum() * 0.97)
    want = orc.proj.project_l1_Duchi(w.copy(), tau)
    got = _P(sip, "l1", 0.0, tau, TF, n=(500, 2))(w.copy())
    assert relerr(got, want) < 20 * np.finfo(TF).eps and got[7] == 0


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_histogram_projector(sip, orc, TF):
    """project_histogram_relaxed.jl:9-26 on the device (stable radix sort in isless order + clamp through the
    permutation): the reference's own tests (test_projectors.jl:275-289) and bit-exactness against the oracle,
    including ties and signed zeros."""
    rng = np.random.default_rng(21)
    ref = np.sort(rng.standard_normal(100)).astype(TF)
    x = rng.standard_normal(100).astype(TF)
    got = _P(sip, "histogram", ref, ref, TF)(x.copy())
    assert np.array_equal(np.sort(got), ref)                                  # exact histogram: ref == sort(x)
    LB = np.sort(rng.standard_normal(100)).astype(TF)
    UB = (LB + TF(0.7)).astype(TF)
    got = _P(sip, "histogram", LB, UB, TF)(x.copy())
    assert np.all(np.sort(got) <= UB) and np.all(np.sort(got) >= LB)          # relaxed histogram
    assert np.array_equal(got, orc.proj.project_histogram_relaxed(x.copy(), LB, UB))
    # many ties, signed zeros: the stable order decides which bound an entry meets
    t = np.round(rng.standard_normal(5000) * 2).astype(TF)
    t[::7] = TF(-0.0)
    LB = np.sort(rng.standard_normal(5000)).astype(TF)
    UB = (LB + TF(0.05)).astype(TF)
    got = _P(sip, "histogram", LB, UB, TF, n=(2500, 2))(t.copy())
    want = orc.proj.project_histogram_relaxed(t.copy(), LB, UB)
    assert np.array_equal(got, want) and np.array_equal(np.signbit(got), np.signbit(want))


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_parsdmm_with_histogram_set(sip, orc, TF):
    """bounds ∩ relaxed histogram of the model ∩ TV-l1 through the full iteration (two-pass y/l update).
    The set is a union of permuted boxes (non-convex): its projection jumps when two entries swap rank, so Float32
    runs of the device and the oracle separate after a few dozen iterations (measured: identical logs for 12+
    iterations, different stopping iteration at 54 vs 60).  Float64 runs to the stopping rule; Float32 compares the
    first 12 iterations."""
    n = (40, 36)
    spec = pr.spec_config1(n, TF)
    m = spec["m"]
    other = np.sort(pr.synthetic_model(n, TF, seed=77))
    spec["sets"] = [("bounds", "identity", 1500.0, 4500.0), ("histogram", "identity", other - TF(40), other + TF(40)),
                    ("l1", "TV", 0.0, 0.6 * pr.tv_l1(n, spec["d"], TF, m))]

    def tw(o):
        o.maxit = 60 if TF == np.float64 else 12
    o, s = run_both(sip, orc, spec, tw)
    check_parity(o, s, TF)


def test_rejected_sets(sip):
    for st in ("rank", "nuclear", "subspace"):
        with pytest.raises(NotImplementedError):
            _P(sip, st, 0.0, 3.0, np.float32)
    with pytest.raises(NotImplementedError):
        cons = [sip.set_definitions("bounds", "DFT", 0.0, 1.0, ("matrix", ""))]
        sip.setup_constraints(cons, sip.compgrid((1.0, 1.0), (8, 8)), np.float32)
    with pytest.raises(Exception):
        _P(sip, "l1", 0.0, -1.0, np.float32)(np.ones(100, dtype=np.float32))


# ---------------------------------------------------------------------------------------------
# full PARSDMM
# ---------------------------------------------------------------------------------------------
def run_both(sip, orc, spec, tweak=None, **kw):
    o_opt = orc.PARSDMM_options()
    s_opt = sip.PARSDMM_options()
    if tweak:
        tweak(o_opt)
        tweak(s_opt)
    ob = pr.build(orc, copy.deepcopy(spec), o_opt)
    sb = pr.build(sip, copy.deepcopy(spec), s_opt)
    m = spec["m"]
    xo, lo, ll, yy = orc.PARSDMM(m.copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
    xs, ls, l2, y2 = sip.PARSDMM(m.copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], **kw)
    return (xo, lo, ll, yy, ob), (xs, ls, l2, y2, sb)


def check_parity(o, s, TF, iters_exact=True):
    xo, lo, ll, yy, ob = o
    xs, ls, l2, y2, sb = s
    assert np.array_equal(ob["set_Prop"].AtA_offsets[1], sb["set_Prop"].AtA_offsets[1])
    dev = getattr(sb["AtA"], "_device", None)
    if dev is not None:       # Q_offsets in the reference's first-appearance order (integer work, bit-exact)
        import oracle.operators as oops
        assert np.array_equal(dev.q_offsets, oops.assemble_Q(ob["AtA"], ob["set_Prop"].AtA_offsets, np.ones(len(ob["AtA"])))[1])
    if iters_exact:
        assert len(ls.obj) == len(lo.obj), (len(ls.obj), len(lo.obj))
        assert np.array_equal(ls.cg_it, lo.cg_it), (ls.cg_it, lo.cg_it)
        assert ls.set_feasibility.shape == lo.set_feasibility.shape
        tol = TOL[TF]
        assert np.allclose(ls.set_feasibility, lo.set_feasibility, rtol=50 * tol, atol=1e-12)
        assert np.allclose(ls.rho, lo.rho, rtol=50 * tol)
        assert np.allclose(ls.gamma, lo.gamma, rtol=50 * tol)
        assert np.allclose(ls.obj, lo.obj, rtol=50 * tol)
        assert np.allclose(ls.r_pri, lo.r_pri, rtol=1e-2, atol=1e-6 * np.abs(lo.r_pri).max())
        assert np.allclose(ls.r_dual, lo.r_dual, rtol=1e-2, atol=1e-5 * np.abs(lo.r_dual).max())
        assert np.allclose(ls.r_dual_total, lo.r_dual_total, rtol=1e-2, atol=1e-5 * np.abs(lo.r_dual_total).max())
        assert np.allclose(ls.r_pri_total, lo.r_pri_total, rtol=1e-2, atol=1e-6 * np.abs(lo.r_pri_total).max())
        assert np.allclose(ls.evol_x[1:], lo.evol_x[1:], rtol=1e-2, atol=1e-9)
        for a, b, yb in zip(l2, ll, yy):
            # multipliers of inactive sets are pure rounding noise: absolute floor relative to ||y||
            floor = 1e4 * np.finfo(TF).eps * np.linalg.norm(yb.astype(np.float64))
            assert np.linalg.norm(a.astype(np.float64) - b) <= 100 * tol * np.linalg.norm(b.astype(np.float64)) + floor
        for a, b in zip(y2, yy):
            assert relerr(a, b) < 100 * tol
    assert relerr(xs, xo) < TOL[TF], relerr(xs, xo)


@pytest.mark.parametrize("TF", [np.float64, np.float32])
def test_parsdmm_config1_2d(sip, orc, TF):
    o, s = run_both(sip, orc, pr.spec_config1((64, 64), TF))
    check_parity(o, s, TF)
    assert np.isnan(s[1].evol_x[0]) and s[1].cg_it[0] == 0       # quirk: iteration 1 has rhs == 0


def test_parsdmm_config1_f64_accurate(sip, orc):
    """test_PARSDMM.jl:97-111-style accurate settings: every set feasible to 1.5*feas_tol."""
    def tw(o):
        o.obj_tol, o.feas_tol, o.evol_rel_tol, o.maxit = 1e-9, 1e-9, 1e-12, 400
    o, s = run_both(sip, orc, pr.spec_config1((32, 32), np.float64), tw)
    check_parity(o, s, np.float64)


@pytest.mark.parametrize("n", [(24, 24, 24), (40, 28, 20)])
def test_parsdmm_config2_3d_f32(sip, orc, n):
    def tw(o):
        o.evol_rel_tol = 10 * np.finfo(np.float32).eps       # examples/test_scaling_3D.jl:25
        o.maxit = 60
    o, s = run_both(sip, orc, pr.spec_config2(n, np.float32), tw)
    check_parity(o, s, np.float32)


def test_parsdmm_config3_cardinality(sip, orc):
    def tw(o):
        o.maxit = 40
    o, s = run_both(sip, orc, pr.spec_config3((20, 20, 20), np.float32), tw)
    check_parity(o, s, np.float32)
    # non-convex overrides (PARSDMM_initialize.jl:107-114): gamma fixed at 0.75
    assert np.all(s[1].gamma == np.float32(0.75))
    # bit-exact support of the cardinality-projected auxiliary vector
    assert np.array_equal(s[3][2] != 0, o[3][2] != 0)


def test_parsdmm_config4_bounds_only(sip, orc):
    def tw(o):
        o.rho_ini = [1.0, 1000.0, 1000.0, 1000.0, 1.0]        # examples/test_scaling_3D.jl:97
        o.evol_rel_tol = 10 * np.finfo(np.float32).eps
        o.maxit = 50
    o, s = run_both(sip, orc, pr.spec_config4((20, 22, 24), np.float32), tw)
    check_parity(o, s, np.float32)


def test_parsdmm_feasible_input_returned(sip, orc):
    """test_PARSDMM.jl:17-36: a feasible model comes back unchanged with one-row logs."""
    TF = np.float64
    m = pr.synthetic_model((20, 30), TF)
    spec = dict(n=(20, 30), d=(1.0, 1.0), TF=TF, m=m, sets=[("bounds", "identity", float(m.min()), float(m.max()))],
                mode="matrix")
    o, s = run_both(sip, orc, spec)
    assert np.array_equal(s[0], m) and np.array_equal(o[0], m)
    assert len(s[1].obj) == 1 and s[1].set_feasibility.shape == (1, 1)


def test_parsdmm_option_variants(sip, orc):
    """test_PARSDMM.jl:113-189: option variants keep every set feasible and match the oracle."""
    spec = pr.spec_config1((40, 36), np.float64)
    for kw in (dict(adjust_gamma=False), dict(adjust_rho=False), dict(adjust_rho=False, adjust_gamma=False),
               dict(adjust_feasibility_rho=False), dict(rho_update_frequency=1), dict(gamma_ini=1.5)):
        def tw(o, kw=kw):
            for k, v in kw.items():
                setattr(o, k, v)
            o.maxit = 80
        o, s = run_both(sip, orc, spec, tw)
        check_parity(o, s, np.float64)


def test_parsdmm_warm_start(sip, orc):
    """Warm start x,l,y (zero_ini_guess=false, PARSDMM_initialize.jl:304-313): restart == oracle restart."""
    spec = pr.spec_config1((32, 32), np.float64)
    def tw(o):
        o.maxit = 8
    o, s = run_both(sip, orc, spec, tw)
    xo, lo, ll, yy, ob = o
    xs, ls, l2, y2, sb = s
    ob["opt"].zero_ini_guess = False
    sb["opt"].zero_ini_guess = False
    ob["opt"].maxit = sb["opt"].maxit = 40
    m = spec["m"]
    xo2, lo2, _, _ = orc.PARSDMM(m.copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"], xo.copy(),
                                 [v.copy() for v in ll], [v.copy() for v in yy])
    xs2, ls2, _, _ = sip.PARSDMM(m.copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], xs.copy(),
                                 [v.copy() for v in l2], [v.copy() for v in y2])
    assert len(ls2.obj) == len(lo2.obj) and relerr(xs2, xo2) < 1e-5


def test_parallel_option_rejected(sip):
    spec = pr.spec_config1((16, 16), np.float32)
    opt = sip.PARSDMM_options()
    opt.parallel = True
    with pytest.raises(NotImplementedError):
        pr.build(sip, spec, opt)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_parsdmm_config5_minkowski(sip, orc, TF):
    """Generalized Minkowski set (unknown [x1; x2], block operators, [I I] distance term).

    Float64 runs to the reference's stopping rule with identical iteration counts.  In Float32 this
    under-determined splitting amplifies rounding differences of the adaptation reductions (the device
    and the oracle agree bit-for-bit in all logged scalars for ~40 iterations, 1e-5 in x after 80, 1e-3
    after 150, measured), so the Float32 case compares a fixed 80 iterations."""
    oo, so = orc.PARSDMM_options(), sip.PARSDMM_options()
    ob = pr.build_minkowski(orc, (32, 28), TF, oo)
    sb = pr.build_minkowski(sip, (32, 28), TF, so)
    for o in (ob["opt"], sb["opt"]):
        if TF == np.float32:
            o.maxit, o.feas_tol, o.obj_tol, o.evol_rel_tol = 80, 1e-12, 1e-12, 1e-14
        else:
            o.maxit = 200
    m = ob["m"]
    xo, lo, ll, yy = orc.PARSDMM(m.copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
    xs, ls, l2, y2 = sip.PARSDMM(m.copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])
    assert xs.size == 2 * m.size == xo.size
    assert list(sb["AtA"]._device.q_offsets[:3]) == [0, -m.size, m.size] or m.size in np.abs(sb["AtA"]._device.q_offsets)
    check_parity((xo, lo, ll, yy, ob), (xs, ls, l2, y2, sb), TF)
    if TF == np.float64:
        assert len(ls.obj) < 200        # stopped by the reference's rules, same iteration as the oracle
    tot = xs[: m.size] + xs[m.size:]
    assert tot.min() >= 1500 - 50 and tot.max() <= 4500 + 50


@pytest.mark.parametrize("which", ["config4_bounds", "tv_l1"])
def test_parsdmm_multilevel(sip, orc, which):
    """PARSDMM_multi_level (config 4, examples/test_scaling_3D.jl:144-145): 3 levels, coarsening 2, warm
    starts of x, l, y through nearest-neighbour resampling and carried rho."""
    from oracle import multilevel as om
    TF = np.float32
    spec = pr.spec_config4((32, 24, 16), TF) if which == "config4_bounds" else pr.spec_config2((32, 24, 16), TF)
    res = []
    for api in (orc, sip):
        cg = api.compgrid(tuple(spec["d"]), tuple(spec["n"]))
        cons = [api.set_definitions(st, op, lo, hi, ("tensor", "")) for (st, op, lo, hi) in spec["sets"]]
        opt = api.PARSDMM_options()
        opt.FL = TF
        opt.evol_rel_tol = 10 * np.finfo(TF).eps
        opt.maxit = 40
        if which == "config4_bounds":
            opt.rho_ini = [1.0, 1000.0, 1000.0, 1000.0, 1.0]
        if api is orc:
            lv = om.setup_multi_level_PARSDMM(spec["m"], 3, 2, cg, cons, opt, orc.types)
            res.append(om.PARSDMM_multi_level(spec["m"].copy(), *lv[:5], opt))
        else:
            lv = sip.setup_multi_level_PARSDMM(spec["m"], 3, 2, cg, cons, opt)
            res.append(sip.PARSDMM_multi_level(spec["m"].copy(), *lv[:5], opt))                      # device resampling
            host = sip.PARSDMM_multi_level(spec["m"].copy(), *lv[:5], opt, device_resample=False)   # host resampling
            assert np.array_equal(host[0], res[-1][0]) and host[1].timing["level_iterations"] == res[-1][1].timing["level_iterations"]
            assert all(np.array_equal(a, b) for a, b in zip(host[3], res[-1][3]))
        assert [float(v) for v in opt.rho_ini] == ([1.0, 1000.0, 1000.0, 1000.0, 1.0] if which == "config4_bounds" else [10.0])
    (xo, lo, ll, yy), (xs, ls, l2, y2) = res
    assert [len(g.obj) for g in lo.levels] == ls.timing["level_iterations"]
    assert np.array_equal(ls.cg_it, lo.cg_it)
    assert relerr(xs, xo) < TOL[TF]
    for a, b in zip(y2, yy):
        assert relerr(a, b) < 100 * TOL[TF]


# ---------------------------------------------------------------------------------------------
# BASELINE-size checks through size-independent properties (the oracle would need minutes here)
# ---------------------------------------------------------------------------------------------
def test_full_size_config2_properties(sip):
    """configs[1] at its full size (200^3 Float32): operator adjointness and linearity, projector
    idempotence / feasibility, and every constraint set satisfied by the PARSDMM result to
    1.5*feas_tol (the reference's own acceptance test, test_PARSDMM.jl:86-89)."""
    TF = np.float32
    n = (200, 200, 200)
    spec = pr.spec_config2(n, TF)
    opt = sip.PARSDMM_options()
    opt.evol_rel_tol = 10 * float(np.finfo(TF).eps)
    sb = pr.build(sip, spec, opt)
    rng = np.random.default_rng(0)
    N = int(np.prod(n))
    TV = sb["TD_OP"][1]
    x1, x2 = rng.standard_normal(N).astype(TF), rng.standard_normal(N).astype(TF)
    v = rng.standard_normal(TV.rows).astype(TF)
    Ax1, Ax2 = TV @ x1, TV @ x2
    # <A x, v> == <x, A' v>
    lhs = float(np.dot(Ax1.astype(np.float64), v.astype(np.float64)))
    rhs = float(np.dot(x1.astype(np.float64), (TV.T @ v).astype(np.float64)))
    assert abs(lhs - rhs) <= 1e-5 * (abs(lhs) + abs(rhs) + 1.0)
    # linearity
    comb = TV @ (TF(2.0) * x1 - TF(0.5) * x2)
    assert relerr(comb, TF(2.0) * Ax1 - TF(0.5) * Ax2) < 1e-5
    # the CDS form of A'A applied by the SpMV kernel agrees with A'(A x)
    R, off = sb["AtA"][1], sb["set_Prop"].AtA_offsets[1]
    assert relerr(sip.CDS_MVp(N, R.shape[1], R, off, x1, np.zeros(N, dtype=TF)), TV.T @ Ax1) < 1e-5
    # projectors: idempotent, inside the set
    P_l1 = sb["P_sub"][1]
    assert np.array_equal(P_l1(Ax1.copy()), Ax1)                    # inside the ball: untouched
    big = (Ax1 * TF(100.0 * float(P_l1.max) / float(np.abs(Ax1.astype(np.float64)).sum()))).astype(TF)
    y = P_l1(big.copy())
    assert abs(float(np.abs(y.astype(np.float64)).sum()) - float(P_l1.max)) <= 1e-5 * float(P_l1.max)
    assert relerr(P_l1(y.copy()), y) < 1e-6
    # the projection itself
    x, log, _, _ = sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"],
                               return_ly=False)
    assert 5 < len(log.obj) < opt.maxit and log.cg_it[0] == 0
    for i in range(len(sb["P_sub"])):
        s = sb["TD_OP"][i] @ x
        ps = sb["P_sub"][i](s.copy())
        feas = np.linalg.norm(ps.astype(np.float64) - s) / np.linalg.norm(s.astype(np.float64))
        assert feas <= 1.5 * float(opt.feas_tol), (i, feas)
    assert abs(log.obj[-1] - 0.5 * np.linalg.norm(x.astype(np.float64) - spec["m"]) ** 2) <= 1e-3 * log.obj[-1]


# ---------------------------------------------------------------------------------------------
# fiber application modes: project_bounds!.jl:38-88, project_cardinality!.jl:23-113
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n,kind", [((50, 37), "identity"), ((50, 37), "D_x"), ((50, 37), "D_z"), ((23, 17, 11), "identity"),
                                    ((23, 17, 11), "D_y"), ((23, 17, 11), "D_z")])
def test_fiber_projectors_match_oracle(sip, orc, TF, n, kind):
    rng = np.random.default_rng(11)
    dirs = ("x", "z") if len(n) == 2 else ("x", "y", "z")
    for direction in dirs:
        for st in ("cardinality", "bounds"):
            P = {}
            for api in (orc, sip):
                cg = api.compgrid((1.0,) * len(n), n)
                TD_n = api.get_TD_operator(cg, kind, TF)[3]
                ax = dirs.index(direction)
                if st == "bounds":
                    lo = np.sort(rng.standard_normal(TD_n[ax])) - 0.5 if api is orc else lo
                    hi = lo + np.abs(rng.standard_normal(TD_n[ax])) if api is orc else hi
                    cons = [api.set_definitions("bounds", kind, lo.copy(), hi.copy(), ("fiber", direction))]
                else:
                    cons = [api.set_definitions("cardinality", kind, 0, 3, ("fiber", direction))]
                P[api] = api.setup_constraints(cons, cg, TF)[0][0]
            M = int(np.prod(TD_n))
            v = np.round(rng.standard_normal(M) * 2, 1).astype(TF)        # rounded: plenty of exact ties
            got, want = P[sip](v.copy()), P[orc](v.copy())
            assert np.array_equal(got, want), (kind, direction, st)
            if st == "cardinality":
                G = got.reshape(TD_n, order="F")
                assert np.all(np.count_nonzero(G, axis=dirs.index(direction)) <= 3)


def test_fiber_cardinality_counts_like_reference_test(sip):
    """test_setup_constraints.jl:108-135: exactly k non-zeros per column / row for random data."""
    n = (50, 100)
    X = np.random.default_rng(3).standard_normal(n)
    for direction, k, axis in (("x", 7, 0), ("z", 11, 1)):
        cons = [sip.set_definitions("cardinality", "identity", 0, k, ("fiber", direction))]
        P = sip.setup_constraints(cons, sip.compgrid((1.0, 1.0), n), np.float64)[0][0]
        Y = P(X.ravel(order="F").copy()).reshape(n, order="F")
        assert np.all(np.count_nonzero(Y, axis=axis) == k)


@pytest.mark.parametrize("TF", [np.float64, np.float32])
def test_parsdmm_with_fiber_sets(sip, orc, TF):
    """PARSDMM with a per-fiber cardinality set on D_x / D_z (examples/constrained_freq_FWI_simple.jl:286-302)
    and per-fiber bounds."""
    n, d = (40, 36), (25.0, 6.0)
    m = pr.synthetic_model(n, TF)
    lo = np.linspace(1400.0, 1600.0, n[1])
    hi = np.linspace(3000.0, 4700.0, n[1])
    res = []
    for api in (orc, sip):
        cg = api.compgrid(d, n)
        cons = [api.set_definitions("bounds", "identity", lo.copy(), hi.copy(), ("fiber", "z")),
                api.set_definitions("cardinality", "D_x", 0, 6, ("fiber", "x")),
                api.set_definitions("cardinality", "D_z", 0, 8, ("fiber", "z"))]
        opt = api.PARSDMM_options()
        opt.FL, opt.maxit = TF, 30
        P_sub, TD_OP, set_Prop = api.setup_constraints(cons, cg, TF)
        TD_OP, AtA, l, y = api.PARSDMM_precompute_distribute(TD_OP, set_Prop, cg, opt)
        res.append(api.PARSDMM(m.copy(), AtA, TD_OP, set_Prop, P_sub, cg, opt) + (set_Prop,))
    (xo, lo_, ll, yy, spo), (xs, ls, l2, y2, sps) = res
    assert spo.ncvx == sps.ncvx == [False, True, True, False]
    assert len(ls.obj) == len(lo_.obj) and np.array_equal(ls.cg_it, lo_.cg_it)
    assert relerr(xs, xo) < TOL[TF]
    assert np.array_equal(y2[1] != 0, yy[1] != 0) and np.array_equal(y2[2] != 0, yy[2] != 0)     # supports bit exact
    assert np.allclose(ls.set_feasibility, lo_.set_feasibility, rtol=50 * TOL[TF], atol=1e-12)


# ---------------------------------------------------------------------------------------------
# feasibility problems and error behaviour of the boundary
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("TF", [np.float64, np.float32])
def test_parsdmm_feasibility_only(sip, orc, TF):
    """options.feasibility_only = true drops the distance term (PARSDMM_precompute_distribute.jl:17,
    PARSDMM.jl:55-56, update_y_l.jl:90): every operator is a constraint set, p == pp."""
    spec = pr.spec_config1((36, 30), TF)
    def tw(o):
        o.feasibility_only = True
        o.maxit = 60
    o, s = run_both(sip, orc, spec, tw)
    assert len(s[4]["TD_OP"]) == len(spec["sets"]) and s[1].rho.shape[1] == len(spec["sets"])
    check_parity(o, s, TF)


def test_boundary_error_behaviour(sip):
    """Errors surface as exceptions with the library's message, never as silent fallbacks."""
    import ctypes as C
    L = sip._lib
    lib = L.load()
    spec = pr.spec_config1((16, 16), np.float32)
    b = pr.build(sip, spec)
    # wrong length of m
    with pytest.raises(ValueError):
        sip.PARSDMM(spec["m"][:-1].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
    # P_sub / TD_OP mismatch
    with pytest.raises(ValueError):
        sip.PARSDMM(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"][:-1], b["cg"], b["opt"])
    # complex input (PARSDMM.jl:50-52)
    with pytest.raises((TypeError, ValueError)):
        sip.PARSDMM(spec["m"].astype(np.complex64), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
    # C ABI: unsupported set kind and call-order violations return codes + messages
    pb = C.c_void_p()
    n = (C.c_int64 * 3)(8, 8, 1)
    h = (C.c_double * 3)(1.0, 1.0, 1.0)
    assert lib.sipb_problem_create(L.ctx(), L.SIPB_F32, 2, n, h, 0, 0, C.byref(pb)) == 0
    d = L.SetDesc()
    d.set_kind = 42
    assert lib.sipb_problem_add_set(pb, C.byref(d)) == L.SIPB_E_UNSUPPORTED and b"outside the device hot path" in lib.sipb_last_error()
    assert lib.sipb_problem_finalize(pb) == L.SIPB_E_INVALID          # no sets
    d.set_kind, d.op_kind = L.SET_L1, L.OP_TV
    d.max = -1.0
    assert lib.sipb_problem_add_set(pb, C.byref(d)) == L.SIPB_E_INVALID and b"Radius of L1 ball" in lib.sipb_last_error()
    d.max = 1.0
    assert lib.sipb_problem_add_set(pb, C.byref(d)) == 0
    assert lib.sipb_problem_finalize(pb) == L.SIPB_E_STATE            # AtA missing
    lib.sipb_problem_destroy(pb)
    # a diagonal of AtA missing in Q: CDS_scaled_add!.jl:18-20
    R = np.ones((16, 1), dtype=np.float32, order="F")
    with pytest.raises(L.SipbError) as e:
        sip.CDS_scaled_add(R.copy(order="F"), np.ones((16, 1), dtype=np.float32, order="F"), np.array([0]), np.array([1]), 1.0)
    assert e.value.code == L.SIPB_E_MISSING_DIAG
    # 3-D-only operator on a 2-D grid
    with pytest.raises(ValueError):
        sip.get_TD_operator(sip.compgrid((1.0, 1.0), (8, 8)), "D_y", np.float32)


# ---------------------------------------------------------------------------------------------
# the two device forms of Q (CDS arrays / stencil-class tables) are the same arithmetic
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["config1_f64", "config2_f32", "odd_grid_f32", "minkowski_f64"])
def test_q_class_tables_bit_identical_to_arrays(sip, case, monkeypatch):
    """sipb_problem_q_form: every AtA of get_TD_operator.jl has one value per stencil class, so the solver keeps
    class tables; SIPB_Q_CLASSES=0 forces the CDS arrays of the reference (Q_update!.jl:45-49, cg.jl:84).
    Both run the same multiply-adds in the same order: x, l, y and every log entry must agree bit for bit."""
    def solve():
        opt = sip.PARSDMM_options()
        if case == "minkowski_f64":
            b = pr.build_minkowski(sip, (32, 28), np.float64, opt)
            b["opt"].maxit = 60
            m = b["m"]
        else:
            spec = {"config1_f64": pr.spec_config1((48, 40), np.float64), "config2_f32": pr.spec_config2((24, 20, 16), np.float32),
                    "odd_grid_f32": pr.spec_config2((7, 5, 3), np.float32)}[case]
            b = pr.build(sip, copy.deepcopy(spec), opt)
            m = spec["m"]
        x, log, l, y = sip.PARSDMM(m.copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
        return x, log, l, y, b["AtA"]._device.q_form

    xa, la, l_a, y_a, form_a = solve()
    monkeypatch.setenv("SIPB_Q_CLASSES", "0")
    xb, lb, l_b, y_b, form_b = solve()
    assert form_a == "classes" and form_b == "arrays"
    assert np.array_equal(xa, xb)
    assert np.array_equal(la.cg_it, lb.cg_it) and np.array_equal(la.obj, lb.obj)
    assert np.array_equal(la.cg_relres, lb.cg_relres, equal_nan=True) and np.array_equal(la.rho, lb.rho)
    for a, b in zip(l_a + y_a, l_b + y_b):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_q_class_detection_rejects_irregular_matrix(sip):
    """An AtA that is not constant per stencil class (one perturbed entry) must stay in array form."""
    spec = pr.spec_config2((12, 10, 8), np.float32)
    b = pr.build(sip, copy.deepcopy(spec), sip.PARSDMM_options())
    b["AtA"][1][137, 0] *= np.float32(1.5)
    x, log, l, y = sip.PARSDMM(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
    assert b["AtA"]._device.q_form == "arrays"
    assert np.all(np.isfinite(x))


# ---------------------------------------------------------------------------------------------
# custom (explicit sparse) transform-domain operators   (setup_constraints.jl:70-72)
# ---------------------------------------------------------------------------------------------
def _weighted_gradient(orc, n, d, TF, seed=5):
    """A user-style operator: the discrete gradient with smoothly varying row weights (banded A'A, values that
    are NOT constant per stencil class)."""
    import scipy.sparse as sps
    A, *_ = orc.get_TD_operator(orc.compgrid(d, n), "TV", TF)
    w = (1.0 + 0.5 * np.sin(np.arange(A.shape[0]) * 0.37 + seed)).astype(TF)
    W = sps.csc_matrix(sps.diags(w).astype(TF) @ A).astype(TF)
    W.sort_indices()
    return W


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_sparse_operator_apply_bit_exact(sip, orc, TF):
    n, d = (13, 9, 5), (2.0, 3.0, 1.5)
    W = _weighted_gradient(orc, n, d, TF)
    op = sip.SparseOperator(W, n, d, TF)
    assert op.shape == W.shape
    x = pr.splitmix_uniform(3, W.shape[1]).astype(TF)
    v = pr.splitmix_uniform(4, W.shape[0]).astype(TF)
    assert np.array_equal(op @ x, orc.ops.spmv(W, x))
    assert np.array_equal(op.T @ v, orc.ops.spmv_t(W, v))
    with pytest.raises(ValueError):
        op @ v


@pytest.mark.parametrize("TF", [np.float64, np.float32])
@pytest.mark.parametrize("order", ["custom_middle", "custom_first"])
def test_parsdmm_custom_sparse_operator(sip, orc, TF, order):
    """examples/ConstraintSetupExamples.jl:125-146: l1 constraint in the domain of a user-supplied sparse matrix,
    flagged by hand (AtA_diag=false, dense=false, banded=true), next to stencil-operator sets."""
    n, d = (32, 28), (25.0, 6.0)
    m = pr.synthetic_model(n, TF).ravel(order="F")
    W = _weighted_gradient(orc, n, d, TF)
    tau = 0.4 * float(np.abs(W.astype(np.float64) @ m.astype(np.float64)).sum())

    def build(api):
        cg = api.compgrid(d, n)
        sets = [api.set_definitions("bounds", "identity", 1500.0, 4500.0, ("matrix", "")),
                api.set_definitions("l1", "identity", 0.0, tau, ("matrix", ""), (W.copy(), False)),
                api.set_definitions("bounds", "D_z", 0.0, 1e6, ("matrix", ""))]
        if order == "custom_first":
            sets = [sets[1], sets[0], sets[2]]
        ic = 1 if order == "custom_middle" else 0
        opt = api.PARSDMM_options()
        opt.FL = TF
        P_sub, TD_OP, set_Prop = api.setup_constraints(sets, cg, TF)
        set_Prop.AtA_diag[ic], set_Prop.dense[ic], set_Prop.banded[ic] = False, False, True
        TD_OP, AtA, l, y = api.PARSDMM_precompute_distribute(TD_OP, set_Prop, cg, opt)
        return dict(cg=cg, opt=opt, P_sub=P_sub, TD_OP=TD_OP, set_Prop=set_Prop, AtA=AtA)

    ob, sb = build(orc), build(sip)
    assert np.array_equal(ob["set_Prop"].AtA_offsets[0], sb["set_Prop"].AtA_offsets[0])
    xo, lo, ll, yy = orc.PARSDMM(m.copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
    xs, ls, l2, y2 = sip.PARSDMM(m.copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])
    assert sb["AtA"]._device.q_form == "arrays"           # weighted rows: not constant per stencil class
    check_parity((xo, lo, ll, yy, ob), (xs, ls, l2, y2, sb), TF)
    assert len(ls.obj) > 5
    s = W @ xs
    assert np.abs(s).sum() <= tau * (1 + 50 * ls.set_feasibility[-2].max() + 1e-3)


# ---------------------------------------------------------------------------------------------
# slice-mode cardinality on 3-D tensors   (project_cardinality!.jl:115-146, test_projectors.jl:81-93)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("TF", [np.float64, np.float32])
def test_slice_cardinality_projector(sip, orc, TF):
    n = (13, 9, 7)
    X = np.random.default_rng(11).standard_normal(n).astype(TF)
    X[2, :, :] = np.sign(X[2, :, :])            # slices full of ties: the stable order decides
    X[:, 4, :] = TF(0.5) * np.sign(X[:, 4, :])
    X[:, :, 3] = TF(2.0)
    for direction, axis, k in (("x", 0, 7), ("y", 1, 6), ("z", 2, 5)):
        outs = []
        for api in (orc, sip):
            cons = [api.set_definitions("cardinality", "identity", 0, k, ("slice", direction))]
            P = api.setup_constraints(cons, api.compgrid((1.0, 1.0, 1.0), n), TF)[0][0]
            v = X.ravel(order="F").copy()
            out = P(v)
            outs.append((out, v))
        (oo, vo), (so, vs) = outs
        assert np.array_equal(so, oo)                                        # values and support bit exact (stable ties)
        assert np.all(np.count_nonzero(so.reshape(n, order="F"), axis=tuple(a for a in range(3) if a != axis)) == k)
        # x / y slices: the reference works on a permuted copy (input untouched); z: in place
        assert np.array_equal(vs, vo) and (np.array_equal(vs, X.ravel(order="F")) == (direction != "z"))


def test_parsdmm_with_slice_cardinality(sip, orc):
    """PARSDMM with bounds and a per-x-slice cardinality set on D_z.  The in-loop feasibility of an x / y slice set
    is identically 0 in the reference (its projector does not mutate the argument update_y_l.jl:93 relies on):
    reproduced by oracle and device."""
    TF = np.float64
    n, d = (12, 10, 8), (25.0, 25.0, 10.0)
    m = pr.synthetic_model(n, TF)
    res = []
    for api in (orc, sip):
        cg = api.compgrid(d, n)
        cons = [api.set_definitions("bounds", "identity", 1500.0, 4600.0, ("tensor", "")),
                api.set_definitions("cardinality", "D_z", 0, 25, ("slice", "x")),
                api.set_definitions("cardinality", "D_x", 0, 30, ("slice", "z"))]
        opt = api.PARSDMM_options()
        opt.FL, opt.maxit = TF, 40
        P_sub, TD_OP, set_Prop = api.setup_constraints(cons, cg, TF)
        TD_OP, AtA, l, y = api.PARSDMM_precompute_distribute(TD_OP, set_Prop, cg, opt)
        res.append(api.PARSDMM(m.copy(), AtA, TD_OP, set_Prop, P_sub, cg, opt))
    (xo, lo_, ll, yy), (xs, ls, l2, y2) = res
    assert len(ls.obj) == len(lo_.obj) and np.array_equal(ls.cg_it, lo_.cg_it)
    assert relerr(xs, xo) < TOL[TF]
    assert np.array_equal(y2[1] != 0, yy[1] != 0) and np.array_equal(y2[2] != 0, yy[2] != 0)
    assert np.allclose(ls.set_feasibility, lo_.set_feasibility, rtol=50 * TOL[TF], atol=1e-12)
    assert ls.set_feasibility[0, 1] > 0 and np.all(ls.set_feasibility[1:-1, 1] == 0)        # the quirk
    yz = y2[2].reshape((n[0] - 1, n[1], n[2]), order="F")
    assert np.all(np.count_nonzero(yz, axis=(0, 1)) <= 30)


# ---------------------------------------------------------------------------------------------
# BASELINE-size parity: the device against the CPU oracle at the grid sizes BASELINE.json names
# (the threaded C/OpenMP port of the oracle where it applies: tests/test_cpu_baseline.py pins it to the NumPy oracle)
# ---------------------------------------------------------------------------------------------
def _cpu_parsdmm(orc, spec, tweak):
    from oracle import cpu_baseline as cb
    cb.use_all_cores()
    opt = orc.PARSDMM_options()
    tweak(opt)
    ob = pr.build(orc, copy.deepcopy(spec), opt)
    return cb.PARSDMM(spec["m"].copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"], constraint=ob["cons"])


def _dev_parsdmm(sip, spec, tweak):
    opt = sip.PARSDMM_options()
    tweak(opt)
    sb = pr.build(sip, copy.deepcopy(spec), opt)
    return sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])


def _same_run(dev, cpu, TF):
    xs, ls, l2, y2 = dev
    xo, lo, ll, yy = cpu
    assert len(ls.obj) == len(lo.obj), (len(ls.obj), len(lo.obj))
    assert np.array_equal(ls.cg_it, lo.cg_it)
    assert relerr(xs, xo) < TOL[TF], relerr(xs, xo)
    assert np.allclose(ls.set_feasibility, lo.set_feasibility, rtol=50 * TOL[TF], atol=1e-12)
    assert np.allclose(ls.rho, lo.rho, rtol=50 * TOL[TF]) and np.allclose(ls.obj, lo.obj, rtol=50 * TOL[TF])
    for a, b in zip(y2, yy):
        assert relerr(a, b) < 100 * TOL[TF]


def test_baseline_size_config1_256x256_f64(sip, orc):
    """BASELINE configs[0]: 2-D 256x256 Float64, bounds ∩ TV l1-ball ∩ vertical slope bounds, to the stopping rules."""
    def tw(o):
        o.maxit = 500
    spec = pr.spec_config1((256, 256), np.float64)
    _same_run(_dev_parsdmm(sip, spec, tw), _cpu_parsdmm(orc, spec, tw), np.float64)


def test_baseline_size_config2_200cubed_f32(sip, orc):
    """BASELINE configs[1]: 3-D 200^3 Float32, bounds ∩ anisotropic TV ∩ lateral smoothness, to the stopping rules."""
    def tw(o):
        o.maxit, o.evol_rel_tol = 200, 10 * float(np.finfo(np.float32).eps)
    spec = pr.spec_config2((200, 200, 200), np.float32)
    _same_run(_dev_parsdmm(sip, spec, tw), _cpu_parsdmm(orc, spec, tw), np.float32)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n,frac,seed", [((72, 60), 0.4, 5), ((96, 64), 0.3, 6), ((333, 257), 0.55, 7), ((640, 512), 0.4, 8),
                                         ((640, 512), 0.12, 9)])
def test_parsdmm_cardinality_ties_in_the_loop(sip, orc, TF, n, frac, seed):
    """Threshold ties inside the y/l update: with an integer-valued model and Q a multiple of the identity the CG scales
    every entry alike, so in iteration 2 the k-th largest magnitude is shared by hundreds to tens of thousands of rows
    (708 ties for a quota of 679 at 72 x 60; 52936 for 51431 at 640 x 512) and the stable order of sortperm
    (project_cardinality!.jl:18-19) decides the support — across many row chunks of the select kernels (speculative
    levels in pass 1, per-chunk tie prefix applied by pass 2, the chunk on the quota boundary settled in place).
    Two iterations only: from iteration 3 on, classes that are EQUAL in exact arithmetic are ordered by rounding, and the
    CPU restatements themselves (NumPy f64acc / native reductions, C port) end in different supports
    (tests/checks/tie_sensitivity.py) — nothing to compare against."""
    rng = np.random.default_rng(seed)
    N = int(np.prod(n))
    m = np.round(rng.standard_normal(N) * 3).astype(TF)
    k = int(frac * N)
    spec = dict(n=n, d=(1.0, 1.0), TF=TF, m=m, sets=[("cardinality", "identity", 0, k)], mode="matrix")
    def tw(o):
        o.maxit = 2
    o, s = run_both(sip, orc, spec, tw)
    assert np.count_nonzero(o[3][0]) == k                          # the ties were cut at the quota, not kept wholesale
    assert np.array_equal(o[0] != 0, s[0] != 0)                   # support of x
    for a, b in zip(s[3], o[3]):
        assert np.array_equal(a != 0, b != 0)                     # support of every y
    check_parity(o, s, TF)


def test_config3_128cubed_f32(sip, orc):
    """BASELINE configs[2] (bounds ∩ TV l1 ∩ cardinality of the gradient) at 128^3, 30 iterations (the CPU oracle's
    sparse set-up and stable sort of 512^3 take tens of minutes; tests/checks/parity_fullsize.py runs larger grids).

    The cardinality set is non-convex: an entry whose magnitude sits at the k-th largest value flips in or out of the
    support on a one-ulp change of its input.  Measured at this size (tests/checks/config3_divergence.py): device and CPU runs are
    BIT-IDENTICAL in every logged scalar for 15 iterations; at iteration 16 one rho differs by one ulp (an adaptation
    sum rounds differently) and the supports drift apart.  After 30 iterations x differs by 1.20e-3 (device vs the
    C/OpenMP port), 1.13e-3 (device vs the NumPy oracle) and 0.91e-3 (the two CPU restatements AGAINST EACH OTHER),
    with 2.1 % / 1.9 % / 1.2 % of the k support entries different: no pair of faithful Float32 implementations meets
    1e-3 here.  The test pins what is stable: iteration and CG counts, exactly k survivors, x to 3e-3, supports to
    4 % (bit-identical x at 96^3 / 25 iterations: bench.py's parity object)."""
    def tw(o):
        o.maxit, o.evol_rel_tol = 30, 10 * float(np.finfo(np.float32).eps)
    spec = pr.spec_config3((128, 128, 128), np.float32)
    (xs, ls, l2, y2), (xo, lo, ll, yy) = _dev_parsdmm(sip, spec, tw), _cpu_parsdmm(orc, spec, tw)
    assert len(ls.obj) == len(lo.obj) and np.array_equal(ls.cg_it, lo.cg_it)
    assert relerr(xs, xo) < 3e-3, relerr(xs, xo)
    assert np.allclose(ls.obj, lo.obj, rtol=5e-2)
    k = np.count_nonzero(yy[2])
    assert np.count_nonzero(y2[2]) == k                              # exactly k entries survive on both sides
    assert np.count_nonzero((y2[2] != 0) != (yy[2] != 0)) <= 4e-2 * k


def test_baseline_size_config4_multilevel_100cubed(sip, orc):
    """BASELINE configs[3]-style multilevel PARSDMM (3 levels, coarsening 2) with the constraints of
    examples/test_scaling_3D.jl:41-74 at 100^3 -> 50^3 -> 25^3 (400^3 needs minutes of NumPy oracle time)."""
    from oracle import multilevel as om
    TF = np.float32
    spec = pr.spec_config4((100, 100, 100), TF)
    res = []
    for api in (orc, sip):
        cg = api.compgrid(tuple(spec["d"]), tuple(spec["n"]))
        cons = [api.set_definitions(st, op, lo, hi, ("tensor", "")) for (st, op, lo, hi) in spec["sets"]]
        opt = api.PARSDMM_options()
        opt.FL, opt.evol_rel_tol, opt.maxit = TF, 10 * np.finfo(TF).eps, 40
        opt.rho_ini = [1.0, 1000.0, 1000.0, 1000.0, 1.0]
        if api is orc:
            lv = om.setup_multi_level_PARSDMM(spec["m"], 3, 2, cg, cons, opt, orc.types)
            res.append(om.PARSDMM_multi_level(spec["m"].copy(), *lv[:5], opt))
        else:
            lv = sip.setup_multi_level_PARSDMM(spec["m"], 3, 2, cg, cons, opt)
            res.append(sip.PARSDMM_multi_level(spec["m"].copy(), *lv[:5], opt))
    (xo, lo, ll, yy), (xs, ls, l2, y2) = res
    assert [len(g.obj) for g in lo.levels] == ls.timing["level_iterations"]
    assert np.array_equal(ls.cg_it, lo.cg_it)
    assert relerr(xs, xo) < TOL[TF]


def test_baseline_size_config5_minkowski_1024x1024_f32(sip, orc):
    """BASELINE configs[4]: generalized Minkowski set at 1024x1024 Float32 with the options of
    example_2D_Minkowski_projection.jl:25-29, to the reference's stopping rules: the device and the oracle must stop at
    the same iteration with x inside the Float32 tolerance."""
    TF = np.float32
    oo, so = orc.PARSDMM_options(), sip.PARSDMM_options()
    ob = pr.build_minkowski(orc, (1024, 1024), TF, oo)
    sb = pr.build_minkowski(sip, (1024, 1024), TF, so)
    for o in (ob["opt"], sb["opt"]):
        o.maxit = 500
    m = ob["m"]
    xs, ls, l2, y2 = sip.PARSDMM(m.copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])
    xo, lo, ll, yy = orc.PARSDMM(m.copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
    assert len(ls.obj) == len(lo.obj), (len(ls.obj), len(lo.obj))
    assert len(ls.obj) < 500
    assert np.array_equal(ls.cg_it, lo.cg_it)
    assert relerr(xs, xo) < TOL[TF], relerr(xs, xo)


# ---------------------------------------------------------------------------------------------
# batched independent projections: examples/Constraint_examples_2D.jl:221-226 (one PARSDMM per RGB
# channel), examples/Dykstra_parallel_vs_PARSDMM.jl:134,149 (PARSDMM as an inner projector)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("TF", [np.float64, np.float32])
def test_parsdmm_batch_matches_independent_solves(sip, orc, TF):
    n = (48, 40)
    B = 5
    specs = []
    for b in range(B):
        sp_ = pr.spec_config1(n, TF)
        sp_["m"] = pr.synthetic_model(n, TF, seed=100 + b)
        # every "channel" has its own bounds and TV budget
        sp_["sets"] = [("bounds", "identity", 1500.0 + 20 * b, 4500.0 - 15 * b),
                       ("l1", "TV", 0.0, (0.4 + 0.05 * b) * pr.tv_l1(n, sp_["d"], TF, sp_["m"])), ("bounds", "D_z", 0.0, 1e6)]
        specs.append(sp_)
    opt = sip.PARSDMM_options()
    builds = [pr.build(sip, copy.deepcopy(sp_), copy.deepcopy(opt)) for sp_ in specs]
    res = sip.PARSDMM_batch([sp_["m"].copy() for sp_ in specs], [b["AtA"] for b in builds], [b["TD_OP"] for b in builds],
                            [b["set_Prop"] for b in builds], [b["P_sub"] for b in builds], builds[0]["cg"], builds[0]["opt"])
    assert len(res) == B
    for b in range(B):
        # bit for bit the result of a separate device call ...
        sb = pr.build(sip, copy.deepcopy(specs[b]), sip.PARSDMM_options())
        xs, ls, l2, y2 = sip.PARSDMM(specs[b]["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])
        xb, lb, lb2, yb2 = res[b]
        assert np.array_equal(xb, xs) and np.array_equal(lb.cg_it, ls.cg_it) and np.array_equal(lb.obj, ls.obj)
        assert all(np.array_equal(a, c) for a, c in zip(lb2 + yb2, l2 + y2))
        # ... and within tolerance of the oracle's independent solve
        ob = pr.build(orc, copy.deepcopy(specs[b]), orc.PARSDMM_options())
        xo, lo, _, _ = orc.PARSDMM(specs[b]["m"].copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
        assert len(lb.obj) == len(lo.obj) and np.array_equal(lb.cg_it, lo.cg_it)
        assert relerr(xb, xo) < TOL[TF]
    # one shared problem definition, several models; a second call reuses the device problems
    shared = builds[0]
    ms = [pr.synthetic_model(n, TF, seed=300 + b) for b in range(3)]
    for _ in range(2):
        out = sip.PARSDMM_batch(ms, shared["AtA"], shared["TD_OP"], shared["set_Prop"], shared["P_sub"], shared["cg"], shared["opt"],
                                return_ly=False)
        for b in range(3):
            xs, ls, _, _ = sip.PARSDMM(ms[b].copy(), shared["AtA"], shared["TD_OP"], shared["set_Prop"], shared["P_sub"], shared["cg"],
                                       shared["opt"])
            assert np.array_equal(out[b][0], xs) and len(out[b][1].obj) == len(ls.obj)


def test_multilevel_feasible_coarse_level(sip, orc):
    """A coarse level that is already feasible returns x = m_coarse, l = y = 0 (PARSDMM.jl:63-82) and must leave exactly
    that resident on the device, because the device-side multilevel warm start (sipb_problem_warm_from) reads x, l, y of
    the coarse problem from its device buffers.  The slope bound holds on the coarse grid and fails on the fine one.

    (Through PARSDMM_multi_level itself this situation ends in NaN — in the reference too: the one-row log of a feasible
    level has rho = 0 and PARSDMM_multi_level.jl:57 carries it into the next level, where update_y_l.jl:34 divides by it;
    oracle and device agree on that, first assertion.  The warm start is therefore also exercised by hand with the
    original rho_ini.)"""
    from oracle import multilevel as om
    from sip_b200 import multilevel as sm
    from sip_b200.solver import device_problem
    TF = np.float32
    n = (32, 24, 16)
    spec = pr.spec_config4(n, TF)
    # a component alternating from plane to plane: steep on the fine grid, nearly invisible after nearest-neighbour coarsening
    m = spec["m"].reshape(n, order="F").astype(np.float64) + 200.0 * ((-1.0) ** np.arange(n[2]))[None, None, :]
    m = np.ascontiguousarray(m.ravel(order="F").astype(TF))
    sets = [("bounds", "identity", 0.0, 1e5), ("bounds", "D_z", -100.0, 100.0)]
    cg0 = sip.compgrid(tuple(spec["d"]), n)
    cons0 = [sip.set_definitions(st, op, lo, hi, ("tensor", "")) for (st, op, lo, hi) in sets]
    lv0 = sip.setup_multi_level_PARSDMM(m, 2, 2, cg0, cons0, sip.PARSDMM_options())
    cgc = lv0[4][1]
    mc = sip.resample_nn(m, n, cgc.n).reshape(tuple(cgc.n), order="F").astype(np.float64)
    coarse = np.abs(np.diff(mc, axis=2)).max() / float(cgc.d[2])
    bound = 1.001 * coarse           # just feasible on the coarse grid; ~40 % relative infeasibility on the fine one
    sets[1] = ("bounds", "D_z", -bound, bound)

    def setup(api):
        cg = api.compgrid(tuple(spec["d"]), n)
        cons = [api.set_definitions(st, op, lo, hi, ("tensor", "")) for (st, op, lo, hi) in sets]
        opt = api.PARSDMM_options()
        opt.FL, opt.maxit = TF, 40
        if api is orc:
            return om.setup_multi_level_PARSDMM(m, 2, 2, cg, cons, opt, orc.types), opt
        return sip.setup_multi_level_PARSDMM(m, 2, 2, cg, cons, opt), opt

    # 1. the drivers themselves: the reference's rho = 0 carry-over makes both NaN from the second level on
    lvo, oo = setup(orc)
    xo, lo, _, _ = om.PARSDMM_multi_level(m.copy(), *lvo[:5], oo)
    lvs, so = setup(sip)
    xs, ls, _, _ = sip.PARSDMM_multi_level(m.copy(), *lvs[:5], so)
    assert ls.timing["levels"][0]["stopped_feasible"] and [len(g.obj) for g in lo.levels] == ls.timing["level_iterations"]
    assert np.isnan(xo).all() and np.isnan(xs).all()

    # 2. the warm start by hand with the original rho_ini: device-resident path == host path == oracle
    TD, AT, PS, SP, CG = lvs[:5]
    m_c = sip.resample_nn(m, n, CG[1].n)
    so.zero_ini_guess = True
    xc, lc, l_c, y_c = sip.PARSDMM(m_c, AT[1], TD[1], SP[1], PS[1], CG[1], so)
    assert lc.timing["stopped_feasible"] and np.array_equal(xc, m_c) and all(not v.any() for v in l_c + y_c)
    so.zero_ini_guess = False
    devs = [device_problem(m.dtype, AT[q], TD[q], SP[q], PS[q], CG[q], so) for q in (0, 1)]
    sm._device_warm_start(devs[0], devs[1], sm.warm_start_segments(SP, CG, True, 0))
    x_dev, log_dev, l_dev, y_dev = sip.PARSDMM(m, AT[0], TD[0], SP[0], PS[0], CG[0], so, warm_resident=True)
    x0 = sip.resample_nn(xc, CG[1].n, CG[0].n)
    l0, y0 = sip.interpolate_y_l(l_c, y_c, SP, CG, True, 0)
    x_host, log_host, l_host, y_host = sip.PARSDMM(m, AT[0], TD[0], SP[0], PS[0], CG[0], so, x0.copy(), l0, y0)
    assert np.isfinite(x_dev).all() and not log_dev.timing["stopped_feasible"]
    assert np.array_equal(x_dev, x_host) and np.array_equal(log_dev.cg_it, log_host.cg_it)
    assert all(np.array_equal(a, b) for a, b in zip(l_dev + y_dev, l_host + y_host))
    TDo, ATo, PSo, SPo, CGo = lvo[:5]
    oo.zero_ini_guess = False
    x_or, log_or, _, _ = orc.PARSDMM(m.copy(), ATo[0], TDo[0], SPo[0], PSo[0], CGo[0], oo, x0.copy(),
                                     [np.zeros(v.size, dtype=TF) for v in l0], [np.zeros(v.size, dtype=TF) for v in y0])
    assert len(log_dev.obj) == len(log_or.obj) and np.array_equal(log_dev.cg_it, log_or.cg_it)
    assert relerr(x_dev, x_or) < TOL[TF]


def test_device_side_loops_bit_identical_to_host_loops(sip):
    """The CG iteration and the l1 threshold search as CUDA-graph WHILE nodes (default) against the host-driven loops
    (SIPB_GRAPH_LOOPS=0; a fresh context reads the switch): the same kernels run in the same order, so x, l, y and every
    log entry agree bit for bit; and the graph path needs far fewer host launches."""
    import ctypes as C
    import os
    L = sip._lib
    out = {}
    for mode in ("graph", "host"):
        os.environ["SIPB_GRAPH_LOOPS"] = "1" if mode == "graph" else "0"
        try:
            h = C.c_void_p()
            L.check(L.load().sipb_ctx_create(0, C.byref(h)))
        finally:
            os.environ.pop("SIPB_GRAPH_LOOPS", None)
        old = L._ctx.get(0)
        L._ctx[0] = h
        try:
            res = []
            for spec in (pr.spec_config2((24, 20, 16), np.float32), pr.spec_config1((48, 40), np.float64),
                         pr.spec_config3((16, 12, 10), np.float32)):
                opt = sip.PARSDMM_options()
                opt.maxit = 40
                b = pr.build(sip, copy.deepcopy(spec), opt)
                x, log, l, y = sip.PARSDMM(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
                res.append((x, log, l, y))
                del b
            out[mode] = res
        finally:
            L._ctx[0] = old
            import gc
            gc.collect()
            L.load().sipb_ctx_destroy(h)
    for (xa, la, l_a, y_a), (xb, lb, l_b, y_b) in zip(out["graph"], out["host"]):
        assert np.array_equal(xa, xb) and np.array_equal(la.cg_it, lb.cg_it) and np.array_equal(la.obj, lb.obj)
        assert np.array_equal(la.cg_relres, lb.cg_relres, equal_nan=True) and np.array_equal(la.rho, lb.rho)
        assert all(np.array_equal(a, b) for a, b in zip(l_a + y_a, l_b + y_b))
    # config 2 at 24x20x16: one graph launch per x-minimisation and per l1 search instead of a launch per kernel
    assert out["graph"][0][1].timing["total_launches"] > 0
