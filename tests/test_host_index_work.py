"""CPU tests of the host-side index work of the product (bit-exact against the oracle) and of the
C-ABI library's exported surface.  No CUDA compute is invoked."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import problems as pr

orc = pr.OracleAPI()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GRIDS = [((9, 6), (0.99, 1.123)), ((30, 20), (25.0, 25.0)), ((4, 6, 5), (0.99, 1.123, 1.0)),
         ((13, 7, 9), (25.0, 12.5, 6.0)), ((5, 4, 1), (1.0, 2.0, 3.0))]


def same_csc(A, B):
    A, B = sp.csc_matrix(A), sp.csc_matrix(B)
    A.sort_indices()
    B.sort_indices()
    return (A.shape == B.shape and A.dtype == B.dtype and np.array_equal(A.indptr, B.indptr) and
            np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data))


@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n,d", GRIDS)
def test_operators_structure_and_values_bit_exact(sip, TF, n, d):
    """get_TD_operator / get_discrete_Grad: same CSC structure and values as the Kronecker construction."""
    kinds = ["identity", "D_x", "D_z", "TV"] + (["D_y"] if (len(n) == 3 and n[2] > 1) else ["D_xz"])
    for kind in kinds:
        A, diag_o, dense_o, TDn_o, band_o = orc.get_TD_operator(orc.compgrid(d, n), kind, TF)
        op, diag_s, dense_s, TDn_s, band_s = sip.get_TD_operator(sip.compgrid(d, n), kind, TF)
        assert same_csc(op.tosparse(), A), kind
        assert (diag_o, dense_o, tuple(TDn_o), band_o) == (diag_s, dense_s, tuple(TDn_s), band_s)
        assert op.shape == A.shape


@pytest.mark.parametrize("TF", [np.float32, np.float64])
@pytest.mark.parametrize("n,d", GRIDS)
def test_ata_cds_bit_exact(sip, TF, n, d):
    """A'A directly in CDS form == mat2CDS(A'*A): offsets (integer work) and values bit-exact."""
    kinds = ["identity", "D_x", "D_z", "TV"] + (["D_y"] if (len(n) == 3 and n[2] > 1) else ["D_xz"])
    for kind in kinds:
        A = orc.get_TD_operator(orc.compgrid(d, n), kind, TF)[0]
        Ro, oo = orc.mat2CDS(orc.ops.AtA_sparse(A))
        Rs, os_ = sip.get_TD_operator(sip.compgrid(d, n), kind, TF)[0].ata_cds()
        assert np.array_equal(oo, os_) and os_.dtype == np.int64, kind
        assert Rs.dtype == TF and np.array_equal(Ro, Rs), kind
        R2, o2 = sip.mat2CDS(orc.ops.AtA_sparse(A))
        assert np.array_equal(o2, oo) and np.array_equal(R2, Ro)


def test_mat2cds_random_and_stored_zeros(sip):
    A = sp.random(60, 60, density=0.05, random_state=3, format="csc")
    A.data[::7] = 0.0                      # stored zeros still define diagonals (findnz semantics)
    Ro, oo = orc.mat2CDS(A)
    Rs, os_ = sip.mat2CDS(A)
    assert np.array_equal(oo, os_) and np.array_equal(Ro, Rs)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_setup_and_precompute_match_oracle(sip, TF):
    for spec in (pr.spec_config1((12, 10), TF), pr.spec_config2((6, 5, 4), TF), pr.spec_config3((6, 5, 4), TF),
                 pr.spec_config4((6, 5, 4), TF)):
        ob = pr.build(orc, dict(spec))
        sb = pr.build(sip, dict(spec))
        assert len(ob["TD_OP"]) == len(sb["TD_OP"]) == len(spec["sets"]) + 1
        for Ao, As in zip(ob["TD_OP"], sb["TD_OP"]):
            assert same_csc(As.tosparse(), Ao)
        for Ro, Rs in zip(ob["AtA"], sb["AtA"]):
            assert np.array_equal(Ro, Rs)
        po, ps = ob["set_Prop"], sb["set_Prop"]
        for a, b in zip(po.AtA_offsets, ps.AtA_offsets):
            assert np.array_equal(a, b)
        assert po.ncvx == ps.ncvx and po.tag == ps.tag and po.banded == ps.banded and po.AtA_diag == ps.AtA_diag
        assert [tuple(t) for t in po.TD_n] == [tuple(t) for t in ps.TD_n]
        for a, b in zip(ob["l"] + ob["y"], sb["l"] + sb["y"]):
            assert a.shape == b.shape and a.dtype == b.dtype and not b.any()


def test_ncvx_flags(sip):
    """setup_constraints.jl:89-97."""
    cg = sip.compgrid((1.0, 1.0), (8, 8))
    cons = [sip.set_definitions("bounds", "D_z", 0.5, 1.0, ("matrix", "")),
            sip.set_definitions("bounds", "D_z", 0.0, 1.0, ("matrix", "")),
            sip.set_definitions("bounds", "identity", 0.5, 1.0, ("matrix", "")),
            sip.set_definitions("cardinality", "TV", 0, 3, ("matrix", ""))]
    _, _, sp_ = sip.setup_constraints(cons, cg, np.float32)
    assert sp_.ncvx == [True, False, False, True]
    assert isinstance(cons[0].min, np.float32) and isinstance(cons[3].max, int)     # ints untouched (:31-43)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_minkowski_precompute_matches_oracle(sip, TF):
    """PARSDMM_precompute_distribute_Minkowski.jl:3-157: block operators and 2N x 2N AtA in CDS form."""
    n, d = (9, 7), (2.0, 3.0)
    out = []
    for api in (orc, sip):
        cg = api.compgrid(d, n)
        c1 = [api.set_definitions("bounds", "identity", -1.0, 1.0, ("matrix", "")),
              api.set_definitions("bounds", "D_z", 0.0, 5.0, ("matrix", ""))]
        c2 = [api.set_definitions("bounds", "identity", -2.0, 2.0, ("matrix", "")),
              api.set_definitions("l1", "TV", 0.0, 3.0, ("matrix", ""))]
        cs = [api.set_definitions("bounds", "identity", 0.0, 9.0, ("matrix", "")),
              api.set_definitions("bounds", "D_x", -1.0, 1.0, ("matrix", ""))]
        P1, T1, S1 = api.setup_constraints(c1, cg, TF)
        P2, T2, S2 = api.setup_constraints(c2, cg, TF)
        P3, T3, S3 = api.setup_constraints(cs, cg, TF)
        opt = api.PARSDMM_options()
        opt.Minkowski = True
        out.append(api.PARSDMM_precompute_distribute_Minkowski(T1, T2, T3, S1, S2, S3, cg, opt))
    (TDo, SPo, AtAo, lo, yo), (TDs, SPs, AtAs, ls, ys) = out
    assert len(TDo) == len(TDs) == 7
    for Ao, As in zip(TDo, TDs):
        assert same_csc(As.tosparse(), Ao)
    for i, (Ro, Rs) in enumerate(zip(AtAo, AtAs)):
        assert np.array_equal(SPo.AtA_offsets[i], SPs.AtA_offsets[i]), i
        assert np.array_equal(Ro, Rs), i
    N = 63
    assert list(SPs.AtA_offsets[-1]) == [-N, 0, N]
    assert SPo.tag == SPs.tag and SPo.ncvx == SPs.ncvx and SPo.AtA_diag == SPs.AtA_diag


def test_options_defaults_and_conversion(sip):
    o = sip.PARSDMM_options()
    assert (o.maxit, o.rho_update_frequency, o.rho_ini, o.gamma_ini) == (200, 2, [10.0], 1.0)
    assert (o.evol_rel_tol, o.feas_tol, o.obj_tol) == (1e-3, 5e-2, 1e-3)
    sip.convert_options(o, np.float32)
    assert isinstance(o.feas_tol, np.float32) and isinstance(o.rho_ini[0], np.float32)
    o2 = sip.default_PARSDMM_options(sip.PARSDMM_options(), np.float64)
    assert isinstance(o2.obj_tol, np.float64) and o2.FL is np.float64


# ---- C ABI surface ---------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol(sip):
    """libsipb200.so loads and exports every function include/sipb200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "sipb200.h")).read()
    declared = set(re.findall(r"\b(sipb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(sip._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    bound = {s[0] for s in sip._lib.SYMBOLS}
    assert declared == bound
    assert sip._lib.load().sipb_abi_version() == 1
    names = [sip._lib.load().sipb_kernel_class_name(i).decode() for i in range(sip._lib.N_KERNEL_CLASSES)]
    assert "cds_spmv_dot" in names and "yl_update_fused" in names


def test_struct_sizes_match_header(sip):
    """ctypes mirrors of sipb_set_desc / sipb_options / sipb_log have the C layout (checked against gcc)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "sipb200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(sipb_set_desc), sizeof(sipb_options), sizeof(sipb_log));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    L = sip._lib
    assert sizes == [ctypes.sizeof(L.SetDesc), ctypes.sizeof(L.Options), ctypes.sizeof(L.Log)]


def test_no_cpu_fallback_without_device(sip):
    """Without a CUDA device the compute entry points fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    op = sip.get_TD_operator(sip.compgrid((1.0, 1.0), (8, 8)), "TV", np.float32)[0]
    with pytest.raises(sip._lib.SipbError):
        op @ np.ones(64, dtype=np.float32)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "setintersectionprojection.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


# ---- multilevel index work ----------------------------------------------------------------------------
def test_nearest_neighbour_tables_and_resampling(sip):
    """Interpolations.jl BSpline(Constant()) at range(1, stop=n_src, length=n_dst): nearest neighbour with
    half-way positions rounded up; product tables == oracle tables == the float formula away from ties."""
    from oracle import multilevel as om
    for ns, nd in [(400, 200), (200, 400), (199, 399), (399, 199), (200, 100), (99, 199), (201, 100), (7, 3), (3, 7), (5, 5)]:
        a = om.nn_index(ns, nd)
        assert np.array_equal(a, sip.multilevel._nearest_table(ns, nd))
        pos = 1 + np.arange(nd) * (ns - 1) / (nd - 1)
        tie = np.isclose(pos % 1.0, 0.5)
        assert np.array_equal(a[~tie], np.round(pos[~tie]).astype(int) - 1)
        assert np.array_equal(a[tie], np.floor(pos[tie] + 0.5).astype(int) - 1)
        assert a[0] == 0 and a[-1] == ns - 1
    rng = np.random.default_rng(0)
    v = rng.standard_normal(6 * 5 * 4).astype(np.float32)
    for dst in ((12, 10, 8), (3, 3, 2), (6, 5, 4)):
        assert np.array_equal(sip.resample_nn(v, (6, 5, 4), dst), om.resample(v, (6, 5, 4), dst))
    assert np.array_equal(sip.resample_nn(v, (6, 5, 4), (6, 5, 4)), v)
    for src, dst in (((7, 9), (14, 17)), ((14, 17), (7, 9)), ((5, 6, 7), (11, 3, 7)), ((9,), (4,)), ((4, 1, 5), (8, 1, 3))):
        w = rng.standard_normal(int(np.prod(src)))
        assert np.array_equal(sip.resample_nn(w, src, dst), om.resample(w, src, dst))
        w2 = rng.standard_normal(2 * int(np.prod(src)))[::2]                    # strided input
        assert np.array_equal(sip.resample_nn(w2, src, dst), om.resample(np.ascontiguousarray(w2), src, dst))


def test_multilevel_setup_matches_oracle(sip):
    """setup_multi_level_PARSDMM + constraint2coarse: grids, spacings, rescaled constraints, operators."""
    from oracle import multilevel as om
    TF = np.float32
    spec = pr.spec_config2((16, 12, 8), TF)
    out = []
    for api, setup in ((orc, lambda m, cg, cons, opt: om.setup_multi_level_PARSDMM(m, 3, 2, cg, cons, opt, orc.types)),
                       (sip, lambda m, cg, cons, opt: sip.setup_multi_level_PARSDMM(m, 3, 2, cg, cons, opt))):
        cg = api.compgrid(tuple(spec["d"]), tuple(spec["n"]))
        cons = [api.set_definitions(st, op, lo, hi, ("tensor", "")) for (st, op, lo, hi) in spec["sets"]]
        opt = api.PARSDMM_options()
        out.append(setup(spec["m"], cg, cons, opt))
    (TDo, AtAo, Po, SPo, CGo, Co), (TDs, AtAs, Ps, SPs, CGs, Cs) = out
    assert [tuple(g.n) for g in CGo] == [tuple(g.n) for g in CGs] == [(16, 12, 8), (8, 6, 4), (4, 3, 2)]
    for go, gs in zip(CGo, CGs):
        assert np.allclose(go.d, gs.d, rtol=0, atol=0)
    for co, cs in zip(Co, Cs):
        assert co.set_type == cs.set_type and np.array_equal(np.asarray(co.max), np.asarray(cs.max))
    assert Cs[1].max == np.float32(np.float32(spec["sets"][1][3]) / 8 / 8)      # l1 radius / cf^3 per level
    for lev in range(3):
        for Ao, As in zip(TDo[lev], TDs[lev]):
            assert same_csc(As.tosparse(), Ao)
        for Ro, Rs in zip(AtAo[lev], AtAs[lev]):
            assert np.array_equal(Ro, Rs)


def test_interpolate_y_l_matches_oracle(sip):
    from oracle import multilevel as om
    TF = np.float64
    rng = np.random.default_rng(1)
    for n, sets, mode in (((12, 10, 8), pr.spec_config3((12, 10, 8), TF)["sets"], "tensor"),
                          ((14, 10), pr.spec_config1((14, 10), TF)["sets"], "matrix")):
        res = []
        for api, interp, setup in ((orc, om.interpolate_y_l, lambda m, cg, cons, opt: om.setup_multi_level_PARSDMM(m, 2, 2, cg, cons, opt, orc.types)),
                                   (sip, sip.interpolate_y_l, lambda m, cg, cons, opt: sip.setup_multi_level_PARSDMM(m, 2, 2, cg, cons, opt))):
            cg = api.compgrid((1.0,) * len(n), n)
            cons = [api.set_definitions(st, op, lo, hi, (mode, "")) for (st, op, lo, hi) in sets]
            TD, AtA, P, SP, CG, C = setup(np.zeros(int(np.prod(n)), dtype=TF), cg, cons, api.PARSDMM_options())
            r = np.random.default_rng(5)
            l = [r.standard_normal(A.shape[0]) for A in TD[1]]
            y = [r.standard_normal(A.shape[0]) for A in TD[1]]
            l, y = interp(l, y, SP, CG, len(n) == 3, 0)
            assert [v.size for v in l] == [A.shape[0] for A in TD[0]]
            res.append((l, y))
        for a, b in zip(res[0][0] + res[0][1], res[1][0] + res[1][1]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("TF", [np.float32, np.float64])
def test_custom_sparse_operator_host_side(sip, TF):
    """custom_TD_OP (setup_constraints.jl:70-72): the explicit matrix replaces A; PARSDMM_precompute_distribute
    forms mat2CDS(A'A) from it like the oracle, honouring the caller's set_Prop flags."""
    n, d = (11, 9), (2.0, 3.0)
    A, *_ = orc.get_TD_operator(orc.compgrid(d, n), "TV", TF)
    w = (1.0 + 0.25 * np.cos(np.arange(A.shape[0]))).astype(TF)
    W = sp.csc_matrix(sp.diags(w).astype(TF) @ A).astype(TF)

    def run(api):
        cons = [api.set_definitions("l1", "identity", 0.0, 10.0, ("matrix", ""), (W.copy(), False)),
                api.set_definitions("bounds", "identity", 0.0, 1.0, ("matrix", ""))]
        opt = api.PARSDMM_options()
        P_sub, TD_OP, sP = api.setup_constraints(cons, api.compgrid(d, n), TF)
        sP.AtA_diag[0], sP.dense[0], sP.banded[0] = False, False, True
        TD_OP, AtA, l, y = api.PARSDMM_precompute_distribute(TD_OP, sP, api.compgrid(d, n), opt)
        return TD_OP, AtA, sP, l, y

    To, Ao, So, lo, yo = run(orc)
    Ts, As, Ss, ls, ys = run(sip)
    assert isinstance(Ts[0], sip.SparseOperator) and same_csc(Ts[0].tosparse(), To[0])
    for a, b in zip(Ao, As):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    for a, b in zip(So.AtA_offsets, Ss.AtA_offsets):
        assert np.array_equal(a, b)
    assert [v.size for v in ys] == [v.size for v in yo] == [W.shape[0], W.shape[1], W.shape[1]]
    st = Ts[0].sparse_struct()
    assert (st.rows, st.cols, st.nnz) == (W.shape[0], W.shape[1], W.nnz)
    # the caller did not flag the operator as banded: the reference would take its sparse-Q path, rejected here
    cons = [sip.set_definitions("l1", "identity", 0.0, 10.0, ("matrix", ""), (W.copy(), False))]
    P_sub, TD_OP, sP = sip.setup_constraints(cons, sip.compgrid(d, n), TF)
    sP.AtA_diag[0], sP.banded[0] = False, False
    with pytest.raises(NotImplementedError):
        sip.PARSDMM_precompute_distribute(TD_OP, sP, sip.compgrid(d, n), sip.PARSDMM_options())
    with pytest.raises(ValueError):
        sip.SparseOperator(W[:, :-1], n, d, TF)


def test_ata_class_table_equals_cds_rows():
    """TDOperator.ata_class_table (what the device keeps instead of the N x nd array of mat2CDS(A'A),
    PARSDMM_precompute_distribute.jl:44-55): same offsets, and every row of the CDS array equals — bit for bit — the
    table row of its stencil class, for every operator kind, 2-D / 3-D, and the three Minkowski block placements."""
    import sip_b200  # noqa: F401
    from sip_b200 import _lib as lib, operators as ops

    def row_class(r, n, npts):
        half = 1 if r >= npts else 0
        c = r - half * npts
        n3 = list(n) + [1] * (3 - len(n))
        i, q = c % n3[0], c // n3[0]
        j, k = q % n3[1], q // n3[1]
        ac = lambda idx, nn: 0 if idx == 0 else (2 if idx == nn - 1 else 1)      # noqa: E731
        return ((half * 3 + ac(k, n3[2])) * 3 + ac(j, n3[1])) * 3 + ac(i, n3[0])

    for TF in (np.float32, np.float64):
        for n, h in (((7, 5), (25.0, 6.0)), ((5, 4, 6), (25.0, 12.5, 6.0)), ((3, 3, 3), (1.0, 2.0, 3.0)), ((3, 9), (2.0, 3.0))):
            for kind in ["identity", "D_x", "D_z", "TV"] + (["D_y"] if len(n) == 3 else ["D_xz"]):
                for bm in (lib.BLOCK_PLAIN, lib.BLOCK_LEFT, lib.BLOCK_RIGHT, lib.BLOCK_BOTH):
                    op = ops.TDOperator(kind, n, h, TF, bm)
                    R, offs = op.ata_cds()
                    tab, o2 = op.ata_class_table()
                    assert np.array_equal(offs, o2)
                    for r in range(R.shape[0]):
                        assert R[r, :].tobytes() == tab[row_class(r, n, op.npts), :].tobytes(), (kind, n, bm, r)
    assert ops.TDOperator("TV", (2, 5), (1.0, 1.0), np.float32).ata_class_table() is None       # axis shorter than 3


def test_cds_list_is_lazy_until_indexed():
    """PARSDMM_precompute_distribute returns AtA as a list whose stencil entries are formed on first access."""
    import sip_b200 as sip
    spec = pr.spec_config2((6, 5, 4), np.float32)
    b = pr.build(sip, spec, sip.PARSDMM_options())
    A = b["AtA"]
    assert len(A) == 5 and all(A.is_lazy(i) for i in range(5)) and not A.materialized()
    R1 = A[1]
    assert R1.shape == (6 * 5 * 4, 7) and not A.is_lazy(1) and A.materialized()
    want, offs = b["TD_OP"][1].ata_cds()
    assert np.array_equal(R1, want) and np.array_equal(b["set_Prop"].AtA_offsets[1], offs)
    assert [M.shape[1] for M in A] == [1, 7, 3, 3, 1]
