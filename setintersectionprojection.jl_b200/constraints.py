"""setup_constraints / get_projector — host side (setup_constraints.jl:17-102, get_projector.jl:3-103).

The reference returns closures that capture min/max (get_projector.jl:10,33,41,90); `set_Prop` does not
carry the bounds.  Here `P_sub[i]` is a callable `Projector` functor: `P_sub[i](v)` still projects a
vector in place (on the GPU, through the C ABI — there is no CPU fallback) and additionally exposes
{kind, min, max, k} so that `PARSDMM` can hand a descriptor to the device solver.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .operators import SPECIAL_OPERATORS, SparseOperator, get_TD_operator
from .types import set_properties

_REJECTED = {
    "rank": "SVD-based rank constraints are out of scope on the GPU path",
    "nuclear": "SVD-based nuclear-norm constraints are out of scope on the GPU path",
    "subspace": "subspace constraints are outside the device hot path",
}


class Projector:
    """Device projector functor: P(v) mutates v in place and returns it, like the Julia `!` functions."""

    def __init__(self, set_kind: int, TF, lo=0.0, hi=0.0, k=0, lo_vec=None, hi_vec=None, name="", fiber_axis=0,
                 td_n=None):
        self.set_kind = set_kind
        self.fiber_axis = int(fiber_axis)
        self.td_n = None if td_n is None else tuple(int(v) for v in td_n) + (1,) * (3 - len(td_n))
        self.TF = np.dtype(TF).type
        self.min = lo
        self.max = hi
        self.k = int(k)
        # private copies: the device problem built from this projector is cached and keyed on these arrays, so a caller
        # who later changes the bound vectors in place must not silently keep the stale device copy
        self.min_vec = None if lo_vec is None else np.array(lo_vec, dtype=self.TF, order="C", copy=True)
        self.max_vec = None if hi_vec is None else np.array(hi_vec, dtype=self.TF, order="C", copy=True)
        for v in (self.min_vec, self.max_vec):
            if v is not None:
                v.setflags(write=False)
        self.name = name

    def descriptor(self, op_kind=_lib.OP_IDENTITY, block_mode=_lib.BLOCK_PLAIN, ncvx=False) -> _lib.SetDesc:
        d = _lib.SetDesc()
        d.set_kind, d.op_kind, d.block_mode, d.ncvx = self.set_kind, op_kind, block_mode, int(bool(ncvx))
        d.min = float(self.min) if np.ndim(self.min) == 0 else 0.0
        d.max = float(self.max) if np.ndim(self.max) == 0 else 0.0
        d.k = self.k
        d.min_vec = self.min_vec.ctypes.data if self.min_vec is not None else None
        d.max_vec = self.max_vec.ctypes.data if self.max_vec is not None else None
        d.fiber_axis = self.fiber_axis
        if self.td_n is not None:
            d.td_n[:] = self.td_n
        return d

    def __call__(self, v: np.ndarray) -> np.ndarray:
        if not isinstance(v, np.ndarray) or v.dtype != self.TF or not v.flags.c_contiguous:
            raise TypeError("projector expects a contiguous %s vector" % self.TF.__name__)
        if self.td_n is not None:
            if int(np.prod(self.td_n)) != v.size:
                raise ValueError("input has %d entries, the transform-domain grid %s" % (v.size, self.td_n))
        elif self.min_vec is not None and self.min_vec.size != v.size:
            raise ValueError("vector bounds have %d entries, input has %d" % (self.min_vec.size, v.size))
        d = self.descriptor()
        lib = _lib.load()
        if self.set_kind == _lib.SET_CARD_SLICE and self.fiber_axis != 2:
            # the reference's x / y slice modes work on a permuted COPY and return it; the argument is left untouched
            # (project_cardinality!.jl:115-118 "code currently does not mutate the input for slice projections")
            v = v.copy()
        _lib.check(lib.sipb_project(_lib.ctx(), _lib.dtype_code(self.TF), C.byref(d), v.size, v.ctypes.data, None))
        return v

    def __repr__(self):
        return "Projector(%s)" % self.name


def get_projector(constraint, comp_grid, special_operator_list, A, TD_n, TF) -> Projector:
    """get_projector.jl:3-103 for the sets of the device hot path (matrix/tensor application mode)."""
    st = constraint.set_type
    if st in _REJECTED:
        raise NotImplementedError(_REJECTED[st] + " and are rejected (no CPU fallback)")
    if constraint.TD_OP in special_operator_list:
        raise NotImplementedError("JOLI transform operators (%s) are outside the device CDS path" % constraint.TD_OP)
    lo, hi = constraint.min, constraint.max
    if constraint.app_mode[0] not in ("matrix", "tensor"):
        # fiber application modes (get_projector.jl:12-18,92-98; project_bounds!.jl:38-88,
        # project_cardinality!.jl:23-113) and the slice mode of cardinality on 3-D tensors (:115-146)
        if constraint.app_mode[0] == "slice" and st == "cardinality":
            if len(TD_n) != 3:
                raise NotImplementedError("slice modes exist for 3-D tensors only")
            if constraint.TD_OP in ("TV", "D2D", "D3D"):
                raise ValueError("slice modes need a single-block operator (the TV output is not a grid)")
            try:
                axis = {"x": 0, "y": 1, "z": 2}[constraint.app_mode[1]]
            except KeyError:
                raise ValueError("slice direction %r is not valid" % (constraint.app_mode[1],))
            return Projector(_lib.SET_CARD_SLICE, TF, k=int(hi), name="cardinality(slice)", fiber_axis=axis, td_n=TD_n)
        if constraint.app_mode[0] == "slice" and st == "bounds":
            raise NotImplementedError("bound constraints per slice of a tensor currently not implemented, yet...")   # project_bounds!.jl:83
        if constraint.app_mode[0] != "fiber" or st not in ("bounds", "cardinality"):
            raise NotImplementedError("only the fiber modes of bounds / cardinality and the slice mode of cardinality are "
                                      "on the device path")
        if constraint.TD_OP in ("TV", "D2D", "D3D"):
            raise ValueError("fiber modes need a single-block operator (the TV output is not a grid)")
        nd = len(TD_n)
        try:
            axis = ({"x": 0, "z": 1} if nd == 2 else {"x": 0, "y": 1, "z": 2})[constraint.app_mode[1]]
        except KeyError:
            raise ValueError("fiber direction %r is not valid for a %d-D grid" % (constraint.app_mode[1], nd))
        if st == "bounds":
            if np.ndim(lo) == 0 or np.size(lo) != TD_n[axis] or np.size(hi) != TD_n[axis]:
                raise ValueError("fiber bounds need one (min, max) pair per point of the fiber (%d)" % TD_n[axis])
            return Projector(_lib.SET_BOUNDS_FIBER, TF, lo_vec=lo, hi_vec=hi, name="bounds(fiber)", fiber_axis=axis, td_n=TD_n)
        return Projector(_lib.SET_CARD_FIBER, TF, k=int(hi), name="cardinality(fiber)", fiber_axis=axis, td_n=TD_n)
    if st == "bounds":
        if np.ndim(lo) == 0:
            return Projector(_lib.SET_BOUNDS_SCALAR, TF, lo, hi, name="bounds")          # :10
        return Projector(_lib.SET_BOUNDS_VECTOR, TF, lo_vec=lo, hi_vec=hi, name="bounds(vector)")
    if st == "prox_l1":
        return Projector(_lib.SET_PROX_L1, TF, 0.0, hi, name="prox_l1")                  # :25
    if st == "l1":
        return Projector(_lib.SET_L1, TF, 0.0, hi, name="l1")                            # :33
    if st == "l2":
        return Projector(_lib.SET_L2, TF, 0.0, hi, name="l2")                            # :41
    if st == "annulus":
        return Projector(_lib.SET_ANNULUS, TF, lo, hi, name="annulus")                   # :49
    if st == "cardinality":
        return Projector(_lib.SET_CARDINALITY, TF, k=int(hi), name="cardinality")        # :90 convert(Integer, max)
    if st == "histogram":                                                                # :77-82, vector-valued min / max
        if np.ndim(lo) == 0 or np.ndim(hi) == 0 or np.size(lo) != np.size(hi):
            raise ValueError("histogram constraints need vector-valued min and max of equal length (sorted bounds)")
        return Projector(_lib.SET_HISTOGRAM, TF, lo_vec=lo, hi_vec=hi, name="histogram")
    raise ValueError("unknown set type %r" % st)


def setup_constraints(constraint, comp_grid, TF):
    """setup_constraints.jl:17-102 -> (P_sub, TD_OP, set_Prop).  Mutates `constraint` (min/max cast to TF,
    :31-43) like the reference."""
    TF = np.dtype(TF).type
    for c in constraint:
        if np.ndim(c.min) == 0:
            if not (isinstance(c.min, (int, np.integer)) and not isinstance(c.min, bool)):
                c.min, c.max = TF(c.min), TF(c.max)
        else:
            c.min, c.max = np.asarray(c.min, dtype=TF), np.asarray(c.max, dtype=TF)
    P_sub, TD_OP = [], []
    sp_ = set_properties()
    special = list(SPECIAL_OPERATORS)
    for c in constraint:
        if c.set_type in ("nuclear", "rank") and c.app_mode[0] in ("matrix", "tensor") and len(comp_grid.n) == 3:
            raise ValueError("requested rank or nuclear norm constraints on a tensor, use mode=(slice,x) e.t.c. to "
                             "define constraints per slice")                              # :60-62
        if c.set_type in ("l1", "l2") and c.app_mode[0] in ("slice", "fiber"):
            raise ValueError("l1 and l2 constraints only available for matrix or tensor mode, currently")   # :65-67
        A, AtA_diag, dense, TD_n, banded = get_TD_operator(comp_grid, c.TD_OP, TF)        # :69
        custom = c.custom_TD_OP[0]
        if c.set_type != "subspace" and not (isinstance(custom, (list, tuple)) and len(custom) == 0):   # :70-72
            if c.app_mode[0] not in ("matrix", "tensor"):
                raise NotImplementedError("custom operators are on the device path in matrix/tensor mode only")
            if not (hasattr(custom, "tocsc") or isinstance(custom, np.ndarray)):
                raise NotImplementedError("custom_TD_OP must be an explicit (SciPy sparse / dense) matrix; JOLI-style "
                                          "operators are rejected on the device path")
            A = SparseOperator(custom, comp_grid.n, comp_grid.d, TF)
        P_sub.append(get_projector(c, comp_grid, special, A, TD_n, TF))                   # :74
        TD_OP.append(A)
        sp_.AtA_diag.append(AtA_diag)
        sp_.dense.append(dense)
        sp_.TD_n.append(TD_n)
        sp_.banded.append(banded)
        sp_.tag.append((c.set_type, c.TD_OP, c.app_mode[0], c.app_mode[1]))               # :86
        sp_.AtA_offsets.append(None)
        if c.set_type in ("rank", "cardinality"):                                         # :89-97
            sp_.ncvx.append(True)
        elif c.set_type in ("bounds", "histogram") and c.TD_OP != "identity" and TF(np.max(c.min)) > TF(0.0):
            sp_.ncvx.append(True)
        else:
            sp_.ncvx.append(False)
    return P_sub, TD_OP, sp_
