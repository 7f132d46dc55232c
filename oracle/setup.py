"""Oracle restatement of the problem set-up ("front end") of the hot path (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/setup_constraints.jl:17-102, get_projector.jl:3-103,
PARSDMM_precompute_distribute.jl:6-77 and PARSDMM_precompute_distribute_Minkowski.jl:3-157.
Only the sets/operators of the CDS hot path are restated (bounds, l1, l2, annulus, cardinality,
prox_l1 on identity/D_x/D_y/D_z/TV/D_xz in matrix/tensor mode).
"""
from __future__ import annotations

import copy

import numpy as np
import scipy.sparse as sp

from . import operators as ops
from . import projectors as proj
from .sip_types import set_properties

SPECIAL_OPERATORS = ["DFT", "DCT", "wavelet", "curvelet"]   # setup_constraints.jl:54


def get_projector(constraint, comp_grid, A, TD_n, TF):
    """get_projector.jl:3-103 restricted to matrix/tensor application mode on sparse operators."""
    st = constraint.set_type
    if constraint.TD_OP in SPECIAL_OPERATORS:
        raise NotImplementedError("oracle: JOLI transform operators are outside the CDS hot path")
    cmin, cmax = constraint.min, constraint.max
    if constraint.app_mode[0] not in ("matrix", "tensor"):
        mode = tuple(constraint.app_mode)
        if st == "bounds":                                                   # get_projector.jl:16
            return lambda x: proj.project_bounds_fiber(x, cmin, cmax, TD_n, mode)
        if st == "cardinality" and mode[0] == "slice":                       # get_projector.jl:96, slice modes
            return lambda x: proj.project_cardinality_slice(x, int(cmax), TD_n, mode)
        if st == "cardinality":                                              # get_projector.jl:96
            return lambda x: proj.project_cardinality_fiber(x, int(cmax), TD_n, mode)
        raise NotImplementedError("oracle: fiber/slice modes exist for bounds and cardinality only")
    if st == "bounds":
        return lambda x: proj.project_bounds(x, cmin, cmax)                 # :10
    if st == "prox_l1":
        return lambda x: proj.prox_l1(x, cmax)                              # :25
    if st == "l1":
        return lambda x: proj.project_l1_Duchi(x, cmax)                     # :33
    if st == "l2":
        return lambda x: proj.project_l2(x, cmax)                           # :41
    if st == "annulus":
        return lambda x: proj.project_annulus(x, cmin, cmax)                # :49
    if st == "cardinality":
        return lambda x: proj.project_cardinality(x, int(cmax))             # :90
    if st == "histogram":
        return lambda x: proj.project_histogram_relaxed(x, cmin, cmax)      # :81
    raise NotImplementedError("oracle: set type %r is outside the hot path" % st)


def setup_constraints(constraint, comp_grid, TF):
    """setup_constraints.jl:17-102.  Returns (P_sub, TD_OP, set_Prop).  Mutates `constraint`
    (min/max converted to TF) exactly like the reference (:31-43)."""
    nr = len(constraint)
    for c in constraint:                                                     # :31-43
        if np.ndim(c.min) == 0:
            if isinstance(c.min, (int, np.integer)) and not isinstance(c.min, bool):
                pass
            else:
                c.min = TF(c.min)
                c.max = TF(c.max)
        else:
            c.min = np.asarray(c.min, dtype=TF)
            c.max = np.asarray(c.max, dtype=TF)

    P_sub = [None] * nr
    TD_OP = [None] * nr
    set_Prop = set_properties([False] * nr, [False] * nr, [False] * nr, [None] * nr, [None] * nr,
                              [False] * nr, [None] * nr)
    for i, c in enumerate(constraint):
        if c.set_type in ("nuclear", "rank") and c.app_mode[0] in ("matrix", "tensor") and len(comp_grid.n) == 3:
            raise ValueError("requested rank or nuclear norm constraints on a tensor, use mode=(slice,x) e.t.c. "
                             "to define constraints per slice")                                        # :60-62
        if c.set_type in ("l1", "l2") and c.app_mode[0] in ("slice", "fiber"):
            raise ValueError("l1 and l2 constraints only available for matrix or tensor mode, currently")  # :65-67
        A, AtA_diag, dense, TD_n, banded = ops.get_TD_operator(comp_grid, c.TD_OP, TF)                 # :69
        custom = c.custom_TD_OP[0]
        if c.set_type != "subspace" and not (isinstance(custom, (list, tuple)) and len(custom) == 0):  # :70-72
            A = sp.csc_matrix(custom).astype(TF)
        P_sub[i] = get_projector(c, comp_grid, A, TD_n, TF)                                            # :74
        TD_OP[i] = A                                                                                   # :79
        set_Prop.AtA_diag[i] = AtA_diag
        set_Prop.dense[i] = dense
        set_Prop.TD_n[i] = TD_n
        set_Prop.banded[i] = banded
        set_Prop.tag[i] = (c.set_type, c.TD_OP, c.app_mode[0], c.app_mode[1])                          # :86
        if c.set_type in ("rank", "cardinality"):                                                      # :89-97
            set_Prop.ncvx[i] = True
        elif c.set_type in ("bounds", "histogram") and c.TD_OP != "identity" and TF(np.max(c.min)) > TF(0.0):
            set_Prop.ncvx[i] = True
        else:
            set_Prop.ncvx[i] = False
    return P_sub, TD_OP, set_Prop


def PARSDMM_precompute_distribute(TD_OP, set_Prop, comp_grid, options):
    """PARSDMM_precompute_distribute.jl:6-77.  MUTATES TD_OP and set_Prop (push of the distance term)."""
    TF = TD_OP[0].dtype.type
    N = int(np.prod(comp_grid.n))
    if not options.feasibility_only:                                         # :17-26
        TD_OP.append(sp.identity(N, dtype=TF, format="csc"))
        set_Prop.TD_n.append(tuple(comp_grid.n))
        set_Prop.AtA_offsets.append(np.array([0], dtype=np.int64))
        set_Prop.banded.append(True)
        set_Prop.AtA_diag.append(True)
        set_Prop.ncvx.append(False)
        set_Prop.dense.append(False)
        set_Prop.tag.append(("distance squared", "identity", "matrix", ""))
    p = len(TD_OP)
    AtA = [None] * p
    for i in range(p):                                                       # :40-49
        if set_Prop.AtA_diag[i]:
            AtA[i] = sp.identity(N, dtype=TF, format="csc")
        else:
            AtA[i] = ops.AtA_sparse(TD_OP[i])
    if all(set_Prop.banded[:p]):                                             # :52-59
        for i in range(p):
            AtA[i], off = ops.mat2CDS(AtA[i])
            set_Prop.AtA_offsets[i] = off.astype(np.int64)
        set_Prop.AtA_offsets = set_Prop.AtA_offsets[:p]
    y = [np.zeros(TD_OP[i].shape[0], dtype=TF) for i in range(p)]            # :62-68
    l = [np.zeros(TD_OP[i].shape[0], dtype=TF) for i in range(p)]
    return TD_OP, AtA, l, y


def PARSDMM_precompute_distribute_Minkowski(TD_OP_c1, TD_OP_c2, TD_OP_sum, set_Prop_c1, set_Prop_c2,
                                            set_Prop_sum, comp_grid, options):
    """PARSDMM_precompute_distribute_Minkowski.jl:3-157 (sparse operators only).
    Returns (TD_OP, set_Prop, AtA, l, y).  Mutates the three operator lists and set_Prop_sum."""
    TF = TD_OP_c1[0].dtype.type if len(TD_OP_c1) else TD_OP_c2[0].dtype.type
    N = int(np.prod(comp_grid.n))
    p, q, r = len(TD_OP_c1), len(TD_OP_c2), len(TD_OP_sum)
    s = p + q + r if options.feasibility_only else p + q + r + 1            # :19-23
    AtA = [None] * s
    Z = sp.csc_matrix((N, N), dtype=TF)
    Id = sp.identity(N, dtype=TF, format="csc")

    def blk(a, b, c, d):
        M = sp.bmat([[a, b], [c, d]], format="csc", dtype=TF)
        M.sort_indices()
        return M

    for i in range(p):                                                       # :32-46
        if set_Prop_c1.dense[i]:
            if set_Prop_c1.AtA_diag[i]:
                AtA[i] = blk(Id, Z, Z, Z)
            else:
                raise ValueError("provided a dense non orthogoal transform-domain operator")
        else:
            AtA[i] = blk(ops.AtA_sparse(TD_OP_c1[i]), Z, Z, Z)
    for i in range(q):                                                       # :47-61
        if set_Prop_c2.dense[i]:
            if set_Prop_c2.AtA_diag[i]:
                AtA[p + i] = blk(Z, Z, Z, Id)
            else:
                raise ValueError("provided a dense non orthogoal transform-domain operator")
        else:
            AtA[p + i] = blk(Z, Z, Z, ops.AtA_sparse(TD_OP_c2[i]))
    for i in range(r):                                                       # :62-74
        if set_Prop_sum.dense[i]:
            if set_Prop_sum.AtA_diag[i]:
                AtA[i + p + q] = blk(Id, Id, Id, Id)
            else:
                raise ValueError("provided a dense non orthogoal transform-domain operator")
        else:
            B = ops.AtA_sparse(TD_OP_sum[i])
            AtA[i + p + q] = blk(B, B, B, B)

    def hcat(a, b):
        M = sp.hstack([a, b], format="csc", dtype=TF)
        M.sort_indices()
        return M

    for i in range(p):                                                       # :78-81
        TD_OP_c1[i] = hcat(TD_OP_c1[i], sp.csc_matrix((TD_OP_c1[i].shape[0], N), dtype=TF))
    for i in range(q):                                                       # :82-85
        TD_OP_c2[i] = hcat(sp.csc_matrix((TD_OP_c2[i].shape[0], N), dtype=TF), TD_OP_c2[i])
    for i in range(r):                                                       # :86-88
        TD_OP_sum[i] = hcat(TD_OP_sum[i], TD_OP_sum[i])

    if not options.feasibility_only:                                         # :91-101
        TD_OP_sum.append(hcat(Id, Id))
        set_Prop_sum.TD_n.append(tuple(comp_grid.n))
        set_Prop_sum.AtA_offsets.append(np.array([0], dtype=np.int64))
        set_Prop_sum.banded.append(True)
        set_Prop_sum.AtA_diag.append(False)
        set_Prop_sum.dense.append(False)
        set_Prop_sum.ncvx.append(False)
        set_Prop_sum.tag.append(("distance squared", "identity", "matrix", ""))
        AtA[s - 1] = blk(Id, Id, Id, Id)

    set_Prop = copy.deepcopy(set_Prop_c1)                                    # :104-120
    for other in (set_Prop_c2, set_Prop_sum):
        set_Prop.AtA_diag += list(other.AtA_diag)
        set_Prop.AtA_offsets += list(other.AtA_offsets)
        set_Prop.TD_n += list(other.TD_n)
        set_Prop.banded += list(other.banded)
        set_Prop.dense += list(other.dense)
        set_Prop.ncvx += list(other.ncvx)
        set_Prop.tag += list(other.tag)

    if all(set_Prop.banded[:s]):                                             # :123-131
        for i in range(s):
            AtA[i], off = ops.mat2CDS(AtA[i])
            set_Prop.AtA_offsets[i] = off.astype(np.int64)
        set_Prop.AtA_offsets = set_Prop.AtA_offsets[:s]

    TD_OP = list(TD_OP_c1) + list(TD_OP_c2) + list(TD_OP_sum)                # :138-141
    y = [np.zeros(TD_OP[i].shape[0], dtype=TF) for i in range(s)]
    l = [np.zeros(TD_OP[i].shape[0], dtype=TF) for i in range(s)]
    return TD_OP, set_Prop, AtA, l, y
