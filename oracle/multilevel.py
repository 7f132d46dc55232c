"""Oracle restatement of the multilevel driver (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/PARSDMM_multi_level.jl:8-89, setup_multi_level_PARSDMM.jl:7-137,
constraint2coarse.jl:8-104 and interpolate_y_l.jl:7-97.

Third-party arithmetic: Interpolations.jl (compat "0.13", not vendored) — `interpolate(A, BSpline(Constant()))`
evaluated at `range(1, stop=n_src, length=n_dst)` is nearest-neighbour sampling; v0.13 rounds half-way
positions up (`floor(x + 1/2)`).  PARITY UNPINNED: the reference's only multilevel test is disabled
(test/runtests.jl:47) and no Julia is available; for the BASELINE grids (400->200->100 and the derivative
axes 399->199->99) no sample lands on a half-way position, so the rounding rule is moot there.
"""
from __future__ import annotations

import copy

import numpy as np

from . import parsdmm as par
from . import setup as stp


def nn_index(n_src: int, n_dst: int) -> np.ndarray:
    """0-based source indices of itp(range(1, stop=n_src, length=n_dst)) with round-half-up.
    position_k = 1 + k (n_src-1)/(n_dst-1); index = floor(position + 1/2), in exact integer arithmetic."""
    if n_dst == 1:
        return np.zeros(1, dtype=np.int64)
    k = np.arange(n_dst, dtype=np.int64)
    den = 2 * (n_dst - 1)
    num = 2 * (n_dst - 1) + 2 * k * (n_src - 1) + (n_dst - 1)
    return num // den - 1


def resample(v: np.ndarray, n_src, n_dst) -> np.ndarray:
    """vec(itp(range...)) of a column-major array of shape n_src sampled to shape n_dst."""
    A = v.reshape(tuple(n_src), order="F")
    idx = [nn_index(int(s), int(d)) for s, d in zip(n_src, n_dst)]
    return np.ascontiguousarray(A[np.ix_(*idx)].ravel(order="F"))


def constraint2coarse(constraint, comp_grid, coarsening_factor):
    """constraint2coarse.jl:8-104 (mutates and returns `constraint`)."""
    n = comp_grid.n
    for c in constraint:
        if c.set_type == "rank":
            c.max = min(c.max, min(n))
        if c.set_type == "cardinality":
            c.max = min(c.max, int(np.prod(n)))
    dim3 = len(n) == 3 and n[2] > 1
    for c in constraint:
        if c.set_type == "l1":
            c.max = c.max / (coarsening_factor ** 3 if dim3 else coarsening_factor ** 2)
        if c.set_type == "l2":
            c.max = c.max / (np.sqrt(coarsening_factor ** 3) if dim3 else coarsening_factor)
        if c.set_type == "nuclear" and not dim3:
            c.max = c.max / 2.7
    return constraint


def setup_multi_level_PARSDMM(m, n_levels, coarsening_factor, comp_grid, constraint, options, types):
    """setup_multi_level_PARSDMM.jl:7-137 -> (TD_OP_levels, AtA_levels, P_sub_levels, set_Prop_levels,
    comp_grid_levels, constraint_level)."""
    TF = m.dtype.type
    TD_OP_levels, AtA_levels, P_sub_levels, set_Prop_levels, comp_grid_levels = [], [], [], [], []
    P_sub, TD_OP, set_Prop = stp.setup_constraints(constraint, comp_grid, TF)                    # :45
    TD_OP, AtA, l, y = stp.PARSDMM_precompute_distribute(TD_OP, set_Prop, comp_grid, options)     # :46
    TD_OP_levels.append(TD_OP); AtA_levels.append(AtA); P_sub_levels.append(P_sub)
    set_Prop_levels.append(set_Prop); comp_grid_levels.append(comp_grid)
    constraint_level = copy.deepcopy(constraint)                                                  # :61
    for i in range(2, n_levels + 1):
        # round.(Int, n ./ cf^(i-1)): Julia rounds half to even
        n = tuple(int(np.round(v / coarsening_factor ** (i - 1))) for v in comp_grid.n)           # :66
        d = tuple((vn / nn) * vd for vn, nn, vd in zip(comp_grid.n, n, comp_grid.d))              # :82
        cg = types.compgrid(d, n)
        comp_grid_levels.append(cg)
        constraint_level = constraint2coarse(constraint_level, cg, coarsening_factor)             # :87
        P_l, TD_l, SP_l = stp.setup_constraints(constraint_level, cg, TF)                         # :91
        TD_l, AtA_l, _, _ = stp.PARSDMM_precompute_distribute(TD_l, SP_l, cg, options)            # :92
        TD_OP_levels.append(TD_l); AtA_levels.append(AtA_l); P_sub_levels.append(P_l); set_Prop_levels.append(SP_l)
    return TD_OP_levels, AtA_levels, P_sub_levels, set_Prop_levels, comp_grid_levels, constraint_level


def interpolate_y_l(l, y, set_Prop_levels, comp_grid_levels, dim3, i):
    """interpolate_y_l.jl:7-97; `i` is the 0-based index of the FINER level (levels[i+1] is the coarser).

    For TV operators the vectors are split into blocks AS IF ordered (D_x, D_y, D_z) with shapes
    (n1-1,n2,n3), (n1,n2-1,n3), (n1,n2,n3-1) although the operator is ordered (D_z, D_y, D_x) — a heuristic
    warm start of the reference, replicated verbatim."""
    nc = comp_grid_levels[i + 1].n
    nf = comp_grid_levels[i].n
    for j in range(len(l)):
        tag = set_Prop_levels[i].tag[j][1]
        if tag in ("TV", "D2D", "D3D"):
            if dim3:
                shapes_c = [(nc[0] - 1, nc[1], nc[2]), (nc[0], nc[1] - 1, nc[2]), (nc[0], nc[1], nc[2] - 1)]
                shapes_f = [(nf[0] - 1, nf[1], nf[2]), (nf[0], nf[1] - 1, nf[2]), (nf[0], nf[1], nf[2] - 1)]
            else:
                shapes_c = [(nc[0] - 1, nc[1]), (nc[0], nc[1] - 1)]
                shapes_f = [(nf[0] - 1, nf[1]), (nf[0], nf[1] - 1)]
            ends = np.cumsum([int(np.prod(s)) for s in shapes_c])
            starts = np.concatenate(([0], ends[:-1]))
            ends[-1] = l[j].size                       # y[j][p2e+1:end]
            l[j] = np.concatenate([resample(l[j][a:b], sc, sf) for a, b, sc, sf in zip(starts, ends, shapes_c, shapes_f)])
            y[j] = np.concatenate([resample(y[j][a:b], sc, sf) for a, b, sc, sf in zip(starts, ends, shapes_c, shapes_f)])
        else:
            s = tuple(a - b for a, b in zip(nf, set_Prop_levels[i].TD_n[j]))                      # :78
            src = tuple(set_Prop_levels[i + 1].TD_n[j])
            dst = tuple(a - b for a, b in zip(nf, s))
            l[j] = resample(l[j], src, dst)
            y[j] = resample(y[j], src, dst)
    return l, y


def PARSDMM_multi_level(m, TD_OP_levels, AtA_levels, P_sub_levels, set_Prop_levels, comp_grid_levels, options,
                        x_ini=None, l_ini=None, y_ini=None, solver=None):
    """PARSDMM_multi_level.jl:8-89.  `solver` defaults to the oracle PARSDMM."""
    solve = solver or par.PARSDMM
    TF = m.dtype.type
    n_levels = len(TD_OP_levels)
    rho_orig = copy.deepcopy(options.rho_ini)                                                     # :30
    n0 = comp_grid_levels[0].n
    dim3 = len(n0) == 3 and n0[2] > 1
    m_levels = [m] + [resample(m, n0, comp_grid_levels[i].n) for i in range(1, n_levels)]         # :40-48
    i = n_levels - 1
    options.zero_ini_guess = True                                                                 # :53
    x_ini = np.zeros(int(np.prod(comp_grid_levels[-1].n)), dtype=TF) if x_ini is None else x_ini
    x, log, l, y = solve(m_levels[i], AtA_levels[i], TD_OP_levels[i], set_Prop_levels[i], P_sub_levels[i],
                         comp_grid_levels[i], options, x_ini, l_ini, y_ini)                       # :56
    options.rho_ini = [TF(v) for v in log.rho[-1, :]]                                             # :57
    logs = [log]
    for i in range(n_levels - 2, -1, -1):
        x = resample(x, comp_grid_levels[i + 1].n, comp_grid_levels[i].n)                         # :61-67
        l, y = interpolate_y_l(l, y, set_Prop_levels, comp_grid_levels, dim3, i)                  # :74
        options.zero_ini_guess = False                                                            # :81
        x, log, l, y = solve(m_levels[i], AtA_levels[i], TD_OP_levels[i], set_Prop_levels[i], P_sub_levels[i],
                             comp_grid_levels[i], options, x, l, y)                               # :82
        options.rho_ini = [TF(v) for v in log.rho[-1, :]]                                         # :83
        logs.append(log)
    options.rho_ini = rho_orig                                                                    # :87
    log.levels = logs
    return x, log, l, y
