"""PARSDMM(m,AtA,TD_OP,set_Prop,P_sub,comp_grid,options[,x,l,y]) — host wrapper over the C ABI.

Drop-in for PARSDMM.jl:25-258: same arguments, same return tuple `(x, log_PARSDMM, l, y)`, same log
fields (trimmed exactly like output_check_PARSDMM, PARSDMM.jl:261-278) and the seven timing sections.
The whole iteration runs on the GPU inside `sipb_solve`; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from . import distributed as dd
from .constraints import Projector
from .operators import TDOperator
from .types import convert_options, log_type_PARSDMM


class _DeviceProblem:
    def __init__(self, handle, key, p, pp, N, rows, q_offsets):
        self.handle, self.key, self.p, self.pp, self.N, self.rows, self.q_offsets = handle, key, p, pp, N, rows, q_offsets
        self._fin = weakref.finalize(self, _destroy, handle)


def _destroy(handle):
    try:
        _lib.load().sipb_problem_destroy(handle)
    except Exception:
        pass


def _problem_key(TF, TD_OP, P_sub, set_Prop, options):
    items = [np.dtype(TF).str, bool(options.feasibility_only), bool(options.Minkowski)]
    for A in TD_OP:
        items.append((A.kind, A.n, tuple(float(v) for v in A.h), A.block_mode, id(A) if A.kind == "custom" else None))
    for P in P_sub:
        items.append((P.set_kind, float(P.min) if np.ndim(P.min) == 0 else None,
                      float(P.max) if np.ndim(P.max) == 0 else None, P.k,
                      None if P.min_vec is None else P.min_vec.ctypes.data, P.fiber_axis, P.td_n))
    items.append(tuple(bool(v) for v in set_Prop.ncvx))
    items.append((dd.rank(), dd.world()) if dd.active() else None)
    return tuple(items)


def build_device_problem(TF, AtA, TD_OP, set_Prop, P_sub, comp_grid, options, ctx=None) -> _DeviceProblem:
    """Upload operators descriptors + AtA (CDS) once; replaces the data movement of
    PARSDMM_precompute_distribute.jl and the allocation part of PARSDMM_initialize.jl."""
    lib = _lib.load()
    p = len(TD_OP)
    pp = p if options.feasibility_only else p - 1
    if len(P_sub) != pp:
        raise ValueError("P_sub must hold one projector per constraint set (%d), got %d" % (pp, len(P_sub)))
    if len(AtA) != p:
        raise ValueError("AtA must hold one CDS matrix per operator")
    for A in TD_OP:
        if not isinstance(A, TDOperator):
            raise NotImplementedError("operator %r is not a banded TDOperator; rejected on the device path" % (A,))
    for P in P_sub:
        if not isinstance(P, Projector):
            raise NotImplementedError("P_sub entries must be the Projector functors returned by setup_constraints")
    op0 = TD_OP[0]
    slab = None
    if dd.active():
        if op0.ndim != 3 or options.Minkowski:
            raise NotImplementedError("multi-GPU slabs need a 3-D, non-Minkowski problem")
        slab = dd.slab_range(op0.n[2])
        if getattr(AtA, "slab", None) not in (None, slab):
            raise ValueError("AtA was built for another slab than this rank's")
    n = (C.c_int64 * 3)(*(list(op0.n) + [1] * (3 - op0.ndim)))
    h = (C.c_double * 3)(*([float(v) for v in op0.h] + [1.0] * (3 - op0.ndim)))
    handle = C.c_void_p()
    _lib.check(lib.sipb_problem_create(ctx if ctx is not None else _lib.ctx(), _lib.dtype_code(TF), op0.ndim, n, h,
                                       int(bool(options.Minkowski)),
                                       int(bool(options.feasibility_only)), C.byref(handle)))
    # one row per stencil class instead of the N x nd arrays, unless the caller formed (and maybe changed) an array
    use_tables = (hasattr(AtA, "is_lazy") and all(AtA.is_lazy(i) for i in range(p)) and not AtA.materialized()
                  and os.environ.get("SIPB_Q_CLASSES", "1") != "0")
    try:
        for i in range(p):
            keep = None
            if i < pp:
                d = P_sub[i].descriptor(TD_OP[i].op_kind, TD_OP[i].block_mode, set_Prop.ncvx[i])
                if slab is not None and P_sub[i].min_vec is not None and P_sub[i].set_kind == _lib.SET_BOUNDS_VECTOR:
                    # vector bounds: this rank's rows only (fiber bounds are indexed by the fiber coordinate: whole vector)
                    keep = (dd.scatter_td(P_sub[i].min_vec, TD_OP[i], *slab), dd.scatter_td(P_sub[i].max_vec, TD_OP[i], *slab))
                    d.min_vec, d.max_vec = keep[0].ctypes.data, keep[1].ctypes.data
            else:
                d = _lib.SetDesc()
                d.set_kind, d.op_kind, d.block_mode, d.ncvx = _lib.SET_DISTANCE, TD_OP[i].op_kind, TD_OP[i].block_mode, 0
            if TD_OP[i].op_kind == _lib.OP_SPARSE:          # custom operator: CSR + CSC arrays travel with the descriptor
                if slab is not None:
                    raise NotImplementedError("custom operators are single-GPU (no slab decomposition)")
                d.sparse = C.pointer(TD_OP[i].sparse_struct())
            _lib.check(lib.sipb_problem_add_set(handle, C.byref(d)))
            if use_tables:
                tab, offs = AtA.class_table(i)
                offs = np.ascontiguousarray(offs, dtype=np.int64)
                tab = np.ascontiguousarray(tab, dtype=TF)
                _lib.check(lib.sipb_problem_set_ata_classes(handle, i, tab.ctypes.data,
                                                            offs.ctypes.data_as(C.POINTER(C.c_int64)), offs.size))
                continue
            R = AtA[i]
            if slab is not None and getattr(AtA, "slab", None) is None:       # global CDS given: keep this rank's rows
                plane = op0.n[0] * op0.n[1]
                R = R[plane * slab[0]: plane * slab[1], :]
            R = np.asfortranarray(R, dtype=TF)
            offs = np.ascontiguousarray(set_Prop.AtA_offsets[i], dtype=np.int64)
            if R.shape[1] != offs.size:
                raise ValueError("AtA[%d] has %d diagonals but %d offsets" % (i, R.shape[1], offs.size))
            _lib.check(lib.sipb_problem_set_ata(handle, i, R.ctypes.data, R.shape[0],
                                                offs.ctypes.data_as(C.POINTER(C.c_int64)), offs.size))
        _lib.check(lib.sipb_problem_finalize(handle))
        nd = C.c_int(0)
        _lib.check(lib.sipb_problem_num_q_offsets(handle, C.byref(nd)))
        qo = np.zeros(nd.value, dtype=np.int64)
        _lib.check(lib.sipb_problem_q_offsets(handle, qo.ctypes.data_as(C.POINTER(C.c_int64))))
        qf = C.c_int(0)
        _lib.check(lib.sipb_problem_q_form(handle, C.byref(qf)))
    except Exception:
        lib.sipb_problem_destroy(handle)
        raise
    if slab is None:
        N = op0.npts * (2 if options.Minkowski else 1)
        rows = [A.rows for A in TD_OP]
    else:
        N = op0.n[0] * op0.n[1] * (slab[1] - slab[0])
        rows = [sum(b - a for a, b in dd.local_td_slices(A, *slab)) for A in TD_OP]
    dev = _DeviceProblem(handle, _problem_key(TF, TD_OP, P_sub, set_Prop, options), p, pp, N, rows, qo)
    dev.slab = slab
    dev.q_form = "classes" if qf.value == 1 else "arrays"     # how the device holds Q (sipb_problem_q_form)
    dev.N_global = op0.npts * (2 if options.Minkowski else 1)
    dev.rows_global = [A.rows for A in TD_OP]
    return dev


def device_problem(m_dtype, AtA, TD_OP, set_Prop, P_sub, comp_grid, options):
    """The cached device problem of (AtA, TD_OP, P_sub, ...) — built and uploaded on first use."""
    TF = np.dtype(m_dtype).type
    key = _problem_key(TF, TD_OP, P_sub, set_Prop, options)
    dev = getattr(AtA, "_device", None)
    if dev is None or dev.key != key:
        dev = build_device_problem(TF, AtA, TD_OP, set_Prop, P_sub, comp_grid, options)
        try:
            AtA._device = dev
        except AttributeError:
            pass
    return dev


class _Call:
    """Marshalled arguments of one sipb_solve call (kept alive until the call returns)."""


def _marshal(dev, m, TD_OP, options, x, l, y, profile_kernels, fixed_iterations, return_ly, resident_io, warm_resident) -> _Call:
    TF = m.dtype.type
    p, pp, N = dev.p, dev.pp, dev.N
    m = np.ascontiguousarray(m)
    slab = getattr(dev, "slab", None)
    if slab is not None:
        op0 = TD_OP[0]
        if m.size == dev.N_global:
            m = dd.scatter_model(m, op0.n, *slab)
        if x is not None and np.size(x) == dev.N_global:
            x = dd.scatter_model(np.ascontiguousarray(x, dtype=TF), op0.n, *slab)
        if (l is not None and y is not None and len(l) == p and len(y) == p
                and all(np.size(l[i]) == dev.rows_global[i] for i in range(p))):
            l = [dd.scatter_td(np.ascontiguousarray(l[i], dtype=TF), TD_OP[i], *slab) for i in range(p)]
            y = [dd.scatter_td(np.ascontiguousarray(y[i], dtype=TF), TD_OP[i], *slab) for i in range(p)]
    if options.Minkowski:
        if m.size * 2 != N:
            raise ValueError("Minkowski problems need length(m) == N/2")
    elif m.size != N:
        raise ValueError("m has %d entries, the operators have %d columns" % (m.size, N))
    zero_guess = bool(options.zero_ini_guess)
    if x is None:
        x_out = np.zeros(N, dtype=TF)
    else:
        x_out = np.ascontiguousarray(x, dtype=TF)
        if x_out.size != N:
            if options.Minkowski and x_out.size * 2 == N:
                x_out = np.concatenate([x_out, np.zeros(x_out.size, dtype=TF)])            # PARSDMM.jl:85-89
            else:
                raise ValueError("x has the wrong length")
    have_ly = l is not None and len(l) > 0 and y is not None and len(y) > 0
    if warm_resident and not zero_guess and not return_ly:
        have_ly, l, y = False, None, None
    if not zero_guess and not have_ly and not warm_resident:
        # PARSDMM_initialize.jl:120-127 allocates zeros when l / y are empty
        l = [np.zeros(r, dtype=TF) for r in dev.rows]
        y = [np.zeros(r, dtype=TF) for r in dev.rows]
        have_ly = True
    if return_ly and not have_ly:
        l = [np.zeros(r, dtype=TF) for r in dev.rows]
        y = [np.zeros(r, dtype=TF) for r in dev.rows]
        have_ly = True
    c = _Call()
    c.lp = c.yp = None
    if have_ly:
        l = [np.ascontiguousarray(v, dtype=TF) for v in l]
        y = [np.ascontiguousarray(v, dtype=TF) for v in y]
        for i in range(p):
            if l[i].size != dev.rows[i] or y[i].size != dev.rows[i]:
                raise ValueError("l[%d] / y[%d] have the wrong length" % (i, i))
        c.lp = (C.c_void_p * p)(*[v.ctypes.data for v in l])
        c.yp = (C.c_void_p * p)(*[v.ctypes.data for v in y])

    maxit = int(options.maxit)
    c.rho_ini = np.ascontiguousarray([float(v) for v in options.rho_ini], dtype=np.float64)
    o = _lib.Options()
    o.maxit, o.rho_update_frequency = maxit, int(options.rho_update_frequency)
    o.adjust_rho, o.adjust_gamma = int(bool(options.adjust_rho)), int(bool(options.adjust_gamma))
    o.adjust_feasibility_rho, o.zero_ini_guess = int(bool(options.adjust_feasibility_rho)), int(zero_guess)
    o.n_rho_ini, o.profile_kernels = c.rho_ini.size, int(bool(profile_kernels))
    o.evol_rel_tol, o.feas_tol, o.obj_tol = float(options.evol_rel_tol), float(options.feas_tol), float(options.obj_tol)
    o.gamma_ini = float(options.gamma_ini)
    o.rho_ini = c.rho_ini.ctypes.data_as(C.POINTER(C.c_double))
    o.fixed_iterations, o.return_ly = int(fixed_iterations), int(bool(return_ly and have_ly))
    o.resident_io = int(bool(resident_io))      # benchmark mode: no H2D/D2H (see include/sipb200.h)
    o.warm_resident = int(bool(warm_resident and not zero_guess))

    c.arr = {
        "set_feasibility": np.zeros((maxit + 2, max(pp, 1))), "r_dual": np.zeros((maxit, p)),
        "r_pri": np.zeros((maxit, p)), "r_dual_total": np.zeros(maxit), "r_pri_total": np.zeros(maxit),
        "obj": np.zeros(maxit), "evol_x": np.zeros(maxit), "rho": np.zeros((maxit, p)), "gamma": np.zeros((maxit, p)),
        "cg_relres": np.zeros(maxit),
    }
    c.cg_it = np.zeros(maxit, dtype=np.int32)
    lg = _lib.Log()
    for name, a in c.arr.items():
        setattr(lg, name, a.ctypes.data_as(C.POINTER(C.c_double)))
    lg.cg_it = c.cg_it.ctypes.data_as(C.POINTER(C.c_int32))
    c.m, c.x_in, c.x_out, c.l, c.y, c.o, c.lg, c.dev = m, x, x_out, l, y, o, lg, dev
    c.zero_guess, c.slab = zero_guess, slab
    return c


def _collect(c: _Call, options, gather_result, resident_io):
    lib = _lib.load()
    lg, arr, dev = c.lg, c.arr, c.dev
    pp = dev.pp
    it = max(int(lg.iters), 1)       # feasible input: logs trimmed to row 1 (PARSDMM.jl:70-80)
    rows_f = int(lg.feas_rows)
    timing = {name: float(lg.phase_seconds[i]) for i, name in enumerate(_lib.PHASE_NAMES)}
    timing["solve_seconds"] = float(lg.solve_seconds)
    timing["device_seconds"] = float(lg.device_seconds)
    timing["total_launches"] = int(lg.total_launches)
    timing["h2d_bytes"], timing["d2h_bytes"] = int(lg.h2d_bytes), int(lg.d2h_bytes)
    timing["kernels"] = {lib.sipb_kernel_class_name(i).decode(): (int(lg.kernel_launches[i]), float(lg.kernel_ms[i]))
                         for i in range(_lib.N_KERNEL_CLASSES) if lg.kernel_launches[i]}
    timing["kernel_bytes"] = {lib.sipb_kernel_class_name(i).decode(): float(lg.kernel_bytes[i])
                              for i in range(_lib.N_KERNEL_CLASSES) if lg.kernel_bytes[i]}
    timing["stopped_feasible"] = bool(lg.stopped_feasible)
    log = log_type_PARSDMM(
        set_feasibility=arr["set_feasibility"][:rows_f, :pp], r_dual=arr["r_dual"][:it], r_pri=arr["r_pri"][:it],
        r_dual_total=arr["r_dual_total"][:it], r_pri_total=arr["r_pri_total"][:it], obj=arr["obj"][:it],
        evol_x=arr["evol_x"][:it], rho=arr["rho"][:it], gamma=arr["gamma"][:it], cg_it=c.cg_it[:it].astype(np.int64),
        cg_relres=arr["cg_relres"][:it], timing=timing)
    x, x_out, l, y = c.x_in, c.x_out, c.l, c.y
    if lg.stopped_feasible and c.zero_guess and l is not None:
        # feasible input: the reference returns the zero start vectors (PARSDMM_initialize.jl:304-313, PARSDMM.jl:81)
        for v in list(l) + list(y):
            v[:] = 0
    if x is not None and isinstance(x, np.ndarray) and x.size == x_out.size and x is not x_out:
        x[:] = x_out                                                                       # in-place like the reference
        x_out = x
    if c.slab is not None and gather_result and not resident_io:
        x_out = dd.gather_model(x_out)
    return x_out, log, l, y


def _check_input(m, x, options):
    if not isinstance(m, np.ndarray) or m.dtype not in (np.float32, np.float64) or m.ndim != 1:
        raise TypeError("m must be a Float32/Float64 vector")
    if np.iscomplexobj(m) or (x is not None and np.iscomplexobj(x)):
        raise ValueError("input for PARSDMM is not real")                                  # PARSDMM.jl:50-52
    if getattr(options, "parallel", False):
        raise NotImplementedError("options.parallel=true is rejected on the device path (use slab decomposition)")


def PARSDMM(m, AtA, TD_OP, set_Prop, P_sub, comp_grid, options, x=None, l=None, y=None, *,
            profile_kernels=False, fixed_iterations=0, return_ly=True, resident_io=False, gather_result=True,
            warm_resident=False):
    """Project m onto the intersection of the sets; see PARSDMM.jl:25-35 for the arguments.
    Returns (x, log_PARSDMM, l, y).

    Multi-GPU slabs (after `distributed.init`, 3-D problems): every rank passes the same global `m`
    (or its own slab of it) and receives the global `x` (host-side gather) unless `gather_result=False`;
    `l`, `y` are this rank's slabs (see `distributed.gather_td`).

    `warm_resident=True` (with options.zero_ini_guess == False): the start vectors were already placed in the
    device buffers by `sipb_problem_warm_from` (multilevel driver); x, l, y are not uploaded."""
    _check_input(m, x, options)
    TF = m.dtype.type
    convert_options(options, TF)                                                           # PARSDMM.jl:43
    dev = device_problem(TF, AtA, TD_OP, set_Prop, P_sub, comp_grid, options)
    c = _marshal(dev, m, TD_OP, options, x, l, y, profile_kernels, fixed_iterations, return_ly, resident_io, warm_resident)
    lib = _lib.load()
    _lib.check(lib.sipb_solve(dev.handle, c.m.ctypes.data, c.x_out.ctypes.data, c.lp, c.yp, C.byref(c.o), C.byref(c.lg)))
    return _collect(c, options, gather_result, resident_io)


def PARSDMM_batch(ms, AtA, TD_OP, set_Prop, P_sub, comp_grid, options, *, return_ly=True):
    """B independent projections of the models `ms[b]` at once (one GPU): the per-channel PARSDMM calls of
    examples/Constraint_examples_2D.jl:221-226, or PARSDMM as the projector inside an outer loop
    (examples/Dykstra_parallel_vs_PARSDMM.jl:134,149).  `AtA, TD_OP, set_Prop, P_sub` are either ONE problem definition
    shared by all models or lists of B definitions (each channel its own constraints); the options are shared.
    Returns a list of B `(x, log, l, y)` tuples — the results of B separate `PARSDMM` calls, bit for bit.
    Every problem gets its own device context (stream); one host thread per problem drives it (sipb_solve_batch)."""
    B = len(ms)
    if B == 0:
        return []
    if dd.active():
        raise NotImplementedError("batched projections run as replicas: one GPU per process, no slab decomposition")
    shared = not (isinstance(AtA, (list, tuple)) and len(AtA) == B and isinstance(TD_OP[0], (list, tuple)))
    defs = [(AtA, TD_OP, set_Prop, P_sub)] * B if shared else list(zip(AtA, TD_OP, set_Prop, P_sub))
    TF = ms[0].dtype.type
    for m in ms:
        _check_input(m, None, options)
        if m.dtype.type != TF:
            raise TypeError("all models of a batch must share one float type")
    convert_options(options, TF)
    device = int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else 0
    calls = []
    for b, (A_b, T_b, S_b, P_b) in enumerate(defs):
        key = _problem_key(TF, T_b, P_b, S_b, options)
        cache = getattr(A_b, "_batch_devices", None)
        if cache is None:
            cache = {}
            try:
                A_b._batch_devices = cache
            except AttributeError:
                pass
        dev = cache.get(b)
        if dev is None or dev.key != key:
            dev = build_device_problem(TF, A_b, T_b, S_b, P_b, comp_grid, options, ctx=_lib.batch_ctx(device, b))
            cache[b] = dev
        calls.append(_marshal(dev, ms[b], T_b, options, None, None, None, False, 0, return_ly, False, False))
    lib = _lib.load()
    VP = C.c_void_p
    pbs = (VP * B)(*[c.dev.handle for c in calls])
    mp = (VP * B)(*[c.m.ctypes.data for c in calls])
    xp = (VP * B)(*[c.x_out.ctypes.data for c in calls])
    PVP = C.POINTER(VP)
    lps = (PVP * B)(*[C.cast(c.lp, PVP) if c.lp is not None else PVP() for c in calls])
    yps = (PVP * B)(*[C.cast(c.yp, PVP) if c.yp is not None else PVP() for c in calls])
    logs = (C.POINTER(_lib.Log) * B)(*[C.pointer(c.lg) for c in calls])
    rcs = (C.c_int * B)()
    _lib.check(lib.sipb_solve_batch(pbs, B, mp, xp, lps, yps, C.byref(calls[0].o), logs, rcs))
    return [_collect(c, options, False, False) for c in calls]
