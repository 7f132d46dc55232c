"""Oracle restatement of the reference's public types (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/SetIntersectionProjection.jl:95-149,
default_PARSDMM_options.jl:6-34 and convert_options!.jl:6-15.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Tuple

import numpy as np

# "f64acc": accumulate reductions in float64 and round once to TF; "native": NumPy TF reductions.
REDUCTION_MODE = "f64acc"


def set_reduction_mode(mode: str) -> None:
    global REDUCTION_MODE
    assert mode in ("f64acc", "native")
    REDUCTION_MODE = mode


def dot(a: np.ndarray, b: np.ndarray):
    """dot(a,b) returning a TF scalar (LinearAlgebra.dot -> BLAS dot; order unpinned)."""
    TF = a.dtype.type
    if REDUCTION_MODE == "f64acc":
        return TF(np.dot(a.astype(np.float64, copy=False), b.astype(np.float64, copy=False)))
    return TF(np.dot(a, b))


def norm2(a: np.ndarray):
    """norm(a) returning a TF scalar (BLAS nrm2; order unpinned)."""
    TF = a.dtype.type
    if REDUCTION_MODE == "f64acc":
        a64 = a.astype(np.float64, copy=False)
        return TF(np.sqrt(np.dot(a64, a64)))
    return TF(np.sqrt(TF(np.dot(a, a))))


def norm1(a: np.ndarray):
    """norm(a,1) returning a TF scalar (BLAS asum)."""
    TF = a.dtype.type
    if REDUCTION_MODE == "f64acc":
        return TF(np.sum(np.abs(a), dtype=np.float64))
    return TF(np.sum(np.abs(a)))


def eps(TF) -> Any:
    return TF(np.finfo(TF).eps)


@dataclass
class compgrid:
    """Duck-typed computational grid (test/runtests.jl:18-21)."""
    d: Tuple
    n: Tuple


@dataclass
class PARSDMM_options:
    """SetIntersectionProjection.jl:110-128 (defaults identical)."""
    x_min_solver: str = "CG_normal"
    maxit: int = 200
    evol_rel_tol: Any = 1e-3
    feas_tol: Any = 5e-2
    obj_tol: Any = 1e-3
    rho_ini: Any = field(default_factory=lambda: [10.0])
    rho_update_frequency: int = 2
    gamma_ini: Any = 1.0
    adjust_rho: bool = True
    adjust_gamma: bool = True
    adjust_feasibility_rho: bool = True
    Blas_active: bool = True
    feasibility_only: bool = False
    FL: Any = np.float32
    parallel: bool = False
    zero_ini_guess: bool = True
    Minkowski: bool = False


def default_PARSDMM_options(options: PARSDMM_options, TF) -> PARSDMM_options:
    """default_PARSDMM_options.jl:6-34 (note :30 assigns a local `Minkowski`, not the field)."""
    options.x_min_solver = "CG_normal"
    options.maxit = 200
    options.evol_rel_tol = TF(1e-3)
    options.feas_tol = TF(5e-2)
    options.obj_tol = TF(1e-3)
    options.rho_ini = [TF(10.0)]
    options.rho_update_frequency = 2
    options.gamma_ini = TF(1.0)
    options.adjust_rho = True
    options.adjust_gamma = True
    options.adjust_feasibility_rho = True
    options.Blas_active = True
    options.feasibility_only = False
    options.FL = TF
    options.parallel = False
    options.zero_ini_guess = True
    return options


def convert_options(options: PARSDMM_options, TF) -> None:
    """convert_options!.jl:6-15."""
    options.evol_rel_tol = TF(options.evol_rel_tol)
    options.feas_tol = TF(options.feas_tol)
    options.obj_tol = TF(options.obj_tol)
    options.rho_ini = [TF(r) for r in options.rho_ini]
    options.gamma_ini = TF(options.gamma_ini)


@dataclass
class set_definitions:
    """SetIntersectionProjection.jl:142-149."""
    set_type: str
    TD_OP: str
    min: Any
    max: Any
    app_mode: Tuple[str, str]
    custom_TD_OP: Tuple[Any, bool] = ((), False)


@dataclass
class set_properties:
    """SetIntersectionProjection.jl:132-140."""
    ncvx: List[bool]
    AtA_diag: List[bool]
    dense: List[bool]
    TD_n: List[Tuple]
    tag: List[Tuple[str, str, str, str]]
    banded: List[bool]
    AtA_offsets: List[Any]


@dataclass
class log_type_PARSDMM:
    """SetIntersectionProjection.jl:95-108.  Entries hold TF values embedded in float64 arrays."""
    set_feasibility: np.ndarray
    r_dual: np.ndarray
    r_pri: np.ndarray
    r_dual_total: np.ndarray
    r_pri_total: np.ndarray
    obj: np.ndarray
    evol_x: np.ndarray
    rho: np.ndarray
    gamma: np.ndarray
    cg_it: np.ndarray
    cg_relres: np.ndarray
    timing: dict
