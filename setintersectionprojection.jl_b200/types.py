"""Host-side mirror of the reference's public types (SetIntersectionProjection.jl:95-149).

Same names and field meanings as the Julia structs so that scripts written against the reference read
the same here; Julia's `TF` (Float32/Float64) is a NumPy scalar type.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Tuple

import numpy as np


@dataclass
class compgrid:
    """Computational grid: spacing `d` and point counts `n` (duck-typed in the reference,
    test/runtests.jl:18-21)."""
    d: Tuple
    n: Tuple


@dataclass
class PARSDMM_options:
    """SetIntersectionProjection.jl:110-128."""
    x_min_solver: str = "CG_normal"
    maxit: int = 200
    evol_rel_tol: Any = 1e-3
    feas_tol: Any = 5e-2
    obj_tol: Any = 1e-3
    rho_ini: List[Any] = field(default_factory=lambda: [10.0])
    rho_update_frequency: int = 2
    gamma_ini: Any = 1.0
    adjust_rho: bool = True
    adjust_gamma: bool = True
    adjust_feasibility_rho: bool = True
    Blas_active: bool = True           # both code paths of the reference are one device formula
    feasibility_only: bool = False
    FL: Any = np.float32
    parallel: bool = False             # set-parallel Julia workers: rejected on the device path
    zero_ini_guess: bool = True
    Minkowski: bool = False


def default_PARSDMM_options(options: PARSDMM_options, TF) -> PARSDMM_options:
    """default_PARSDMM_options.jl:6-34 (line 30 of the reference assigns a local variable, so the
    Minkowski field keeps its value)."""
    d = PARSDMM_options()
    for name in ("x_min_solver", "maxit", "rho_update_frequency", "adjust_rho", "adjust_gamma",
                 "adjust_feasibility_rho", "Blas_active", "feasibility_only", "parallel", "zero_ini_guess"):
        setattr(options, name, getattr(d, name))
    options.evol_rel_tol = TF(1e-3)
    options.feas_tol = TF(5e-2)
    options.obj_tol = TF(1e-3)
    options.rho_ini = [TF(10.0)]
    options.gamma_ini = TF(1.0)
    options.FL = TF
    return options


def convert_options(options: PARSDMM_options, TF) -> None:
    """convert_options!.jl:6-15: cast the float-valued options to TF (in place)."""
    for name in ("evol_rel_tol", "feas_tol", "obj_tol", "gamma_ini"):
        setattr(options, name, TF(getattr(options, name)))
    options.rho_ini = [TF(v) for v in options.rho_ini]


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
    """SetIntersectionProjection.jl:132-140 (one entry per set in every list)."""
    ncvx: List[bool] = field(default_factory=list)
    AtA_diag: List[bool] = field(default_factory=list)
    dense: List[bool] = field(default_factory=list)
    TD_n: List[Tuple] = field(default_factory=list)
    tag: List[Tuple[str, str, str, str]] = field(default_factory=list)
    banded: List[bool] = field(default_factory=list)
    AtA_offsets: List[Any] = field(default_factory=list)


@dataclass
class log_type_PARSDMM:
    """SetIntersectionProjection.jl:95-108.  `timing` maps the reference's seven TimerOutputs section
    names (PARSDMM.jl:40,100,105,113,152,163,229) to seconds and carries the device-side extras
    (kernel table, launch count, transfer bytes)."""
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
    timing: Dict[str, Any]
