"""B200-native PARSDMM projection iteration behind the API of SetIntersectionProjection.jl.

Host-side mirror of the reference interface for the hot path only: the names, argument
meanings and error behaviour of `setup_constraints`, `PARSDMM_precompute_distribute[_Minkowski]`,
`PARSDMM`, the option / set / log types.  All compute goes through the C ABI of `libsipb200.so`
(hand-written sm_100a CUDA kernels); there is no CPU fallback.

The directory name contains a dot, so import it through the `sip_b200` shim at the repository root:

    import sip_b200 as sip
    P_sub, TD_OP, set_Prop = sip.setup_constraints(constraint, comp_grid, np.float32)
    TD_OP, AtA, l, y = sip.PARSDMM_precompute_distribute(TD_OP, set_Prop, comp_grid, options)
    x, log, l, y = sip.PARSDMM(m, AtA, TD_OP, set_Prop, P_sub, comp_grid, options)
"""
from . import _lib
from .types import (PARSDMM_options, compgrid, convert_options, default_PARSDMM_options, log_type_PARSDMM,
                    set_definitions, set_properties)
from .operators import (CDS_MVp, CDS_MVp_MT, CDS_scaled_add, SparseOperator, TDOperator, cg, get_TD_operator, get_discrete_Grad,
                        mat2CDS)
from .constraints import Projector, get_projector, setup_constraints
from .precompute import PARSDMM_precompute_distribute, PARSDMM_precompute_distribute_Minkowski
from .solver import PARSDMM, PARSDMM_batch
from .multilevel import (PARSDMM_multi_level, constraint2coarse, interpolate_y_l, resample_nn,
                         setup_multi_level_PARSDMM)

__all__ = [
    "PARSDMM", "PARSDMM_batch", "PARSDMM_multi_level", "PARSDMM_options", "constraint2coarse", "interpolate_y_l", "resample_nn",
    "setup_multi_level_PARSDMM", "PARSDMM_precompute_distribute", "PARSDMM_precompute_distribute_Minkowski",
    "Projector", "TDOperator", "SparseOperator", "CDS_MVp", "CDS_MVp_MT", "CDS_scaled_add", "cg", "compgrid", "convert_options",
    "default_PARSDMM_options", "get_TD_operator", "get_discrete_Grad", "get_projector", "log_type_PARSDMM",
    "mat2CDS", "set_definitions", "set_properties", "setup_constraints",
]
