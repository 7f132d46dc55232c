"""CPU oracle for the PARSDMM projection iteration — TEST INFRASTRUCTURE ONLY.

This package is a line-faithful NumPy/SciPy restatement of the hot path of
slimgroup/SetIntersectionProjection.jl v0.2.5 (reference tree at /root/reference,
pure Julia).  Every function cites the reference file:line it follows.

Rules (tier contract):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything from here;
  * the product package (``setintersectionprojection.jl_b200``) never imports it and never
    falls back to it — it fails loudly when the CUDA library is missing.

Parity status: the reference ships no stored PARSDMM output vectors and Julia is not
installed in this image (nor on the GPU box), so the reference binary cannot be executed.
The oracle is pinned against every deterministic known-answer test the reference's own
test-suite holds for this path (see tests/test_oracle_kat.py: test_TD_OPs.jl, test_prox_l2s!.jl,
test_projectors.jl literals, test_cg.jl, test_CDS_Mvp.jl, test_CDS_scaled_add.jl, test_Q_update.jl,
test_update_y_l.jl formulas, test_PARSDMM.jl feasibility properties).  For full PARSDMM runs
"parity unpinned" applies in the strict sense: there is no reference output to compare with,
only the restated algorithm.

Floating-point discipline: ``TF`` is np.float32 or np.float64.  All vector arithmetic stays in
TF; the Float64 literals the Julia code mixes in (``0.1*…`` argmin_x.jl:34, ``rho .+ 1.0``
prox_l2s!.jl:4) are reproduced by explicit np.float64 promotion.  Reductions (dot, norm, asum)
are summation-order-unpinned in the reference (OpenBLAS); ``REDUCTION_MODE`` selects
"f64acc" (accumulate in float64, round once to TF — the correctly rounded member of that family,
default) or "native" (NumPy's own TF pairwise reduction).
"""
from . import sip_types, operators, projectors, setup, parsdmm, multilevel  # noqa: F401
