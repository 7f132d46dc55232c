#!/usr/bin/env python
"""Why tests/test_gpu_parity.py::test_parsdmm_cardinality_ties_in_the_loop compares TWO iterations only.

Problem: cardinality(identity) on an integer-valued model — Q is a multiple of the identity, the CG scales every entry
alike, so whole classes of rows stay bit-identical and the k-th largest magnitude is shared by hundreds of rows.  From
iteration 3 on, classes that are EQUAL in exact arithmetic reach the sort through different roundings and their order —
hence the support — depends on the last ulp of the reductions: the three CPU restatements of the reference (NumPy oracle
with Float64-accumulated reductions, with NumPy's native reductions, and the C/OpenMP port) disagree with EACH OTHER.
This script prints, per iteration count, whether they end in the same support.  CPU only (oracle/ is test infrastructure).

  python tests/checks/tie_sensitivity.py [maxit ...]          default: 2 3 4 8
"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import problems as pr  # noqa: E402
from oracle import cpu_baseline as cb  # noqa: E402
from oracle import sip_types as stt  # noqa: E402

orc = pr.OracleAPI()


def spec_ties(n, TF, frac, seed):
    rng = np.random.default_rng(seed)
    N = int(np.prod(n))
    m = np.round(rng.standard_normal(N) * 3).astype(TF)
    return dict(n=n, d=(1.0, 1.0), TF=TF, m=m, sets=[("cardinality", "identity", 0, int(frac * N))], mode="matrix")


def run(spec, mode, maxit, api=None):
    stt.REDUCTION_MODE = mode
    opt = orc.PARSDMM_options()
    opt.maxit = maxit
    b = pr.build(orc, copy.deepcopy(spec), opt)
    f = orc.PARSDMM if api is None else api
    return f(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])


def main():
    its = [int(a) for a in sys.argv[1:]] or [2, 3, 4, 8]
    for maxit in its:
        for TF in (np.float32, np.float64):
            for n, frac, seed in (((72, 60), 0.4, 5), ((96, 64), 0.3, 6), ((640, 512), 0.4, 8)):
                spec = spec_ties(n, TF, frac, seed)
                a = run(spec, "f64acc", maxit)
                b = run(spec, "native", maxit)
                c = run(spec, "f64acc", maxit, cb.PARSDMM)
                same = [bool(np.array_equal(a[3][0] != 0, o[3][0] != 0)) for o in (b, c)]
                print("maxit %2d %-8s %-10s y support equal to the f64acc oracle: native %s, C port %s" %
                      (maxit, np.dtype(TF).name, "x".join(map(str, n)), same[0], same[1]), flush=True)
    stt.REDUCTION_MODE = "f64acc"


if __name__ == "__main__":
    main()
