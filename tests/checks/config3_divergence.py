#!/usr/bin/env python
"""BASELINE configs[2] (non-convex: cardinality of the gradient) at n^3, 30 iterations: device vs the C/OpenMP port vs the
NumPy oracle, iteration by iteration (obj, r_pri, rho) and the supports of the cardinality set's y at the end — the
measurement quoted in tests/test_gpu_parity.py::test_config3_128cubed_f32 and DESIGN.md §7.  Needs a GPU.

  python tests/checks/config3_divergence.py [n=128]
"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import problems as pr  # noqa: E402
import sip_b200 as sip  # noqa: E402
from oracle import cpu_baseline as cb  # noqa: E402

cb.use_all_cores()
orc = pr.OracleAPI()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
spec = pr.spec_config3((n, n, n), np.float32)
def run(api, maxit, cpu=False):
    opt = api.PARSDMM_options(); opt.maxit = maxit; opt.evol_rel_tol = 10*float(np.finfo(np.float32).eps)
    b = pr.build(api, copy.deepcopy(spec), opt)
    if cpu:
        return cb.PARSDMM(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"], constraint=b["cons"])
    return api.PARSDMM(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
re = lambda a, b: float(np.linalg.norm(a.astype(np.float64)-b)/np.linalg.norm(b.astype(np.float64)))
xs, ls, l2, y2 = run(sip, 30)
xc, lc, lc2, yc2 = run(orc, 30, cpu=True)
print("dev vs C-baseline: relx", re(xs, xc), "cg", np.array_equal(ls.cg_it, lc.cg_it))
for i in range(len(ls.obj)):
    print(i+1, "obj rel", abs(ls.obj[i]-lc.obj[i])/abs(lc.obj[i]), "rpri", np.abs(ls.r_pri[i]-lc.r_pri[i]).max()/np.abs(lc.r_pri[i]).max(), "rho", ls.rho[i], lc.rho[i])
print("support mismatch", np.count_nonzero((y2[2]!=0)!=(yc2[2]!=0)), "k", np.count_nonzero(yc2[2]), np.count_nonzero(y2[2]))
if n <= 128:
    xo, lo, lo2, yo2 = run(orc, 30)
    print("dev vs numpy-oracle relx", re(xs, xo), " C vs numpy relx", re(xc, xo))
    print("support mismatch dev-numpy", np.count_nonzero((y2[2]!=0)!=(yo2[2]!=0)), " C-numpy", np.count_nonzero((yc2[2]!=0)!=(yo2[2]!=0)))
