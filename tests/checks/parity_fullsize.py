#!/usr/bin/env python
"""Parity of the device path against the CPU oracle (threaded C/OpenMP port) at grid sizes beyond what the test-suite
runs: BASELINE configs[2] (bounds ∩ TV l1 ∩ cardinality of the gradient) at --size^3 for --iters PARSDMM iterations.
The oracle's sparse set-up and stable sorts dominate the run time (minutes at 256^3).  One JSON line per run; keep the
output under profiles/.   python tests/checks/parity_fullsize.py --size 256 --iters 12"""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import problems as pr  # noqa: E402
import sip_b200 as sip  # noqa: E402
from oracle import cpu_baseline as cb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--workload", default="config3", choices=["config3", "config2"])
args = ap.parse_args()
cb.use_all_cores()
orc = pr.OracleAPI()
n = (args.size,) * 3
spec = (pr.spec_config3 if args.workload == "config3" else pr.spec_config2)(n, np.float32)


def opts(api):
    o = api.PARSDMM_options()
    o.maxit, o.evol_rel_tol = args.iters, 10 * float(np.finfo(np.float32).eps)
    return o


t0 = time.perf_counter()
sb = pr.build(sip, copy.deepcopy(spec), opts(sip))
xs, ls, l2, y2 = sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
ob = pr.build(orc, copy.deepcopy(spec), opts(orc))
xo, lo, ll, yy = cb.PARSDMM(spec["m"].copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"], constraint=ob["cons"])
t_cpu = time.perf_counter() - t0
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))   # noqa: E731
same = [i for i in range(min(len(ls.obj), len(lo.obj))) if ls.obj[i] == lo.obj[i] and np.array_equal(ls.rho[i], lo.rho[i])]
first_diff = next((i + 1 for i in range(min(len(ls.obj), len(lo.obj))) if ls.obj[i] != lo.obj[i]), None)
out = {"workload": args.workload, "grid": list(n), "iterations_device": len(ls.obj), "iterations_oracle": len(lo.obj),
       "cg_it_equal": bool(np.array_equal(ls.cg_it, lo.cg_it)), "rel_l2_x": rel(xs, xo),
       "iterations_with_bit_identical_obj_and_rho": len(same), "first_iteration_with_different_obj": first_diff,
       "device_wall_s_incl_setup": round(t_dev, 2), "oracle_wall_s_incl_setup": round(t_cpu, 2), "oracle_threads": cb.threads()}
if args.workload == "config3":
    k = int(np.count_nonzero(yy[2]))
    out["cardinality_k"] = k
    out["support_entries_different"] = int(np.count_nonzero((y2[2] != 0) != (yy[2] != 0)))
print(json.dumps(out))
