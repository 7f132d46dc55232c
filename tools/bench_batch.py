#!/usr/bin/env python
"""Batched independent small projections: B x BASELINE configs[0] (2-D 256x256 Float64, bounds ∩ TV
l1-ball ∩ vertical slope bounds) through sip.PARSDMM_batch vs B sequential sip.PARSDMM calls.  One JSON line."""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import problems as pr  # noqa: E402
import sip_b200 as sip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, nargs="+", default=[1, 4, 16, 64])
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
TF = np.float64
n = (args.n, args.n)
peak = 6534.1
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
spec = pr.spec_config1(n, TF)
opt = sip.PARSDMM_options()
opt.maxit = 500
sb = pr.build(sip, copy.deepcopy(spec), opt)
Bmax = max(args.batch)
ms = [pr.synthetic_model(n, TF, seed=1000 + b) for b in range(Bmax)]
call = lambda m: sip.PARSDMM(m, sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], return_ly=False)   # noqa: E731
call(ms[0])
t0 = time.perf_counter()
seq_its = 0
for b in range(min(Bmax, 16)):
    seq_its += len(call(ms[b])[1].obj)
t_seq = (time.perf_counter() - t0) / min(Bmax, 16)
out = {"workload": "B x 2D %dx%d Float64 bounds ∩ TV l1 ∩ D_z slope bounds (BASELINE configs[0])" % n,
       "sequential": {"ms_per_projection": 1e3 * t_seq, "iterations_per_s": seq_its / (t_seq * min(Bmax, 16))}, "batched": []}
for B in args.batch:
    f = lambda: sip.PARSDMM_batch(ms[:B], sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], return_ly=False)   # noqa: E731
    f()
    t0 = time.perf_counter()
    its, nbytes, launches = 0, 0.0, 0
    for _ in range(args.reps):
        for x, lg, _, _ in f():
            its += len(lg.obj)
            nbytes += sum(lg.timing["kernel_bytes"].values())
            launches += lg.timing["total_launches"]
    t = (time.perf_counter() - t0) / args.reps
    out["batched"].append({"B": B, "ms_per_batch": 1e3 * t, "ms_per_projection": 1e3 * t / B, "projections_per_s": B / t,
                           "iterations_per_s": its / args.reps / t, "speedup_vs_sequential": t_seq * B / t,
                           "algorithmic_gbs": nbytes / args.reps / t / 1e9, "frac_of_measured_peak": nbytes / args.reps / t / 1e9 / peak,
                           "launches_per_batch": launches // args.reps})
print(json.dumps(out))
