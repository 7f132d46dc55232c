#!/usr/bin/env python
"""CDS SpMV(+dot) micro-benchmark through sipb_bench_spmv2: matrix form (arrays / classes) x kernel (generic / tiled).
Prints one JSON line per combination: CUDA-event ms per launch, algorithmic GB/s, fraction of the measured HBM peak.
  python tools/bench_spmv.py [--n 200 512] [--form arrays classes] [--kernel generic tiled] [--flush 0 1] [--reps 30]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sip_b200 as sip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[200, 512])
ap.add_argument("--form", nargs="+", default=["arrays", "classes"])
ap.add_argument("--kernel", nargs="+", default=["generic", "tiled"])
ap.add_argument("--flush", type=int, nargs="+", default=[0])
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--dtype", nargs="+", default=["f32"])
args = ap.parse_args()
L = sip._lib
peak = 6534.1
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
for n in args.n:
    for name in args.dtype:
        for form in args.form:
            for kern in args.kernel:
                for flush in args.flush:
                    ms, nb = C.c_double(0.0), C.c_int64(0)
                    n3 = (C.c_int64 * 3)(n, n, n)
                    L.check(L.load().sipb_bench_spmv2(L.ctx(), 0 if name == "f32" else 1, 3, n3, args.warmup, args.reps, flush,
                                                      1 if form == "classes" else 0, 1 if kern == "tiled" else 0,
                                                      C.byref(ms), C.byref(nb)))
                    gbs = nb.value / (ms.value * 1e-3) / 1e9
                    print(json.dumps({"grid": n, "dtype": name, "form": form, "kernel": kern, "l2_flush": bool(flush),
                                      "ms": round(ms.value, 4), "algorithmic_mb": round(nb.value / 1e6, 1), "gbs": round(gbs, 1),
                                      "frac_of_measured_peak": round(gbs / peak, 3)}), flush=True)
