#!/usr/bin/env python
"""Run the five BASELINE.json configurations once on the device and print one JSON line each
(iterations, wall seconds through the public API incl. H2D/D2H, device seconds, iterations/s, feasibility).
Usage: python tools/run_configs.py [1 2 3 4 5] [--small]"""
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


def timed(fn, reps=2):
    out = None
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return out, best


def report(name, grid, TF, log, wall, extra=None):
    it = len(log.obj)
    line = {"config": name, "grid": list(grid), "dtype": np.dtype(TF).name, "iterations": it, "wall_s": wall,
            "device_s": log.timing.get("device_seconds"), "iterations_per_s": it / wall,
            "cg_iterations": int(np.sum(log.cg_it)), "launches": log.timing.get("total_launches"),
            "final_feasibility": [float(v) for v in log.set_feasibility[max(log.set_feasibility.shape[0] - 2, 0)]]}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def main():
    small = "--small" in sys.argv
    which = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 3, 4, 5]
    if 1 in which:
        n = (64, 64) if small else (256, 256)
        spec = pr.spec_config1(n, np.float64)
        opt = sip.PARSDMM_options()
        opt.maxit = 500
        sb = pr.build(sip, spec, opt)
        f = lambda: sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])   # noqa: E731
        (x, log, _, _), wall = timed(f, 3)
        report("1: 2D bounds ∩ TV-l1 ∩ D_z slope bounds", n, np.float64, log, wall)
    if 2 in which:
        n = (48,) * 3 if small else (200,) * 3
        spec = pr.spec_config2(n, np.float32)
        opt = sip.PARSDMM_options()
        opt.evol_rel_tol = 10 * float(np.finfo(np.float32).eps)
        sb = pr.build(sip, spec, opt)
        f = lambda: sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], return_ly=False)   # noqa: E731
        (x, log, _, _), wall = timed(f)
        report("2: 3D bounds ∩ anisotropic TV ∩ lateral smoothness", n, np.float32, log, wall)
    if 3 in which:
        n = (48,) * 3 if small else (512,) * 3
        spec = pr.spec_config3(n, np.float32)
        opt = sip.PARSDMM_options()
        sb = pr.build(sip, spec, opt)
        f = lambda: sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], return_ly=False)   # noqa: E731
        (x, log, _, _), wall = timed(f, 1)
        report("3: 3D bounds ∩ TV-l1 ∩ cardinality(TV), single GPU", n, np.float32, log, wall)
    if 4 in which:
        n = (48,) * 3 if small else (400,) * 3
        spec = pr.spec_config4(n, np.float32)
        cg = sip.compgrid(tuple(spec["d"]), tuple(spec["n"]))
        cons = [sip.set_definitions(st, op, lo, hi, ("tensor", "")) for (st, op, lo, hi) in spec["sets"]]
        opt = sip.PARSDMM_options()
        opt.FL = np.float32
        opt.evol_rel_tol = 10 * float(np.finfo(np.float32).eps)
        opt.rho_ini = [1.0, 1000.0, 1000.0, 1000.0, 1.0]
        t = time.perf_counter()
        lv = sip.setup_multi_level_PARSDMM(spec["m"], 3, 2, cg, cons, opt)
        t_setup = time.perf_counter() - t
        f = lambda: sip.PARSDMM_multi_level(spec["m"].copy(), *lv[:5], opt)   # noqa: E731
        (x, log, _, _), wall = timed(f)
        report("4: multilevel PARSDMM, 3 levels, coarsening 2", n, np.float32, log, wall,
               {"level_iterations": log.timing["level_iterations"], "setup_s": t_setup,
                "device_s_levels": [lt["device_seconds"] for lt in log.timing["levels"]]})
    if 5 in which:
        n = (128, 128) if small else (1024, 1024)
        opt = sip.PARSDMM_options()
        opt.maxit = 500
        sb = pr.build_minkowski(sip, n, np.float32, opt)
        f = lambda: sip.PARSDMM(sb["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"], return_ly=False)   # noqa: E731
        (x, log, _, _), wall = timed(f)
        report("5: generalized Minkowski set, 2D bounds + TV decomposition", n, np.float32, log, wall)


if __name__ == "__main__":
    main()
