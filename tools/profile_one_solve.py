#!/usr/bin/env python
"""One PARSDMM projection of a bench workload (default BASELINE configs[2] at 512^3, --workload config2 for configs[1]) and nothing else — the process to put under
`ncu` for a launch list / a full capture without paying for bench.py's warm-up, e2e and profiled solves.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_one_solve.py [--size 200] [--solves 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import problems as pr  # noqa: E402
import sip_b200 as sip  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--solves", type=int, default=1)
    ap.add_argument("--workload", default="config3", choices=["config3", "config2"])
    ap.add_argument("--maxit", type=int, default=200)
    a = ap.parse_args()
    size = a.size or (512 if a.workload == "config3" else 200)
    spec = (pr.spec_config3 if a.workload == "config3" else pr.spec_config2)((size,) * 3, np.float32)
    opt = sip.PARSDMM_options()
    opt.maxit = a.maxit
    opt.evol_rel_tol = 10 * float(np.finfo(np.float32).eps)       # examples/test_scaling_3D.jl:25
    sb = pr.build(sip, spec, opt)
    for _ in range(a.solves):
        x, log, _, _ = sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"],
                                   return_ly=False)
    print("iterations", len(log.obj), "cg", int(np.sum(log.cg_it)), "launches", log.timing.get("total_launches"),
          "q_form", sb["AtA"]._device.q_form)


if __name__ == "__main__":
    main()
