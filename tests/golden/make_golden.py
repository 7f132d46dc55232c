#!/usr/bin/env python
"""Generates the golden fixtures in this directory with the CPU oracle (oracle/, a restatement of the reference's
algorithm; the Julia reference itself cannot run here and ships no stored PARSDMM outputs — see DESIGN.md §2).

    python tests/golden/make_golden.py

Fixtures (small .npz files, committed): inputs are regenerated from seeds by tests/problems.py, the files hold
the oracle's final x, the iteration count and the logged scalars.  `reference_kats.json` holds the literal
known answers of the reference's own tests (with file:line)."""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as pr  # noqa: E402

CASES = {
    "config1_f64_32x32": (lambda: pr.spec_config1((32, 32), np.float64), {}),
    "config1_f32_48x40": (lambda: pr.spec_config1((48, 40), np.float32), {}),
    "config2_f32_16x14x12": (lambda: pr.spec_config2((16, 14, 12), np.float32),
                             {"evol_rel_tol": 10 * float(np.finfo(np.float32).eps), "maxit": 60}),
    "config3_f32_14x12x10": (lambda: pr.spec_config3((14, 12, 10), np.float32), {"maxit": 30}),
    "config4_f64_12x14x10": (lambda: pr.spec_config4((12, 14, 10), np.float64),
                             {"rho_ini": [1.0, 1000.0, 1000.0, 1000.0, 1.0], "maxit": 50}),
}


def run_case(api, name):
    make, kw = CASES[name]
    spec = make()
    opt = api.PARSDMM_options()
    for k, v in kw.items():
        setattr(opt, k, v)
    b = pr.build(api, copy.deepcopy(spec), opt)
    x, log, l, y = api.PARSDMM(spec["m"].copy(), b["AtA"], b["TD_OP"], b["set_Prop"], b["P_sub"], b["cg"], b["opt"])
    return spec, b, x, log, l, y


def main():
    orc = pr.OracleAPI()
    for name in CASES:
        spec, b, x, log, l, y = run_case(orc, name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, iterations=len(log.obj), cg_it=log.cg_it,
                            rho=log.rho, gamma=log.gamma, obj=log.obj, set_feasibility=log.set_feasibility,
                            r_pri=log.r_pri, m_checksum=np.float64(np.sum(spec["m"].astype(np.float64))),
                            q_offsets=orc.ops.assemble_Q(b["AtA"], b["set_Prop"].AtA_offsets, np.ones(len(b["AtA"])))[1],
                            ata_offsets_1=b["set_Prop"].AtA_offsets[1], y_last_support=(y[-2] != 0))
        print(name, "iterations", len(log.obj), "sum cg", int(log.cg_it.sum()))
    kats = {
        "prox_l2s": {"x": [2.0], "m": [1.0], "rho": 3.0, "expect": [1.75], "source": "test/test_prox_l2s!.jl:15-19"},
        "cardinality": [
            {"x": [0, 0, 1, 2, 3], "k": 2, "expect": [0, 0, 0, 2, 3], "source": "test/test_projectors.jl:49-52"},
            {"x": [0, 0, -1, 2, -3], "k": 2, "expect": [0, 0, 0, 2, -3], "source": "test/test_projectors.jl:54-56"}],
        "tv_2d_cross": {"n": [9, 6], "h": [0.99, 1.123], "ones_col": 3, "ones_row": 4,
                        "expect": "D_x*vec(x) == diff(x,dims=1)./h1 and D_z*vec(x) == diff(x,dims=2)./h2 exactly; "
                                  "TV = vcat(D_z, D_x)", "source": "test/test_TD_OPs.jl:5-40"},
        "cg_exact_start": {"expect": "iter == 1 and x == xt", "source": "test/test_cg.jl:25-29"},
        "q_offsets_order": {"expect": "unique() over the zero-padded 999x99 table, column major",
                            "source": "src/PARSDMM_initialize.jl:217-221"},
    }
    json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
