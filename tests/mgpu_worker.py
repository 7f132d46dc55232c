"""Worker of the multi-GPU slab parity tests; launched with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py
Every rank runs the slab-decomposed device PARSDMM; rank 0 also runs the CPU oracle on the full problem
and compares (iteration counts, logs, x, and the gathered y / l in the reference's global ordering)."""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems as pr  # noqa: E402
import sip_b200 as sip  # noqa: E402
from sip_b200 import distributed as dd  # noqa: E402

TOL = {np.float32: 1e-3, np.float64: 1e-5}


def relerr(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    dd.init(rank, world, local)
    orc = pr.OracleAPI() if rank == 0 else None
    if rank == 0:
        print("[mgpu %d ranks] peer-memory path: %s" % (world, dd.peer_path()), flush=True)
    cases = [
        ("config2_f32", pr.spec_config2((24, 20, 17), np.float32), dict(maxit=60, evol_rel_tol=10 * np.finfo(np.float32).eps)),
        ("config2_f64", pr.spec_config2((16, 12, 11), np.float64), dict(maxit=80)),
        ("config4_dz", pr.spec_config4((20, 18, 13), np.float32), dict(maxit=50, rho_ini=[1.0, 1000.0, 1000.0, 1000.0, 1.0])),
        ("config3_card", pr.spec_config3((20, 16, 15), np.float32), dict(maxit=40)),
    ]
    ok = True
    for name, spec, kw in cases:
        TF = spec["TF"]
        s_opt = sip.PARSDMM_options()
        for k, v in kw.items():
            setattr(s_opt, k, v)
        sb = pr.build(sip, copy.deepcopy(spec), s_opt)
        assert sb["AtA"].slab == dd.slab_range(spec["n"][2])
        xs, ls, l2, y2 = sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"])
        yg = [dd.gather_td(v, A) for v, A in zip(y2, sb["TD_OP"])]
        lg = [dd.gather_td(v, A) for v, A in zip(l2, sb["TD_OP"])]
        if rank == 0:
            o_opt = orc.PARSDMM_options()
            for k, v in kw.items():
                setattr(o_opt, k, v)
            ob = pr.build(orc, copy.deepcopy(spec), o_opt)
            xo, lo, ll, yy = orc.PARSDMM(spec["m"].copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
            tol = TOL[TF]
            checks = {
                "iters": len(ls.obj) == len(lo.obj),
                "cg_it": np.array_equal(ls.cg_it, lo.cg_it),
                "x": relerr(xs, xo) < tol,
                "feas": np.allclose(ls.set_feasibility, lo.set_feasibility, rtol=50 * tol, atol=1e-12) if ls.set_feasibility.shape == lo.set_feasibility.shape else False,
                "rho": np.allclose(ls.rho, lo.rho, rtol=50 * tol) if ls.rho.shape == lo.rho.shape else False,
                "obj": np.allclose(ls.obj, lo.obj, rtol=50 * tol) if ls.obj.shape == lo.obj.shape else False,
                "r_dual": np.allclose(ls.r_dual, lo.r_dual, rtol=1e-2, atol=1e-5 * np.abs(lo.r_dual).max()) if ls.r_dual.shape == lo.r_dual.shape else False,
                "y": all(relerr(a, b) < 100 * tol for a, b in zip(yg, yy)),
                "l": all(np.linalg.norm(a.astype(np.float64) - b) <= 100 * tol * np.linalg.norm(b.astype(np.float64)) +
                         1e4 * np.finfo(TF).eps * np.linalg.norm(yb.astype(np.float64)) for a, b, yb in zip(lg, ll, yy)),
            }
            if name == "config3_card":
                checks["support"] = bool(np.array_equal(yg[2] != 0, yy[2] != 0))
            good = all(checks.values())
            ok = ok and good
            print("[mgpu %d ranks] %-14s %s iters=%d/%d relerr_x=%.2e %s" % (
                world, name, "OK " if good else "FAIL", len(ls.obj), len(lo.obj), relerr(xs, xo),
                "" if good else str({k: v for k, v in checks.items() if not v})), flush=True)
    # fiber / slice application modes whose fibers stay inside a plane (SURVEY §8f-1 on slabs): per-fiber bounds along z
    # (indexed by the global plane), per-fiber cardinality along x on D_x, per-slice cardinality of the z slices
    for TF in (np.float32, np.float64):
        n, d = (20, 16, 13), (25.0, 25.0, 12.5)
        m = pr.synthetic_model(n, TF)
        lo_v, hi_v = np.linspace(1400.0, 1700.0, n[2]), np.linspace(3200.0, 4700.0, n[2])
        res = []
        for api in ([sip] if rank else [sip, orc]):
            cg = api.compgrid(d, n)
            cons = [api.set_definitions("bounds", "identity", lo_v.copy(), hi_v.copy(), ("fiber", "z")),
                    api.set_definitions("cardinality", "D_x", 0, 5, ("fiber", "x")),
                    api.set_definitions("cardinality", "D_y", 0, 90, ("slice", "z"))]
            opt = api.PARSDMM_options()
            opt.FL, opt.maxit = TF, 25
            P_sub, TD_OP, set_Prop = api.setup_constraints(cons, cg, TF)
            TD_OP, AtA, l0, y0 = api.PARSDMM_precompute_distribute(TD_OP, set_Prop, cg, opt)
            out = api.PARSDMM(m.copy(), AtA, TD_OP, set_Prop, P_sub, cg, opt)
            if api is sip:
                out = out[:3] + ([dd.gather_td(v, A) for v, A in zip(out[3], TD_OP)],)
            res.append(out)
        if rank == 0:
            (xs, ls, _, yg), (xo, lo, _, yy) = res
            checks = {"iters": len(ls.obj) == len(lo.obj), "cg_it": bool(np.array_equal(ls.cg_it, lo.cg_it)),
                      "x": relerr(xs, xo) < TOL[TF],
                      "support": bool(np.array_equal(yg[1] != 0, yy[1] != 0) and np.array_equal(yg[2] != 0, yy[2] != 0))}
            good = all(checks.values())
            ok = ok and good
            print("[mgpu %d ranks] %-14s %s iters=%d/%d relerr_x=%.2e %s" % (
                world, "fiber_" + np.dtype(TF).name, "OK " if good else "FAIL", len(ls.obj), len(lo.obj), relerr(xs, xo),
                "" if good else str({k: v for k, v in checks.items() if not v})), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
