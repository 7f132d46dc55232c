#!/usr/bin/env python
"""bench.py — PARSDMM iterations/s on the 3-D intersection-projection workload of BASELINE.json.

A "step" is one complete PARSDMM projection (initial feasibility check, Q assembly, iterations until the
reference's stopping rules fire) of the configs[1] workload: 3-D 200^3 Float32, bounds ∩ anisotropic-TV
l1 ball ∩ lateral slope bounds (examples/test_scaling_3D.jl-style).  value = PARSDMM iterations / second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 200] [--impl reference]

* device arm (default): `value` with the problem and m resident in HBM (CUDA-event time of the solves),
  `e2e` through the public API `sip_b200.PARSDMM(m, ...)` with host buffers (H2D of m, D2H of x and log
  inside the timed region), the roofline of the dominant kernel class (largest share of the solve; its
  algorithmic bytes are counted by the library at every launch) from a CUDA-event kernel table, and a
  bounded CPU baseline (the NumPy oracle, rank 0, N=1 only).
* --impl reference: the reference's CPU algorithm (oracle port; Julia is not installable here) on the
  host cores, same workload/metric, each step a bounded sample (a few PARSDMM iterations).

Multi-GPU (N>1, torchrun): one process per GPU, the volume is slab-partitioned along its slowest axis
(NCCL halo planes + Float64 all-reduces inside the C library).  Default is weak scaling: the grid is
n x n x (n*N), i.e. one n^3 slab per GPU, and `value` counts slab-iterations per second (N x the PARSDMM
iterations/s of the N-times larger problem; equal to iterations/s at N=1).  `--scaling strong` keeps the
n^3 grid fixed.  Time = max over ranks of the CUDA-event time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def env_int(name, default):
    return int(os.environ.get(name, default))


WORKLOAD = {"name": "config2"}


def workload(n, TF=np.float32, nz=None):
    import problems as pr
    if WORKLOAD["name"] == "config3":      # BASELINE configs[2]: bounds ∩ TV-l1 ∩ cardinality(TV), k = 5 % of the rows
        return pr.spec_config3((n, n, nz or n), TF)
    return pr.spec_config2((n, n, nz or n), TF)


def tweak_options(o):
    o.evol_rel_tol = 10 * float(np.finfo(np.float32).eps)      # examples/test_scaling_3D.jl:25
    o.maxit = 200
    return o


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = "/tmp/sipb_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                t = [v.strip() for v in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1]))
                    mx.append(float(t[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out["sm_mhz"] = float(np.median(busy))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference algorithm)
# --------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


_CPU_MODE = {"threaded": None}


def cpu_threaded_available():
    """The C/OpenMP baseline needs gcc (or the prebuilt oracle/_cbuild/libsipref.so); without it the CPU legs fall
    back to the plain NumPy oracle and say so."""
    if _CPU_MODE["threaded"] is None:
        try:
            from oracle import cpu_baseline as cb
            cb.use_all_cores()           # torchrun exports OMP_NUM_THREADS=1; the baseline gets every host core
            _CPU_MODE["threaded"] = True
        except Exception as e:       # noqa: BLE001
            print("bench.py: threaded CPU baseline unavailable (%s); timing the NumPy oracle instead" % e, file=sys.stderr)
            _CPU_MODE["threaded"] = False
    return _CPU_MODE["threaded"]


def cpu_threads():
    if not cpu_threaded_available():
        return 1
    from oracle import cpu_baseline as cb
    return cb.threads()


def cpu_impl():
    return CPU_IMPL if cpu_threaded_available() else "NumPy/SciPy oracle (single-threaded apart from BLAS dot/norm)"


CPU_IMPL = ("threaded C/OpenMP restatement of the reference's vector phases (oracle/c/ref_kernels.c: per-diagonal CDS "
            "passes, BLAS-1 style CG passes, sort-based l1 projection) driven by the NumPy oracle's control flow")


def cpu_sample(n, iters):
    """Run `iters` PARSDMM iterations of the n^3 workload with the threaded CPU baseline (oracle/cpu_baseline.py:
    the oracle's control flow and scalar rules, vector phases in C/OpenMP on all host cores); returns (its/s over
    the iteration phases, seconds of the iteration phases, setup + initialization seconds, iterations done).
    The operator set-up is cached between calls (it is outside the PARSDMM call in the reference too)."""
    import problems as pr
    orc = pr.OracleAPI()
    t0 = time.perf_counter()
    if n not in _CPU_CACHE:
        spec = workload(n)
        opt = tweak_options(orc.PARSDMM_options())
        _CPU_CACHE[n] = (spec, pr.build(orc, spec, opt))
    spec, ob = _CPU_CACHE[n]
    t_setup = time.perf_counter() - t0
    if cpu_threaded_available():
        from oracle import cpu_baseline as cb
        x, log, _, _ = cb.PARSDMM(spec["m"].copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"],
                                  constraint=ob["cons"], max_iterations=iters if iters > 0 else None)
    else:
        ob["opt"].maxit = min(iters, 4) if iters > 0 else 4
        x, log, _, _ = orc.PARSDMM(spec["m"].copy(), ob["AtA"], ob["TD_OP"], ob["set_Prop"], ob["P_sub"], ob["cg"], ob["opt"])
    t_iter = sum(v for k, v in log.timing.items() if k != "initialization")
    done = len(log.obj)
    return done / t_iter, t_iter, t_setup + log.timing.get("initialization", 0.0), done


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU algorithm (NumPy/SciPy oracle port — Julia cannot be installed
    here) on the host cores.  Every step is a bounded sample: the first `--cpu-iters` PARSDMM iterations
    of one n^3 slab of the workload (for --gpus N the device arm's grid is N such slabs; the CPU processes
    one slab at a time, so its slab-iterations/s do not depend on N)."""
    if rank != 0:
        return
    n = args.n
    cores = cpu_threads()
    iters = args.cpu_iters
    secs, its = [], []
    t_start = time.perf_counter()
    budget_s = float(os.environ.get("SIPB_REFERENCE_BUDGET_S", "280"))      # keep the whole arm within a few minutes
    warm = min(args.warmup, 1)       # the CPU needs no more than one warm-up (page faults / caches of the set-up)
    for s in range(warm + args.steps):
        v, t_iter, t_setup, done = cpu_sample(n, iters)
        if s >= warm:
            secs.append(t_iter)
            its.append(done)
            if time.perf_counter() - t_start > budget_s:
                break
    value = float(sum(its) / sum(secs))
    nz = n * args.gpus if args.scaling == "weak" else n
    sample = ("%s of one %d^3 Float32 slab per step; %s; the rate counts the iteration phases only — operator set-up "
              "and PARSDMM_initialize are excluded, which favours the CPU"
              % ("one full projection (%d PARSDMM iterations)" % its[-1] if iters <= 0 else "first %d PARSDMM iterations" % iters,
                 n, cpu_impl()))
    line = {
        "impl": "reference", "metric": "parsdmm_iterations_per_second", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": len(secs), "warmup": warm, "steps_requested": args.steps,
        "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": args.scaling if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "3D %dx%dx%d Float32 bounds ∩ anisotropic TV l1 ∩ D_x,D_y slope bounds (BASELINE configs[1], "
                               "test_scaling_3D-style)" % (n, n, nz), "grid": [n, n, nz],
                   "note": "CPU restatement of the reference algorithm on %d OpenMP threads; the Julia reference cannot "
                           "be installed in this image (no Julia, no network)" % cores},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# device arm
# --------------------------------------------------------------------------------------------------
def run_device(args, rank, world, local_rank):
    import torch
    import sip_b200 as sip
    import problems as pr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    nz = n * world if (world > 1 and args.scaling == "weak") else n
    N = n * n * nz
    units = world if (world > 1 and args.scaling == "weak") else 1       # slabs of n^3 per PARSDMM iteration
    if world > 1:
        from sip_b200 import distributed as dd
        dd.init(rank, world, local_rank)
    spec = workload(n, nz=nz)
    opt = tweak_options(sip.PARSDMM_options())
    sb = pr.build(sip, spec, opt)
    m = spec["m"]
    call = lambda **kw: sip.PARSDMM(m, sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"],   # noqa: E731
                                    return_ly=False, gather_result=False, **kw)
    # first call uploads the operators (the "distribute" of PARSDMM_precompute_distribute)
    x, log, _, _ = call()
    iters_per_step = len(log.obj)
    for _ in range(max(args.warmup - 1, 0)):
        call(resident_io=True)

    # ---- timed: resident (value) ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_s, its, launches = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, lg, _, _ = call(resident_io=True)
        dev_s += lg.timing["device_seconds"]
        its += len(lg.obj)
        launches += lg.timing["total_launches"]
    barrier()
    wall_resident = time.perf_counter() - t0
    # ---- timed: end to end through the public API with host buffers -------------------------------
    # inputs come from pinned host memory and the result lands in pinned host memory (this rank's slab when N > 1)
    dev = sb["AtA"]._device
    slab = getattr(dev, "slab", None)
    m_loc = m if slab is None else m[n * n * slab[0]: n * n * slab[1]]
    m_pin = torch.empty(m_loc.size, dtype=torch.float32).pin_memory().numpy()
    m_pin[:] = m_loc
    x_pin = torch.empty(dev.N, dtype=torch.float32).pin_memory().numpy()
    call_e2e = lambda: sip.PARSDMM(m_pin, sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"],   # noqa: E731
                                   x=x_pin, return_ly=False, gather_result=False)
    call_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_its, h2d, d2h = 0, 0, 0
    for _ in range(args.steps):
        xk, lg, _, _ = call_e2e()
        e2e_its += len(lg.obj)
        h2d += lg.timing["h2d_bytes"]
        d2h += lg.timing["d2h_bytes"]
    barrier()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop()

    # max over ranks (device time for `value`, wall for e2e)
    tm = torch.tensor([dev_s, wall_e2e, wall_resident], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([its, e2e_its, launches], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_s_max, wall_e2e_max, wall_res_max = [float(v) for v in tm.tolist()]
    its_all, e2e_its_all, launches_all = [float(v) for v in cnt.tolist()]
    # every rank counted the same global iterations: work units = iterations x slabs
    its_all, e2e_its_all = its_all / world * units, e2e_its_all / world * units

    # ---- roofline of the dominant kernel from one profiled solve (CUDA events around every launch) --
    roof = None
    kernels = {}
    _, lgp, _, _ = call(resident_io=True, profile_kernels=True)      # every rank: the solve contains collectives
    if rank == 0:
        kernels = lgp.timing["kernels"]
        kbytes = lgp.timing["kernel_bytes"]         # algorithmic bytes per class, counted by the library at launch
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        else:
            peak, which = 6650.0, "fallback"
        tot_ms = sum(v[1] for v in kernels.values())
        streaming = {k: v for k, v in kernels.items() if kbytes.get(k) and v[1] > 0}
        if streaming:
            top = max(streaming, key=lambda k: streaming[k][1])       # the class with the largest share of the solve
            cnt_k, ms_k = streaming[top]
            alg = kbytes[top] / cnt_k
            avg_ms = ms_k / cnt_k
            ach = alg / (avg_ms * 1e-3) / 1e9
            traffic = None       # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
            tpath = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
            if os.path.exists(tpath) and world == 1 and args.workload == "config2":
                tj = json.load(open(tpath))
                if tj.get("grid") == [n, n, nz] and top in tj.get("kernels", {}):
                    traffic = tj["kernels"][top]["dram_bytes_per_launch"]
            roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "peak_source": which,
                    "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg,
                    "avg_launch_ms": avg_ms, "launches": cnt_k, "share_of_kernel_time": ms_k / tot_ms if tot_ms else None,
                    "q_form": sb["AtA"]._device.q_form,
                    "per_kernel": {k: {"launches": v[0], "ms": round(v[1], 3), "gbs": round(kbytes[k] / (v[1] * 1e-3) / 1e9, 1),
                                       "frac": round(kbytes[k] / (v[1] * 1e-3) / 1e9 / peak, 3)} for k, v in streaming.items()},
                    "whole_solve": {"algorithmic_gb": round(sum(kbytes.values()) / 1e9, 3), "kernel_ms": round(tot_ms, 3),
                                    "gbs": round(sum(kbytes.values()) / (tot_ms * 1e-3) / 1e9, 1),
                                    "frac": round(sum(kbytes.values()) / (tot_ms * 1e-3) / 1e9 / peak, 3)}}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # BASELINE.json's second metric, "CDS-SpMV HBM GB/s": the compressed-diagonal SpMV + dot kernel of cg.jl on the
    # CDS ARRAYS of this grid (nd = 7 diagonals), timed alone back to back (sipb_bench_spmv, CUDA events) — the
    # solve itself no longer streams the matrix when Q is held as stencil-class tables (roofline.q_form).
    cds = None
    if world == 1 and roof is not None:
        try:
            import ctypes as C
            L = sip._lib
            n3 = (C.c_int64 * 3)(n, n, nz)
            ms_, nb_ = C.c_double(0.0), C.c_int64(0)
            L.check(L.load().sipb_bench_spmv(L.ctx(), 0, 3, n3, 3, 20, 0, C.byref(ms_), C.byref(nb_)))
            if ms_.value > 0:
                gbs = nb_.value / (ms_.value * 1e-3) / 1e9
                cds = {"form": "arrays", "avg_launch_ms": ms_.value, "algorithmic_bytes_per_launch": nb_.value,
                       "gbs": gbs, "frac": gbs / roof["peak"], "launches": 20}
        except Exception as e:       # noqa: BLE001 - an auxiliary figure must never take the bench line down
            print("bench.py: CDS SpMV unit timing skipped (%s)" % e, file=sys.stderr)
        roof["cds_spmv_arrays"] = cds

    cpu = None
    if world == 1 and not args.no_cpu:
        v, t_iter, t_setup, done = cpu_sample(n, args.cpu_iters)
        cpu = {"value": v, "unit": "iterations/s", "cores": cpu_threads(), "kind": "port",
               "sample": "%s of the same %d^3 workload (%d PARSDMM iterations, %.1f s of iteration phases; %.1f s of "
                         "set-up and initialization excluded); %s"
                         % ("one full projection" if args.cpu_iters <= 0 else "the first iterations", n, done, t_iter, t_setup,
                            cpu_impl())}

    value = its_all / dev_s_max
    line = {
        "metric": "parsdmm_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s_max / args.steps,
        "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": ("3D %dx%dx%d Float32 bounds ∩ anisotropic TV l1 ∩ D_x,D_y slope bounds (BASELINE configs[1], "
                                "test_scaling_3D-style)" if args.workload == "config2" else
                                "3D %dx%dx%d Float32 bounds ∩ TV l1 ∩ cardinality of the discrete gradient (BASELINE configs[2])")
                               % (n, n, nz), "grid": [n, n, nz],
                   "value_units": "PARSDMM iterations/s" if units == 1 else
                                  "slab-iterations/s = %d x PARSDMM iterations/s (one %d^3 slab per GPU)" % (units, n),
                   "step": "one full PARSDMM projection to the reference's stopping rules",
                   "parsdmm_iterations_per_step": iters_per_step, "time_to_tolerance_ms": 1e3 * dev_s_max / args.steps,
                   "cache": "working set %.1f GB (x-side vectors + 8 vectors per set) >> 126 MB L2, no flush needed" % (N * 4 * 60 / 1e9),
                   "parallelism": "single GPU" if world == 1 else
                                  "z-slabs over %d GPUs (%s), %s scaling" % (
                                      world, "peer-memory CG reductions + neighbour-plane loads over NVLink (CUDA IPC), NCCL for the "
                                      "per-iteration halos / batched all-reduce" if dd.peer_path() else
                                      "NCCL send/recv halo planes + Float64 all-reduce", args.scaling)},
        "e2e": {"value": e2e_its_all / wall_e2e_max, "unit": "iterations/s", "h2d_bytes_per_step": h2d // args.steps,
                "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": 1e3 * wall_e2e_max / args.steps},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernel_table_ms": {k: [v[0], round(v[1], 3)] for k, v in kernels.items()},
        "resident_wall_ms_per_step": 1e3 * wall_res_max / args.steps,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="device", choices=["device", "reference"])
    ap.add_argument("--size", "--n", dest="n", type=int, default=200,
                    help="grid width (BASELINE configs[1] uses 200); use --size under torchrun (its parser claims --n)")
    ap.add_argument("--cpu-iters", type=int, default=0,
                    help="PARSDMM iterations in the bounded CPU sample (0 = one full projection to the stopping rules)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 = BASELINE configs[1] (default, the bench line); config3 = configs[2] (TV cardinality)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = n x n x (n*N) grid (one n^3 slab per GPU), strong = fixed n^3 grid")
    args = ap.parse_args()
    WORKLOAD["name"] = args.workload
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_device(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
