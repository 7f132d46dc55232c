#!/usr/bin/env python
"""Soak test of the peer-memory protocol (mailbox all-reduces, neighbour-plane loads, small peer all-reduces, device-side
CG loops) on slabs: many back-to-back solves of small, UNEVEN slab problems; every solve must reproduce the first one bit
for bit and no bounded spin may time out.  Launch with torchrun:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/soak_mgpu.py --solves 1000"""
import argparse
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import problems as pr  # noqa: E402
import sip_b200 as sip  # noqa: E402
from sip_b200 import distributed as dd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--solves", type=int, default=1000)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("cpu:gloo,cuda:nccl")
dd.init(rank, world, local)
cases = [("config2", pr.spec_config2((32, 24, 2 * world + 3), np.float32), 25),      # uneven: 2-3 planes per rank
         ("config3", pr.spec_config3((24, 20, 3 * world + 1), np.float32), 20)]
t0 = time.perf_counter()
total = 0
for name, spec, maxit in cases:
    opt = sip.PARSDMM_options()
    opt.maxit, opt.evol_rel_tol = maxit, 10 * float(np.finfo(np.float32).eps)
    sb = pr.build(sip, copy.deepcopy(spec), opt)
    ref = None
    for k in range(args.solves // len(cases)):
        x, log, l, y = sip.PARSDMM(spec["m"].copy(), sb["AtA"], sb["TD_OP"], sb["set_Prop"], sb["P_sub"], sb["cg"], sb["opt"],
                                   gather_result=False)
        sig = (x.tobytes(), tuple(log.cg_it), tuple(log.obj))
        if ref is None:
            ref = sig
        elif sig != ref:
            raise SystemExit("[soak rank %d] %s: solve %d differs from the first solve" % (rank, name, k))
        total += 1
    if rank == 0:
        print("[soak %d ranks] %-8s %d solves identical (iterations %d, cg %d)" % (world, name, args.solves // len(cases), len(log.obj),
                                                                                 int(sum(log.cg_it))), flush=True)
dist.barrier()
if rank == 0:
    print("[soak %d ranks] %d solves, peer path %s, %.1f s, no time-outs" % (world, total, dd.peer_path(), time.perf_counter() - t0), flush=True)
dist.destroy_process_group()
