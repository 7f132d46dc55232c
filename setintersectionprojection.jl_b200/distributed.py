"""Multi-GPU slabs: one process per GPU, z-slab domain decomposition.

Replaces the reference's own parallelism (one Julia worker per constraint set over Distributed /
DistributedArrays: update_y_l_parallel.jl, adapt_rho_gamma_parallel.jl) by slabs of the 3-D volume along
its slowest axis.  torch.distributed is only the plumbing that carries the 128-byte NCCL unique id to
every rank and gathers results on the host; halo planes and scalar reductions move inside the C library
(ncclSend/ncclRecv of contiguous planes, Float64 ncclAllReduce) on the solver's stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

_state = {"active": False, "rank": 0, "world": 1}


def init(rank: int | None = None, world: int | None = None, local_rank: int | None = None) -> None:
    """Create the library's NCCL communicator.  torch.distributed must be initialised (any backend)."""
    import torch.distributed as dist          # import torch first: its bundled libnccl is then the one in use
    if not dist.is_initialized():
        raise RuntimeError("call torch.distributed.init_process_group(...) first (it carries the NCCL unique id)")
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
    lib = _lib.load()
    ctx = _lib.ctx(local_rank)
    if world > 1:
        uid = (C.c_char * 128)()
        if rank == 0:
            _lib.check(lib.sipb_comm_unique_id(C.cast(uid, C.c_void_p)))
        box = [bytes(uid.raw) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        buf = C.create_string_buffer(box[0], 128)
        _lib.check(lib.sipb_comm_init(ctx, rank, world, C.cast(buf, C.c_void_p)))
    _state.update(active=world > 1, rank=rank, world=world, device=local_rank)


def active() -> bool:
    return bool(_state["active"])


def peer_path() -> bool:
    """True when the CG uses the peer-memory (CUDA IPC / NVLink) collectives instead of NCCL."""
    if not active():
        return False
    flag = C.c_int(0)
    _lib.check(_lib.load().sipb_comm_peer_path(_lib.ctx(_state["device"]), C.byref(flag)))
    return bool(flag.value)


def rank() -> int:
    return int(_state["rank"])


def world() -> int:
    return int(_state["world"])


def slab_range(n_last: int, rank_: int | None = None, world_: int | None = None):
    """Planes [k0,k1) of the slowest axis owned by a rank (same rule as sipb_slab_range)."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    return (n_last * r) // w, (n_last * (r + 1)) // w


def block_shapes(op):
    """[(shape of row block b, index of its slowest axis)] of a 3-D operator, in row order."""
    n = op.n
    if op.kind == "identity":
        return [tuple(n)]
    return [tuple(v - 1 if a == axis else v for a, v in enumerate(n)) for axis in op._axes()]


def local_td_slices(op, k0: int, k1: int):
    """For each row block: (global start, global stop) of the rows owned by planes [k0,k1): block rows are
    stored plane by plane, so a slab is one contiguous range per block (a D_z block owns the planes
    [k0, min(k1, n3-1)))."""
    out, start = [], 0
    for shp in block_shapes(op):
        plane = shp[0] * shp[1]
        hi = min(k1, shp[2])
        out.append((start + plane * k0, start + plane * max(hi, k0)))
        start += plane * shp[2]
    return out


def scatter_model(m: np.ndarray, n, k0: int, k1: int) -> np.ndarray:
    plane = int(n[0]) * int(n[1])
    return np.ascontiguousarray(m[plane * k0: plane * k1])


def scatter_td(v: np.ndarray, op, k0: int, k1: int) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([v[a:b] for a, b in local_td_slices(op, k0, k1)]))


def _all_gather(arr: np.ndarray):
    import torch.distributed as dist
    out = [None] * world()
    dist.all_gather_object(out, arr)
    return out


def gather_model(x_local: np.ndarray) -> np.ndarray:
    """Host-side gather of the slabs of a model-sized vector (planes are rank ordered)."""
    return np.concatenate(_all_gather(x_local))


def gather_td(v_local: np.ndarray, op) -> np.ndarray:
    """Host-side gather of a transform-domain vector into the reference's global (row-block major) order."""
    parts = _all_gather(v_local)
    n_last = op.n[2]
    total = op.rows
    out = np.empty(total, dtype=v_local.dtype)
    for r, part in enumerate(parts):
        k0, k1 = slab_range(n_last, r, world())
        pos = 0
        for a, b in local_td_slices(op, k0, k1):
            out[a:b] = part[pos: pos + (b - a)]
            pos += b - a
    return out
