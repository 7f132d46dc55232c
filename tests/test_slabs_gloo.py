"""World-size-2 (and 3) gloo tests of the host-side slab logic on CPU: partition ranges, per-rank CDS
rows, the (row-block, plane) ordering of transform-domain vectors, and the halo planes the device path
exchanges — emulated here with NumPy operators and torch.distributed send/recv."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sip_b200 as sip
    from sip_b200 import distributed as dd
    import problems as pr
    dd._state.update(active=True, rank=rank, world=world)      # host logic only: no NCCL communicator on CPU
    orc = pr.OracleAPI()
    TF = np.float64
    d = (2.0, 3.0, 5.0)
    k0, k1 = dd.slab_range(n[2])
    ranges = [dd.slab_range(n[2], r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n[2] and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    plane = n[0] * n[1]
    rng = np.random.default_rng(7)
    x = rng.standard_normal(plane * n[2])
    for kind in ("identity", "D_x", "D_y", "D_z", "TV"):
        op = sip.get_TD_operator(sip.compgrid(d, n), kind, TF)[0]
        A = orc.get_TD_operator(orc.compgrid(d, n), kind, TF)[0]
        # per-rank CDS rows == rows of the global CDS (integer work, bit exact)
        Rg, og = op.ata_cds()
        Rl, ol = op.ata_cds((k0, k1))
        assert np.array_equal(og, ol) and np.array_equal(Rl, Rg[plane * k0: plane * k1])
        # scatter / gather of transform-domain vectors reproduces the reference's global ordering
        v = rng.standard_normal(A.shape[0])
        vl = dd.scatter_td(v, op, k0, k1)
        assert np.array_equal(dd.gather_td(vl, op), v)
        # forward operator on a slab needs exactly one upper halo plane of x
        xl = x[plane * k0: plane * k1]
        halo_hi = np.zeros(plane)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(xl[:plane].copy()), rank - 1))
        if rank < world - 1:
            t = torch.zeros(plane, dtype=torch.float64)
            dist.recv(t, rank + 1)
            halo_hi = t.numpy()
        for r in reqs:
            r.wait()
        s_glob = np.asarray(A @ x).ravel()
        x_ext = np.concatenate([xl, halo_hi])
        # rows owned by this rank, evaluated from local planes + halo only
        got = []
        for (a, b), shp in zip(dd.local_td_slices(op, k0, k1), dd.block_shapes(op)):
            rows = np.arange(a, b)
            sub = A[rows, :]
            cols = sub.tocoo().col
            assert cols.size == 0 or (cols.min() >= plane * k0 and cols.max() < plane * (k1 + 1)), kind
            xz = np.zeros_like(x)
            hi = min(plane * (k1 + 1), x.size)
            xz[plane * k0: hi] = x_ext[: hi - plane * k0]
            got.append(np.asarray(sub @ xz).ravel())
        assert np.array_equal(np.concatenate(got) if got else np.zeros(0), dd.scatter_td(s_glob, op, k0, k1))
        # adjoint on a slab needs the previous rank's last row plane of a D_z block (lower halo)
        t_glob = np.asarray(A.T @ v).ravel()
        cols = np.arange(plane * k0, plane * k1)
        subT = A[:, cols]
        rws = subT.tocoo().row
        own = np.zeros(A.shape[0], dtype=bool)
        for a, b in dd.local_td_slices(op, k0, k1):
            own[a:b] = True
        foreign = np.setdiff1d(np.unique(rws), np.nonzero(own)[0])
        if kind in ("D_z", "TV") and rank > 0:
            zb = dd.local_td_slices(op, *ranges[rank - 1])[0]          # D_z block is the first block
            assert np.array_equal(foreign, np.arange(zb[1] - plane, zb[1])), kind
        else:
            assert foreign.size == 0, kind
        assert np.allclose(np.asarray(subT.T @ v).ravel(), t_glob[cols])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, (5, 4, 7)), (3, (4, 3, 8))])
def test_slab_host_logic_gloo(world, n):
    port = 29600 + world
    mp.spawn(_worker, args=(world, port, n), nprocs=world, join=True)
