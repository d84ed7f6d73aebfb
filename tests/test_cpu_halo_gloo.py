"""world_size-2 (3, 4 and 8) gloo runs of the x-y halo exchange on CPU: the strips come from the library's own halo plan
(cgfd_b200_halo_plan, the code the NCCL path uses), the transport is torch.distributed gloo send/recv.
After the exchange every ghost strip an operator needs must equal the neighbour's physical values."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgfd3d_b200 import decomp

GNI, GNJ, NK, NC = 22, 17, 6, 3


def _field(gi, gj, k, c):
    return (c * 1000003 + k * 10007 + gj * 101 + gi).astype(np.float32)


def _worker(rank, world, px, py, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gi0, ni, gj0, nj = decomp.local_block(rank, px, py, GNI, GNJ)
    neigh = decomp.neighbours(rank, px, py)
    nx, ny, nz = ni + 6, nj + 6, NK + 6
    grid = dict(nx=nx, ny=ny, nz=nz, ni1=3, ni2=3 + ni - 1, nj1=3, nj2=3 + nj - 1, nk1=3, nk2=3 + NK - 1)
    I = np.arange(nx)[None, None, None, :] - 3 + gi0
    J = np.arange(ny)[None, None, :, None] - 3 + gj0
    K = np.arange(nz)[None, :, None, None]
    Cc = np.arange(NC)[:, None, None, None]
    truth = _field(I, J, K, Cc) + np.zeros((NC, nz, ny, nx), np.float32)
    errors = []
    for dirx in (0, 1):
        for diry in (0, 1):
            w = np.zeros_like(truth)
            w[:, 3:3 + NK, 3:3 + nj, 3:3 + ni] = truth[:, 3:3 + NK, 3:3 + nj, 3:3 + ni]

            def sendrecv(peer, side, out, shape):
                got = torch.empty(shape, dtype=torch.float32)
                # lower rank sends first: a deadlock-free order for blocking gloo point-to-point
                if rank < peer:
                    dist.send(torch.from_numpy(out), peer)
                    dist.recv(got, peer)
                else:
                    dist.recv(got, peer)
                    dist.send(torch.from_numpy(out), peer)
                return got.numpy()

            decomp.exchange_host(w, grid, neigh, dirx, diry, sendrecv)
            lx, rx = (3, 1) if dirx else (1, 3)
            ly, ry = (3, 1) if diry else (1, 3)
            checks = []
            if neigh[0] >= 0:
                checks.append((slice(3, 3 + NK), slice(3, 3 + nj), slice(3 - lx, 3)))
            if neigh[1] >= 0:
                checks.append((slice(3, 3 + NK), slice(3, 3 + nj), slice(3 + ni, 3 + ni + rx)))
            if neigh[2] >= 0:
                checks.append((slice(3, 3 + NK), slice(3 - ly, 3), slice(3, 3 + ni)))
            if neigh[3] >= 0:
                checks.append((slice(3, 3 + NK), slice(3 + nj, 3 + nj + ry), slice(3, 3 + ni)))
            for ck in checks:
                sl = (slice(None),) + ck
                if not np.array_equal(w[sl], truth[sl]):
                    errors.append((rank, dirx, diry, ck))
            # nothing but the planned strips may have been touched
            mask = np.ones_like(w, bool)
            mask[:, 3:3 + NK, 3:3 + nj, 3:3 + ni] = False
            for ck in checks:
                mask[(slice(None),) + ck] = False
            if np.any(w[mask] != 0):
                errors.append((rank, dirx, diry, "wrote outside the plan"))
    q.put((rank, errors, len(neigh) - list(neigh).count(-1)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("px,py", [(2, 1), (1, 2), (2, 2), (3, 1), (4, 2)])   # (4, 2) = the 8-GPU process grid: ranks with neighbours on both x sides
def test_halo_exchange_gloo(px, py):
    world = px * py
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, px, py, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, errors, nn in res:
        assert not errors, errors
        assert nn >= 1


def test_split_matches_reference_rule():
    # remainder points go to the first blocks (forward/gd_t.c:2775-2849)
    assert [decomp.split(10, 3, w) for w in range(3)] == [(0, 4), (4, 3), (7, 3)]
    assert decomp.neighbours(0, 2, 2) == (-1, 2, -1, 1)
    assert decomp.neighbours(3, 2, 2) == (1, -1, 2, -1)
