#!/usr/bin/env python
"""N-rank vs 1-rank parity of the x-y decomposed hot path (launch with torchrun, one rank per GPU):
the same global problem is advanced nt steps (a) on rank 0 alone and (b) split over all ranks with the NCCL halo
exchange; the gathered wavefields must agree to round-off (the arithmetic per point is identical; only the
order in which a source shared by several... is not an issue here: one source, one owner)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgfd3d_b200 import decomp, hostsetup as hs, solver  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    px, py = {2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))
    if len(sys.argv) > 1 and sys.argv[1] == "y":
        px, py = py, px
    gni, gnj, nk, nt = 72, 60, 40, 48
    kw = dict(topo="hill", hill=(300.0, 800.0), pml_layers=6)
    G = hs.build_problem(gni, gnj, nk, **kw)
    gsi, gsj, gsk = gni // 2 - 1, gnj // 2 + 2, nk - 1 - 8
    gi0, ni, gj0, nj = decomp.local_block(rank, px, py, gni, gnj)
    nb = decomp.neighbours(rank, px, py)
    P = hs.build_problem(ni, nj, nk, sub=(gi0, gj0, gni, gnj, nb), dt=G.dt, **kw)
    for key in list(P.pml):   # same PML profiles as the global run (a per-rank L0 estimate would differ slightly)
        P.pml[key] = G.pml[key]
    if gi0 <= gsi < gi0 + ni and gj0 <= gsj < gj0 + nj:
        hs.make_source(P, gsi - gi0, gsj - gj0, gsk, nt_total=nt, spatial="point", fc=3.0, t0=0.3, stf_len=0.8)
    S = solver.Solver(P, device=local)
    uid = [solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    S.comm_init(uid[0], rank, world)
    S.run(nt)
    w = S.get_wavefield()
    phys = np.ascontiguousarray(w[:, 3:-3, 3:-3, 3:-3])
    out = [None] * world
    dist.all_gather_object(out, (gi0, ni, gj0, nj, phys))
    res = None
    if rank == 0:
        hs.make_source(G, gsi, gsj, gsk, nt_total=nt, spatial="point", fc=3.0, t0=0.3, stf_len=0.8)
        S1 = solver.Solver(G, device=local)
        S1.run(nt)
        w1 = S1.get_wavefield()[:, 3:-3, 3:-3, 3:-3]
        full = np.zeros_like(w1)
        for (a, n1, b, n2, ph) in out:
            full[:, :, b:b + n2, a:a + n1] = ph
        errs = []
        for c in range(9):
            d = float(np.abs(full[c] - w1[c]).max()); m = float(np.abs(w1[c]).max())
            errs.append(d / m if m > 0 else d)
        res = {"world": world, "grid": "%dx%d" % (px, py), "max_rel_err": max(errs), "amp": float(np.abs(w1[2]).max()), "ok": bool(max(errs) <= 1e-5)}
        print("MULTI_GPU_CHECK " + json.dumps(res), flush=True)
    S.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not res["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
