"""N-rank vs 1-rank value check of the x-y decomposed hot path, on the process grid a multi-GPU run really uses.

The same global problem is advanced `nt` steps (a) split over all ranks with the NCCL halo exchange and (b) on rank 0 alone;
the gathered wavefields must agree to round-off (the arithmetic per point is identical: only which kernel launch -- boundary
phase or interior phase -- computes a point differs). Called by bench.py before the timed region of every multi-GPU run
(`"parity_nrank"` in its JSON line), by scripts/multi_gpu_check.py and tests/test_gpu_multi.py.

Every rank's block is at least three 32 x 8 tiles wide in both directions, so that boundary-phase tiles, interior tiles, the
free-surface rows, CFS-PML faces and ranks with neighbours on both sides (4 x 2 grid) all occur; the source sits next to an
inter-rank face (its footprint is injected by the boundary phase of one rank and reaches the neighbour through the exchange).
"""
from __future__ import annotations

import numpy as np

from . import decomp, hostsetup as hs, solver


def check(rank: int, world: int, local: int, px: int, py: int, dist, medium: str = "iso", nt: int = 40, tol: float = 1e-5,
          block=(96, 40, 36)):
    """dist = an initialised torch.distributed module (any backend that can all_gather_object / broadcast_object_list).
    Returns the result dict on rank 0, None elsewhere."""
    bi, bj, nk = block
    gni, gnj = bi * px, bj * py
    kw = dict(topo="hill", hill=(300.0, 0.12 * max(gni, gnj) * 100.0), pml_layers=6, medium=medium, seed=3 if medium != "iso" else None,
              free_top=(medium == "iso"))
    if medium != "iso":
        kw["pml_faces"] = ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1))
    G = hs.build_problem(gni, gnj, nk, dt=0.008, **kw)
    # source one point left of the first inter-rank face in x (or in y on a 1 x N grid), a few rows below the surface
    gsi = bi - 2 if px > 1 else gni // 2
    gsj = bj - 1 if (py > 1 and px == 1) else gnj // 2 + 1
    gsk = nk - 1 - 7
    gi0, ni, gj0, nj = decomp.local_block(rank, px, py, gni, gnj)
    nb = decomp.neighbours(rank, px, py)
    P = hs.build_problem(ni, nj, nk, sub=(gi0, gj0, gni, gnj, nb), dt=G.dt, **kw)
    if medium != "iso":
        # the random perturbation of the synthetic media must be the global field's, not a per-rank one
        sl = (slice(None), slice(gj0, gj0 + nj + 6), slice(gi0, gi0 + ni + 6))
        P.media = [np.ascontiguousarray(a[sl]) for a in G.media]
    for key in list(P.pml):   # same PML profiles as the global run (a per-rank slab-length estimate differs slightly)
        P.pml[key] = G.pml[key]
    src_kw = dict(nt_total=nt, spatial="gauss", fc=3.0, t0=0.25, stf_len=0.6, mech=(1e16, 0.7e16, 1.2e16, 0.3e16, -0.2e16, 0.1e16))
    # a Gaussian footprint is 7 points wide: every rank it reaches gets its share (the reference does the same per rank,
    # forward/src_t.c:482-496, with the footprint clipped to the rank's own points)
    if gi0 - 3 <= gsi < gi0 + ni + 3 and gj0 - 3 <= gsj < gj0 + nj + 3:
        hs.make_source(P, gsi - gi0, gsj - gj0, gsk, **src_kw)
    S = solver.Solver(P, device=local)
    uid = [solver.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    S.comm_init(uid[0], rank, world)
    S.run(nt)
    w = S.get_wavefield()
    S.close()
    phys = np.ascontiguousarray(w[:9, 3:-3, 3:-3, 3:-3])
    out = [None] * world
    dist.all_gather_object(out, (gi0, ni, gj0, nj, phys))
    if rank != 0:
        return None
    hs.make_source(G, gsi, gsj, gsk, **src_kw)
    S1 = solver.Solver(G, device=local)
    S1.run(nt)
    w1 = S1.get_wavefield()[:9, 3:-3, 3:-3, 3:-3]
    S1.close()
    full = np.zeros_like(w1)
    for (a, n1, b, n2, ph) in out:
        full[:, :, b:b + n2, a:a + n1] = ph
    errs = []
    for c in range(9):
        d = float(np.abs(full[c] - w1[c]).max())
        m = float(np.abs(w1[c]).max())
        errs.append(d / m if m > 0 else d)
    return {"grid": "%dx%d" % (px, py), "medium": medium, "global_size": "%dx%dx%d" % (gni, gnj, nk), "steps": nt,
            "max_rel_err": max(errs), "amp_vz": float(np.abs(w1[2]).max()), "tol": tol, "ok": bool(max(errs) <= tol and np.isfinite(full).all())}
