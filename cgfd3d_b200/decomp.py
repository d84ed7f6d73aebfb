"""x-y domain decomposition of the hot path over ranks (one rank per GPU).

Mirrors the reference's topology and index bookkeeping:
  rank <-> (px, py) : MPI_Cart_create, row-major, y fastest, non-periodic (forward/mympi_t.c:32-49)
  block sizes       : gd_indx_set (forward/gd_t.c:2775-2849): points are dealt out evenly, the remainder
                      goes to the first ranks; z is never split (forward/gd_t.c:2851-2853)
  halo strips       : the library's own plan (cgfd_b200_halo_plan), see solver.halo_plan
"""
from __future__ import annotations

import numpy as np

from . import solver

NG = 3


def rank_coords(rank: int, px: int, py: int):
    return rank // py, rank % py


def neighbours(rank: int, px: int, py: int):
    """(x1, x2, y1, y2) ranks, -1 at a physical boundary."""
    ix, iy = rank_coords(rank, px, py)

    def rk(a, b):
        return a * py + b if (0 <= a < px and 0 <= b < py) else -1

    return rk(ix - 1, iy), rk(ix + 1, iy), rk(ix, iy - 1), rk(ix, iy + 1)


def split(n: int, parts: int, which: int):
    """(start, count) of block `which` when n points are dealt to `parts` blocks, remainder to the first ones."""
    base, rem = divmod(n, parts)
    cnt = base + (1 if which < rem else 0)
    start = which * base + min(which, rem)
    return start, cnt


def ref_split(n: int, parts: int, which: int, nlay1: int = 0, nlay2: int = 0):
    """(start, count) of block `which` along one axis exactly as gd_indx_set deals them out (forward/gd_t.c:2775-2849): the
    absorbing layers of the two physical faces count twice as load (n + nlay1 + nlay2 points are dealt evenly, the remainder to
    the first blocks), and the blocks that own a physical face give their layers back."""
    n_et = n + nlay1 + nlay2
    avg, left = divmod(n_et, parts)
    cnt = avg
    if which == 0:
        cnt -= nlay1
    if which == parts - 1:
        cnt -= nlay2
    if which < left:
        cnt += 1
    start = 0 if which == 0 else which * avg - nlay1
    if left:
        start += which if which < left else left
    return start, cnt


def local_block(rank, px, py, gni, gnj):
    ix, iy = rank_coords(rank, px, py)
    gi0, ni = split(gni, px, ix)
    gj0, nj = split(gnj, py, iy)
    return gi0, ni, gj0, nj


def exchange_host(w: np.ndarray, grid: dict, neigh, dirx: int, diry: int, sendrecv):
    """Halo exchange of a host array w[ncmp][nz][ny][nx] driven by the library's halo plan.
    sendrecv(peer, send_array, recv_shape) -> received array. Used by the gloo tests (and as documentation of
    what the GPU path does with pack kernel -> ncclSend/ncclRecv -> unpack kernel)."""
    for side in range(4):
        peer = neigh[side]
        if peer < 0:
            continue
        sb, rb = solver.halo_plan(grid, dirx, diry, side)
        si, sni, sj, snj, sk, snk = sb
        ri, rni, rj, rnj, rk_, rnk = rb
        out = np.ascontiguousarray(w[:, sk:sk + snk, sj:sj + snj, si:si + sni])
        got = sendrecv(peer, side, out, (w.shape[0], rnk, rnj, rni))
        w[:, rk_:rk_ + rnk, rj:rj + rnj, ri:ri + rni] = got
    return w
