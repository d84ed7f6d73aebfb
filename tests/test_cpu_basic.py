"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the ctypes mirror
matches the compiled struct, the compute path refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from cgfd3d_b200 import abi, hostsetup as hs, solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = solver.load_library()
    header = open(os.path.join(ROOT, "include", "cgfd3d_b200.h")).read()
    declared = set(re.findall(r"\b(cgfd_b200_[a-z0-9_]+)\s*\(", header))
    declared.discard("cgfd_b200_ctx")
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), "library does not export " + name
    assert set(solver.SYMBOLS) <= declared


def test_struct_layout_matches_library():
    L = solver.load_library()
    assert L.cgfd_b200_abi_version() == abi.ABI_VERSION
    assert L.cgfd_b200_sizeof_problem() == ctypes.sizeof(abi.Problem)


def test_no_cpu_fallback():
    if solver.device_count() > 0:
        pytest.skip("a GPU is present")
    prob = hs.build_problem(12, 12, 12, pml_layers=2)
    with pytest.raises(solver.CgfdError, match="no CUDA device"):
        solver.Solver(prob)


def test_fd_tables():
    fd = hs.fd_macdrp()
    # every stage flips all three axes (forward/fd_t.c:233-236)
    for p in range(8):
        for s in range(3):
            for a in range(3):
                assert fd.dir[p][s][a] + fd.dir[p][s + 1][a] == 1
    assert abs(sum(fd.coef[0])) < 1e-3 and abs(sum(fd.coef[1])) < 1e-3  # the reference rounds its coefficients
    assert [fd.indx[0][n] for n in range(5)] == [-1, 0, 1, 2, 3]
    assert [fd.indx[1][n] for n in range(5)] == [-3, -2, -1, 0, 1]


def test_device_setup_matches_host_setup():
    """devsetup (torch expressions, run on the GPU for the big bench blocks) == hostsetup (numpy), here both on the CPU"""
    import numpy as np
    from cgfd3d_b200 import devsetup as ds, hostsetup as hs
    # an edge block with one neighbour, and a block of the 4x2 grid with neighbours on both x sides and one y side
    for sub in ((40, 0, 80, 36, (0, -1, -1, -1)), (40, 36, 160, 72, (0, 4, 1, -1))):
        a = hs.build_problem(40, 36, 30, topo="hill", hill=(300.0, 800.0), pml_layers=6, dt=0.01, sub=sub)
        b = ds.build_problem(40, 36, 30, device="cpu", hill=(300.0, 800.0), pml_layers=6, dt=0.01, sub=sub)
        for m in range(10):
            np.testing.assert_allclose(b.metric[m].numpy(), a.metric[m], rtol=1e-6, atol=0)
        assert set(a.pml) == set(b.pml)
        for k in a.pml:
            for q in (1, 2, 3):
                np.testing.assert_array_equal(a.pml[k][q], b.pml[k][q])
        for n in ("matVx2Vz", "matVy2Vz", "matF2Vz"):
            np.testing.assert_allclose(b.mats[n], a.mats[n], rtol=1e-6, atol=1e-30)


def test_launch_plan_longest_job_first():
    """the block order of the interior kernel (pure host logic of the library): a permutation of tiles x z-chunks in which every
    tile that meets an x / y PML slab comes before every tile that does not; chunks cover the rows below the free-surface rows"""
    from cgfd3d_b200 import solver
    ni, nj, nk, nl = 400, 400, 200, 10
    grid = dict(nx=ni + 6, ny=nj + 6, nz=nk + 6, ni1=3, ni2=ni + 2, nj1=3, nj2=nj + 2, nk1=3, nk2=nk + 2)
    ntx, nty = (ni + 31) // 32, (nj + 7) // 8
    zchunk, order = solver.launch_plan(grid, ((nl, nl), (nl, nl), (nl, 0)), 1, (0, ntx, 0, nty))
    nzc = -(-(nk - 4) // zchunk)
    assert 16 <= zchunk <= 49 and len(order) == ntx * nty * nzc
    assert sorted(order) == list(range(ntx * nty * nzc))

    def is_pml(b):
        x, y = b % ntx, (b // ntx) % nty
        i0, j0 = 3 + 32 * x, 3 + 8 * y
        return i0 <= 3 + nl or i0 + 31 >= ni + 2 - nl or j0 <= 3 + nl or j0 + 7 >= nj + 2 - nl
    flags = [is_pml(b) for b in order]
    first_fast = flags.index(False)
    assert all(flags[:first_fast]) and not any(flags[first_fast:])
    assert 0 < first_fast < len(order)
    # within a band of one wave (296 tiles) the chunks of a tile follow each other in the direction of the march
    ntile = ntx * nty
    for dz in (1, 0):
        _, od = solver.launch_plan(grid, ((nl, nl), (nl, nl), (nl, 0)), 1, (0, ntx, 0, nty), dz=dz)
        assert sorted(od) == list(range(ntile * nzc))
        pos = {b: n for n, b in enumerate(od)}
        t = od[first_fast] % ntile
        seq = [pos[z * ntile + t] for z in (range(nzc) if dz else range(nzc - 1, -1, -1))]
        assert seq == sorted(seq) and seq[1] - seq[0] <= 296
    # the one-tile-wide rectangles of the boundary phase (tiles next to an inter-rank face) are permutations too
    for rect in ((0, 1, 0, nty), (ntx - 1, ntx, 0, nty), (1, ntx - 1, 0, 1), (1, ntx - 1, nty - 1, nty)):
        for dz in (0, 1):
            zc, od = solver.launch_plan(grid, ((0, nl), (nl, 0), (nl, 0)), 1, rect, dz=dz)
            nb = (rect[1] - rect[0]) * (rect[3] - rect[2]) * -(-(nk - 4) // zc)
            assert sorted(od) == list(range(nb)), (rect, dz)
    # no PML, one chunk-major band when there are fewer tiles than one wave
    zc2, order2 = solver.launch_plan(dict(grid, nx=70, ni2=66), ((0, 0), (0, 0), (0, 0)), 0, (0, 2, 0, nty))
    assert order2 == list(range(len(order2)))


def test_launch_plan_chunk_length_does_not_grow_with_the_grid():
    """z chunks stay short whatever the grid (bands of tiles restart together at every chunk boundary, which keeps the halo lines
    neighbouring tiles share in L2: 800x800x400 ran 12 % slower with the 198-row chunks the wave count alone allows, profiles/
    r2_experiments.txt r2p): ~18 rows with two resident blocks per SM, ~25 with one; small rectangles get more, shorter chunks,
    never under 16 rows"""
    from cgfd3d_b200 import solver
    nl = 10
    for (ni, nj, nk) in ((400, 400, 200), (800, 800, 400), (1200, 600, 600), (150, 300, 300)):
        grid = dict(nx=ni + 6, ny=nj + 6, nz=nk + 6, ni1=3, ni2=ni + 2, nj1=3, nj2=nj + 2, nk1=3, nk2=nk + 2)
        ntx, nty = (ni + 31) // 32, (nj + 7) // 8
        for bps, lo, hi in ((2, 16, 19), (1, 16, 26)):
            for free_top in (0, 1):
                zc, od = solver.launch_plan(grid, ((nl, nl), (nl, nl), (nl, 0)), free_top, (0, ntx, 0, nty), blocks_per_sm=bps)
                rows = nk - 4 * free_top
                assert lo <= zc <= hi, (ni, nj, nk, bps, zc)
                assert len(od) == ntx * nty * -(-rows // zc)
    # a one-tile-wide boundary rectangle of a small block: chunks shrink towards 16 rows to make enough blocks, not below
    grid = dict(nx=106, ny=106, nz=66, ni1=3, ni2=102, nj1=3, nj2=102, nk1=3, nk2=62)
    zc, od = solver.launch_plan(grid, ((0, nl), (nl, nl), (nl, 0)), 0, (0, 1, 0, 13))
    assert zc == 20 and len(od) == 13 * 3   # 4 chunks would be 15 rows
    # fewer rows than one chunk: a single chunk
    grid = dict(nx=46, ny=46, nz=17, ni1=3, ni2=42, nj1=3, nj2=42, nk1=3, nk2=13)
    zc, od = solver.launch_plan(grid, ((0, 0), (0, 0), (0, 0)), 0, (0, 2, 0, 5))
    assert zc == 11 and len(od) == 10


def test_bench_rank_problems_for_the_8_gpu_grid():
    """bench.build_rank_problem on the 4x2 process grid of the 8-GPU run (small blocks, numpy set-up): neighbours follow the
    reference's row-major cart (forward/mympi_t.c:32-40), inter-rank faces carry no PML, exactly one rank owns the source, and the
    ghost metrics of an inter-rank face equal the neighbour's physical values."""
    import bench
    size, n = (24, 16, 12), 8
    assert bench.proc_grid(n) == (4, 2)
    probs = [bench.build_rank_problem(size, r, n) for r in range(n)]
    owners = [r for r, p in enumerate(probs) if p.src and p.src.get("total_number", 0) > 0]
    assert len(owners) == 1
    px, py = 4, 2
    for r, p in enumerate(probs):
        ix, iy = r // py, r % py
        want = (r - py if ix > 0 else -1, r + py if ix < px - 1 else -1, r - 1 if iy > 0 else -1, r + 1 if iy < py - 1 else -1)
        assert tuple(p.neigh) == want
        for side in range(4):
            has_pml = (side // 2, side % 2) in p.pml
            assert has_pml == (want[side] < 0)
        assert (2, 0) in p.pml and p.free_top == 1
    # x2 ghosts of rank 0 (ix=0, iy=0) = first physical columns of rank 2 (ix=1, iy=0), for every metric array
    a, b = probs[0], probs[2]
    for m in range(10):
        np.testing.assert_array_equal(a.metric[m][3:-3, 3:-3, -3:], b.metric[m][3:-3, 3:-3, 3:6])
