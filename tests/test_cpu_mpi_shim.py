"""The fork / shared-memory MPI stand-in (oracle/shims/mpi_shim.c, TEST INFRASTRUCTURE) runs the reference program on several
ranks of one host: N ranks must reproduce 1 rank. That pins (a) the stand-in itself (persistent-request halo exchange,
MPI_Sendrecv with vector datatypes in the metric exchange, collectives), (b) cgfd3d_b200.decomp.ref_split = gd_indx_set's rule
(forward/gd_t.c:2775-2849: PML layers count as load) -- a wrong block size would make the per-rank coordinate files the wrong
shape -- and gives bench.py's reference arm its multi-rank CPU run."""
import os
import tempfile

import numpy as np
import pytest

from cgfd3d_b200 import decomp
from oracle import harness as H

need = pytest.mark.skipif(not H.have_ref("ref_main_zero"), reason="oracle/_ref/ref_main_zero not built (make -C oracle ref)")


def test_ref_split_deals_out_every_point_once():
    for n, parts, l1, l2 in ((400, 4, 10, 10), (100, 3, 10, 0), (61, 2, 6, 6), (50, 1, 6, 6), (403, 4, 0, 0)):
        nxt = 0
        for w in range(parts):
            start, cnt = decomp.ref_split(n, parts, w, l1, l2)
            assert start == nxt and cnt > 0
            nxt = start + cnt
        assert nxt == n
    # the edge blocks give their absorbing layers back: 420 / 4 = 105 -> 95, 105, 105, 95
    assert [decomp.ref_split(400, 4, w, 10, 10)[1] for w in range(4)] == [95, 105, 105, 95]


def _run(size, px, py, nt, hill, **kw):
    wd = tempfile.mkdtemp(prefix="cgfd_mpi_")
    if hill:
        H.write_multirank_hill_case(wd, size, px, py, nt, 0.015, hill=hill, pml_layers=6, **kw)
    else:
        par = H.make_par(wd, size[0], size[1], size[2], nt, 0.02, pml_layers=6, src_spatial="point", lines=kw.get("lines"))
        par["number_of_mpiprocs_x"], par["number_of_mpiprocs_y"] = px, py
        H.write_case(wd, par, kw["src"], [("r1", 0, 1, 25, 20, 0)])
    env = dict(os.environ, CGFD_SHIM_NPROCS=str(px * py))
    H.run(H.ref_binary("ref_main_zero"), wd, timeout=900, env=env)
    return H.read_sac_dir(os.path.join(wd, "OUT"))


def _worst(a, b):
    w = 0.0
    assert set(a) == set(b) and len(a) > 30
    for k in a:
        if float(np.abs(a[k]).max()) > 0:
            w = max(w, float(np.linalg.norm(a[k].astype(np.float64) - b[k]) / np.linalg.norm(a[k].astype(np.float64))))
    return w


@need
def test_cartesian_ranks_bit_identical():
    size, nt = (48, 44, 32), 50
    kw = dict(src=H.moment_src(22, 23, 9, m=(1e16, 0.5e16, 2e16, 0.2e16, -0.1e16, 0.3e16)),
              lines=[{"name": "L1", "grid_index_start": [6, 8, 31], "grid_index_incre": [7, 6, 0], "grid_index_count": 5}])
    one = _run(size, 1, 1, nt, None, **kw)
    assert max(float(np.abs(v).max()) for v in one.values()) > 0
    for (px, py) in ((2, 2), (3, 1)):
        assert _worst(one, _run(size, px, py, nt, None, **kw)) == 0.0, (px, py)


@need
def test_hill_ranks_match_one_rank():
    """curvilinear grid through per-rank coord_px?_py?.nc files: the ranks estimate the PML slab length on their own part of the
    slab (bdry_cal_abl_len_dh), so the profiles differ in the last digits and the seismograms by ~1e-6"""
    size, nt = (61, 50, 30), 40
    kw = dict(src=H.moment_src(30, 26, 8, m=(1e16, 0.5e16, 2e16, 0.2e16, -0.1e16, 0.3e16)),
              lines=[{"name": "L1", "grid_index_start": [6, 8, 29], "grid_index_incre": [9, 6, 0], "grid_index_count": 6}])
    one = _run(size, 1, 1, nt, (600.0, 1200.0), **kw)
    w = _worst(one, _run(size, 3, 2, nt, (600.0, 1200.0), **kw))
    assert w <= 5e-5, w


@need
def test_reference_arm_times_the_loop_from_the_programs_own_step_lines():
    """bench.py's reference arm: one run of the reference program with verbose > 10, the '-> it=' lines of its driver
    (forward/drv_rk_curv_col.c:172) stamped on arrival. Every step must be seen, in order, and the outputs must be there."""
    import bench
    cb = bench.cpu_baseline(2, steps=3, warm=1, size=(48, 40, 24), hill=(400.0, 800.0))
    assert cb["kind"] == "reference" and cb["cores"] == 2 and cb["ranks"] == "2x1" and cb["finite"]
    assert cb["value"] > 0 and cb["seconds_per_step"] > 0
    import shutil
    if shutil.which("stdbuf"):
        lo, hi = cb["seconds_per_step_min_max"]
        assert 0 < lo <= cb["seconds_per_step"] <= hi
        assert "stamped on arrival" in cb["sample"]
    wd = tempfile.mkdtemp(prefix="cgfd_mpi_")
    H.write_multirank_hill_case(wd, (48, 40, 24), 1, 1, 4, 0.015, hill=(400.0, 800.0), pml_layers=6, src=H.moment_src(24, 20, 8))
    wall, stamps = H.run_timed(H.ref_binary("ref_main_zero"), wd, timeout=300, env=dict(os.environ, CGFD_SHIM_NPROCS="1"))
    if shutil.which("stdbuf"):
        # number_of_time_steps = 4 -> nt_total = 5 steps (forward/main_curv_col_el_3d.c:615)
        assert sorted(stamps) == [0, 1, 2, 3, 4] and all(stamps[n + 1] > stamps[n] for n in range(4)) and stamps[4] < wall
