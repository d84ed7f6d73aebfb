#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the reference program itself
(oracle/_ref/ref_main_zero: unmodified CGFD3D sources + single-rank MPI / dense-container NetCDF stand-ins,
traction image with the ZERO guard). Needs /root/reference (to build oracle/_ref); the outputs are
committed so that the GPU box can check parity without it.

  config1 : BASELINE.json configs[0]: isotropic elastic halfspace, flat free surface (Cartesian), 100x100x60, explosive point
            source, CFS-PML on 5 sides, 1000 steps; line receivers + surface snapshots.
  small   : the same at 48x44x32, 200 steps.
  hill100 : the physics of configs[1] on the reference's CURVILINEAR input route: Gaussian-hill grid imported as
            coord_px0_py0.nc (gd_curv_coord_import, forward/gd_t.c:741-783), metrics by the reference's gd_curv_metric_cal,
            100x100x60, 1000 steps; line receivers, a station at a FRACTIONAL grid position (8-corner interpolation of
            io_recv_keep, forward/io_funcs.c:1584-1623, incl. its strain outputs), surface snapshots.
  hill200 : 200x200x100 hill, 100 steps; line receivers, station, surface snapshot and x / y / z slices
            (io_slice_nc_put, forward/io_funcs.c:991-1107).

  python tests/golden/make_golden.py [config1|small|hill100|hill200]
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402

CASES = {
    # n = (ni, nj, nk); src = (i, j, depth); line = (start, incr, count); tinc / sinc = snapshot time / space stride
    "config1": dict(n=(100, 100, 60), nt=1000, dt=0.025, src=(50, 50, 20), line=((20, 30, 59), (15, 10, 0), 5), tinc=100, sinc=4),
    "small": dict(n=(48, 44, 32), nt=200, dt=0.025, src=(24, 22, 10), line=((8, 10, 31), (8, 6, 0), 5), tinc=20, sinc=2),
    "hill100": dict(n=(100, 100, 60), nt=1000, dt=0.019, src=(50, 50, 20), line=((20, 30, 59), (15, 10, 0), 5), tinc=100, sinc=4,
                    hill=(1000.0, 1500.0), station=("r1", 0, 1, 60.3, 55.6, 2.4)),
    "hill200": dict(n=(200, 200, 100), nt=100, dt=0.02, src=(100, 100, 12), line=((40, 60, 99), (30, 20, 0), 5), tinc=10, sinc=8,
                    hill=(1500.0, 3000.0), station=("r1", 0, 1, 110.25, 95.5, 1.75),
                    slices={"x_index": [104], "y_index": [93], "z_index": [80]}, slice_keep=(49, 99), slice_cmps=("Vx", "Vz", "Txz"), slice_stride=2,
                    src_m=(1e16, 0.6e16, 1.4e16, 0.2e16, -0.3e16, 0.1e16)),
}
SLICE_FILES = {"x_index": "slicex_i%d_px0_py0.nc", "y_index": "slicey_j%d_px0_py0.nc", "z_index": "slicez_k%d_px0_py0.nc"}


def case_files(name, workdir=None):
    """(par, src text, station list, coords or None) of a case"""
    c = CASES[name]
    ni, nj, nk = c["n"]
    (l0, linc, lcnt) = c["line"]
    coords, grid = None, None
    if "hill" in c:
        from cgfd3d_b200 import hostsetup as hs
        coords = hs.hill_coords(ni, nj, nk, height=c["hill"][0], sigma=c["hill"][1])
        grid = {"import": "IN"}
    par = H.make_par(workdir, ni, nj, nk, c["nt"], c["dt"], pml_layers=10, grid=grid, slices=c.get("slices"),
                     lines=[{"name": "L1", "grid_index_start": list(l0), "grid_index_incre": list(linc), "grid_index_count": lcnt}],
                     snapshots=[{"name": "surf", "grid_index_start": [0, 0, nk - 1], "grid_index_count": [ni // c["sinc"], nj // c["sinc"], 1],
                                 "grid_index_incre": [c["sinc"], c["sinc"], 1], "time_index_start": 0, "time_index_incre": c["tinc"],
                                 "save_velocity": 1, "save_stress": 0, "save_strain": 0}])
    src = H.moment_src(*c["src"], m=c.get("src_m", (1e16, 1e16, 1e16, 0, 0, 0)))
    stations = [c.get("station", ("r1", 0, 1, ni // 2 + 10, nj // 2 + 5, 0))]
    return par, src, stations, coords


def collect(name, outdir):
    """everything of a finished run that the fixtures keep / the tests compare, as a flat dict of arrays:
    evt1_L1_no<n>_<cmp> line seismograms, sta_<cmp> the station's seismograms (9 wavefield + 6 strain components),
    snap_<V> surface snapshots, slice<x|y|z>_<cmp> selected frames of the slices (strided)"""
    c = CASES[name]
    sac = H.read_sac_dir(outdir)
    keep = {k.replace(".", "_"): v for k, v in sac.items() if ".L1." in k}
    if "station" in c:
        # hazard 5 of SURVEY.md 8c: the station name may be lost by the reference's overlapping sprintf; take whatever it wrote
        sta = {k: v for k, v in sac.items() if ".L1." not in k}
        for k, v in sta.items():
            keep["sta_" + k.split(".")[-1]] = v
        assert len(sta) == 15, sorted(sta)
    snap = H.read_cgnc(os.path.join(outdir, "surf_px0_py0.nc"))
    for v in ("Vx", "Vy", "Vz"):
        keep["snap_" + v] = snap["vars"][v]
    keep["snap_time"] = snap["vars"]["time"]
    for key, idxs in (c.get("slices") or {}).items():
        sl = H.read_cgnc(os.path.join(outdir, SLICE_FILES[key] % idxs[0]))
        st = c["slice_stride"]
        for v in c["slice_cmps"]:
            keep["slice%s_%s" % (key[0], v)] = np.ascontiguousarray(sl["vars"][v][list(c["slice_keep"])][:, ::st, ::st])
    return keep


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "config1"
    wd = tempfile.mkdtemp(prefix="cgfd_golden_")
    par, src, stations, coords = case_files(name, wd)
    H.write_case(wd, par, src, stations, coords)
    wall, out = H.run(H.ref_binary("ref_main_zero"), wd, timeout=7200)
    keep = collect(name, os.path.join(wd, "OUT"))
    keep["wall_s"] = np.array([wall])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name), **keep)
    print("wrote ref_%s.npz: %d arrays, reference wall %.1f s" % (name, len(keep), wall))
    shutil.rmtree(wd)


if __name__ == "__main__":
    main()
