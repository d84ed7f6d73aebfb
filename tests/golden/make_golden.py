#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the reference program itself
(oracle/_ref/ref_main_zero: unmodified CGFD3D sources + single-rank MPI / dense-container NetCDF stand-ins,
traction image with the ZERO guard) on BASELINE.json configs[0]:
isotropic elastic halfspace, flat free surface, 100x100x60, explosive point source, CFS-PML on 5 sides,
1000 time steps. Needs /root/reference (to build oracle/_ref); the outputs are committed so that the GPU box
can check parity without it.

  python tests/golden/make_golden.py [config1|small]
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
    # name: (ni, nj, nk, nt, dt, source (i,j,depth), line start/incr/count, snapshot stride)
    "config1": dict(n=(100, 100, 60), nt=1000, dt=0.025, src=(50, 50, 20), line=((20, 30, 59), (15, 10, 0), 5), tinc=100, sinc=4),
    "small": dict(n=(48, 44, 32), nt=200, dt=0.025, src=(24, 22, 10), line=((8, 10, 31), (8, 6, 0), 5), tinc=20, sinc=2),
}


def case_files(name, workdir):
    c = CASES[name]
    ni, nj, nk = c["n"]
    (l0, linc, lcnt) = c["line"]
    par = H.make_par(workdir, ni, nj, nk, c["nt"], c["dt"], pml_layers=10,
                     lines=[{"name": "L1", "grid_index_start": list(l0), "grid_index_incre": list(linc), "grid_index_count": lcnt}],
                     snapshots=[{"name": "surf", "grid_index_start": [0, 0, nk - 1], "grid_index_count": [ni // c["sinc"], nj // c["sinc"], 1],
                                 "grid_index_incre": [c["sinc"], c["sinc"], 1], "time_index_start": 0, "time_index_incre": c["tinc"],
                                 "save_velocity": 1, "save_stress": 0, "save_strain": 0}])
    src = H.moment_src(*c["src"])
    stations = [("r1", 0, 1, ni // 2 + 10, nj // 2 + 5, 0)]
    return par, src, stations


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "config1"
    wd = tempfile.mkdtemp(prefix="cgfd_golden_")
    par, src, stations = case_files(name, wd)
    H.write_case(wd, par, src, stations)
    wall, out = H.run(H.ref_binary("ref_main_zero"), wd, timeout=7200)
    sac = H.read_sac_dir(os.path.join(wd, "OUT"))
    keep = {k.replace(".", "_"): v for k, v in sac.items() if ".L1." in k}
    snap = H.read_cgnc(os.path.join(wd, "OUT", "surf_px0_py0.nc"))
    for v in ("Vx", "Vy", "Vz"):
        keep["snap_" + v] = snap["vars"][v]
    keep["snap_time"] = snap["vars"]["time"]
    keep["wall_s"] = np.array([wall])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name), **keep)
    print("wrote ref_%s.npz: %d arrays, reference wall %.1f s" % (name, len(keep), wall))
    shutil.rmtree(wd)


if __name__ == "__main__":
    main()
