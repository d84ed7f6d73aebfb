#!/usr/bin/env python
"""The drop-in driver's own rate on BASELINE.json configs[1]: the reference program (its main, par file, set-up code and io_*
writers) linked against integration/drv_rk_curv_col_b200.c + libcgfd3d_b200.so, 400x400x200 Gaussian-hill grid through
gd_curv_coord_import, CFS-PML 10 x 5 + free surface, one moment source, the example's outputs: a surface Vx/Vy/Vz snapshot EVERY
step (example/cgfd3d.example.sh:323-334), a receiver line and a station. Prints the driver's "GPU time loop" line.
  python scripts/dropin_config1.py [nsteps] [NIxNJxNK]"""
import os
import re
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402


def main():
    nt = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    size = tuple(int(v) for v in sys.argv[2].split("x")) if len(sys.argv) > 2 else (400, 400, 200)
    ni, nj, nk = size
    wd = tempfile.mkdtemp(prefix="cgfd_cfg1_")
    t0 = time.time()
    H.write_multirank_hill_case(
        wd, size, 1, 1, nt, 0.012, hill=(1000.0, 0.1 * max(ni, nj) * 100.0), pml_layers=10, src=H.moment_src(ni // 2, nj // 2, 20),
        lines=[{"name": "L1", "grid_index_start": [ni // 8, nj // 2, nk - 1], "grid_index_incre": [ni // 16, 0, 0], "grid_index_count": 12}],
        snapshots=[{"name": "surf", "grid_index_start": [0, 0, nk - 1], "grid_index_count": [ni, nj, 1], "grid_index_incre": [1, 1, 1],
                    "time_index_start": 0, "time_index_incre": 1, "save_velocity": 1, "save_stress": 0, "save_strain": 0}])
    t_in = time.time() - t0
    wall, out = H.run(H.ref_binary("cgfd_main_b200"), wd, timeout=3000)
    m = re.search(r"GPU time loop: (\d+) steps in ([0-9.]+) s = ([0-9.]+) Gpoint", out)
    snap = os.path.getsize(os.path.join(wd, "OUT", "surf_px0_py0.nc"))
    print("DROPIN_CONFIG1 size %dx%dx%d steps %d: time loop %s s = %s Gpoint-updates/s; program wall %.1f s (input files %.1f s); snapshot file %.1f MB"
          % (size + (nt, m.group(2) if m else "?", m.group(3) if m else "?", wall, t_in, snap / 1e6)))
    if not m:
        print(out[-2000:])


if __name__ == "__main__":
    main()
