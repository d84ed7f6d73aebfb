#!/usr/bin/env python
"""One small run of the hot path for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python scripts/sanitize_case.py <iso|vti|aniso|visco> [nsteps]
40x36x30 hill block, CFS-PML on five faces + free surface, Gaussian moment source, record points, a streaming snapshot,
whole-wavefield transfers. ni = 40 leaves a partial tile in x, nj = 36 one in y. Prints max|w| so a run that computed nothing shows."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgfd3d_b200 import hostsetup as hs, solver  # noqa: E402


def main():
    medium = sys.argv[1] if len(sys.argv) > 1 else "iso"
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    ni, nj, nk = 40, 36, 30
    prob = hs.build_problem(ni, nj, nk, topo="hill", hill=(300.0, 600.0), pml_layers=6, free_top=True, dt_safety=0.9, medium=medium,
                            seed=None if medium == "iso" else 5)
    if medium != "iso":
        # any well-conditioned matrices exercise the same memory accesses (the real ones are checked by the parity suite)
        eye = np.tile(np.eye(3, dtype=np.float32).reshape(-1) * 0.1, prob.nx * prob.ny)
        prob.mats = dict(matVx2Vz=eye.copy(), matVy2Vz=eye.copy(), matF2Vz=eye.copy(), matD=np.tile(np.eye(3, dtype=np.float32).reshape(-1), prob.nx * prob.ny))
    hs.make_source(prob, ni // 2, nj // 2, nk - 8, nt_total=nt, kind="moment", spatial="gauss", fc=3.0, t0=0.3, stf_len=0.8)
    S = solver.Solver(prob)
    S.set_record_points([prob.iptr(5 + 3 * n, 7 + 2 * n, nk - 1) for n in range(6)], nt)
    sid, frames = S.add_snapshot((0, 1, 2), (3, ni // 2, 2, 3, nj // 2, 2, prob.nz - 4, 1, 1), max_frames=nt)
    S.run(nt)
    w = S.get_wavefield()
    S.set_wavefield(w)
    S.run(2, it0=nt)
    rec = S.get_record(0, nt)
    S.close()
    print("sanitize_case %s: max|Vz| %.4e  max|rec| %.4e  frames %d  finite %s" % (medium, float(np.abs(w[2]).max()), float(np.abs(rec).max()),
                                                                               S.snapshot_frames(sid) if False else len(frames), bool(np.isfinite(w).all())))


if __name__ == "__main__":
    main()
