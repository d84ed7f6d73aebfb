import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgfd3d_b200 import abi, solver, hostsetup as hs
from oracle import ref_flat
from tests import util

def scenario(medium, nt, zero=True):
    prob = util.small_problem(ni=70, nj=30, nk=28, pml_layers=5, nt_total=24, seed=11, medium=medium)
    if zero:
        for m in (abi.M_XIY, abi.M_XIZ, abi.M_ETX, abi.M_ETZ):
            prob.metric[m][...] = 0.0
    if medium == "iso":
        mvx, mvy, mf = hs.dvh2dvz_iso(prob.metric, prob.media[0], prob.media[1], prob.grid)
        prob.mats = dict(matVx2Vz=mvx, matVy2Vz=mvy, matF2Vz=mf, matD=np.zeros_like(mf))
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    wr, _, _ = R.run(nt)
    G = solver.Solver(prob)
    G.run(nt)
    w = G.get_wavefield()
    G.close()
    errs = [util.rel_l2(w[c], wr[c]) for c in range(9)]
    print(medium, "nt", nt, "gz", os.environ.get("CGFD_GZ"), "max err vs ref %.2e" % max(errs), flush=True)
    if max(errs) > 1e-4:
        d = np.abs(w.astype(np.float64) - wr)
        for c in range(9):
            s = max(float(np.abs(wr[c]).max()), 1e-30)
            dk = d[c].max(axis=(1, 2)) / s; dj = d[c].max(axis=(0, 2)) / s; di = d[c].max(axis=(0, 1)) / s
            print("  cmp", c, "k:", np.nonzero(dk > 1e-5)[0][[0, -1]] if (dk > 1e-5).any() else None,
                  "j:", np.nonzero(dj > 1e-5)[0][[0, -1]] if (dj > 1e-5).any() else None,
                  "i:", np.nonzero(di > 1e-5)[0][[0, -1]] if (di > 1e-5).any() else None, "max rel %.2e" % (d[c].max() / s))
    return w

order = sys.argv[1:] or ["iso", "vti"]
for spec in order:
    med, nt = spec.split(":") if ":" in spec else (spec, "1")
    scenario(med, int(nt))
