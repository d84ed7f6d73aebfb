import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgfd3d_b200 import abi, solver, hostsetup as hs
from oracle import ref_flat
from tests import util
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 24
prob = util.small_problem(ni=70, nj=30, nk=28, pml_layers=5, nt_total=nt, seed=11)
for m in (abi.M_XIY, abi.M_XIZ, abi.M_ETX, abi.M_ETZ):
    prob.metric[m][...] = 0.0
mvx, mvy, mf = hs.dvh2dvz_iso(prob.metric, prob.media[0], prob.media[1], prob.grid)
prob.mats = dict(matVx2Vz=mvx, matVy2Vz=mvy, matF2Vz=mf, matD=np.zeros_like(mf))
wr, _, _ = ref_flat.RefSolver(prob).run(nt)
out = {}
for gz in ("1", "0", "1"):
    os.environ["CGFD_GZ"] = gz
    G = solver.Solver(prob)
    print("gz", gz, "class", G.grid_class())
    G.run(nt)
    w = G.get_wavefield()
    G.close()
    if gz in out:
        print("repeat identical:", np.array_equal(out[gz], w))
    out[gz] = w
    print(" vs ref:", ["%.2e" % util.rel_l2(w[c], wr[c]) for c in range(9)])
d = np.abs(out["1"].astype(np.float64) - out["0"])
print("gz1 vs gz0 rel_l2:", ["%.2e" % util.rel_l2(out["1"][c], out["0"][c]) for c in range(9)])
idx = np.unravel_index(np.argmax(d), d.shape)
print("max diff at", idx, d[idx], out["1"][idx], out["0"][idx], "nz,ny,nx", prob.nz, prob.ny, prob.nx)
for c in range(9):
    dc = d[c]
    nzk = np.nonzero(dc.max(axis=(1, 2)) > 0)[0]
    print(c, "k range with diffs", (nzk.min(), nzk.max()) if len(nzk) else None, "count", int((dc > 0).sum()))
