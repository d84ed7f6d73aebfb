#!/usr/bin/env python
"""finite-ness probe of a big single-GPU block: run in chunks, print max |V| on a few planes after every chunk."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cgfd3d_b200 import solver

size = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "800x800x400").split("x"))
torch.cuda.set_device(0)
prob = bench.build_rank_problem(size, 0, 1, device="cuda:0")
S = solver.Solver(prob, device=0)
prob.metric = prob.media = None
torch.cuda.empty_cache()
ni, nj, nk = size
it = 0
for chunk in range(12):
    S.run(4, it0=it); it += 4
    out = []
    for (c, k) in ((2, prob.nz - 4), (0, prob.nz - 4), (3, prob.nz - 30), (2, 3 + nk // 2), (2, 3), (5, 8)):
        b = S.get_box(c, 0, prob.nx, 1, 0, prob.ny, 1, k, 1, 1)
        bad = ~np.isfinite(b)
        where = np.argwhere(bad)[:3].tolist() if bad.any() else []
        out.append("c%d k%d max %.3e nbad %d %s" % (c, k, float(np.nanmax(np.abs(b))), int(bad.sum()), where))
    print("it", it, " | ".join(out), flush=True)
