#!/usr/bin/env python
"""which planes / tiles hold non-finite values after a few steps of a big block"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cgfd3d_b200 import solver

size = tuple(int(v) for v in sys.argv[1].split("x"))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.cuda.set_device(0)
prob = bench.build_rank_problem(size, 0, 1, device="cuda:0")
S = solver.Solver(prob, device=0)
prob.metric = prob.media = None
torch.cuda.empty_cache()
for it in range(nsteps):
    S.run(1, it0=it)
    for comp in (2, 3):
        bad_k = []
        for k in list(range(0, 8)) + list(range(prob.nz - 26, prob.nz)):
            b = S.get_box(comp, 0, prob.nx, 1, 0, prob.ny, 1, k, 1, 1)[0]
            bad = ~np.isfinite(b)
            if bad.any():
                jj, ii = np.nonzero(bad)
                bad_k.append((k, int(bad.sum()), int(ii.min()), int(ii.max()), int(jj.min()), int(jj.max())))
        print("size", size, "after step", it, "comp", comp, "bad planes", len(bad_k), flush=True)
        for r in bad_k[:6] + bad_k[-3:]:
            print("  k=%d nbad=%d i[%d..%d] j[%d..%d]" % r)
    if bad_k:
        break
