"""Set-up arrays of a synthetic subdomain built ON THE GPU with torch (input generation only, not the hot path).

hostsetup.build_problem needs ~45 GB of host memory and minutes of single-threaded numpy for one 800x800x400 block
(BASELINE.json configs[2], one such block per GPU); this module evaluates the same formulas
(forward/gd_t.c:190-402 metrics with mirrored ghosts, the Gaussian-hill coordinates of hostsetup.hill_coords) as torch
tensor expressions on the rank's own GPU and hands DEVICE pointers to cgfd_b200_create (the library accepts host or
device pointers for the metric / media arrays). The small pieces (PML profiles, free-surface matrices, dt) reuse the
numpy code of hostsetup on slabs / planes copied back to the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import abi, hostsetup as hs

NG = hs.NG


def _cdiff(a, axis):
    nz, ny, nx = a.shape
    sl = [slice(NG, nz - NG), slice(NG, ny - NG), slice(NG, nx - NG)]
    out = None
    for off, c in zip(hs.FDC_INDX, hs.FDC_COEF):
        s = list(sl)
        s[axis] = slice(NG + off, a.shape[axis] - NG + off)
        term = a[tuple(s)] * float(np.float32(c))
        out = term if out is None else out.add_(term)
    return out


def metric_from_coords(x, y, z):
    """torch twin of hostsetup.metric_from_coords; x, y, z float32 tensors [nz][ny][nx]. Returns 10 tensors."""
    nz, ny, nx = x.shape
    v1 = [_cdiff(a, 2) for a in (x, y, z)]
    v2 = [_cdiff(a, 1) for a in (x, y, z)]
    v3 = [_cdiff(a, 0) for a in (x, y, z)]

    def cross(a, b):
        return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

    g = cross(v1, v2)
    jac = g[0] * v3[0] + g[1] * v3[1] + g[2] * v3[2]
    out = []

    def finish(p):
        full = torch.zeros((nz, ny, nx), dtype=torch.float32, device=x.device)
        full[NG:nz - NG, NG:ny - NG, NG:nx - NG] = p
        for g_ in range(NG):
            full[:, :, NG - 1 - g_] = full[:, :, NG + g_]
            full[:, :, nx - NG + g_] = full[:, :, nx - NG - 1 - g_]
        for g_ in range(NG):
            full[:, NG - 1 - g_, :] = full[:, NG + g_, :]
            full[:, ny - NG + g_, :] = full[:, ny - NG - 1 - g_, :]
        for g_ in range(NG):
            full[NG - 1 - g_, :, :] = full[NG + g_, :, :]
            full[nz - NG + g_, :, :] = full[nz - NG - 1 - g_, :, :]
        return full

    out.append(finish(jac))
    for a, b in ((v2, v3), (v3, v1), (v1, v2)):   # xi, eta, zeta rows
        for comp in cross(a, b):
            out.append(finish(comp / jac))
    return out


def hill_coords(ni, nj, nk, dh, height, sigma, gi0, gj0, gni, gnj, device):
    nx, ny, nz = ni + 2 * NG, nj + 2 * NG, nk + 2 * NG
    x1 = (torch.arange(nx, dtype=torch.float64, device=device) - NG + gi0) * dh[0]
    y1 = (torch.arange(ny, dtype=torch.float64, device=device) - NG + gj0) * dh[1]
    xc = 0.5 * (gni - 1) * dh[0]
    yc = 0.5 * (gnj - 1) * dh[1]
    r2 = (x1[None, :] - xc) ** 2 + (y1[:, None] - yc) ** 2
    ztop = height * torch.exp(-r2 / (2.0 * sigma * sigma))
    zbot = -(nk - 1) * dh[2]
    s = ((torch.arange(nz, dtype=torch.float64, device=device) - NG) / float(nk - 1))[:, None, None]
    z = (zbot + s * (ztop[None, :, :] - zbot)).to(torch.float32)
    x = x1.to(torch.float32)[None, None, :].expand(nz, ny, nx).contiguous()
    y = y1.to(torch.float32)[None, :, None].expand(nz, ny, nx).contiguous()
    return x, y, z


class _Slab:
    """numpy view of the part of a device array hostsetup.pml_profiles reads (only the slab is copied)."""

    def __init__(self, t):
        self.t = t

    def __getitem__(self, idx):
        return self.t[idx].cpu().numpy()


def build_problem(ni, nj, nk, *, device, dh=(100.0, 100.0, 100.0), hill=(1000.0, 2000.0), vp=3000.0, vs=2000.0, rho=1500.0,
                  pml_layers=10, pml_faces=((0, 0), (0, 1), (1, 0), (1, 1), (2, 0)), free_top=True, dt=0.012, sub=None, medium="iso",
                  nmaxwell=3):
    """Device twin of hostsetup.build_problem(topo='hill'): the returned HostProblem holds torch CUDA tensors in
    .metric / .media (float32, contiguous)."""
    if sub is None:
        gi0 = gj0 = 0
        gni, gnj = ni, nj
        neigh = (-1, -1, -1, -1)
    else:
        gi0, gj0, gni, gnj, neigh = sub
    dev = torch.device(device)
    ex = [NG if neigh[n] >= 0 else 0 for n in range(4)]
    eni, enj = ni + ex[0] + ex[1], nj + ex[2] + ex[3]
    x, y, z = hill_coords(eni, enj, nk, dh, hill[0], hill[1], gi0 - ex[0], gj0 - ex[2], gni, gnj, dev)
    metric = metric_from_coords(x, y, z)
    if any(ex):
        cs = (slice(None), slice(ex[2], ex[2] + nj + 2 * NG), slice(ex[0], ex[0] + ni + 2 * NG))
        x, y, z = (a[cs].contiguous() for a in (x, y, z))
        metric = [m[cs].contiguous() for m in metric]
    prob = hs.HostProblem(ni=ni, nj=nj, nk=nk, dt=float(np.float32(dt)), free_top=1 if free_top else 0, neigh=tuple(neigh))
    prob.metric = metric
    shape = (prob.nz, prob.ny, prob.nx)
    mu = np.float32(rho * vs * vs)
    lam = np.float32(rho * vp * vp - 2.0 * rho * vs * vs)
    prob.media = [torch.full(shape, float(v), dtype=torch.float32, device=dev) for v in (lam, mu, np.float32(1.0) / np.float32(rho))]
    g = prob.grid
    for (idim, iside) in pml_faces:
        if idim < 2 and neigh[idim * 2 + iside] >= 0:
            continue
        A, B, D = hs.pml_profiles(_Slab(x), _Slab(y), _Slab(z), g, idim, iside, pml_layers)
        prob.pml[(idim, iside)] = (pml_layers, A, B, D)
    if medium != "iso":
        # the homogeneous test media of hostsetup.set_test_medium (values after forward/md_t.c:501-595, 912-952) as device fills
        slw = float(np.float32(1.0 / rho))
        def fill(v):
            return torch.full(shape, float(np.float32(v)), dtype=torch.float32, device=dev)
        if medium == "vti":
            prob.medium_type = abi.MEDIUM_ELASTIC_VTI
            prob.media = [fill(v) for v in (25.2e9, 10.962e9, 18.0e9, 5.12e9, 7.168e9)] + [fill(slw)]
        elif medium == "aniso":
            prob.medium_type = abi.MEDIUM_ELASTIC_ANISO
            c11, c13, c33, c55, c66 = 25.2e9, 10.962e9, 18.0e9, 5.12e9, 7.168e9
            c12, e = c11 - 2 * c66, 0.15e9
            C = [c11, c12, c13, e, -e, 0.5 * e, c11, c13, -0.5 * e, e, 0.7 * e, c33, 0.3 * e, -0.6 * e, e, c55, 0.4 * e, -0.2 * e, c55, 0.8 * e, c66]
            prob.media = [fill(v) for v in C] + [fill(slw)]
        elif medium == "visco":
            prob.medium_type = abi.MEDIUM_VISCOELASTIC_ISO
            prob.nmaxwell = nmaxwell
            prob.media = prob.media[:2] + [fill(slw)] + [fill(0.03 + 0.01 * n) for n in range(nmaxwell)] + [fill(0.05 + 0.01 * n) for n in range(nmaxwell)]
            import math
            prob.visco_wl = tuple(float(np.float32(2.0 * math.pi * 10 ** (-1.0 + 2.0 * n / max(nmaxwell - 1, 1)))) for n in range(nmaxwell))
        else:
            raise ValueError(medium)
        if free_top:
            # free-surface matrices by the library itself (cgfd_b200_dvh2dvz takes the device arrays as they are)
            from . import solver
            prob.coords = (x, y, z)
            prob.mats = solver.dvh2dvz(prob, device=dev.index or 0)
            prob.coords = None
    elif free_top:
        k = g["nk2"]
        # the free-surface matrices need the k = nk2 plane only
        met2 = [None] + [m[k:k + 1].cpu().numpy() for m in metric[1:]]
        lam2 = np.full((1, prob.ny, prob.nx), lam, np.float32)
        mu2 = np.full((1, prob.ny, prob.nx), mu, np.float32)
        g2 = dict(g)
        g2["nk2"] = 0
        mvx, mvy, mf = hs.dvh2dvz_iso(met2, lam2, mu2, g2)
        prob.mats = dict(matVx2Vz=mvx, matVy2Vz=mvy, matF2Vz=mf, matD=np.zeros_like(mf))
    del x, y, z
    return prob
