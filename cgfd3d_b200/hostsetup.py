"""Host-side set-up of one subdomain for the hot path: the pieces of the reference's set-up
pipeline that produce the arrays the time-stepping kernels consume.

The drop-in C driver (integration/drv_rk_curv_col_b200.c) takes these arrays from the reference's
own structs; this module builds the same arrays for synthetic problems (tests, bench.py) with numpy,
following the reference formulas:

  fd_macdrp()              forward/fd_t.c:27-49, 75-114, 182-289   scheme tables
  metric_from_coords()     forward/gd_t.c:190-402                  7-point centred metrics + mirrored ghosts
  pml_profiles()           forward/bdry_t.c:238-268, 333-366, 371-503
  dvh2dvz_iso()            forward/sv_curv_col_el_iso.c:1258-1375  free-surface 3x3 matrices
  estimate_dt()            forward/blk_t.c:2840-2930
  stf tables               forward/src_t.c:984-1174, 2159-2166

Everything is float32, x fastest: arrays have shape [nz][ny][nx] including 3 ghost layers.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import abi

NG = 3  # ghost layers per side (forward/fd_t.c:64-66)
f32 = np.float32

# interior MacCormack/DRP 5-point ops and near-surface variants (forward/fd_t.c:75-114)
MAC_INDX = {0: (-1, 0, 1, 2, 3), 1: (-3, -2, -1, 0, 1)}
MAC_COEF = {0: (-0.30874, -0.6326, 1.233, -0.3334, 0.04168), 1: (-0.04168, 0.3334, -1.233, 0.6326, 0.30874)}
LAY_INDX = {1: {0: (0, 1), 1: (-1, 0)}, 2: {0: (0, 1, 2), 1: (-2, -1, 0)}}
LAY_COEF = {1: {0: (-1.0, 1.0), 1: (-1.0, 1.0)}, 2: {0: (-7.0 / 6.0, 8.0 / 6.0, -1.0 / 6.0), 1: (1.0 / 6.0, -8.0 / 6.0, 7.0 / 6.0)}}
# centred 7-point op used for the metrics (forward/fd_t.c:165-171)
FDC_INDX = (-3, -2, -1, 0, 1, 2, 3)
FDC_COEF = (-0.02084, 0.1667, -0.7709, 0.0, 0.7709, -0.1667, 0.02084)
# 8 direction pairs (forward/fd_t.c:185-195)
FD_FLAGS = ((0, 0, 0), (1, 1, 0), (1, 1, 1), (0, 0, 1), (0, 1, 0), (1, 0, 0), (1, 0, 1), (0, 1, 1))
RK_A = (0.5, 0.5, 1.0, 0.0)
RK_B = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)
RK_RHS_TIME = (0.0, 0.5, 0.5, 1.0)
CFL = 1.3


def stage_dir(ipair: int, istage: int) -> tuple[int, int, int]:
    """Direction index per axis of the operator of [ipair][istage] (forward/fd_t.c:233-236)."""
    return tuple((FD_FLAGS[ipair][a] + istage) % 2 for a in range(3))


def fd_macdrp() -> abi.Fd:
    fd = abi.Fd()
    for s in range(4):
        fd.rk_a[s] = RK_A[s]
        fd.rk_b[s] = RK_B[s]
    for p in range(8):
        for s in range(4):
            d = stage_dir(p, s)
            for a in range(3):
                fd.dir[p][s][a] = d[a]
    for d in (0, 1):
        for n in range(5):
            fd.indx[d][n] = MAC_INDX[d][n]
            fd.coef[d][n] = MAC_COEF[d][n]
        for lay in (1, 2):
            fd.lay_len[lay][d] = len(LAY_INDX[lay][d])
            for n in range(len(LAY_INDX[lay][d])):
                fd.lay_indx[lay][d][n] = LAY_INDX[lay][d][n]
                fd.lay_coef[lay][d][n] = LAY_COEF[lay][d][n]
    return fd


# ---------------------------------------------------------------------------------------------
# grid and metrics
# ---------------------------------------------------------------------------------------------
def cartesian_coords(ni, nj, nk, dh=(100.0, 100.0, 100.0), origin=None):
    """Coordinates incl. ghosts; the top physical row sits at z = 0 unless origin is given
    (forward/gd_t.c:551-583)."""
    nx, ny, nz = ni + 2 * NG, nj + 2 * NG, nk + 2 * NG
    if origin is None:
        origin = (0.0, 0.0, -(nk - 1) * dh[2])
    x1 = (origin[0] + (np.arange(nx) - NG) * dh[0]).astype(f32)
    y1 = (origin[1] + (np.arange(ny) - NG) * dh[1]).astype(f32)
    z1 = (origin[2] + (np.arange(nz) - NG) * dh[2]).astype(f32)
    x = np.broadcast_to(x1[None, None, :], (nz, ny, nx)).copy()
    y = np.broadcast_to(y1[None, :, None], (nz, ny, nx)).copy()
    z = np.broadcast_to(z1[:, None, None], (nz, ny, nx)).copy()
    return x, y, z


def hill_coords(ni, nj, nk, dh=(100.0, 100.0, 100.0), height=1000.0, sigma=2000.0, gi0=0, gj0=0, gni=None, gnj=None):
    """Gaussian-hill topography: z_top(x,y) = H exp(-r^2 / 2 sigma^2) centred on
    the GLOBAL domain of gni x gnj points, of which this subdomain holds [gi0, gi0+ni) x [gj0, gj0+nj).
    Columns are stretched linearly between a flat bottom and the surface, ghosts by the same formula."""
    gni = gni or ni
    gnj = gnj or nj
    nx, ny, nz = ni + 2 * NG, nj + 2 * NG, nk + 2 * NG
    x1 = ((np.arange(nx) - NG + gi0) * dh[0]).astype(np.float64)
    y1 = ((np.arange(ny) - NG + gj0) * dh[1]).astype(np.float64)
    xc = 0.5 * (gni - 1) * dh[0]
    yc = 0.5 * (gnj - 1) * dh[1]
    r2 = (x1[None, :] - xc) ** 2 + (y1[:, None] - yc) ** 2
    ztop = height * np.exp(-r2 / (2.0 * sigma * sigma))
    zbot = -(nk - 1) * dh[2]
    s = ((np.arange(nz) - NG) / float(nk - 1))[:, None, None]
    z = (zbot + s * (ztop[None, :, :] - zbot)).astype(f32)
    x = np.broadcast_to(x1.astype(f32)[None, None, :], (nz, ny, nx)).copy()
    y = np.broadcast_to(y1.astype(f32)[None, :, None], (nz, ny, nx)).copy()
    return x, y, z


def _cdiff(a, axis):
    """7-point centred difference over the physical range, evaluated left to right in float32
    like M_FD_SHIFT (forward/fd_t.h:11-15). Returns an array of the physical shape."""
    nz, ny, nx = a.shape
    sl = [slice(NG, nz - NG), slice(NG, ny - NG), slice(NG, nx - NG)]
    out = None
    for off, c in zip(FDC_INDX, FDC_COEF):
        s = list(sl)
        s[axis] = slice(NG + off, a.shape[axis] - NG + off)
        term = f32(c) * a[tuple(s)]
        out = term if out is None else out + term
    return out


def metric_from_coords(x, y, z):
    """jac, xi_x..zeta_z with mirrored ghosts (forward/gd_t.c:190-402). Returns a list of 10 arrays."""
    nz, ny, nx = x.shape
    d = {}
    for name, a in (("x", x), ("y", y), ("z", z)):
        d[name + "_xi"] = _cdiff(a, 2)
        d[name + "_et"] = _cdiff(a, 1)
        d[name + "_zt"] = _cdiff(a, 0)
    v1 = (d["x_xi"], d["y_xi"], d["z_xi"])
    v2 = (d["x_et"], d["y_et"], d["z_et"])
    v3 = (d["x_zt"], d["y_zt"], d["z_zt"])

    def cross(a, b):
        return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

    g = cross(v1, v2)
    jac = (f32(0.0) + g[0] * v3[0]) + g[1] * v3[1] + g[2] * v3[2]
    xi = cross(v2, v3)
    et = cross(v3, v1)
    zt = cross(v1, v2)
    phys = [jac] + [c / jac for c in xi] + [c / jac for c in et] + [c / jac for c in zt]
    out = []
    for p in phys:
        full = np.zeros((nz, ny, nx), dtype=f32)
        full[NG:nz - NG, NG:ny - NG, NG:nx - NG] = p
        # mirror about the mid-point between the last physical and first ghost point, x then y then z
        for g_ in range(NG):
            full[:, :, NG - 1 - g_] = full[:, :, NG + g_]
            full[:, :, nx - NG + g_] = full[:, :, nx - NG - 1 - g_]
        for g_ in range(NG):
            full[:, NG - 1 - g_, :] = full[:, NG + g_, :]
            full[:, ny - NG + g_, :] = full[:, ny - NG - 1 - g_, :]
        for g_ in range(NG):
            full[NG - 1 - g_, :, :] = full[NG + g_, :, :]
            full[nz - NG + g_, :, :] = full[nz - NG - 1 - g_, :, :]
        out.append(full)
    return out


def estimate_dt(x, y, z, vp_max: float) -> float:
    """dtmax = CFL / Vp * (min distance from a point to the 8 planes through its neighbours)
    (forward/blk_t.c:2840-2930), for a homogeneous Vp."""
    nz, ny, nx = x.shape
    P = np.stack([x, y, z], axis=-1).astype(np.float64)
    c = P[NG:nz - NG, NG:ny - NG, NG:nx - NG]
    lmin = np.inf
    for kk in (-1, 1):
        for jj in (-1, 1):
            for ii in (-1, 1):
                p1 = P[NG:nz - NG, NG:ny - NG, NG - ii:nx - NG - ii]
                p2 = P[NG:nz - NG, NG - jj:ny - NG - jj, NG:nx - NG]
                p3 = P[NG - kk:nz - NG - kk, NG:ny - NG, NG:nx - NG]
                n = np.cross(p2 - p1, p3 - p1)
                L = np.abs(np.einsum("...i,...i", n, c - p1)) / np.sqrt(np.einsum("...i,...i", n, n))
                lmin = min(lmin, float(L.min()))
    return CFL / vp_max * lmin


# ---------------------------------------------------------------------------------------------
# CFS-PML coefficient profiles
# ---------------------------------------------------------------------------------------------
def _abl_len_dh(x, y, z, rng, idim):
    """mean arc length per cell and slab length along idim (forward/bdry_t.c:371-503)."""
    (i1, i2), (j1, j2), (k1, k2) = rng
    X = x[k1:k2 + 1, j1:j2 + 1, i1:i2 + 1].astype(np.float64)
    Y = y[k1:k2 + 1, j1:j2 + 1, i1:i2 + 1].astype(np.float64)
    Z = z[k1:k2 + 1, j1:j2 + 1, i1:i2 + 1].astype(np.float64)
    ax = 2 - idim
    seg = np.sqrt(np.diff(X, axis=ax) ** 2 + np.diff(Y, axis=ax) ** 2 + np.diff(Z, axis=ax) ** 2)
    dh = f32(seg.sum() / seg.size)
    n = (i2 - i1, j2 - j1, k2 - k1)[idim]
    return f32(dh * f32(n)), dh


def pml_profiles(x, y, z, grid, idim, iside, nlay, alpha_max=3.14, beta_max=2.0, ref_vel=7000.0):
    """Transformed A, B, D of one face: D <- d/beta, A <- alpha + d/beta, B <- 1/beta
    (forward/bdry_t.c:238-268, 333-366)."""
    rng = [[grid["ni1"], grid["ni2"]], [grid["nj1"], grid["nj2"]], [grid["nk1"], grid["nk2"]]]
    if iside == 0:
        rng[idim][1] = rng[idim][0] + nlay
    else:
        rng[idim][0] = rng[idim][1] - nlay
    L0, dh = _abl_len_dh(x, y, z, rng, idim)
    npts = nlay + 1
    num_lay = f32(npts - 1)
    Rpp = f32(math.pow(10, -((math.log10(float(num_lay)) - 1.0) / math.log10(2.0) + 4.0)))
    dmax = f32(-float(f32(ref_vel)) / (2.0 * float(L0)) * math.log(float(Rpp)) * (2.0 + 1.0))
    A = np.zeros(npts, f32)
    B = np.zeros(npts, f32)
    D = np.zeros(npts, f32)
    for n in range(npts):
        L = f32(f32(n) * dh)
        i = npts - 1 - n if iside == 0 else n
        xl = float(f32(L / L0))
        d = f32(float(dmax) * math.pow(xl, 2.0))
        a = f32(float(f32(alpha_max)) * (1.0 - math.pow(xl, 1.0)))
        b = f32(1.0 + (float(f32(beta_max)) - 1.0) * math.pow(xl, 2.0))
        D[i] = f32(d / b)
        A[i] = f32(a + D[i])
        B[i] = f32(1.0 / float(b))
    return A, B, D


# ---------------------------------------------------------------------------------------------
# free-surface matrices, isotropic
# ---------------------------------------------------------------------------------------------
def dvh2dvz_iso(metric, lam, mu, grid):
    """matVx2Vz = A^-1 B, matVy2Vz = A^-1 C, matF2Vz = A^-1 at k = nk2
    (forward/sv_curv_col_el_iso.c:1258-1375), stored [(j*nx+i)*9 + row*3 + col]."""
    nz, ny, nx = lam.shape
    k = grid["nk2"]
    e = [[metric[1 + 3 * r + c][k].astype(f32) for c in range(3)] for r in range(3)]  # e[r][c]: r=xi,et,zt
    l = lam[k].astype(f32)
    m = mu[k].astype(f32)
    l2m = l + f32(2.0) * m
    e1, e2, e3 = e[0], e[1], e[2]

    def build(ea, eb, sign):
        # M[r][c] = lam*ea[r]*eb[c] + mu*ea[c]*eb[r]  (+ mu*(ea.eb - ea[r]eb[r]) + (lam+2mu) on the diagonal)
        M = [[None] * 3 for _ in range(3)]
        for r in range(3):
            for c in range(3):
                if r == c:
                    others = [q for q in range(3) if q != r]
                    M[r][c] = l2m * ea[r] * eb[r] + m * (ea[others[0]] * eb[others[0]] + ea[others[1]] * eb[others[1]])
                else:
                    M[r][c] = l * ea[r] * eb[c] + m * ea[c] * eb[r]
                if sign < 0:
                    M[r][c] = -M[r][c]
        return M

    A = build(e3, e3, +1)
    Bm = build(e3, e1, -1)
    Cm = build(e3, e2, -1)
    # adjugate / determinant inverse (lib/fdlib_math.c:7-35)
    inv = [[None] * 3 for _ in range(3)]
    inv[0][0] = A[1][1] * A[2][2] - A[2][1] * A[1][2]
    inv[0][1] = A[2][1] * A[0][2] - A[0][1] * A[2][2]
    inv[0][2] = A[0][1] * A[1][2] - A[0][2] * A[1][1]
    inv[1][0] = A[1][2] * A[2][0] - A[1][0] * A[2][2]
    inv[1][1] = A[0][0] * A[2][2] - A[2][0] * A[0][2]
    inv[1][2] = A[1][0] * A[0][2] - A[0][0] * A[1][2]
    inv[2][0] = A[1][0] * A[2][1] - A[1][1] * A[2][0]
    inv[2][1] = A[2][0] * A[0][1] - A[0][0] * A[2][1]
    inv[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0]
    det = inv[0][0] * A[0][0] + inv[0][1] * A[1][0] + inv[0][2] * A[2][0]
    with np.errstate(divide="ignore", invalid="ignore"):
        rdet = f32(1.0) / det
    Ai = [[inv[r][c] * rdet for c in range(3)] for r in range(3)]

    def matmul(P, Q):
        return [[(f32(0.0) + P[r][0] * Q[0][c]) + P[r][1] * Q[1][c] + P[r][2] * Q[2][c] for c in range(3)] for r in range(3)]

    AB = matmul(Ai, Bm)
    AC = matmul(Ai, Cm)
    outs = []
    mask = np.zeros((ny, nx), bool)
    mask[grid["nj1"]:grid["nj2"] + 1, grid["ni1"]:grid["ni2"] + 1] = True
    for M in (AB, AC, Ai):
        arr = np.zeros((ny, nx, 3, 3), f32)
        for r in range(3):
            for c in range(3):
                arr[:, :, r, c] = np.where(mask, M[r][c], f32(0.0))
        outs.append(arr.reshape(-1).copy())
    return outs  # matVx2Vz, matVy2Vz, matF2Vz


# ---------------------------------------------------------------------------------------------
# source time functions
# ---------------------------------------------------------------------------------------------
def fun_ricker(t, fc, t0):
    """forward/src_t.c:2159-2166 (float in, double inside, float out)."""
    u = f32((float(f32(t)) - float(f32(t0))) * 2.0 * math.pi * float(f32(fc)))
    u = float(u)
    return f32((1 - u * u / 2.0) * math.exp(-u * u / 4.0))


def fun_ricker_deriv(t, fc, t0):
    """forward/src_t.c:2169-2176 (float in, double inside, float out)."""
    u = float(f32((float(f32(t)) - float(f32(t0))) * 2.0 * math.pi * float(f32(fc))))
    return f32(u * (-3 + 1.0 / 2.0 * u * u) * math.exp(-u * u / 4.0) * math.pi * float(f32(fc)))


# ---------------------------------------------------------------------------------------------
# exponential sponge (ablexp)
# ---------------------------------------------------------------------------------------------
def _ablexp_mask(i, vel, dt, num_lay, dh):
    """bdry_ablexp_cal_mask (forward/bdry_t.c:814-832), float32 arithmetic."""
    ln = f32(f32(num_lay) * f32(dh))
    num_step = int(f32(f32(ln / f32(vel)) / f32(dt)))
    total = f32(0.0)
    for n in range(num_step):
        total = f32(total + f32(math.pow(float(f32(f32(f32(n) * f32(dt)) * f32(vel)) / ln), 2.0)))
    alpha = f32(0.6 / float(total))
    return f32(math.exp(float(f32(-alpha * f32(math.pow(float(f32(f32(i) / f32(num_lay))), 2.0))))))


def ablexp_profiles(x, y, z, grid, layers, dt, vel=7000.0):
    """The six shell blocks and the 1-D damping profiles of bdry_ablexp_set (forward/bdry_t.c:513-796).
    layers[idim][iside] = number of sponge layers on that face (0 = none; inter-rank faces must be 0).
    Returns (blk[6][7] int32 = enable, ni1, ni2, nj1, nj2, nk1, nk2; Ex[nx]; Ey[ny]; Ez[nz])."""
    g = grid
    ni1, ni2, nj1, nj2, nk1, nk2 = g["ni1"], g["ni2"], g["nj1"], g["nj2"], g["nk1"], g["nk2"]
    ni, nj, nk = ni2 - ni1 + 1, nj2 - nj1 + 1, nk2 - nk1 + 1
    E = [np.ones(g["nx"], f32), np.ones(g["ny"], f32), np.ones(g["nz"], f32)]
    blk = np.zeros((6, 7), np.int32)
    blk[:, 2::2] = -1
    L = layers
    nix, njy = L[0][0] + L[0][1], L[1][0] + L[1][1]
    # (count, first) per axis of the six blocks x1 x2 y1 y2 z1 z2: y blocks exclude the x shells, z blocks the x and y shells
    spec = [
        ((L[0][0], ni1), (nj, nj1), (nk, nk1)),
        ((L[0][1], ni2 - L[0][1] + 1), (nj, nj1), (nk, nk1)),
        ((ni - nix, ni1 + L[0][0]), (L[1][0], nj1), (nk, nk1)),
        ((ni - nix, ni1 + L[0][0]), (L[1][1], nj2 - L[1][1] + 1), (nk, nk1)),
        ((ni - nix, ni1 + L[0][0]), (nj - njy, nj1 + L[1][0]), (L[2][0], nk1)),
        ((ni - nix, ni1 + L[0][0]), (nj - njy, nj1 + L[1][0]), (L[2][1], nk2 - L[2][1] + 1)),
    ]
    for n, sp in enumerate(spec):
        idim, iside = n // 2, n % 2
        rng = [[a1, a1 + cnt - 1] for (cnt, a1) in sp]
        blk[n, 1:] = [rng[0][0], rng[0][1], rng[1][0], rng[1][1], rng[2][0], rng[2][1]]
        if min(cnt for cnt, _ in sp) <= 0:
            continue
        blk[n, 0] = 1
        _, dh = _abl_len_dh(x, y, z, rng, idim)
        cnt, a1 = sp[idim]
        for a in range(a1, a1 + cnt):
            # the first point of the layer gets the weakest damping on the "2" sides, the strongest on the "1" sides
            i = cnt - (a - a1) if iside == 0 else a - a1 + 1
            E[idim][a] = _ablexp_mask(i, vel, dt, cnt, dh)
    return blk, E[0], E[1], E[2]


@dataclass
class HostProblem:
    """numpy arrays of one subdomain + conversion to the C ABI struct."""
    ni: int
    nj: int
    nk: int
    dt: float
    medium_type: int = abi.MEDIUM_ELASTIC_ISO
    nmaxwell: int = 0
    metric: list = field(default_factory=list)
    media: list = field(default_factory=list)
    visco_wl: tuple = ()
    free_top: int = 1
    timg_mode: int = abi.TIMG_ZERO
    pml: dict = field(default_factory=dict)   # (idim, iside) -> (nlay, A, B, D)
    mats: dict = field(default_factory=dict)  # matVx2Vz, matVy2Vz, matF2Vz, matD
    src: dict = field(default_factory=dict)
    neigh: tuple = (-1, -1, -1, -1)
    coords: tuple | None = None
    graves_Qs: np.ndarray | None = None   # Qs [nz][ny][nx]: Graves' attenuation of an elastic medium (None = off)
    graves_Qs_freq: float = 1.0
    ablexp: tuple | None = None   # (blk[6][7], Ex, Ey, Ez) of ablexp_profiles(), None = no sponge
    _keep: list = field(default_factory=list, repr=False)

    @property
    def nx(self):
        return self.ni + 2 * NG

    @property
    def ny(self):
        return self.nj + 2 * NG

    @property
    def nz(self):
        return self.nk + 2 * NG

    @property
    def ncmp(self):
        return 9 + 6 * self.nmaxwell

    @property
    def nvol(self):
        return self.nx * self.ny * self.nz

    @property
    def grid(self):
        return dict(nx=self.nx, ny=self.ny, nz=self.nz, ni1=NG, ni2=NG + self.ni - 1, nj1=NG, nj2=NG + self.nj - 1,
                    nk1=NG, nk2=NG + self.nk - 1)

    def iptr(self, i, j, k):
        """flat index of LOCAL physical point (i,j,k), 0-based without ghosts."""
        return (i + NG) + (j + NG) * self.nx + (k + NG) * self.nx * self.ny

    def pml_slab(self, idim, iside):
        g = self.grid
        r = [[g["ni1"], g["ni2"]], [g["nj1"], g["nj2"]], [g["nk1"], g["nk2"]]]
        nlay = self.pml[(idim, iside)][0] if (idim, iside) in self.pml else 0
        if iside == 0:
            r[idim][1] = r[idim][0] + nlay
        else:
            r[idim][0] = r[idim][1] - nlay
        return r

    def pml_aux_shape(self, idim, iside):
        r = self.pml_slab(idim, iside)
        return (9, r[2][1] - r[2][0] + 1, r[1][1] - r[1][0] + 1, r[0][1] - r[0][0] + 1)

    def to_c(self) -> abi.Problem:
        p = abi.Problem()
        keep = self._keep
        p.abi_version = abi.ABI_VERSION
        for k_, v in self.grid.items():
            setattr(p.grid, k_, v)
        p.fd = fd_macdrp()
        p.dt = self.dt
        p.medium_type = self.medium_type
        p.nmaxwell = self.nmaxwell
        p.ncmp = self.ncmp
        assert len(self.metric) == 10
        for m in range(10):
            p.metric[m] = abi.as_f(self.metric[m])
        p.nmedia = len(self.media)
        for m, a in enumerate(self.media):
            p.media[m] = abi.as_f(a)
        for n, w in enumerate(self.visco_wl):
            p.visco_wl[n] = w
        p.free_top = self.free_top
        p.timg_mode = self.timg_mode
        for (idim, iside), (nlay, A, B, D) in self.pml.items():
            f = p.pml[idim][iside]
            f.enabled = 1
            f.nlay = nlay
            f.A, f.B, f.D = abi.as_f(A), abi.as_f(B), abi.as_f(D)
        for name in ("matVx2Vz", "matVy2Vz", "matF2Vz", "matD"):
            setattr(p, name, abi.as_f(self.mats.get(name)))
        s = self.src
        if s and s.get("total_number", 0) > 0:
            c = p.src
            for name in ("total_number", "max_nt", "max_stage", "is_surface_force_strict", "total_number_surface_force",
                         "itype_spatial_ext", "ext_half_npoint", "force_actived", "moment_actived"):
                setattr(c, name, int(s.get(name, 0)))
            c.ext_func_coef = float(s.get("ext_func_coef", 1.5))
            for name in ("si", "sj", "sk", "it_begin", "it_end", "force_rate_indx"):
                setattr(c, name, abi.as_i(s.get(name)))
            for name in ("si_inc", "sj_inc", "sk_inc", "Fx", "Fy", "Fz", "Mxx", "Myy", "Mzz", "Mxz", "Myz", "Mxy",
                         "Fx_rate", "Fy_rate", "Fz_rate"):
                setattr(c, name, abi.as_f(s.get(name)))
        else:
            p.src.itype_spatial_ext = abi.SRC_SPATIAL_POINT
            p.src.max_stage = 4
        for n in range(4):
            p.neigh[n] = self.neigh[n]
        if self.ablexp is not None:
            blk, Ex, Ey, Ez = self.ablexp
            p.ablexp_enabled = 1
            for n in range(6):
                for q in range(7):
                    p.ablexp_blk[n][q] = int(blk[n][q])
            p.ablexp_Ex, p.ablexp_Ey, p.ablexp_Ez = abi.as_f(Ex), abi.as_f(Ey), abi.as_f(Ez)
        p.graves_Qs = abi.as_f(self.graves_Qs)
        p.graves_Qs_freq = float(self.graves_Qs_freq)
        keep.append(p)
        return p


def make_source(prob: HostProblem, si, sj, sk, *, nt_total, kind="moment", mech=(1e16, 1e16, 1e16, 0, 0, 0),
                fc=2.0, t0=0.5, stf_len=1.0, spatial="point", inc=(0.0, 0.0, 0.0), t_start=0.0, strict=1):
    """One point source by LOCAL physical index (0-based), Ricker STF; tables as forward/src_t.c:984-1174.
    mech = (Mxx,Myy,Mzz,Myz,Mxz,Mxy) for kind='moment' (the .src file order), (Fx,Fy,Fz) for 'force'.
    A force whose footprint reaches the free surface (sk + half extent >= top row; half extent 0 for a point source, 3 for a
    Gaussian one, forward/src_t.c:361-371, 416-421) is a SURFACE force when strict = 1 (par source_surface_force_strict): it gets
    a rate table Fx_rate.. = d(stf)/dt * F (forward/src_t.c:1083-1090) and is applied through the traction / velocity slices of
    src_set_surface_layer_for_force (forward/src_t.c:153-314)."""
    dt = f32(prob.dt)
    it_begin = int(t_start / float(dt))
    max_nt = int(float(f32(stf_len)) / float(dt) + 0.5)  # forward/src_t.c:550-552
    it_end = it_begin + max_nt - 1                        # forward/src_t.c:1035-1036
    max_stage = 4
    tab = {n: np.zeros((1, max_nt, max_stage), f32) for n in ("Fx", "Fy", "Fz", "Mxx", "Myy", "Mzz", "Mxz", "Myz", "Mxy")}
    rate = {n: np.zeros((1, max_nt, max_stage), f32) for n in ("Fx_rate", "Fy_rate", "Fz_rate")}
    t_shift = f32(f32(t_start) - f32(f32(it_begin) * dt + f32(0.0)))
    half = 3 if spatial == "gauss" else 0
    surf = bool(kind == "force" and strict and prob.free_top and sk + half >= prob.nk - 1)
    for it in range(max_nt):
        for st in range(max_stage):
            t = f32(f32(f32(it) * dt + f32(RK_RHS_TIME[st]) * dt) - t_shift)
            v = fun_ricker(t, fc, t0)
            if kind == "moment":
                mxx, myy, mzz, myz, mxz, mxy = [f32(q) for q in mech]
                tab["Mxx"][0, it, st] = v * mxx
                tab["Myy"][0, it, st] = v * myy
                tab["Mzz"][0, it, st] = v * mzz
                tab["Myz"][0, it, st] = v * myz
                tab["Mxz"][0, it, st] = v * mxz
                tab["Mxy"][0, it, st] = v * mxy
            else:
                fx, fy, fz = [f32(q) for q in mech[:3]]
                tab["Fx"][0, it, st] = v * fx
                tab["Fy"][0, it, st] = v * fy
                tab["Fz"][0, it, st] = v * fz
                if surf:
                    r = fun_ricker_deriv(t, fc, t0)
                    rate["Fx_rate"][0, it, st] = r * fx
                    rate["Fy_rate"][0, it, st] = r * fy
                    rate["Fz_rate"][0, it, st] = r * fz
    src = dict(total_number=1, max_nt=max_nt, max_stage=max_stage,
               si=np.array([si + NG], np.int32), sj=np.array([sj + NG], np.int32), sk=np.array([sk + NG], np.int32),
               si_inc=np.array([inc[0]], f32), sj_inc=np.array([inc[1]], f32), sk_inc=np.array([inc[2]], f32),
               it_begin=np.array([it_begin], np.int32), it_end=np.array([it_end], np.int32),
               is_surface_force_strict=int(strict), total_number_surface_force=1 if surf else 0, force_rate_indx=np.zeros(1, np.int32),
               itype_spatial_ext=abi.SRC_SPATIAL_GAUSSIAN if spatial == "gauss" else abi.SRC_SPATIAL_POINT,
               ext_half_npoint=3, ext_func_coef=1.5,
               force_actived=1 if kind == "force" else 0, moment_actived=1 if kind == "moment" else 0)
    src.update(tab)
    src.update(rate)
    prob.src = src
    return src


def build_problem(ni, nj, nk, *, dh=(100.0, 100.0, 100.0), topo="flat", hill=(1000.0, 2000.0),
                  vp=3000.0, vs=2000.0, rho=1500.0, pml_layers=10,
                  pml_faces=((0, 0), (0, 1), (1, 0), (1, 1), (2, 0)), free_top=True, dt=None, dt_safety=1.0,
                  timg_mode=abi.TIMG_ZERO, sub=None, medium="iso", nmaxwell=3, seed=None) -> HostProblem:
    """Homogeneous isotropic half-space (medium 'code' of forward/md_t.c:456-463: Vp 3000, Vs 2000, rho 1500)
    on a Cartesian or Gaussian-hill grid with CFS-PML + free top.
    sub = (gi0, gj0, gni, gnj, neigh) places this block inside a global x-y decomposition."""
    if sub is None:
        gi0 = gj0 = 0
        gni, gnj = ni, nj
        neigh = (-1, -1, -1, -1)
    else:
        gi0, gj0, gni, gnj, neigh = sub
    # Inter-rank faces: the ghost metrics there must be the neighbour's own values (gd_curv_metric_exchange,
    # forward/gd_t.c:407-475), not mirrored ones. The coordinates are analytic, so build a block that is 3 points
    # wider on every side that has a neighbour, compute the metrics on it and crop: the cropped ghosts then hold
    # exactly what the neighbour computes for its physical points; physical faces keep the mirrored ghosts.
    ex = [NG if neigh[n] >= 0 else 0 for n in range(4)]
    eni, enj = ni + ex[0] + ex[1], nj + ex[2] + ex[3]
    egi0, egj0 = gi0 - ex[0], gj0 - ex[2]
    if topo == "flat":
        x, y, z = cartesian_coords(eni, enj, nk, dh, origin=(egi0 * dh[0], egj0 * dh[1], -(nk - 1) * dh[2]))
    else:
        x, y, z = hill_coords(eni, enj, nk, dh, hill[0], hill[1], egi0, egj0, gni, gnj)
    metric = metric_from_coords(x, y, z)
    if any(ex):
        cs = (slice(None), slice(ex[2], ex[2] + nj + 2 * NG), slice(ex[0], ex[0] + ni + 2 * NG))
        x, y, z = (np.ascontiguousarray(a[cs]) for a in (x, y, z))
        metric = [np.ascontiguousarray(m[cs]) for m in metric]
    vp_max = vp if medium in ("iso", "visco") else math.sqrt(25.2e9 * 1.1 / rho)
    if dt is None:
        dt = estimate_dt(x, y, z, vp_max) * dt_safety
    prob = HostProblem(ni=ni, nj=nj, nk=nk, dt=float(f32(dt)), free_top=1 if free_top else 0, timg_mode=timg_mode,
                       neigh=tuple(neigh), coords=(x, y, z))
    prob.metric = metric
    shape = (prob.nz, prob.ny, prob.nx)
    mu = f32(rho * vs * vs)
    lam = f32(rho * vp * vp - 2.0 * rho * vs * vs)
    prob.media = [np.full(shape, lam, f32), np.full(shape, mu, f32), np.full(shape, f32(1.0) / f32(rho), f32)]
    if medium != "iso":
        set_test_medium(prob, medium, rho=rho, nmaxwell=nmaxwell, seed=seed)
    g = prob.grid
    for (idim, iside) in pml_faces:
        if idim < 2 and neigh[idim * 2 + iside] >= 0:
            continue  # inter-rank face: no PML (forward/bdry_t.c:154-158)
        A, B, D = pml_profiles(x, y, z, g, idim, iside, pml_layers)
        prob.pml[(idim, iside)] = (pml_layers, A, B, D)
    if free_top and medium == "iso":
        mvx, mvy, mf = dvh2dvz_iso(metric, prob.media[0], prob.media[1], g)
        prob.mats = dict(matVx2Vz=mvx, matVy2Vz=mvy, matF2Vz=mf, matD=np.zeros_like(mf))
    elif free_top:
        # vti / aniso / visco: the 3x3 surface matrices come from the reference's own one-shot *_dvh2dvz set-up code
        # (kept as host code by the drop-in driver); callers fill prob.mats from it
        z9 = np.zeros(prob.nx * prob.ny * 9, f32)
        prob.mats = dict(matVx2Vz=z9.copy(), matVy2Vz=z9.copy(), matF2Vz=z9.copy(), matD=z9.copy())
    return prob


def set_test_medium(prob: HostProblem, medium: str, rho=1500.0, nmaxwell=3, seed=None):
    """Synthetic media of the other three constitutive laws (values after forward/md_t.c:501-595, 912-952), optionally with
    a +-10 % point-wise random perturbation so that every array matters."""
    shape = (prob.nz, prob.ny, prob.nx)
    rng = np.random.default_rng(seed if seed is not None else 0)

    def fld(v, amp=0.1):
        a = np.full(shape, v, np.float64)
        if seed is not None:
            a *= 1.0 + amp * rng.uniform(-1, 1, shape)
        return a.astype(f32)

    slw = fld(1.0 / rho)
    if medium == "vti":
        prob.medium_type = abi.MEDIUM_ELASTIC_VTI
        prob.media = [fld(25.2e9), fld(10.962e9), fld(18.0e9), fld(5.12e9), fld(7.168e9), slw]   # c11 c13 c33 c55 c66 1/rho
    elif medium == "aniso":
        prob.medium_type = abi.MEDIUM_ELASTIC_ANISO
        c11, c13, c33, c55, c66 = 25.2e9, 10.962e9, 18.0e9, 5.12e9, 7.168e9
        c12 = c11 - 2 * c66
        # upper triangle row by row; the off-diagonal couplings a VTI medium lacks get small non-zero values
        e = 0.15e9
        C = [c11, c12, c13, e, -e, 0.5 * e,
             c11, c13, -0.5 * e, e, 0.7 * e,
             c33, 0.3 * e, -0.6 * e, e,
             c55, 0.4 * e, -0.2 * e,
             c55, 0.8 * e,
             c66]
        prob.media = [fld(v) for v in C] + [slw]
    elif medium == "visco":
        prob.medium_type = abi.MEDIUM_VISCOELASTIC_ISO
        prob.nmaxwell = nmaxwell
        lam, mu = prob.media[0], prob.media[1]
        if seed is not None:
            lam = (lam * (1.0 + 0.1 * rng.uniform(-1, 1, shape))).astype(f32)
            mu = (mu * (1.0 + 0.1 * rng.uniform(-1, 1, shape))).astype(f32)
        ylam = [fld(0.03 + 0.01 * n) for n in range(nmaxwell)]
        ymu = [fld(0.05 + 0.01 * n) for n in range(nmaxwell)]
        prob.media = [lam, mu, slw] + ylam + ymu
        # relaxation frequencies log-spaced over 0.1 .. 10 Hz (md_vis_GMB_cal_Y, forward/md_t.c:642-792)
        prob.visco_wl = tuple(float(f32(2.0 * math.pi * 10 ** (-1.0 + 2.0 * n / max(nmaxwell - 1, 1)))) for n in range(nmaxwell))
    else:
        raise ValueError(medium)
    return prob
