"""Shared helpers of the parity tests."""
import numpy as np

from cgfd3d_b200 import abi, hostsetup as hs

CMP = abi.CMP_NAMES


def small_problem(ni=26, nj=22, nk=20, topo="hill", pml_layers=4, free_top=True, pml_faces=None, src="moment",
                  spatial="point", nt_total=40, timg_mode=abi.TIMG_ZERO, dt=None, seed=None, medium="iso"):
    if pml_faces is None:
        pml_faces = ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0)) if free_top else ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1))
    prob = hs.build_problem(ni, nj, nk, topo=topo, hill=(300.0, 600.0), pml_layers=pml_layers, pml_faces=pml_faces,
                            free_top=free_top, timg_mode=timg_mode, dt=dt, dt_safety=0.9, medium=medium, seed=seed)
    if seed is not None and medium == "iso":
        # heterogeneous medium: +-20 % smooth-free random perturbation keeps the scheme stable for short runs
        rng = np.random.default_rng(seed)
        shape = prob.media[0].shape
        prob.media[0] = (prob.media[0] * (1.0 + 0.2 * rng.uniform(-1, 1, shape))).astype(np.float32)
        prob.media[1] = (prob.media[1] * (1.0 + 0.2 * rng.uniform(-1, 1, shape))).astype(np.float32)
        prob.media[2] = (prob.media[2] * (1.0 + 0.2 * rng.uniform(-1, 1, shape))).astype(np.float32)
        if free_top:
            mvx, mvy, mf = hs.dvh2dvz_iso(prob.metric, prob.media[0], prob.media[1], prob.grid)
            prob.mats = dict(matVx2Vz=mvx, matVy2Vz=mvy, matF2Vz=mf, matD=np.zeros_like(mf))
    if src == "moment":
        hs.make_source(prob, ni // 2, nj // 2 - 1, nk - 1 - 6, nt_total=nt_total, kind="moment",
                       mech=(1e16, 0.7e16, 1.2e16, 0.3e16, -0.2e16, 0.1e16), spatial=spatial, inc=(0.2, -0.3, 0.1),
                       fc=3.0, t0=0.3, stf_len=0.8)
    elif src == "force":
        hs.make_source(prob, ni // 2 + 1, nj // 2, nk - 1 - 5, nt_total=nt_total, kind="force", mech=(1e16, -2e16, 3e16),
                       spatial=spatial, inc=(0.1, 0.2, -0.4), fc=3.0, t0=0.3, stf_len=0.8)
    return prob


def fill_surface_matrices(prob, R):
    """free-surface matrices of a non-isotropic problem from the reference's own *_dvh2dvz (R = ref_flat.RefSolver(prob))"""
    if prob.free_top and prob.medium_type != abi.MEDIUM_ELASTIC_ISO:
        prob.mats = R.dvh2dvz()
    return prob


def random_state(prob, seed=12345, scale_t=1.0e6):
    """wavefield U(-0.5,0.5) (stress scaled so both halves of the RHS matter), zero in the ghosts of
    physical faces like the reference keeps them; random PML aux of matching magnitudes."""
    rng = np.random.default_rng(seed)
    w = np.zeros((prob.ncmp, prob.nz, prob.ny, prob.nx), np.float32)
    ph = (slice(None), slice(3, prob.nz - 3), slice(3, prob.ny - 3), slice(3, prob.nx - 3))
    w[ph] = rng.uniform(-0.5, 0.5, w[ph].shape)
    w[3:9] *= scale_t
    if prob.ncmp > 9:
        w[9:] *= 1.0e-4   # memory variables are strains
    aux = {}
    for key in prob.pml:
        shp = prob.pml_aux_shape(*key)
        a = rng.uniform(-0.5, 0.5, shp).astype(np.float32)
        a[0:3] *= 1e-4
        a[3:9] *= 1e7
        aux[key] = a
    return w, aux


def rel_max(a, b):
    """max|a-b| / max|b| (b = reference); 0/0 -> 0."""
    d = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))
    m = float(np.max(np.abs(b)))
    return 0.0 if d == 0.0 else (d / m if m > 0 else float("inf"))


def rel_l2(a, b):
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    nb = np.linalg.norm(b)
    na = np.linalg.norm(a - b)
    return 0.0 if na == 0.0 else (na / nb if nb > 0 else float("inf"))


def surface_force_problem(spatial="point", nt_total=40, **kw):
    """hill problem with a strict surface force: a point force ON the top row, or a Gaussian one two rows below it whose
    footprint reaches the surface (forward/src_t.c:361-371) -- both go through the traction / velocity slices of
    src_set_surface_layer_for_force (forward/src_t.c:153-314) and the matF2Vz term (forward/sv_curv_col_el_iso.c:583-592)"""
    prob = small_problem(src=None, nt_total=nt_total, **kw)
    sk = prob.nk - 1 if spatial == "point" else prob.nk - 3
    hs.make_source(prob, prob.ni // 2 + 1, prob.nj // 2, sk, nt_total=nt_total, kind="force", mech=(1e12, -2e12, 3e12),
                   spatial=spatial, inc=(0.1, 0.2, -0.3), fc=3.0, t0=0.3, stf_len=0.8)
    assert prob.src["total_number_surface_force"] == 1
    return prob


def sponge_problem(layers=6, mixed=False, **kw):
    """exponential sponge (bdry_ablexp_apply, forward/bdry_t.c:840-890; the example script's default boundary) on the five
    non-free faces; mixed = CFS-PML on the x faces and the sponge on y and bottom"""
    kw.setdefault("pml_faces", ((0, 0), (0, 1)) if mixed else ())
    prob = small_problem(**kw)
    lay = [[0, 0] if mixed else [layers, layers], [layers, layers], [layers, 0]]
    prob.ablexp = hs.ablexp_profiles(*prob.coords, prob.grid, lay, prob.dt)
    return prob
