"""GPU parity of the VTI, general anisotropic and visco-elastic (GMB) hot paths against the UNMODIFIED reference functions
(sv_curv_col_el_vti_onestage, sv_curv_col_el_aniso_onestage, sv_curv_col_vis_iso_onestage and drv_rk_curv_col_allstep of
oracle/_ref/libcgfd_ref_flat.so), through the C ABI. The 3x3 free-surface matrices are produced by the reference's own
one-shot *_dvh2dvz set-up code (host code in the drop-in driver as well).

Tolerances (float32): one RHS evaluation max|gpu-ref| <= 2e-5 max|ref| per component -- the anisotropic Hooke law is
evaluated through the physical velocity gradient (63 products) instead of the reference's term-by-term expansion (216),
same value up to re-association; N-step runs relative L2 <= 1e-4 (BASELINE.json north_star).
"""
import numpy as np
import pytest

from cgfd3d_b200 import solver
from oracle import ref_flat
from tests import util

pytestmark = pytest.mark.gpu
TOL_STAGE = 2e-5
TOL_RUN = 1e-4
MEDIA = ["vti", "aniso", "visco"]


def _need():
    if not ref_flat.available():
        pytest.fail("oracle/_ref/libcgfd_ref_flat.so is missing (build it with make -C oracle ref where /root/reference exists)")
    if solver.device_count() < 1:
        pytest.fail("no CUDA device: the hot path has no CPU fallback")


def _names(ncmp):
    return util.CMP + ["J%d" % n for n in range(ncmp - 9)]


def _check_stage(prob, it, ipair, istage, seed):
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    G = solver.Solver(prob)
    w, aux = util.random_state(prob, seed)
    for key, a in aux.items():
        R.set_pml_aux(key[0], key[1], a.ravel())
        G.set_pml_aux(key[0], key[1], a.ravel())
    rr = R.onestage(it, ipair, istage, w)
    rg = G.onestage(it, ipair, istage, w)
    names = _names(prob.ncmp)
    bad = []
    for c in range(prob.ncmp):
        e = util.rel_max(rg[c], rr[c])
        if not e <= TOL_STAGE:
            bad.append((names[c], e))
    for key in aux:
        ar = R.get_pml_aux_rhs(*key).reshape(9, -1)
        ag = G.get_pml_aux_rhs(*key).reshape(9, -1)
        for c in range(9):
            e = util.rel_max(ag[c], ar[c])
            if not e <= TOL_STAGE:
                bad.append(("aux%s.%s" % (key, util.CMP[c]), e))
    G.close()
    assert float(np.abs(rr[3]).max()) > 0
    assert not bad, "ipair=%d istage=%d: %s" % (ipair, istage, bad)


@pytest.mark.parametrize("medium", MEDIA)
@pytest.mark.parametrize("ipair", [0, 2, 5, 7])
def test_onestage_media(medium, ipair):
    """hill topography, PML on 5 faces, free surface, heterogeneous medium; stages 0 and 1 of 4 pairs = all 8 direction kernels"""
    _need()
    prob = util.small_problem(seed=11, medium=medium)
    for istage in (0, 1):
        _check_stage(prob, it=3, ipair=ipair, istage=istage, seed=300 + ipair)


@pytest.mark.parametrize("medium", MEDIA)
@pytest.mark.parametrize("case", ["pml6", "gauss_force"])
def test_onestage_media_variants(medium, case):
    _need()
    kw = dict(seed=5, medium=medium)
    if case == "pml6":
        kw.update(free_top=False, ni=37, nj=19, nk=23, pml_layers=5)
    else:
        kw.update(src="force", spatial="gauss")
    prob = util.small_problem(**kw)
    _check_stage(prob, it=2, ipair=1, istage=2, seed=8)
    _check_stage(prob, it=2, ipair=6, istage=3, seed=9)


@pytest.mark.parametrize("medium", MEDIA)
def test_run_media_matches_reference_driver(medium):
    """60 RK4 steps from rest with a moment source: wavefield (incl. memory variables), PML aux and surface traces against
    drv_rk_curv_col_allstep run on the same flat arrays"""
    _need()
    nt = 60
    prob = util.small_problem(ni=40, nj=36, nk=30, pml_layers=8, nt_total=nt, medium=medium)
    rec = [prob.iptr(10 + 5 * n, 12 + 3 * n, prob.nk - 1) for n in range(5)] + [prob.iptr(20, 17, prob.nk - 8)]
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    wr, recr, _ = R.run(nt, rec_iptr=rec)
    G = solver.Solver(prob)
    G.set_record_points(rec, nt)
    G.run(nt)
    wg = G.get_wavefield()
    recg = G.get_record(0, nt)
    names = _names(prob.ncmp)
    bad = []
    for c in range(prob.ncmp):
        e = util.rel_l2(wg[c], wr[c])
        if not e <= TOL_RUN:
            bad.append(("w." + names[c], e))
    for c in range(3):
        for ip in range(len(rec)):
            e = util.rel_l2(recg[:, c, ip], recr[:, c, ip])
            if not e <= TOL_RUN:
                bad.append(("rec%d.%s" % (ip, util.CMP[c]), e))
    for key in prob.pml:
        e = util.rel_l2(G.get_pml_aux(*key), R.get_pml_aux(*key))
        if not e <= TOL_RUN:
            bad.append(("aux%s" % (key,), e))
    assert np.isfinite(wg).all()
    assert float(np.abs(wr[0]).max()) > 0
    G.close()
    assert not bad, bad


@pytest.mark.parametrize("medium", ["iso", "vti"])
def test_graves_qs_attenuation(medium):
    """Graves' constant-Q attenuation (md->visco_type = graves_Qs): w_end *= exp(-pi f0 dt / Qs) after the last stage
    (forward/sv_curv_col_el.c:638-666), fused here into the last stage's epilogue. 60 steps against the reference driver with a
    heterogeneous Qs in [15, 60]; the run must also differ visibly from the unattenuated one."""
    _need()
    nt = 60
    prob = util.small_problem(ni=42, nj=35, nk=30, pml_layers=6, nt_total=nt, medium=medium, seed=4)
    rng = np.random.default_rng(9)
    prob.graves_Qs = rng.uniform(15.0, 60.0, (prob.nz, prob.ny, prob.nx)).astype(np.float32)
    prob.graves_Qs_freq = 2.5
    rec = [prob.iptr(10 + 5 * n, 12 + 3 * n, prob.nk - 1) for n in range(4)]
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    wr, recr, _ = R.run(nt, rec_iptr=rec)
    G = solver.Solver(prob)
    G.set_record_points(rec, nt)
    G.run(nt)
    wg = G.get_wavefield()
    recg = G.get_record(0, nt)
    G.close()
    bad = [(util.CMP[c], util.rel_l2(wg[c], wr[c])) for c in range(9) if not util.rel_l2(wg[c], wr[c]) <= TOL_RUN]
    for c in range(3):
        for ip in range(len(rec)):
            e = util.rel_l2(recg[:, c, ip], recr[:, c, ip])
            if not e <= TOL_RUN:
                bad.append(("rec%d.%s" % (ip, util.CMP[c]), e))
    assert float(np.abs(wr[0]).max()) > 0 and not bad, bad
    prob.graves_Qs = None
    G = solver.Solver(prob)
    G.run(nt)
    w0 = G.get_wavefield()
    G.close()
    assert util.rel_l2(wg[2], w0[2]) > 1e-2   # the attenuation is really applied


@pytest.mark.parametrize("case", ["mij", "vi+mij"])
def test_distributed_sources(case):
    """finite-fault style dd sources (sv_curv_col_el_rhs_srcdd, forward/sv_curv_col_el.c:486-632): 40 points on a dipping plane,
    time functions delivered in blocks of 7 steps the way src_dd_accit_loadstf reads them; 30 steps against the reference driver
    fed from files with the same tables. No other source is present."""
    _need()
    nt, nb = 30, 7
    prob = util.small_problem(ni=44, nj=34, nk=30, pml_layers=6, nt_total=nt, src=None, seed=8)
    rng = np.random.default_rng(3)
    pts = [(12 + q % 10, 10 + (q // 10) * 3, prob.nk - 6 - (q % 10)) for q in range(40)]
    indx = np.array([prob.iptr(*p) for p in pts], np.int64)
    t = (np.arange(nt)[:, None] + np.array([0.0, 0.5, 0.5, 1.0])[None, :]) * prob.dt          # stage times
    stf = np.exp(-((t - 0.08) / 0.03) ** 2)[:, :, None, None] * (1.0 + 0.3 * rng.uniform(-1, 1, (1, 1, len(pts), 1)))
    mij = (stf * 1e15 * rng.uniform(-1, 1, (1, 1, len(pts), 6))).astype(np.float32)
    vi = (stf * 1e9 * rng.uniform(-1, 1, (1, 1, len(pts), 3))).astype(np.float32) if case == "vi+mij" else None
    R = ref_flat.RefSolver(prob)
    R.set_dd(indx, vi, mij, nb)
    wr, _, _ = R.run(nt)
    G = solver.Solver(prob)
    G.dd_set_points(indx, vi is not None, True, nb)
    for it0 in range(0, nt, nb):
        n = min(nb, nt - it0)
        G.dd_load_block(it0, None if vi is None else vi[it0:it0 + n], mij[it0:it0 + n])
        G.run(n, it0=it0)
    wg = G.get_wavefield()
    G.close()
    assert float(np.abs(wr[0]).max()) > 0 and np.isfinite(wg).all()
    bad = [(util.CMP[c], util.rel_l2(wg[c], wr[c])) for c in range(9) if not util.rel_l2(wg[c], wr[c]) <= TOL_RUN]
    assert not bad, bad


@pytest.mark.parametrize("medium", ["iso", "vti", "aniso", "visco"])
def test_dvh2dvz_on_device(medium):
    """cgfd_b200_dvh2dvz = the reference's *_dvh2dvz of the four constitutive laws (iso.c:1258-1375, vti.c:1085-1205,
    aniso.c:1267-1422, vis_iso.c:353-507) on the GPU: bit-identical matrices on a hill grid with heterogeneous media and
    perturbed x-y grid lines (every metric array matters)."""
    _need()
    from cgfd3d_b200 import hostsetup as hs
    prob = util.small_problem(ni=37, nj=29, nk=23, seed=4, medium=medium)
    x, y, z = (a.copy() for a in prob.coords)
    rng = np.random.default_rng(2)
    x += rng.uniform(-8, 8, x.shape).astype(np.float32)
    y += rng.uniform(-8, 8, y.shape).astype(np.float32)
    prob.coords = (x, y, z)
    prob.metric = hs.metric_from_coords(x, y, z)
    ref = ref_flat.RefSolver(prob).dvh2dvz()
    got = solver.dvh2dvz(prob)
    assert float(np.abs(ref["matVx2Vz"]).max()) > 0
    names = ["matVx2Vz", "matVy2Vz"] + (["matF2Vz"] if medium == "iso" else []) + (["matD"] if medium == "visco" else [])
    for k in names:
        assert float(np.abs(ref[k]).max()) > 0, k
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
