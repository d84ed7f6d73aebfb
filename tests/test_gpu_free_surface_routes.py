"""The two routes of the free-surface rows (rhs_timg_z2 / rhs_vlow_z2, forward/sv_curv_col_el.c:30-305, sv_curv_col_el_iso.c:451-634):
planes of the interior kernel's top z chunk (TOPK launch of k_main_tma: the default for the isotropic medium) and the separate
k_top launch (the default for VTI / general anisotropic / visco-elastic). Every other GPU test runs a medium's DEFAULT route; the tests
here run the OTHER one (CGFD_FUSE_TOP is read when a context is created), against the same unmodified reference functions and with
the same tolerances, and check that the library reports the route it runs.
"""
import os

import numpy as np
import pytest

from cgfd3d_b200 import abi, solver
from oracle import ref_flat
from tests import util
from tests.test_gpu_media import _check_stage as check_stage_media
from tests.test_gpu_iso import _check_stage as check_stage_iso

pytestmark = pytest.mark.gpu
TOL_RUN = 1e-4
DEFAULT_FUSED = {"iso": 1, "vti": 0, "aniso": 0, "visco": 0}


def _need():
    if not ref_flat.available():
        pytest.fail("oracle/_ref/libcgfd_ref_flat.so is missing (build it with make -C oracle ref where /root/reference exists)")
    if solver.device_count() < 1:
        pytest.fail("no CUDA device: the hot path has no CPU fallback")


@pytest.fixture
def other_route(request):
    medium = request.param
    old = os.environ.get("CGFD_FUSE_TOP")
    os.environ["CGFD_FUSE_TOP"] = str(1 - DEFAULT_FUSED[medium])
    yield medium
    if old is None:
        del os.environ["CGFD_FUSE_TOP"]
    else:
        os.environ["CGFD_FUSE_TOP"] = old


@pytest.mark.parametrize("medium", list(DEFAULT_FUSED))
def test_default_route_is_reported(medium):
    _need()
    assert "CGFD_FUSE_TOP" not in os.environ
    prob = util.small_problem(seed=2, medium=medium)
    G = solver.Solver(prob)
    assert G.top_fused() == DEFAULT_FUSED[medium]
    G.close()
    prob = util.small_problem(seed=2, medium=medium, free_top=False)
    G = solver.Solver(prob)
    assert G.top_fused() == 0
    G.close()


@pytest.mark.parametrize("other_route", list(DEFAULT_FUSED), indirect=True)
@pytest.mark.parametrize("ipair", [0, 3, 5, 6])
def test_onestage_other_route(other_route, ipair):
    """one RHS evaluation, hill topography, PML on 5 faces, free surface: stages 0 and 1 of 4 pairs = all 8 direction kernels"""
    _need()
    medium = other_route
    prob = util.small_problem(seed=11, medium=medium)
    G = solver.Solver(prob)
    assert G.top_fused() == 1 - DEFAULT_FUSED[medium]
    G.close()
    for istage in (0, 1):
        if medium == "iso":
            check_stage_iso(prob, it=3, ipair=ipair, istage=istage, seed=400 + ipair)
        else:
            check_stage_media(prob, it=3, ipair=ipair, istage=istage, seed=400 + ipair)


@pytest.mark.parametrize("other_route", list(DEFAULT_FUSED), indirect=True)
@pytest.mark.parametrize("case", ["gauss_force", "mirror", "thin"])
def test_run_other_route(other_route, case):
    """40 RK4 steps from rest against drv_rk_curv_col_allstep: Gaussian force whose footprint reaches the surface rows, the MIRROR
    traction image (isotropic: against the MIRROR copy of the reference), a grid with fewer rows than one z chunk"""
    _need()
    medium = other_route
    nt = 40
    kw = dict(ni=40, nj=36, nk=30, pml_layers=8, nt_total=nt, medium=medium)
    if case == "gauss_force":
        kw.update(src="force", spatial="gauss")
    elif case == "mirror":
        if medium != "iso":
            pytest.skip("the MIRROR copy of the reference is pinned for the isotropic medium")
        if not ref_flat.available(mirror=True):
            pytest.fail("oracle/_ref/libcgfd_ref_flat_mirror.so is missing")
        kw.update(timg_mode=abi.TIMG_MIRROR)
    else:
        kw.update(ni=37, nj=19, nk=11, pml_layers=3)
    prob = util.small_problem(**kw)
    rec = [prob.iptr(10 + 2 * n, 5 + 2 * n, prob.nk - 1) for n in range(4)]
    R = ref_flat.RefSolver(prob)   # picks the MIRROR copy of the reference from prob.timg_mode
    if medium != "iso":
        util.fill_surface_matrices(prob, R)
    wr, recr, _ = R.run(nt, rec_iptr=rec)
    G = solver.Solver(prob)
    assert G.top_fused() == 1 - DEFAULT_FUSED[medium]
    G.set_record_points(rec, nt)
    G.run(nt)
    wg = G.get_wavefield()
    recg = G.get_record(0, nt)
    bad = []
    for c in range(prob.ncmp):
        e = util.rel_l2(wg[c], wr[c])
        if not e <= TOL_RUN:
            bad.append(("w%d" % c, e))
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
