"""GPU parity of the isotropic hot path against the UNMODIFIED reference functions
(oracle/_ref/libcgfd_ref_flat.so = reference sources + the zero-guarded traction image), through the C ABI.

Tolerances (float32; differences come from FMA contraction and the order the source term enters the RK axpy):
  one RHS evaluation : max|gpu-ref| <= 2e-5 * max|ref| per component
  N-step runs        : relative L2 <= 1e-4 per component / per trace (BASELINE.json north_star)
"""
import numpy as np
import pytest

from cgfd3d_b200 import abi, solver
from oracle import ref_flat
from tests import util

pytestmark = pytest.mark.gpu

TOL_STAGE = 2e-5
TOL_RUN = 1e-4


def _need():
    if not ref_flat.available():
        pytest.fail("oracle/_ref/libcgfd_ref_flat.so is missing (build it with make -C oracle ref where /root/reference exists)")
    if solver.device_count() < 1:
        pytest.fail("no CUDA device: the hot path has no CPU fallback")


def _check_stage(prob, it, ipair, istage, seed):
    w, aux = util.random_state(prob, seed)
    R = ref_flat.RefSolver(prob)
    G = solver.Solver(prob)
    for key, a in aux.items():
        R.set_pml_aux(key[0], key[1], a.ravel())
        G.set_pml_aux(key[0], key[1], a.ravel())
    rr = R.onestage(it, ipair, istage, w)
    rg = G.onestage(it, ipair, istage, w)
    bad = []
    for c in range(9):
        e = util.rel_max(rg[c], rr[c])
        if not e <= TOL_STAGE:
            bad.append((util.CMP[c], e))
    for key in aux:
        ar = R.get_pml_aux_rhs(*key).reshape(9, -1)
        ag = G.get_pml_aux_rhs(*key).reshape(9, -1)
        for c in range(9):
            e = util.rel_max(ag[c], ar[c])
            if not e <= TOL_STAGE:
                bad.append(("aux%s.%s" % (key, util.CMP[c]), e))
    G.close()
    assert not bad, "ipair=%d istage=%d: %s" % (ipair, istage, bad)


@pytest.mark.parametrize("ipair", range(8))
def test_onestage_all_direction_pairs(ipair):
    """every (x,y,z) direction combination: 8 pairs x stages 0,1 cover all 8 kernels twice"""
    _need()
    prob = util.small_problem(seed=7)
    for istage in (0, 1):
        _check_stage(prob, it=3, ipair=ipair, istage=istage, seed=100 + ipair)


@pytest.mark.parametrize("case", ["flat", "nopml", "pml6", "gauss", "force", "odd"])
def test_onestage_variants(case):
    _need()
    kw = dict(seed=3)
    if case == "flat":
        kw.update(topo="flat")
    elif case == "nopml":
        kw.update(pml_faces=())
    elif case == "pml6":
        kw.update(free_top=False)
    elif case == "gauss":
        kw.update(spatial="gauss")
    elif case == "force":
        kw.update(src="force", spatial="gauss")
    elif case == "odd":
        kw.update(ni=37, nj=19, nk=23, pml_layers=5)
    prob = util.small_problem(**kw)
    _check_stage(prob, it=2, ipair=2, istage=1, seed=5)
    _check_stage(prob, it=2, ipair=5, istage=2, seed=6)


@pytest.mark.parametrize("case", ["hill_pml_free", "flat_pml6", "gauss_src"])
def test_run_matches_reference_driver(case):
    """N RK4 steps from rest with a point source: full wavefield, PML aux and recorded traces against
    drv_rk_curv_col_allstep run on the same flat arrays."""
    _need()
    nt = 60
    kw = dict(ni=40, nj=36, nk=30, pml_layers=8, nt_total=nt)
    if case == "flat_pml6":
        kw.update(topo="flat", free_top=False)
    if case == "gauss_src":
        kw.update(spatial="gauss")
    prob = util.small_problem(**kw)
    rec = [prob.iptr(10 + 5 * n, 12 + 3 * n, prob.nk - 1) for n in range(5)] + [prob.iptr(20, 17, prob.nk - 8)]
    R = ref_flat.RefSolver(prob)
    wr, recr, _ = R.run(nt, rec_iptr=rec)
    G = solver.Solver(prob)
    G.set_record_points(rec, nt)
    G.run(nt)
    wg = G.get_wavefield()
    recg = G.get_record(0, nt)
    bad = []
    for c in range(9):
        e = util.rel_l2(wg[c], wr[c])
        if not e <= TOL_RUN:
            bad.append(("w." + util.CMP[c], e))
    for c in range(3):  # velocity traces (surface stresses vanish, their relative error is meaningless)
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


def test_box_and_pg_taps():
    _need()
    nt = 20
    prob = util.small_problem(ni=30, nj=28, nk=24, pml_layers=5, nt_total=nt)
    G = solver.Solver(prob)
    G.run(nt)
    w = G.get_wavefield()
    box = G.get_box(2, 4, 10, 2, 5, 8, 3, prob.nz - 4, 1, 1)
    np.testing.assert_array_equal(box, w[2, prob.nz - 4:prob.nz - 3, 5:5 + 8 * 3:3, 4:4 + 10 * 2:2])
    pg = G.get_pg()
    assert pg.shape[0] == 15 and float(pg[0].max()) > 0
    # PGVz is the running max of |Vz| on the surface: at least the final value
    assert (pg[4, 3:-3, 3:-3] + 1e-30 >= np.abs(w[2, prob.nz - 4, 3:-3, 3:-3])).all()
    G.close()


@pytest.mark.parametrize("medium", ["iso", "vti"])
def test_interior_tiles_without_pml(medium):
    """A block wide enough that some 32x8 tiles meet no PML slab: those run the PML-free copy of the loop body
    (the small cases above never do). One RHS evaluation and a 30-step run against the reference."""
    _need()
    nt = 30
    prob = util.small_problem(ni=100, nj=44, nk=40, pml_layers=5, nt_total=nt, seed=21, medium=medium)
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    G = solver.Solver(prob)
    w, aux = util.random_state(prob, 77)
    for key, a in aux.items():
        R.set_pml_aux(key[0], key[1], a.ravel())
        G.set_pml_aux(key[0], key[1], a.ravel())
    for (ipair, istage) in ((0, 0), (3, 1), (6, 2)):
        rr, rg = R.onestage(1, ipair, istage, w), G.onestage(1, ipair, istage, w)
        bad = [(util.CMP[c], util.rel_max(rg[c], rr[c])) for c in range(9) if not util.rel_max(rg[c], rr[c]) <= TOL_STAGE]
        assert not bad, (ipair, istage, bad)
    G.close()
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    wr, _, _ = R.run(nt)
    G = solver.Solver(prob)
    G.run(nt)
    wg = G.get_wavefield()
    G.close()
    assert np.isfinite(wg).all() and float(np.abs(wr[0]).max()) > 0
    bad = [(util.CMP[c], util.rel_l2(wg[c], wr[c])) for c in range(9) if not util.rel_l2(wg[c], wr[c]) <= TOL_RUN]
    assert not bad, bad


@pytest.mark.parametrize("medium", ["iso", "vti"])
def test_vertically_deformed_grid_kernels(medium, monkeypatch):
    """Grids whose xi_y, xi_z, eta_x, eta_z vanish identically run kernels that never read those arrays (GZ, cgfd_dev.cuh).
    They must agree with the general kernels to float32 round-off (same terms, the zero ones dropped; only the compiler's FMA
    contraction differs: rel L2 <= 2e-6 after the run) and match the reference."""
    _need()
    nt = 24
    prob = util.small_problem(ni=70, nj=30, nk=28, pml_layers=5, nt_total=nt, seed=11, medium=medium)
    for m in (abi.M_XIY, abi.M_XIZ, abi.M_ETX, abi.M_ETZ):
        prob.metric[m][...] = 0.0
    if medium == "iso":
        from cgfd3d_b200 import hostsetup as hs
        mvx, mvy, mf = hs.dvh2dvz_iso(prob.metric, prob.media[0], prob.media[1], prob.grid)
        prob.mats = dict(matVx2Vz=mvx, matVy2Vz=mvy, matF2Vz=mf, matD=np.zeros_like(mf))
    R = ref_flat.RefSolver(prob)
    util.fill_surface_matrices(prob, R)
    wr, _, _ = R.run(nt)
    out = {}
    for gz in ("1", "0"):
        monkeypatch.setenv("CGFD_GZ", gz)
        G = solver.Solver(prob)
        assert G.grid_class() == int(gz)
        G.run(nt)
        out[gz] = G.get_wavefield()
        G.close()
    assert float(np.abs(wr[0]).max()) > 0
    e10 = [util.rel_l2(out["1"][c], out["0"][c]) for c in range(9)]
    e1r = [util.rel_l2(out["1"][c], wr[c]) for c in range(9)]
    e0r = [util.rel_l2(out["0"][c], wr[c]) for c in range(9)]
    msg = "gz vs general %s | gz vs ref %s | general vs ref %s" % (["%.1e" % e for e in e10], ["%.1e" % e for e in e1r], ["%.1e" % e for e in e0r])
    assert max(e10) <= 2e-6 and max(e1r) <= TOL_RUN and max(e0r) <= TOL_RUN, msg
    # the test grids built from coordinates carry round-off in those four arrays, like the reference's own: general kernels
    monkeypatch.delenv("CGFD_GZ")
    G = solver.Solver(util.small_problem(seed=1))
    assert G.grid_class() == 0
    G.close()


def test_streaming_snapshot_and_staged_wavefield_copies():
    """add_snapshot frames written while run() advances equal per-step get_box gathers; set/get_wavefield (staged,
    re-pitched on the device) round-trip bit-exactly, ghosts included."""
    _need()
    nt = 12
    prob = util.small_problem(ni=45, nj=28, nk=24, pml_layers=5, nt_total=nt)
    rng = np.random.default_rng(5)
    G = solver.Solver(prob)
    w0 = rng.uniform(-1, 1, G.shape).astype(np.float32)
    G.set_wavefield(w0)
    np.testing.assert_array_equal(G.get_wavefield(), w0)
    G.close()
    # reference sequence: one step at a time, synchronous gathers
    box_a = (3, prob.ni, 1, 3, prob.nj, 1, prob.nz - 4, 1, 1)      # surface, every point
    box_b = (4, 10, 3, 5, 6, 2, 6, 4, 3)                            # strided volume
    G = solver.Solver(prob)
    fa, fb = [], []
    for it in range(nt):
        G.run(1, it0=it)
        fa.append(np.stack([G.get_box(c, *box_a) for c in (0, 1, 2)]))
        if it >= 2 and (it - 2) % 3 == 0:
            fb.append(np.stack([G.get_box(c, *box_b) for c in (2, 5)]))
    G.close()
    G = solver.Solver(prob)
    ia, oa = G.add_snapshot((0, 1, 2), box_a, max_frames=nt)
    ib, ob = G.add_snapshot((2, 5), box_b, max_frames=3, it1=2, tinv=3)   # fewer frames allowed than due: extra ones are dropped
    G.run(nt)
    assert G.snapshot_frames(ia) == nt and G.snapshot_frames(ib) == 3
    np.testing.assert_array_equal(oa, np.stack(fa))
    np.testing.assert_array_equal(ob, np.stack(fb[:3]))
    assert float(np.abs(oa).max()) > 0
    G.close()


def test_metric_from_coords_on_device():
    """cgfd_b200_metric_from_coords = gd_curv_metric_cal (forward/gd_t.c:190-402) on the GPU: bit-identical to the reference
    function (products and sums rounded separately, ghosts mirrored in the same order) on a hill grid with perturbed x-y lines."""
    _need()
    prob = util.small_problem(ni=37, nj=29, nk=23)
    x, y, z = (a.copy() for a in prob.coords)
    rng = np.random.default_rng(2)
    x += rng.uniform(-8, 8, x.shape).astype(np.float32)      # a genuinely curvilinear grid: every metric array matters
    y += rng.uniform(-8, 8, y.shape).astype(np.float32)
    ref = ref_flat.RefSolver(prob).metric_from_coords(x, y, z)
    got = solver.metric_from_coords(prob.grid, x, y, z)
    assert float(np.abs(ref[2]).max()) > 0
    for m in range(10):
        np.testing.assert_array_equal(got[m], ref[m])


def test_async_blocks_match_plain_run():
    """run_async / wait_block / snapshot_set_output (the calls the drop-in driver advances with: block b+1 enqueued before block b is
    waited for, two host buffers alternating): frames, record samples and the final wavefield equal those of one plain run()."""
    _need()
    nt, B = 29, 8
    prob = util.small_problem(ni=45, nj=28, nk=24, pml_layers=5, nt_total=nt)
    rec = [prob.iptr(10 + 5 * n, 6 + 3 * n, prob.nk - 1) for n in range(5)]
    box = (3, prob.ni // 2, 2, 3, prob.nj // 2, 2, prob.nz - 4, 1, 1)
    G = solver.Solver(prob)
    G.set_record_points(rec, nt)
    sid, frames = G.add_snapshot((0, 1, 2), box, max_frames=nt, it1=2, tinv=3)
    G.run(nt)
    w_ref, rec_ref, nfr = G.get_wavefield(), G.get_record(0, nt), G.snapshot_frames(sid)
    G.close()
    G = solver.Solver(prob)
    G.set_record_points(rec, nt)
    bufs = [np.zeros((B // 3 + 2,) + frames.shape[1:], np.float32) for _ in range(2)]
    recb = [np.zeros((B, 9, len(rec)), np.float32) for _ in range(2)]
    sid, _ = G.add_snapshot((0, 1, 2), box, max_frames=bufs[0].shape[0], it1=2, tinv=3, out=bufs[0])
    got_frames, got_rec, blocks = [], [], []
    it = 0
    while it < nt or blocks:
        if it < nt:
            n = min(B, nt - it)
            cur = (it // B) % 2
            G.snapshot_set_output(sid, bufs[cur])
            G.run_async(n, it0=it, rec_out=recb[cur])
            blocks.append((it, n, cur))
            it += n
        if len(blocks) == 2 or it >= nt:
            b0, n0, c0 = blocks.pop(0)
            G.wait_block(b0 + n0 - 1)
            nf = sum(1 for q in range(b0, b0 + n0) if q >= 2 and (q - 2) % 3 == 0)
            got_frames.append(bufs[c0][:nf].copy())
            got_rec.append(recb[c0][:n0].copy())
    G.sync()
    w = G.get_wavefield()
    G.close()
    np.testing.assert_array_equal(w, w_ref)
    np.testing.assert_array_equal(np.concatenate(got_rec), rec_ref)
    np.testing.assert_array_equal(np.concatenate(got_frames), frames[:nfr])
    assert nfr == 9 and float(np.abs(frames).max()) > 0
