"""Size-independent property at the full size of BASELINE.json configs[1] (400x400x200, hill topography, CFS-PML + free surface):
the numerical domain of dependence. One RK stage reaches at most 3 points along every axis, one step 12. After nt steps a point
further than 12*nt from every boundary of a block cannot know where the block ends, so a window around a deep source must come out
BIT-IDENTICAL on the full grid and on a small block cut out of it (same metric / media values, other tile alignment, other launch
plan, no PML, no free surface) - the reference CPU code cannot serve as the checker at this size within the test budget."""
import numpy as np
import pytest

from cgfd3d_b200 import hostsetup as hs, solver

pytestmark = pytest.mark.gpu
NI, NJ, NK = 400, 400, 200


@pytest.fixture(scope="module")
def big_problem():
    if solver.device_count() < 1:
        pytest.fail("no CUDA device: the hot path has no CPU fallback")
    return hs.build_problem(NI, NJ, NK, topo="hill", hill=(1000.0, 4000.0), pml_layers=10, free_top=True, dt=0.012)


def test_linearity_400x400x200(big_problem):
    """The scheme is linear in the source: doubling the moment tensor doubles every value of the wavefield EXACTLY (a factor 2
    commutes with every float32 product, sum and FMA as long as no denormals are involved). Checked on the free-surface plane, an x-PML slab, a y-PML slab and an
    interior box after 6 steps, i.e. through the interior, PML and free-surface code paths at full size."""
    prob = big_problem
    boxes = [(3, NI, 1, 3, NJ, 1, NK + 2, 1, 1),            # free surface (array indices: physical + 3)
             (3, 11, 1, 14, 48, 1, 150, 52, 1),              # x1 PML slab up to the surface rows
             (14, 48, 1, 3, 11, 1, 150, 52, 1),              # y1 PML slab up to the surface rows
             (14, 80, 1, 14, 80, 1, 140, 62, 1)]             # interior + free-surface rows around the source
    res = []
    for amp in (1.0, 2.0):
        hs.make_source(prob, 30, 30, NK - 1 - 12, nt_total=100, kind="moment",   # 20 points from the x1 and y1 slabs (a step reaches >= 4), 12 below the surface
                       mech=tuple(amp * m for m in (1e16, 0.6e16, 1.3e16, 0.2e16, -0.4e16, 0.3e16)), fc=2.0, t0=0.0, stf_len=1.0)
        G = solver.Solver(prob)
        G.run(6)
        res.append([np.stack([G.get_box(c, *b) for c in range(9)]) for b in boxes])
        G.close()
    assert all(float(np.abs(r).max()) > 0 for r in res[0])   # the wave has reached every box
    for a, b in zip(*res):
        assert np.isfinite(a).all()
        # exact wherever the values are ordinary floats; at the leading edge of the numerical precursor they fall into the
        # denormal range, where rounding is absolute and a factor 2 no longer commutes with it
        big_enough = np.abs(a) > 1e-25
        assert int(big_enough.sum()) > a.size // 100
        np.testing.assert_array_equal(2.0 * a[big_enough], b[big_enough])
        assert float(np.abs(b[~big_enough] - 2.0 * a[~big_enough]).max(initial=0.0)) <= 1e-24


def test_domain_of_dependence_400x400x200(big_problem):
    ni, nj, nk, nt, n_small, rad = NI, NJ, NK, 3, 96, 10
    big = big_problem
    si, sj, sk = 203, 197, 101                      # deep source, off the tile grid
    kw = dict(nt_total=100, kind="moment", mech=(1e16, 0.6e16, 1.3e16, 0.2e16, -0.4e16, 0.3e16), fc=2.0, t0=0.0, stf_len=1.0)
    hs.make_source(big, si, sj, sk, **kw)
    # the block [o, o + n_small) of the physical range around the source, ghosts cut out of the big arrays as they are
    h = n_small // 2
    oi, oj, ok = si - h, sj - h, sk - h
    cut = (slice(ok, ok + n_small + 6), slice(oj, oj + n_small + 6), slice(oi, oi + n_small + 6))
    small = hs.HostProblem(ni=n_small, nj=n_small, nk=n_small, dt=big.dt, free_top=0, timg_mode=big.timg_mode, neigh=(-1, -1, -1, -1))
    small.metric = [np.ascontiguousarray(m[cut]) for m in big.metric]
    small.media = [np.ascontiguousarray(m[cut]) for m in big.media]
    hs.make_source(small, h, h, h, **kw)
    win_b = (si + 3 - rad, 2 * rad + 1, 1, sj + 3 - rad, 2 * rad + 1, 1, sk + 3 - rad, 2 * rad + 1, 1)
    win_s = (h + 3 - rad, 2 * rad + 1, 1, h + 3 - rad, 2 * rad + 1, 1, h + 3 - rad, 2 * rad + 1, 1)
    assert h - rad > 12 * nt and min(si, sj, sk, ni - si, nj - sj, nk - sk) - rad > 12 * nt
    out = []
    for prob, win in ((big, win_b), (small, win_s)):
        G = solver.Solver(prob)
        G.run(nt)
        out.append(np.stack([G.get_box(c, *win) for c in range(9)]))
        G.close()
    assert float(np.abs(out[0][2]).max()) > 0 and np.isfinite(out[0]).all()
    for c in range(9):
        np.testing.assert_array_equal(out[0][c], out[1][c])


def test_onestage_matches_reference_at_full_size(big_problem):
    """The bench workload itself against the UNMODIFIED reference: one RHS evaluation of sv_curv_col_el_iso_onestage on the
    400x400x200 hill problem (random wavefield and PML auxiliary state, moment source active), two operator pairs / stages with
    opposite zeta directions -- every interior tile, every PML face and the free-surface rows at the size the metric is quoted on.
    Tolerance as everywhere: max|gpu - ref| <= 2e-5 max|ref| per component."""
    from oracle import ref_flat
    from tests import util
    if not ref_flat.available():
        pytest.fail("oracle/_ref/libcgfd_ref_flat.so is missing")
    prob = big_problem
    hs.make_source(prob, NI // 2, NJ // 2, NK - 1 - 20, nt_total=100, kind="moment", mech=(1e16, 0.6e16, 1.3e16, 0.2e16, -0.4e16, 0.3e16),
                   fc=2.0, t0=0.1, stf_len=1.0)
    w, aux = util.random_state(prob, seed=2024)
    R = ref_flat.RefSolver(prob)
    G = solver.Solver(prob)
    for key, a in aux.items():
        R.set_pml_aux(key[0], key[1], a.ravel())
        G.set_pml_aux(key[0], key[1], a.ravel())
    bad = []
    for (it, ipair, istage) in ((5, 1, 1), (6, 6, 2)):
        rr = R.onestage(it, ipair, istage, w)
        rg = G.onestage(it, ipair, istage, w)
        for c in range(9):
            e = util.rel_max(rg[c], rr[c])
            if not e <= 2e-5:
                bad.append((ipair, istage, util.CMP[c], e))
        for key in aux:
            ar = R.get_pml_aux_rhs(*key).reshape(9, -1)
            ag = G.get_pml_aux_rhs(*key).reshape(9, -1)
            for c in range(9):
                e = util.rel_max(ag[c], ar[c])
                if not e <= 2e-5:
                    bad.append((ipair, istage, "aux%s.%s" % (key, util.CMP[c]), e))
        assert float(np.abs(rr[0]).max()) > 0 and float(np.abs(rr[5]).max()) > 0
        del rr, rg
    G.close()
    R.close()
    assert not bad, bad
