"""Pins the C restatement (oracle/cgfd_oracle.c) against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref): one RHS evaluation per direction pair on random fields, and multi-step runs.
Both are built without FMA contraction, so agreement is at round-off of float32 re-association
(the restatement sums a few terms in loops the reference unrolls): <= 2e-6 of max|ref|."""
import numpy as np
import pytest

from oracle import port, ref_flat
from tests import util

need = pytest.mark.skipif(not (port.available() and ref_flat.available()),
                          reason="oracle libraries not built (make -C oracle && make -C oracle ref)")
TOL = 2e-6


def _stage(prob, it, ipair, istage, seed):
    w, aux = util.random_state(prob, seed)
    R, P = ref_flat.RefSolver(prob), port.PortSolver(prob)
    for key, a in aux.items():
        R.set_pml_aux(key[0], key[1], a.ravel())
        P.set_pml_aux(key[0], key[1], a.ravel())
    rr, rp = R.onestage(it, ipair, istage, w), P.onestage(it, ipair, istage, w)
    bad = [(util.CMP[c], util.rel_max(rp[c], rr[c])) for c in range(9) if not util.rel_max(rp[c], rr[c]) <= TOL]
    for key in aux:
        ar, ap = R.get_pml_aux_rhs(*key).reshape(9, -1), P.get_pml_aux_rhs(*key).reshape(9, -1)
        bad += [("aux%s.%d" % (key, c), util.rel_max(ap[c], ar[c])) for c in range(9) if not util.rel_max(ap[c], ar[c]) <= TOL]
    assert not bad, (ipair, istage, bad)


@need
@pytest.mark.parametrize("ipair", range(8))
def test_port_onestage_matches_reference(ipair):
    prob = util.small_problem(seed=7)
    _stage(prob, 3, ipair, 0, 100 + ipair)
    _stage(prob, 3, ipair, 1, 200 + ipair)


@need
@pytest.mark.parametrize("case", ["nopml", "pml6", "gauss", "force"])
def test_port_onestage_variants(case):
    kw = dict(seed=3)
    if case == "nopml":
        kw.update(pml_faces=())
    elif case == "pml6":
        kw.update(free_top=False)
    elif case == "gauss":
        kw.update(spatial="gauss")
    elif case == "force":
        kw.update(src="force", spatial="gauss")
    prob = util.small_problem(**kw)
    _stage(prob, 2, 2, 1, 5)
    _stage(prob, 2, 5, 2, 6)


@need
def test_port_run_matches_reference_driver():
    nt = 40
    prob = util.small_problem(ni=30, nj=26, nk=24, pml_layers=6, nt_total=nt)
    rec = [prob.iptr(8 + 3 * n, 9 + 2 * n, prob.nk - 1) for n in range(5)]
    wr, recr, _ = ref_flat.RefSolver(prob).run(nt, rec_iptr=rec)
    wp, recp, _ = port.PortSolver(prob).run(nt, rec_iptr=rec)
    assert float(np.abs(wr[0]).max()) > 0
    for c in range(9):
        assert util.rel_l2(wp[c], wr[c]) <= 1e-5, (c, util.rel_l2(wp[c], wr[c]))
    for c in range(3):
        for ip in range(len(rec)):
            assert util.rel_l2(recp[:, c, ip], recr[:, c, ip]) <= 1e-5
