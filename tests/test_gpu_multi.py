"""x-y decomposition over several GPUs with the NCCL halo exchange must reproduce the single-GPU run
(needs >= 2 GPUs on the box; skipped otherwise -- the strip logic itself is covered on CPU by test_cpu_halo_gloo)."""
import os
import subprocess
import sys

import pytest

from cgfd3d_b200 import solver

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("split,medium", [("x", "iso"), ("y", "iso"), ("x", "visco"), ("y", "vti")])
def test_two_ranks_match_one(split, medium):
    n = solver.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs, found %d" % n)
    port = 29631 + ["x", "y"].index(split) + 2 * ["iso", "vti", "visco"].index(medium)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "multi_gpu_check.py"), split, medium]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_GPU_CHECK" in p.stdout and '"ok": true' in p.stdout, p.stdout[-2000:]


@pytest.mark.parametrize("case", ["iso_hill", "visco", "ablexp"])
def test_dropin_two_ranks_match_reference_two_ranks(case):
    """The drop-in binary as a 2-rank program (ranks forked by the MPI stand-in of oracle/shims/mpi_shim.c, rank r on GPU r, NCCL id
    broadcast through MPI_Bcast exactly as integration/drv_rk_curv_col_b200.c does under a real MPI) against the reference program on
    the same two ranks: per-rank line seismograms and surface snapshots. Covers what only exists with neighbours: the halo exchange
    inside the reference-facing driver, the sponge applied AFTER the exchange (forward/drv_rk_curv_col.c:483-485) and the
    visco-elastic surface rotation after the halo send."""
    import tempfile
    import numpy as np
    from oracle import harness as H
    from tests import util
    if solver.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    BIN = H.ref_binary("cgfd_main_b200")
    if not os.path.isfile(BIN):
        pytest.fail("oracle/_ref/cgfd_main_b200 missing")
    size, nt = (64, 40, 30), 100
    kw = dict(lines=[{"name": "L1", "grid_index_start": [6, 8, size[2] - 1], "grid_index_incre": [9, 5, 0], "grid_index_count": 6}],
              snapshots=[{"name": "surf", "grid_index_start": [0, 0, size[2] - 1], "grid_index_count": [size[0] // 2, size[1] // 2, 1],
                          "grid_index_incre": [2, 2, 1], "time_index_start": 0, "time_index_incre": 10,
                          "save_velocity": 1, "save_stress": 0, "save_strain": 0}],
              src=H.moment_src(31, 20, 8, m=(1e16, 0.5e16, 2e16, 0.2e16, -0.1e16, 0.3e16)))
    if case == "visco":
        kw.update(medium_type="viscoelastic_iso", visco={"type": "gmb", "number_of_maxwell": 3, "max_freq": 10.0, "min_freq": 0.1, "refer_freq": 1.0})
    if case == "ablexp":
        kw.update(pml_sides=(), ablexp_sides=("x_left", "x_right", "y_front", "y_back", "z_bottom"))
    out = {}
    for name, binary in (("ref", H.ref_binary("ref_main_zero")), ("gpu", BIN)):
        wd = tempfile.mkdtemp(prefix="cgfd_multi_")
        H.write_multirank_hill_case(wd, size, 2, 1, nt, 0.01, hill=(400.0, 1000.0), pml_layers=6, **kw)
        env = dict(os.environ, CGFD_SHIM_NPROCS="2")
        wall, txt = H.run(binary, wd, timeout=900, env=env)
        sac = H.read_sac_dir(os.path.join(wd, "OUT"))
        snaps = {r: H.read_cgnc(os.path.join(wd, "OUT", "surf_px%d_py0.nc" % r))["vars"] for r in (0, 1)}
        out[name] = (sac, snaps)
    sr, sg = out["ref"][0], out["gpu"][0]
    assert set(sr) == set(sg) and len(sr) >= 54
    bad = []
    for ir in range(6):
        for names in (("Vx", "Vy", "Vz"), ("Txx", "Tyy", "Tzz", "Tyz", "Txz", "Txy")):
            ks = ["evt1.L1.no%d.%s" % (ir, c) for c in names]
            num = np.sqrt(sum(float(np.sum((sg[k].astype(np.float64) - sr[k]) ** 2)) for k in ks))
            den = np.sqrt(sum(float(np.sum(sr[k].astype(np.float64) ** 2)) for k in ks))
            assert den > 0
            if not num <= 1e-4 * den:
                bad.append((ir, names[0], num / den))
    for r in (0, 1):
        for v in ("Vx", "Vy", "Vz"):
            e = util.rel_l2(out["gpu"][1][r][v], out["ref"][1][r][v])
            if not e <= 1e-4:
                bad.append(("snap rank %d %s" % (r, v), e))
    assert not bad, bad
