#!/usr/bin/env python
"""bench.py -- Gpoint-updates/s per RK4 step of the CGFD3D time-stepping hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size NIxNJxNK]

N = 1 : BASELINE.json configs[1]: isotropic elastic, Gaussian-hill topography (curvilinear grid),
        400x400x200, CFS-PML (10 layers, 5 faces) + traction-image free surface, one moment source.
N > 1 : launched by torchrun, one rank per GPU; BASELINE.json configs[2]: weak scaling, 800x800x400 per
        GPU (overridable with --size), x-y decomposition, NCCL halo exchange.

One JSON line on stdout (rank 0). `value` = physical grid points advanced one RK4 step per second with
everything resident in HBM, timed with CUDA events on the solver's stream (max over ranks);
`e2e` = the same through the public API from HOST buffers: initial wavefield H2D, every step the
receiver samples and a surface Vx/Vy/Vz snapshot streamed D2H while the steps run, final wavefield D2H.
`--impl reference` times the unmodified reference CPU code (oracle/_ref, built from /root/reference
by oracle/Makefile) on a bounded sample of the same workload with one replica per host core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# load every kernel image when the context is created, not at first launch inside the timed region
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

METRIC = "Gpoint-updates/s per RK4 step"
BYTES_PER_POINT_STEP_ISO = 768.0  # 4*(16*C + 4*M), C = 9, M = 12
# the same formula for the other media (visco: 3 Maxwell bodies, C = 27, M = 9 + 3 + 6)
BYTES_PER_POINT_STEP = {"iso": 768.0, "vti": 816.0, "aniso": 1072.0, "visco": 2016.0}
WORKLOAD = {"iso": "isotropic elastic", "vti": "VTI elastic", "aniso": "general anisotropic elastic", "visco": "visco-elastic isotropic (GMB, 3 Maxwell bodies)"}


def parse_size(s):
    a = [int(v) for v in s.lower().split("x")]
    assert len(a) == 3
    return tuple(a)


def proc_grid(n):
    """process grid of the weak-scaling runs: 2x1, 2x2, 4x2."""
    return {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(n) or (n, 1)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons of one GPU while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.stop = False
        self.t = None

    def _loop_nvml(self, nv, h):
        # NVML in-process: a sample every ~5 ms (an nvidia-smi call takes > 100 ms, too coarse for a sub-second region)
        R = nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown, \
            nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx), "%.2f" % pw] + ["Active" if rs & r else "Not Active" for r in R])
            except Exception:
                pass
            time.sleep(0.005)

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:   # NVML enumerates physical devices
                tok = vis.split(",")[self.index].strip()
                if tok.isdigit():
                    idx = int(tok)
                else:
                    uuid = tok
            h = nv.nvmlDeviceGetHandleByUUID(uuid) if uuid else nv.nvmlDeviceGetHandleByIndex(idx)
            self.how = "nvml"
            return self._loop_nvml(nv, h)
        except Exception:
            self.how = "nvidia-smi"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "how": getattr(self, "how", None)}


def build_rank_problem(size, rank, nranks, device=None, medium="iso"):
    """One rank's block of the global hill problem. device=None: numpy on the host (hostsetup); device='cuda:N': the same
    formulas as torch expressions on that GPU (devsetup) -- used for the 800x800x400-per-GPU runs, where the numpy route
    needs ~45 GB of host memory and minutes per rank."""
    from cgfd3d_b200 import hostsetup as hs
    ni, nj, nk = size
    px, py = proc_grid(nranks)
    ix, iy = rank // py, rank % py   # row-major, y fastest (MPI_Cart_create, forward/mympi_t.c:32-40)
    def rk(a, b):
        return a * py + b if (0 <= a < px and 0 <= b < py) else -1
    neigh = (rk(ix - 1, iy), rk(ix + 1, iy), rk(ix, iy - 1), rk(ix, iy + 1))
    gni, gnj = ni * px, nj * py
    # hill spans the global domain (sigma ~ 1/10 of it, height ~ 10 cells)
    dh = (100.0, 100.0, 100.0)
    sigma = 0.1 * max(gni, gnj) * dh[0]
    # dt below the CFL bound of the stretched grid (estimate_dt on the full array is slow; checked in tests)
    if medium != "iso":
        # the other constitutive laws (BASELINE.json configs[3], [4] are parity cases, not bench lines): CFS-PML on all six
        # faces, because their free-surface matrices come from the reference's own host set-up code
        prob = hs.build_problem(ni, nj, nk, dh=dh, topo="hill", hill=(1000.0, sigma), pml_layers=10, free_top=False,
                                pml_faces=((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)), dt=0.008,
                                sub=(ix * ni, iy * nj, gni, gnj, neigh), medium=medium)
    elif device is None:
        prob = hs.build_problem(ni, nj, nk, dh=dh, topo="hill", hill=(1000.0, sigma), pml_layers=10, free_top=True,
                                dt=0.012, sub=(ix * ni, iy * nj, gni, gnj, neigh))
    else:
        from cgfd3d_b200 import devsetup
        prob = devsetup.build_problem(ni, nj, nk, device=device, dh=dh, hill=(1000.0, sigma), pml_layers=10, free_top=True,
                                      dt=0.012, sub=(ix * ni, iy * nj, gni, gnj, neigh))
    # one explosive moment source under the hill top, on the rank that owns it
    gsi, gsj = gni // 2, gnj // 2
    if ix * ni <= gsi < (ix + 1) * ni and iy * nj <= gsj < (iy + 1) * nj:
        hs.make_source(prob, gsi - ix * ni, gsj - iy * nj, nk - 1 - 20, nt_total=100000, kind="moment",
                       mech=(1e16, 1e16, 1e16, 0, 0, 0), fc=2.0, t0=0.5, stf_len=1.0)
    return prob


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cgfd3d_b200 import solver

    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if nranks != args.gpus:
        if nranks == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    if nranks > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = args.size or ((400, 400, 200) if nranks == 1 else (800, 800, 400))
    t0 = time.time()
    on_device = size[0] * size[1] * size[2] > 64e6 and args.medium == "iso"   # big blocks: set-up arrays built on the GPU (devsetup.py)
    prob = build_rank_problem(size, rank, nranks, device=("cuda:%d" % local) if on_device else None, medium=args.medium)
    t_host = time.time() - t0
    t0 = time.time()
    S = solver.Solver(prob, device=local)
    t_upload = time.time() - t0
    if on_device:   # the library holds its own padded copies
        prob.metric = prob.media = None
        torch.cuda.empty_cache()
    if nranks > 1:
        uid = [solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        S.comm_init(uid[0], rank, nranks)
    if args.variant:
        S.set_variant(args.variant)
    ni, nj, nk = size
    npts = ni * nj * nk
    K, W = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    # the 8 operator pairs (it % 8) use 8 different kernel instantiations: touch all of them before timing
    S.run(max(W, 8), it0=0)
    W = max(W, 8)
    S.set_profiling(True)
    barrier()
    with ClockSampler(local) as clk:
        tw0 = time.time()
        S.run(K, it0=W)
        barrier()
        tw1 = time.time()
    ms = S.last_run_ms()
    main_ms, main_n, launches = S.get_profile()
    S.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if nranks > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = npts * nranks * K / (ms_max * 1e-3) / 1e9

    # ---- end to end from host buffers -------------------------------------------------------------
    # the call sequence of a production run: initial wavefield from pinned host memory, K steps with the outputs the
    # reference's example writes (receiver traces + a surface Vx/Vy/Vz snapshot EVERY step, example/cgfd3d.example.sh
    # :323-334) streamed to pinned host buffers while the steps run, final wavefield back to the host
    rec = [prob.iptr(ni // 2 + 5 * n, nj // 2 + 3 * n, nk - 1) for n in range(-4, 5)]
    S.set_record_points(rec, K + 8)
    w_host = torch.zeros(S.shape, dtype=torch.float32).pin_memory().numpy()
    ncmp = S.shape[0]
    snap = torch.zeros((K, 3, 1, nj, ni), dtype=torch.float32).pin_memory().numpy()
    S.add_snapshot((0, 1, 2), (3, ni, 1, 3, nj, 1, prob.nz - 4, 1, 1), max_frames=K, it1=W + K, out=snap)
    h2d = w_host.nbytes / K
    d2h = w_host.nbytes / K + len(rec) * ncmp * 4 + snap.nbytes / K
    barrier()
    te0 = time.time()
    S.set_wavefield(w_host)
    te1 = time.time()
    S.run(K, it0=W + K)
    traces = S.get_record(0, K)
    te2 = time.time()
    S.get_wavefield(out=w_host)
    barrier()
    te3 = time.time()
    te = torch.tensor([te3 - te0], dtype=torch.float64, device="cuda")
    if nranks > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = npts * nranks * K / float(te.item()) / 1e9
    e2e_ms = {"set_wavefield": round((te1 - te0) * 1e3, 2), "steps_with_outputs": round((te2 - te1) * 1e3, 2),
              "get_wavefield": round((te3 - te2) * 1e3, 2)}
    finite = bool(np.isfinite(w_host).all()) and bool(np.isfinite(snap).all()) and bool(np.isfinite(traces).all())

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        kb = (main_ms / max(main_n, 1)) * 1e-3
        # algorithmic bytes of one interior-kernel launch = 768/4 B per point-stage x the points it covers
        bpps = BYTES_PER_POINT_STEP[args.medium]
        main_pts = ni * nj * (nk - 4 if prob.free_top else nk)   # the free-surface kernel owns the top 4 rows
        ach = (bpps / 4.0) * main_pts / kb / 1e9 if main_n else None
        out = {
            "metric": METRIC, "value": round(value, 4), "unit": "Gpoint-updates/s", "n_gpus": nranks, "steps": K, "warmup": W,
            "ms_per_step": round(ms_max / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("%s, Gaussian-hill topography (curvilinear), %dx%dx%d per GPU, " % ((WORKLOAD[args.medium],) + tuple(size)))
                                   + ("CFS-PML 10 layers x 5 faces, traction-image free surface, 1 moment source" if prob.free_top
                                      else "CFS-PML 10 layers x 6 faces, 1 moment source"),
                       "proc_grid": "%dx%d" % proc_grid(nranks), "l2": "working set >> 126 MB L2 (no flush needed)",
                       "variant": args.variant or "default",
                       "kernels": "vertically-deformed-grid (4 metric arrays identically zero)" if S.grid_class() else "general curvilinear"},
            "e2e": {"value": round(e2e, 4), "unit": "Gpoint-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "set_wavefield (H2D) + run(K) with receiver traces and a surface Vx/Vy/Vz snapshot streamed to pinned host "
                            "memory every step + get_wavefield (D2H)", "ms": e2e_ms},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": "k_main_tma", "achieved": None if ach is None else round(ach, 1), "peak": peak,
                         "unit": "GB/s", "frac": None if ach is None else round(ach / peak, 4), "traffic": ncu_traffic(size),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "launches_timed": int(main_n), "avg_launch_ms": round(main_ms / max(main_n, 1), 4),
                         "algorithmic_bytes_per_launch": int((bpps / 4.0) * main_pts),
                         "whole_step_frac": round(bpps * npts / (ms_max / K * 1e-3) / 1e9 / peak, 4)},
            "setup_s": {"arrays": round(t_host, 2), "arrays_built_on": "gpu (torch)" if on_device else "host (numpy)", "upload": round(t_upload, 2)},
            "wall_s_timed": round(tw1 - tw0, 4), "finite": finite,
        }
        if nranks == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1, steps=args.cpu_steps)
    S.close()
    if nranks > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def ncu_traffic(size):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
    capture (profiles/traffic.json), valid for the workload it was captured on; None otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t["bytes_per_launch"] if list(size) == t["size"] else None
    except Exception:
        return None


# ---- reference CPU arm --------------------------------------------------------------------------
# The reference program itself (oracle/_ref/ref_main_zero = unmodified CGFD3D sources, compiled by oracle/Makefile where
# /root/reference exists) on the bench workload: the SAME 400x400x200 Gaussian-hill problem, imported through the reference's own
# coord_px?_py?.nc route and split over one MPI rank per host core (x-y blocks by gd_indx_set, real halo exchange every RK stage).
# There is no MPI in the image: the ranks are forked by the shared-memory stand-in of oracle/shims/mpi_shim.c.
REF_SIZE = (400, 400, 200)
REF_HILL = (1000.0, 4000.0)     # the hill of build_rank_problem at this size (sigma = 0.1 * 400 * 100 m)
REF_DT = 0.012


def _rank_grid(n):
    """px x py = n ranks, as square as possible, px >= py"""
    py = int(n ** 0.5)
    while n % py:
        py -= 1
    return n // py, py


def _ref_program_seconds(size, px, py, nt, hill):
    """wall seconds of one run of the reference program with px*py ranks (set-up + nt time steps)"""
    import shutil
    import tempfile
    from oracle import harness as H
    wd = tempfile.mkdtemp(prefix="cgfd_refarm_")
    try:
        H.write_multirank_hill_case(wd, size, px, py, nt, REF_DT, hill=hill, pml_layers=10, src=H.moment_src(size[0] // 2, size[1] // 2, 20))
        env = dict(os.environ, CGFD_SHIM_NPROCS=str(px * py), CGFD_SHIM_MSG_MB="64")
        wall, out = H.run(H.ref_binary("ref_main_zero"), wd, verbose=1, timeout=3000, env=env)
        sac = H.read_sac_dir(os.path.join(wd, "OUT"))
        finite = all(bool(np.isfinite(v).all()) for v in sac.values()) and len(sac) > 0
        return wall, finite
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def cpu_baseline(cores, steps=4, warm=1, size=REF_SIZE, hill=REF_HILL):
    """Gpoint-updates/s of the reference program on `cores` ranks: (wall of a run of warm + steps steps) - (wall of a run of warm
    steps), i.e. the time loop alone (the program's own timer has a resolution of one second)."""
    from oracle import harness as H
    if not H.have_ref("ref_main_zero"):
        return {"value": None, "unit": "Gpoint-updates/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    px, py = _rank_grid(cores)
    t_w, ok1 = _ref_program_seconds(size, px, py, warm, hill)
    t_wk, ok2 = _ref_program_seconds(size, px, py, warm + steps, hill)
    secs = max(t_wk - t_w, 1e-6)
    npts = size[0] * size[1] * size[2]
    return {"value": round(npts * steps / secs / 1e9, 6), "unit": "Gpoint-updates/s", "cores": cores, "kind": "reference",
            "ranks": "%dx%d" % (px, py), "seconds_per_step": round(secs / steps, 4), "setup_s": round(t_w, 2), "finite": bool(ok1 and ok2),
            "sample": "%d RK4 steps of the bench workload itself (isotropic, Gaussian hill %dx%dx%d through gd_curv_coord_import, CFS-PML 10x5, "
                      "free surface, 1 moment source) by the unmodified reference program on %dx%d ranks with halo exchange "
                      "(fork/shared-memory MPI stand-in: no MPI in the image); time = wall(%d steps) - wall(%d steps)"
                      % ((steps,) + tuple(size) + (px, py, warm + steps, warm))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K = max(1, min(args.steps, 6))
    t0 = time.time()
    cb = cpu_baseline(cores, steps=K, warm=1)
    # one core, for scale: a 200x200x100 cut of the same physics (the full block needs minutes per step on one core)
    one = cpu_baseline(1, steps=2, warm=1, size=(200, 200, 100), hill=(1000.0, 2000.0)) if args.one_core else None
    wall = time.time() - t0
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Gpoint-updates/s", "n_gpus": args.gpus, "steps": K,
           "warmup": 1, "ms_per_step": None if cb["value"] is None else round(cb["seconds_per_step"] * 1e3, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "isotropic elastic, Gaussian-hill topography (curvilinear), %dx%dx%d, CFS-PML 10 layers x 5 faces, "
                                  "traction-image free surface, 1 moment source; reference CPU program on %s host ranks" % (REF_SIZE + (cb.get("ranks"),))},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "Gpoint-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": round(wall, 2)}
    if one:
        out["cpu_one_core"] = one
    emit(out)


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything any library prints on fd 1 during the run (NCCL's version banner, ...) goes to stderr; the one JSON line is
    written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size", type=parse_size, default=None)
    ap.add_argument("--variant", default="")
    ap.add_argument("--medium", default="iso", choices=["iso", "vti", "aniso", "visco"],
                    help="constitutive law (default iso = the BASELINE.json metric; the others are side measurements)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--one-core", action="store_true", help="--impl reference: also time one core on a 200x200x100 cut")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
