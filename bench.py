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


def build_rank_problem(size, rank, nranks, device=None, medium="iso", global_size=None):
    """One rank's block of the global hill problem. size = per-rank block (weak scaling: the global domain is px x py blocks), or
    global_size = the whole domain, dealt out evenly over the process grid (BASELINE.json configs[3], [4]).
    device=None: numpy on the host (hostsetup); device='cuda:N': the same formulas as torch expressions on that GPU (devsetup) --
    used for the 800x800x400-per-GPU runs, where the numpy route needs ~45 GB of host memory and minutes per rank."""
    from cgfd3d_b200 import decomp, hostsetup as hs
    px, py = proc_grid(nranks)
    ix, iy = rank // py, rank % py   # row-major, y fastest (MPI_Cart_create, forward/mympi_t.c:32-40)
    neigh = decomp.neighbours(rank, px, py)
    if global_size is None:
        ni, nj, nk = size
        gni, gnj = ni * px, nj * py
        gi0, gj0 = ix * ni, iy * nj
    else:
        gni, gnj, nk = global_size
        gi0, ni, gj0, nj = decomp.local_block(rank, px, py, gni, gnj)
    # hill spans the global domain (sigma ~ 1/10 of it, height ~ 10 cells)
    dh = (100.0, 100.0, 100.0)
    sigma = 0.1 * max(gni, gnj) * dh[0]
    sub = (gi0, gj0, gni, gnj, neigh)
    # dt below the CFL bound of the stretched grid (estimate_dt on the full array is slow; checked in tests)
    dt = 0.012 if medium == "iso" else 0.008
    # cost-attribution side runs (scripts/gpu_r2h.sh): BENCH_DIAG="pml=xz,free=0" keeps only the named PML axes / drops the free
    # surface; the line then says so in config.diag and is not a bench line of the named workload
    diag = dict(kv.split("=") for kv in os.environ.get("BENCH_DIAG", "").split(",") if "=" in kv)
    kw = {}
    if "pml" in diag:
        kw["pml_faces"] = tuple((ax, sd) for ax, sd in ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0)) if "xyz"[ax] in diag["pml"])
    free_top = diag.get("free", "1") != "0"
    if device is not None:
        from cgfd3d_b200 import devsetup
        prob = devsetup.build_problem(ni, nj, nk, device=device, dh=dh, hill=(1000.0, sigma), pml_layers=10, free_top=free_top, dt=dt, sub=sub, medium=medium, **kw)
    else:
        prob = hs.build_problem(ni, nj, nk, dh=dh, topo="hill", hill=(1000.0, sigma), pml_layers=10, free_top=free_top, dt=dt, sub=sub, medium=medium, **kw)
        if medium != "iso":
            # free-surface matrices of the other constitutive laws: computed by the library on the device (cgfd_b200_dvh2dvz)
            from cgfd3d_b200 import solver
            prob.mats = solver.dvh2dvz(prob, device=int(os.environ.get("LOCAL_RANK", "0")))
    # one explosive moment source under the hill top, on the rank that owns it
    gsi, gsj = gni // 2, gnj // 2
    if gi0 <= gsi < gi0 + ni and gj0 <= gsj < gj0 + nj:
        hs.make_source(prob, gsi - gi0, gsj - gj0, nk - 1 - 20, nt_total=100000, kind="moment",
                       mech=(1e16, 1e16, 1e16, 0, 0, 0), fc=2.0, t0=0.5, stf_len=1.0)
    return prob


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cgfd3d_b200 import nrank_check, solver

    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if nranks != args.gpus:
        if nranks == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    if nranks > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if nranks > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, max(args.warmup, 8)   # the 8 operator pairs (it % 8) are 8 kernel instantiations: touch all before timing
    px, py = proc_grid(nranks)

    # ---- N ranks reproduce 1 rank: value check on this run's own process grid, before anything is timed --------------------
    parity = None
    if nranks > 1 and not args.no_parity:
        parity = nrank_check.check(rank, nranks, local, px, py, dist, medium=args.medium)
        ok = [parity["ok"] if rank == 0 else None]
        dist.broadcast_object_list(ok, src=0)
        if not ok[0]:
            if rank == 0:
                emit({"error": "N-rank run does not reproduce the 1-rank run", "parity_nrank": parity})
            dist.destroy_process_group()
            raise SystemExit(1)

    def measure(size, global_size=None, e2e=True, sample_clocks=True):
        """device-resident (and end-to-end) Gpoint-updates/s of one problem; returns a dict"""
        t0 = time.time()
        gs = global_size
        big = (size[0] * size[1] * size[2] if gs is None else gs[0] * gs[1] * gs[2] / nranks) > 64e6
        on_device = big or args.medium != "iso"   # big blocks / many media arrays: set-up arrays built on the GPU (devsetup.py)
        prob = build_rank_problem(size, rank, nranks, device=("cuda:%d" % local) if on_device else None, medium=args.medium, global_size=gs)
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
        ni, nj, nk = prob.ni, prob.nj, prob.nk
        npts_all = (size[0] * size[1] * size[2] * nranks) if gs is None else gs[0] * gs[1] * gs[2]
        S.run(W, it0=0)
        S.set_profiling(True)
        barrier()
        clk = ClockSampler(local) if sample_clocks else None
        if clk:
            clk.__enter__()
        tw0 = time.time()
        S.run(K, it0=W)
        barrier()
        tw1 = time.time()
        if clk:
            clk.__exit__()
        ms_max = allmax(S.last_run_ms())
        main_ms, main_n, launches = S.get_profile()
        S.set_profiling(False)
        R = {"value": npts_all * K / (ms_max * 1e-3) / 1e9, "ms_per_step": ms_max / K, "main_ms": main_ms, "main_n": main_n,
             "launches": launches, "clocks": clk.summary() if clk else None, "wall": tw1 - tw0, "npts_all": npts_all,
             "block": (ni, nj, nk), "free_top": prob.free_top, "top_fused": S.top_fused(), "gz": S.grid_class(),
             "pml_slab_points": sum(int(np.prod(prob.pml_aux_shape(*f)[1:])) for f in prob.pml), "t_host": t_host, "t_upload": t_upload,
             "on_device": on_device, "e2e": []}
        if e2e:
            # ---- end to end from host buffers: the call sequence of a production run -- initial wavefield from pinned host memory,
            # Ke steps with the outputs the reference's example writes (receiver traces + a surface Vx/Vy/Vz snapshot EVERY step,
            # example/cgfd3d.example.sh:323-334) streamed to pinned host buffers while the steps run, final wavefield back to the host.
            # The state transfer is a one-off: the rate depends on the number of steps it is spread over, so two lengths are reported.
            rec = [prob.iptr(ni // 2 + 5 * n, nj // 2 + 3 * n, nk - 1) for n in range(-4, 5)]
            w_host = torch.zeros(S.shape, dtype=torch.float32).pin_memory().numpy()
            ncmp = S.shape[0]
            it0 = W + K
            for Ke in ((K,) if args.short_e2e else (K, 100)):
                S.set_record_points(rec, Ke)
                snap = torch.zeros((Ke, 3, 1, nj, ni), dtype=torch.float32).pin_memory().numpy()
                S.add_snapshot((0, 1, 2), (3, ni, 1, 3, nj, 1, prob.nz - 4, 1, 1), max_frames=Ke, it1=it0, out=snap)
                barrier()
                te0 = time.time()
                S.set_wavefield(w_host)
                te1 = time.time()
                S.run(Ke, it0=it0)
                traces = S.get_record(0, Ke)
                te2 = time.time()
                S.get_wavefield(out=w_host)
                barrier()
                te3 = time.time()
                secs = allmax(te3 - te0)
                R["e2e"].append({"steps": Ke, "value": npts_all * Ke / secs / 1e9, "h2d": w_host.nbytes / Ke,
                                 "d2h": w_host.nbytes / Ke + len(rec) * ncmp * 4 + snap.nbytes / Ke,
                                 "ms": {"set_wavefield": round((te1 - te0) * 1e3, 2), "steps_with_outputs": round((te2 - te1) * 1e3, 2),
                                        "get_wavefield": round((te3 - te2) * 1e3, 2)},
                                 "finite": bool(np.isfinite(w_host).all()) and bool(np.isfinite(snap).all()) and bool(np.isfinite(traces).all())})
                it0 += Ke
                del snap
        S.close()
        del S, prob
        torch.cuda.empty_cache()
        return R

    # ---- the measured workload --------------------------------------------------------------------------------------------
    gs = args.global_size
    size = None if gs else (args.size or ((400, 400, 200) if nranks == 1 else (800, 800, 400)))
    R = measure(size, gs, e2e=not args.no_e2e)

    # ---- weak-scaling base: the per-GPU block of the multi-GPU runs on ONE GPU, so that the efficiency of an N-GPU line can be
    # recomputed from the lines themselves. N = 1: a second problem in this process. N > 1: rank 0 starts a one-GPU bench.py on its
    # own device once the N-rank problem is released; the other ranks wait at the barrier.
    weak = None
    if not args.no_weak_base and args.medium == "iso" and gs is None and args.size is None:
        if nranks == 1:
            Rw = measure((800, 800, 400), None, e2e=False, sample_clocks=False)
            weak = {"size": "800x800x400", "value": round(Rw["value"], 4), "ms_per_step": round(Rw["ms_per_step"], 4),
                    "what": "BASELINE.json configs[2] block on one GPU (all four x-y faces CFS-PML), same process, %d steps" % K}
        else:
            if rank == 0:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                dev = vis.split(",")[local] if vis else str(local)
                env = {k: v for k, v in os.environ.items()
                       if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "LOCAL_WORLD_SIZE", "GROUP_RANK",
                                    "ROLE_RANK", "ROLE_WORLD_SIZE", "GROUP_WORLD_SIZE", "TORCHELASTIC_RUN_ID")}
                env["CUDA_VISIBLE_DEVICES"] = dev
                pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--gpus", "1", "--steps", str(K), "--warmup", str(W), "--size",
                                     "800x800x400", "--no-cpu-baseline", "--short-e2e", "--no-weak-base"], capture_output=True, text=True, env=env, timeout=900)
                try:
                    j = json.loads(pr.stdout.strip().splitlines()[-1])
                    weak = {"size": "800x800x400", "value": j["value"], "ms_per_step": j["ms_per_step"],
                            "what": "the per-GPU block of this run on ONE GPU (all four x-y faces CFS-PML), measured by rank 0 on its own device right after the timed region"}
                except Exception as e:   # noqa: BLE001
                    weak = {"error": "weak base run failed: %r %s" % (e, pr.stderr[-300:])}
            barrier()

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        main_ms, main_n = R["main_ms"], R["main_n"]
        kb = (main_ms / max(main_n, 1)) * 1e-3
        # algorithmic bytes of one interior-kernel launch = (bytes per point-step / 4 stages) x the points it covers
        bpps = BYTES_PER_POINT_STEP[args.medium]
        ni, nj, nk = R["block"]
        # a separate free-surface launch (CGFD_FUSE_TOP=0) owns the top 4 rows; by default they are planes of the interior kernel
        main_pts = ni * nj * (nk - 4 if (R["free_top"] and not R["top_fused"]) else nk)
        ach = (bpps / 4.0) * main_pts / kb / 1e9 if main_n else None
        # CFS-PML auxiliary variables: 16 x 9 floats per slab point and step, i.e. 144 B per launch on average.
        # NOT part of `achieved` / `frac` (the survey's interior formula); reported beside them
        pml_bytes = 144 * R["pml_slab_points"]
        ach_pml = ((bpps / 4.0) * main_pts + pml_bytes) / kb / 1e9 if main_n else None
        e0 = R["e2e"][0] if R["e2e"] else None
        workload = "%s, Gaussian-hill topography (curvilinear), " % WORKLOAD[args.medium]
        if gs is None:
            workload += "%dx%dx%d per GPU%s, " % (tuple(size) + (" (weak scaling)" if nranks > 1 else "",))
        else:
            workload += "%dx%dx%d global over %d GPUs (rank 0 block %dx%dx%d), " % (tuple(gs) + (nranks, ni, nj, nk))
        workload += ("CFS-PML 10 layers x 5 faces, traction-image free surface, 1 moment source" if R["free_top"]
                     else "CFS-PML 10 layers x 6 faces, 1 moment source")
        out = {
            "metric": METRIC, "value": round(R["value"], 4), "unit": "Gpoint-updates/s", "n_gpus": nranks, "steps": K, "warmup": W,
            "ms_per_step": round(R["ms_per_step"], 4), "higher_is_better": True, "scaling": "weak" if gs is None else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "proc_grid": "%dx%d" % (px, py), "l2": "working set >> 126 MB L2 (no flush needed)",
                       "variant": args.variant or "default", "medium": args.medium,
                       **({"diag": os.environ["BENCH_DIAG"]} if os.environ.get("BENCH_DIAG") else {}),
                       "kernels": "vertically-deformed-grid (4 metric arrays identically zero)" if R["gz"] else "general curvilinear"},
            "e2e": None if e0 is None else {"value": round(e0["value"], 4), "unit": "Gpoint-updates/s", "h2d_bytes_per_step": int(e0["h2d"]), "d2h_bytes_per_step": int(e0["d2h"]),
                    "steps": e0["steps"],
                    "what": "set_wavefield (H2D) + run(steps) with receiver traces and a surface Vx/Vy/Vz snapshot streamed to pinned host "
                            "memory every step + get_wavefield (D2H); the whole-wavefield transfers are a one-off, so the rate grows with "
                            "the number of steps they are spread over (e2e_k100 = the same over 100 steps)", "ms": e0["ms"]},
            "gpu_launches": int(R["launches"]),
            "clocks": R["clocks"],
            "roofline": {"bound": "hbm", "kernel": "k_main_tma", "achieved": None if ach is None else round(ach, 1), "peak": peak,
                         "unit": "GB/s", "frac": None if ach is None else round(ach / peak, 4), "traffic": ncu_traffic(R["block"], args.medium),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "launches_timed": int(main_n), "avg_launch_ms": round(main_ms / max(main_n, 1), 4),
                         "algorithmic_bytes_per_launch": int((bpps / 4.0) * main_pts),
                         "free_surface_rows": "planes of k_main_tma's top z chunk" if R["top_fused"] else ("separate k_top launch" if R["free_top"] else "none"),
                         "pml_aux_bytes_per_launch": int(pml_bytes), "frac_incl_pml_aux": None if ach_pml is None else round(ach_pml / peak, 4),
                         "whole_step_frac": round(bpps * (R["npts_all"] / nranks) / (R["ms_per_step"] * 1e-3) / 1e9 / peak, 4)},
            "setup_s": {"arrays": round(R["t_host"], 2), "arrays_built_on": "gpu (torch)" if R["on_device"] else "host (numpy)", "upload": round(R["t_upload"], 2)},
            "wall_s_timed": round(R["wall"], 4), "finite": all(e["finite"] for e in R["e2e"]),
        }
        if len(R["e2e"]) > 1:
            e1 = R["e2e"][1]
            out["e2e_k100"] = {"value": round(e1["value"], 4), "steps": e1["steps"], "h2d_bytes_per_step": int(e1["h2d"]),
                               "d2h_bytes_per_step": int(e1["d2h"]), "ms": e1["ms"]}
        if nranks > 1:
            out["per_gpu"] = round(R["value"] / nranks, 4)
            out["parity_nrank"] = parity
        if weak is not None:
            out["weak_base"] = weak
            if nranks > 1 and "value" in weak:
                out["efficiency_vs_weak_base"] = round(R["value"] / nranks / weak["value"], 4)
        if nranks == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(os.cpu_count() or 1, steps=args.cpu_steps)
    if nranks > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def ncu_traffic(size, medium="iso"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json, one entry per medium), valid for the workload it was captured on; None otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(medium)
        return e["bytes_per_launch"] if e and list(size) == list(e["size"]) else None
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


def _ref_program_step_stamps(size, px, py, nt, hill):
    """one run of nt steps with the start of every step stamped on arrival of the driver's '-> it=' line (oracle/harness.run_timed);
    returns (wall, {step: seconds}, outputs finite)"""
    import shutil
    import tempfile
    from oracle import harness as H
    wd = tempfile.mkdtemp(prefix="cgfd_refarm_")
    try:
        H.write_multirank_hill_case(wd, size, px, py, nt, REF_DT, hill=hill, pml_layers=10, src=H.moment_src(size[0] // 2, size[1] // 2, 20))
        env = dict(os.environ, CGFD_SHIM_NPROCS=str(px * py), CGFD_SHIM_MSG_MB="64")
        wall, stamps = H.run_timed(H.ref_binary("ref_main_zero"), wd, timeout=3000, env=env)
        sac = H.read_sac_dir(os.path.join(wd, "OUT"))
        finite = all(bool(np.isfinite(v).all()) for v in sac.values()) and len(sac) > 0
        return wall, stamps, finite
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def cpu_baseline(cores, steps=4, warm=1, size=REF_SIZE, hill=REF_HILL):
    """Gpoint-updates/s of the reference program on `cores` ranks over `steps` RK4 steps after `warm` untimed ones. One run of
    warm + steps + 1 steps (number_of_time_steps = warm + steps) whose driver reports the start of every step (verbose > 10, forward/drv_rk_curv_col.c:172): the time
    between the start of step `warm` and the start of step `warm + steps`, i.e. the time loop alone. Where the lines cannot be
    stamped on arrival: (wall of a run of warm + steps steps) - (wall of a run of warm steps) (the program's own timer has a
    resolution of one second)."""
    from oracle import harness as H
    if not H.have_ref("ref_main_zero"):
        return {"value": None, "unit": "Gpoint-updates/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    px, py = _rank_grid(cores)
    npts = size[0] * size[1] * size[2]
    # the reference runs number_of_time_steps + 1 steps (forward/main_curv_col_el_3d.c:615): steps 0 .. warm + steps start, the last one is untimed
    wall, stamps, ok = _ref_program_step_stamps(size, px, py, warm + steps, hill)
    if warm in stamps and warm + steps in stamps:
        secs = max(stamps[warm + steps] - stamps[warm], 1e-6)
        per = sorted(stamps[n + 1] - stamps[n] for n in range(warm, warm + steps))
        return {"value": round(npts * steps / secs / 1e9, 6), "unit": "Gpoint-updates/s", "cores": cores, "kind": "reference",
                "ranks": "%dx%d" % (px, py), "seconds_per_step": round(secs / steps, 4), "seconds_per_step_min_max": [round(per[0], 4), round(per[-1], 4)],
                "setup_s": round(stamps[0], 2), "finite": bool(ok),
                "sample": "%d RK4 steps of the bench workload itself (isotropic, Gaussian hill %dx%dx%d through gd_curv_coord_import, CFS-PML 10x5, "
                          "free surface, 1 moment source) by the unmodified reference program on %dx%d ranks with halo exchange "
                          "(fork/shared-memory MPI stand-in: no MPI in the image); time = start of step %d to start of step %d of one run, "
                          "from the driver's own per-step lines stamped on arrival"
                          % ((steps,) + tuple(size) + (px, py, warm, warm + steps))}
    t_w, ok1 = _ref_program_seconds(size, px, py, warm, hill)
    t_wk, ok2 = _ref_program_seconds(size, px, py, warm + steps, hill)
    secs = max(t_wk - t_w, 1e-6)
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
    ap.add_argument("--size", type=parse_size, default=None, help="per-GPU block NIxNJxNK (weak scaling)")
    ap.add_argument("--global-size", type=parse_size, default=None, help="whole domain NIxNJxNK dealt out over the GPUs (BASELINE.json configs[3], [4])")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the N-rank-vs-1-rank value check before timing")
    ap.add_argument("--no-weak-base", action="store_true", help="skip the one-GPU measurement of the weak-scaling block")
    ap.add_argument("--short-e2e", action="store_true", help="only the K-step end-to-end leg (not the 100-step one)")
    ap.add_argument("--no-e2e", action="store_true", help="side measurements of very large blocks: skip the end-to-end leg (no pinned host copy of the state)")
    ap.add_argument("--variant", default="")
    ap.add_argument("--medium", default="iso", choices=["iso", "vti", "aniso", "visco"],
                    help="constitutive law (default iso = the BASELINE.json metric; the others are side measurements)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=8)
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
