"""Harness for running the compiled reference (oracle/_ref) and reading its outputs.

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.

It writes the reference's own input formats (par JSON, .src, .station — formats from
/root/reference/example/cgfd3d.example.sh:77-379, parser forward/par_t.c:84-1044), runs a
binary built by oracle/Makefile, and parses SAC seismograms (lib/sacLib.c:343-377: 632-byte
header + npts float32) and the dense "CGNC1" container our NetCDF stand-in writes
(oracle/shims/netcdf_shim.c).
"""
from __future__ import annotations

import json
import os
import struct
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def ref_binary(name: str = "ref_main_zero") -> str:
    return os.path.join(REF_DIR, name)


def have_ref(name: str = "ref_main_zero") -> bool:
    return os.path.isfile(ref_binary(name))


CFSPML = {"number_of_layers": 10, "alpha_max": 3.14, "beta_max": 2.0, "ref_vel": 7000.0}


def make_par(
    workdir: str,
    nx: int,
    ny: int,
    nz: int,
    nt: int,
    dt: float,
    *,
    pml_layers: int = 10,
    pml_sides=("x_left", "x_right", "y_front", "y_back", "z_bottom"),
    ablexp_sides=(),
    free_top: bool = True,
    grid: dict | None = None,
    medium_type: str = "elastic_iso",
    visco: dict | None = None,
    src_spatial: str = "point",
    lines: list | None = None,
    snapshots: list | None = None,
    slices: dict | None = None,
    export: bool = False,
    check_stability: int = 1,
    ddsource: dict | None = None,
) -> dict:
    """Build the par dict. Paths are relative to workdir/cwd."""
    par = {
        "number_of_total_grid_points_x": nx,
        "number_of_total_grid_points_y": ny,
        "number_of_total_grid_points_z": nz,
        "number_of_mpiprocs_x": 1,
        "number_of_mpiprocs_y": 1,
        "size_of_time_step": dt,
        "number_of_time_steps": nt,
        "check_stability": check_stability,
    }
    names = {
        "x_left": "boundary_x_left", "x_right": "boundary_x_right",
        "y_front": "boundary_y_front", "y_back": "boundary_y_back",
        "z_bottom": "boundary_z_bottom", "z_top": "boundary_z_top",
    }
    for s in pml_sides:
        c = dict(CFSPML)
        c["number_of_layers"] = pml_layers
        par[names[s]] = {"cfspml": c}
    for s in ablexp_sides:
        par[names[s]] = {"ablexp": {"number_of_layers": pml_layers, "ref_vel": 7000.0}}
    if free_top:
        par["boundary_z_top"] = {"free": "timg"}
    if grid is None:
        grid = {"cartesian": {"origin": [0.0, 0.0, -(nz - 1) * 100.0], "inteval": [100.0, 100.0, 100.0]}}
    par["grid_generation_method"] = grid
    par["is_export_grid"] = 1 if export else 0
    par["grid_export_dir"] = "OUT"
    par["metric_calculation_method"] = {"calculate": 1}
    par["is_export_metric"] = 1 if export else 0
    par["medium"] = {"type": medium_type, "input_way": "code", "code": "x", "equivalent_medium_method": "loc"}
    if visco is not None:
        par["visco_config"] = visco
    par["is_export_media"] = 1 if export else 0
    par["media_export_dir"] = "OUT"
    par["in_source_file"] = "case.src"
    par["is_export_source"] = 0
    par["source_export_dir"] = "OUT"
    par["source_spatial_distribution_type"] = src_spatial
    par["output_dir"] = "OUT"
    par["tmp_dir"] = "OUT"
    par["in_station_file"] = "case.station"
    if lines:
        par["receiver_line"] = lines
    if snapshots:
        par["snapshot"] = snapshots
    if slices:
        par["slice"] = slices
    if ddsource:
        # distributed (finite-fault) sources: nc file of write_ddsource(), read block-wise by src_dd_read2local / src_dd_accit_loadstf
        par["in_ddsource_file"] = ddsource.get("file", "case_dd.nc")
        par["ddsource_add_at_point"] = 1
        par["ddsource_nt_per_read"] = int(ddsource.get("nt_per_read", 100))
    par["check_nan_every_nummber_of_steps"] = 0
    par["output_all"] = 0
    return par


def write_case(workdir: str, par: dict, src_text: str, stations: list, coords=None) -> str:
    """stations: list of (name, is_coord, is_depth, x, y, z). coords = (x, y, z) arrays [nz][ny][nx] incl. ghosts: written as
    IN/coord_px0_py0.nc for "grid_generation_method": {"import": "IN"} (gd_curv_coord_import, forward/gd_t.c:741-783)."""
    os.makedirs(os.path.join(workdir, "OUT"), exist_ok=True)
    if coords is not None:
        os.makedirs(os.path.join(workdir, "IN"), exist_ok=True)
        x, y, z = coords
        nz, ny, nx = x.shape
        write_cgnc(os.path.join(workdir, "IN", "coord_px0_py0.nc"), {"k": nz, "j": ny, "i": nx},
                   {"x": (("k", "j", "i"), x), "y": (("k", "j", "i"), y), "z": (("k", "j", "i"), z)})
    with open(os.path.join(workdir, "case.json"), "w") as f:
        json.dump(par, f, indent=1)
    with open(os.path.join(workdir, "case.src"), "w") as f:
        f.write(src_text)
    with open(os.path.join(workdir, "case.station"), "w") as f:
        f.write("%d\n" % len(stations))
        for s in stations:
            f.write("%s %d %d %g %g %g\n" % tuple(s))
    return os.path.join(workdir, "case.json")


def moment_src(i, j, depth_k, m=(1e16, 1e16, 1e16, 0, 0, 0), f0=2.0, t0=0.5, stf_len=1.0) -> str:
    """One moment-tensor point source by grid index, Ricker STF (example.sh:79-112)."""
    return (
        "evt1\n1\n0 %g\n2 0\n0 1\n%d %d %d\n0.0 ricker %g %g\n%g %g %g %g %g %g\n"
        % ((stf_len, i, j, depth_k, f0, t0) + tuple(m))
    )


def force_src(i, j, depth_k, fvec=(0.0, 0.0, 1e16), f0=2.0, t0=0.5, stf_len=1.0) -> str:
    """One force point source by grid index, Ricker STF."""
    return (
        "evt1\n1\n0 %g\n1 0\n0 1\n%d %d %d\n0.0 ricker %g %g\n%g %g %g\n"
        % ((stf_len, i, j, depth_k, f0, t0) + tuple(fvec))
    )


def run(binary: str, workdir: str, verbose: int = 1, timeout: float = 3600, env=None) -> tuple[float, str]:
    """Run `<binary> case.json <verbose>` in workdir. Returns (wall seconds, stdout)."""
    t0 = time.time()
    p = subprocess.run(
        [binary, "case.json", str(verbose)], cwd=workdir, capture_output=True, text=True, timeout=timeout, env=env
    )
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(
            "%s failed (rc=%d)\nstdout tail:\n%s\nstderr tail:\n%s"
            % (binary, p.returncode, p.stdout[-3000:], p.stderr[-3000:])
        )
    return wall, p.stdout


def run_timed(binary: str, workdir: str, timeout: float = 3600, env=None):
    """Run `<binary> case.json 11` (verbose > 10: the reference driver prints '-> it=<n>, t=...' when a step STARTS,
    forward/drv_rk_curv_col.c:172) with line-buffered stdout and stamp every such line as it arrives. Returns (wall seconds,
    {step: seconds since program start}); the dict is empty when the lines could not be stamped (no stdbuf: the pipe is then block
    buffered and everything arrives at exit)."""
    import re
    import shutil
    stdbuf = shutil.which("stdbuf")
    cmd = ([stdbuf, "-oL"] if stdbuf else []) + [binary, "case.json", "11"]
    t0 = time.time()
    stamps = {}
    with open(os.path.join(workdir, "stderr.log"), "w") as ferr:
        p = subprocess.Popen(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=ferr, text=True, env=env)
        tail = []
        try:
            for line in p.stdout:
                m = re.match(r"-> it=(\d+),", line)
                if m:
                    stamps.setdefault(int(m.group(1)), time.time() - t0)
                tail.append(line)
                if len(tail) > 200:
                    del tail[:100]
                if time.time() - t0 > timeout:
                    p.kill()
                    raise RuntimeError("%s: timeout after %.0f s" % (binary, timeout))
            rc = p.wait()
        finally:
            if p.poll() is None:
                p.kill()
    wall = time.time() - t0
    if rc != 0:
        raise RuntimeError("%s failed (rc=%d)\nstdout tail:\n%s\nstderr tail:\n%s"
                           % (binary, rc, "".join(tail)[-3000:], open(os.path.join(workdir, "stderr.log")).read()[-3000:]))
    if not stdbuf or len(stamps) < 2 or max(stamps.values()) - min(stamps.values()) < 1e-3:
        stamps = {}
    return wall, stamps


def read_sac(path: str) -> np.ndarray:
    return np.fromfile(path, dtype="<f4", offset=632)


def read_sac_dir(outdir: str) -> dict:
    out = {}
    for fn in sorted(os.listdir(outdir)):
        if fn.endswith(".sac"):
            out[fn[:-4]] = read_sac(os.path.join(outdir, fn))
    return out


def read_cgnc(path: str) -> dict:
    """Parse the CGNC1 container (netcdf_shim.c). Returns {'dims':{}, 'vars':{name: ndarray}, 'atts':{}}."""
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:6] == b"CGNC1\n", "%s is not a CGNC1 container" % path
    pos = 6

    def i32():
        nonlocal pos
        v = struct.unpack_from("<i", buf, pos)[0]
        pos += 4
        return v

    def i64():
        nonlocal pos
        v = struct.unpack_from("<q", buf, pos)[0]
        pos += 8
        return v

    def s():
        nonlocal pos
        n = i32()
        v = buf[pos:pos + n].decode()
        pos += n
        return v

    def atts():
        nonlocal pos
        out = {}
        for _ in range(i32()):
            name = s()
            n = i32()
            out[name] = np.frombuffer(buf, "<i4", n, pos).copy()
            pos += 4 * n
        return out

    dims = []
    for _ in range(i32()):
        name = s()
        ln = i64()
        unl = i32()
        dims.append((name, ln, unl))
    res = {"dims": {d[0]: d[1] for d in dims}, "vars": {}, "var_atts": {}}
    for _ in range(i32()):
        name = s()
        typ = i32()
        nd = i32()
        dimids = [i32() for _ in range(nd)]
        res["var_atts"][name] = atts()
        n = i64()
        dt = "<f4" if typ == 5 else "<i4"
        arr = np.frombuffer(buf, dt, n, pos).copy()
        pos += 4 * n
        shape = [dims[d][1] for d in dimids]
        res["vars"][name] = arr.reshape(shape) if nd else arr
    res["atts"] = atts()
    return res


def write_cgnc(path: str, dims: dict, variables: dict, atts: dict | None = None) -> None:
    """Write a CGNC1 container the shim's nc_open can import.
    dims: ordered {name: len}; variables: {name: (dim-name tuple, float32 ndarray)}."""
    names = list(dims)
    with open(path, "wb") as f:
        f.write(b"CGNC1\n")
        f.write(struct.pack("<i", len(names)))
        for n in names:
            b = n.encode()
            f.write(struct.pack("<i", len(b)) + b + struct.pack("<qi", dims[n], 0))
        f.write(struct.pack("<i", len(variables)))
        for vn, (vd, arr) in variables.items():
            b = vn.encode()
            arr = np.ascontiguousarray(arr, dtype="<f4")
            f.write(struct.pack("<i", len(b)) + b + struct.pack("<ii", 5, len(vd)))
            for d in vd:
                f.write(struct.pack("<i", names.index(d)))
            f.write(struct.pack("<i", 0))
            f.write(struct.pack("<q", arr.size))
            f.write(arr.tobytes())
        atts = atts or {}
        f.write(struct.pack("<i", len(atts)))
        for an, av in atts.items():
            b = an.encode()
            av = np.asarray(av, dtype="<i4").ravel()
            f.write(struct.pack("<i", len(b)) + b + struct.pack("<i", av.size) + av.tobytes())


def write_multirank_hill_case(workdir: str, size, px: int, py: int, nt: int, dt: float, hill=(1000.0, 2000.0), pml_layers: int = 10,
                              src=None, lines=None, snapshots=None, medium_type: str = "elastic_iso", visco: dict | None = None,
                              pml_sides=("x_left", "x_right", "y_front", "y_back", "z_bottom"), ablexp_sides=()) -> str:
    """A px x py-rank run of the reference program on a Gaussian-hill grid (BASELINE.json configs[1] / [2] physics): the grid goes
    in through gd_curv_coord_import, one coord_px?_py?.nc per rank holding that rank's block incl. ghosts, blocks dealt out by
    gd_indx_set's rule (cgfd3d_b200.decomp.ref_split). Run it with CGFD_SHIM_NPROCS = px * py (oracle/shims/mpi_shim.c)."""
    from cgfd3d_b200 import decomp, hostsetup as hs
    ni, nj, nk = size
    par = make_par(workdir, ni, nj, nk, nt, dt, pml_layers=pml_layers, grid={"import": "IN"}, lines=lines, snapshots=snapshots,
                   medium_type=medium_type, visco=visco, pml_sides=pml_sides, ablexp_sides=ablexp_sides)
    par["number_of_mpiprocs_x"], par["number_of_mpiprocs_y"] = px, py
    os.makedirs(os.path.join(workdir, "IN"), exist_ok=True)
    sides = set(pml_sides) | set(ablexp_sides)
    lx = (pml_layers if "x_left" in sides else 0, pml_layers if "x_right" in sides else 0)
    ly = (pml_layers if "y_front" in sides else 0, pml_layers if "y_back" in sides else 0)
    for ix in range(px):
        gi0, lni = decomp.ref_split(ni, px, ix, *lx)
        for iy in range(py):
            gj0, lnj = decomp.ref_split(nj, py, iy, *ly)
            x, y, z = hs.hill_coords(lni, lnj, nk, height=hill[0], sigma=hill[1], gi0=gi0, gj0=gj0, gni=ni, gnj=nj)
            nz_, ny_, nx_ = x.shape
            write_cgnc(os.path.join(workdir, "IN", "coord_px%d_py%d.nc" % (ix, iy)), {"k": nz_, "j": ny_, "i": nx_},
                       {"x": (("k", "j", "i"), x), "y": (("k", "j", "i"), y), "z": (("k", "j", "i"), z)})
    write_case(workdir, par, src or moment_src(ni // 2, nj // 2, 20), [("r1", 0, 1, ni // 2 + 10, nj // 2 + 5, 0)])
    return os.path.join(workdir, "case.json")


def write_ddsource(path: str, t, xyz, force=None, moment_rate=None) -> None:
    """The distributed-source input of src_dd_read2local (forward/src_t.c:1181-1927): dims time / number, global attributes
    location_is_axis = 0 (grid index) and z_is_depth = 1, vars time(time), x / y / z(number) and Fx Fy Fz (number, time) and / or
    Mxx_rate Myy_rate Mzz_rate Myz_rate Mxz_rate Mxy_rate (number, time).
    t [nt]; xyz [n][3] = (i, j, depth below the surface) in grid indices; force [n][3][nt]; moment_rate [n][6][nt] (xx yy zz yz xz xy)."""
    xyz = np.asarray(xyz, np.float32)
    n, nt = xyz.shape[0], len(t)
    var = {"time": (("time",), np.asarray(t, np.float32)),
           "x": (("number",), xyz[:, 0].copy()), "y": (("number",), xyz[:, 1].copy()), "z": (("number",), xyz[:, 2].copy())}
    if force is not None:
        for c, name in enumerate(("Fx", "Fy", "Fz")):
            var[name] = (("number", "time"), np.ascontiguousarray(force[:, c, :], np.float32))
    if moment_rate is not None:
        for c, name in enumerate(("Mxx_rate", "Myy_rate", "Mzz_rate", "Myz_rate", "Mxz_rate", "Mxy_rate")):
            var[name] = (("number", "time"), np.ascontiguousarray(moment_rate[:, c, :], np.float32))
    write_cgnc(path, {"time": nt, "number": n}, var, {"location_is_axis": [0], "z_is_depth": [1]})


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 (b = reference). 0 if both are identically zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    if nb == 0.0:
        return 0.0 if np.linalg.norm(a) == 0.0 else float("inf")
    return float(np.linalg.norm(a - b) / nb)
