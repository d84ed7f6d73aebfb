"""Python binding of the C ABI (include/cgfd3d_b200.h) through ctypes.

`Solver` mirrors what the reference's driver does with its structs
(drv_rk_curv_col_allstep, forward/drv_rk_curv_col.c:27-556): upload once, run RK4 steps on the
device, feed the output taps. There is no CPU path: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# CGFD_LIB selects another in-tree build of the same library (tile-shape experiments, scripts/gpu_tiles.sh)
LIB_PATH = os.environ.get("CGFD_LIB") or os.path.join(_HERE, "libcgfd3d_b200.so")

SYMBOLS = [
    "cgfd_b200_last_error", "cgfd_b200_abi_version", "cgfd_b200_sizeof_problem", "cgfd_b200_device_count", "cgfd_b200_create", "cgfd_b200_destroy",
    "cgfd_b200_set_wavefield", "cgfd_b200_get_wavefield", "cgfd_b200_set_pml_aux", "cgfd_b200_get_pml_aux",
    "cgfd_b200_pml_aux_size", "cgfd_b200_onestage", "cgfd_b200_get_pml_aux_rhs", "cgfd_b200_run",
    "cgfd_b200_set_record_points", "cgfd_b200_get_record", "cgfd_b200_get_box", "cgfd_b200_get_pg",
    "cgfd_b200_comm_unique_id", "cgfd_b200_comm_init", "cgfd_b200_halo_plan", "cgfd_b200_set_profiling", "cgfd_b200_get_profile",
    "cgfd_b200_last_run_ms", "cgfd_b200_set_variant", "cgfd_b200_grid_class", "cgfd_b200_top_fused",
    "cgfd_b200_add_snapshot", "cgfd_b200_snapshot_frames", "cgfd_b200_dd_set_points", "cgfd_b200_dd_load_block",
    "cgfd_b200_metric_from_coords", "cgfd_b200_launch_plan", "cgfd_b200_dvh2dvz",
    "cgfd_b200_run_async", "cgfd_b200_sync", "cgfd_b200_wait_block", "cgfd_b200_snapshot_set_output", "cgfd_b200_host_alloc", "cgfd_b200_host_free",
]

_lib = None


class CgfdError(RuntimeError):
    pass


def load_library():
    """Load libcgfd3d_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CgfdError("%s is missing: build it with `make -C cgfd3d_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ci, fp = C.c_void_p, C.c_int, abi.fptr
    L.cgfd_b200_last_error.restype = C.c_char_p
    L.cgfd_b200_sizeof_problem.restype = C.c_size_t
    L.cgfd_b200_create.argtypes = [C.POINTER(abi.Problem), ci, C.POINTER(vp)]
    L.cgfd_b200_destroy.argtypes = [vp]
    L.cgfd_b200_destroy.restype = None
    L.cgfd_b200_set_wavefield.argtypes = [vp, fp]
    L.cgfd_b200_get_wavefield.argtypes = [vp, fp]
    L.cgfd_b200_set_pml_aux.argtypes = [vp, ci, ci, fp]
    L.cgfd_b200_get_pml_aux.argtypes = [vp, ci, ci, fp]
    L.cgfd_b200_get_pml_aux_rhs.argtypes = [vp, ci, ci, fp]
    L.cgfd_b200_pml_aux_size.argtypes = [vp, ci, ci]
    L.cgfd_b200_pml_aux_size.restype = C.c_size_t
    L.cgfd_b200_onestage.argtypes = [vp, ci, ci, ci, fp, fp]
    L.cgfd_b200_run.argtypes = [vp, ci, ci]
    L.cgfd_b200_set_record_points.argtypes = [vp, ci, C.POINTER(C.c_int64), ci]
    L.cgfd_b200_get_record.argtypes = [vp, ci, ci, fp]
    L.cgfd_b200_get_box.argtypes = [vp] + [ci] * 10 + [fp]
    L.cgfd_b200_get_pg.argtypes = [vp, fp]
    L.cgfd_b200_comm_unique_id.argtypes = [C.c_char_p]
    L.cgfd_b200_comm_init.argtypes = [vp, C.c_char_p, ci, ci]
    L.cgfd_b200_halo_plan.argtypes = [C.POINTER(abi.Grid), ci, ci, ci, C.POINTER(ci * 6), C.POINTER(ci * 6)]
    L.cgfd_b200_set_profiling.argtypes = [vp, ci]
    L.cgfd_b200_get_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.cgfd_b200_last_run_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.cgfd_b200_set_variant.argtypes = [vp, C.c_char_p]
    L.cgfd_b200_grid_class.argtypes = [vp]
    L.cgfd_b200_top_fused.argtypes = [vp]
    L.cgfd_b200_add_snapshot.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(ci * 9), ci, ci, ci, fp]
    L.cgfd_b200_snapshot_frames.argtypes = [vp, ci]
    L.cgfd_b200_dd_set_points.argtypes = [vp, ci, C.POINTER(C.c_int64), ci, ci, ci, ci]
    L.cgfd_b200_dd_load_block.argtypes = [vp, ci, ci, fp, fp]
    L.cgfd_b200_launch_plan.argtypes = [C.POINTER(abi.Grid), C.POINTER((ci * 2) * 3), ci, ci, ci, C.POINTER(ci * 4), C.POINTER(ci), C.POINTER(ci), ci]
    L.cgfd_b200_metric_from_coords.argtypes = [ci, C.POINTER(abi.Grid), fp, fp, fp, ci, C.POINTER(ci), fp, C.POINTER(fp * 10)]
    L.cgfd_b200_run_async.argtypes = [vp, ci, ci, fp]
    L.cgfd_b200_sync.argtypes = [vp]
    L.cgfd_b200_wait_block.argtypes = [vp, ci]
    L.cgfd_b200_snapshot_set_output.argtypes = [vp, ci, fp, ci]
    L.cgfd_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.cgfd_b200_host_free.argtypes = [vp]
    L.cgfd_b200_host_free.restype = None
    L.cgfd_b200_dvh2dvz.argtypes = [ci, C.POINTER(abi.Problem), fp, fp, fp, ci, C.POINTER(ci), fp, fp, fp, fp, fp]
    _lib = L
    return L


def device_count() -> int:
    return load_library().cgfd_b200_device_count()


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(abi.fptr)


class Solver:
    """One subdomain resident on one GPU."""

    def __init__(self, prob, device: int = 0):
        self.L = load_library()
        self.prob = prob
        self._c = prob.to_c()
        h = C.c_void_p()
        self._chk(self.L.cgfd_b200_create(C.byref(self._c), device, C.byref(h)))
        self.h = h
        self.ncmp = prob.ncmp
        self.shape = (self.ncmp, prob.nz, prob.ny, prob.nx)
        self.nrec = 0

    def _chk(self, rc):
        if rc != 0:
            raise CgfdError(self.L.cgfd_b200_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.cgfd_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state
    def set_wavefield(self, w):
        w = np.ascontiguousarray(w, np.float32)
        assert w.shape == self.shape
        self._chk(self.L.cgfd_b200_set_wavefield(self.h, _f(w)))

    def get_wavefield(self, out=None):
        w = np.empty(self.shape, np.float32) if out is None else out
        self._chk(self.L.cgfd_b200_get_wavefield(self.h, _f(w)))
        return w

    def pml_aux_size(self, idim, iside):
        return self.L.cgfd_b200_pml_aux_size(self.h, idim, iside)

    def set_pml_aux(self, idim, iside, aux):
        aux = np.ascontiguousarray(aux, np.float32)
        assert aux.size == self.pml_aux_size(idim, iside)
        self._chk(self.L.cgfd_b200_set_pml_aux(self.h, idim, iside, _f(aux)))

    def get_pml_aux(self, idim, iside):
        out = np.empty(self.pml_aux_size(idim, iside), np.float32)
        self._chk(self.L.cgfd_b200_get_pml_aux(self.h, idim, iside, _f(out)))
        return out

    def get_pml_aux_rhs(self, idim, iside):
        out = np.empty(self.pml_aux_size(idim, iside), np.float32)
        self._chk(self.L.cgfd_b200_get_pml_aux_rhs(self.h, idim, iside, _f(out)))
        return out

    # -- compute
    def onestage(self, it, ipair, istage, w_cur):
        w_cur = np.ascontiguousarray(w_cur, np.float32)
        assert w_cur.shape == self.shape
        rhs = np.empty(self.shape, np.float32)
        self._chk(self.L.cgfd_b200_onestage(self.h, it, ipair, istage, _f(w_cur), _f(rhs)))
        return rhs

    def run(self, nsteps, it0=0):
        self._chk(self.L.cgfd_b200_run(self.h, it0, nsteps))

    def run_async(self, nsteps, it0=0, rec_out=None):
        """enqueue nsteps steps and return; rec_out [nsteps][ncmp][nrec] receives their record samples (see sync())"""
        self._chk(self.L.cgfd_b200_run_async(self.h, it0, nsteps, abi.fptr() if rec_out is None else _f(rec_out)))

    def sync(self):
        self._chk(self.L.cgfd_b200_sync(self.h))

    def wait_block(self, it_last):
        self._chk(self.L.cgfd_b200_wait_block(self.h, it_last))

    def snapshot_set_output(self, sid, out):
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        self._chk(self.L.cgfd_b200_snapshot_set_output(self.h, sid, _f(out), out.shape[0]))
        self._snap_keep = getattr(self, "_snap_keep", []) + [out]

    # -- taps
    def set_record_points(self, iptr, max_nt):
        idx = np.ascontiguousarray(iptr, np.int64)
        self._chk(self.L.cgfd_b200_set_record_points(self.h, len(idx), idx.ctypes.data_as(C.POINTER(C.c_int64)), max_nt))
        self.nrec = len(idx)

    def get_record(self, it_first, nt):
        out = np.empty((nt, self.ncmp, self.nrec), np.float32)
        self._chk(self.L.cgfd_b200_get_record(self.h, it_first, nt, _f(out)))
        return out

    def get_box(self, icmp, i1, ni, di, j1, nj, dj, k1, nk, dk, out=None):
        o = np.empty((nk, nj, ni), np.float32) if out is None else out
        self._chk(self.L.cgfd_b200_get_box(self.h, icmp, i1, ni, di, j1, nj, dj, k1, nk, dk, _f(o)))
        return o

    def dd_set_points(self, indx, vi_actived, mij_actived, nt_per_block, max_stage=4):
        """distributed-source points (flat host indices) and the size of a time-function block (include/cgfd3d_b200.h)"""
        a = np.ascontiguousarray(indx, np.int64)
        self._chk(self.L.cgfd_b200_dd_set_points(self.h, len(a), a.ctypes.data_as(C.POINTER(C.c_int64)), int(vi_actived), int(mij_actived),
                                                 max_stage, nt_per_block))

    def dd_load_block(self, it_first, vi=None, mij=None):
        """time functions of steps it_first .. it_first+nt-1: vi[nt][stage][n][3], mij[nt][stage][n][6]"""
        vi = None if vi is None else np.ascontiguousarray(vi, np.float32)
        mij = None if mij is None else np.ascontiguousarray(mij, np.float32)
        nt = (vi if vi is not None else mij).shape[0]
        self._chk(self.L.cgfd_b200_dd_load_block(self.h, it_first, nt, abi.as_f(vi), abi.as_f(mij)))

    def add_snapshot(self, cmps, box, max_frames, it1=0, tinv=1, out=None):
        """Stream frames of the strided sub-box `box` = (i1, ni, di, j1, nj, dj, k1, nk, dk) of components `cmps` to host
        memory while run() advances (include/cgfd3d_b200.h). Returns (id, out) with out[frame][cmp][nk][nj][ni]."""
        ni, nj, nk = box[1], box[4], box[7]
        if out is None:
            out = np.zeros((max_frames, len(cmps), nk, nj, ni), np.float32)
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.size == max_frames * len(cmps) * nk * nj * ni
        cm = (C.c_int * len(cmps))(*cmps)
        bx = (C.c_int * 9)(*box)
        sid = self.L.cgfd_b200_add_snapshot(self.h, len(cmps), cm, C.byref(bx), it1, tinv, max_frames, _f(out))
        if sid < 0:
            raise CgfdError(self.L.cgfd_b200_last_error().decode())
        self._snap_keep = getattr(self, "_snap_keep", []) + [out]   # the library writes into it during run()
        return sid, out

    def snapshot_frames(self, sid):
        return int(self.L.cgfd_b200_snapshot_frames(self.h, sid))

    def get_pg(self):
        out = np.empty((15, self.prob.ny, self.prob.nx), np.float32)
        self._chk(self.L.cgfd_b200_get_pg(self.h, _f(out)))
        return out

    # -- multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        self._chk(self.L.cgfd_b200_comm_init(self.h, unique_id, rank, nranks))

    # -- measurement
    def set_profiling(self, on=True):
        self._chk(self.L.cgfd_b200_set_profiling(self.h, 1 if on else 0))

    def get_profile(self):
        ms, n1, n2 = C.c_double(), C.c_int64(), C.c_int64()
        self._chk(self.L.cgfd_b200_get_profile(self.h, C.byref(ms), C.byref(n1), C.byref(n2)))
        return ms.value, n1.value, n2.value

    def last_run_ms(self):
        ms = C.c_double()
        self._chk(self.L.cgfd_b200_last_run_ms(self.h, C.byref(ms)))
        return ms.value

    def grid_class(self) -> int:
        """1: kernels specialised for vertically deformed grids (four metric arrays identically zero) are in use."""
        return int(self.L.cgfd_b200_grid_class(self.h))

    def top_fused(self) -> int:
        """1: the free-surface rows are planes of the interior kernel (no separate launch)."""
        return int(self.L.cgfd_b200_top_fused(self.h))

    def set_variant(self, name: str):
        self._chk(self.L.cgfd_b200_set_variant(self.h, name.encode()))


def metric_from_coords(grid: dict, x, y, z, fd_indx=None, fd_coef=None, device=0):
    """jac, xi_x .. zeta_z (list of 10 arrays [nz][ny][nx]) from the coordinates, computed on the GPU by the library
    (gd_curv_metric_cal, forward/gd_t.c:190-402). Default operator: fd->fdc_indx / fdc_coef of fd_set_macdrp (forward/fd_t.c:292-301)."""
    from . import hostsetup
    fd_indx = hostsetup.FDC_INDX if fd_indx is None else fd_indx
    fd_coef = hostsetup.FDC_COEF if fd_coef is None else fd_coef
    L = load_library()
    g = abi.Grid(**grid)
    x, y, z = (np.ascontiguousarray(a, np.float32) for a in (x, y, z))
    out = [np.empty_like(x) for _ in range(10)]
    po = (abi.fptr * 10)(*[_f(o) for o in out])
    ii = (C.c_int * len(fd_indx))(*fd_indx)
    cc = np.asarray(fd_coef, np.float32)
    if L.cgfd_b200_metric_from_coords(device, C.byref(g), _f(x), _f(y), _f(z), len(fd_indx), ii, _f(cc), C.byref(po)) != 0:
        raise CgfdError(L.cgfd_b200_last_error().decode())
    return out


def dvh2dvz(prob, device=0):
    """The free-surface conversion matrices of `prob` (a hostsetup.HostProblem) computed by the library on the GPU
    (*_dvh2dvz of the four constitutive laws, include/cgfd3d_b200.h): dict(matVx2Vz, matVy2Vz, matF2Vz, matD) of flat
    [ny*nx*9] arrays, ready for prob.mats. The visco-elastic medium needs prob.coords."""
    from . import hostsetup
    L = load_library()
    n = prob.nx * prob.ny * 9
    out = {k: np.zeros(n, np.float32) for k in ("matVx2Vz", "matVy2Vz", "matF2Vz", "matD")}
    saved, prob.mats = prob.mats, {}
    try:
        c = prob.to_c()
    finally:
        prob.mats = saved
    null = abi.fptr()
    vis = prob.medium_type == abi.MEDIUM_VISCOELASTIC_ISO
    if vis:
        keep = [a if hasattr(a, "data_ptr") else np.ascontiguousarray(a, np.float32) for a in prob.coords]   # torch (device) or numpy
        xyz = tuple(abi.as_f(a) for a in keep)
    else:
        xyz = (null, null, null)
    ii = (C.c_int * len(hostsetup.FDC_INDX))(*hostsetup.FDC_INDX)
    cc = np.asarray(hostsetup.FDC_COEF, np.float32)
    rc = L.cgfd_b200_dvh2dvz(device, C.byref(c), xyz[0], xyz[1], xyz[2], len(hostsetup.FDC_INDX), ii, _f(cc),
                             _f(out["matVx2Vz"]), _f(out["matVy2Vz"]), _f(out["matF2Vz"]), _f(out["matD"]))
    if rc != 0:
        raise CgfdError(L.cgfd_b200_last_error().decode())
    return out


def launch_plan(grid: dict, pml_nlay, free_top: int, rect, blocks_per_sm: int = 2, dz: int = 1):
    """(zchunk, order) of the interior kernel for the tile rectangle rect = (bx0, bx1, by0, by1); pure host logic of the library."""
    L = load_library()
    g = abi.Grid(**grid)
    nl = ((C.c_int * 2) * 3)(*[(C.c_int * 2)(*row) for row in pml_nlay])
    rc = (C.c_int * 4)(*rect)
    zc = C.c_int(0)
    n = L.cgfd_b200_launch_plan(C.byref(g), C.byref(nl), free_top, blocks_per_sm, dz, C.byref(rc), C.byref(zc), None, 0)
    if n < 0:
        raise CgfdError("launch_plan: bad arguments")
    order = (C.c_int * max(n, 1))()
    L.cgfd_b200_launch_plan(C.byref(g), C.byref(nl), free_top, blocks_per_sm, dz, C.byref(rc), C.byref(zc), order, n)
    return zc.value, list(order)[:n]


def halo_plan(grid: dict, dirx: int, diry: int, side: int):
    """(send_box, recv_box) of one side as (i1, ni, j1, nj, k1, nk); pure host logic of the library."""
    L = load_library()
    g = abi.Grid(**grid)
    sb, rb = (C.c_int * 6)(), (C.c_int * 6)()
    if L.cgfd_b200_halo_plan(C.byref(g), dirx, diry, side, C.byref(sb), C.byref(rb)) != 0:
        raise CgfdError(L.cgfd_b200_last_error().decode())
    return tuple(sb), tuple(rb)


def comm_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.cgfd_b200_comm_unique_id(buf) != 0:
        raise CgfdError(L.cgfd_b200_last_error().decode())
    return buf.raw
