"""ctypes wrapper over oracle/_ref/libcgfd_ref_flat.so (oracle/ref_flat_adapter.c): the UNMODIFIED
reference functions of the hot path, callable on the flat problem description of the C ABI.

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libcgfd_ref_flat.so")
# the same adapter over the MIRROR restatement of the traction image (oracle/patch_timg.py): oracle of timg_mode = CGFD_TIMG_MIRROR
LIB_MIRROR = os.path.join(HERE, "_ref", "libcgfd_ref_flat_mirror.so")
fptr = C.POINTER(C.c_float)


def available(mirror: bool = False) -> bool:
    return os.path.isfile(LIB_MIRROR if mirror else LIB)


_libs = {}


def lib(mirror: bool = False):
    if mirror not in _libs:
        L = C.CDLL(LIB_MIRROR if mirror else LIB)
        L.cgfd_ref_create.restype = C.c_void_p
        L.cgfd_ref_create.argtypes = [C.c_void_p]
        L.cgfd_ref_ncmp.argtypes = [C.c_void_p]
        L.cgfd_ref_set_coords.argtypes = [C.c_void_p, fptr, fptr, fptr]
        L.cgfd_ref_pml_aux_size.restype = C.c_size_t
        L.cgfd_ref_pml_aux_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cgfd_ref_set_pml_aux.argtypes = [C.c_void_p, C.c_int, C.c_int, fptr]
        L.cgfd_ref_get_pml_aux.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, fptr]
        L.cgfd_ref_dvh2dvz.argtypes = [C.c_void_p, fptr, fptr, fptr, fptr]
        L.cgfd_ref_onestage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, fptr, fptr]
        L.cgfd_ref_metric_from_coords.argtypes = [C.c_void_p, fptr, fptr, fptr, fptr]
        L.cgfd_ref_set_dd.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int, fptr, fptr, C.c_char_p]
        L.cgfd_ref_run.argtypes = [C.c_void_p, C.c_int, fptr, C.c_int, C.POINTER(C.c_int64), fptr, C.c_char_p,
                                   C.POINTER(C.c_double)]
        _libs[mirror] = L
    return _libs[mirror]


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(fptr)


class _RefSolverImpl:
    """The reference CPU implementation behind the same calls as cgfd3d_b200.solver.Solver (in-process)."""

    def __init__(self, prob):
        self.prob = prob
        self._c = prob.to_c()
        self.L = lib(mirror=bool(getattr(prob, "timg_mode", 0)))
        self.h = self.L.cgfd_ref_create(C.byref(self._c))
        if not self.h:
            raise RuntimeError("cgfd_ref_create failed")
        self.ncmp = self.L.cgfd_ref_ncmp(self.h)
        if getattr(prob, "coords", None) is not None:
            x, y, z = (np.ascontiguousarray(a, np.float32) for a in prob.coords)
            assert self.L.cgfd_ref_set_coords(self.h, _f(x), _f(y), _f(z)) == 0
        self.shape = (self.ncmp, prob.nz, prob.ny, prob.nx)

    def pml_aux_size(self, idim, iside):
        return self.L.cgfd_ref_pml_aux_size(self.h, idim, iside)

    def set_pml_aux(self, idim, iside, aux):
        """aux: the 9 components in use; the reference allocates ncmp per level (forward/bdry_t.c:300), the rest stays 0"""
        aux = np.ascontiguousarray(aux, np.float32).ravel()
        full = np.zeros(self.pml_aux_size(idim, iside), np.float32)
        assert aux.size == full.size // self.ncmp * 9
        full[:aux.size] = aux
        assert self.L.cgfd_ref_set_pml_aux(self.h, idim, iside, _f(full)) == 0

    def get_pml_aux(self, idim, iside, level=0):
        out = np.zeros(self.pml_aux_size(idim, iside), np.float32)
        assert self.L.cgfd_ref_get_pml_aux(self.h, idim, iside, level, _f(out)) == 0
        return out[:out.size // self.ncmp * 9].copy()

    def get_pml_aux_rhs(self, idim, iside):
        return self.get_pml_aux(idim, iside, 2)

    def metric_from_coords(self, x, y, z):
        """the reference's gd_curv_metric_cal on this grid: array [10][nz][ny][nx]"""
        x, y, z = (np.ascontiguousarray(a, np.float32) for a in (x, y, z))
        out = np.zeros((10,) + x.shape, np.float32)
        assert self.L.cgfd_ref_metric_from_coords(self.h, _f(x), _f(y), _f(z), _f(out)) == 0
        return out

    def set_dd(self, indx, vi, mij, nt_per_read):
        """distributed sources: vi [nt][stage][n][3] and / or mij [nt][stage][n][6] (None = not active)"""
        indx = np.ascontiguousarray(indx, np.int64)
        nt = (vi if vi is not None else mij).shape[0]
        vi = None if vi is None else np.ascontiguousarray(vi, np.float32)
        mij = None if mij is None else np.ascontiguousarray(mij, np.float32)
        self._ddtmp = tempfile.TemporaryDirectory()
        null = fptr()
        rc = self.L.cgfd_ref_set_dd(self.h, len(indx), indx.ctypes.data_as(C.POINTER(C.c_int64)), int(vi is not None), int(mij is not None),
                                   nt, nt_per_read, _f(vi) if vi is not None else null, _f(mij) if mij is not None else null,
                                   os.path.join(self._ddtmp.name, "dd").encode())
        assert rc == 0

    def dvh2dvz(self):
        n = self.prob.nx * self.prob.ny * 9
        outs = [np.zeros(n, np.float32) for _ in range(4)]
        assert self.L.cgfd_ref_dvh2dvz(self.h, *[_f(o) for o in outs]) == 0
        return dict(matVx2Vz=outs[0], matVy2Vz=outs[1], matF2Vz=outs[2], matD=outs[3])

    def onestage(self, it, ipair, istage, w_cur):
        w_cur = np.ascontiguousarray(w_cur, np.float32)
        rhs = np.zeros(self.shape, np.float32)
        assert self.L.cgfd_ref_onestage(self.h, it, ipair, istage, _f(w_cur), _f(rhs)) == 0
        return rhs

    def run(self, nsteps, w0=None, rec_iptr=None, outdir=None):
        """Returns (w_final, record[it][icmp][ip], seconds inside drv_rk_curv_col_allstep)."""
        w = np.zeros(self.shape, np.float32) if w0 is None else np.array(w0, np.float32, order="C", copy=True)
        nrec = 0 if rec_iptr is None else len(rec_iptr)
        idx = np.ascontiguousarray(rec_iptr if nrec else [0], np.int64)
        rec = np.zeros((nsteps, self.ncmp, max(nrec, 1)), np.float32)
        secs = C.c_double(0.0)
        tmp = None
        if outdir is None:
            tmp = tempfile.TemporaryDirectory()
            outdir = tmp.name
        rc = self.L.cgfd_ref_run(self.h, nsteps, _f(w), nrec, idx.ctypes.data_as(C.POINTER(C.c_int64)), _f(rec),
                                outdir.encode(), C.byref(secs))
        assert rc == 0
        if tmp is not None:
            tmp.cleanup()
        return w, (rec if nrec else rec[:, :, :0]), secs.value


def _serve(conn, prob):
    """child process: one reference instance, commands over a pipe"""
    try:
        R = _RefSolverImpl(prob)
        conn.send(("ok", (R.ncmp, R.shape)))
    except Exception as e:   # pragma: no cover
        conn.send(("err", repr(e)))
        return
    while True:
        try:
            name, args, kw = conn.recv()
        except EOFError:
            return
        if name == "__close__":
            return
        try:
            conn.send(("ok", getattr(R, name)(*args, **kw)))
        except Exception as e:
            conn.send(("err", repr(e)))


class RefSolver:
    """Proxy of _RefSolverImpl living in a forked child process, one per instance.

    The reference keeps process-wide state and is known to touch memory it does not own (SURVEY.md 8c hazards); run inside
    the test process, an earlier instance can change what a later one -- or the arrays of the problem under test -- sees
    (observed: a VTI case wrong only when an isotropic case ran before it in the same pytest process). A child forked per
    instance starts from the pristine library image and a copy-on-write view of `prob`, so nothing leaks either way."""

    def __init__(self, prob):
        import multiprocessing as mp
        lib(mirror=bool(getattr(prob, "timg_mode", 0)))   # dlopen before the fork, the child inherits the mapping
        ctx = mp.get_context("fork")
        self._conn, child = ctx.Pipe()
        self._p = ctx.Process(target=_serve, args=(child, prob), daemon=True)
        import warnings
        with warnings.catch_warnings():
            # the child runs only the C reference and numpy; the parent's CUDA / BLAS threads are never touched there
            warnings.simplefilter("ignore", DeprecationWarning)
            self._p.start()
        child.close()
        st, val = self._conn.recv()
        if st != "ok":
            raise RuntimeError("reference instance failed: %s" % val)
        self.prob = prob
        self.ncmp, self.shape = val

    def _call(self, name, *args, **kw):
        self._conn.send((name, args, kw))
        st, val = self._conn.recv()
        if st != "ok":
            raise RuntimeError("reference call %s failed: %s" % (name, val))
        return val

    def pml_aux_size(self, idim, iside):
        return self._call("pml_aux_size", idim, iside)

    def set_pml_aux(self, idim, iside, aux):
        return self._call("set_pml_aux", idim, iside, np.ascontiguousarray(aux, np.float32))

    def get_pml_aux(self, idim, iside, level=0):
        return self._call("get_pml_aux", idim, iside, level)

    def get_pml_aux_rhs(self, idim, iside):
        return self._call("get_pml_aux_rhs", idim, iside)

    def set_dd(self, indx, vi, mij, nt_per_read):
        return self._call("set_dd", indx, vi, mij, nt_per_read)

    def metric_from_coords(self, x, y, z):
        return self._call("metric_from_coords", x, y, z)

    def dvh2dvz(self):
        return self._call("dvh2dvz")

    def onestage(self, it, ipair, istage, w_cur):
        return self._call("onestage", it, ipair, istage, np.ascontiguousarray(w_cur, np.float32))

    def run(self, nsteps, w0=None, rec_iptr=None, outdir=None):
        return self._call("run", nsteps, w0=w0, rec_iptr=rec_iptr, outdir=outdir)

    def close(self):
        if getattr(self, "_p", None) is not None:
            try:
                self._conn.send(("__close__", (), {}))
            except Exception:
                pass
            self._p.join(timeout=10)
            if self._p.is_alive():
                self._p.kill()
            self._p = None

    def __del__(self):
        self.close()
