"""ctypes wrapper over oracle/libcgfd_oracle.so (oracle/cgfd_oracle.c): the plain-C restatement of the
isotropic hot path. TEST INFRASTRUCTURE, NOT PRODUCT (see the header of cgfd_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcgfd_oracle.so")
fptr = C.POINTER(C.c_float)
_lib = None


def available() -> bool:
    return os.path.isfile(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.cgfd_oracle_create.restype = C.c_void_p
        L.cgfd_oracle_create.argtypes = [C.c_void_p]
        L.cgfd_oracle_pml_aux_size.restype = C.c_size_t
        L.cgfd_oracle_pml_aux_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cgfd_oracle_set_pml_aux.argtypes = [C.c_void_p, C.c_int, C.c_int, fptr]
        L.cgfd_oracle_get_pml_aux.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, fptr]
        L.cgfd_oracle_onestage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, fptr, fptr]
        L.cgfd_oracle_set_dd.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int, fptr, fptr]
        L.cgfd_oracle_run.argtypes = [C.c_void_p, C.c_int, fptr, C.c_int, C.POINTER(C.c_int64), fptr, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _f(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(fptr)


class PortSolver:
    def __init__(self, prob):
        self.prob = prob
        self._c = prob.to_c()
        self.h = lib().cgfd_oracle_create(C.byref(self._c))
        if not self.h:
            raise RuntimeError("cgfd_oracle_create failed (isotropic elastic only)")
        self.ncmp = 9
        self.shape = (9, prob.nz, prob.ny, prob.nx)

    def pml_aux_size(self, idim, iside):
        return lib().cgfd_oracle_pml_aux_size(self.h, idim, iside)

    def set_pml_aux(self, idim, iside, aux):
        aux = np.ascontiguousarray(aux, np.float32)
        assert aux.size == self.pml_aux_size(idim, iside)
        assert lib().cgfd_oracle_set_pml_aux(self.h, idim, iside, _f(aux)) == 0

    def get_pml_aux(self, idim, iside, level=0):
        out = np.zeros(self.pml_aux_size(idim, iside), np.float32)
        assert lib().cgfd_oracle_get_pml_aux(self.h, idim, iside, level, _f(out)) == 0
        return out

    def get_pml_aux_rhs(self, idim, iside):
        return self.get_pml_aux(idim, iside, 2)

    def set_dd(self, indx, vi, mij):
        """distributed sources: vi [nt][4][n][3] and / or mij [nt][4][n][6] (None = not active)"""
        indx = np.ascontiguousarray(indx, np.int64)
        nt = (vi if vi is not None else mij).shape[0]
        vi = None if vi is None else np.ascontiguousarray(vi, np.float32)
        mij = None if mij is None else np.ascontiguousarray(mij, np.float32)
        null = fptr()
        assert lib().cgfd_oracle_set_dd(self.h, len(indx), indx.ctypes.data_as(C.POINTER(C.c_int64)), nt,
                                        _f(vi) if vi is not None else null, _f(mij) if mij is not None else null) == 0

    def onestage(self, it, ipair, istage, w_cur):
        w_cur = np.ascontiguousarray(w_cur, np.float32)
        rhs = np.zeros(self.shape, np.float32)
        assert lib().cgfd_oracle_onestage(self.h, it, ipair, istage, _f(w_cur), _f(rhs)) == 0
        return rhs

    def run(self, nsteps, w0=None, rec_iptr=None):
        w = np.zeros(self.shape, np.float32) if w0 is None else np.array(w0, np.float32, order="C", copy=True)
        nrec = 0 if rec_iptr is None else len(rec_iptr)
        idx = np.ascontiguousarray(rec_iptr if nrec else [0], np.int64)
        rec = np.zeros((nsteps, 9, max(nrec, 1)), np.float32)
        secs = C.c_double(0.0)
        rc = lib().cgfd_oracle_run(self.h, nsteps, _f(w), nrec, idx.ctypes.data_as(C.POINTER(C.c_int64)), _f(rec), C.byref(secs))
        assert rc == 0
        return w, (rec if nrec else rec[:, :, :0]), secs.value
