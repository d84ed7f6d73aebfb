"""ctypes mirror of include/cgfd3d_b200.h (the C ABI of the hot path).

Field order and types must match the header exactly; tests/test_cpu_basic.py checks sizeof(cgfd_problem_t)
against the value the compiled library reports.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 2

MEDIUM_ELASTIC_ISO = 2
MEDIUM_ELASTIC_VTI = 3
MEDIUM_ELASTIC_ANISO = 4
MEDIUM_VISCOELASTIC_ISO = 5
SRC_SPATIAL_POINT = 1
SRC_SPATIAL_GAUSSIAN = 2
TIMG_ZERO = 0
TIMG_MIRROR = 1

MAX_MEDIA = 24
MAX_MAXWELL = 8
NUM_PAIRS = 8
NUM_STAGES = 4
NUM_METRIC = 10

VX, VY, VZ, TXX, TYY, TZZ, TYZ, TXZ, TXY = range(9)
# metric array order of cgfd_problem_t.metric (forward/gd_t.c:101-180)
M_JAC, M_XIX, M_XIY, M_XIZ, M_ETX, M_ETY, M_ETZ, M_ZTX, M_ZTY, M_ZTZ = range(10)
CMP_NAMES = ["Vx", "Vy", "Vz", "Txx", "Tyy", "Tzz", "Tyz", "Txz", "Txy"]
JAC, XI_X, XI_Y, XI_Z, ET_X, ET_Y, ET_Z, ZT_X, ZT_Y, ZT_Z = range(10)

fptr = C.POINTER(C.c_float)
iptr = C.POINTER(C.c_int32)


class Grid(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nx", "ny", "nz", "ni1", "ni2", "nj1", "nj2", "nk1", "nk2")]


class Fd(C.Structure):
    _fields_ = [
        ("rk_a", C.c_float * NUM_STAGES),
        ("rk_b", C.c_float * NUM_STAGES),
        ("dir", ((C.c_int32 * 3) * NUM_STAGES) * NUM_PAIRS),
        ("indx", (C.c_int32 * 5) * 2),
        ("coef", (C.c_float * 5) * 2),
        ("lay_len", (C.c_int32 * 2) * 3),
        ("lay_indx", ((C.c_int32 * 5) * 2) * 3),
        ("lay_coef", ((C.c_float * 5) * 2) * 3),
    ]


class PmlFace(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("nlay", C.c_int32), ("A", fptr), ("B", fptr), ("D", fptr)]


class Src(C.Structure):
    _fields_ = [
        ("total_number", C.c_int32),
        ("max_nt", C.c_int32),
        ("max_stage", C.c_int32),
        ("si", iptr), ("sj", iptr), ("sk", iptr),
        ("si_inc", fptr), ("sj_inc", fptr), ("sk_inc", fptr),
        ("it_begin", iptr), ("it_end", iptr),
        ("is_surface_force_strict", C.c_int32),
        ("total_number_surface_force", C.c_int32),
        ("force_rate_indx", iptr),
        ("itype_spatial_ext", C.c_int32),
        ("ext_half_npoint", C.c_int32),
        ("ext_func_coef", C.c_float),
        ("force_actived", C.c_int32),
        ("moment_actived", C.c_int32),
        ("Fx", fptr), ("Fy", fptr), ("Fz", fptr),
        ("Mxx", fptr), ("Myy", fptr), ("Mzz", fptr), ("Mxz", fptr), ("Myz", fptr), ("Mxy", fptr),
        ("Fx_rate", fptr), ("Fy_rate", fptr), ("Fz_rate", fptr),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("grid", Grid),
        ("fd", Fd),
        ("dt", C.c_float),
        ("medium_type", C.c_int32),
        ("nmaxwell", C.c_int32),
        ("ncmp", C.c_int32),
        ("metric", fptr * NUM_METRIC),
        ("nmedia", C.c_int32),
        ("media", fptr * MAX_MEDIA),
        ("visco_wl", C.c_float * MAX_MAXWELL),
        ("free_top", C.c_int32),
        ("timg_mode", C.c_int32),
        ("pml", (PmlFace * 2) * 3),
        ("matVx2Vz", fptr), ("matVy2Vz", fptr), ("matF2Vz", fptr), ("matD", fptr),
        ("ablexp_enabled", C.c_int32),
        ("ablexp_blk", (C.c_int32 * 7) * 6),
        ("ablexp_Ex", fptr), ("ablexp_Ey", fptr), ("ablexp_Ez", fptr),
        ("src", Src),
        ("neigh", C.c_int32 * 4),
        ("graves_Qs", fptr),
        ("graves_Qs_freq", C.c_float),
    ]


def as_f(a: np.ndarray | None):
    """float32 C-contiguous ndarray -> POINTER(c_float) (NULL for None). Caller keeps `a` alive."""
    if a is None:
        return fptr()
    if hasattr(a, "data_ptr"):   # torch tensor (host or device memory; the library takes either for metric / media)
        assert str(a.dtype) == "torch.float32" and a.is_contiguous()
        return C.cast(a.data_ptr(), fptr)
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(fptr)


def as_i(a: np.ndarray | None):
    if a is None:
        return iptr()
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(iptr)
