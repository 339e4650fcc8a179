"""Builds and loads tests/host_math_harness.cpp (product HD arithmetic compiled for the host)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_math_harness.cpp")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libhostmath.so")
DEPS = [SRC, os.path.join(ROOT, "constriction_b200", "csrc", "coder_math.cuh"),
        os.path.join(ROOT, "constriction_b200", "csrc", "model_math.cuh")]

_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-Wno-unknown-pragmas", "-o", SO, SRC])
    L = C.CDLL(SO)
    u32p, i32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
    L.h_divmod.argtypes = [C.c_uint64, C.c_uint32, u64p, u32p]
    L.h_divmod.restype = None
    L.h_divmod_check.argtypes = [u64p, u32p, C.c_uint64]
    L.h_divmod_check.restype = C.c_uint64
    L.h_ans_encode.argtypes = [i32p, C.c_uint64, u32p, C.c_int32, C.c_uint64, u32p, u64p, C.c_int]
    L.h_f64div_check.argtypes = [u64p, u32p, C.c_uint64]
    L.h_f64div_check.restype = C.c_uint64
    L.h_ans_encode.restype = C.c_uint64
    L.h_ans_decode.argtypes = [u32p, C.c_uint64, i32p, C.c_uint64, u32p, C.c_uint32, C.c_int32]
    L.h_ans_decode.restype = None
    L.h_range_encode.argtypes = [i32p, C.c_uint64, u32p, C.c_int32, u32p]
    L.h_range_encode.restype = C.c_uint64
    L.h_range_encode_split.argtypes = [i32p, C.c_uint64, C.c_uint64, u32p, C.c_int32, u32p]
    L.h_range_encode_split.restype = C.c_uint64
    L.h_range_decode.argtypes = [u32p, C.c_uint64, i32p, C.c_uint64, u32p, C.c_uint32, C.c_int32]
    L.h_range_quantile_check.argtypes = [u64p, u64p, C.c_uint64]
    L.h_range_quantile_check.restype = C.c_uint64
    L.h_erf.argtypes = [C.c_double]
    L.h_erf.restype = C.c_double
    L.h_exp.argtypes = [C.c_double]
    L.h_exp.restype = C.c_double
    L.h_qgauss_cdf.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_double, u32p]
    _LIB = L
    return L
