"""ctypes binding of libconstriction_b200.so (the C ABI in include/constriction_b200.h).

There is no CPU fallback: if the shared library cannot be built / loaded this module raises, and
every compute entry point returns CTR_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

OK = 0
ERR_IMPOSSIBLE_SYMBOL = 1
ERR_INVALID_DATA = 2
ERR_TRAILING_ZERO = 3
ERR_NOT_SEALED = 4
ERR_BAD_MODEL = 5
ERR_SEEK = 6
ERR_OUT_OF_SPACE = 7
ERR_BAD_ARGUMENT = 8
ERR_CUDA = 9

INDEX_NONE, INDEX_PER_SYMBOL, INDEX_PER_STREAM = 0, 1, 2
FLAG_RAW = 1
FLAG_CHECKPOINTS = 2

u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class Layout(C.Structure):
    """ctr_layout"""
    _fields_ = [
        ("n_streams", C.c_uint64),
        ("n_symbols", C.c_uint64),
        ("sym_offsets_dev", vp),
        ("model_index_dev", vp),
        ("model_index_mode", C.c_int32),
        ("flags", C.c_uint32),
        ("checkpoint_every", C.c_uint32),
        ("reserved", C.c_uint32),
        ("ckpt_offsets_dev", vp),
        ("checkpoints_dev", vp),
    ]


class ContainerView(C.Structure):
    """ctr_container_view"""
    _fields_ = [("coder", C.c_uint32), ("word_bits", C.c_uint32), ("precision", C.c_uint32), ("checkpoint_every", C.c_uint32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32), ("n_streams", C.c_uint64), ("n_symbols", C.c_uint64),
                ("total_words", C.c_uint64), ("n_records", C.c_uint64), ("sym_offsets", vp), ("offsets", vp),
                ("ckpt_offsets", vp), ("records", vp), ("words", vp)]


# every symbol include/constriction_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ctr_abi_version": (C.c_int, []),
    "ctr_status_string": (C.c_char_p, [C.c_int]),
    "ctr_last_cuda_error": (C.c_char_p, []),
    "ctr_device_count": (C.c_int, []),
    "ctr_model_quantized_gaussian": (C.c_int, [C.c_int32, C.c_int32, vp, vp, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_categorical_f32": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_categorical_f64": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_categorical_perfect_f32": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_categorical_perfect_f64": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_quantized": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_binomial": (C.c_int, [vp, vp, C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_from_cdf": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int32, vp, C.POINTER(vp)]),
    "ctr_model_uniform": (C.c_int, [C.c_uint32, vp, C.POINTER(vp)]),
    "ctr_model_destroy": (C.c_int, [vp]),
    "ctr_model_info": (C.c_int, [vp, u32p, u32p, i32p]),
    "ctr_model_cdf_dev": (vp, [vp]),
    "ctr_model_copy_cdf_host": (C.c_int, [vp, vp, vp]),
    "ctr_ans_encode_workspace_bytes": (C.c_size_t, [C.POINTER(Layout)]),
    "ctr_ans_max_compressed_words": (C.c_uint64, [C.POINTER(Layout)]),
    "ctr_ans_encode_reverse": (C.c_int, [vp, vp, C.POINTER(Layout), vp, vp, C.c_size_t, vp, C.c_uint64, vp, vp, vp, vp]),
    "ctr_ans_decode": (C.c_int, [vp, vp, vp, C.POINTER(Layout), vp, vp, vp, vp, vp, vp]),
    "ctr_range_encode_workspace_bytes": (C.c_size_t, [C.POINTER(Layout)]),
    "ctr_range_max_compressed_words": (C.c_uint64, [C.POINTER(Layout)]),
    "ctr_range_encode": (C.c_int, [vp, vp, C.POINTER(Layout), vp, vp, C.c_size_t, vp, C.c_uint64, vp, vp, vp, vp]),
    "ctr_range_decode": (C.c_int, [vp, vp, vp, C.POINTER(Layout), vp, vp, vp, vp, vp, vp]),
    "ctr_checkpoint_max_records": (C.c_uint64, [C.POINTER(Layout)]),
    "ctr_checkpoint_offsets": (C.c_int, [C.POINTER(Layout), vp, vp]),
    "ctr_ans_encode_reverse_gaussian": (C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, C.POINTER(Layout), vp, vp, C.c_size_t,
                                                  vp, C.c_uint64, vp, vp, vp, vp]),
    "ctr_ans_decode_gaussian": (C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, vp, C.POINTER(Layout), vp, vp, vp, vp, vp, vp]),
    "ctr_range_encode_gaussian": (C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, C.POINTER(Layout), vp, vp, C.c_size_t,
                                            vp, C.c_uint64, vp, vp, vp, vp]),
    "ctr_range_decode_gaussian": (C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, vp, C.POINTER(Layout), vp, vp, vp, vp, vp, vp]),
    "ctr_small_model_from_cdf": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int32, vp, vp, C.POINTER(vp)]),
    "ctr_small_model_categorical_f32": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, vp, C.POINTER(vp)]),
    "ctr_small_model_categorical_f64": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, vp, C.POINTER(vp)]),
    "ctr_small_model_destroy": (C.c_int, [vp]),
    "ctr_small_model_copy_cdf_host": (C.c_int, [vp, vp, vp]),
    "ctr_small_encode_workspace_bytes": (C.c_size_t, [C.POINTER(Layout)]),
    "ctr_small_max_compressed_words": (C.c_uint64, [C.POINTER(Layout)]),
    "ctr_small_ans_encode_reverse": (C.c_int, [vp, vp, C.POINTER(Layout), vp, C.c_size_t, vp, C.c_uint64, vp, vp, vp]),
    "ctr_small_ans_decode": (C.c_int, [vp, vp, vp, C.POINTER(Layout), vp, vp, vp]),
    "ctr_small_range_encode": (C.c_int, [vp, vp, C.POINTER(Layout), vp, C.c_size_t, vp, C.c_uint64, vp, vp, vp]),
    "ctr_small_range_decode": (C.c_int, [vp, vp, vp, C.POINTER(Layout), vp, vp, vp]),
    "ctr_ans_encode_reverse_host": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.c_uint64, vp,
                                              C.POINTER(C.c_int), u64p]),
    "ctr_ans_decode_host": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.POINTER(C.c_int), u64p]),
    "ctr_range_encode_host": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.c_uint64, vp,
                                        C.POINTER(C.c_int), u64p]),
    "ctr_range_decode_host": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.POINTER(C.c_int), u64p]),
    "ctr_ans_encode_reverse_host_async": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.c_uint64, vp,
                                                    C.POINTER(C.c_int), u64p, C.POINTER(vp)]),
    "ctr_ans_decode_host_async": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.POINTER(C.c_int), u64p,
                                            C.POINTER(vp)]),
    "ctr_range_encode_host_async": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.c_uint64, vp,
                                              C.POINTER(C.c_int), u64p, C.POINTER(vp)]),
    "ctr_range_decode_host_async": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_int32, vp, C.POINTER(C.c_int), u64p,
                                              C.POINTER(vp)]),
    "ctr_host_job_wait": (C.c_int, [vp]),
    "ctr_stream_write_value32": (C.c_int, [vp, C.c_uint32, vp]),
    "ctr_stream_wait_value32": (C.c_int, [vp, C.c_uint32, vp]),
    "ctr_container_size": (C.c_size_t, [C.POINTER(ContainerView)]),
    "ctr_container_pack": (C.c_int, [C.POINTER(ContainerView), vp, C.c_size_t]),
    "ctr_container_unpack": (C.c_int, [vp, C.c_size_t, C.POINTER(ContainerView)]),
    "ctr_gather_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "ctr_gather_destroy": (C.c_int, [vp]),
    "ctr_gather_begin_turn": (C.c_int, [vp, u32p, C.POINTER(vp), u64p, C.POINTER(vp)]),
    "ctr_gather_slot": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(vp), C.POINTER(vp)]),
    "ctr_gather_push": (C.c_int, [vp, C.c_uint32, C.c_uint64, vp]),
    "ctr_gather_wait": (C.c_int, [vp, C.c_uint32, vp]),
    "ctr_gather_release": (C.c_int, [vp, C.c_uint32, vp]),
    "ctr_gather_sync": (C.c_int, [vp]),
    "ctr_gather_compressed_nccl": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp,
                                             vp, vp, vp]),
    "ctr_kernel_launch_count": (C.c_uint64, []),
    "ctr_profile_enable": (None, [C.c_int]),
    "ctr_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), u64p]),
}

_LIB = None


def library_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Loads (building first if stale and nvcc is available) the CUDA library.  Raises on failure."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this machine: use the prebuilt library if there is one
            if not os.path.exists(path):
                raise RuntimeError(
                    "libconstriction_b200.so is missing and could not be built; constriction_b200 has no "
                    f"CPU fallback ({exc})") from exc
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.ctr_abi_version() != 3:
        raise RuntimeError("libconstriction_b200.so: ABI version mismatch")
    _LIB = lib
    return lib


class CtrError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = load().ctr_status_string(code).decode()
        if code == ERR_CUDA:
            detail = (detail + " " + load().ctr_last_cuda_error().decode()).strip()
        super().__init__(f"{msg} {detail}".strip())


def raise_for(code: int, detail: str = "") -> None:
    """Maps status codes to the exceptions the reference's Python API raises
    (pybindings/stream/mod.rs:83-91, stack.rs:230-235, queue.rs:677-685)."""
    if code == OK:
        return
    msg = load().ctr_status_string(code).decode()
    if detail:
        msg = f"{msg} ({detail})"
    if code == ERR_IMPOSSIBLE_SYMBOL:
        raise KeyError(msg)
    if code in (ERR_INVALID_DATA, ERR_NOT_SEALED):
        raise AssertionError(msg)
    if code in (ERR_TRAILING_ZERO, ERR_BAD_MODEL, ERR_SEEK, ERR_BAD_ARGUMENT):
        raise ValueError(msg)
    if code == ERR_OUT_OF_SPACE:
        raise MemoryError(msg)
    raise CtrError(code, detail)
