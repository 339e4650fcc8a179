"""constriction_b200 -- B200-native batched entropy coding behind constriction's stream API.

`constriction_b200.batch`   many independent coders per call, device tensors in / out (the hot path)
`constriction_b200.stream`  mirror of `constriction.stream.{stack,queue,model}` (one coder per object),
                            every symbol of which is coded by the same CUDA kernels through the C ABI
`constriction_b200.dist`    shard a batch over the GPUs of one node and gather the compressed words

All compute runs in libconstriction_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/constriction_b200.h).  There is no CPU implementation in this package.
"""
from . import _native  # noqa: F401

__all__ = ["batch", "stream", "dist", "native_library_path"]


def native_library_path() -> str:
    return _native.library_path()


def __getattr__(name):
    if name in ("batch", "stream", "dist"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
