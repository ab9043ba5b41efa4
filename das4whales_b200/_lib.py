"""cffi (ABI-mode) binding of libd4w.so.  The cdef is the block between D4W_CDEF_BEGIN and
D4W_CDEF_END of include/d4w.h, read verbatim, so header and binding cannot drift apart.

There is no CPU fallback: if the library is missing or fails to load, every op raises.
"""
import os
import re
import threading

import cffi

HERE = os.path.dirname(os.path.abspath(__file__))
# canonical header: <repo>/include/d4w.h; das4whales_b200/d4w.h is the copy that ships inside the package (refreshed by
# _build.build_library, checked identical by tests/test_abi.py) so that an installed / copied package still imports
_ROOT_HEADER = os.path.normpath(os.path.join(HERE, "..", "include", "d4w.h"))
HEADER = _ROOT_HEADER if os.path.exists(_ROOT_HEADER) else os.path.join(HERE, "d4w.h")
LIBPATH = os.path.join(HERE, "libd4w.so")

ffi = cffi.FFI()
_lock = threading.Lock()
_lib = None


def header_cdef():
    with open(HEADER) as f:
        text = f.read()
    m = re.search(r"/\* D4W_CDEF_BEGIN \*/(.*?)/\* D4W_CDEF_END \*/", text, re.S)
    if not m:
        raise RuntimeError("include/d4w.h: D4W_CDEF markers not found")
    return m.group(1)


def declared_symbols():
    """Names of every function include/d4w.h declares (used by the CPU symbol-export test)."""
    return re.findall(r"\b(d4w_[a-z0-9_]+)\s*\(", header_cdef())


ffi.cdef(header_cdef())


class D4WError(RuntimeError):
    pass


def lib():
    """dlopen libd4w.so (once). Fails loudly -- the CUDA extension IS the product path."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIBPATH):
                    raise D4WError(
                        f"{LIBPATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the das4whales_b200 kernels)")
                _lib = ffi.dlopen(LIBPATH)
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = ffi.string(lib().d4w_last_error()).decode(errors="replace")
        if rc == 2:
            raise ValueError(f"{what}: {msg} (das4whales_b200.dsp.supported_shape(nx, ns) suggests the nearest supported shape)")
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        raise D4WError(f"{what}: {msg} (status {rc})")


def ptr(t, ctype="void*"):
    """device pointer of a torch tensor as a cffi pointer"""
    return ffi.cast(ctype, t.data_ptr())


def stream_ptr():
    import torch
    return ffi.cast("void*", torch.cuda.current_stream().cuda_stream)
