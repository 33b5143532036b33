"""ctypes binding of libfrtm_b200.so — the only way device work is issued by this package.

The prototypes are parsed from ``include/frtm_b200.h`` so the header stays the single source of truth for
the C ABI.  There is deliberately no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised (the product path must never silently run on PyTorch/CPU).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("FRTM_B200_LIB") or os.path.join(_HERE, "libfrtm_b200.so")   # override: instrumented builds (make timing)
HEADER_PATH = os.path.join(ROOT, "include", "frtm_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
}


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[object, List[object]]]:
    """name -> (restype, argtypes) for every ``frtm_*`` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s*(frtm_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CTYPES[a.replace("const", "").split()[0]])
        protos[name] = (restype, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "libfrtm_b200.so not found at %s — build it with `make -C frtm_vos_b200/csrc` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU/PyTorch fallback." % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)       # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self._status_fns = {n for n, (r, _) in self.protos.items() if r is ctypes.c_int and n not in ("frtm_version",)}

    def __getattr__(self, name):
        full = "frtm_" + name
        fn = getattr(self.cdll, full)
        if full in self._status_fns:
            def checked(*args, _fn=fn, _n=full):
                rc = _fn(*args)
                if rc != 0:
                    raise RuntimeError("%s failed (%d): %s" % (_n, rc, self.cdll.frtm_last_error().decode()))
            setattr(self, name, checked)
            return checked
        setattr(self, name, fn)
        return fn


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """Raw handle of the current CUDA stream (this is on the launch path: ~3400 calls per 65-frame sequence)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: frtm_vos_b200 has no CPU path (got device %s)" % (what, t.device))
