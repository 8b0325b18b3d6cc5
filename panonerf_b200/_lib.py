"""ctypes binding of the C ABI declared in include/panonerf_b200.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the declaration, and
`declared_symbols()` lets the tests check that the shared library exports every one of them.  There is no CPU
fallback: if the library is missing the import of any op fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "panonerf_b200.h")
LIB_PATH = os.environ.get("PNB_LIB_PATH") or os.path.join(HERE, "libpanonerf_b200.so")   # (override: A/B experiments)

PNB_F32, PNB_BF16 = 0, 1
EPI_BIAS, EPI_RELU, EPI_MASK, EPI_ACCUM = 1, 2, 4, 8

_CTYPE = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float}


def _parse_header() -> Dict[str, Tuple[object, List[object]]]:
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(pnb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if ret == "const char*" else _CTYPE[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPE[base])
        protos[name] = (restype, argtypes)
    return protos


PROTOTYPES = _parse_header()


def declared_symbols() -> List[str]:
    return sorted(PROTOTYPES)


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded shared library (loads on first use; raises LibraryMissing if it was never built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m panonerf_b200.build` "
                "(there is no CPU fallback for the CUDA hot path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            if os.environ.get("PNB_LIB_PATH") and not hasattr(handle, name):
                continue                        # A/B against an older build of the library (experiments only)
            fn = getattr(handle, name)          # AttributeError here == header/library mismatch: fail loudly
            fn.restype, fn.argtypes = restype, argtypes
        if handle.pnb_abi_version() != 1:
            raise RuntimeError("panonerf_b200: ABI version mismatch between header and library")
        _lib = handle
    return _lib


class KernelError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().pnb_last_error().decode("utf-8", "replace")
        raise KernelError(f"panonerf_b200 {what} failed (code {rc}): {msg}")
