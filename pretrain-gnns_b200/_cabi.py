"""ctypes binding of libpgnn_b200.so, generated from include/pgnn_b200.h at import time.

The prototypes are parsed from the header itself, so the Python side cannot drift from the C ABI:
every `PGNN_API` declaration becomes a typed ctypes function (all pointer types map to `c_void_p`,
i.e. raw device addresses from `tensor.data_ptr()`).  There is NO fallback: if the library has not
been built, or a call returns a negative code, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "pgnn_b200.h")
LIB_PATH = os.environ.get("PGNN_LIB") or os.path.join(_HERE, "libpgnn_b200.so")  # PGNN_LIB: development builds (tools/)

_SCALARS = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "float": ctypes.c_float, "double": ctypes.c_double,
            "void": None}


def _ctype(decl: str):
    decl = re.sub(r"/\*.*?\*/", "", decl).strip()
    if "*" in decl:
        return ctypes.c_char_p if re.match(r"const\s+char\s*\*", decl) else ctypes.c_void_p
    base = decl.replace("const", "").split()
    return _SCALARS[base[0]]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every PGNN_API declaration."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"PGNN_API\s+([\w\s\*]+?)\s*\b(pgnn_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        protos[name] = (_ctype(ret + (" " if "*" not in ret else "")), argtypes)
    return protos


class PgnnError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise PgnnError(
                f"{LIB_PATH} is missing: build it with `python pretrain-gnns_b200/build.py` "
                "(or __graft_entry__.build()). There is no CPU / eager fallback for this path.")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(dll, name)  # AttributeError here = header/library drift: fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def __getattr__(self, name):
        return getattr(self.load(), name)


lib = _Lib()


def check(rc: int, what: str = ""):
    if rc is not None and rc < 0:
        dll = lib.load()
        msg = dll.pgnn_error_string(rc).decode()
        extra = f" (cudaError {dll.pgnn_last_cuda_error()})" if rc == -2 else ""
        raise PgnnError(f"libpgnn_b200 {what} failed: {msg}{extra}")
    return rc
