"""ctypes binding of libpplp_b200.so, generated from the declarations in include/pplp_b200.h.

Every prototype in the header becomes a ctypes function with matching argument types, so the header stays the single
source of truth for the boundary.  Pointer parameters are `c_void_p`: pass `tensor.data_ptr()`, `ndarray.ctypes.data`
or None.  `check(rc)` turns a negative return code into PplpError carrying pplp_last_error().
"""
import ctypes as C
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
HEADER = os.path.join(ROOT, "include", "pplp_b200.h")
LIB_PATH = os.environ.get("PPLP_B200_LIB") or os.path.join(PKG, "libpplp_b200.so")   # the override is for A/B builds of experiments

_SCALARS = {
    "int": C.c_int, "size_t": C.c_size_t, "uint64_t": C.c_uint64, "uint32_t": C.c_uint32, "double": C.c_double,
    "void": None,
}


class PplpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pplp error {code}: {msg}")
        self.code = code
        self.msg = msg


def _ctype(decl):
    decl = decl.strip()
    if "*" in decl or "[" in decl:
        if decl.replace("const", "").replace(" ", "").startswith("char*"):
            return C.c_char_p
        return C.c_void_p
    base = decl.replace("const", "").split()
    return _SCALARS[base[0]]


def parse_header(path=HEADER):
    """Returns {name: (restype, [argtypes])} for every function prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    text = text.replace('extern "C" {', " ")
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pplp_\w+)\s*\(([^;{}]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier) unless the declaration is just a type
                mm = re.match(r"(.*?)(\b\w+)(\s*\[\d*\])?$", a)
                typ = (mm.group(1) + (mm.group(3) or "")) if mm and mm.group(1).strip() else a
                argtypes.append(_ctype(typ))
        protos[name] = (_ctype(ret), argtypes)
    return protos


def declared_symbols():
    return sorted(parse_header().keys())


_lib = None


def lib():
    """Loads the CUDA library.  Raises if it is missing — the package has no other execution path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"pplp_b200: {LIB_PATH} is missing. Build it with `python -m pplp_b200.build` (needs nvcc); "
            "there is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in parse_header().items():
        f = getattr(L, name)   # AttributeError here means the header declares a symbol the library lacks
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc is not None and rc < 0:
        raise PplpError(rc, lib().pplp_last_error().decode(errors="replace"))
    return rc
