"""ctypes binding of include/b2o.h (libb2o.so).  There is NO CPU fallback: if the CUDA library is
missing or no B200 is present, every operator constructor raises.

The prototypes below are generated from the declarations in include/b2o.h so that the header stays
the single source of truth (tests/test_abi.py checks every declared symbol is exported)."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2o.so")
HEADER = os.path.join(HERE, "..", "include", "b2o.h")

B2O_OK, B2O_ESHAPE, B2O_EARG, B2O_ECUDA, B2O_ENCCL, B2O_ESTATE, B2O_ENOMEM, B2O_EUNSUPPORTED = range(8)
B2O_F64, B2O_F32, B2O_BF16 = 0, 1, 2


class LinearOperatorException(Exception):
    """Mirror of LinearOperatorException (src/abstract.jl:17-19)."""


class B2OError(RuntimeError):
    """CUDA / NCCL / argument failure reported by libb2o."""


class ErrorException(Exception):
    """Mirror of Julia's ErrorException (wrong push! variant, src/lbfgs.jl:296-298,332-334)."""


_CTYPES = {
    "int": ctypes.c_int, "double": ctypes.c_double, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "size_t": ctypes.c_size_t,
}


def declared_functions(header=HEADER):
    """[(name, restype, [argtypes])] parsed from b2o.h."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = []
    for m in re.finditer(r"\b(int|const char \*)\s*(b2o_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace("const ", "").split()[0]
                    argtypes.append(_CTYPES[base])
        out.append((name, ctypes.c_char_p if "char" in ret else ctypes.c_int, argtypes))
    return out


_lib = None


def load():
    """Load libb2o.so (once) and attach prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libb2o.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` -- "
            "there is no CPU fallback for the operator-apply path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, restype, argtypes in declared_functions():
        fn = getattr(lib, name)  # AttributeError if the header declares something the library lacks
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().b2o_last_error().decode("utf-8", "replace")


def check(status):
    """Map a b2o_status to the exception type the reference would throw."""
    if status == B2O_OK:
        return
    msg = last_error()
    if status == B2O_ESHAPE:
        raise LinearOperatorException(msg)
    if status == B2O_ESTATE:
        if msg.startswith("only the diagonal"):
            raise LinearOperatorException(msg)
        if msg.startswith("Cannot"):
            raise ErrorException(msg)
        raise ErrorException(msg)
    if status == B2O_EARG and msg.startswith("indices should be between"):
        raise LinearOperatorException(msg)
    raise B2OError("libb2o status %d: %s" % (status, msg))
