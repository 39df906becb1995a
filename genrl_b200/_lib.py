"""ctypes binding of libgenrl_b200.so, generated from include/genrl_b200.h (the C-ABI is the single source of truth).

The product path fails loudly when the CUDA library is missing or cannot be built: there is no CPU fallback.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "genrl_b200.h")
LIB_PATH = os.path.join(HERE, "lib", "libgenrl_b200.so")

_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
}


def parse_header(path=HEADER):
    """Returns {name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int|long long)\s+(genrl_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "const char*": ctypes.c_char_p}[ret]
        argtypes, argnames = [], []
        if args:
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(a.split("*")[-1].strip())
                else:
                    ty, nm = a.rsplit(" ", 1)
                    ty = ty.replace("const ", "").strip()
                    argtypes.append(_CTYPES[ty])
                    argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


class GenrlError(RuntimeError):
    pass


_NO_LAUNCH = ("genrl_abi_version", "genrl_pack_entry_size", "genrl_device_info", "genrl_rssm_scan_stream_supported")


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()
        self.launches = 0          # number of kernel-launching C-ABI calls made (bench.py reports it)
        self.profile = None        # when a list: (name, start_event, end_event, args) per call (bench.py roofline pass)

    def load(self):
        if self._dll is not None:
            return self._dll
        # always go through build(): a no-op when build.stamp matches the digest of csrc/ + the header, a rebuild when the library
        # is missing or stale (a stale binary with a changed argument list but the same symbol names would otherwise be called
        # with the wrong arguments and never be detected)
        from . import build as _build
        try:
            _build.build()
        except Exception as e:
            if not os.path.exists(LIB_PATH):
                raise GenrlError(f"{LIB_PATH} is missing and could not be built ({e}); the genrl_b200 CUDA path has no fallback")
            raise GenrlError(f"{LIB_PATH} is stale (csrc/ or include/ changed) and could not be rebuilt: {e}")
        if not os.path.exists(LIB_PATH):
            raise GenrlError(f"{LIB_PATH} is missing and could not be built; the genrl_b200 CUDA path has no fallback")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes, _) in self.protos.items():
            fn = getattr(dll, name)          # AttributeError here = header/library mismatch: fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def __getattr__(self, name):
        if name.startswith("genrl_"):
            fn = getattr(self.load(), name)
            restype = self.protos[name][0]
            if restype is not ctypes.c_int or name in _NO_LAUNCH:
                return fn

            def call(*args, _fn=fn, _name=name):
                self.launches += 1
                if self.profile is not None:
                    import torch
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = _fn(*args)
                    e1.record()
                    self.profile.append((_name, e0, e1, args[1:4], args))
                    if rc != 0:
                        raise GenrlError(f"{_name} failed: {self._dll.genrl_last_error().decode()}")
                    return rc
                rc = _fn(*args)
                if rc != 0:
                    raise GenrlError(f"{_name} failed: {self._dll.genrl_last_error().decode()}")
                return rc

            self.__dict__[name] = call
            return call
        raise AttributeError(name)


lib = _Lib()
