"""ctypes binding of libwaveglow_b200.so (the C ABI declared in include/waveglow_b200.h).

Prototypes are parsed from the header itself so the binding cannot drift from the declaration.
There is NO fallback: if the library is missing and cannot be built, or a call fails, a
RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading
from typing import Dict, List, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "waveglow_b200.h")
LIB_PATH = os.path.join(HERE, "libwaveglow_b200.so")

_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "long long": ctypes.c_longlong,
}
_lock = threading.Lock()
_lib = None
_protos: Dict[str, Tuple[object, List[object]]] = {}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """{name: (return type, [argument types])} for every WGB_API declaration."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"WGB_API\s+([\w\s\*]+?)\s*(wgb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    types.append("ptr")
                else:
                    types.append(" ".join(a.split(" ")[:-1]))      # drop the parameter name
        out[name] = (ret, types)
    return out


def _load():
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        # a library built from other sources than this tree (whose header the prototypes below are parsed from) is
        # rebuilt before it is loaded; without nvcc that raises instead of calling stale code through new prototypes
        from . import build as _build
        if _build.needs_build():
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (ret, types) in parse_header().items():
            fn = getattr(lib, name)            # AttributeError -> the .so is stale: fail loudly
            fn.restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
            fn.argtypes = [ctypes.c_void_p if t == "ptr" else _CTYPES[t] for t in types]
            _protos[name] = (fn, types)
        if lib.wgb_abi_version() != 1 or lib.wgb_source_hash().decode() != _build.source_hash():
            raise RuntimeError("libwaveglow_b200.so does not match this source tree; rebuild with "
                               "python -m text2speech_b200.build --force")
        _lib = lib
        return lib


def lib():
    return _load()


def last_error() -> str:
    return _load().wgb_last_error().decode()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke a C-ABI entry point; tensors become device pointers; non-zero status raises."""
    _load()
    fn, types = _protos[name]
    if len(args) != len(types):
        raise TypeError(f"{name} takes {len(types)} arguments, got {len(args)}")
    conv = []
    for a, t in zip(args, types):
        if t == "ptr":
            if a is None:
                conv.append(None)
            elif isinstance(a, torch.Tensor):
                if not a.is_cuda:
                    raise RuntimeError(f"{name}: tensor argument must live on a CUDA device (no CPU path exists)")
                if not a.is_contiguous():
                    raise RuntimeError(f"{name}: tensor argument must be contiguous")
                conv.append(a.data_ptr())
            elif isinstance(a, (str, bytes)):                     # const char* (wgb_set_tuning's key)
                conv.append(ctypes.c_char_p(a.encode() if isinstance(a, str) else a))
            else:
                conv.append(int(a))
        else:
            conv.append(a)
    status = fn(*conv)
    if status != 0:
        raise RuntimeError(f"{name} failed (status {status}): {last_error()}")


def require_b200(device: torch.device) -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("text2speech_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    status = _load().wgb_device_check(idx)
    if status != 0:
        raise RuntimeError(f"wgb_device_check failed: {last_error()}")
