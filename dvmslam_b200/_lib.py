"""ctypes loader for libdvmslam_b200.so (the C-ABI declared in include/dvmslam_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this raises, and
every entry point that needs a GPU returns DVM_ERR_NO_DEVICE (-4) without one.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdvmslam_b200.so")
_LIB = None

DVM_OK, DVM_ERR_INVALID, DVM_ERR_CUDA, DVM_ERR_CAPACITY, DVM_ERR_NO_DEVICE, DVM_ERR_NUMERIC = 0, -1, -2, -3, -4, -5


class DvmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dvmslam_b200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> None:
    """Compile the CUDA extension for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", os.path.join(_HERE, "csrc")],
                       capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libdvmslam_b200.so failed")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.dvm_last_error.restype = C.c_char_p
        L.dvm_version.restype = C.c_char_p
        L.dvm_kernel_launch_count.restype = C.c_uint64
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != DVM_OK:
        raise DvmError(rc, lib().dvm_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(lib().dvm_kernel_launch_count())
