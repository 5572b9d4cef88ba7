"""The C-ABI library loads and exports every symbol include/dvmslam_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dvmslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"DVM_API\s+[\w\s\*]+?\b(dvm_\w+)\s*\(", src)))


def test_header_is_plain_c():
    import subprocess

    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "dvmslam_b200.h")],
                   check=True)


def test_library_exports_every_declared_symbol():
    from dvmslam_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.dvm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.dvm_version()


def test_no_gpu_fails_loudly():
    """Without a GPU the create call must return DVM_ERR_NO_DEVICE, never silently fall back."""
    import torch

    from dvmslam_b200 import _lib

    if torch.cuda.is_available():
        return
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.dvm_orb_create(ctypes.byref(h), 0, 1000, ctypes.c_float(1.2), 8, 20, 7, 640, 480)
    assert rc == _lib.DVM_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.dvm_last_error() or b"sm_100a" in L.dvm_last_error()
