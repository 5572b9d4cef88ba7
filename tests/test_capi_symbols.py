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


def test_fundamental_from_poses_is_the_oracles():
    """dvm_fundamental_from_poses is host arithmetic (no GPU): T12 = T1w * T2w^-1 by Sophus' normalising quaternion
    product, F12 = K1^-T [t12]x R12 K2^-1 by Eigen's 3x3 inverse and left-to-right products, the epipole -- bit-equal to
    the oracle restatement, which tests/test_ref_matchers.py pins to the reference's sources."""
    import numpy as np

    from dvmslam_b200.matching import fundamental_from_poses
    from oracle.bow import fundamental_from_poses as oracle_fundamental

    rng = np.random.default_rng(0)
    for _ in range(200):
        q1, q2 = rng.normal(size=4), rng.normal(size=4)
        q1, q2 = (q1 / np.linalg.norm(q1)).astype(np.float32), (q2 / np.linalg.norm(q2)).astype(np.float32)
        t1, t2 = rng.normal(size=3).astype(np.float32), rng.normal(size=3).astype(np.float32)
        K1 = np.array([500 + rng.uniform(-50, 50), 500 + rng.uniform(-50, 50), 320 + rng.uniform(-5, 5), 240], np.float32)
        K2 = np.array([994.3, 993.4, 638.0, 372.6], np.float32)
        a, b = fundamental_from_poses(q1, t1, q2, t2, K1, K2), oracle_fundamental(q1, t1, q2, t2, K1, K2)
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
