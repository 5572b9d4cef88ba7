"""The reference's own ORBextractor.cc, compiled where it lies against oracle/cvshim (oracle/_ref,
`make -C oracle ref`), must agree bit for bit with the oracle restatement: this pins the in-tree logic
(octree with libstdc++ sort/list, IC_Angle, rotated BRIEF with libm cosf/sinf, output ordering,
constructor tables) to the actual reference source.  Skipped where neither the built library nor
/root/reference is available."""
import ctypes as C
import os

import numpy as np
import pytest

from dvmslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_orb.so")


def _ref():
    if not os.path.exists(REF_SO):
        if not os.path.isdir("/root/reference/src/slam_system/orb_slam3"):
            pytest.skip("oracle/_ref not built and /root/reference absent")
        import oracle

        oracle.build(ref=True)
    L = C.CDLL(REF_SO)
    L.ref_orb_create.restype = C.c_void_p
    L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ref_orb_destroy.argtypes = [C.c_void_p]
    L.ref_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.ref_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    return L


@pytest.mark.parametrize("w,h,nf,seed,lap", [(640, 480, 1000, 0, (0, 1000)), (1280, 720, 2000, 1, (0, 1000)),
                                             (752, 480, 1200, 2, (0, 0)), (1241, 376, 2000, 3, (300, 700)),
                                             (480, 640, 500, 4, (0, 1000)), (320, 240, 5000, 5, (0, 1000))])
def test_reference_source_equals_oracle(w, h, nf, seed, lap):
    from oracle.orb import KP_DTYPE, OrbOracle

    L = _ref()
    img = synth.frame(w, h, seed)
    cap = nf * 2 + 64
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    mono = C.c_int()
    hd = L.ref_orb_create(nf, 1.2, 8, 20, 7)
    n = L.ref_orb_extract(hd, img.ctypes.data, w, h, w, lap[0], lap[1], kps.ctypes.data, desc.ctypes.data, cap,
                          C.byref(mono))
    sc, inv, s2, is2 = (np.zeros(8, np.float32) for _ in range(4))
    L.ref_orb_tables(hd, sc.ctypes.data, inv.ctypes.data, s2.ctypes.data, is2.ctypes.data)
    L.ref_orb_destroy(hd)
    assert n > 0
    orc = OrbOracle(nf)
    k0, d0, m0 = orc.extract(img, lap)
    assert n == len(k0) and mono.value == m0
    assert np.array_equal(kps[:n], k0)
    assert np.array_equal(desc[:n], d0)
    T = orc.tables()
    for a, b in ((sc, "scale"), (inv, "inv_scale"), (s2, "sigma2"), (is2, "inv_sigma2")):
        assert np.array_equal(a, T[b])
