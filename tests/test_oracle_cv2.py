"""Pins the oracle: its OpenCV-primitive models and its whole ORB pipeline must equal the cv2 4.13
primitives the reference calls (the reference itself holds no fixtures)."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from dvmslam_b200 import synth  # noqa: E402
from oracle import lib  # noqa: E402
from oracle.orb import OrbOracle, extract_with_cv2  # noqa: E402

u8p = C.POINTER(C.c_uint8)


def P(a):
    return a.ctypes.data_as(u8p)


def setup_module(_):
    cv2.setNumThreads(1)


@pytest.mark.parametrize("w,h", [(640, 480), (1280, 720), (533, 400), (357, 201), (752, 480), (1241, 376), (161, 97)])
def test_resize_model(w, h):
    L = lib()
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    inv = np.float32(1) / np.float32(1.2)
    dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
    for (tw, th) in [(dw, dh), (w // 2, h // 2), (w + 5, h + 3)]:
        ref = cv2.resize(img, (tw, th), interpolation=cv2.INTER_LINEAR)
        out = np.zeros((th, tw), np.uint8)
        L.cvm_resize_linear_u8(P(img), w, h, w, P(out), tw, th, tw)
        assert np.array_equal(ref, out)


@pytest.mark.parametrize("w,h", [(131, 97), (640, 480), (7, 7), (5, 9), (20, 4)])
def test_gaussian_model(w, h):
    L = lib()
    rng = np.random.default_rng(w + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    out = np.zeros_like(img)
    L.cvm_gaussian7_u8(P(img), w, h, w, P(out), w)
    assert np.array_equal(ref, out)


def test_fast_atan2_model():
    L = lib()
    L.cvm_fast_atan2.restype = C.c_float
    L.cvm_fast_atan2.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(3)
    ys = rng.integers(-3000000, 3000000, 20000)
    xs = rng.integers(-3000000, 3000000, 20000)
    ys[:100] = 0
    xs[100:200] = 0
    ys[200:300] = xs[200:300]
    for y, x in zip(ys, xs):
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(L.cvm_fast_atan2(float(y), float(x)))


@pytest.mark.parametrize("w,h,th", [(150, 120, 20), (150, 120, 7), (42, 43, 20), (41, 44, 7), (7, 7, 7), (300, 200, 35)])
def test_fast_model(w, h, th):
    L = lib()

    class KP(C.Structure):
        _fields_ = [("x", C.c_int), ("y", C.c_int), ("r", C.c_int)]

    img = synth.texture(max(w, 64), max(h, 64), seed=w + th)[:h, :w].copy()
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)]
    buf = (KP * 100000)()
    n = L.cvm_fast_detect(P(img), w, h, w, th, buf, 100000)
    assert ref == [(buf[i].x, buf[i].y, buf[i].r) for i in range(n)]


@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 0), (1280, 720, 2000, 1), (752, 480, 1200, 2), (480, 640, 500, 3)])
def test_pipeline_equals_cv2_backed_pipeline(w, h, nf, seed):
    img = synth.frame(w, h, seed)
    k, d, m = OrbOracle(nf).extract(img)
    k2, d2, m2 = extract_with_cv2(img, nf)
    assert m == m2 and np.array_equal(k, k2) and np.array_equal(d, d2)
    assert len(k) >= nf * 0.9


def test_constant_tables():
    """Known-answer constants the reference's code implies."""
    T = OrbOracle(1000).tables()
    assert list(T["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(T["per_level"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(OrbOracle(2000).tables()["per_level"]) == [434, 362, 302, 251, 209, 175, 145, 122]
