"""GPU parity of the ORB extractor: the CUDA path, called through the C-ABI, must reproduce the
oracle bit for bit -- pyramid, per-cell FAST candidates, octree selection (set and order), keypoints
(coordinates, angle, response, octave, size) and descriptors, in the reference's output order."""
import numpy as np
import pytest

from dvmslam_b200 import synth

pytestmark = pytest.mark.gpu

CASES = [  # (w, h, nfeatures, seed)
    (640, 480, 1000, 0),   # config C1
    (640, 480, 2000, 1),
    (1280, 720, 2000, 2),  # config C2 frame size
    (752, 480, 1200, 3),   # EuRoC
    (1241, 376, 2000, 4),  # KITTI: nIni = 4 root nodes
    (480, 640, 500, 5),    # portrait: round(448/608) = 1 root node
]


def _sorted_rows(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])] if len(a) else a


@pytest.fixture(scope="module")
def extractors():
    cache = {}
    yield cache
    for e in cache.values():
        e.close()


def _get(cache, nf, w, h):
    from dvmslam_b200.extractor import ORBextractor

    key = (nf, w, h)
    if key not in cache:
        cache[key] = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h)
    return cache[key]


@pytest.mark.parametrize("w,h,nf,seed", CASES)
def test_extract_matches_oracle(extractors, w, h, nf, seed):
    from oracle.orb import OrbOracle

    img = synth.frame(w, h, seed)
    orc = OrbOracle(nf)
    k0, d0, m0 = orc.extract(img)
    ext = _get(extractors, nf, w, h)
    k1, d1, m1 = ext(img)
    # stage-wise first, so a failure names the stage
    for l in range(8):
        assert np.array_equal(orc.level_image(l), ext.level_image(l)), f"pyramid level {l}"
    for l in range(8):
        c0 = orc.level_candidates(l)
        ref = np.stack([c0["x"].astype(int) + 16, c0["y"].astype(int) + 16, c0["response"].astype(int)], 1)
        got = ext.level_keypoints(l, 0)
        assert np.array_equal(_sorted_rows(ref), _sorted_rows(got)), f"FAST candidates level {l}"
    for l in range(8):
        s0 = orc.level_selected(l)
        ref = np.stack([s0["x"].astype(int), s0["y"].astype(int), s0["response"].astype(int)], 1)
        got = ext.level_keypoints(l, 1)
        assert np.array_equal(ref, got), f"octree selection level {l} (set and order)"
    for l in range(8):
        if len(orc.level_selected(l)):
            assert np.array_equal(orc.level_image(l, True), ext.level_image(l, True)), f"blur level {l}"
    assert m0 == m1
    assert len(k0) == len(k1)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(k0[f], k1[f]), f
    assert np.array_equal(k0["angle"], k1["angle"]), "angle"
    assert np.array_equal(d0, d1), "descriptors"


def test_repeatable_and_handle_reuse(extractors):
    """Same handle, different frames and back: results do not depend on history."""
    from oracle.orb import OrbOracle

    ext = _get(extractors, 1000, 640, 480)
    a = synth.frame(640, 480, 10)
    b = synth.frame(640, 480, 11)
    ka, da, ma = ext(a)
    ext(b)
    ka2, da2, ma2 = ext(a)
    assert ma == ma2 and np.array_equal(ka, ka2) and np.array_equal(da, da2)
    k0, d0, m0 = OrbOracle(1000).extract(a)
    assert np.array_equal(k0, ka) and np.array_equal(d0, da)


def test_smaller_image_on_larger_handle(extractors):
    from oracle.orb import OrbOracle

    ext = _get(extractors, 2000, 1280, 720)
    img = synth.frame(800, 600, 12)
    k1, d1, m1 = ext(img)
    k0, d0, m0 = OrbOracle(2000).extract(img)
    assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)
    ext(synth.frame(1280, 720, 2))  # and back to the full size


def test_strided_input_and_lapping_area(extractors):
    from oracle.orb import OrbOracle

    ext = _get(extractors, 1000, 640, 480)
    big = synth.frame(700, 480, 13)
    view = big[:, 30:670]  # stride 700, width 640
    for lap in [(0, 1000), (0, 0), (200, 400), (-5, -1)]:
        k1, d1, m1 = ext(view, lap)
        k0, d0, m0 = OrbOracle(1000).extract(np.ascontiguousarray(view), lap)
        assert m0 == m1, lap
        assert np.array_equal(k0, k1) and np.array_equal(d0, d1), lap


def test_flat_and_noise_images(extractors):
    from oracle.orb import OrbOracle

    ext = _get(extractors, 1000, 640, 480)
    flat = np.full((480, 640), 77, np.uint8)
    k, d, m = ext(flat)
    assert len(k) == 0 and m == 0
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)  # the densest candidate field
    k1, d1, m1 = ext(noise)
    k0, d0, m0 = OrbOracle(1000).extract(noise)
    assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)
    half = flat.copy()
    half[:, 320:] = synth.frame(640, 480, 14)[:, 320:]
    k1, d1, m1 = ext(half)
    k0, d0, m0 = OrbOracle(1000).extract(half)
    assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)


def test_empty_image_returns_minus_one(extractors):
    ext = _get(extractors, 1000, 640, 480)
    k, d, m = ext(np.zeros((0, 0), np.uint8))
    assert m == -1 and len(k) == 0


def test_init_extractor_5x_features():
    """Mono initialisation builds a second extractor with 5*nFeatures (O3/src/Tracking.cc:581)."""
    from dvmslam_b200.extractor import ORBextractor
    from oracle.orb import OrbOracle

    img = synth.frame(1280, 720, 20)
    ext = ORBextractor(10000, 1.2, 8, 20, 7, max_width=1280, max_height=720)
    k1, d1, m1 = ext(img)
    k0, d0, m0 = OrbOracle(10000).extract(img)
    ext.close()
    assert m0 == m1 and np.array_equal(k0, k1) and np.array_equal(d0, d1)


def test_tables_match_oracle():
    from dvmslam_b200.extractor import ORBextractor
    from oracle.orb import OrbOracle

    for nf in (500, 1000, 1200, 2000, 10000):
        ext = ORBextractor(nf, 1.2, 8, 20, 7, max_width=640, max_height=480)
        T1, T0 = ext.tables(), OrbOracle(nf).tables()
        ext.close()
        for key in ("scale", "inv_scale", "sigma2", "inv_sigma2", "per_level"):
            assert np.array_equal(T0[key], T1[key]), (nf, key)


def test_unaligned_device_image_takes_the_byte_gather_pyramid(extractors):
    """A caller's image read in place from an odd address with an odd pitch (a cv::Mat ROI): the word-load pyramid kernel does
    not apply, the byte-gather one must give the same levels, candidates and selection."""
    import torch
    from oracle.orb import OrbOracle

    w, h, stride = 640, 480, 701
    img = synth.frame(w, h, 21)
    orc = OrbOracle(1000)
    orc.extract(img)
    ext = _get(extractors, 1000, 640, 480)
    buf = torch.zeros(h * stride + 16, dtype=torch.uint8, device="cuda")
    view = buf[1:1 + h * stride].view(h, stride)
    view[:, :w] = torch.from_numpy(img).cuda()
    torch.cuda.synchronize()
    ext.extract_device(buf.data_ptr() + 1, w, h, stride)
    ext.sync()
    for l in range(8):
        assert np.array_equal(orc.level_image(l), ext.level_image(l)), f"pyramid level {l}"
    for l in range(8):
        s0 = orc.level_selected(l)
        ref = np.stack([s0["x"].astype(int), s0["y"].astype(int), s0["response"].astype(int)], 1)
        assert np.array_equal(ref, ext.level_keypoints(l, 1)), f"octree selection level {l}"
