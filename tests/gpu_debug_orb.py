"""Stage-by-stage diagnostic dump (not a test): python tests/gpu_debug_orb.py > gpurun_out/orb_debug.txt"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvmslam_b200 import synth
from dvmslam_b200.extractor import ORBextractor
from oracle.orb import OrbOracle


def srt(a):
    return a[np.lexsort(a.T[::-1])] if len(a) else a


for (w, h, nf, seed) in [(640, 480, 1000, 0), (1280, 720, 2000, 2), (1241, 376, 2000, 4)]:
    img = synth.frame(w, h, seed)
    orc = OrbOracle(nf)
    k0, d0, m0 = orc.extract(img)
    ext = ORBextractor(nf, max_width=w, max_height=h)
    t = time.time()
    try:
        k1, d1, m1 = ext(img)
    except Exception as e:
        print("EXTRACT FAILED", e)
        continue
    print(f"== {w}x{h} nf={nf}: first call {1e3*(time.time()-t):.2f} ms; n oracle {len(k0)} gpu {len(k1)}; mono {m0} {m1}")
    for l in range(8):
        a, b = orc.level_image(l), ext.level_image(l)
        print(f" L{l} pyr equal={np.array_equal(a,b)} ndiff={(a!=b).sum() if a.shape==b.shape else 'shape'}")
        c0 = orc.level_candidates(l)
        ref = np.stack([c0['x'].astype(int)+16, c0['y'].astype(int)+16, c0['response'].astype(int)], 1)
        got = ext.level_keypoints(l, 0)
        eq = np.array_equal(srt(ref), srt(got))
        print(f"    cand oracle {len(ref)} gpu {len(got)} equal(set)={eq}")
        if not eq:
            sr, sg = set(map(tuple, ref)), set(map(tuple, got))
            print("     only-oracle", sorted(sr - sg)[:8], " only-gpu", sorted(sg - sr)[:8])
        s0 = orc.level_selected(l)
        ref = np.stack([s0['x'].astype(int), s0['y'].astype(int), s0['response'].astype(int)], 1)
        got = ext.level_keypoints(l, 1)
        eqo = np.array_equal(ref, got)
        eqs = np.array_equal(srt(ref), srt(got))
        print(f"    sel oracle {len(ref)} gpu {len(got)} equal(order)={eqo} equal(set)={eqs}")
        if not eqo and len(ref) and len(got):
            n = min(len(ref), len(got))
            bad = np.nonzero((ref[:n] != got[:n]).any(1))[0]
            print("     first diffs at", bad[:5], ref[bad[:3]].tolist(), got[bad[:3]].tolist())
        if len(s0):
            a, b = orc.level_image(l, True), ext.level_image(l, True)
            print(f"    blur equal={np.array_equal(a,b)}")
    if len(k0) == len(k1):
        for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
            print(f"  field {f}: equal={np.array_equal(k0[f], k1[f])} ndiff={(k0[f]!=k1[f]).sum()}")
        print(f"  desc equal={np.array_equal(d0,d1)} rows differing={(d0!=d1).any(1).sum()} bits={np.unpackbits(d0^d1).sum()}")
    # timing of repeated host calls
    for _ in range(3):
        ext(img)
    t = time.time()
    for _ in range(20):
        ext(img, copy=False)
    print(f"  host-call latency {(time.time()-t)/20*1e3:.3f} ms/frame")
    ext.close()
