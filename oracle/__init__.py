"""CPU oracle for the DVM-SLAM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (dvmslam_b200) never does.

Parity status: the reference holds no tests or golden vectors for this path
(SURVEY.md section 8c).  The oracle is pinned instead against (1) the cv2 4.13.0
primitives the reference calls (tests/test_oracle_cv2.py, golden vectors under
tests/golden/ made by tests/golden/make_golden.py) and (2) the reference's own
sources compiled unmodified where they lie into oracle/_ref/ (`make ref`):
ORBextractor.cc against oracle/cvshim (tests/test_ref_build.py), ORBmatcher.cc with
the Frame / KeyFrame / MapPoint / Pinhole bodies it calls against oracle/slamshim
(tests/test_ref_matchers.py), and the optimisation functions of Optimizer.cc with
OptimizableTypes.cpp and the vendored g2o against the mini Eigen of oracle/g2oshim
(tests/test_ref_optimizer.py), and the vendored DBoW2 (the ORBVocabulary) against the stand-in
opencv2/core of oracle/dbowshim (tests/test_ref_dbow.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(ref: bool = False) -> None:
    subprocess.run(["make", "-C", _HERE, "-j8"] + (["ref"] if ref else []), check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        try:
            _LIB = ctypes.CDLL(path)
        except OSError:
            build()
            _LIB = ctypes.CDLL(path)
    return _LIB
