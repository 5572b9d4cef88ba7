"""dvmslam_b200 -- B200-native (sm_100a) ORB front end and local back end of DVM-SLAM.

The product is the C-ABI shared library built from csrc/ (include/dvmslam_b200.h); the C++
drop-in adapters live in host/.  This Python package is only the thin ctypes mirror of the
reference's operator surface (ORBextractor / ORBmatcher / Optimizer) that the parity tests and
bench.py drive, plus the synthetic-input generators.
"""
from ._lib import DvmError, build, launch_count, lib  # noqa: F401
