#!/bin/bash
# the byte-gather pyramid kernel (fallback for unaligned caller images) against the plan with 4-aligned regions
DVM_PYRAMID_SCALAR=1 timeout 100 python -m pytest tests/test_orb_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -2
