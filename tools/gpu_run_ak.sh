#!/bin/bash
timeout 100 python -m pytest tests/test_orb_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -3
