#!/bin/bash
timeout 70 python -m pytest tests/test_lba_gpu.py -m gpu -q -x --timeout 60 2>&1 | tail -2
