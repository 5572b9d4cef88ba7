#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lba_gpu.py -m gpu -x -q > gpurun_out/r2i_pytest_lba.log 2>&1; tail -25 gpurun_out/r2i_pytest_lba.log
