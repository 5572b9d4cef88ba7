#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bow_gpu.py -m gpu -x -q -k "sim3" > gpurun_out/r2e_pytest_sim3.log 2>&1
tail -30 gpurun_out/r2e_pytest_sim3.log
