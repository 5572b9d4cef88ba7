#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sim3_gpu.py tests/test_lba_gpu.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -30 gpurun_out/r2j_pytest.log
