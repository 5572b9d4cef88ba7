#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_adapters.py -m gpu -x -q > gpurun_out/r2f_pytest_adapters.log 2>&1
tail -30 gpurun_out/r2f_pytest_adapters.log
