#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_track_gpu.py tests/test_bow_gpu.py -m gpu -q -x --timeout 60 > gpurun_out/r2o_pytest.log 2>&1; tail -4 gpurun_out/r2o_pytest.log
DVM_MATCH_PROFILE=1 timeout 120 python tests/gpu_profile_track.py 4 0 > gpurun_out/r2o_chain_phases.log 2>&1; tail -12 gpurun_out/r2o_chain_phases.log
