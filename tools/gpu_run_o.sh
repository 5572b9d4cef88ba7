#!/bin/bash
mkdir -p gpurun_out
DVM_MATCH_PROFILE=1 timeout 120 python tests/gpu_profile_track.py 4 0 > gpurun_out/r2o_chain_phases.log 2>&1; tail -16 gpurun_out/r2o_chain_phases.log
