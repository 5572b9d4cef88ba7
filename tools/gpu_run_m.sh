#!/bin/bash
# in-kernel phase times of the tracking chain (globaltimer stamps; diagnostics)
mkdir -p gpurun_out
DVM_MATCH_PROFILE=1 DVM_POSE_PROFILE=1 timeout 600 python tests/gpu_profile_track.py 12 0 > gpurun_out/r2m_chain_phases.log 2>&1; tail -40 gpurun_out/r2m_chain_phases.log
