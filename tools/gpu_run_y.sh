#!/bin/bash
mkdir -p gpurun_out
DVM_LBA_PROFILE=1 timeout 120 python tests/gpu_profile_track.py 1 3 2>&1 | grep -E "chol us|L2 us|lba phases|lba host|lba iters" | tail -9
