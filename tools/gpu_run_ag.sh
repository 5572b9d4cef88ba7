#!/bin/bash
# ncu --set full of the vectorised pyramid kernel (for profiles/traffic.json)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"pyramid4" -s 2 -c 1 -o gpurun_out/r2ag_full_pyramid python tests/gpu_profile_track.py 3 0 > gpurun_out/r2ag_ncu_pyr.log 2>&1; tail -2 gpurun_out/r2ag_ncu_pyr.log
ls -la gpurun_out/r2ag*.ncu-rep
