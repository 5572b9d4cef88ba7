#!/bin/bash
mkdir -p gpurun_out
for mode in 2 3 4; do timeout 120 python tests/gpu_bench_knn.py 64 64 $mode 2>&1 | grep -v "1x1" | tee -a gpurun_out/r2c_knn_probe.log; done
