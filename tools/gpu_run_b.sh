#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bow_gpu.py -m gpu -x -q -k "hamming" > gpurun_out/r2b_pytest_hamming.log 2>&1
tail -15 gpurun_out/r2b_pytest_hamming.log
for mode in 1 2 3; do timeout 120 python tests/gpu_bench_knn.py 64 64 $mode 2>&1 | tee -a gpurun_out/r2b_knn_bench.log; done
