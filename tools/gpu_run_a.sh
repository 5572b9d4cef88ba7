#!/bin/bash
# round-2 run A: baseline GPU tests, compute-sanitizer over the BA tests, FP64 peaks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
./build/dmma_peak > gpurun_out/r2a_dmma_peak.jsonl 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
for tool in memcheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_lba_gpu.py -m gpu -q > gpurun_out/r2a_sanitizer_$tool.log 2>&1
done
tail -3 gpurun_out/r2a_pytest.log; tail -4 gpurun_out/r2a_sanitizer_*.log; cat gpurun_out/r2a_dmma_peak.jsonl
