#!/bin/bash
# compute-sanitizer over the bundle-adjustment tests after this round's kernel work (pair lists, look-ahead solve)
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 100 compute-sanitizer --tool $tool python -m pytest tests/test_lba_gpu.py -m gpu -q -x --timeout 90 > gpurun_out/r2ai_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'passed|failed|ERROR SUMMARY' gpurun_out/r2ai_sanitizer_$tool.log | tr '\n' ' ')"
done
