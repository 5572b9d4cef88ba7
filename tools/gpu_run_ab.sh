#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_track_gpu.py tests/test_host_adapters.py -m gpu -q -x --timeout 120 2>&1 | tail -2
DVM_POSE_PROFILE=1 timeout 120 python tests/gpu_profile_track.py 3 0 2>&1 | grep -E "pose-opt" | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ab_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frame_us", d["roofline"].get("frame_us"), d["roofline"]["chain_us"])
PY
tail -2 gpurun_out/r2ab_bench.err
