#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_orb_gpu.py tests/test_track_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frame_us", d["roofline"].get("frame_us"), "extract", d["roofline"].get("extraction_stage_us"))
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pyramid" -s 2 -c 3 python tests/gpu_profile_track.py 6 0 2>&1 | grep -E "pyramid|gpu__time" | head -8
