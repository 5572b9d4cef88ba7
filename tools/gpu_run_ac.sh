#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_orb_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ac_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frame_us", d["roofline"].get("frame_us"), d["roofline"]["extraction_stage_us"])
PY
tail -2 gpurun_out/r2ac_bench.err
