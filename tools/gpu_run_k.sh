#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("c5"), indent=1))
print("value", d["value"], "e2e", d["e2e"]["value"])
PY
tail -5 gpurun_out/r2k_bench.err
