#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_track_gpu.py tests/test_host_adapters.py -m gpu -q -x --timeout 120 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2aa_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "lba", d["lba"]["value"])
PY
tail -2 gpurun_out/r2aa_bench.err
