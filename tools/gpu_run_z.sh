#!/bin/bash
# intermediate evidence: full GPU suite, smoke, bench N=1
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2z_pytest.log 2>&1; tail -4 gpurun_out/r2z_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "lba", {k: d["lba"][k] for k in ("value", "ms_per_ba_e2e", "ms_per_ba_kernel")}, "c5", d["c5"]["value"], d["c5"]["kernel_ms"])
PY
tail -2 gpurun_out/r2z_bench.err
