#!/bin/bash
# final tree: full GPU suite, smoke, bench N=1, reference arm, launch list
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2ah_pytest.log 2>&1; tail -3 gpurun_out/r2ah_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ah_smoke.log 2>&1; tail -1 gpurun_out/r2ah_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ah_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "lba", {k: d["lba"][k] for k in ("value", "ms_per_ba_e2e", "ms_per_ba_kernel")}, "c5", d["c5"]["value"], "stages", d["roofline"]["extraction_stage_us"], "hbm", d["roofline"]["hbm_stages"]["pyramid_resize_chain"])
PY
tail -2 gpurun_out/r2ah_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2ah_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ah_ncu_bench.log 2>&1; tail -1 gpurun_out/r2ah_ncu_bench.log | head -c 100
