#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_lba_gpu.py tests/test_sim3_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/r2x_pytest.log 2>&1; tail -5 gpurun_out/r2x_pytest.log
DVM_LBA_PROFILE=1 timeout 120 python tests/gpu_profile_track.py 1 3 2>&1 | grep -E "chol us|lba phases|lba host|lba iters" | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "lba", {k: d["lba"][k] for k in ("value", "ms_per_ba_e2e", "ms_per_ba_kernel")}, "c5", d["c5"]["value"], d["c5"]["kernel_ms"])
PY
tail -2 gpurun_out/r2x_bench.err
