#!/bin/bash
# matcher rewrite (flattened warp walk, compact query list, monotone claims, fused retry, prior folded into pose-opt)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_track_gpu.py tests/test_bow_gpu.py tests/test_host_adapters.py tests/test_ref_matchers.py -m gpu -q -x --timeout 60 > gpurun_out/r2n_pytest.log 2>&1; tail -15 gpurun_out/r2n_pytest.log
DVM_MATCH_PROFILE=1 timeout 120 python tests/gpu_profile_track.py 6 0 > gpurun_out/r2n_chain_phases.log 2>&1; tail -12 gpurun_out/r2n_chain_phases.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "chain", d["roofline"].get("chain_us"), "frame_us", d["roofline"].get("frame_us"))
PY
tail -3 gpurun_out/r2n_bench.err
