#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -3 gpurun_out/r2g_smoke.log
timeout 600 python -m pytest tests/test_exchange_gpu.py tests/test_lba_gpu.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -5 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 6000 gpurun_out/r2g_bench.json; tail -5 gpurun_out/r2g_bench.err
