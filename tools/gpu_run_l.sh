#!/bin/bash
# round-2 baseline of the restored tree: full GPU suite, smoke, bench, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; tail -5 gpurun_out/r2l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; tail -3 gpurun_out/r2l_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -c 7000 gpurun_out/r2l_bench.json; tail -5 gpurun_out/r2l_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2l_ncu_bench.log 2>&1; tail -2 gpurun_out/r2l_ncu_bench.log
