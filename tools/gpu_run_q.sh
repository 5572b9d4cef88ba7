#!/bin/bash
# round-2 evidence: full GPU suite, smoke, bench, launch list, ncu --set full of the chain / extractor / BA / Hamming kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2q_pytest.log 2>&1; tail -4 gpurun_out/r2q_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; tail -2 gpurun_out/r2q_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 600 gpurun_out/r2q_bench.json; tail -3 gpurun_out/r2q_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2q_ncu_bench.log 2>&1; tail -1 gpurun_out/r2q_ncu_bench.log | head -c 300
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"pose_opt|match_|fast_cells|describe|octree|resize|grid_build|lba_kernel" -s 60 -c 36 -o gpurun_out/r2q_full_chain python tests/gpu_profile_track.py 6 2 > gpurun_out/r2q_ncu_chain.log 2>&1; tail -2 gpurun_out/r2q_ncu_chain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hamming_tc|expand_desc" -s 2 -c 2 -o gpurun_out/r2q_full_knn python tests/gpu_bench_knn.py 16 16 > gpurun_out/r2q_ncu_knn.log 2>&1; tail -2 gpurun_out/r2q_ncu_knn.log
ls -la gpurun_out/*.ncu-rep
