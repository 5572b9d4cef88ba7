#!/bin/bash
# round-2 final evidence: full GPU suite, smoke, bench, launch list, ncu --set full of the chain / extractor / BA / Hamming kernels
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2ad_pytest.log 2>&1; tail -4 gpurun_out/r2ad_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ad_smoke.log 2>&1; tail -2 gpurun_out/r2ad_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; tail -c 400 gpurun_out/r2ad_bench.json; tail -3 gpurun_out/r2ad_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ad_bench_reference.json 2> gpurun_out/r2ad_bench_reference.err; tail -c 300 gpurun_out/r2ad_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2ad_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ad_ncu_bench.log 2>&1; tail -1 gpurun_out/r2ad_ncu_bench.log | head -c 200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"pose_opt|match_|pyramid|fast_cells|describe|octree|grid_build" -s 22 -c 22 -o gpurun_out/r2ad_full_chain python tests/gpu_profile_track.py 6 0 > gpurun_out/r2ad_ncu_chain.log 2>&1; tail -2 gpurun_out/r2ad_ncu_chain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lba_kernel" -s 1 -c 1 -o gpurun_out/r2ad_full_lba python tests/gpu_profile_track.py 1 2 > gpurun_out/r2ad_ncu_lba.log 2>&1; tail -2 gpurun_out/r2ad_ncu_lba.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hamming_tc" -s 1 -c 1 -o gpurun_out/r2ad_full_knn python tests/gpu_bench_knn.py 64 64 > gpurun_out/r2ad_ncu_knn.log 2>&1; tail -2 gpurun_out/r2ad_ncu_knn.log
ls -la gpurun_out/r2ad*.ncu-rep
