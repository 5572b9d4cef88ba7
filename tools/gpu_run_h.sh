#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q > gpurun_out/r2h_pytest_exchange_2gpu.log 2>&1; tail -8 gpurun_out/r2h_pytest_exchange_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; tail -c 1500 gpurun_out/r2h_bench_n2.json; tail -5 gpurun_out/r2h_bench_n2.err
