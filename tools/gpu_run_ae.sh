#!/bin/bash
# 2 GPUs: the NCCL exchange test and the bench at N = 2
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_exchange_gpu.py -m gpu -q --timeout 120 > gpurun_out/r2ae_pytest_n2.log 2>&1; tail -3 gpurun_out/r2ae_pytest_n2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ae_bench_n2.json 2> gpurun_out/r2ae_bench_n2.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ae_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"]["value"], "lba", d["lba"]["value"], "exchange", {k: d["exchange"][k] for k in ("value", "ms_per_round", "bytes_exchanged_per_rank", "candidate_sets_correct_all_ranks")})
PY
tail -2 gpurun_out/r2ae_bench_n2.err
