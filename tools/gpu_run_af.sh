#!/bin/bash
# 8 GPUs: the bench at N = 8 (one agent per GPU; exchange round over NCCL among the eight)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2af_bench_n8.json 2> gpurun_out/r2af_bench_n8.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2af_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "e2e", d["e2e"]["value"], "lba", {k: d["lba"][k] for k in ("value", "ms_per_ba_e2e", "ms_per_ba_kernel")}, "exchange", {k: d["exchange"][k] for k in ("value", "ms_per_round", "bytes_exchanged_per_rank", "candidate_sets_correct_all_ranks")})
PY
tail -2 gpurun_out/r2af_bench_n8.err
