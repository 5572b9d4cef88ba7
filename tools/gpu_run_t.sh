#!/bin/bash
# scaling: the bench at N GPUs (one agent per GPU), launched as the driver does
N=$1
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2t_bench_n$N.json 2> gpurun_out/r2t_bench_n$N.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2t_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "lba", d["lba"]["value"], d["lba"]["ms_per_ba_e2e"], "exchange", {k: d["exchange"][k] for k in ("value", "ms_per_round", "bytes_exchanged_per_rank", "candidate_sets_correct_all_ranks")}, "c5", d["c5"]["value"])
PY
tail -2 gpurun_out/r2t_bench_n$N.err
