"""World-size-2 test of the inter-agent exchange step (config C3) on CPU: `gloo` carries the all-to-all,
the oracle's exhaustive Hamming search is injected as the matcher (the product has no CPU matcher), so
what is tested is the host logic -- who sends what to whom, ragged counts, ids, thresholds."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DVM_ROOT"])
from dvmslam_b200 import synth
from dvmslam_b200.exchange import LoopClosureExchange
from oracle.bow import hamming_knn

def oracle_counts(a, b, th_low, nnratio):
    a, b = a.numpy(), b.numpy()
    out = np.zeros((len(a), len(b)), np.int32)
    for i in range(len(a)):
        for j in range(len(b)):
            idx, d1, d2 = hamming_knn(a[i], b[j])
            out[i, j] = int(((d1 <= th_low) & (d1.astype(np.float32) < np.float32(nnratio) * d2.astype(np.float32))).sum())
    return out

dist.init_process_group("gloo")
r = dist.get_rank()
N = 120
A = synth.keyframe_blocks(9, N, seed=1)                      # agent 0
B = synth.keyframe_blocks(9, N, seed=2, shared_from=A)       # agent 1: keyframe k overlaps agent 0's keyframe k
mine = A if r == 0 else B
ex = LoopClosureExchange(n_feat=N, max_keyframes=32, matcher=oracle_counts, min_matches=15)
log = {}
ex.add_keyframes(mine[:3] if r == 0 else mine[:6])           # ragged: 3 vs 6 new keyframes
log["round1"] = ex.exchange(); log["sent1"] = ex.last_bytes_sent
log["round2"] = ex.exchange(); log["sent2"] = ex.last_bytes_sent     # nothing new (or below MIN_BOW_SHARE_SIZE)
ex.add_keyframes(mine[3:9] if r == 0 else mine[6:9])         # agent 1 now has 3 new: below the share size of 5
log["round3"] = ex.exchange(); log["sent3"] = ex.last_bytes_sent
ex.add_keyframes(mine[:2])                                   # agent 1: 5 unsent (ids 6..10) -> sent
log["round4"] = ex.exchange(); log["sent4"] = ex.last_bytes_sent
json.dump(log, open(os.path.join(os.environ["DVM_OUT"], f"rank{r}.json"), "w"))
dist.destroy_process_group()
'''


def test_exchange_world2_gloo(tmp_path):
    import json

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DVM_ROOT=ROOT, DVM_OUT=str(tmp_path), OMP_NUM_THREADS="1")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    l0 = json.load(open(tmp_path / "rank0.json"))
    l1 = json.load(open(tmp_path / "rank1.json"))
    row = 120 * 32
    # rank 1 (higher id) sends, rank 0 (pair owner) matches; rank 1 never reports candidates
    assert all(l1[f"round{k}"] == [] for k in (1, 2, 3, 4))
    assert l0["sent1"] == 0 and l1["sent1"] == 6 * row
    # round 1: agent 0 holds keyframes 0..2, receives 1's keyframes 0..5: overlaps are (k, k) for k < 3
    got = {(p, a, b) for p, a, b, n in l0["round1"]}
    assert got == {(1, 0, 0), (1, 1, 1), (1, 2, 2)}
    assert all(n >= 15 for _, _, _, n in l0["round1"])
    assert l0["round1"] == sorted(l0["round1"], key=lambda c: (-c[3], c[0], c[1], c[2]))
    # round 2: nothing new to send
    assert l0["round2"] == [] and l1["sent2"] == 0
    # round 3: agent 1 has only 3 new keyframes (< MIN_BOW_SHARE_SIZE): held back
    assert l0["round3"] == [] and l1["sent3"] == 0
    # round 4: agent 1 sends ids 6..10 (5 keyframes); ids 9, 10 are copies of its keyframes 0, 1, which
    # overlap agent 0's keyframes 0, 1 and their copies 9, 10 (agent 0 also re-added its first two);
    # ids 6..8 overlap agent 0's 6..8 (added in round 3)
    assert l1["sent4"] == 5 * row
    got = {(p, a, b) for p, a, b, n in l0["round4"]}
    assert got == {(1, 6, 6), (1, 7, 7), (1, 8, 8), (1, 9, 0), (1, 10, 1), (1, 9, 9), (1, 10, 10)}


def test_exchange_needs_cuda_or_matcher():
    from dvmslam_b200.exchange import LoopClosureExchange, pair_owner

    with pytest.raises(RuntimeError, match="no CPU"):
        LoopClosureExchange(n_feat=10, max_keyframes=4)
    assert pair_owner(3, 1) == 1 and pair_owner(0, 7) == 0
