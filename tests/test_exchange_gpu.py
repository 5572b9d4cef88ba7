"""Inter-agent exchange step on B200s: the CUDA matcher behind LoopClosureExchange against the oracle, and
(when the box has two GPUs) the whole step over NCCL with one agent per GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_counts(A, B, th_low, nnratio):
    from oracle.bow import hamming_knn

    out = np.zeros((len(A), len(B)), np.int32)
    for i in range(len(A)):
        for j in range(len(B)):
            _, d1, d2 = hamming_knn(A[i], B[j])
            out[i, j] = int(((d1 <= th_low) & (d1.astype(np.float32) < np.float32(nnratio) * d2.astype(np.float32))).sum())
    return out


def test_match_counts_single_gpu():
    import torch

    from dvmslam_b200 import synth
    from dvmslam_b200.exchange import LoopClosureExchange

    N = 2000
    A = synth.keyframe_blocks(5, N, seed=1)
    B = synth.keyframe_blocks(7, N, seed=2, shared_from=np.concatenate([A, A[:2]]))
    ex = LoopClosureExchange(n_feat=N, max_keyframes=16, device=torch.device("cuda", 0))
    ex.add_keyframes(B)
    got = ex.match_counts(torch.from_numpy(A).cuda())
    want = _oracle_counts(A, B, 50, 0.75)
    assert np.array_equal(got, want)
    assert (np.diag(got[:5, :5]) > 400).all() and got[0, 5] > 400 and got[0, 1] < 20
    assert ex.exchange() == []   # world of one: nobody to exchange with
    ex.close()


WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DVM_ROOT"])
from dvmslam_b200 import synth
from dvmslam_b200.exchange import LoopClosureExchange
r = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
N = 2000
A = synth.keyframe_blocks(8, N, seed=1)
B = synth.keyframe_blocks(8, N, seed=2, shared_from=A)
ex = LoopClosureExchange(n_feat=N, max_keyframes=32, device=torch.device("cuda", r))
ex.add_keyframes(A[:6] if r == 0 else B)
c = ex.exchange()
json.dump({"cands": c, "sent": ex.last_bytes_sent}, open(os.path.join(os.environ["DVM_OUT"], f"rank{r}.json"), "w"))
ex.close()
dist.destroy_process_group()
'''


def test_exchange_two_gpus_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DVM_ROOT=ROOT, DVM_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29713", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    l0 = json.load(open(tmp_path / "rank0.json"))
    l1 = json.load(open(tmp_path / "rank1.json"))
    assert l1["cands"] == [] and l1["sent"] == 8 * 2000 * 32 and l0["sent"] == 0
    assert {(p, a, b) for p, a, b, n in l0["cands"]} == {(1, k, k) for k in range(6)}
    assert all(n > 300 for *_, n in l0["cands"])
