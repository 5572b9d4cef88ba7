"""Inter-agent loop-closure detection step (BASELINE.json config C3), one agent per rank/GPU.

Reference behaviour this replaces (W/ = src/slam_system/): every agent pushes the BoW vectors of its
new keyframes to its peers (`sendNewKeyFrameBows`, W/src/orb_slam3_wrapper.cpp:457-534; at least
MIN_BOW_SHARE_SIZE new keyframes, each keyframe sent to a peer once), and the agent with the LOWER id of
a pair looks for merge candidates among its own keyframes (`receiveNewKeyFrameBows` :536-600,
`isLeadNodeInGroup` :1238-1243) before descriptors are compared by SearchByBoW
(O3/src/ORBmatcher.cc:709-834) inside place recognition.

B200-native form: the keyframes' descriptor blocks u8[K, N, 32] live in HBM; the one real
exchange step is an all-to-all (`torch.distributed.all_to_all_single`: NCCL over NVLink on GPUs, gloo in
the CPU tests) in which each rank sends its not-yet-sent blocks to the ranks that own the pair; the owner
compares every received keyframe with every keyframe of its database by exhaustive nearest /
second-nearest Hamming search (dvm_hamming_knn_device) and reports the keyframe pairs whose number of
accepted descriptor matches reaches `min_matches`.  Control decisions stay on the host, as in the
reference.  There is no CPU matcher in the product: without CUDA tensors a `matcher` has to be injected
(the CPU tests inject the oracle to exercise the exchange logic).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

MIN_BOW_SHARE_SIZE = 5   # W/src/orb_slam3_wrapper.cpp:37: fewer new keyframes than this are not worth a message
N_BOW_MATCHES = 20       # O3/src/LoopClosing.cc:647,751: descriptor matches a candidate pair needs


def pair_owner(i: int, j: int, policy: str = "lead") -> int:
    """The rank that matches the keyframes of agents i and j.  "lead": the lower id, the reference's rule
    (only the lead node attempts a merge).  "balanced": the lower id when the ids differ by an odd number,
    else the higher one, so that every rank owns about (world - 1) / 2 pairs."""
    lo, hi = min(i, j), max(i, j)
    if policy == "lead" or (hi - lo) % 2 == 1:
        return lo
    return hi


class LoopClosureExchange:
    def __init__(self, n_feat: int = 2000, max_keyframes: int = 1024, device: Optional[torch.device] = None,
                 group=None, matcher: Optional[Callable] = None, th_low: int = 50, nnratio: float = 0.75,
                 min_matches: int = N_BOW_MATCHES, min_share: int = MIN_BOW_SHARE_SIZE, owner: str = "lead"):
        self.group = group
        self.owner = owner
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.n_feat, self.cap = n_feat, max_keyframes
        self.th_low, self.nnratio, self.min_matches, self.min_share = th_low, nnratio, min_matches, min_share
        self.db = torch.empty((max_keyframes, n_feat, 32), dtype=torch.uint8, device=self.device)
        self.n_kf = 0
        self.sent_upto = [0] * self.world      # per peer: keyframes [0, sent_upto) were already sent
        self._recv_upto = [0] * self.world     # per peer: keyframes received so far (= the next one's id there)
        self._matcher = matcher
        self._knn = None
        self.last_bytes_sent = 0
        if matcher is None:
            if self.device.type != "cuda":
                raise RuntimeError("LoopClosureExchange: no CUDA device and no matcher injected -- there is no CPU "
                                   "fallback for the Hamming search")
            from .matching import HammingKnn

            self._knn = HammingKnn(self.device.index or 0, torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ database
    def add_keyframes(self, desc_blocks) -> range:
        """Appends keyframes (u8[K, n_feat, 32], numpy or torch) to this agent's database; returns their ids."""
        t = torch.as_tensor(desc_blocks)
        if t.dtype != torch.uint8 or t.dim() != 3 or tuple(t.shape[1:]) != (self.n_feat, 32):
            raise ValueError(f"expected u8[K, {self.n_feat}, 32], got {t.dtype} {tuple(t.shape)}")
        k = t.shape[0]
        if self.n_kf + k > self.cap:
            raise ValueError("keyframe database is full")
        self.db[self.n_kf:self.n_kf + k].copy_(t, non_blocking=True)
        ids = range(self.n_kf, self.n_kf + k)
        self.n_kf += k
        return ids

    # ------------------------------------------------------------------ the exchange step
    def _plan(self) -> Tuple[List[int], List[int]]:
        """send counts per peer (only to the owner of the pair, only unsent keyframes, only when enough are new)."""
        send = [0] * self.world
        for p in range(self.world):
            if p == self.rank or pair_owner(self.rank, p, self.owner) != p:
                continue
            new = self.n_kf - self.sent_upto[p]
            send[p] = new if new >= self.min_share else 0
        counts = torch.tensor(send, dtype=torch.int64, device=self.device)
        recv = torch.empty_like(counts)
        if self.world > 1:
            dist.all_to_all_single(recv, counts, group=self.group)
        else:
            recv.copy_(counts)
        return send, [int(x) for x in recv.tolist()]

    def exchange(self) -> List[Tuple[int, int, int, int]]:
        """One exchange + matching round.  Returns merge candidates (peer rank, peer keyframe id, own
        keyframe id, accepted descriptor matches), best first, for the pairs this rank owns."""
        send, recv = self._plan()
        row = self.n_feat * 32
        # the keyframes not yet sent differ per peer only by their start: pack [start_p, n_kf) per peer
        parts = [self.db[self.sent_upto[p]:self.sent_upto[p] + send[p]] for p in range(self.world)]
        sendbuf = torch.cat(parts).reshape(-1) if sum(send) else torch.empty(0, dtype=torch.uint8, device=self.device)
        recvbuf = torch.empty(sum(recv) * row, dtype=torch.uint8, device=self.device)
        if self.world > 1:
            dist.all_to_all_single(recvbuf, sendbuf, [r * row for r in recv], [s * row for s in send], group=self.group)
        self.last_bytes_sent = int(sendbuf.numel())
        for p in range(self.world):
            self.sent_upto[p] += send[p]
        # ids of the received keyframes in the sender's numbering: the sender's sent_upto[self.rank] before this round
        peer_first = self._peer_first_ids(recv)
        out: List[Tuple[int, int, int, int]] = []
        off = 0
        for p in range(self.world):
            k = recv[p]
            if k == 0:
                continue
            blocks = recvbuf[off * row:(off + k) * row].view(k, self.n_feat, 32)
            off += k
            if self.n_kf == 0:
                continue
            counts = self.match_counts(blocks, self.db[:self.n_kf])
            ia, ib = np.nonzero(counts >= self.min_matches)
            out += [(p, peer_first[p] + int(a), int(b), int(counts[a, b])) for a, b in zip(ia, ib)]
        out.sort(key=lambda c: (-c[3], c[0], c[1], c[2]))
        return out

    def _peer_first_ids(self, recv: List[int]) -> List[int]:
        first = list(self._recv_upto)
        for p in range(self.world):
            self._recv_upto[p] += recv[p]
        return first

    # ------------------------------------------------------------------ matching
    def match_counts(self, a_blocks: torch.Tensor, b_blocks: torch.Tensor) -> np.ndarray:
        """int32[ka, kb]: per keyframe pair, the descriptors of a whose nearest descriptor in b is accepted
        (distance <= th_low and < nnratio * second-nearest distance)."""
        ka, kb = a_blocks.shape[0], b_blocks.shape[0]
        if self._matcher is not None:
            return np.asarray(self._matcher(a_blocks, b_blocks, self.th_low, self.nnratio), np.int32).reshape(ka, kb)
        a, b = a_blocks.contiguous(), b_blocks.contiguous()
        # keys are scratch here (only the counts leave the GPU); chunk the batch so they stay small
        chunk = max(1, min(ka, (64 << 20) // max(1, kb * self.n_feat * 8)))
        counts = torch.empty((ka, kb), dtype=torch.int32, device=a.device)
        k1 = torch.empty((chunk, kb, self.n_feat), dtype=torch.int32, device=a.device)
        k2 = torch.empty_like(k1)
        for s in range(0, ka, chunk):
            n = min(chunk, ka - s)
            self._knn.knn_device(a[s:].data_ptr(), n, self.n_feat, b.data_ptr(), kb, self.n_feat, k1.data_ptr(),
                                 k2.data_ptr(), counts[s:].data_ptr(), self.th_low, self.nnratio)
        self._knn.sync()
        return counts.cpu().numpy()

    def close(self):
        if self._knn is not None:
            self._knn.close()
            self._knn = None
