"""Inter-agent loop-closure detection step (BASELINE.json config C3), one agent per rank/GPU -- ctypes mirror of
dvm_exchange_* (include/dvmslam_b200.h, csrc/exchange.cu).

Reference behaviour this replaces (W/ = src/slam_system/): every agent pushes the BoW vectors of its new keyframes to its
peers (`sendNewKeyFrameBows`, W/src/orb_slam3_wrapper.cpp:457-534; at least MIN_BOW_SHARE_SIZE new keyframes, each keyframe
sent to a peer once), and the agent with the LOWER id of a pair looks for merge candidates among its own keyframes
(`receiveNewKeyFrameBows` :536-618, `isLeadNodeInGroup` :1238-1243).

All of the step lives in the library: the plan (who sends what to whom), grouped ncclSend / ncclRecv of the descriptor blocks
straight out of the HBM-resident database over NVLink, the exhaustive Hamming search (tcgen05 int8 kernel) of every received
keyframe against the local database, thresholds and ordering of the candidates.  This module only distributes the
ncclUniqueId (over torch.distributed's store -- the reference's agents would use their DDS channel) and exposes the calls.

Without GPUs the library has no matcher and no transport of its own: the CPU tests inject both through dvm_exchange_hooks
(`matcher=` here, the all-to-all carried by torch.distributed's gloo backend) and so exercise the library's host logic.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import numpy as np

from ._lib import check, lib

MIN_BOW_SHARE_SIZE = 5   # W/src/orb_slam3_wrapper.cpp:37: fewer new keyframes than this are not worth a message
N_BOW_MATCHES = 20       # O3/src/LoopClosing.cc:647,751: descriptor matches a candidate pair needs
_vp = C.c_void_p

_ALLTOALL = C.CFUNCTYPE(C.c_int, _vp, _vp, C.POINTER(C.c_size_t), _vp, C.POINTER(C.c_size_t))
_MATCH = C.CFUNCTYPE(C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp)


class _Hooks(C.Structure):
    _fields_ = [("ctx", _vp), ("alltoall", _ALLTOALL), ("match_counts", _MATCH)]


def pair_owner(i: int, j: int, policy: str = "lead") -> int:
    """The rank that matches the keyframes of agents i and j.  "lead": the lower id, the reference's rule (only the lead
    node attempts a merge).  "balanced": the lower id when the ids differ by an odd number, else the higher one, so that
    every rank owns about (world - 1) / 2 pairs.  (Mirrors pair_owner in csrc/exchange.cu.)"""
    lo, hi = min(i, j), max(i, j)
    if policy == "lead" or (hi - lo) % 2 == 1:
        return lo
    return hi


def _bind(L):
    if getattr(L, "_exchange_bound", False):
        return
    L.dvm_exchange_unique_id.argtypes = [_vp]
    L.dvm_exchange_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]
    L.dvm_exchange_destroy.argtypes = [_vp]
    L.dvm_exchange_destroy.restype = None
    L.dvm_exchange_set_policy.argtypes = [_vp, C.c_int, C.c_float, C.c_int, C.c_int]
    L.dvm_exchange_add_keyframes.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp]
    L.dvm_exchange_keyframes.argtypes = [_vp]
    L.dvm_exchange_reset.argtypes = [_vp]
    L.dvm_exchange_database.argtypes = [_vp]
    L.dvm_exchange_database.restype = _vp
    L.dvm_exchange_match_counts.argtypes = [_vp, _vp, C.c_int, _vp]
    L.dvm_exchange_round.argtypes = [_vp, _vp, C.c_int, _vp]
    L.dvm_exchange_last_bytes_sent.argtypes = [_vp]
    L.dvm_exchange_last_bytes_sent.restype = C.c_size_t
    L.dvm_exchange_last_match_ms.argtypes = [_vp]
    L.dvm_exchange_last_match_ms.restype = C.c_float
    L._exchange_bound = True


class LoopClosureExchange:
    def __init__(self, n_feat: int = 2000, max_keyframes: int = 1024, device=None, group=None,
                 matcher: Optional[Callable] = None, th_low: int = 50, nnratio: float = 0.75,
                 min_matches: int = N_BOW_MATCHES, min_share: int = MIN_BOW_SHARE_SIZE, owner: str = "lead"):
        import torch
        import torch.distributed as dist

        self.L = lib()
        _bind(self.L)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_feat, self.cap = n_feat, max_keyframes
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.h = _vp()
        self._keep = []
        hooks_ptr = None
        ident = None
        if matcher is not None:
            hooks_ptr = self._make_hooks(matcher)
        elif self.device.type != "cuda":
            raise RuntimeError("LoopClosureExchange: no CUDA device and no matcher injected -- there is no CPU fallback for "
                               "the Hamming search")
        elif self.world > 1:
            buf = np.zeros(128, np.uint8)
            if self.rank == 0:
                check(self.L.dvm_exchange_unique_id(buf.ctypes.data))
            obj = [buf.tobytes()]
            dist.broadcast_object_list(obj, src=0, group=group)   # the ncclUniqueId travels over the host-side channel
            ident = np.frombuffer(obj[0], np.uint8).copy()
            self._keep.append(ident)
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None
        check(self.L.dvm_exchange_create(C.byref(self.h), self.device.index or 0, self.rank, self.world,
                                         ident.ctypes.data if ident is not None else None, n_feat, max_keyframes,
                                         1 if owner == "balanced" else 0, _vp(stream) if stream else None, hooks_ptr))
        check(self.L.dvm_exchange_set_policy(self.h, th_low, nnratio, min_matches, min_share))
        self.min_matches = min_matches

    # ------------------------------------------------------------------ test hooks (CPU, gloo)
    def _make_hooks(self, matcher):
        import torch
        import torch.distributed as dist

        world, group, n_feat = self.world, self.group, self.n_feat

        def alltoall(ctx, send, send_bytes, recv, recv_bytes):
            try:
                sb = [send_bytes[p] for p in range(world)]
                rb = [recv_bytes[p] for p in range(world)]
                if sum(sb):
                    s = torch.from_numpy(np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), (sum(sb),)).copy())
                else:
                    s = torch.empty(0, dtype=torch.uint8)
                r = torch.empty(sum(rb), dtype=torch.uint8)
                dist.all_to_all_single(r, s, rb, sb, group=group)
                if sum(rb):
                    C.memmove(recv, r.numpy().ctypes.data, sum(rb))
                return 0
            except Exception:   # pragma: no cover
                import traceback

                traceback.print_exc()
                return 1

        def match_counts(ctx, a, ka, db, kb, nf, th_low, nnratio, counts):
            try:
                A = np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_uint8)), (ka, nf, 32))
                B = np.ctypeslib.as_array(C.cast(db, C.POINTER(C.c_uint8)), (kb, nf, 32))
                out = np.asarray(matcher(torch.from_numpy(A.copy()), torch.from_numpy(B.copy()), th_low, nnratio), np.int32).reshape(ka, kb)
                C.memmove(counts, np.ascontiguousarray(out).ctypes.data, ka * kb * 4)
                return 0
            except Exception:   # pragma: no cover
                import traceback

                traceback.print_exc()
                return 1

        del n_feat
        cb_a, cb_m = _ALLTOALL(alltoall), _MATCH(match_counts)
        hooks = _Hooks(None, cb_a, cb_m)
        self._keep += [cb_a, cb_m, hooks]
        return C.byref(hooks)

    # ------------------------------------------------------------------ database
    @property
    def n_kf(self) -> int:
        return int(self.L.dvm_exchange_keyframes(self.h))

    @property
    def db_ptr(self) -> int:
        """Device pointer (host pointer with hooks) of the database u8[max_keyframes][n_feat][32]."""
        return int(self.L.dvm_exchange_database(self.h) or 0)

    @property
    def last_bytes_sent(self) -> int:
        return int(self.L.dvm_exchange_last_bytes_sent(self.h))

    @property
    def last_match_ms(self) -> float:
        """Device time of the last match_counts call (CUDA events on the exchange's stream)."""
        return float(self.L.dvm_exchange_last_match_ms(self.h))

    def add_keyframes(self, desc_blocks) -> range:
        """Appends keyframes (u8[K, n_feat, 32], numpy or torch, host or device) to this agent's database; returns their ids."""
        import torch

        t = torch.as_tensor(desc_blocks)
        if t.dtype != torch.uint8 or t.dim() != 3 or tuple(t.shape[1:]) != (self.n_feat, 32):
            raise ValueError(f"expected u8[K, {self.n_feat}, 32], got {t.dtype} {tuple(t.shape)}")
        t = t.contiguous()
        if self.n_kf + t.shape[0] > self.cap:
            raise ValueError("keyframe database is full")
        first = C.c_int()
        check(self.L.dvm_exchange_add_keyframes(self.h, _vp(t.data_ptr()), int(t.shape[0]), int(t.is_cuda), C.addressof(first)))
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        return range(first.value, first.value + int(t.shape[0]))

    # ------------------------------------------------------------------ the exchange step
    def exchange(self) -> List[Tuple[int, int, int, int]]:
        """One exchange + matching round (collective).  Returns merge candidates (peer rank, peer keyframe id, own keyframe
        id, accepted descriptor matches), best first, for the pairs this rank owns."""
        cap = 4096
        while True:
            buf = np.zeros((cap, 4), np.int32)
            n = C.c_int()
            check(self.L.dvm_exchange_round(self.h, buf.ctypes.data, cap, C.addressof(n)))
            if n.value <= cap:
                return [tuple(int(v) for v in row) for row in buf[:n.value]]
            raise RuntimeError(f"{n.value} candidates exceed the buffer of {cap}")   # a round cannot be repeated

    def match_counts(self, a_blocks, b_blocks=None) -> np.ndarray:
        """int32[ka, n_kf]: per keyframe pair, the descriptors of a whose nearest descriptor in the DATABASE keyframe is
        accepted (distance <= th_low and < nnratio * second-nearest distance).  b_blocks is accepted for symmetry with the
        earlier interface and must be the database."""
        import torch

        a = torch.as_tensor(a_blocks).contiguous()
        out = np.zeros((a.shape[0], self.n_kf), np.int32)
        check(self.L.dvm_exchange_match_counts(self.h, _vp(a.data_ptr()), int(a.shape[0]), out.ctypes.data))
        return out

    def reset(self):
        """Empties the database and forgets what was sent / received (collective in effect: every rank does the same); the
        communicator stays."""
        check(self.L.dvm_exchange_reset(self.h))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_exchange_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
