#!/usr/bin/env python
"""bench.py -- headline benchmark of the DVM-SLAM hot path on B200 (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (config C2 of BASELINE.json, "tracking only"): one agent per GPU, a 1280x720 synthetic
stream over a textured plane, 2000 features per frame, a pre-built local map of 6000 points.  One
"step" = one batch of FRAMES_PER_STEP consecutive frames, each pushed through the whole per-frame hot
path in the reference's order: ExtractORB -> Frame grid -> SearchByProjection(last frame) ->
PoseOptimization -> isInFrustum + SearchByProjection(local map) -> PoseOptimization.
`value` = tracked frames/sec over all agents with the frames already resident in HBM (320 distinct
frames = 281 MB per agent, larger than the 126 MB L2, cycled; constant-velocity prior computed on the
device, no host synchronisation inside a step); `e2e` = the same frames through the C-ABI call with
pinned HOST images, the pose and inlier counts read back every frame (H2D and D2H inside the timed
region).  The line also carries `lba`: local-BA LM iterations/sec on config C4 (50 free + 10 fixed
keyframes, ~4900 points, ~57k observations) through dvm_local_ba with host arrays.
`--impl reference` times the CPU oracle (the restated reference path) on the host cores.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT = 1280, 720, 2000
FRAMES_PER_STEP = 64
RESIDENT_FRAMES = 320
C3_KEYFRAMES = 64   # keyframes per agent in the inter-agent exchange step
METRIC = "frames/sec/agent tracked (1280x720, 2000 features, tracking only) + local-BA iter/sec in `lba`"
UNIT = "frames/s"
WORKLOAD = "C2: single agent per GPU, 1280x720 synthetic stream, 2000 feats/frame"

# Algorithmic bytes per 1280x720 / 2000-keypoint frame, per stage:
#   pyramid: read levels 0..6 + write levels 1..7;  fast: read all levels;
#   describe: blur read+write of all levels + 749 B/kp orientation + 512 B/kp taps + 60 B/kp output
STAGE_NAMES = ["pyramid_resize_chain", "fast_cells", "octree_select", "orient_blur_brief"]
STAGE_BYTES = [2781331 + 1931488, 2853088, None, 5706176 + 1498000 + 1024000 + 120000]
FRAME_BYTES_TOTAL = 15914083


STAGE_KERNELS = ["pyramid4_kernel", "fast_cells_kernel", "octree_kernel", "describe_kernel"]


def ncu_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by profiles/summarize_ncu.py); None if absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(STAGE_KERNELS[stage])
    except Exception:
        return None


def ncu_traffic_of(kernel):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except Exception:
        return None


def measured_peak_bf16():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        return 1590.0   # fallback of B200_PROFILING.md


def measured_peak_fp64():
    """FP64 tensor-pipe (DMMA m8n8k4) peak measured by tools/dmma_peak.cu on this pool's B200 (profiles/r2_a_fp64_peak.jsonl)."""
    try:
        rows = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r2_a_fp64_peak.jsonl")) if l.strip()]
        return max(r["dmma_tflops"] for r in rows), "measured (tools/dmma_peak.cu, profiles/r2_a_fp64_peak.jsonl)"
    except Exception:
        return 37.0, "nominal"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


MAP_KEYFRAMES = list(range(0, RESIDENT_FRAMES, 40))
MAP_POINTS = 6000
BOUNDS = (0.0, 0.0, float(W), float(H))


def make_stream(seed: int):
    from dvmslam_b200 import synth

    return synth.OrbitStream(W, H, seed=seed, period=RESIDENT_FRAMES)


def make_frames(seed: int, n: int, stream=None) -> np.ndarray:
    s = stream or make_stream(seed)
    out = np.empty((n, H, W), np.uint8)
    for k in range(n):
        out[k] = s.frame(k)
    return out


def bench_config():
    return {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP, "local_map_points": MAP_POINTS,
            "l2": f"{RESIDENT_FRAMES} distinct resident frames ({RESIDENT_FRAMES * W * H >> 20} MB) cycled: inputs larger than L2"}


def lba_scene(seed: int = 0):
    from dvmslam_b200 import synth

    S = synth.ba_scene(50, 10, 5000, seed=seed)
    return S, (S["cam_q"], S["cam_t"], S["cam_fixed"], S["pts"], S["edge_cam"], S["edge_pt"], S["edge_obs"], S["edge_w"], S["K"])


# ------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    """One CPU agent: builds its map, bootstraps, then tracks n_frames frames with the oracle chain."""
    seed, n_frames, reps = args
    from dvmslam_b200 import synth
    from oracle.orb import OrbOracle
    from oracle.track import TrackerOracle

    S = make_stream(seed)
    o = OrbOracle(NFEAT)
    T = o.tables()
    M = synth.plane_map(S, o.extract, MAP_KEYFRAMES, T["scale"], MAP_POINTS)
    trk = TrackerOracle(o.extract, T, S.K, BOUNDS, M)
    R, t = S.pose(0)
    q = synth.quat_from_R(R).astype(np.float32)
    frames = make_frames(seed, n_frames + 1, S)
    trk.bootstrap(frames[0], q, t)
    pq, pt = q, np.asarray(t, np.float32)
    dt = 0.0
    for _ in range(reps):
        for k in range(1, n_frames + 1):
            t0 = time.perf_counter()
            r = trk.track(frames[k], pq, pt)
            dt += time.perf_counter() - t0
            pq, pt = r["q"], r["t"]
    return dt, n_frames * reps


def _lba_worker(args):
    seed, reps = args
    from oracle.lba import local_ba

    _, a = lba_scene(seed)
    its, dt = 0, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        r = local_ba(*a)
        dt += time.perf_counter() - t0
        its += r["iters"]
    return dt, its


def cpu_lba_iters_per_sec(cores: int, reps: int = 1):
    import multiprocessing as mp

    if cores == 1:
        dt, n = _lba_worker((0, reps))
        return n / dt
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_lba_worker, [(i % 3, reps) for i in range(cores)])
    return sum(n / dt for dt, n in res)


def cpu_stage_times(frame: np.ndarray, reps: int = 3):
    """Per-stage CPU time (ms per 1280x720 frame, single thread) of the OpenCV primitives the reference's front end calls --
    the resize chain of ComputePyramid, cv::FAST (threshold 20, NMS) and the 7x7 GaussianBlur over the 8 pyramid levels -- for
    the C models the CPU arm runs (oracle/cvmodels.c, bit-exact with cv2) AND for cv2's own SIMD code on the same images, so
    that the distance of the port from the real library is on record.  cv2 may be absent on a box: then only the models."""
    import ctypes as C

    from oracle import lib as oracle_lib

    L = oracle_lib()

    class KP(C.Structure):
        _fields_ = [("x", C.c_int), ("y", C.c_int), ("r", C.c_int)]

    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    sizes = [(W, H)]
    inv = np.float32(1)
    for _ in range(7):
        inv = inv * (np.float32(1) / np.float32(1.2))
        sizes.append((int(np.rint(np.float32(W) * inv)), int(np.rint(np.float32(H) * inv))))
    out = {}
    buf = (KP * 400000)()

    def timed(fn):
        best = 1e9
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t)
        return 1e3 * best

    levels = [frame]

    def chain_port():
        lv = [frame]
        for (w, h) in sizes[1:]:
            o = np.empty((h, w), np.uint8)
            L.cvm_resize_linear_u8(P(lv[-1]), lv[-1].shape[1], lv[-1].shape[0], lv[-1].shape[1], P(o), w, h, w)
            lv.append(o)
        levels[:] = lv

    out["resize_chain"] = {"port_ms": timed(chain_port)}
    out["fast_th20_nms"] = {"port_ms": timed(lambda: [L.cvm_fast_detect(P(l), l.shape[1], l.shape[0], l.shape[1], 20, buf, 400000) for l in levels])}

    def blur_port():
        for l in levels:
            o = np.empty_like(l)
            L.cvm_gaussian7_u8(P(l), l.shape[1], l.shape[0], l.shape[1], P(o), l.shape[1])

    out["gaussian_blur_7x7"] = {"port_ms": timed(blur_port)}
    if cv2 is not None:
        def chain_cv():
            lv = [frame]
            for (w, h) in sizes[1:]:
                lv.append(cv2.resize(lv[-1], (w, h), interpolation=cv2.INTER_LINEAR))

        det = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        out["resize_chain"]["cv2_ms"] = timed(chain_cv)
        out["fast_th20_nms"]["cv2_ms"] = timed(lambda: [det.detect(l) for l in levels])
        out["gaussian_blur_7x7"]["cv2_ms"] = timed(lambda: [cv2.GaussianBlur(l, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101) for l in levels])
        out["cv2_version"] = cv2.__version__
    return out


def cpu_oracle_fps(cores: int, frames_per_core: int, reps: int = 1):
    """frames/sec of the CPU oracle (restated reference front end) using `cores` processes."""
    import multiprocessing as mp

    if cores == 1:
        dt, n = _ref_worker((0, frames_per_core, reps))
        return n / dt
    with mp.get_context("fork").Pool(cores) as pool:
        t = time.perf_counter()
        res = pool.map(_ref_worker, [(i, frames_per_core, reps) for i in range(cores)])
        wall = time.perf_counter() - t
    # each worker times its own loop (frame synthesis excluded); aggregate = sum of per-worker rates
    del wall
    return sum(n / dt for dt, n in res)


def run_reference(args):
    """The CPU arm on this arm's config: one agent per --gpus slot, each agent's frame stream tracked by ONE thread --
    the reference's Tracking thread is single-threaded (W/src/ros_mono.cpp:30-43 calls TrackMonocular on the image
    callback thread; cv::FAST on ~40 px cells and g2o with OpenMP off do not fan out), so "all the host threads the
    path can use" for N agents is N.  The aggregate over every host core (one independent agent per core) is
    reported next to it as `all_cores` for scale."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    agents = max(1, min(args.gpus, cores))
    per_step = 16  # tracked frames per agent per step: a bounded sample of the same workload
    for _ in range(args.warmup and 1):
        cpu_oracle_fps(agents, 1)
    t = time.perf_counter()
    vals = [cpu_oracle_fps(agents, per_step) for _ in range(args.steps)]
    wall = time.perf_counter() - t
    v = float(np.mean(vals))
    lba_v = cpu_lba_iters_per_sec(agents, 1)
    all_fps = cpu_oracle_fps(cores, 2) if cores > agents else v
    all_lba = cpu_lba_iters_per_sec(cores, 1) if cores > agents else lba_v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_step * agents / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            # the same config block as our arm (the CPU arm times a bounded SAMPLE of that workload per step: cpu_baseline.sample)
            "config": bench_config(),
            # every step also rebuilds the agent's map and synthesises its frames (untimed, as in our arm)
            "wall_ms_per_step": 1e3 * wall / max(args.steps, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": agents, "kind": "port",
                             "sample": f"{agents} agent(s), one oracle process (thread) per agent, {per_step} tracked "
                                       f"frames/agent/step x {args.steps} steps"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "lba": {"value": lba_v, "unit": "LM iterations/s", "cores": agents, "kind": "port",
                    "sample": "one C4 local BA per agent (oracle/lba_oracle.cpp)"},
            "all_cores": {"cores": cores, "frames_per_s": all_fps, "lba_iters_per_s": all_lba,
                          "note": "one independent agent per host core: the box's aggregate CPU capacity, not one "
                                  "agent's rate"}}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch

    from dvmslam_b200 import launch_count, synth
    from dvmslam_b200.extractor import ORBextractor
    from dvmslam_b200.optimizer import LocalBA
    from dvmslam_b200.tracking import Tracker

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # this agent's synthetic stream, resident in HBM and mirrored in pinned host memory
    S = make_stream(seed=rank)
    frames_np = make_frames(rank, RESIDENT_FRAMES, S)
    host = torch.from_numpy(frames_np).pin_memory()
    dev = host.cuda(non_blocking=False)
    host_np = host.numpy()
    frame_bytes = W * H

    ext = ORBextractor(NFEAT, 1.2, 8, 20, 7, max_width=W, max_height=H, device=local_rank)
    T = ext.tables()
    world_map = synth.plane_map(S, lambda im: ext(im), MAP_KEYFRAMES, T["scale"], MAP_POINTS)
    trk = Tracker(ext, S.K, BOUNDS, world_map)
    stream = torch.cuda.ExternalStream(trk.stream(), device=local_rank)   # tracking chain (extraction: ext.stream())
    base = dev.data_ptr()
    R0, t0 = S.pose(0)
    q0 = synth.quat_from_R(R0).astype(np.float32)

    def bootstrap():
        return trk.bootstrap(host_np[0], q0, t0)

    inliers = []

    dev_pending = [0]

    def step_device(s):
        # the same hand-over as the host loop below (two frames ahead on the extraction streams), images already in HBM
        for i in range(FRAMES_PER_STEP):
            k = (1 + s * FRAMES_PER_STEP + i) % RESIDENT_FRAMES
            if dev_pending[0] == 0:
                trk.track((base + k * frame_bytes, W, H, W), sync=False)
            else:
                trk.track(None, sync=False)
                dev_pending[0] -= 1
            while dev_pending[0] < 2:
                kk = (k + 1 + dev_pending[0]) % RESIDENT_FRAMES
                trk.prefetch((base + kk * frame_bytes, W, H, W))
                dev_pending[0] += 1

    pending = [0]  # frames whose upload + extraction is already enqueued (dvm_tracker_prefetch; at most two)

    def step_host(s):
        # a camera-driven loop: frames k+1 and k+2 are handed to the two extraction streams while frame k's chain
        # runs; every frame's pose and counts reach the host (the chain's last kernel writes them into a pinned ring) and
        # are read one frame behind the enqueue, so the GPU is never idle between two frames of the agent
        for i in range(FRAMES_PER_STEP):
            k = (1 + s * FRAMES_PER_STEP + i) % RESIDENT_FRAMES
            if pending[0] == 0:
                trk.track(host_np[k], sync=False)
            else:
                trk.track(None, sync=False)
                pending[0] -= 1
            while pending[0] < 2:
                trk.prefetch(host_np[(k + 1 + pending[0]) % RESIDENT_FRAMES])
                pending[0] += 1
            if i > 0:
                _, _, c = trk.result(lag=1)
                inliers.append(c[3])
        _, _, c = trk.result()
        inliers.append(c[3])

    # ---- device-resident throughput (`value`) ----
    bootstrap()
    for s in range(args.warmup):
        step_device(s)
    q, t, c = trk.result()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in range(args.steps):
        step_device(args.warmup + s)
    e1.record(stream)
    q, t, c_dev = trk.result()
    barrier()
    launches = launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- end to end through the C-ABI host call (`e2e`) ----
    bootstrap()
    pending[0] = 0
    for s in range(min(args.warmup, 1)):
        step_host(s)
    inliers.clear()
    barrier()
    tw0 = time.perf_counter()
    for s in range(args.steps):
        step_host(1 + s)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - tw0], device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clock_info = clocks.stop() if rank == 0 else None
    e2e_total = float(e2e_s.item())
    tracked_ok = float(np.mean(np.array(inliers) >= 30)) if inliers else 0.0

    # ---- per-stage device time of the extractor for the roofline (separate profiled pass) ----
    ext.set_profiling(True)
    for i in range(FRAMES_PER_STEP):
        ext.extract_device(base + (i % RESIDENT_FRAMES) * frame_bytes, W, H, W)
    nprof, stage_ms = ext.get_profile()
    ext.set_profiling(False)
    stage_us = [1e3 * float(m) / max(nprof, 1) for m in stage_ms]

    # ---- per-segment device time of the tracking chain (the critical path), CUDA events on the chain's stream ----
    bootstrap()
    trk.set_profiling(True)
    for i in range(FRAMES_PER_STEP):
        trk.track((base + ((1 + i) % RESIDENT_FRAMES) * frame_bytes, W, H, W), sync=False)
    nchain, chain_ms = trk.get_profile()
    _, _, c_prof = trk.result()
    trk.set_profiling(False)
    CHAIN_NAMES = ["prior_and_reset", "search_by_projection_last", "pose_optimization_1", "search_local_points", "pose_optimization_2"]
    chain_us = [1e3 * float(m) / max(nchain, 1) for m in chain_ms]

    # ---- local BA (config C4) ----
    lba = None
    if rank == 0 or world > 1:
        solver = LocalBA(64, device=local_rank)
        _, a = lba_scene(rank % 3)
        for _ in range(3):
            solver.LocalBundleAdjustment(*a)
        barrier()
        reps, its, kern_ms = 20, 0, 0.0
        tl0 = time.perf_counter()
        for _ in range(reps):
            r = solver.LocalBundleAdjustment(*a)
            its += r["iters"]
            kern_ms += r["kernel_ms"]
        lba_s = torch.tensor([time.perf_counter() - tl0], device="cuda")
        if dist is not None:
            dist.all_reduce(lba_s, op=dist.ReduceOp.MAX)
        ne = len(a[4])
        lba = {"value": world * its / float(lba_s.item()), "unit": "LM iterations/s", "iterations_per_ba": its / reps,
               "ms_per_ba_e2e": 1e3 * float(lba_s.item()) / reps, "ms_per_ba_kernel": kern_ms / reps,
               "workload": f"C4: 50 free + 10 fixed keyframes, {len(a[3])} points, {ne} observations",
               "kernel_iters_per_s": its / (kern_ms * 1e-3),
               # ~29 MB and ~75 MFLOP per LM iteration at this size: latency-bound regime
               "achieved_GBps_kernel": 29e6 * its / (kern_ms * 1e-3) / 1e9}
        solver.close()

    # ---- config C5: the optimisation sequence of a map merge / loop closure on the MERGED map of 8 agents x 50 keyframes
    # (LoopClosing::MergeLocal -> welding BA, then OptimizeEssentialGraph, then the global BA; O3/src/LoopClosing.cc:1262-1810).
    # The reference runs it on the lead agent only: every rank computes a replica here ("replicas only") ----
    from dvmslam_b200.optimizer import EssentialGraphOptimizer

    c5 = None
    if rank == 0 or world > 1:
        W5 = synth.ba_scene(30, 10, 3000, seed=20 + rank % 3)                      # welding window: 30 adjustable + 10 fixed keyframes
        wa = (W5["cam_q"], W5["cam_t"], W5["cam_fixed"], W5["pts"], W5["edge_cam"], W5["edge_pt"], W5["edge_obs"], W5["edge_w"], W5["K"])
        G5 = synth.ba_scene_large(399, 1, 20000, seed=30 + rank % 3)              # the merged map: 400 keyframes
        ga = (G5["cam_q"], G5["cam_t"], G5["cam_fixed"], G5["pts"], G5["edge_cam"], G5["edge_pt"], G5["edge_obs"], G5["edge_w"], G5["K"])
        eg_in = synth.loop_pose_graph(400, seed=40 + rank % 3)
        big = LocalBA(400, device=local_rank)
        ego = EssentialGraphOptimizer(device=local_rank)
        big.MergeBundleAdjustment(*wa); ego.OptimizeEssentialGraph(*eg_in[:5]); big.BundleAdjustment(*ga, nIterations=10)   # warm-up
        barrier()
        reps5 = 3
        t50 = time.perf_counter()
        kms = np.zeros(3)
        for _ in range(reps5):
            rw = big.MergeBundleAdjustment(*wa)
            re = ego.OptimizeEssentialGraph(*eg_in[:5])
            rg = big.BundleAdjustment(*ga, nIterations=10)
            kms += [rw["kernel_ms"], re["kernel_ms"], rg["kernel_ms"]]
        c5_s = torch.tensor([time.perf_counter() - t50], device="cuda")
        if dist is not None:
            dist.all_reduce(c5_s, op=dist.ReduceOp.MAX)
        err_before = float(np.abs(eg_in[0][:, 4:7] - eg_in[5][:, 4:7]).max())
        err_after = float(np.abs(re["sim3"][:, 4:7] - eg_in[5][:, 4:7]).max())
        c5 = {"value": world * reps5 / float(c5_s.item()), "unit": "merge optimisation sequences/s",
              "ms_per_sequence_e2e": 1e3 * float(c5_s.item()) / reps5,
              "kernel_ms": {"welding_ba": kms[0] / reps5, "essential_graph": kms[1] / reps5, "global_ba": kms[2] / reps5},
              "workload": f"C5: welding BA 30+10 keyframes / {len(W5['pts'])} points; essential graph 400 keyframes / {len(eg_in[2])} "
                          f"edges (2793 unknowns, dense); global BA 399 free keyframes / {len(G5['pts'])} points / {len(G5['edge_cam'])} "
                          "observations (2394 x 2394 reduced system); replicas on every rank",
              "welding_ba": {"iters": rw["iters"], "chi_first": rw["chi_first"], "chi_last": rw["chi_last"]},
              "essential_graph": {"iters": re["iters"], "trials": re["trials"], "chi_first": re["chi_first"], "chi_last": re["chi_last"],
                                  "max_position_error_before_m": err_before, "max_position_error_after_m": err_after},
              "global_ba": {"iters": rg["iters"], "chi_first": rg["chi_first"], "chi_last": rg["chi_last"],
                            # ~n^3/3 FLOP per factorisation of the reduced system, one per LM trial
                            "reduced_solve_gflop_per_trial": 2400 ** 3 / 3 / 1e9},
              "checks": {"chi2_decreases": bool(rw["chi_last"] < rw["chi_first"] and re["chi_last"] < re["chi_first"] and rg["chi_last"] < rg["chi_first"])}}
        big.close()
        ego.close()

    # ---- inter-agent loop-closure exchange step (config C3): dvm_exchange_round = grouped ncclSend / ncclRecv of the new
    # keyframe descriptor blocks to the owner of each agent pair + exhaustive Hamming matching (tcgen05 int8 kernel)
    # against the local keyframe database; at N = 1 the matching alone.  Every agent's keyframe k shares 30 % of its
    # descriptors (with bit noise) with agent 0's keyframe k, so the candidate SET is known: it is checked, not counted ----
    from dvmslam_b200.exchange import LoopClosureExchange, pair_owner
    from dvmslam_b200.matching import HammingKnn

    base_kf = synth.keyframe_blocks(C3_KEYFRAMES, NFEAT, seed=100)
    own = base_kf if rank == 0 else synth.keyframe_blocks(C3_KEYFRAMES, NFEAT, seed=100 + rank, shared_from=base_kf)
    own_dev = torch.from_numpy(own).cuda()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xms, mms, ncand, cand_ok, nl0 = [], [], 0, True, launch_count()
    ex = LoopClosureExchange(n_feat=NFEAT, max_keyframes=C3_KEYFRAMES, device=torch.device("cuda", local_rank), owner="balanced")
    for rep in range(4):   # first repetition = warm-up (NCCL connection setup, buffer growth)
        ex.reset()         # a new session on the same communicator: empty database, nothing sent yet
        ex.add_keyframes(own_dev)
        barrier()
        x0.record()
        if world > 1:
            cands = ex.exchange()
            got = {(p, a, b) for p, a, b, _ in cands}
            want = {(p, k, k) for p in range(world) if p != rank and pair_owner(rank, p, "balanced") == rank
                    for k in range(C3_KEYFRAMES)}
            cand_ok = cand_ok and got == want
        else:
            cnt = ex.match_counts(own_dev)
            cands = np.argwhere(cnt >= ex.min_matches)
            cand_ok = cand_ok and {tuple(c) for c in cands} == {(k, k) for k in range(C3_KEYFRAMES)}
        x1.record()
        torch.cuda.synchronize()
        if rep:
            xms.append(x0.elapsed_time(x1))
            mms.append(ex.last_match_ms)
            ncand = len(cands)
    ex.close()
    xt = torch.tensor([float(np.mean(xms))], device="cuda")
    okt = torch.tensor([1.0 if cand_ok else 0.0], device="cuda")
    if dist is not None:
        dist.all_reduce(xt, op=dist.ReduceOp.MAX)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    kf_pairs = C3_KEYFRAMES * C3_KEYFRAMES * (world * (world - 1) // 2 if world > 1 else 1)
    exchange = {"value": kf_pairs / (float(xt.item()) * 1e-3), "unit": "keyframe pairs/s", "ms_per_round": float(xt.item()),
                "workload": f"C3: {C3_KEYFRAMES} keyframes x {NFEAT} descriptors per agent, every agent pair matched once "
                            f"(balanced ownership), {'grouped ncclSend/ncclRecv + ' if world > 1 else ''}exhaustive Hamming "
                            "(tcgen05 int8) through dvm_exchange_*",
                "descriptor_pairs_per_s": kf_pairs * NFEAT * NFEAT / (float(xt.item()) * 1e-3),
                "match_kernel_ms_rank0": float(np.mean(mms)),
                "bytes_exchanged_per_rank": C3_KEYFRAMES * NFEAT * 32 if world > 1 else 0,
                "candidates_rank0": int(ncand), "candidate_sets_correct_all_ranks": bool(okt.item() > 0.5),
                "gpu_launches": int(launch_count() - nl0)}
    # ---- the Hamming kernel alone (64 x 64 keyframe pairs of 2000 x 2000 descriptors), for its tensor-pipe roofline ----
    hk = HammingKnn(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    k1 = torch.empty((C3_KEYFRAMES, C3_KEYFRAMES, NFEAT), dtype=torch.int32, device="cuda")
    k2 = torch.empty_like(k1)
    cn = torch.empty((C3_KEYFRAMES, C3_KEYFRAMES), dtype=torch.int32, device="cuda")
    ham_ms = {}
    for mode in (2, 3, 1):   # tcgen05 kernel, its MMA-only probe, the popcount kernel it replaced
        hk.set_mode(mode)
        for it in range(2 if mode != 1 else 1):
            hk.knn_device(own_dev.data_ptr(), C3_KEYFRAMES, NFEAT, own_dev.data_ptr(), C3_KEYFRAMES, NFEAT, k1.data_ptr(),
                          k2.data_ptr(), cn.data_ptr(), 50, 0.75)
        torch.cuda.synchronize()
        reps_h = 5 if mode != 1 else 2
        x0.record()
        for it in range(reps_h):
            hk.knn_device(own_dev.data_ptr(), C3_KEYFRAMES, NFEAT, own_dev.data_ptr(), C3_KEYFRAMES, NFEAT, k1.data_ptr(),
                          k2.data_ptr(), cn.data_ptr(), 50, 0.75)
        x1.record()
        torch.cuda.synchronize()
        ham_ms[mode] = x0.elapsed_time(x1) / reps_h
    hk.close()
    del k1, k2, cn

    if rank == 0:
        frames = args.steps * FRAMES_PER_STEP
        value = world * frames / (ms_total * 1e-3)
        e2e = world * frames / e2e_total
        peak, peak_src = measured_peak_hbm()
        frame_us = 1e3 * ms_total / frames
        # The kernel with the largest share of the critical path: the frame rate is bounded by the tracking chain (pose k ->
        # prior k + 1), extraction runs ahead on its own streams.  PoseOptimization (two launches per frame) is that kernel:
        # one 8-CTA cluster running 4 x 10 dependent LM passes over ~1000 edges, i.e. latency-bound by construction.  Its
        # algorithmic bytes: 24 B per edge (world point 12 + keypoint 8 + weight 4), read once.
        pose_us = 0.5 * (chain_us[2] + chain_us[4])
        pose_edges = max(int(c_prof[1]), 1)
        pose_bytes = 24 * pose_edges
        pose_gbs = pose_bytes / (pose_us * 1e-6) / 1e9
        roofline = {"bound": "latency", "kernel": "pose_opt_kernel", "achieved": pose_gbs, "peak": peak, "unit": "GB/s",
                    "frac": pose_gbs / peak, "traffic": ncu_traffic_of("pose_opt_kernel"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": pose_bytes, "avg_launch_us": pose_us, "launches_per_frame": 2,
                    "share_of_chain": (chain_us[2] + chain_us[4]) / max(sum(chain_us), 1e-9),
                    "note": "critical-path kernel; bound by its dependent LM passes (40 per launch), not by HBM",
                    "chain_us": dict(zip(CHAIN_NAMES, chain_us)), "chain_us_total": sum(chain_us), "frame_us": frame_us,
                    "extraction_stage_us": dict(zip(STAGE_NAMES, stage_us)),
                    # the HBM-facing stages of the frame with their own algorithmic bytes, measured in
                    # the profiled extraction pass; `traffic` = DRAM bytes per launch from the committed ncu capture
                    "hbm_stages": {STAGE_NAMES[i]: {"algorithmic_bytes": STAGE_BYTES[i],
                                                    "achieved_GBps": STAGE_BYTES[i] / (stage_us[i] * 1e-6) / 1e9,
                                                    "frac": STAGE_BYTES[i] / (stage_us[i] * 1e-6) / 1e9 / peak,
                                                    "traffic": ncu_traffic(i)}
                                   for i in range(4) if STAGE_BYTES[i] is not None},
                    "whole_frame": {"algorithmic_bytes": FRAME_BYTES_TOTAL,
                                    "achieved_GBps": FRAME_BYTES_TOTAL * value / world / 1e9,
                                    "frac": FRAME_BYTES_TOTAL * value / world / 1e9 / peak}}
        # ---- per-leg rooflines ----
        fp64_peak, fp64_src = measured_peak_fp64()
        if lba is not None:
            # ~75 MFLOP of FP64 per LM iteration at C4 (linearise + Schur ~65, reduced solve ~9)
            lba_tf = 75e6 * lba["kernel_iters_per_s"] / 1e12
            lba["roofline"] = {"bound": "latency", "kernel": "lba_kernel", "achieved": lba_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                               "frac": lba_tf / fp64_peak, "peak_source": fp64_src,
                               "algorithmic_flop_per_iteration": 75e6, "algorithmic_bytes_per_iteration": 29e6,
                               "achieved_GBps": 29e6 * lba["kernel_iters_per_s"] / 1e9,
                               "traffic": ncu_traffic_of("lba_kernel"),
                               "note": "FP64 (DMMA m8n8k4 in the reduced solve); a 300 x 300 system is ~9 MFLOP: the kernel is "
                                       "bound by its dependent phases, not by the FP64 pipe"}
        int8_peak = 2.0 * measured_peak_bf16()
        ham_ops = 2.0 * 256 * (C3_KEYFRAMES * NFEAT) ** 2
        roofline_hamming = {"bound": "tensor", "kernel": "hamming_tc_kernel", "achieved": ham_ops / (ham_ms[2] * 1e-3) / 1e12,
                            "peak": int8_peak, "unit": "TOP/s", "frac": ham_ops / (ham_ms[2] * 1e-3) / 1e12 / int8_peak,
                            "peak_source": "2 x measured bf16 cuBLAS peak (MEASURED_PEAKS.json): the int8 tcgen05 rate is twice bf16's",
                            "ms": ham_ms[2], "mma_only_probe_ms": ham_ms[3], "popcount_kernel_ms": ham_ms[1],
                            "workload": f"{C3_KEYFRAMES} x {C3_KEYFRAMES} keyframe pairs of {NFEAT} x {NFEAT} descriptors, 256-bit",
                            "descriptor_pairs_per_s": (C3_KEYFRAMES * NFEAT) ** 2 / (ham_ms[2] * 1e-3),
                            "traffic": ncu_traffic_of("hamming_tc_kernel")}
        cpu = None
        if world == 1:
            tc = time.perf_counter()
            cpu_frames, cpu_bas = 192, 24   # ~8 s + ~2 s of single-thread CPU work
            cpu_fps = cpu_oracle_fps(1, cpu_frames)
            cpu_lba = cpu_lba_iters_per_sec(1, cpu_bas)
            cpu = {"value": cpu_fps, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{cpu_frames} tracked frames of the same stream + {cpu_bas} C4 local BAs, single thread, "
                             f"{time.perf_counter() - tc:.1f} s including map and frame synthesis", "lba_iters_per_s": cpu_lba,
                   # whole-level primitives on one frame of the stream: the C models of the port beside cv2's own code
                   "opencv_primitives_ms_per_frame": cpu_stage_times(frames_np[1])}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8 (extract/match) + f64 (pose, BA)",
                "data": "synthetic",
                "config": bench_config(),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": FRAMES_PER_STEP * frame_bytes,
                        "d2h_bytes_per_step": FRAMES_PER_STEP * 48, "tracked_ok_frac": tracked_ok,
                        "median_inliers": float(np.median(inliers)) if inliers else 0.0},
                "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu,
                "lba": lba, "exchange": exchange, "c5": c5, "roofline_hamming": roofline_hamming,
                "last_counts_device_run": list(c_dev)}
        print(json.dumps(line), flush=True)
    trk.close()
    ext.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
