"""Seeded synthetic inputs for parity tests and bench.py.

Nothing here is on the product path: it only manufactures the 8-bit grey
images, the plane-sweep camera stream and the bundle-adjustment scene that
BASELINE.json's configs name ("synthetic frame", "synthetic stream").
"""
from __future__ import annotations

import numpy as np

# rpi_cam.yaml intrinsics with the distortion zeroed
# (reference: src/slam_system/configs/rpi_cam.yaml:23-26)
RPI_K = (994.3, 993.4, 638.0, 372.6)


def _box(a: np.ndarray, r: int) -> np.ndarray:
    """(2r+1)^2 box filter with edge replication, float32, via integral image."""
    if r <= 0:
        return a
    p = np.pad(a, r, mode="edge").astype(np.float64)
    ii = np.zeros((p.shape[0] + 1, p.shape[1] + 1))
    ii[1:, 1:] = p.cumsum(0).cumsum(1)
    k = 2 * r + 1
    s = ii[k:, k:] - ii[:-k, k:] - ii[k:, :-k] + ii[:-k, :-k]
    return (s / (k * k)).astype(np.float32)


def texture(w: int, h: int, seed: int = 0, n_shapes: int = 400, flat_frac: float = 0.06) -> np.ndarray:
    """8-bit grey texture: 3 octaves of box-filtered uniform noise plus random filled
    rectangles/discs (corners for FAST at several scales) and a few flat patches so
    that the minThFAST fallback and empty cells are exercised."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    for r, amp in ((1, 38.0), (3, 52.0), (8, 64.0)):
        img += amp * (_box(rng.random((h, w), dtype=np.float32), r) - 0.5) * (2 * r + 1) * 0.55
    img += 128.0
    yy, xx = np.mgrid[0:h, 0:w]
    n_shapes = int(n_shapes * (w * h) / (640.0 * 480.0))
    for _ in range(n_shapes):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        s = int(rng.integers(4, max(5, min(w, h) // 8)))
        val = float(rng.integers(20, 236))
        x0, x1 = max(cx - s, 0), min(cx + s, w)
        y0, y1 = max(cy - s, 0), min(cy + s, h)
        if rng.random() < 0.6:
            sub = img[y0:y1, x0:x1]
            sub[...] = 0.35 * sub + 0.65 * val
        else:
            m = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= s * s
            sub = img[y0:y1, x0:x1]
            sub[m] = 0.35 * sub[m] + 0.65 * val
    # fine speckle: small high-contrast blobs, the bulk of the FAST-20 corners
    n_speck = int(0.012 * w * h)
    sx = rng.integers(2, w - 8, n_speck)
    sy = rng.integers(2, h - 8, n_speck)
    ss = rng.integers(2, 7, n_speck)
    sv = rng.integers(-90, 91, n_speck).astype(np.float32)
    for x, y, q, v in zip(sx, sy, ss, sv):
        img[y:y + q, x:x + q] += v
    n_flat = int(flat_frac * 10)
    for _ in range(n_flat):
        fw, fh = int(rng.integers(w // 16, w // 6)), int(rng.integers(h // 16, h // 6))
        x0, y0 = int(rng.integers(0, w - fw)), int(rng.integers(0, h - fh))
        base = float(rng.integers(60, 200))
        # low-contrast patch: only FAST-7 corners (or none) survive here
        img[y0:y0 + fh, x0:x0 + fw] = base + 0.08 * (img[y0:y0 + fh, x0:x0 + fw] - 128.0)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frame(w: int = 640, h: int = 480, seed: int = 0) -> np.ndarray:
    """Config C1: one synthetic frame."""
    return texture(w, h, seed)


class PlaneStream:
    """Config C2: camera translating 2 cm and yawing 0.2 deg per frame over a textured
    plane at 3 m (plane z = 3 in the first camera's frame), 1280x720 crop of a 2x master
    texture, warped with cv2.warpPerspective(INTER_LINEAR)."""

    def __init__(self, w: int = 1280, h: int = 720, seed: int = 0, K=RPI_K, depth: float = 3.0):
        self.w, self.h, self.K, self.depth = w, h, K, depth
        self.master = texture(2 * w, 2 * h, seed)
        fx, fy, cx, cy = K
        self.Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])

    def pose(self, k: int):
        """World-to-camera (Rcw, tcw) of frame k; world = frame-0 camera frame."""
        yaw = np.deg2rad(0.2) * k
        c, s = np.cos(yaw), np.sin(yaw)
        Rwc = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        twc = np.array([0.02 * k, 0.0, 0.0])
        Rcw = Rwc.T
        return Rcw, -Rcw @ twc

    def plane_to_world(self, u, v):
        """Master-texture pixel (u,v) -> 3-D world point on the plane.  Frame 0 sees the
        master region [w/2, 3w/2) x [h/2, 3h/2) one to one."""
        fx, fy, cx, cy = self.K
        u = np.asarray(u, np.float64)
        v = np.asarray(v, np.float64)
        X = (u - self.w / 2 - cx) * self.depth / fx
        Y = (v - self.h / 2 - cy) * self.depth / fy
        return np.stack([X, Y, np.full_like(X, self.depth)], -1)

    def homography(self, k: int) -> np.ndarray:
        """Maps master pixels to frame-k pixels."""
        Rcw, tcw = self.pose(k)
        # master pixel -> world: Xw = A @ [u, v, 1]
        p00 = self.plane_to_world(0.0, 0.0)
        p10 = self.plane_to_world(1.0, 0.0)
        p01 = self.plane_to_world(0.0, 1.0)
        A = np.stack([p10 - p00, p01 - p00, p00], 1)
        return self.Km @ (Rcw @ A + np.outer(tcw, [0, 0, 1.0]))

    def frame(self, k: int) -> np.ndarray:
        import cv2

        H = self.homography(k)
        return cv2.warpPerspective(self.master, H, (self.w, self.h), flags=cv2.INTER_LINEAR,
                                   borderMode=cv2.BORDER_REFLECT_101)


def backproject_to_plane(stream: "PlaneStream", k: int, xy: np.ndarray) -> np.ndarray:
    """World points (float32 [n,3]) where the pixel rays of frame k hit the textured plane."""
    fx, fy, cx, cy = stream.K
    Rcw, tcw = stream.pose(k)
    Rwc, twc = Rcw.T, -Rcw.T @ tcw
    rays = np.stack([(xy[:, 0] - cx) / fx, (xy[:, 1] - cy) / fy, np.ones(len(xy))], 1) @ Rwc.T
    d = (stream.depth - twc[2]) / rays[:, 2]
    return (twc[None, :] + rays * d[:, None]).astype(np.float32)


def quat_from_R(R: np.ndarray) -> np.ndarray:
    """(x, y, z, w) unit quaternion of a rotation matrix."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def tracking_case(stream: "PlaneStream", k: int, extract, n_local: int = 4000, seed: int = 0,
                  pose_noise=(0.004, 0.002)):
    """Inputs of one tracked frame (SURVEY.md 8d "tracking-only workload"): frame k-1 with a map point
    behind every keypoint (plane geometry is known), frame k, a perturbed pose prior for frame k, and
    a local map of `n_local` plane points observed from frame max(k-8, 0) (descriptor = that first
    observation).  `extract(img) -> (kps, desc, mono)` is either the oracle's or the GPU extractor."""
    rng = np.random.default_rng(seed * 1000 + k)
    last_img, cur_img = stream.frame(k - 1), stream.frame(k)
    lk, ld, _ = extract(last_img)
    ck, cd, _ = extract(cur_img)
    Xw = backproject_to_plane(stream, k - 1, np.stack([lk["x"], lk["y"]], 1).astype(np.float64))
    n_last = len(lk)
    has_mp = (rng.random(n_last) < 0.85).astype(np.uint8)       # not every keypoint holds a map point
    outlier = (rng.random(n_last) < 0.03).astype(np.uint8)      # a few flagged by the previous PoseOptimization
    obs_pos = (rng.random(n_last) < 0.97).astype(np.uint8)      # Observations() > 0 for nearly all
    Rcw, tcw = stream.pose(k)
    # motion-model prior: true pose with a small perturbation
    w = rng.normal(0, pose_noise[1], 3)
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    dR = np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx
    Rp = (dR @ Rcw).astype(np.float32)
    tp = (tcw + rng.normal(0, pose_noise[0], 3)).astype(np.float32)
    # local map: keypoints of an earlier frame lifted to the plane
    k0 = max(k - 8, 0)
    mk, md, _ = extract(stream.frame(k0))
    sel = rng.permutation(len(mk))[:n_local]
    sel.sort()
    Xm = backproject_to_plane(stream, k0, np.stack([mk["x"][sel], mk["y"][sel]], 1).astype(np.float64))
    qp = quat_from_R(Rp.astype(np.float64)).astype(np.float32)   # the prior as an SE3f would hold it
    return dict(last_kps=lk, last_desc=ld, cur_kps=ck, cur_desc=cd, cur_img=cur_img, last_Xw=Xw, has_mp=has_mp,
                outlier=outlier, obs_pos=obs_pos, Rcw_prior=Rp, qcw_prior=qp, tcw_prior=tp, Rcw_true=Rcw, tcw_true=tcw,
                K=np.array(stream.K, np.float32), map_Xw=Xm, map_desc=md[sel].copy(), map_octave=mk["octave"][sel].copy(),
                bounds=(0.0, 0.0, float(stream.w), float(stream.h)))


def ba_scene(n_free: int = 50, n_fixed: int = 10, n_points: int = 5000, seed: int = 0, K=RPI_K, w: int = 1280,
             h: int = 720, max_obs: int = 12, pix_sigma: float = 1.0, outlier_frac: float = 0.05,
             pose_noise=(0.02, np.deg2rad(0.5)), point_noise: float = 0.05):
    """Config C4: cameras on a 12 m circle looking inward, points uniform in a 6 m cube,
    every point observed by up to `max_obs` of the cameras that see it, pixel noise sigma*1.2^octave with
    octave ~ U{0..7}, gross outliers, perturbed initial poses/points.  Returns flat float32 arrays in the
    layout dvm_local_ba takes; cameras [0, n_free) are free, the rest fixed."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    nc = n_free + n_fixed
    ang = np.sort(rng.uniform(0, 2 * np.pi, nc))
    rng.shuffle(ang)
    Rs, ts = [], []
    for a in ang:
        C = np.array([6.0 * np.cos(a), rng.uniform(-0.5, 0.5), 6.0 * np.sin(a)])
        z = -C / np.linalg.norm(C)                      # look at the origin
        x = np.cross(np.array([0.0, 1.0, 0.0]), z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        Rcw = np.stack([x, y, z], 0)
        Rs.append(Rcw)
        ts.append(-Rcw @ C)
    P = rng.uniform(-3.0, 3.0, (n_points, 3))
    e_cam, e_pt, e_obs, e_w = [], [], [], []
    for j in range(n_points):
        vis = []
        for c in range(nc):
            Xc = Rs[c] @ P[j] + ts[c]
            if Xc[2] < 0.5:
                continue
            u, v = fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy
            if 0 <= u < w and 0 <= v < h:
                vis.append((c, u, v))
        if len(vis) > max_obs:
            idx = rng.choice(len(vis), max_obs, replace=False)
            vis = [vis[i] for i in sorted(idx)]
        for c, u, v in vis:
            octv = int(rng.integers(0, 8))
            s = pix_sigma * 1.2 ** octv
            du, dv = rng.normal(0, s, 2)
            if rng.random() < outlier_frac:
                du, dv = rng.uniform(-50, 50, 2)
            e_cam.append(c)
            e_pt.append(j)
            e_obs.append((u + du, v + dv))
            e_w.append(np.float32(1.0) / (np.float32(1.2) ** octv) ** 2)
    # drop points nobody sees (they would make Hll singular, which the reference never builds)
    e_pt = np.array(e_pt, np.int32)
    seen = np.unique(e_pt)
    remap = -np.ones(n_points, np.int64)
    remap[seen] = np.arange(len(seen))
    e_pt = remap[e_pt].astype(np.int32)
    P = P[seen]
    q = np.zeros((nc, 4), np.float32)
    t = np.zeros((nc, 3), np.float32)
    q_true = np.zeros((nc, 4))
    t_true = np.zeros((nc, 3))
    for c in range(nc):
        q_true[c], t_true[c] = quat_from_R(Rs[c]), ts[c]
        R, tt = Rs[c], ts[c]
        if c < n_free:
            wv = rng.normal(0, 1, 3)
            wv *= pose_noise[1] / np.linalg.norm(wv)
            th = np.linalg.norm(wv)
            Kx = np.array([[0, -wv[2], wv[1]], [wv[2], 0, -wv[0]], [-wv[1], wv[0], 0]])
            dR = np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx
            R = dR @ R
            tt = tt + rng.normal(0, pose_noise[0] / np.sqrt(3), 3)
        q[c], t[c] = quat_from_R(R), tt
    fixed = np.zeros(nc, np.uint8)
    fixed[n_free:] = 1
    pts = (P + rng.normal(0, point_noise / np.sqrt(3), P.shape)).astype(np.float32)
    return dict(cam_q=q, cam_t=t, cam_fixed=fixed, pts=pts, edge_cam=np.array(e_cam, np.int32), edge_pt=e_pt,
                edge_obs=np.array(e_obs, np.float32), edge_w=np.array(e_w, np.float32), K=np.array(K, np.float32),
                q_true=q_true, t_true=t_true, pts_true=P)


def ba_scene_large(n_free: int = 399, n_fixed: int = 1, n_points: int = 20000, seed: int = 0, K=RPI_K, w: int = 1280, h: int = 720,
                   max_obs: int = 12, pix_sigma: float = 1.0, outlier_frac: float = 0.02, pose_noise=(0.02, np.deg2rad(0.5)),
                   point_noise: float = 0.05):
    """The geometry of ba_scene for MERGED maps (config C5: hundreds of keyframes), generated with array operations instead of
    per-point loops (not stream-compatible with ba_scene's seeds).  Same layout of the returned arrays; edges grouped by point."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    nc = n_free + n_fixed
    ang = rng.permutation(np.sort(rng.uniform(0, 2 * np.pi, nc)))
    C = np.stack([6.0 * np.cos(ang), rng.uniform(-0.5, 0.5, nc), 6.0 * np.sin(ang)], 1)
    z = -C / np.linalg.norm(C, axis=1, keepdims=True)
    x = np.cross(np.array([0.0, 1.0, 0.0]), z)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y = np.cross(z, x)
    Rs = np.stack([x, y, z], 1)                     # [nc, 3, 3] rows = camera axes
    ts = -np.einsum("cij,cj->ci", Rs, C)
    P = rng.uniform(-3.0, 3.0, (n_points, 3))
    e_cam, e_pt, e_u, e_v = [], [], [], []
    for lo in range(0, n_points, 2048):             # chunks bound the [points, cameras] temporaries
        Pc = P[lo:lo + 2048]
        Xc = np.einsum("cij,pj->pci", Rs, Pc) + ts[None]
        u = fx * Xc[..., 0] / Xc[..., 2] + cx
        v = fy * Xc[..., 1] / Xc[..., 2] + cy
        vis = (Xc[..., 2] >= 0.5) & (u >= 0) & (u < w) & (v >= 0) & (v < h)
        key = np.where(vis, rng.random(vis.shape), 2.0)       # keep at most max_obs random visible cameras per point
        kth = np.sort(key, axis=1)[:, min(max_obs, nc) - 1][:, None]
        keep = vis & (key <= kth)
        pi, ci = np.nonzero(keep)
        e_pt.append(pi + lo); e_cam.append(ci); e_u.append(u[pi, ci]); e_v.append(v[pi, ci])
    e_pt, e_cam = np.concatenate(e_pt), np.concatenate(e_cam).astype(np.int32)
    ne = len(e_pt)
    octv = rng.integers(0, 8, ne)
    sig = pix_sigma * 1.2 ** octv
    noise = rng.normal(0, 1, (ne, 2)) * sig[:, None]
    out = rng.random(ne) < outlier_frac
    noise[out] = rng.uniform(-50, 50, (int(out.sum()), 2))
    obs = (np.stack([np.concatenate(e_u), np.concatenate(e_v)], 1) + noise).astype(np.float32)
    e_w = (np.float32(1.0) / (np.float32(1.2) ** octv.astype(np.float32)) ** 2).astype(np.float32)
    seen = np.unique(e_pt)
    remap = -np.ones(n_points, np.int64)
    remap[seen] = np.arange(len(seen))
    e_pt = remap[e_pt].astype(np.int32)
    P = P[seen]
    q = np.zeros((nc, 4), np.float32)
    t = np.zeros((nc, 3), np.float32)
    for c in range(nc):
        R, tt = Rs[c], ts[c]
        if c < n_free:
            wv = rng.normal(0, 1, 3)
            wv *= pose_noise[1] / np.linalg.norm(wv)
            th = np.linalg.norm(wv)
            Kx = np.array([[0, -wv[2], wv[1]], [wv[2], 0, -wv[0]], [-wv[1], wv[0], 0]])
            R = (np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx) @ R
            tt = tt + rng.normal(0, pose_noise[0] / np.sqrt(3), 3)
        q[c], t[c] = quat_from_R(R), tt
    fixed = np.zeros(nc, np.uint8)
    fixed[n_free:] = 1
    pts = (P + rng.normal(0, point_noise / np.sqrt(3), P.shape)).astype(np.float32)
    return dict(cam_q=q, cam_t=t, cam_fixed=fixed, pts=pts, edge_cam=e_cam, edge_pt=e_pt, edge_obs=obs.reshape(-1), edge_w=e_w,
                K=np.array(K, np.float32), pts_true=P)


def sim3_mul(a, b):
    """Sim3 product on (q xyzw, t, s) rows, g2o::Sim3::operator*."""
    def qmul(p, q):
        x1, y1, z1, w1 = p
        x2, y2, z2, w2 = q
        return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                         w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])

    def rot(q, v):
        uv = 2 * np.cross(q[:3], v)
        return v + q[3] * uv + np.cross(q[:3], uv)
    return np.concatenate([qmul(a[:4], b[:4]), a[7] * rot(a[:4], b[4:7]) + a[4:7], [a[7] * b[7]]])


def sim3_inv(a):
    qc = np.array([-a[0], -a[1], -a[2], a[3]])
    v = (-1.0 / a[7]) * a[4:7]
    uv = 2 * np.cross(qc[:3], v)
    return np.concatenate([qc, v + qc[3] * uv + np.cross(qc[:3], uv), [1.0 / a[7]]])


def _sim3_from(rotvec, t, log_s):
    th = np.linalg.norm(rotvec)
    q = np.array([0.0, 0.0, 0.0, 1.0]) if th < 1e-12 else np.concatenate([np.sin(th / 2) * np.asarray(rotvec) / th, [np.cos(th / 2)]])
    return np.concatenate([q, np.asarray(t, np.float64), [np.exp(log_s)]])


def loop_pose_graph(n: int = 400, seed: int = 0, drift=(0.002, 0.005, 0.0015), covis_edges: int = 2):
    """Config C5's essential graph: `n` keyframes of a closed trajectory (8 agents x 50 keyframes after the merge), estimates
    with accumulating similarity drift, spanning-tree edges taken from the drifted estimates, `covis_edges` extra edges
    per keyframe to earlier neighbours (covisibility >= 100) and one loop edge from the truth.
    -> (sim3 [n, 8] Scw estimates, fixed [n], vi, vj, meas [ne, 8] = Sji, truth [n, 8])."""
    rng = np.random.default_rng(seed)
    true = [_sim3_from([0, 2 * np.pi * k / n, 0], [3 * np.cos(2 * np.pi * k / n), 0.1 * np.sin(6 * np.pi * k / n), 3 * np.sin(2 * np.pi * k / n)], 0.0)
            for k in range(n)]
    est = [true[0].copy()]
    D = _sim3_from([0, 0, 0], [0, 0, 0], 0.0)
    for k in range(1, n):
        D = sim3_mul(_sim3_from(rng.normal(0, drift[0], 3), rng.normal(0, drift[1], 3), rng.normal(0, drift[2])), D)
        est.append(sim3_mul(D, true[k]))
    vi, vj, meas = [], [], []
    for k in range(1, n):                       # vertex(0) = child k, vertex(1) = parent k - 1, Sji = Sjw * Swi
        vi.append(k); vj.append(k - 1); meas.append(sim3_mul(est[k - 1], sim3_inv(est[k])))
        for d in range(2, 2 + covis_edges):
            if k - d >= 0:
                vi.append(k); vj.append(k - d); meas.append(sim3_mul(est[k - d], sim3_inv(est[k])))
    vi.append(n - 1); vj.append(0); meas.append(sim3_mul(true[0], sim3_inv(true[n - 1])))   # the loop closure
    fixed = np.zeros(n, np.uint8)
    fixed[0] = 1
    return np.array(est), fixed, np.array(vi, np.int32), np.array(vj, np.int32), np.array(meas), np.array(true)


class OrbitStream(PlaneStream):
    """Bounded variant of the C2 stream for long runs: the camera sways over the same textured plane
    (x = A sin, yaw = B sin, period `period` frames) so that frame `period` equals frame 0 and the
    resident frame set of bench.py can be cycled.  Peak per-frame motion is 2 cm / 0.2 deg like C2."""

    def __init__(self, w: int = 1280, h: int = 720, seed: int = 0, K=RPI_K, depth: float = 3.0, period: int = 320):
        super().__init__(w, h, seed, K, depth)
        self.period = period
        self.amp_t = 0.02 * period / (2 * np.pi)
        self.amp_yaw = np.deg2rad(0.2) * period / (2 * np.pi)

    def pose(self, k: int):
        ph = 2 * np.pi * k / self.period
        yaw = self.amp_yaw * np.sin(ph + 1.0)
        c, s = np.cos(yaw), np.sin(yaw)
        Rwc = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        twc = np.array([self.amp_t * np.sin(ph), 0.05 * np.sin(2 * ph), 0.1 * np.sin(ph + 0.5)])
        Rcw = Rwc.T
        return Rcw, -Rcw @ twc


def plane_map(stream: "PlaneStream", extract, keyframes, scale_factors, max_points: int = 6000, seed: int = 0):
    """Ground-truth local map for the tracking-only workload: keypoints of the given keyframes lifted to
    the plane; descriptor = first observation; normal / distance limits as MapPoint::UpdateNormalAndDepth
    computes them (O3/src/MapPoint.cc:469-520) for a single observation."""
    rng = np.random.default_rng(seed)
    Xs, Ds, Ns, mind, maxd = [], [], [], [], []
    nlev = len(scale_factors)
    for k in keyframes:
        kps, desc, _ = extract(stream.frame(k))
        X = backproject_to_plane(stream, k, np.stack([kps["x"], kps["y"]], 1).astype(np.float64))
        Rcw, tcw = stream.pose(k)
        Ow = (-Rcw.T @ tcw).astype(np.float32)
        PC = X - Ow[None, :]
        dist = np.linalg.norm(PC, axis=1).astype(np.float32)
        mx = dist * scale_factors[kps["octave"]]
        Xs.append(X); Ds.append(desc); Ns.append(PC / dist[:, None])
        maxd.append(mx); mind.append(mx / scale_factors[nlev - 1])
    X, D, N = np.concatenate(Xs), np.concatenate(Ds), np.concatenate(Ns)
    mind, maxd = np.concatenate(mind), np.concatenate(maxd)
    if len(X) > max_points:
        sel = np.sort(rng.permutation(len(X))[:max_points])
        X, D, N, mind, maxd = X[sel], D[sel], N[sel], mind[sel], maxd[sel]
    return dict(xw=X.astype(np.float32), desc=np.ascontiguousarray(D), normal=N.astype(np.float32),
                min_dist=mind.astype(np.float32), max_dist=maxd.astype(np.float32))


# ------------------------------------------------------------------ inputs of the descriptor matchers
def toy_feature_vector(desc: np.ndarray, bits: int = 7) -> dict:
    """A stand-in for DBoW2's FeatureVector (node id at level L-4 -> feature indices in index order): the
    node of a descriptor is the majority bit of `bits` disjoint groups of its 256 bits, so that noisy
    copies of a descriptor mostly share a node, as words under a common vocabulary ancestor do.  The
    matchers take the feature vector as an input; only its structure matters for parity."""
    d = np.unpackbits(np.ascontiguousarray(desc, np.uint8), axis=1, bitorder="little")
    g = 256 // bits
    node = np.zeros(len(d), np.int64)
    for b in range(bits):
        node |= (d[:, b * g:(b + 1) * g].sum(1) * 2 > g).astype(np.int64) << b
    fv: dict = {}
    for i, n in enumerate(node):
        fv.setdefault(int(n) * 3 + 11, []).append(i)   # sparse, non-contiguous node ids like a real vocabulary
    return fv


def noisy_copy(desc: np.ndarray, flip_p: float, rng) -> np.ndarray:
    """Bernoulli(flip_p) bit flips of 32-byte descriptors."""
    flips = np.packbits(rng.random((len(desc), 256)) < flip_p, axis=1, bitorder="little")
    return np.ascontiguousarray(desc ^ flips)


def keyframe_blocks(n_kf: int, n_feat: int, seed: int, shared_from: np.ndarray | None = None, shared_frac: float = 0.3,
                    flip_p: float = 0.08) -> np.ndarray:
    """Config C3 input: `n_kf` keyframes x `n_feat` descriptors u8[n_kf, n_feat, 32].  With
    `shared_from` (another agent's blocks of the same shape) the first `shared_frac` of every keyframe's
    descriptors are noisy copies of that agent's keyframe with the same index (true matches, expected
    distance 256 * flip_p), placed at shuffled positions; the rest is uniform random (expected 128)."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, (n_kf, n_feat, 32), dtype=np.uint8)
    if shared_from is not None:
        ns = int(shared_frac * n_feat)
        for k in range(n_kf):
            src = rng.choice(n_feat, ns, replace=False)
            dst = rng.choice(n_feat, ns, replace=False)
            out[k, dst] = noisy_copy(shared_from[k, src], flip_p, rng)
    return out


def toy_vocabulary(k: int = 10, L: int = 3, seed: int = 0, ragged: bool = False, stop_frac: float = 0.05):
    """A synthetic DBoW2 vocabulary tree in loadFromTextFile order (breadth first, parents before children):
    -> dict(k, L, scoring, weighting, parent, is_leaf, desc, weight).  Child descriptors are noisy copies of their
    parent's so that descents are meaningful; with `ragged` some inner nodes have fewer children and some branches
    end in a leaf above level L; a few words are stopped (weight 0)."""
    rng = np.random.default_rng(seed)
    parent, is_leaf, desc, weight = [], [], [], []
    frontier = [(0, rng.integers(0, 256, 32, dtype=np.uint8), 0)]   # (node id, descriptor, level)
    next_id = 1
    while frontier:
        new = []
        for pid, pdesc, lvl in frontier:
            nchild = k if not ragged else int(rng.integers(2, k + 1))
            for _ in range(nchild):
                d = noisy_copy(pdesc[None, :], 0.25 / (lvl + 1), rng)[0]
                leaf = lvl + 1 == L or (ragged and lvl + 1 >= 2 and rng.random() < 0.15)
                parent.append(pid); is_leaf.append(int(leaf)); desc.append(d)
                weight.append(0.0 if (leaf and rng.random() < stop_frac) else (float(rng.uniform(0.5, 9.0)) if leaf else 0.0))
                if not leaf:
                    new.append((next_id, d, lvl + 1))
                next_id += 1
        frontier = new
    return dict(k=k, L=L, scoring=0, weighting=0, parent=np.array(parent, np.int64), is_leaf=np.array(is_leaf, np.int64),
                desc=np.array(desc, np.uint8), weight=np.array(weight, np.float64))


def write_vocabulary_text(path: str, v: dict) -> None:
    """ORBvoc.txt format (DBoW2 saveToTextFile)."""
    with open(path, "w") as f:
        f.write(f"{v['k']} {v['L']}  {v['scoring']} {v['weighting']}\n")
        for p, leaf, d, w in zip(v["parent"], v["is_leaf"], v["desc"], v["weight"]):
            f.write(f"{int(p)} {int(leaf)} " + " ".join(str(int(x)) for x in d) + f" {float(w)!r}\n")


def sim3_scene(n: int = 200, seed: int = 0, K=RPI_K, w: int = 1280, h: int = 720, scale: float = 1.3, pix_sigma: float = 1.0,
               outlier_frac: float = 0.1, perturb=(np.deg2rad(1.0), 0.05, 0.03), not_in_kf2_frac: float = 0.2):
    """Correspondences for Optimizer::OptimizeSim3: two keyframes of two maps whose frames differ by a similarity.  Point
    i has camera-frame positions P1c (map 1, keyframe 1) and P2c (map 2, keyframe 2) with P1c = S12 * P2c, keypoints
    obs1 / obs2 with octave noise and gross outliers; a fraction of the points has no keypoint in keyframe 2 (the
    reference then uses the normalised projection of P2c as measurement and octave 0, Optimizer.cc:2118-2125).  Returns
    float32 arrays in dvm_optimize_sim3's layout, the true S12 and a perturbed initial guess (double)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    # true S12: moderate rotation / translation, the given scale
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = np.deg2rad(rng.uniform(5, 25))
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R12 = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    t12 = rng.uniform(-0.5, 0.5, 3)
    p2, p1, o1, o2, w1, w2 = [], [], [], [], [], []
    while len(p2) < n:
        X2 = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(2, 10)])
        X1 = scale * (R12 @ X2) + t12
        if X1[2] < 0.5:
            continue
        u1, v1 = fx * X1[0] / X1[2] + cx, fy * X1[1] / X1[2] + cy
        u2, v2 = fx * X2[0] / X2[2] + cx, fy * X2[1] / X2[2] + cy
        if not (0 <= u1 < w and 0 <= v1 < h and 0 <= u2 < w and 0 <= v2 < h):
            continue
        oc1, oc2 = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        d1 = rng.normal(0, pix_sigma * 1.2 ** oc1, 2)
        d2 = rng.normal(0, pix_sigma * 1.2 ** oc2, 2)
        if rng.random() < outlier_frac:
            d1 = rng.uniform(-60, 60, 2)
        if rng.random() < not_in_kf2_frac:
            oc2 = 0
            ob2 = (np.float32(X2[0]) * (np.float32(1) / np.float32(X2[2])), np.float32(X2[1]) * (np.float32(1) / np.float32(X2[2])))
        else:
            ob2 = (u2 + d2[0], v2 + d2[1])
        p2.append(X2); p1.append(X1)
        o1.append((u1 + d1[0], v1 + d1[1])); o2.append(ob2)
        w1.append(np.float32(1.0) / (np.float32(1.2) ** oc1) ** 2)
        w2.append(np.float32(1.0) / (np.float32(1.2) ** oc2) ** 2)
    q_true = quat_from_R(R12)
    # initial guess: exp(noise) * S12
    dax = rng.normal(size=3); dax /= np.linalg.norm(dax)
    da = perturb[0]
    Kd = np.array([[0, -dax[2], dax[1]], [dax[2], 0, -dax[0]], [-dax[1], dax[0], 0]])
    Rd = np.eye(3) + np.sin(da) * Kd + (1 - np.cos(da)) * Kd @ Kd
    R0 = Rd @ R12
    t0 = t12 + rng.normal(0, perturb[1], 3)
    s0 = scale * (1 + perturb[2])
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(p1c=f32(p1), p2c=f32(p2), obs1=f32(o1), obs2=f32(o2), w1=f32(w1), w2=f32(w2), K=f32(K),
                q_true=q_true, t_true=t12, s_true=scale, q0=quat_from_R(R0).astype(np.float64), t0=t0.astype(np.float64), s0=float(s0))
