"""Seeded Sim3 pose-graph cases shared by tests/test_sim3_oracle.py and tests/golden/make_golden.py (numpy only)."""
import numpy as np


def sim3_mul(a, b):
    """Sim3 product on (q xyzw, t, s) rows, g2o::Sim3::operator*."""
    def qmul(p, q):
        x1, y1, z1, w1 = p; x2, y2, z2, w2 = q
        return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                         w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    def rot(q, v):
        qv = q[:3]; uv = 2 * np.cross(qv, v)
        return v + q[3] * uv + np.cross(qv, uv)
    return np.concatenate([qmul(a[:4], b[:4]), a[7] * rot(a[:4], b[4:7]) + a[4:7], [a[7] * b[7]]])


def sim3_inv(a):
    qc = np.array([-a[0], -a[1], -a[2], a[3]])
    qv = qc[:3]; v = (-1.0 / a[7]) * a[4:7]; uv = 2 * np.cross(qv, v)
    return np.concatenate([qc, v + qc[3] * uv + np.cross(qv, uv), [1.0 / a[7]]])


def loop_graph(n=24, seed=0, drift=(0.004, 0.01, 0.003)):
    """A closed trajectory: true poses on a circle, estimates with accumulating similarity drift, odometry edges taken from
    the drifted estimates (zero error at the start), one loop edge from the truth."""
    from oracle.sim3 import sim3_exp
    rng = np.random.default_rng(seed)
    true = []
    for k in range(n):
        a = 2 * np.pi * k / n
        true.append(sim3_exp([0, a, 0, 3 * np.cos(a), 0.1 * np.sin(3 * a), 3 * np.sin(a), 0]))
    est = [true[0].copy()]
    D = sim3_exp(np.zeros(7))
    for k in range(1, n):
        step = np.concatenate([rng.normal(0, drift[0], 3), rng.normal(0, drift[1], 3), [rng.normal(0, drift[2])]])
        D = sim3_mul(sim3_exp(step), D)
        est.append(sim3_mul(D, true[k]))
    vi, vj, meas = [], [], []
    for k in range(1, n):                      # spanning-tree edge: vertex 0 = child k, vertex 1 = parent k-1, Sji = Sjw * Swi
        vi.append(k); vj.append(k - 1); meas.append(sim3_mul(est[k - 1], sim3_inv(est[k])))
    vi.append(n - 1); vj.append(0); meas.append(sim3_mul(true[0], sim3_inv(true[n - 1])))   # the loop closure
    fixed = np.zeros(n, np.uint8); fixed[0] = 1
    return np.array(est), fixed, np.array(vi, np.int32), np.array(vj, np.int32), np.array(meas), np.array(true)
