"""Deterministic inputs of the GMS cases whose outputs were recorded from the reference's own machine code
(tests/golden/gms_dll.npz, written by tests/golden/make_gms_dll_golden.py from
/root/reference/SfM-GMS/bin/opencv_xfeatures2d452.dll, hosted by oracle/dllref/).

The .npz holds only what the DLL returned; this module rebuilds the inputs (from the committed cv2 fixtures,
from seeded generators and from hand-made edge sets), so the CPU tests (oracle == DLL) and the GPU tests
(CUDA path == DLL) run where the DLL is absent.  `bf(desc1, desc2) -> train_idx` is supplied by the caller:
the generator and the CPU tests use the cv2-pinned oracle, the GPU tests use the CUDA matcher.
"""
import hashlib

import numpy as np

from conftest import load_golden

FLAGS = [("00", 0, 0), ("10", 1, 0), ("01", 0, 1), ("11", 1, 1)]     # tag, withRotation, withScale
REAL = ["pikabun12", "disparityLR", "view01_2k", "bun12_rot180_3k"]
FACTORS = [0.0, 0.5, 1.0, 3.0, 5.999999, 6.0000001, 7.25, 12.0, 50.0]
N_STRESS = 240
GRID_SEED, GRID_N = 4242, 1 << 20
RIGHT_GRIDS = [20, 10, 14, 28, 40]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _inside(xy, w, h):
    xy = np.asarray(xy, np.float32).copy()
    xy[:, 0] = np.clip(xy[:, 0], 0, np.nextafter(np.float32(w), np.float32(0)))
    xy[:, 1] = np.clip(xy[:, 1], 0, np.nextafter(np.float32(h), np.float32(0)))
    return xy


# ---------------------------------------------------------------------------------------------- grid-index inputs
def edge_values():
    """every f32 within +-3 ulp of a cell edge k/W (all right grids and the left grid) and of a half-cell edge
    (k+0.5)/20 of the shifted left grids, inside [0, 1)"""
    vals = []
    for wgrid in sorted(set(RIGHT_GRIDS)):
        for k2 in range(0, 2 * wgrid + 1):           # k2/2 = k or k + 0.5
            c = np.float32(k2 / (2.0 * wgrid))
            v = c
            for _ in range(3):
                v = np.nextafter(v, np.float32(-1))
            for _ in range(7):
                vals.append(v)
                v = np.nextafter(v, np.float32(2))
    v = np.unique(np.array(vals, np.float32))
    return v[(v >= 0) & (v < 1)]


def grid_edge_points():
    """normalised points: each edge value as x against a few y, and as y against a few x"""
    e = edge_values()
    others = np.array([0.0, 0.26, 0.5, 0.974999, 0.99999994], np.float32)
    a = np.stack(np.meshgrid(e, others, indexing="ij"), -1).reshape(-1, 2)
    b = a[:, ::-1]
    c = np.stack([e, e[::-1]], 1)
    return np.ascontiguousarray(np.concatenate([a, b, c]), np.float32)


def grid_random_points():
    rng = np.random.default_rng(GRID_SEED)
    return rng.random((GRID_N, 2), dtype=np.float32)


# ---------------------------------------------------------------------------------------------- full-run cases
def real_case(name):
    g = load_golden(name)
    n = len(g["bf_train"])
    return dict(size1=tuple(int(v) for v in g["size1"]), size2=tuple(int(v) for v in g["size2"]), kp1=g["kp1"],
                kp2=g["kp2"], q=np.arange(n, dtype=np.int32), t=g["bf_train"].astype(np.int32))


def sift_case():
    g = load_golden("sift_view01_1500")
    n = len(g["l2_train"])
    return dict(size1=tuple(int(v) for v in g["size1"]), size2=tuple(int(v) for v in g["size2"]), kp1=g["kp1"],
                kp2=g["kp2"], q=np.arange(n, dtype=np.int32), t=g["l2_train"].astype(np.int32))


def synth_case(cfg, bf):
    """BASELINE configs 2 / 3 at full size (sfm_gms_b200.synth); matches from `bf`"""
    from sfm_gms_b200 import synth

    d = synth.make_config(cfg)
    t = np.asarray(bf(d["desc1"], d["desc2"]), np.int32)
    return dict(size1=d["size1"], size2=d["size2"], kp1=d["kp1"], kp2=d["kp2"], q=np.arange(len(t), dtype=np.int32), t=t)


def edge_pixels_case():
    """keypoints whose normalised coordinate lands on / next to a cell edge or half-cell edge, a few per cell and 45 % outliers so
    that cells sit around the threshold: the mask then depends on which side of the edge the DLL puts each point.
    200 x 200 image (10-px cells) and a 2nd image of 140 x 280 (right cells of 7 x 14 px at scale 0)."""
    rng = np.random.default_rng(77)
    pts = []
    for k in range(0, 40):
        c = np.float32(k * 5.0)                       # every edge and half edge of the 20-grid on a 200-px axis
        v = c
        for _ in range(2):
            v = np.nextafter(v, np.float32(-1))
        for _ in range(5):
            if 0 <= v < 200:
                pts.append(v)
            v = np.nextafter(v, np.float32(1000))
    e = np.array(pts, np.float32)
    reps = 3
    x = np.repeat(e, reps)
    y = (rng.random(len(x)) * 199).astype(np.float32)
    kp1 = np.concatenate([np.stack([x, y], 1), np.stack([y, x], 1), np.stack([x, x[::-1]], 1)]).astype(np.float32)
    kp1 = _inside(kp1, 200, 200)
    kp2 = _inside(kp1 * np.array([0.7, 1.4], np.float32), 140, 280)
    n = len(kp1)
    bad = rng.random(n) < 0.45                        # outliers: cell populations end up around the threshold
    kp2[bad] = _inside(np.stack([rng.random(n) * 140, rng.random(n) * 280], 1), 140, 280)[bad]
    return dict(size1=(200, 200), size2=(140, 280), kp1=kp1, kp2=kp2, q=np.arange(n, dtype=np.int32),
                t=np.arange(n, dtype=np.int32))


def subset_case():
    """an arbitrary match list over pikabun12: a subset, out of order, with repeated query and train indices"""
    c = real_case("pikabun12")
    rng = np.random.default_rng(31)
    pick = rng.integers(0, len(c["q"]), 7000)
    c["q"], c["t"] = c["q"][pick].copy(), c["t"][pick].copy()
    c["t"][::97] = rng.integers(0, len(c["kp2"]), len(c["t"][::97]))
    return c


def stress_case(i):
    """small random pairs: random image sizes, a similarity-warped inlier subset (one of the 8 x 5 rotation / scale
    hypotheses), arbitrary match lists (repeats, any order), 0 ... 2500 matches — cell populations sit around the
    6*sqrt(mean) threshold, so individual f64 threshold decisions matter"""
    rng = np.random.default_rng(900_000 + i)
    w1, h1, w2, h2 = (int(v) for v in rng.integers(40, 3000, 4))
    n1, n2 = int(rng.integers(1, 1800)), int(rng.integers(1, 1800))
    kp1 = _inside(np.stack([rng.random(n1) * w1, rng.random(n1) * h1], 1), w1, h1)
    kp2 = _inside(np.stack([rng.random(n2) * w2, rng.random(n2) * h2], 1), w2, h2)
    m = min(n1, n2)
    rot_k = int(rng.integers(0, 8))
    scale = [1.0, 0.5, 2 ** -0.5, 2 ** 0.5, 2.0][int(rng.integers(0, 5))]
    th = np.pi / 4 * rot_k
    p = kp1[:m].astype(np.float64) / [w1, h1] - 0.5
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    p = ((p @ R.T) * min(scale, 1.0) + 0.5 + rng.normal(0, 0.004, (m, 2))) * [w2, h2]
    ok = (p[:, 0] >= 0) & (p[:, 0] < w2) & (p[:, 1] >= 0) & (p[:, 1] < h2) & (rng.random(m) < rng.uniform(0.2, 1.0))
    kp2[:m][ok] = _inside(p[ok], w2, h2)
    n = int(rng.integers(0, 2500)) if i % 17 else 0
    q = rng.integers(0, n1, n).astype(np.int32)
    t = np.where((q < m) & (rng.random(n) < 0.8), q, rng.integers(0, n2, n)).astype(np.int32)
    if i % 5 == 0 and n:                                   # sometimes a plain one-match-per-query list
        q = np.arange(min(n, n1), dtype=np.int32)
        t = np.where(q < m, q, rng.integers(0, n2, len(q))).astype(np.int32)
    factor = 6.0 if i % 3 else float(rng.uniform(0.5, 9.0))
    return dict(size1=(w1, h1), size2=(w2, h2), kp1=kp1, kp2=kp2, q=q, t=t, factor=factor)


def micro_inputs():
    """the inputs of tests/microcases.py, captured by running each case with a recording gms function"""
    from microcases import MICROCASES

    out = []

    class Stop(Exception):
        pass

    for name in sorted(MICROCASES):
        calls = []

        def rec(s1, s2, kp1, kp2, q, t, rot, sc, factor):
            calls.append(dict(size1=tuple(s1), size2=tuple(s2), kp1=np.asarray(kp1, np.float32), kp2=np.asarray(kp2, np.float32),
                              q=np.asarray(q, np.int32), t=np.asarray(t, np.int32), rot=int(bool(rot)), sc=int(bool(sc)),
                              factor=float(factor)))
            raise Stop()

        # every call of the case, one at a time: re-run the case, letting the first k calls through with the oracle
        import oracle
        k = 0
        while True:
            calls.clear()
            seen = [0]

            def gate(s1, s2, kp1, kp2, q, t, rot, sc, factor):
                if seen[0] == k:
                    return rec(s1, s2, kp1, kp2, q, t, rot, sc, factor)
                seen[0] += 1
                return oracle.gms(s1, s2, kp1, kp2, q, t, rot, sc, factor)

            try:
                MICROCASES[name](gate)
                break
            except Stop:
                out.append(("%s#%d" % (name, k), calls[0]))
                k += 1
    return out
