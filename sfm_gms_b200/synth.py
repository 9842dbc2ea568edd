"""Synthetic ORB-like inputs of the shapes BASELINE.json names, numpy only.

Used by tests/ and bench.py (there is no network for real datasets).  Deterministic in the seed.
"""
import numpy as np

CONFIGS = {
    # name: (width, height, n_keypoints, with_rotation, with_scale, seed)
    "cfg2_640x480_10k": (640, 480, 10_000, False, False, 2),
    "cfg3_1080p_50k_rs": (1920, 1080, 50_000, True, True, 3),
    "cfg4_4k_200k": (3840, 2160, 200_000, False, False, 4),
}


def _inside(xy, w, h):
    """Clip to the supported GMS domain 0 <= x < w, 0 <= y < h in float32."""
    xy = xy.astype(np.float32)
    xy[:, 0] = np.clip(xy[:, 0], 0, np.nextafter(np.float32(w), np.float32(0)))
    xy[:, 1] = np.clip(xy[:, 1], 0, np.nextafter(np.float32(h), np.float32(0)))
    return xy


def random_keypoints(rng, n, w, h):
    xy = np.stack([rng.random(n, dtype=np.float32) * np.float32(w), rng.random(n, dtype=np.float32) * np.float32(h)], 1)
    return _inside(xy, w, h)


def flip_bits(rng, desc, p):
    """Flip each of the 256 bits independently with probability p (~Binomial(256, p) flips per row)."""
    n = desc.shape[0]
    flips = np.packbits(rng.random((n, 256), dtype=np.float32) < p, axis=1)
    return desc ^ flips


def make_pair(w, h, n, seed, inlier_frac=0.5, rot_k=0, scale=1.0, shift_frac=0.03, dup_frac=0.01, flip_p=0.08):
    """One synthetic image pair.

    Image 2's inliers are image-1 keypoints warped by a similarity (rotation 90deg*rot_k about the image
    centre, isotropic scale, translation shift_frac*size) + N(0,1px) noise, with noisy copies of their
    descriptors; the rest is random.  dup_frac of the train rows are exact duplicates of other train rows
    (tie-break exercise).  Returns dict(size1,size2,kp1,kp2,desc1,desc2).
    """
    rng = np.random.default_rng(seed)
    kp1 = random_keypoints(rng, n, w, h)
    desc1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    n_in = int(n * inlier_frac)
    src = rng.permutation(n)[:n_in]
    c = np.array([w / 2.0, h / 2.0])
    p = kp1[src].astype(np.float64) - c
    th = np.pi / 2 * rot_k
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    p = (p @ R.T) * scale + c + np.array([w, h]) * shift_frac + rng.normal(0, 1.0, (n_in, 2))
    kp2 = random_keypoints(rng, n, w, h)
    desc2 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ok = (p[:, 0] >= 0) & (p[:, 0] < w - 1) & (p[:, 1] >= 0) & (p[:, 1] < h - 1)
    slots = rng.permutation(n)[:n_in]
    kp2[slots[ok]] = _inside(p[ok], w, h)
    desc2[slots[ok]] = flip_bits(rng, desc1[src[ok]], flip_p)
    n_dup = int(n * dup_frac)
    if n_dup:
        a = rng.integers(0, n, n_dup)
        b = rng.integers(0, n, n_dup)
        desc2[a] = desc2[b]
    return dict(size1=(w, h), size2=(w, h), kp1=kp1, kp2=kp2, desc1=desc1, desc2=desc2)


def make_config(name, n_override=None):
    w, h, n, rot, sc, seed = CONFIGS[name]
    if n_override:
        n = n_override
    if name.startswith("cfg3"):
        # 180deg rotation + half scale about the centre: a non-trivial hypothesis (scale idx 1, rotation
        # type 5) must win the 40-way search
        d = make_pair(w, h, n, seed, rot_k=2, scale=0.5, shift_frac=0.0)
    else:
        d = make_pair(w, h, n, seed)
    d["with_rotation"], d["with_scale"] = rot, sc
    return d


def make_pair_batch(n_pairs, w=640, h=480, n=10_000, seed0=2):
    """n_pairs distinct config-2-shaped pairs as one image set: images (2p, 2p+1) form pair p."""
    descs, kps = [], []
    for p in range(n_pairs):
        d = make_pair(w, h, n, seed0 + p)
        descs += [d["desc1"], d["desc2"]]
        kps += [d["kp1"], d["kp2"]]
    offsets = np.arange(2 * n_pairs + 1, dtype=np.int64) * n
    sizes = np.tile(np.array([[w, h]], np.int32), (2 * n_pairs, 1))
    pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(-1, 2)
    return dict(offsets=offsets, desc=np.concatenate(descs), kp=np.concatenate(kps), sizes=sizes, pairs=pairs)


def make_sequence(n_images, n_kp=10_000, w=640, h=480, pool=None, step_frac=0.01, flip_p=0.05, seed0=1000):
    """Config 5: an image sequence over a shared landmark pool.

    A pool of 2*n_kp landmarks (descriptor + position in a wide panorama); image k sees the n_kp landmarks
    of a window sliding by step_frac*pool per image, positions translated accordingly, descriptors with
    per-image bit noise (seed 1000+k).  Returns an image-set dict (offsets, desc, kp, sizes).
    """
    pool = pool or 2 * n_kp
    rng = np.random.default_rng(seed0 - 1)
    pool_desc = rng.integers(0, 256, (pool, 32), dtype=np.uint8)
    pool_x = np.sort(rng.random(pool)) * 2.0 * w            # panorama twice as wide as one frame
    pool_y = rng.random(pool) * h
    step = max(1, int(pool * step_frac))
    descs, kps = [], []
    for k in range(n_images):
        r = np.random.default_rng(seed0 + k)
        start = (k * step) % (pool - n_kp + 1)
        idx = np.arange(start, start + n_kp)
        x = pool_x[idx] - pool_x[start]
        span = max(pool_x[idx[-1]] - pool_x[start], 1e-6)
        xy = np.stack([x / span * (w - 1) + r.normal(0, 0.5, n_kp), pool_y[idx] + r.normal(0, 0.5, n_kp)], 1)
        perm = r.permutation(n_kp)
        kps.append(_inside(xy[perm], w, h))
        descs.append(flip_bits(r, pool_desc[idx][perm], flip_p))
    offsets = np.arange(n_images + 1, dtype=np.int64) * n_kp
    sizes = np.tile(np.array([[w, h]], np.int32), (n_images, 1))
    return dict(offsets=offsets, desc=np.concatenate(descs), kp=np.concatenate(kps), sizes=sizes)


def all_pairs(n_images):
    i, j = np.triu_indices(n_images, 1)
    return np.stack([i, j], 1).astype(np.int32)


def window_pairs(n_images, window):
    out = [(i, j) for i in range(n_images) for j in range(i + 1, min(n_images, i + 1 + window))]
    return np.array(out, np.int32).reshape(-1, 2)
