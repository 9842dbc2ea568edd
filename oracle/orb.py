"""CPU restatement of cv::ORB (compute on provided keypoints, detectAndCompute), SURVEY §8f-4.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Reference call site: DisparityUtil.cpp:107 (`ORB::create()`,
all defaults) and :127-134 (`f2d->compute(img, keypoints, descriptors)` with a KeyPoint at every pixel).  The
arithmetic lives in OpenCV features2d `orb.cpp` (un-vendored dependency, 4.5.x in the reference, 4.13 in this image):

  detectAndCompute(image, noArray(), keypoints, descriptors, useProvidedKeypoints=true)
    1. non-gray input -> cvtColor(BGR2GRAY): (B*3735 + G*19235 + R*9798 + 2^14) >> 15
    2. level 0 of the pyramid = the image with a reflect-101 border (keypoints with octave 0 use no other level)
    3. KeyPointsFilter::runByImageBorder(kp, size, edgeThreshold): keep edge <= cvRound(x) < w-edge, same for y
    4. GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) -- on the pyramid SUB-matrix, which OpenCV does not send
       to its fixed-point 8-bit kernel but to the float separable filter: float32 taps, row pass sequential (FMA in
       the 32-pixel vector loop, two roundings in its scalar remainder), column pass symmetric with FMA, rounded half
       to even (oracle_orb_blur7 in sfmgms_oracle.c, which also lists the evidence).
    5. computeOrbDescriptors: angle (degrees, as given: -1 for a default KeyPoint) -> radians in float,
       a = (float)cos, b = (float)sin; for each of the 512 pattern points  x = px*a - py*b, y = px*b + py*a  in float
       (products and sum rounded separately), sample blurred[cy + cvRound(y), cx + cvRound(x)] around
       (cx, cy) = (cvRound(pt.x), cvRound(pt.y)); bit k of byte i = sample[16i + 2k] < sample[16i + 2k + 1].
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN = np.load(os.path.join(_HERE, "orb_pattern.npy")).astype(np.int32)   # (512, 2): x, y
EDGE_THRESHOLD = 31


def gray_from_bgr(img):
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def gaussian_blur_7(gray):
    """GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) as ORB gets it: see oracle_orb_blur7 in sfmgms_oracle.c."""
    from . import orb_blur7
    return orb_blur7(gray)


def orb_compute(image, pts, angles, octaves=None):
    """-> (kept int32[n], desc uint8[n, 32]).  image: HxW (gray) or HxWx3 (BGR) uint8; pts float32 (N, 2) = (x, y) in
    image coordinates; angles float32 (N,) in degrees; octaves int (N,) pyramid level of each keypoint (None: 0).
    Keypoints not sorted by octave come back regrouped level by level (orb.cpp), `kept` follows that order."""
    image = np.asarray(image)
    gray = gray_from_bgr(image) if image.ndim == 3 else image
    h, w = gray.shape
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    ang = np.asarray(angles, np.float32).reshape(-1)
    octv = np.zeros(len(pts), np.int64) if octaves is None else np.asarray(octaves, np.int64).reshape(-1)
    nlevels = int(octv.max()) + 1 if len(octv) else 1
    sorted_by_level = bool(np.all(np.diff(octv) >= 0))
    cx = np.rint(pts[:, 0]).astype(np.int64)                   # cvRound: half to even
    cy = np.rint(pts[:, 1]).astype(np.int64)
    keep = (cx >= EDGE_THRESHOLD) & (cx < w - EDGE_THRESHOLD) & (cy >= EDGE_THRESHOLD) & (cy < h - EDGE_THRESHOLD)
    kept = np.nonzero(keep)[0].astype(np.int32)
    if not sorted_by_level:
        kept = kept[np.argsort(octv[kept], kind="stable")]
    if len(kept) == 0:
        return kept, np.zeros((0, 32), np.uint8)
    levels = build_pyramid(gray, nlevels)
    blurred = [gaussian_blur_7(img) for img in levels]
    # what OpenCV's pyramid buffer holds around a level: the reflect-101 border of the UNBLURRED level
    padded = [np.pad(b, 32, mode="constant") for b in blurred]
    for l in range(nlevels):
        raw = np.pad(levels[l], 32, mode="reflect")
        raw[32:-32, 32:-32] = blurred[l]
        padded[l] = raw
    f = np.float32
    px = PATTERN[:, 0].astype(f)
    py = PATTERN[:, 1].astype(f)
    desc = np.zeros((len(kept), 32), np.uint8)
    for r, i in enumerate(kept):
        l = int(octv[i])
        inv = f(1) / level_scale(l)
        ccx, ccy = int(np.rint(f(pts[i, 0] * inv))), int(np.rint(f(pts[i, 1] * inv)))
        rad = f(ang[i] * f(np.pi / f(180.0)))                   # float * (float)(CV_PI/180.f)
        a, b = f(np.cos(np.float64(rad))), f(np.sin(np.float64(rad)))
        ix = np.rint((px * a).astype(f) - (py * b).astype(f)).astype(np.int64)
        iy = np.rint((px * b).astype(f) + (py * a).astype(f)).astype(np.int64)
        vals = padded[l][ccy + iy + 32, ccx + ix + 32].astype(np.int32)
        desc[r] = np.packbits((vals[0::2] < vals[1::2]).astype(np.uint8).reshape(32, 8)[:, ::-1], axis=1).ravel()
    return kept, desc


# ---------------------------------------------------------------------------------------------------------
# detectAndCompute (DisparityUtil.cpp:139-140 with ORB::create() defaults; BASELINE config 1 with 10000 / 0).
# Restates orb.cpp computeKeyPoints / HarrisResponses / ICAngles, fast.cpp FAST_t<16> + cornerScore<16>,
# keypoint.cpp runByImageBorder / retainBest and imgproc resize(INTER_LINEAR_EXACT).  Pinned: keypoint list (order,
# pt, size, angle, response, octave) and descriptors identical to cv2 4.13 on the committed fixtures.
# ---------------------------------------------------------------------------------------------------------
import ctypes
import subprocess

SCALE_FACTOR = float(np.float32(1.2))        # ORB::create(scaleFactor = 1.2f), stored as double
N_LEVELS = 8
PATCH = 31
HARRIS_K = np.float32(0.04)
_CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
           (-3, 0), (-3, 1), (-2, 2), (-1, 3)]
_RETAIN = None


def _retain_lib():
    global _RETAIN
    if _RETAIN is None:
        so = os.path.join(_HERE, "libretain_best.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libretain_best.so"])
        _RETAIN = ctypes.CDLL(so)
    return _RETAIN


def retain_best(response, n_points):
    """-> (responses, ids) of the survivors, in the order std::nth_element + std::partition leave them."""
    resp = np.ascontiguousarray(response, np.float32).copy()
    ids = np.arange(len(resp), dtype=np.int32)
    m = _retain_lib().retain_best(resp.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p), len(resp), int(n_points))
    return resp[:m], ids[:m]


def _linear_exact_coeffs(ssize, dsize):
    scale = np.float64(1.0) / (np.float64(dsize) / np.float64(ssize))
    ofs = np.zeros(dsize, np.int64)
    c1 = np.zeros(dsize, np.int64)
    for v in range(dsize):
        f = scale * (np.float64(v) + 0.5) - 0.5
        i = int(np.floor(f))
        if i >= 0 and ssize > 1:
            if i < ssize - 1:
                ofs[v] = i
                c1[v] = int(np.rint((f - i) * 256))
            else:
                ofs[v] = ssize - 1
    return ofs, c1


def resize_linear_exact(src, dw, dh):
    """cv::resize(src, (dw, dh), INTER_LINEAR_EXACT) for 8-bit single channel: 8.8 fixed-point weights."""
    sh, sw = src.shape
    xo, xc = _linear_exact_coeffs(sw, dw)
    yo, yc = _linear_exact_coeffs(sh, dh)
    s = src.astype(np.int64)
    hz = s[:, xo] * (256 - xc) + s[:, np.minimum(xo + 1, sw - 1)] * xc
    v = hz[yo, :] * (256 - yc)[:, None] + hz[np.minimum(yo + 1, sh - 1), :] * yc[:, None]
    return ((v + 32768) >> 16).astype(np.uint8)


def level_scale(level, scale_factor=SCALE_FACTOR):
    """getScale(level, firstLevel = 0, scaleFactor) = (float)pow(scaleFactor, level), scaleFactor held as double"""
    return np.float32(np.float64(scale_factor) ** np.float64(level))


def build_pyramid(gray, nlevels=N_LEVELS, scale_factor=SCALE_FACTOR):
    h, w = gray.shape
    levels, prev = [], gray
    for l in range(nlevels):
        inv = np.float32(1.0) / level_scale(l, scale_factor)
        dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
        cur = gray if l == 0 else resize_linear_exact(prev, dw, dh)
        levels.append(cur)
        prev = cur
    return levels


def fast_detect(img, threshold):
    """FAST-9/16 with non-max suppression -> (xs, ys, scores) in row-major order (fast.cpp)."""
    h, w = img.shape
    if h < 7 or w < 7:
        z = np.zeros(0, np.int64)
        return z, z, z
    I = img.astype(np.int16)
    c = I[3:h - 3, 3:w - 3]
    d = np.stack([c - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in _CIRCLE], 0)
    dd = np.concatenate([d, d[:8]], 0)
    amin = np.full(c.shape, -32768, np.int16)
    amax = np.full(c.shape, 32767, np.int16)
    for s in range(16):                       # the 16 contiguous arcs of 9 pixels
        seg = dd[s:s + 9]
        amin = np.maximum(amin, seg.min(0))
        amax = np.minimum(amax, seg.max(0))
    corner = (amin > threshold) | (amax < -threshold)
    score = np.maximum(np.maximum(amin, threshold), -np.minimum(amax, -threshold)) - 1     # cornerScore<16>
    smap = np.zeros((h, w), np.int32)
    smap[3:h - 3, 3:w - 3] = np.where(corner, score, 0)
    p = np.pad(smap, 1)
    nb = np.stack([p[1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)], 0)
    ys, xs = np.nonzero((smap > 0) & (smap > nb.max(0)))
    return xs, ys, smap[ys, xs]


def features_per_level(nfeatures, nlevels=N_LEVELS, scale_factor=SCALE_FACTOR):
    f = np.float32
    factor = f(1.0 / np.float64(scale_factor))
    nd = f(f(nfeatures) * (f(1) - factor)) / (f(1) - f(np.float64(factor) ** np.float64(nlevels)))
    out, total = [], 0
    for _ in range(nlevels - 1):
        out.append(int(np.rint(f(nd))))
        total += out[-1]
        nd = f(f(nd) * factor)
    out.append(max(nfeatures - total, 0))
    return out


def umax_table(half=PATCH // 2):
    f = np.float32
    vmax = int(np.floor(f(half) * np.sqrt(f(2)) / f(2) + f(1)))
    vmin = int(np.ceil(f(half) * np.sqrt(f(2)) / f(2)))
    um = [0] * (half + 2)
    for v in range(vmax + 1):
        um[v] = int(np.rint(np.sqrt(float(half * half - v * v))))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while um[v0] == um[v0 + 1]:
            v0 += 1
        um[v] = v0
        v0 += 1
    return um


def harris_responses(img, xs, ys, block=7):
    f = np.float32
    I = img.astype(np.int64)
    r = block // 2
    Ix = (I[1:-1, 2:] - I[1:-1, :-2]) * 2 + (I[:-2, 2:] - I[:-2, :-2]) + (I[2:, 2:] - I[2:, :-2])       # [y-1, x-1] <-> (x, y)
    Iy = (I[2:, 1:-1] - I[:-2, 1:-1]) * 2 + (I[2:, :-2] - I[:-2, :-2]) + (I[2:, 2:] - I[:-2, 2:])
    scale = f(1) / (f(4 * block) * f(255))
    s4 = f(f(f(scale * scale) * scale) * scale)
    out = np.zeros(len(xs), np.float32)
    for i, (x, y) in enumerate(zip(xs, ys)):
        wx = Ix[y - r - 1:y + r, x - r - 1:x + r]
        wy = Iy[y - r - 1:y + r, x - r - 1:x + r]
        a, b, c = f(int((wx * wx).sum())), f(int((wy * wy).sum())), f(int((wx * wy).sum()))
        u = f(a + b)
        out[i] = f(f(f(f(a * b) - f(c * c)) - f(f(HARRIS_K * u) * u)) * s4)
    return out


_P1, _P3, _P5, _P7 = (np.float32(c) * np.float32(180 / np.pi) for c in
                      (0.9997878412794807, -0.3258083974640975, 0.1555786518463281, -0.04432655554792128))


def fast_atan2(y, x):
    """cv::fastAtan2(float y, float x): degrees, 7th-order odd polynomial in float."""
    f = np.float32
    ax, ay, eps = f(abs(x)), f(abs(y)), f(2.220446049250313e-16)
    if ax >= ay:
        c = f(ay / f(ax + eps))
        c2 = f(c * c)
        a = f(f(f(f(f(f(f(_P7 * c2) + _P5) * c2) + _P3) * c2) + _P1) * c)
    else:
        c = f(ax / f(ay + eps))
        c2 = f(c * c)
        a = f(f(90) - f(f(f(f(f(f(f(_P7 * c2) + _P5) * c2) + _P3) * c2) + _P1) * c))
    if x < 0:
        a = f(f(180) - a)
    if y < 0:
        a = f(f(360) - a)
    return a


def ic_angle(img, x, y, um, half=PATCH // 2):
    I = img.astype(np.int64)
    m10 = int((np.arange(-half, half + 1) * I[y, x - half:x + half + 1]).sum())
    m01 = 0
    for v in range(1, half + 1):
        d = um[v]
        plus, minus = I[y + v, x - d:x + d + 1], I[y - v, x - d:x + d + 1]
        m01 += v * int((plus - minus).sum())
        m10 += int((np.arange(-d, d + 1) * (plus + minus)).sum())
    return fast_atan2(np.float32(m01), np.float32(m10))


class CvRng:
    """cv::RNG: multiply-with-carry; uniform(a, b) = a + next() % (b - a)."""

    def __init__(self, state):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def sample_pattern(patch=PATCH, wta_k=2):
    """The BRIEF sample points of a configuration (orb.cpp): bit_pattern_31_ for patchSize 31, else makeRandomPattern
    (RNG 0x34985739); WTA_K 3 / 4: initializeOrbPattern draws 128 tuples of distinct points from that pool (RNG 0x12345678)."""
    if patch == PATCH:
        pool = PATTERN
    else:
        rng = CvRng(0x34985739)
        pool = np.array([(rng.uniform(-(patch // 2), patch // 2 + 1), rng.uniform(-(patch // 2), patch // 2 + 1))
                         for _ in range(512)], np.int32)
    if wta_k == 2:
        return pool
    rng = CvRng(0x12345678)
    pat = np.zeros((128 * wta_k, 2), np.int32)
    for i in range(128):
        for k in range(wta_k):
            while True:
                pt = pool[rng.uniform(0, 512)]
                if not any((pat[wta_k * i + k1] == pt).all() for k1 in range(k)):
                    pat[wta_k * i + k] = pt
                    break
    return pat


def _pad_reflect101(img, pad):
    def idx(n):
        i = np.arange(-pad, n + pad)
        if n == 1:
            return np.zeros_like(i)
        for _ in range(8):
            i = np.where(i < 0, -i, i)
            i = np.where(i >= n, 2 * n - 2 - i, i)
        return i
    return img[np.ix_(idx(img.shape[0]), idx(img.shape[1]))]


def orb_detect_and_compute(image, nfeatures=500, fast_threshold=20, nlevels=N_LEVELS, with_descriptors=True,
                           scale_factor=1.2, edge_threshold=EDGE_THRESHOLD, score_type=0, wta_k=2, patch_size=PATCH):
    """ORB::create(nfeatures, scaleFactor, nlevels, edgeThreshold, 0, WTA_K, scoreType, patchSize, fastThreshold)
    .detectAndCompute.  score_type 0 = HARRIS_SCORE, 1 = FAST_SCORE.  Reads that leave a level (small edgeThreshold)
    see what OpenCV's pyramid buffer holds there: the reflect-101 border of the unblurred level.
    -> (kp float64[n, 6] = x, y, size, angle, response, octave  (float32 values), desc uint8[n, 32] or None)"""
    f = np.float32
    B = 48                                                     # >= ceil(31 * sqrt(2)): the widest excursion of any read
    sf = float(np.float32(scale_factor))                     # create() takes a float, the class keeps a double
    image = np.asarray(image)
    gray = gray_from_bgr(image) if image.ndim == 3 else image
    levels = build_pyramid(gray, nlevels, sf)
    scales = [level_scale(l, sf) for l in range(nlevels)]
    nper = features_per_level(nfeatures, nlevels, sf)
    half = patch_size // 2
    um = umax_table(half)
    padded = [_pad_reflect101(img, B) for img in levels]
    picked = []
    for l, img in enumerate(levels):
        h, w = img.shape
        xs, ys, sc = fast_detect(img, fast_threshold)
        m = (xs >= edge_threshold) & (xs < w - edge_threshold) & (ys >= edge_threshold) & (ys < h - edge_threshold)
        xs, ys, sc = xs[m], ys[m], sc[m]
        resp, ids = retain_best(sc.astype(np.float32), (2 if score_type == 0 else 1) * nper[l])
        picked.append((xs[ids], ys[ids], resp))
    rows = []
    for l, (xs, ys, fast_resp) in enumerate(picked):
        if len(xs) == 0:
            continue
        if score_type == 0:
            resp, ids = retain_best(harris_responses(padded[l], xs + B, ys + B), nper[l])
        else:
            resp, ids = fast_resp, np.arange(len(xs))
        for r, i in zip(resp, ids):
            x, y = int(xs[i]), int(ys[i])
            rows.append((f(f(x) * scales[l]), f(f(y) * scales[l]), f(f(patch_size) * scales[l]),
                         ic_angle(padded[l], x + B, y + B, um, half), r, l))
    kp = np.array(rows, np.float64).reshape(-1, 6)
    if not with_descriptors or len(rows) == 0:
        return kp, (np.zeros((0, 32), np.uint8) if with_descriptors else None)
    blurred = []
    for l, img in enumerate(levels):
        b = padded[l].copy()
        b[B:-B, B:-B] = gaussian_blur_7(img)                   # only the level itself is blurred, its border is not
        blurred.append(b)
    pat = sample_pattern(patch_size, wta_k)
    px, py = pat[:, 0].astype(f), pat[:, 1].astype(f)
    desc = np.zeros((len(rows), 32), np.uint8)
    for r, (x, y, _, ang, _, l) in enumerate(rows):
        inv = f(1) / scales[l]
        cx, cy = int(np.rint(f(x * inv))), int(np.rint(f(y * inv)))
        rad = f(ang * f(np.pi / f(180.0)))
        a, b = f(np.cos(np.float64(rad))), f(np.sin(np.float64(rad)))
        ix = np.rint((px * a).astype(f) - (py * b).astype(f)).astype(np.int64)
        iy = np.rint((px * b).astype(f) + (py * a).astype(f)).astype(np.int64)
        v = blurred[l][cy + iy + B, cx + ix + B].astype(np.int32)
        if wta_k == 2:
            desc[r] = np.packbits((v[0::2] < v[1::2]).astype(np.uint8).reshape(32, 8)[:, ::-1], axis=1).ravel()
        elif wta_k == 3:                                       # index of the maximum of 3 (orb.cpp's tie rules)
            t = v.reshape(32, 4, 3)
            t0, t1, t2 = t[..., 0], t[..., 1], t[..., 2]
            k = np.where(t2 > t1, np.where(t2 > t0, 2, 0), (t1 > t0).astype(np.int64))
            desc[r] = (k[:, 0] | (k[:, 1] << 2) | (k[:, 2] << 4) | (k[:, 3] << 6)).astype(np.uint8)
        else:                                                  # index of the maximum of 4
            t = v.reshape(32, 4, 4)
            t0, t1, t2, t3 = t[..., 0], t[..., 1], t[..., 2], t[..., 3]
            k = np.where(np.maximum(t0, t1) > np.maximum(t2, t3), (t1 > t0).astype(np.int64), np.where(t3 > t2, 3, 2))
            desc[r] = (k[:, 0] | (k[:, 1] << 2) | (k[:, 2] << 4) | (k[:, 3] << 6)).astype(np.uint8)
    return kp, desc
