"""Randomised ORB parity against live cv2 (needs cv2 on the box): sizes, channels, nfeatures, FAST thresholds, textures."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
import sfm_gms_b200 as sg

ctx = sg.Context(0)
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
t0 = time.time()
it = bad = 0
while time.time() - t0 < budget:
    h, w = int(rng.integers(70, 700)), int(rng.integers(70, 900))
    ch = int(rng.choice([1, 3]))
    cell = int(rng.choice([2, 3, 5, 8, 13]))
    shape = (h // cell + 2, w // cell + 2) if ch == 1 else (h // cell + 2, w // cell + 2, 3)
    base = rng.integers(0, 256, shape, dtype=np.uint8)
    img = np.kron(base, np.ones((cell, cell) if ch == 1 else (cell, cell, 1), np.uint8))[:h, :w]
    noise = int(rng.choice([0, 2, 10, 40]))
    if noise:
        img = (img.astype(np.int32) + rng.integers(-noise, noise + 1, img.shape)).clip(0, 255)
    if rng.random() < 0.3:
        img = cv2.GaussianBlur(img.astype(np.uint8), (5, 5), 1.2)
    img = np.ascontiguousarray(img.astype(np.uint8))
    nf, thr = int(rng.choice([1, 50, 500, 3000, 20000])), int(rng.choice([0, 5, 20, 60]))
    full = os.environ.get("ORB_STRESS_FULL") is not None       # also randomise the other ORB::create arguments
    sf = float(rng.choice([1.2, 1.1, 1.5, 2.0])) if full else 1.2
    nl = int(rng.choice([8, 1, 3, 12])) if full else 8
    edge = int(rng.choice([31, 0, 3, 7, 19, 40])) if full else 31
    wta = int(rng.choice([2, 3, 4])) if full else 2
    score = int(rng.choice([0, 1])) if full else 0
    patch = int(rng.choice([31, 9, 21, 45, 63])) if full else 31
    orb = cv2.ORB_create(nf, sf, nl, edge, 0, wta, score, patch, thr)
    try:
        rk, rd = orb.detectAndCompute(img, None)
    except cv2.error:                                          # a pyramid level of size 0: OpenCV asserts, we must refuse too
        try:
            ctx.orb_detect_and_compute(img, nf, thr, True, sf, nl, edge, score, 0, wta, patch)
            bad += 1
            print("MISMATCH: cv2 raised, the library did not", (sf, nl, w, h), flush=True)
        except sg.SfmGmsError:
            pass
        continue
    ref = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in rk], np.float32).reshape(-1, 6)
    kp, desc = ctx.orb_detect_and_compute(img, nf, thr, True, sf, nl, edge, score, 0, wta, patch)
    ok = kp.shape == ref.shape and np.array_equal(kp, ref) and (len(rk) == 0 or (rd is not None and np.array_equal(desc, rd)))
    # provided keypoints: random positions / angles / octaves
    n = 400
    pts = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1).astype(np.float32)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    octv = rng.integers(0, 3, n).astype(np.int32) if min(h, w) > 200 else np.zeros(n, np.int32)
    kps = [cv2.KeyPoint(float(p[0]), float(p[1]), 31, float(a), 0, int(o), i) for i, (p, a, o) in enumerate(zip(pts, ang, octv))]
    ck, cd = cv2.ORB_create().compute(img, kps)
    kept, d2 = ctx.orb_compute(img, pts, ang, octv)
    ok2 = [k.class_id for k in ck] == kept.tolist() and (len(ck) == 0 or np.array_equal(cd, d2))
    it += 1
    if not (ok and ok2):
        bad += 1
        nbits = int(np.unpackbits(desc ^ rd).sum()) if ok is False and desc is not None and rd is not None and desc.shape == rd.shape else -1
        nb2 = int(np.unpackbits(cd ^ d2).sum()) if (ck and cd.shape == d2.shape) else -1
        print("MISMATCH it=%d params=%r %dx%dx%d cell=%d noise=%d nf=%d thr=%d detect_ok=%s (n %d vs %d, bits %d) compute_ok=%s (bits %d)" % (
            it, (sf, nl, edge, wta, score, patch), w, h, ch, cell, noise, nf, thr, ok, len(kp), len(ref), nbits, ok2, nb2), flush=True)
print("orb stress: %d iterations, %d mismatching in %.0f s" % (it, bad, time.time() - t0))
