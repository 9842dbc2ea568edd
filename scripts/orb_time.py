"""Wall-clock of ORB detectAndCompute / dense compute through the C ABI vs cv2 on the same host (informational)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_gms_b200 as sg

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "orb_detect.npz"))
tile = g["view0_bgr_img"]                                   # 320 x 400 x 3 crop of a reference image
img = np.tile(tile, (4, 4, 1))[:1110, :1390].copy()         # the size of the reference's view0.png
ctx = sg.Context(0)
try:
    import cv2
except Exception:
    cv2 = None
for nf, thr in ((500, 20), (10000, 0)):
    ctx.orb_detect_and_compute(img, nf, thr)
    t0 = time.perf_counter()
    for _ in range(5):
        kp, desc = ctx.orb_detect_and_compute(img, nf, thr)
    t_gpu = (time.perf_counter() - t0) / 5
    line = "detectAndCompute %dx%d nfeatures=%d fast=%d: %d kpts, %.2f ms/call (C ABI, host image in, host results out)" % (
        img.shape[1], img.shape[0], nf, thr, len(kp), 1e3 * t_gpu)
    if cv2 is not None:
        orb = cv2.ORB_create(nf); orb.setFastThreshold(thr)
        orb.detectAndCompute(img, None)
        t0 = time.perf_counter()
        for _ in range(3):
            k2, d2 = orb.detectAndCompute(img, None)
        t_cpu = (time.perf_counter() - t0) / 3
        same = len(k2) == len(kp) and np.array_equal(d2, desc) and all(
            (a.pt[0], a.pt[1]) == (float(b[0]), float(b[1])) for a, b in zip(k2, kp))
        line += " | cv2 %.1f ms on %d threads, identical=%s" % (1e3 * t_cpu, cv2.getNumThreads(), same)
    print(line, flush=True)
# dense descriptors: a keypoint at every pixel, angle -1 (DisparityUtil.cpp:127-134)
h, w = img.shape[:2]
gx, gy = np.meshgrid(np.arange(w), np.arange(h), indexing="ij")
pts = np.stack([gx.ravel(), gy.ravel()], 1).astype(np.float32)
ctx.orb_compute(img, pts)
t0 = time.perf_counter()
for _ in range(3):
    kept, desc = ctx.orb_compute(img, pts)
t_gpu = (time.perf_counter() - t0) / 3
line = "dense compute: %d keypoints -> %d descriptors, %.1f ms/call" % (len(pts), len(kept), 1e3 * t_gpu)
if cv2 is not None:
    kps = [cv2.KeyPoint(float(x), float(y), 1) for x, y in pts[:: 16]]
    t0 = time.perf_counter()
    k2, d2 = cv2.ORB_create().compute(img, kps)
    t_cpu = time.perf_counter() - t0
    line += " | cv2 on 1/16 of them: %.1f ms (x16 = %.0f ms)" % (1e3 * t_cpu, 16e3 * t_cpu)
print(line, flush=True)
