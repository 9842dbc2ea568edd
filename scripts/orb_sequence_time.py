"""Wall-clock of ORB over an image sequence through sfmgms_set_images_from_pixels (host threads + streams inside)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_gms_b200 as sg

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "orb_detect.npz"))
tile = g["view0_bgr_img"]
rng = np.random.default_rng(0)
imgs = []
for i in range(32):
    im = np.tile(tile, (4, 4, 1))[:1110, :1390].astype(np.int32)
    imgs.append(np.ascontiguousarray((np.roll(im, (3 * i, 5 * i), (0, 1)) + rng.integers(-2, 3, im.shape)).clip(0, 255).astype(np.uint8)))
ctx = sg.Context(0)
for nf, thr in ((500, 20), (10000, 0)):
    ctx.set_images_from_pixels(imgs[:4], nfeatures=nf, fast_threshold=thr)
    t0 = time.perf_counter()
    off = ctx.set_images_from_pixels(imgs, nfeatures=nf, fast_threshold=thr)
    t_seq = time.perf_counter() - t0
    t0 = time.perf_counter()
    for im in imgs:
        ctx.orb_detect_and_compute(im, nf, thr)
    t_one = time.perf_counter() - t0
    print("nfeatures=%d fast=%d: %d images, %d keypoints: sequence call %.1f ms (%.2f ms/image), one call per image %.1f ms (%.2f ms/image)" % (
        nf, thr, len(imgs), off[-1], 1e3 * t_seq, 1e3 * t_seq / len(imgs), 1e3 * t_one, 1e3 * t_one / len(imgs)), flush=True)
