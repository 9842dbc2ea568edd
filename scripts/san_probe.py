"""Small end-to-end run for compute-sanitizer: all three Hamming kernels + both GMS paths on small inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_gms_b200 as sg
from sfm_gms_b200 import api, synth
ctx = sg.Context(0)
d = synth.make_pair(640, 480, 1500, seed=5)
s = synth.make_pair_batch(3, n=700)
for k in (api.HAMMING_POPC, api.HAMMING_TC, api.HAMMING_FP4):
    ctx.set_option(api.OPT_HAMMING_KERNEL, k)
    r = ctx.match_pair(d["desc1"], d["desc2"], d["kp1"], d["kp2"], d["size1"], d["size2"], True, True)
    ctx.set_images(s["offsets"], s["desc"], s["kp"], s["sizes"])
    a = ctx.match_pairs(s["pairs"])
    b = ctx.match_image_set(s["offsets"], s["desc"], s["kp"], s["sizes"], s["pairs"])
    assert np.array_equal(a["mask"], b["mask"])
    p1, p2, n = ctx.inlier_points(0, 700)
    idx, dist, keep = ctx.bf_hamming_crosscheck(d["desc1"][:300], d["desc2"][:500])
    print("kernel", k, "inliers", r["n_inliers"], a["n_inliers"].tolist(), n, int(keep.sum()), flush=True)
ctx.close()
print("done")
