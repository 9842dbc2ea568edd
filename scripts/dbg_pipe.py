import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle
import sfm_gms_b200 as sg
from sfm_gms_b200 import api
from test_gpu_parity import _ragged_set
oracle.set_num_threads(os.cpu_count())
ctx = sg.Context(0)
rng = np.random.default_rng(78)
sizes_n = [2100, 1900, 0, 1777, 2500, 1, 3000, 2048, 2047]
off, desc, kp, wh = _ragged_set(rng, sizes_n)
pairs = np.array([(0, 1), (1, 3), (3, 4), (4, 6), (6, 7), (7, 8), (8, 0), (2, 1), (1, 2), (5, 4), (0, 1), (7, 6),
                  (6, 4), (3, 3), (4, 8), (8, 7), (0, 8), (1, 0)], np.int32)
exp = []
for (a, b) in pairs:
    d1, d2 = desc[off[a]:off[a+1]], desc[off[b]:off[b+1]]
    if len(d2) == 0 or len(d1) == 0: exp.append(0); continue
    oi, od = oracle.bf_hamming(d1, d2)
    exp.append(oracle.gms(wh[a], wh[b], kp[off[a]:off[a+1]], kp[off[b]:off[b+1]], np.arange(len(oi)), oi)["n_inliers"])
print("oracle   ", exp)
for kern in (api.HAMMING_POPC, api.HAMMING_TC):
    ctx.set_option(api.OPT_HAMMING_KERNEL, kern)
    ctx.set_images(off, desc, kp, wh)
    for rep in range(3):
        a = ctx.match_pairs(pairs)
        print("match_pairs", kern, a["n_inliers"].tolist(), flush=True)
    for rep in range(2):
        b = ctx.match_image_set(off, desc, kp, wh, pairs)
        print("pipelined  ", kern, b["n_inliers"].tolist(), flush=True)
