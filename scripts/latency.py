"""Single-pair call latency through the C ABI (host buffers in, host results out) for BASELINE configs 2-4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_gms_b200 as sg
from sfm_gms_b200 import api, synth
ctx = sg.Context(0)
ctx.set_option(api.OPT_TIMING, 1)
for name in ["cfg2_640x480_10k", "cfg3_1080p_50k_rs", "cfg4_4k_200k"]:
    d = synth.make_config(name)
    for flags in ([(d["with_rotation"], d["with_scale"])] + ([(True, True)] if name.startswith("cfg4") else [])):
        ts, dev = [], []
        for i in range(7):
            t0 = time.perf_counter()
            r = ctx.match_pair(d["desc1"], d["desc2"], d["kp1"], d["kp2"], d["size1"], d["size2"], flags[0], flags[1])
            ts.append(time.perf_counter() - t0)
            dev.append(ctx.last_timing())
        print("%s rot=%d scale=%d: call %.3f ms (min %.3f) | device: hamming %.3f ms, gms %.3f ms | inliers %d best_hyp %d" % (
            name, flags[0], flags[1], 1e3 * np.median(ts[2:]), 1e3 * min(ts), np.median([x[0] for x in dev[2:]]),
            np.median([x[1] for x in dev[2:]]), r["n_inliers"], r["best_hyp"]), flush=True)
