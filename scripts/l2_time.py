"""Device time of the L2 brute-force kernels on one 10k x 10k, 128-d pair (integer SIFT-like data and general floats)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_gms_b200 as sg
from sfm_gms_b200 import api
ctx = sg.Context(0)
ctx.set_option(api.OPT_TIMING, 1)
rng = np.random.default_rng(0)
q = rng.integers(0, 160, (10000, 128)).astype(np.float32)
t = rng.integers(0, 160, (10000, 128)).astype(np.float32)
for name, mode, a, b in (("tcgen05 u8 (integer-valued)", 2, q, t), ("DP4A (integer-valued)", 1, q, t),
                         ("fp32 order-exact (forced, integer data)", 3, q, t),
                         ("fp32 order-exact (RootSIFT-like floats; incl. the refused tensor-core attempt)", 0, np.sqrt(q / q.sum(1, keepdims=True)), np.sqrt(t / t.sum(1, keepdims=True)))):
    ctx.set_option(api.OPT_L2_KERNEL, mode)
    ms = []
    for _ in range(5):
        ctx.bf_l2(a, b)
        ms.append(ctx.last_timing()[0])
    print("%-85s %.3f ms" % (name, float(np.median(ms[1:]))), flush=True)
    if mode != 3 and a is q:
        xc = []
        for _ in range(3):
            ctx.bf_l2_crosscheck(a, b)
            xc.append(ctx.last_timing()[0])
        print("%-85s %.3f ms" % ("  cross-check (both directions)", float(np.median(xc[1:]))), flush=True)
