"""GPU probe: tensor-core Hamming kernel vs popc kernel vs oracle on a few shapes, with timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import sfm_gms_b200 as sg
from sfm_gms_b200 import api

oracle.set_num_threads(os.cpu_count())
ctx = sg.Context(0)
rng = np.random.default_rng(0)
shapes = [(128, 128), (256, 128), (300, 1000), (1000, 3000), (10000, 10000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for nq, nt in shapes:
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q[0] = t[nt - 1]; t[nt // 2] = t[0]
    oi, od = oracle.bf_hamming(q, t)
    for name, k in [("popc", api.HAMMING_POPC), ("tc", api.HAMMING_TC), ("fp4", api.HAMMING_FP4)]:
        ctx.set_option(api.OPT_HAMMING_KERNEL, k)
        ctx.set_option(api.OPT_TIMING, 1)
        for rep in range(3):
            idx, dist = ctx.bf_hamming(q, t)
        ms = ctx.last_timing()[0]
        ok = np.array_equal(idx, oi) and np.array_equal(dist, od)
        bad = np.nonzero((idx != oi) | (dist != od))[0]
        print(f"{nq}x{nt} {name}: ok={ok} hamming_ms={ms:.4f} nbad={len(bad)}", flush=True)
        if not ok:
            print("  first bad rows", bad[:8], "got", idx[bad[:8]], dist[bad[:8]], "exp", oi[bad[:8]], od[bad[:8]], flush=True)
