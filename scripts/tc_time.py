"""GPU timing probe for the Hamming stage on a resident batch (no parity check)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_gms_b200 as sg
from sfm_gms_b200 import api
import bench
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
desc, kp = bench.gen_pairs_torch(P, 2, dev)
ctx = sg.Context(0)
ctx.set_option(api.OPT_TIMING, 1)
off = np.arange(2 * P + 1, dtype=np.int64) * bench.N_KP
sizes = np.tile(np.array([[640, 480]], np.int32), (2 * P, 1))
pairs = np.ascontiguousarray(np.arange(2 * P, dtype=np.int32).reshape(-1, 2))
ctx.set_images_raw(off, desc.data_ptr(), kp.data_ptr(), sizes, api.SFMGMS_DEVICE, keepalive=(desc, kp))
o = torch.zeros(P, dtype=torch.int32, device=dev)
for cache in (1, 0):
    ctx.set_option(api.OPT_TC_OPERAND_CACHE, cache)
    ts = []
    for i in range(6):
        ctx.match_pairs_raw(pairs, 0, 0, 6.0, api.SFMGMS_DEVICE, o.data_ptr())
        ts.append(ctx.last_timing())
    print("dbg=%s cache=%d P=%d hamming_ms=%.3f (%.2f us/pair) gms_ms=%.3f launches=%d" % (os.environ.get("SFMGMS_TC_DEBUG", "0"), cache, P, np.mean([t[0] for t in ts[2:]]), 1e3 * np.mean([t[0] for t in ts[2:]]) / P, np.mean([t[1] for t in ts[2:]]), ts[-1][2]), flush=True)
