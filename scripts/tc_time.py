"""GPU timing probe for the Hamming stage on a resident batch (no parity check)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_gms_b200 as sg
from sfm_gms_b200 import api
import bench
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
_s = bench.make_batch("cfg2", P)
desc, kp = torch.from_numpy(_s["desc"]).to(dev), torch.from_numpy(_s["kp"]).to(dev)
ctx = sg.Context(0)
ctx.set_option(api.OPT_TIMING, 1)
KERN = os.environ.get("SFMGMS_KERNEL", "tc")
ctx.set_option(api.OPT_HAMMING_KERNEL, {"tc": api.HAMMING_TC, "fp4": api.HAMMING_FP4, "popc": api.HAMMING_POPC}[KERN])
off = np.arange(2 * P + 1, dtype=np.int64) * 10_000
sizes = np.tile(np.array([[640, 480]], np.int32), (2 * P, 1))
pairs = np.ascontiguousarray(np.arange(2 * P, dtype=np.int32).reshape(-1, 2))
ctx.set_images_raw(off, desc.data_ptr(), kp.data_ptr(), sizes, api.SFMGMS_DEVICE, keepalive=(desc, kp))
o = torch.zeros(P, dtype=torch.int32, device=dev)
for cache in ((1, 0) if not os.environ.get("SFMGMS_TC_DEBUG") else (1,)):
    ctx.set_option(api.OPT_TC_OPERAND_CACHE, cache)
    ts = []
    for i in range(6):
        ctx.match_pairs_raw(pairs, 0, 0, 6.0, api.SFMGMS_DEVICE, o.data_ptr())
        ts.append(ctx.last_timing())
    print(KERN, "dbg=%s cache=%d P=%d hamming_ms=%.3f (%.2f us/pair) gms_ms=%.3f launches=%d" % (os.environ.get("SFMGMS_TC_DEBUG", "0"), cache, P, np.mean([t[0] for t in ts[2:]]), 1e3 * np.mean([t[0] for t in ts[2:]]) / P, np.mean([t[1] for t in ts[2:]]), ts[-1][2]), flush=True)

# ---- sustained run with clock/power sampling (pynvml) ----
import threading, time
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples = []
    stop = False
    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
            time.sleep(0.02)
    th = threading.Thread(target=sampler); th.start()
    ctx.set_option(api.OPT_TC_OPERAND_CACHE, 1)
    t0 = time.time(); ts = []
    while time.time() - t0 < 3.0:
        ctx.match_pairs_raw(pairs, 0, 0, 6.0, api.SFMGMS_DEVICE, o.data_ptr())
        ts.append(ctx.last_timing()[0])
    stop = True; th.join()
    sm = np.array([s[0] for s in samples]); pw = np.array([s[1] for s in samples])
    print(KERN, "dbg=%s lib=%s sustained 3s:" % (os.environ.get("SFMGMS_TC_DEBUG", "0"), os.path.basename(os.environ.get("SFMGMS_LIB", "default"))), " hamming_ms first=%.3f median=%.3f last=%.3f | sm_mhz min/med/max %d/%d/%d | power W med/max %.0f/%.0f | reasons %s" % (
        ts[0], np.median(ts), ts[-1], sm.min(), np.median(sm), sm.max(), np.median(pw), pw.max(), sorted(set(hex(s[2]) for s in samples))), flush=True)
except Exception as ex:
    print("nvml sampling failed:", ex)
