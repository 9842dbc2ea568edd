"""BASELINE config 5: all-pairs matching over a synthetic image sequence (device-resident set, device outputs).

usage: allpairs.py [n_images=512] [n_kp=10000]   — prints pairs/s and a few sanity statistics; with --check N it
verifies N random pairs bit-exactly against the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_gms_b200 as sg
from sfm_gms_b200 import api, synth

n_images = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 512
n_kp = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 10000
n_check = int(sys.argv[sys.argv.index("--check") + 1]) if "--check" in sys.argv else 0
t0 = time.time()
s = synth.make_sequence(n_images, n_kp)
pairs = synth.all_pairs(n_images)
print("generated %d images x %d kpts, %d pairs in %.1f s" % (n_images, n_kp, len(pairs), time.time() - t0), flush=True)
dev = torch.device("cuda", 0)
desc = torch.from_numpy(s["desc"]).to(dev)
kp = torch.from_numpy(s["kp"]).to(dev)
ctx = sg.Context(0)
ctx.set_images_raw(s["offsets"], desc.data_ptr(), kp.data_ptr(), s["sizes"], api.SFMGMS_DEVICE, keepalive=(desc, kp))
P = len(pairs)
ninl = torch.zeros(P, dtype=torch.int32, device=dev)
want_full = n_check > 0
tot = P * n_kp
ti = torch.zeros(tot if want_full else 1, dtype=torch.int32, device=dev)
mk = torch.zeros(tot if want_full else 1, dtype=torch.uint8, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.match_pairs_raw(pairs, 0, 0, 6.0, api.SFMGMS_DEVICE, ninl.data_ptr(), 0, 0, ti.data_ptr() if want_full else 0, 0,
                        mk.data_ptr() if want_full else 0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("run %d: %d pairs in %.3f s = %.0f pairs/s (%.2f us/pair)" % (rep, P, dt, P / dt, 1e6 * dt / P), flush=True)
h = ninl.cpu().numpy()
gap = pairs[:, 1] - pairs[:, 0]
for g in (1, 8, 64, n_images - 1):
    sel = h[gap == g]
    if len(sel):
        print("  image gap %3d: mean inliers %.0f" % (g, sel.mean()))
if n_check:
    import oracle
    oracle.set_num_threads(os.cpu_count())
    rng = np.random.default_rng(0)
    off = s["offsets"]
    for p in rng.choice(P, n_check, replace=False):
        a, b = pairs[p]
        oi, od = oracle.bf_hamming(s["desc"][off[a]:off[a + 1]], s["desc"][off[b]:off[b + 1]])
        o = oracle.gms(s["sizes"][a], s["sizes"][b], s["kp"][off[a]:off[a + 1]], s["kp"][off[b]:off[b + 1]], np.arange(n_kp), oi)
        gi = ti[p * n_kp:(p + 1) * n_kp].cpu().numpy()
        gm = mk[p * n_kp:(p + 1) * n_kp].cpu().numpy().astype(bool)
        assert np.array_equal(gi, oi) and np.array_equal(gm, o["mask"]) and h[p] == o["n_inliers"], p
    print("  %d random pairs verified bit-exactly against the oracle" % n_check)
