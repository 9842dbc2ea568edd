"""Randomised stress: many ragged shapes, all Hamming kernels must agree bit-for-bit with the POPC kernel (itself
oracle-checked in tests); multi-pair batches through both host paths, random chunk sizes of the pair-list runner, the
compacted outputs against the per-match outputs, every 7th batch against the CPU oracle; repeated to shake out rare
synchronisation bugs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_gms_b200 as sg
from sfm_gms_b200 import api, synth
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
ctx = sg.Context(0)
rng = np.random.default_rng(int(time.time()))
t_end = time.time() + budget
it = 0
while time.time() < t_end:
    it += 1
    mode = it % 3
    if mode == 0:
        nq, nt = int(rng.integers(1, 4000)), int(rng.integers(1, 4000))
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 4 if it % 2 else 256, (nt, 32), dtype=np.uint8)
        ref = None
        for k in (api.HAMMING_POPC, api.HAMMING_TC, api.HAMMING_FP4):
            ctx.set_option(api.OPT_HAMMING_KERNEL, k)
            out = ctx.bf_hamming(q, t)
            if ref is None:
                ref = out
            else:
                assert np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]), (it, k, nq, nt)
    else:
        n_img = int(rng.integers(2, 9))
        sizes_n = [int(x) for x in rng.integers(0, 2500, n_img)]
        w = 640
        descs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in sizes_n]
        kps = [synth.random_keypoints(rng, n, w, 480) for n in sizes_n]
        off = np.concatenate([[0], np.cumsum(sizes_n)]).astype(np.int64)
        desc = np.concatenate(descs) if sum(sizes_n) else np.zeros((0, 32), np.uint8)
        kp = np.concatenate(kps) if sum(sizes_n) else np.zeros((0, 2), np.float32)
        wh = np.tile(np.array([[w, 480]], np.int32), (n_img, 1))
        n_pairs = int(rng.integers(1, 40))
        pairs = rng.integers(0, n_img, (n_pairs, 2)).astype(np.int32)
        rot, sc = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        ctx.set_option(api.OPT_CHUNK_ROWS, int(rng.choice([1, 2000, 7000, 1 << 22])))
        ref = None
        for k in (api.HAMMING_POPC, api.HAMMING_FP4, api.HAMMING_TC):
            ctx.set_option(api.OPT_HAMMING_KERNEL, k)
            if mode == 1:
                ctx.set_images(off, desc, kp, wh)
                out = ctx.match_pairs(pairs, rot, sc)
            else:
                out = ctx.match_image_set(off, desc, kp, wh, pairs, rot, sc)
            if ref is None:
                ref = out
            else:
                for key in ("train_idx", "dist", "mask", "n_inliers", "best_hyp", "mask_len"):
                    assert np.array_equal(out[key], ref[key]), (it, k, key)
        # compacted outputs == the inliers of the per-match outputs
        ctx.set_images(off, desc, kp, wh)
        c = ctx.match_pairs_compact(pairs, rot, sc)
        mo = ref["offsets"]
        for p in range(n_pairs):
            m = ref["mask"][mo[p]:mo[p + 1]].astype(bool)
            rows = c["matches"][c["offsets"][p]:c["offsets"][p + 1]]
            assert np.array_equal(rows["queryIdx"], np.nonzero(m)[0]) and np.array_equal(rows["trainIdx"], ref["train_idx"][mo[p]:mo[p + 1]][m]), (it, p)
            assert np.array_equal(c["pts1"][c["offsets"][p]:c["offsets"][p + 1]], kp[off[pairs[p, 0]]:off[pairs[p, 0] + 1]][m]), (it, p)
        if it % 7 == 1:
            import oracle
            for p in range(min(n_pairs, 6)):
                a, b = pairs[p]
                if sizes_n[a] == 0 or sizes_n[b] == 0:
                    continue
                oi, od = oracle.bf_hamming(descs[a], descs[b])
                o = oracle.gms((w, 480), (w, 480), kps[a], kps[b], np.arange(len(oi), dtype=np.int32), oi, rot, sc)
                sl = slice(mo[p], mo[p + 1])
                assert np.array_equal(ref["train_idx"][sl], oi) and np.array_equal(ref["dist"][sl], od), (it, p, "bf vs oracle")
                full = o["mask"] if len(o["mask"]) == len(oi) else np.zeros(len(oi), bool)
                assert np.array_equal(ref["mask"][sl].astype(bool), full) and ref["n_inliers"][p] == o["n_inliers"], (it, p, "gms vs oracle")
print("stress ok: %d iterations in %.0f s" % (it, budget))
