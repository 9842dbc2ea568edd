#!/usr/bin/env python
"""Record what the REFERENCE'S OWN MACHINE CODE answers for stage 2 (GMS)  ->  tests/golden/gms_dll.npz.

Run in the build container only (needs /root/reference/SfM-GMS/bin/opencv_xfeatures2d452.dll).  The DLL is hosted
on Linux by oracle/dllref/ (PE loader + Microsoft-ABI calls; see gms_dll_host.c): every value stored here was
computed by the DLL's .text — the exported cv::xfeatures2d::matchGMS (@VA 0x180048280, the very call of
FeatureMatchUtil.cpp:69), GMSMatcher::run per hypothesis (@VA 0x180048630), getGridIndexLeft / getGridIndexRight
(@VA 0x180047bc0 / 0x180047d60) and the ROT / SCALE tables of the mapped image.  Nothing here comes from the oracle:
the oracle only supplies stage-1 matches (BF-Hamming, itself pinned by cv2) as INPUT for the synthetic cases.
Inputs are rebuilt deterministically by tests/gms_dll_cases.py, so only outputs are stored.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gms_dll_cases as C  # noqa: E402
import oracle  # noqa: E402
from oracle import dllref  # noqa: E402


def dll_mask(c, rot, sc, factor):
    """matchGMS through the DLL's export; imgIdx carries the position of each match so that the mask can be
    rebuilt from matchesGMS (the DLL copies whole 16-byte DMatch records, in input order)."""
    n = len(c["q"])
    out = dllref.match_gms(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], rot, sc, factor, tag_positions=True)
    pos = out["imgIdx"]
    assert np.all(np.diff(pos) > 0) and np.array_equal(out["queryIdx"], c["q"][pos]) and np.array_equal(out["trainIdx"], c["t"][pos])
    m = np.zeros(n, bool)
    m[pos] = True
    return m


def record(out, name, c, factor=6.0, hyp=True, flags=C.FLAGS):
    n = len(c["q"])
    out[name + "/n"] = np.int32(n)
    for tag, rot, sc in flags:
        out["%s/mask_%s" % (name, tag)] = np.packbits(dll_mask(c, rot, sc, factor))
    if hyp:
        counts, masks = dllref.hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], factor, True)
        out[name + "/hyp_counts"] = counts
        out[name + "/hyp_sha"] = np.array(C.sha(masks))
        # getInlierMask's selection rule, re-derived from the per-hypothesis runs, must agree with the export
        best = int(np.argmax(counts)) if counts.max() > 0 else -1
        m11 = np.unpackbits(out[name + "/mask_11"])[:n].astype(bool) if "%s/mask_11" % name in out else None
        if m11 is not None:
            assert np.array_equal(m11, masks[best] if best >= 0 else np.zeros(n, bool)), name


def main():
    oracle.build()
    oracle.set_num_threads(os.cpu_count())
    out = {}
    t0 = time.time()
    rot, sc = dllref.tables()
    out["rot"], out["scale"] = rot, sc

    pe = C.grid_edge_points()
    pr = C.grid_random_points()
    rand_parts = []
    for t in (1, 2, 3, 4):
        out["grid_edge_left%d" % t] = dllref.grid_left(pe, t).astype(np.int16)
        rand_parts.append(dllref.grid_left(pr, t))
    for w in C.RIGHT_GRIDS:
        out["grid_edge_right%d" % w] = dllref.grid_right(pe, w, w).astype(np.int16)
        rand_parts.append(dllref.grid_right(pr, w, w))
    out["grid_rand_sha"] = np.array(C.sha(*rand_parts))
    print("grid: %d edge points, %d random points" % (len(pe), len(pr)))

    bf = lambda a, b: oracle.bf_hamming(a, b)[0]
    for name in C.REAL:
        record(out, name, C.real_case(name))
    record(out, "sift_view01_1500", C.sift_case())
    for cfg in ("cfg2_640x480_10k", "cfg3_1080p_50k_rs"):
        c = C.synth_case(cfg, bf)
        out[cfg + "/t_sha"] = np.array(C.sha(c["t"]))
        record(out, cfg, c)
    record(out, "edge_pixels", C.edge_pixels_case())
    record(out, "subset", C.subset_case())
    for k, f in enumerate(C.FACTORS):
        record(out, "view01_2k_f%d" % k, C.real_case("view01_2k"), factor=f, hyp=(k % 3 == 0))
        record(out, "edge_pixels_f%d" % k, C.edge_pixels_case(), factor=f, hyp=False)
    for i in range(C.N_STRESS):
        c = C.stress_case(i)
        record(out, "stress%d" % i, c, factor=c["factor"], hyp=(i % 4 == 0))
    for name, c in C.micro_inputs():
        tag = "%d%d" % (c["rot"], c["sc"])
        record(out, "micro/" + name, c, factor=c["factor"], hyp=False, flags=[(tag, c["rot"], c["sc"])])
    np.savez_compressed(os.path.join(HERE, "gms_dll.npz"), **out)
    print("wrote gms_dll.npz: %d arrays, %.1f s" % (len(out), time.time() - t0))
    for name in C.REAL + ["cfg2_640x480_10k", "cfg3_1080p_50k_rs", "edge_pixels", "subset"]:
        n = int(out[name + "/n"])
        print(name, n, {tag: int(np.unpackbits(out["%s/mask_%s" % (name, tag)])[:n].sum()) for tag, _, _ in C.FLAGS})


if __name__ == "__main__":
    main()
