#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run in the BUILD container only).

Sources of truth, per field:
  desc*/kp*/size*      cv2 4.13 ORB (nfeatures, fastThreshold=0) on the reference's own images
                       /root/reference/SfM-GMS/SourceImages/* (BASELINE.json configs[0]).
  bf_train/bf_dist     cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).match  -- the EXECUTABLE
                       reference for stage 1 (FeatureMatchUtil.cpp:66-68 call shape).
  xc_*                 cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match (FeatureMatchUtil.cpp:22).
  anchor_*             SURVEY.md Appendix C per-hypothesis inlier counts (survey's independent numpy
                       restatement of the DLL disassembly) -- typed in from SURVEY.md, NOT computed here.
  gms_mask_*           the C oracle's masks (regression pin of the oracle itself).  The EXECUTABLE reference for
                       stage 2 is the vendored DLL's own code: see make_gms_dll_golden.py / gms_dll.npz.
/root/reference is not available on the GPU box; tests only read the .npz files written here.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

R = "/root/reference/SfM-GMS/SourceImages/"

ANCHORS = {
    "pikabun12": dict(default=3704, best=4090, best_hyp=8, counts=[
        3704, 2995, 2379, 2196, 2258, 2255, 2506, 3091, 4090, 3704, 3181, 3140, 3213, 3144, 3401, 3882,
        4056, 3451, 3053, 2868, 2847, 2844, 3159, 3729, 3567, 2666, 1991, 1756, 1786, 1718, 2070, 2864,
        2697, 1727, 928, 921, 928, 967, 983, 1922]),
    "disparityLR": dict(default=6504, best=6740, best_hyp=8, counts=[
        6504, 5926, 5002, 4922, 4928, 4921, 4980, 6145, 6740, 6605, 6105, 6008, 5935, 6143, 6145, 6585,
        6525, 6337, 5474, 5336, 5357, 5389, 5698, 6304, 5907, 5038, 3471, 3275, 3275, 3210, 3569, 4951,
        4753, 3523, 1878, 1806, 1763, 1768, 1805, 3417]),
}


def orb_pair(a, b, nfeat, rot180=False):
    i1, i2 = cv2.imread(R + a), cv2.imread(R + b)
    if rot180:  # main.cpp:36 img_rotate(img2, 180)
        i2 = cv2.rotate(i2, cv2.ROTATE_180)
    orb = cv2.ORB_create(nfeat)
    orb.setFastThreshold(0)
    k1, d1 = orb.detectAndCompute(i1, None)
    k2, d2 = orb.detectAndCompute(i2, None)
    p1 = np.array([k.pt for k in k1], np.float32)
    p2 = np.array([k.pt for k in k2], np.float32)
    return (i1.shape[1], i1.shape[0]), (i2.shape[1], i2.shape[0]), p1, p2, d1, d2


def pack(name, s1, s2, p1, p2, d1, d2, anchors=None):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).match(d1, d2)
    bt = np.array([x.trainIdx for x in m], np.int32)
    bd = np.array([int(x.distance) for x in m], np.int32)
    assert [x.queryIdx for x in m] == list(range(len(m)))
    xm = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(d1, d2)
    xq = np.array([x.queryIdx for x in xm], np.int32)
    xt = np.array([x.trainIdx for x in xm], np.int32)
    out = dict(size1=np.array(s1, np.int32), size2=np.array(s2, np.int32), kp1=p1, kp2=p2, desc1=d1, desc2=d2,
               bf_train=bt, bf_dist=bd, xc_query=xq, xc_train=xt)
    q = np.arange(len(bt), dtype=np.int32)
    for tag, rot, sc in [("00", 0, 0), ("10", 1, 0), ("01", 0, 1), ("11", 1, 1)]:
        r = oracle.gms(s1, s2, p1, p2, q, bt, rot, sc)
        out["gms_mask_" + tag] = np.packbits(r["mask"])
        out["gms_len_" + tag] = np.int32(len(r["mask"]))
        out["gms_n_" + tag] = np.int32(r["n_inliers"])
        out["gms_hyp_" + tag] = r["hyp_counts"]
        out["gms_best_" + tag] = np.int32(r["best_hyp"])
    if anchors:
        out["anchor_counts"] = np.array(anchors["counts"], np.int32)
        out["anchor_default"] = np.int32(anchors["default"])
        out["anchor_best"] = np.int32(anchors["best"])
        out["anchor_best_hyp"] = np.int32(anchors["best_hyp"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, len(p1), len(p2), "default", int(out["gms_n_00"]), "rs", int(out["gms_n_11"]),
          "best_hyp", int(out["gms_best_11"]), "xc", len(xq))


def tie_case():
    """cv2-pinned tie-break fixture: duplicates everywhere."""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    t = np.concatenate([base, base, base])  # every train row triplicated
    q = base[rng.integers(0, 50, 64)].copy()
    q[::3, 5] ^= 0x10
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).match(q, t)
    # low-entropy descriptors: many equal distances
    q2 = rng.integers(0, 2, (257, 32), dtype=np.uint8) * 255
    t2 = rng.integers(0, 2, (1031, 32), dtype=np.uint8) * 255
    m2 = cv2.BFMatcher(cv2.NORM_HAMMING, False).match(q2, t2)
    np.savez_compressed(os.path.join(HERE, "bf_ties.npz"), q=q, t=t,
                        train=np.array([x.trainIdx for x in m], np.int32),
                        dist=np.array([int(x.distance) for x in m], np.int32), q2=q2, t2=t2,
                        train2=np.array([x.trainIdx for x in m2], np.int32),
                        dist2=np.array([int(x.distance) for x in m2], np.int32))
    print("bf_ties", max(x.trainIdx for x in m))


def sift_case():
    """The reference's literal main path (FeatureMatchUtil.cpp:10, 66-69): SIFT + BFMatcher(NORM_L2) + matchGMS(true,true).
    cv2 pins stage 1 (trainIdx + float distance); descriptors are integer-valued, stored as uint8."""
    i1, i2 = cv2.imread(R + "view0.png"), cv2.imread(R + "view1.png")
    sift = cv2.SIFT_create(1500)
    k1, d1 = sift.detectAndCompute(i1, None)
    k2, d2 = sift.detectAndCompute(i2, None)
    assert np.abs(d1 - np.round(d1)).max() == 0 and d1.max() <= 255 and d1.min() >= 0
    m = cv2.BFMatcher(cv2.NORM_L2, False).match(d1, d2)
    bt = np.array([x.trainIdx for x in m], np.int32)
    bd = np.array([x.distance for x in m], np.float32)
    p1 = np.array([k.pt for k in k1], np.float32)
    p2 = np.array([k.pt for k in k2], np.float32)
    s1, s2 = (i1.shape[1], i1.shape[0]), (i2.shape[1], i2.shape[0])
    r = oracle.gms(s1, s2, p1, p2, np.arange(len(bt), dtype=np.int32), bt, True, True)
    # float-tie case: d2 = n+1 at train row 0 and d2 = n at row 1 with sqrtf(n) == sqrtf(n+1) (n >= 2^22):
    # OpenCV compares float distances with strict '<' and must answer row 0.
    n = next(v for v in range(1 << 22, 1 << 23) if np.sqrt(np.float32(v)) == np.sqrt(np.float32(v + 1)))

    def vec_with_norm2(target):
        v = np.zeros(128, np.float32)
        rem = target
        for k in range(128):
            x = min(255, int(np.floor(np.sqrt(rem))))
            v[k] = x
            rem -= x * x
        assert rem == 0, rem
        return v

    tq = np.zeros((1, 128), np.float32)
    tt = np.stack([vec_with_norm2(n + 1), vec_with_norm2(n), vec_with_norm2(n + 1)])
    tm = cv2.BFMatcher(cv2.NORM_L2, False).match(tq, tt)
    assert tm[0].trainIdx == 0, tm[0].trainIdx
    np.savez_compressed(os.path.join(HERE, "sift_view01_1500.npz"), size1=np.array(s1, np.int32), size2=np.array(s2, np.int32),
                        kp1=p1, kp2=p2, desc1=d1.astype(np.uint8), desc2=d2.astype(np.uint8), l2_train=bt, l2_dist=bd,
                        gms_mask_11=np.packbits(r["mask"]), gms_n_11=np.int32(r["n_inliers"]), gms_best_11=np.int32(r["best_hyp"]),
                        gms_len_11=np.int32(len(r["mask"])), tie_q=tq, tie_t=tt, tie_train=np.int32(tm[0].trainIdx),
                        tie_dist=np.float32(tm[0].distance))
    print("sift_view01_1500", len(bt), "gms(rot,scale)", r["n_inliers"], "tie n =", n)


def bruteforce_case():
    """bruteForceMatch (FeatureMatchUtil.cpp:20-31): BFMatcher(NORM_L2, crossCheck=true).match, std::sort by
    distance, prune while front*4 < back, cap 500.  cv2 pins the cross-checked match list; the sort/prune/cap tail
    is restated here with a STABLE sort (std::sort leaves the order of equal distances unspecified; ties keep
    query order).  Inputs are the descriptors already stored in sift_view01_1500.npz / view01_2k.npz."""
    def tail(m, coef=4.0, cap=500):
        m = sorted(m, key=lambda x: x[2])                     # stable: ties stay in queryIdx order
        while m and np.float64(m[0][2]) * coef < np.float64(m[-1][2]):
            m.pop()
        return m[:cap]

    out = {}
    g = np.load(os.path.join(HERE, "sift_view01_1500.npz"))
    d1, d2 = g["desc1"].astype(np.float32), g["desc2"].astype(np.float32)
    h = np.load(os.path.join(HERE, "view01_2k.npz"))
    for name, a, b, norm in (("l2", d1, d2, cv2.NORM_L2), ("l2_rev", d2[:700], d1, cv2.NORM_L2),
                             ("ham", h["desc1"], h["desc2"], cv2.NORM_HAMMING)):
        for xc in (True, False):
            m = [(x.queryIdx, x.trainIdx, np.float32(x.distance)) for x in cv2.BFMatcher(norm, xc).match(a, b)]
            key = "%s_%s" % (name, "xc" if xc else "nn")
            out[key + "_full"] = np.array([(q, t) for q, t, _ in m], np.int32).reshape(-1, 2)
            out[key + "_full_dist"] = np.array([d for _, _, d in m], np.float32)
            for coef, cap in ((4.0, 500), (1.5, 100000), (4.0, 37)):
                r = tail(m, coef, cap)
                k2 = "%s_c%g_m%d" % (key, coef, cap)
                out[k2] = np.array([(q, t) for q, t, _ in r], np.int32).reshape(-1, 2)
                out[k2 + "_dist"] = np.array([d for _, _, d in r], np.float32)
        print("bruteforce", name, len(out[name + "_xc_full"]), len(out[name + "_xc_c4_m500"]), len(out[name + "_xc_c1.5_m100000"]))
    np.savez_compressed(os.path.join(HERE, "bruteforce.npz"), **out)


def l2_float_case():
    """General float descriptors (not integer-valued): cv2.BFMatcher(NORM_L2) pins trainIdx and the float distance,
    which depend on the addition order inside OpenCV's normL2Sqr_.  RootSIFT of the SIFT fixture (dim 128), a
    64-wide slice, and a width with a scalar tail (dim 70); cross-check list for RootSIFT."""
    g = np.load(os.path.join(HERE, "sift_view01_1500.npz"))
    d1, d2 = g["desc1"].astype(np.float32), g["desc2"].astype(np.float32)
    r1 = np.sqrt(d1 / np.maximum(d1.sum(1, keepdims=True), 1e-7)).astype(np.float32)
    r2 = np.sqrt(d2 / np.maximum(d2.sum(1, keepdims=True), 1e-7)).astype(np.float32)
    out = {}
    for name, a, b in (("root128", r1, r2), ("root64", np.ascontiguousarray(r1[:, :64]), np.ascontiguousarray(r2[:, :64])),
                       ("root70", np.ascontiguousarray(r1[:600, 3:73]), np.ascontiguousarray(r2[:900, 3:73]))):
        m = cv2.BFMatcher(cv2.NORM_L2, False).match(a, b)
        out[name + "_train"] = np.array([x.trainIdx for x in m], np.int32)
        out[name + "_dist"] = np.array([x.distance for x in m], np.float32)
    m = cv2.BFMatcher(cv2.NORM_L2, True).match(r1, r2)
    out["root128_xc"] = np.array([(x.queryIdx, x.trainIdx) for x in m], np.int32)
    out["root128_xc_dist"] = np.array([x.distance for x in m], np.float32)
    np.savez_compressed(os.path.join(HERE, "l2_float.npz"), **out)
    print("l2_float", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    oracle.build()
    oracle.set_num_threads(os.cpu_count())
    pack("pikabun12", *orb_pair("PikaBun1.jpg", "PikaBun2.jpg", 10000), anchors=ANCHORS["pikabun12"])
    pack("disparityLR", *orb_pair("Disparity_L.jpg", "Disparity_R.jpg", 10000), anchors=ANCHORS["disparityLR"])
    pack("view01_2k", *orb_pair("view0.png", "view1.png", 2000))
    pack("bun12_rot180_3k", *orb_pair("Bun1.jpg", "Bun2.jpg", 3000, rot180=True))
    tie_case()
    sift_case()
    bruteforce_case()
    l2_float_case()
